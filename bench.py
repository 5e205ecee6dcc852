#!/usr/bin/env python
"""bench.py — stress-step throughput (Mcell-steps/s) of the B200 path, with roofline and CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
  * a "step" = one constitutiveEq::correct() (update + assembly + solve of all valid components, all
    modes + eig/exp/tau + tau BCs) with U/phi already resident in HBM;
  * N = 1 workload = BASELINE.json configs[1]: C2, 2-D 4:1 planar contraction, PTTLog, 971,271 cells;
    N > 1 (torchrun, one rank per GPU) = the same blockMeshDict refined so that every GPU keeps
    ~971k cells (weak scaling), decomposed in x like decomposePar `simple (N 1 1)`;
  * `e2e` = the same metric through rheo_gpu_correct() with pinned HOST buffers (U, U_b, phi up; tau
    down every step);
  * `--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over
    sub-domains = one per host core) on the same workload — the reference binary itself cannot be
    built without OpenFOAM-9/Eigen/MPI.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from rheotool_b200 import abi, cases, mesh  # noqa: E402

WEAK_REFINE = {1: (9, 9), 2: (18, 9), 4: (18, 18), 8: (36, 18)}


def workload(name: str, n_gpus: int, scale: float):
    """(CaseSpec, (px,py,pz), label)"""
    if name == "C2":
        rx, ry = WEAK_REFINE.get(n_gpus, (9 * n_gpus, 9))
        rx, ry = max(1, int(round(rx * scale))), max(1, int(round(ry * scale)))
        spec = cases.contraction_2d(rx, ry)
        return spec, (n_gpus, 1, 1), f"C2 2-D 4:1 planar contraction PTTLog, Contraction41 blocks x({rx},{ry})"
    spec = cases.by_name(name, scale)
    return spec, spec.decomp.get(n_gpus, (n_gpus, 1, 1)), f"{name} {spec.note}"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def measured_traffic(config: str, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture of this
    config (profiles/r1_traffic.json), or None when the capture does not hold this kernel / configuration."""
    p = ROOT / "profiles" / "r1_traffic.json"
    if not p.exists():
        return None, None
    t = json.loads(p.read_text())
    k = t.get("configs", {}).get(config, {}).get(kernel.strip("()"))
    if not k:
        return None, None
    return k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"], "profiles/r1_traffic.json (" + t.get("source", "") + ")"


def kernel_bytes_per_cell(kernel: str, dims: int, n_modes: int, n_colours: int):
    """(algorithmic bytes one launch moves per cell it processes, fraction of the mesh one launch processes).
    DESIGN.md §3; d=8, i=4, f = internal faces per cell (3 in 3-D, 2 in 2-D), K = 2f slots, c1 = solved components of one
    mode, c = c1 x modes.  Gathers from neighbours count once per distinct neighbour cell (they are L2/L1 hits otherwise)."""
    f = 3 if dims == 3 else 2
    c1 = 6 if dims == 3 else 4
    K = 2 * f
    c = c1 * n_modes
    k = kernel.strip("()")
    base = k.split("<")[0]
    half = 1.0 / max(1, n_colours)      # kernels launched per colour (hex meshes: 2 colours -> half the mesh per launch)
    row = K * (4 + 8)                   # matrix row: neighbour table + coefficients
    if base == "k_flux_assemble":
        # theta, U, tile record (nbr, meta, S, W, D, rV, V), flux tile | A, diag+rD, bsrc, corr (one value per face and component), gradU
        return (c1 + 3) * 8 + K * (4 + 4 + 7 * 8) + 16 + K * 8 + K * 8 + 16 + c1 * 8 + f * c1 * 8 + 72, 1.0
    if base == "k_cell_source2":
        # gradU, theta, thetaOld, lam, R, V, bsrc read | bsrc, fFene
        return (9 + 6 + 6 + 3 + 9 + 1) * 8 + c1 * 8 + 6 * 8 + 8, 1.0
    if base == "k_eig_tau":
        return (6 + 1) * 8 + (3 + 9 + 6) * 8, 1.0
    if base == "k_krylov_init":
        # diag, row, psi own + gathered, b, corr (inflow faces) | r, r0
        return 8 + row + c * 8 * 2 + c * 8 + f * c * 8 + 2 * c * 8, 1.0
    if base == "k_sweep":
        # rD, row, gathered y | UPD 1: r, p, v -> p, y;  UPD 2: r, v -> s, z  (average 4.5 vectors); plain: y -> y
        fused = "UPD" in k or k.rstrip(">").endswith(("1", "2"))
        return 8 + row + c * 8 + (4.5 if fused else 2) * c * 8, half
    if base == "k_spmv":
        # diag, rD, row, gathered y, y own, other | v (+ y when the backward substitution is fused: colour 0)
        fuse = k.rstrip(">").rstrip().endswith("1")
        return 16 + row + c * 8 * 4 + (c * 8 if fuse else 0), half
    if base == "k_update_p":
        return 8 + 5 * c * 8, half
    if base == "k_make_s":
        return 8 + 4 * c * 8, half
    if base == "k_update_x_r":
        return 8 * c * 8, 1.0
    if base == "k_sum_psi":
        return c * 8, 1.0
    return None, 1.0


class ClockSampler:
    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = ROOT / "gpurun_out"
        self.path.mkdir(exist_ok=True)
        self.file = self.path / f"clocks_bench_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(self.file, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for line in self.file.read_text().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_rank_mesh(spec, decomp, rank, n_ranks):
    if n_ranks == 1:
        return mesh.tensor_grid(spec.grid)
    return mesh.tensor_grid_part(spec.grid, *decomp, rank)


def initial_state(m, spec):
    """theta0 + its eigen-pairs (restart state) computed with the GPU kernel itself."""
    from rheotool_b200.stress import eig_exp
    U, Ub, phi, theta0 = m.synth_fields(spec.synth)
    return U, Ub, phi, theta0


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on stdout; stdout carries the one JSON line
    import torch
    import torch.distributed as dist
    from rheotool_b200.stress import GpuStressModel, eig_exp

    n = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != n:
        if rank == 0:
            print(f"bench.py: --gpus {n} but WORLD_SIZE={world}; launch with torchrun --nproc-per-node {n}", file=sys.stderr)
        if world == 1 and n > 1:
            sys.exit(2)
        n = world
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the stress step has no CPU fallback")
    torch.cuda.set_device(local)
    if n > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    spec, decomp, label = workload(args.config, n, args.scale)
    spec.schemes.solver = abi.SOLVER[args.solver]
    m = build_rank_mesh(spec, decomp, rank, n)
    U, Ub, phi, theta0 = m.synth_fields(spec.synth)
    # dt for face-CFL 0.2 on the GLOBAL mesh
    rate = torch.tensor([m.max_courant_rate(phi)], dtype=torch.float64, device="cuda")
    cells = torch.tensor([float(m.n_cells)], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(rate, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells, op=dist.ReduceOp.SUM)
    dt = spec.cfl / float(rate.item())
    n_cells_total = int(cells.item())

    g = GpuStressModel(m, spec.models, spec.schemes, local)
    if n > 1:
        uid = [GpuStressModel.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(rank, n, uid[0])
    for mi in range(len(spec.models)):
        th = theta0 * (1.0 + 0.1 * mi)
        vals, vecs = eig_exp(th, local)
        g.upload_state(mi, th, np.zeros_like(th), vals, vecs)
    g.upload_velocity(U, Ub, phi)
    ext = torch.cuda.ExternalStream(g.stream_ptr(), device=torch.device("cuda", local))

    def barrier():
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        g.store_old_time()
        g.correct(dt)

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    l0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = []
    barrier()
    cs0 = g.comm_stats()
    e0.record(ext)
    for _ in range(args.steps):
        one_step()
        iters.append(g.last_iterations())
    e1.record(ext)
    barrier()
    cs1 = g.comm_stats()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if sampler else None
    launches = g.launch_count() - l0
    value = n_cells_total * args.steps / (ms_total * 1e-3) / 1e6

    # ---- e2e through the C-ABI with pinned HOST buffers (upload U,U_b,phi; download tau) every step
    hU = torch.from_numpy(U).pin_memory(); hUb = torch.from_numpy(Ub).pin_memory(); hphi = torch.from_numpy(phi).pin_memory()
    htau = torch.zeros((m.n_cells, 6), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        g.correct_host(hU.data_ptr(), hUb.data_ptr(), hphi.data_ptr(), dt, True, htau.data_ptr())
    barrier()
    tb0 = g.transfer_bytes()
    t0 = time.perf_counter()
    e0.record(ext)
    for _ in range(e2e_steps):
        g.correct_host(hU.data_ptr(), hUb.data_ptr(), hphi.data_ptr(), dt, True, htau.data_ptr())
    e1.record(ext)
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = n_cells_total * e2e_steps / (float(ms2.item()) * 1e-3) / 1e6
    tb1 = g.transfer_bytes()
    h2d = (tb1[0] - tb0[0]) // e2e_steps   # counted inside the library from the copies it issued
    d2h = (tb1[1] - tb0[1]) // e2e_steps

    # ---- roofline of the dominant kernel: per-kernel CUDA-event timing pass (serialised launches)
    roof = None
    phase = None
    if rank == 0 or n > 1:
        g.set_phase_timing(True)
        one_step()
        phase = g.phase_times()
        g.set_phase_timing(False)
        g.set_kernel_timing(True)
        for _ in range(3):
            one_step()
        kt = g.kernel_times()
        g.set_kernel_timing(False)
        if rank == 0 and kt:
            tot = sum(v[1] for v in kt.values())
            top = max(kt.items(), key=lambda kv: kv[1][1])
            name, (cnt, tms) = top
            peak, peak_src = peaks()
            _, cstart = g.renumbering()
            bpc, frac_cells = kernel_bytes_per_cell(name, spec.dims, len(spec.models), len(cstart) - 1)
            cells_per_launch = m.n_cells * frac_cells
            achieved = (bpc * cells_per_launch / (tms / cnt * 1e-3) / 1e9) if bpc else None
            traffic, traffic_src = measured_traffic(args.config, name)
            roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": (bpc * cells_per_launch) if bpc else None, "peak_source": peak_src,
                    "kernel_share_of_step": tms / tot, "avg_launch_ms": tms / cnt,
                    "bytes_per_cell_launch": bpc,
                    "kernels_ms_per_step": {k: round(v[1] / 3, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])}}
            # whole-step algorithmic bytes (SURVEY §8d contract: 1480 + 1304 k (3-D), 1192 + 880 k (2-D), per mode)
            kmean = statistics.mean(iters) if iters else 0
            per_cell = ((1480 + 1304 * kmean) if spec.dims == 3 else (1192 + 880 * kmean)) * len(spec.models)
            roof["step_bytes_per_cell_contract"] = per_cell
            roof["step_achieved_gbs"] = per_cell * n_cells_total / n / (ms_total / args.steps * 1e-3) / 1e9
            roof["step_frac"] = roof["step_achieved_gbs"] / peak

    out = None
    if rank == 0:
        out = {
            "metric": "stress-step Mcell-steps/s (update+assembly+solve)", "value": value, "unit": "Mcell-steps/s",
            "n_gpus": n, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "cells_total": n_cells_total, "cells_per_gpu": m.n_cells, "dt": dt, "cfl": spec.cfl,
                       "krylov_iterations_mean": statistics.mean(iters) if iters else None,
                       "limiter": "cubista", "solver": args.solver + "+DILU", "tolerance": spec.schemes.tolerance,
                       "modes": len(spec.models), "decomposition": list(decomp),
                       "l2": "working set (~1.0 kB/cell) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "Mcell-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches,
            "comm": {"mode": cs1["mode"],
                     "halo_wait_ms_per_step": (cs1["halo_wait_ms"] - cs0["halo_wait_ms"]) / args.steps,
                     "reduce_wait_ms_per_step": (cs1["reduce_wait_ms"] - cs0["reduce_wait_ms"]) / args.steps,
                     "halo_swaps_per_step": (cs1["halo_waits"] - cs0["halo_waits"]) / args.steps,
                     "reductions_per_step": (cs1["reduce_waits"] - cs0["reduce_waits"]) / args.steps},
            "clocks": clocks,
            "roofline": roof,
            "phase_ms": phase,
        }
    g.close()
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out, (spec, label, dt)


def run_cpu_reference(args, n_steps, label_only=False, max_seconds=25.0, scale=None):
    """The reference algorithm on the host cores: oracle, one sub-domain per core (OpenMP over ranks)."""
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    spec, _, label = workload(args.config, 1, args.scale if scale is None else scale)
    m = mesh.tensor_grid(spec.grid)
    U, Ub, phi, theta0 = m.synth_fields(spec.synth)
    dt = spec.cfl / m.max_courant_rate(phi)
    R = max(1, min(cores, 64))
    # tutorials solve theta with PBiCG (fvSolution:32-45); north_star names PBiCGStab: keep the GPU arm's solver
    c2r = m.simple_decomp(R, 1, 1) if R > 1 else np.zeros(m.n_cells, dtype=np.int32)
    subs = [m.decompose(c2r, R, r) for r in range(R)] if R > 1 else [m]
    oc = orc.OracleCase([s.desc for s in subs], spec.models, spec.schemes)
    for r, s in enumerate(subs):
        ca = s.global_cells() if R > 1 else np.arange(m.n_cells)
        _, fa = s.proc_addressing() if R > 1 else (None, None)
        for mi in range(len(spec.models)):
            th = theta0[ca] * (1.0 + 0.1 * mi)
            vals, vecs = orc.calc_eig(th)
            oc.set_state(r, mi, th, np.zeros_like(th), vals, vecs)
        if R > 1:
            ph = np.where(fa > 0, phi[np.abs(fa) - 1], -phi[np.abs(fa) - 1])
            nint = s.n_internal
            oc.set_velocity(r, U[ca], _sub_ub(Ub, fa, nint, m.n_internal), ph)
        else:
            oc.set_velocity(r, U, Ub, phi)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    oc.store_old_time(); oc.step(dt)   # warm-up (page faults, first touch)
    times = []
    t_all = time.perf_counter()
    for _ in range(n_steps):
        t0 = time.perf_counter()
        oc.store_old_time(); oc.step(dt)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > max_seconds:
            break
    sec = sum(times)
    val = m.n_cells * len(times) / sec / 1e6
    return {"value": val, "unit": "Mcell-steps/s", "cores": R, "kind": "port",
            "sample": f"{len(times)} steps of {label} ({m.n_cells} cells), oracle = CPU restatement of rheoTool's algorithm, "
                      f"{R} sub-domains on {R} OpenMP threads, Krylov iterations {oc.last_iterations()}",
            "ms_per_step": sec / len(times) * 1e3, "steps": len(times)}, (m.n_cells, label, dt)


def _sub_ub(Ub, fa, nint, n_int_global):
    """boundary U of a sub-mesh: physical faces take the global patch value, processor faces are ignored."""
    gb = np.abs(fa[nint:]) - 1 - n_int_global
    out = np.zeros((len(gb), 3))
    ok = gb >= 0
    out[ok] = Ub[gb[ok]]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload per direction (tests only; 1.0 = the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--solver", default="PBiCGStab", choices=["PBiCGStab", "PBiCG"],
                    help="Krylov method of the GPU arm (PBiCG: csrc/gpu/pbicg.cuh, one GPU; the headline configuration is PBiCGStab)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb, (ncell, label, dt) = run_cpu_reference(args, args.steps, max_seconds=150.0)
        line = {"impl": "reference", "metric": "stress-step Mcell-steps/s (update+assembly+solve)", "value": cb["value"], "unit": "Mcell-steps/s",
                "n_gpus": args.gpus, "steps": cb["steps"], "warmup": 1, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": label, "cells_total": ncell, "dt": dt},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "Mcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "restated CPU baseline (rheoTool algorithm, not the rheoTool binary: OpenFOAM-9/Eigen/MPI are not installable here); the restatement is pinned on rheoTool's own text compiled into oracle/_ref (tests/test_reference_pin.py), whose stand-in linear solver is not rheoTool's and is therefore not what is timed"}
        print(json.dumps(line))
        return

    out, _ = run_ours(args)
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            cb, _ = run_cpu_reference(args, 5, max_seconds=25.0)
            out["cpu_baseline"] = cb
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))


if __name__ == "__main__":
    main()
