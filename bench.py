#!/usr/bin/env python
"""bench.py — stress-step throughput (Mcell-steps/s) of the B200 path, with roofline and CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
  * a "step" = one constitutiveEq::correct() (update + assembly + solve of all valid components, all
    modes + eig/exp/tau + tau BCs) with U/phi already resident in HBM;
  * workload (every N) = BASELINE.json configs[4]: C5, FENE-PLog lid-driven cavity, 400^3 = 64,000,000 cells (the mesh the
    metric's scaling target is quoted on; ~114 GiB of HBM on one B200).  N > 1 (torchrun, one rank per GPU) = the SAME global
    mesh decomposed like decomposePar `simple` 2:(2,1,1) 4:(2,2,1) 8:(2,2,2): STRONG scaling.  `--config C2|C3|C4` select
    the other BASELINE configurations; `--config C2 --weak` is round 1's weak-scaling family (~971k cells per GPU);
  * every run first solves a shrunk replica of the workload on the same ranks and compares it with the CPU oracle on ONE rank
    (`parity`), and prints a checksum of the full-size state after the timed steps (`state_check`: equal across N);
  * `e2e` = the same metric through the C-ABI with pinned HOST buffers: U, U_b, phi up, rheo_gpu_correct(), then the explicit
    part of constitutiveEq::divTau down (rheo_gpu_div_tau: 3 doubles per cell); `e2e_tau_download` = round 1's variant (tau
    itself, 6 doubles per cell, comes back);
  * `--impl reference` times the CPU restatement of the reference algorithm (oracle/, OpenMP over sub-domains = one per
    host core, thread count set explicitly) on a bounded sample of the same workload (the same model, schemes and CFL on the
    1/8 sub-cube a rank of the 8-GPU run owns) — the reference binary itself cannot be built without OpenFOAM-9/Eigen/MPI
    (DESIGN.md).
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


STRONG_DECOMP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# shrunk replica of each configuration for the on-box parity check (cells: C2 ~12k, C3 ~74k, C4 ~14k x 4 modes, C5 64k)
PARITY_SCALE = {"C1stock": 1.0, "C1": 0.25, "C2": 1 / 9, "C3": 4 / 19, "C4": 24 / 252, "C5": 40 / 400}
# bounded sample of the CPU arm: the workload shrunk per direction (C5: 200^3 = the sub-cube one rank of the 8-GPU run owns)
CPU_SAMPLE_SCALE = {"C1stock": 1.0, "C1": 1.0, "C2": 1.0, "C3": 0.5, "C4": 0.5, "C5": 0.5}


def workload(name: str, n_gpus: int, scale: float, weak: bool = False):
    """(CaseSpec, (px,py,pz), label, "strong"|"weak")"""
    if name == "C2" and weak:
        rx, ry = WEAK_REFINE.get(n_gpus, (9 * n_gpus, 9))
        rx, ry = max(1, int(round(rx * scale))), max(1, int(round(ry * scale)))
        spec = cases.contraction_2d(rx, ry)
        return spec, (n_gpus, 1, 1), f"C2 2-D 4:1 planar contraction PTTLog, Contraction41 blocks x({rx},{ry})", "weak"
    spec = cases.by_name(name, scale)
    if spec.stock and n_gpus != 1:
        raise SystemExit("bench.py: --config C1stock (the 24,894-cell tutorial mesh) runs on one GPU")
    decomp = STRONG_DECOMP.get(n_gpus, (n_gpus, 1, 1)) if spec.dims == 3 else (n_gpus, 1, 1)
    return spec, decomp, f"{name} {spec.note}", "strong"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def measured_traffic(config: str, kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu capture of this config
    (profiles/r2_traffic.json, else round 1's profiles/r1_traffic.json), or None when no capture holds this kernel /
    configuration."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = ROOT / "profiles" / name
        if not p.exists():
            continue
        t = json.loads(p.read_text())
        k = t.get("configs", {}).get(config, {}).get(kernel.strip("()"))
        if k:
            return k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"], f"profiles/{name} (" + t.get("source", "") + ")"
    return None, None


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
    if base == "k_flux3":
        # assembly3.cuh: theta, U, tile record v3 (nbr, meta, G, D, G0, V), flux tile | A, diag+rD, rowsum, inflow mask, bsrc, A theta,
        # corr (one value per face and component), gradU
        return (c1 + 3) * 8 + K * (4 + 4 + 6 * 8) + 32 + K * 8 + K * 8 + 16 + 8 + 4 + c1 * 8 + c1 * 8 + f * c1 * 8 + 72, 1.0
    if base == "k_source_init":
        # gradU, theta, thetaOld, lam, R, V, bsrc, A theta, rowsum, inflow mask, corr (inflow faces) | fFene, r, r0   (one mode per launch)
        return (9 + 6 + 6 + 3 + 9 + 1) * 8 + c1 * 8 + c1 * 8 + 8 + 4 + f * c1 * 8 + 8 + 2 * c1 * 8, 1.0
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
    if base == "k_bsweep":
        # block ordering (blocksweep.cuh): k_bsweep<NR, KT, DIR, UPD, FIRST> over one chunk colour
        # rD, row, in/out-of-chunk gathers (one record per distinct out-of-chunk neighbour: ~0.5/cell) | UPD 1: r, p, v -> p, y;  UPD 2: r, v -> s, z
        # (average 4.5 vectors);  DIR 2 (backward only): y -> y
        args = [a.strip() for a in k[k.index("<") + 1:k.rindex(">")].split(",")]
        backward_only = len(args) >= 3 and args[2] == "2"
        return 8 + row + 0.5 * c * 8 + (2 if backward_only else 4.5) * c * 8, half
    if base == "k_bspmv0":
        # diag, rD, row, out-of-chunk gathers, y own, other | y, v
        return 16 + row + 0.5 * c * 8 + 4 * c * 8, half
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


def synth(spec, m):
    """(U, U_b, phi, theta0, theta_b or None): the synthetic fields of the workload on mesh `m`"""
    if spec.stock:
        return cases.stock_fields(spec, m)
    return (*m.synth_fields(spec.synth), None)


def build_rank_mesh(spec, decomp, rank, n_ranks):
    if spec.stock:
        return cases.stock_mesh(spec)
    if n_ranks == 1:
        return mesh.tensor_grid(spec.grid)
    return mesh.tensor_grid_part(spec.grid, *decomp, rank)


def _tight(schemes, tol=1e-15):
    c = abi.RheoSchemeCtl()
    for f, _ in abi.RheoSchemeCtl._fields_:
        setattr(c, f, getattr(schemes, f))
    c.tolerance = tol
    return c


def _rel_l2(a, b):
    den = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def parity_replica(args, n, rank, local, decomp, dist, steps=2):
    """The workload shrunk to oracle size, decomposed over the SAME ranks with the SAME decomposition and solved by the
    same library calls (halo swaps and reductions over NVLink included); rank 0 gathers theta/tau and compares them with the
    CPU oracle run on ONE rank (the checker; never timed here).  Returns the `parity` object on rank 0, None elsewhere."""
    import torch
    from rheotool_b200.stress import GpuStressModel, eig_exp
    scale = PARITY_SCALE[args.config]
    spec, _, label, _ = workload(args.config, n, scale, False)
    spec.schemes.solver = abi.SOLVER[args.solver]
    sc = _tight(spec.schemes)
    part = build_rank_mesh(spec, decomp, rank, n)
    U, Ub, phi, theta0, theta_b = synth(spec, part)
    rate = torch.tensor([part.max_courant_rate(phi)], dtype=torch.float64, device="cuda")
    if n > 1:
        dist.all_reduce(rate, op=dist.ReduceOp.MAX)
    dt = spec.cfl / float(rate.item())
    g = GpuStressModel(part, spec.models, sc, local)
    if n > 1:
        uid = [GpuStressModel.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(rank, n, uid[0])
    for mi in range(len(spec.models)):
        th = theta0 * (1.0 + 0.1 * mi)
        vals, vecs = eig_exp(th, local)
        g.upload_state(mi, th, np.zeros_like(th), vals, vecs, theta_b=theta_b)
    g.upload_velocity(U, Ub, phi)
    its = []
    for _ in range(steps):
        g.store_old_time(); g.correct(dt)
        its.append(g.last_iterations())
    mine = {"cells": part.global_cells() if n > 1 else np.arange(part.n_cells),
            "theta": [g.theta(mi) for mi in range(len(spec.models))], "tau": [g.tau(mi) for mi in range(len(spec.models))]}
    mode = g.comm_stats()["mode"]
    g.close()
    parts = [mine]
    if n > 1:
        parts = [None] * n if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    from oracle import oracle as orc   # the checker
    full = build_rank_mesh(spec, (1, 1, 1), 0, 1)
    Uf, Ubf, phif, th0, thb = synth(spec, full)
    oc = orc.OracleCase([full.desc], spec.models, sc)
    for mi in range(len(spec.models)):
        th = th0 * (1.0 + 0.1 * mi)
        vals, vecs = orc.calc_eig(th)
        oc.set_state(0, mi, th, np.zeros_like(th), vals, vecs, theta_b=thb)
    oc.set_velocity(0, Uf, Ubf, phif)
    for _ in range(steps):
        oc.store_old_time(); oc.step(dt)
    e_th = e_tau = 0.0
    for mi in range(len(spec.models)):
        for key, fld in (("theta", abi.FIELD_THETA), ("tau", abi.FIELD_TAU)):
            ref = oc.get(0, mi, fld)
            got = np.full_like(ref, np.nan)
            for d in parts:
                got[d["cells"]] = d[key][mi]
            e = _rel_l2(got, ref)
            if key == "theta":
                e_th = max(e_th, e)
            else:
                e_tau = max(e_tau, e)
    ok = bool(np.isfinite(e_th) and np.isfinite(e_tau) and e_th <= 1e-10 and e_tau <= 1e-10)
    return {"relL2_theta": e_th, "relL2_tau": e_tau, "tolerance": 1e-10, "ok": ok, "against": "CPU oracle on one rank (oracle/, pinned on the reference text)",
            "replica": f"{label}, {full.n_cells} cells, {steps} steps, solver tolerance 1e-15", "ranks": n, "decomposition": list(decomp),
            "comm": mode, "krylov_iterations": its}


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
    verbose = rank == 0 and bool(os.environ.get("RHEO_BENCH_VERBOSE"))
    t_start = time.perf_counter()

    spec, decomp, label, scaling = workload(args.config, n, args.scale, args.weak)
    spec.schemes.solver = abi.SOLVER[args.solver]

    # ---- parity first: a run whose small replica disagrees with the oracle prints no throughput
    parity = None if args.no_parity else parity_replica(args, n, rank, local, decomp, dist)
    if rank == 0 and parity is not None and not parity["ok"]:
        print(json.dumps({"error": "parity replica disagrees with the CPU oracle", "parity": parity}))
        sys.exit(3)
    if verbose:
        print(f"bench: parity replica done at {time.perf_counter() - t_start:.1f} s: {parity}", file=sys.stderr)

    m = build_rank_mesh(spec, decomp, rank, n)
    U, Ub, phi, theta0, theta_b = synth(spec, m)
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
        g.upload_state(mi, th, np.zeros_like(th), vals, vecs, theta_b=theta_b)
        del vals, vecs, th
    g.upload_velocity(U, Ub, phi)
    del theta0
    if verbose:
        free_b, total_b = torch.cuda.mem_get_info()
        print(f"bench: setup done at {time.perf_counter() - t_start:.1f} s, HBM used {(total_b - free_b) / 2**30:.1f} GiB of {total_b / 2**30:.1f}", file=sys.stderr)
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
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        one_step()
        iters.append(g.last_iterations())
        step_ev[i].record(ext)   # each step ends with the host reading the solver's control block, so this adds no synchronisation
    e1.record(ext)
    barrier()
    cs1 = g.comm_stats()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    step_ms = [round((e0 if i == 0 else step_ev[i - 1]).elapsed_time(step_ev[i]), 3) for i in range(args.steps)]
    if n > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if sampler else None
    launches = g.launch_count() - l0
    value = n_cells_total * args.steps / (ms_total * 1e-3) / 1e6

    # ---- size-independent check of the full-size state after warmup + steps: the same global mesh, dt and step count give
    # the same fields on 1, 2, 4 or 8 ranks (to the solver tolerance), so these sums must agree across the N of a scaling run
    chk = torch.zeros(3, dtype=torch.float64)
    th = g.theta(0)
    chk[0] = float(np.abs(th).sum()); chk[1] = float(th.sum())
    del th
    ta = g.tau(0)
    chk[2] = float(np.abs(ta).sum())
    finite = bool(np.isfinite(ta).all())
    del ta
    chk = chk.cuda()
    if n > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    chk = chk.cpu().tolist()
    state_check = {"after_steps": args.warmup + args.steps, "theta_abs_sum": chk[0], "theta_sum": chk[1], "tau_abs_sum": chk[2],
                   "finite": finite, "note": "mode 0, summed over all ranks; equal across N for the same --config/--steps/--warmup"}

    # ---- e2e through the C-ABI with pinned HOST buffers, every step: upload U, U_b, phi; correct(); download what the CPU
    # momentum predictor needs.  Two variants of that last leg:
    #   e2e                the explicit part of constitutiveEq::divTau evaluated on the device (rheo_gpu_div_tau, stabilization
    #                      coupling — 47 of the 52 tutorials), 3 doubles per cell come back; tau stays in HBM
    #   e2e_tau_download   round 1's call: tau itself (6 doubles per cell) comes back and the host evaluates divTau
    hU = torch.from_numpy(U).pin_memory(); hUb = torch.from_numpy(Ub).pin_memory(); hphi = torch.from_numpy(phi).pin_memory()
    htau = torch.zeros((m.n_cells, 6), dtype=torch.float64).pin_memory()
    hdiv = torch.zeros((m.n_cells, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_tau():
        g.correct_host(hU.data_ptr(), hUb.data_ptr(), hphi.data_ptr(), dt, True, htau.data_ptr())

    def e2e_div():
        g.correct_host(hU.data_ptr(), hUb.data_ptr(), hphi.data_ptr(), dt, True, None)
        g.div_tau_host(abi.STAB_COUPLING, hdiv.data_ptr())

    def time_e2e(call):
        for _ in range(2):
            call()
        barrier()
        tb0 = g.transfer_bytes()
        t0 = time.perf_counter()
        e0.record(ext)
        for _ in range(e2e_steps):
            call()
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], dtype=torch.float64, device="cuda")
        if n > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        tb1 = g.transfer_bytes()
        # bytes counted inside the library from the copies it issued
        return n_cells_total * e2e_steps / (float(ms2.item()) * 1e-3) / 1e6, (tb1[0] - tb0[0]) // e2e_steps, (tb1[1] - tb0[1]) // e2e_steps

    e2e_tau_value, h2d_tau, d2h_tau = time_e2e(e2e_tau)
    e2e_value, h2d, d2h = time_e2e(e2e_div)
    e2e_div_finite = bool(np.isfinite(hdiv.numpy()).all())

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
            # the committed ncu capture is of the N = 1 run at full size: its bytes per launch say nothing about a rank's share of a
            # decomposed or shrunk mesh
            traffic, traffic_src = measured_traffic(args.config, name) if (n == 1 and args.scale == 1.0) else (None, None)
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
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "cells_total": n_cells_total, "cells_per_gpu": m.n_cells, "dt": dt, "cfl": spec.cfl,
                       "krylov_iterations_mean": statistics.mean(iters) if iters else None,
                       "krylov_iterations_per_step": iters, "ms_per_timed_step": step_ms,
                       "ordering": g.ordering(),
                       "limiter": "cubista", "solver": args.solver + "+DILU", "tolerance": spec.schemes.tolerance,
                       "modes": len(spec.models), "decomposition": list(decomp),
                       "l2": "working set (~1.7 kB/cell) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "Mcell-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "call": "rheo_gpu_correct(U, U_b, phi -> device) + rheo_gpu_div_tau(coupling -> host): 3 doubles per cell back", "finite": e2e_div_finite},
            "e2e_tau_download": {"value": e2e_tau_value, "unit": "Mcell-steps/s", "h2d_bytes_per_step": h2d_tau, "d2h_bytes_per_step": d2h_tau, "steps": e2e_steps,
                                 "call": "rheo_gpu_correct(U, U_b, phi -> device; tau -> host): 6 doubles per cell back (round 1's e2e)"},
            "gpu_launches": launches,
            "parity": parity,
            "state_check": state_check,
            "comm": {"mode": cs1["mode"],
                     "halo_wait_ms_per_step": (cs1["halo_wait_ms"] - cs0["halo_wait_ms"]) / args.steps,
                     "reduce_wait_ms_per_step": (cs1["reduce_wait_ms"] - cs0["reduce_wait_ms"]) / args.steps,
                     "halo_swaps_per_step": (cs1["halo_waits"] - cs0["halo_waits"]) / args.steps,
                     "reductions_per_step": (cs1["reduce_waits"] - cs0["reduce_waits"]) / args.steps},
            "clocks": clocks,
            "roofline": roof,
            "phase_ms": phase,
        }
        if roof is not None:
            peer = sum(v for k, v in roof["kernels_ms_per_step"].items() if k.startswith("k_peer"))
            out["comm"]["peer_kernels_ms_per_step"] = round(peer, 4)
            out["comm"]["peer_kernels_share_of_step"] = peer / max(1e-12, sum(roof["kernels_ms_per_step"].values()))
    g.close()
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out, (spec, label, dt)


def run_cpu_reference(args, n_steps, max_seconds=25.0, warmup=1):
    """The reference algorithm on the host cores: the oracle, one sub-domain per core (OpenMP threads standing in for MPI
    ranks), on a BOUNDED sample of the GPU arm's workload: the same configuration (model, schemes, CFL, synthetic fields)
    shrunk per direction by CPU_SAMPLE_SCALE (C5: the 200^3 sub-cube one rank of the 8-GPU run owns)."""
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    scale = args.scale * CPU_SAMPLE_SCALE[args.config]
    full_spec, _, full_label, _ = workload(args.config, 1, args.scale, False)
    spec, _, label, _ = workload(args.config, 1, scale, False)
    spec.schemes.solver = abi.SOLVER[args.solver]
    m = build_rank_mesh(spec, (1, 1, 1), 0, 1)
    U, Ub, phi, theta0, _ = synth(spec, m)
    dt = spec.cfl / m.max_courant_rate(phi)
    R = max(1, min(cores, 64))
    threads = orc.set_num_threads(R)   # explicit: torchrun exports OMP_NUM_THREADS=1
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
    for _ in range(max(1, warmup)):
        oc.store_old_time(); oc.step(dt)   # warm-up (page faults, first touch)
    times, its = [], []
    t_all = time.perf_counter()
    for _ in range(n_steps):
        t0 = time.perf_counter()
        oc.store_old_time(); oc.step(dt)
        times.append(time.perf_counter() - t0)
        its.append(oc.last_iterations())
        if time.perf_counter() - t_all > max_seconds:
            break
    sec = sum(times)
    val = m.n_cells * len(times) / sec / 1e6
    sample = (f"{len(times)} steps of {label} ({m.n_cells} cells"
              + (f" = the workload {full_label} shrunk x{CPU_SAMPLE_SCALE[args.config]} per direction" if scale != args.scale else "")
              + f"), oracle = CPU restatement of rheoTool's algorithm in the reference's cell order, {R} sub-domains on {threads} OpenMP threads, "
              f"Krylov iterations {max(its)}")
    return {"value": val, "unit": "Mcell-steps/s", "cores": threads, "kind": "port", "sample": sample,
            "ms_per_step": sec / len(times) * 1e3, "steps": len(times), "cells": m.n_cells,
            "krylov_iterations": max(its)}, (full_spec, full_label, dt)


K_SAMPLE_SCALE = {"C1stock": 1.0, "C1": 1.0, "C2": 0.45, "C3": 0.5, "C4": 0.25, "C5": 0.25}


def reference_ordering_iterations(args, steps=2):
    """Krylov iterations the REFERENCE's cell order needs: the oracle on ONE sub-domain (no block-Jacobi cut of DILU) in the
    mesh generator's natural order, on the workload shrunk per direction by K_SAMPLE_SCALE (same model, schemes, CFL)."""
    from oracle import oracle as orc
    spec, _, label, _ = workload(args.config, 1, args.scale * K_SAMPLE_SCALE[args.config], False)
    spec.schemes.solver = abi.SOLVER[args.solver]
    m = build_rank_mesh(spec, (1, 1, 1), 0, 1)
    U, Ub, phi, theta0, _ = synth(spec, m)
    dt = spec.cfl / m.max_courant_rate(phi)
    oc = orc.OracleCase([m.desc], spec.models, spec.schemes)
    for mi in range(len(spec.models)):
        th = theta0 * (1.0 + 0.1 * mi)
        vals, vecs = orc.calc_eig(th)
        oc.set_state(0, mi, th, np.zeros_like(th), vals, vecs)
    oc.set_velocity(0, U, Ub, phi)
    k = 0
    for _ in range(steps):
        oc.store_old_time(); oc.step(dt)
        k = oc.last_iterations()
    return k, f"{label} ({m.n_cells} cells), one sub-domain, {steps} steps"


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
    ap.add_argument("--config", default="C5", choices=["C1", "C1stock", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json configuration; default C5 (64 M cells: the mesh the scaling target names), strong scaling over --gpus")
    ap.add_argument("--weak", action="store_true", help="with --config C2: round 1's weak-scaling family (~971k cells per GPU)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload per direction (tests only; 1.0 = the named config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the shrunk-replica parity check against the oracle")
    ap.add_argument("--solver", default="PBiCGStab", choices=["PBiCGStab", "PBiCG"],
                    help="Krylov method of the GPU arm (PBiCG: csrc/gpu/pbicg.cuh, one GPU; the headline configuration is PBiCGStab)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        cb, (spec, label, dt) = run_cpu_reference(args, args.steps, max_seconds=150.0, warmup=args.warmup)
        n_full = int(np.prod([len(a) - 1 for a in (spec.grid.xs, spec.grid.ys, spec.grid.zs)])) if (spec.grid is not None and len(spec.grid.boxes) == 1) else None
        _, _, _, scaling = workload(args.config, args.gpus, args.scale, args.weak)
        line = {"impl": "reference", "metric": "stress-step Mcell-steps/s (update+assembly+solve)", "value": cb["value"], "unit": "Mcell-steps/s",
                "n_gpus": args.gpus, "steps": cb["steps"], "warmup": max(1, args.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": label, "cells_total": n_full, "cells_timed": cb["cells"], "dt_of_sample": dt,
                           "krylov_iterations_mean": cb["krylov_iterations"]},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "Mcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "restated CPU baseline (rheoTool algorithm, not the rheoTool binary: OpenFOAM-9/Eigen/MPI are not installable here); the restatement is pinned on rheoTool's own text compiled into oracle/_ref (tests/test_reference_pin.py), whose stand-in linear solver is not rheoTool's and is therefore not what is timed. "
                        "Mcell-steps/s of a CPU is size-independent at these sizes (every sub-domain is far larger than the caches), so the bounded sample stands for the full mesh; the speed-up over all host cores is well below linear (memory-bound)"}
        print(json.dumps(line))
        return

    out, _ = run_ours(args)
    if rank == 0:
        if args.gpus == 1 and not args.no_cpu_baseline:
            cb, _ = run_cpu_reference(args, 5, max_seconds=25.0)
            out["cpu_baseline"] = cb
            kref, ksample = reference_ordering_iterations(args)
            out["config"]["krylov_iterations_reference_ordering"] = kref
            out["config"]["krylov_iterations_reference_ordering_note"] = "CPU oracle in the reference's (natural) cell order on " + ksample
            roof = out.get("roofline")
            if roof and roof.get("step_frac") is not None:
                spec_dims = 2 if args.config in ("C1", "C1stock", "C2") else 3
                modes = out["config"]["modes"]
                per_cell_ref = ((1480 + 1304 * kref) if spec_dims == 3 else (1192 + 880 * kref)) * modes
                roof["step_bytes_per_cell_contract_reference_k"] = per_cell_ref
                roof["step_frac_reference_k"] = roof["step_frac"] * per_cell_ref / roof["step_bytes_per_cell_contract"]
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))


if __name__ == "__main__":
    main()
