#!/bin/bash
# registers / spills per kernel of engine.cu (nvcc -Xptxas -v, sm_100a) — demangled, one line per kernel
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  -Xptxas -v -I include -I rheotool_b200/csrc/host -I rheotool_b200/csrc/gpu -c rheotool_b200/csrc/gpu/engine.cu -o /tmp/engine_ptxas.o 2>&1 \
 | awk '/Compiling entry function/ {match($0, /\x27[^\x27]+\x27/); name=substr($0, RSTART+1, RLENGTH-2)}
        /bytes stack frame/ {spill=$0; sub(/^ */,"",spill)}
        /Used [0-9]+ registers/ {match($0,/Used [0-9]+ registers/); print name "\t" substr($0,RSTART,RLENGTH) "\t" spill}' \
 | while IFS=$'\t' read -r n r s; do echo "$(echo "$n" | c++filt | sed 's/(.*//; s/void //; s/rk:://') | $r | $s"; done
