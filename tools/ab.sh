#!/bin/bash
# A/B of engine switches inside ONE box (clocks differ by ~1 % between boxes): ms/step of bench.py per setting, repeated.
out=${1:-gpurun_out/ab.txt}; shift
: > $out
for rep in 1 2; do
  for cfg in "$@"; do
    ms=$(env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_step'])")
    echo "rep$rep $cfg $ms" | tee -a $out
  done
done
