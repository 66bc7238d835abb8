#!/bin/bash
# A/B of multi-GPU switches inside ONE box: ms/step of bench.py under torchrun per setting, repeated.   tools/ab_n2.sh OUT NGPU cfg...
out=${1:-gpurun_out/ab_n2.txt}; n=${2:-2}; shift; shift
: > $out
for rep in 1 2; do
  for cfg in "$@"; do
    ms=$(env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline 2>gpurun_out/ab_n2.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], (d.get('local_bn') or {}).get('ms_per_step'), d['config'].get('grad_allreduce'))")
    echo "rep$rep N=$n $cfg $ms" | tee -a $out
  done
done
