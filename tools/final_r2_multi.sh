#!/bin/bash
# Round-2 multi-GPU evidence: bench.py under torchrun for each N given (weak scaling, SyncBN on = the product default), plus the
# torchrun SyncBN / overlapped all-reduce test when exactly 2 GPUs are visible.   tools/final_r2_multi.sh 8 4
O=gpurun_out/final; mkdir -p $O
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n${n}_bf16x3.json 2> $O/bench_n${n}.err
  tail -c 400 $O/bench_n${n}_bf16x3.json; echo
done
if [ "$(nvidia-smi -L | wc -l)" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_syncbn.py -m gpu -q > $O/tests_gpu_2ranks.log 2>&1; tail -3 $O/tests_gpu_2ranks.log
fi
