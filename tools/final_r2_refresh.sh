#!/bin/bash
# Refresh of the one-GPU evidence after the last kernel changes (no ncu --set full captures: those are per-kernel and unchanged).
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; tail -3 $O/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python tools/probe_hbm.py > $O/hbm_probe.txt 2>&1; cat $O/hbm_probe.txt
python bench.py --steps 20 --warmup 5 --kernel-profile $O/kernel_breakdown_psp_bf16x3.md > $O/bench_n1_bf16x3.json 2> $O/bench_n1_bf16x3.err
python bench.py --model ocr --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --kernel-profile $O/kernel_breakdown_ocr_bf16x3.md > $O/bench_n1_ocr_bf16x3.json 2> $O/bench_ocr.err
python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --kernel-profile $O/kernel_breakdown_psp_bf16.md > $O/bench_n1_bf16.json 2> $O/bench_bf16.err
python tools/bench_conv.py > $O/conv_shapes_bf16x3.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile-run --steps 1 --warmup 1 > $O/launches.log 2>&1
python tools/ncu_summary.py launches $O/launches.csv > $O/launches.md; rm -f $O/launches.csv
ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -c 6 -o $O/conv_tc2_1x1 -f python tools/bench_conv.py --only "l3 1x1 256->1024" --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py full $O/conv_tc2_1x1.ncu-rep > $O/ncu_conv_tc2_1x1.md 2>&1; rm -f $O/conv_tc2_1x1.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -c 2 -o $O/wgrad_tc_rows -f python tools/bench_conv.py --only "stem 3x3 64->128" --iters 1 > /dev/null 2>&1
python tools/ncu_summary.py full $O/wgrad_tc_rows.ncu-rep > $O/ncu_wgrad_tc_rows.md 2>&1; rm -f $O/wgrad_tc_rows.ncu-rep
tail -c 300 $O/bench_n1_bf16x3.json
