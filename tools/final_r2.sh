#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): tests, bench lines, per-kernel tables, ncu launch list and --set full captures.
# Everything lands in gpurun_out/final/ as text (the .ncu-rep files are summarised on the box and removed: 64 MiB merge limit).
O=gpurun_out/final; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; tail -3 $O/tests_gpu.log
python bench.py --steps 20 --warmup 5 --kernel-profile $O/kernel_breakdown_psp_bf16x3.md > $O/bench_n1_bf16x3.json 2> $O/bench_n1_bf16x3.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --model ocr --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --kernel-profile $O/kernel_breakdown_ocr_bf16x3.md > $O/bench_n1_ocr_bf16x3.json 2> $O/bench_ocr.err
python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library-baseline --kernel-profile $O/kernel_breakdown_psp_bf16.md > $O/bench_n1_bf16.json 2> $O/bench_bf16.err
python tools/bench_conv.py > $O/conv_shapes_bf16x3.txt 2>&1
python tools/bench_conv.py --precision bf16 > $O/conv_shapes_bf16.txt 2>&1
python tools/bench_bn.py > $O/bn_shapes.txt 2>&1
# --- ncu: launch list of the bench command (shares, not absolutes) ---
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile-run --steps 1 --warmup 1 > $O/launches.log 2>&1
python tools/ncu_summary.py launches $O/launches.csv > $O/launches.md; rm -f $O/launches.csv
# --- ncu --set full: the dominant kernels ---
$NCU -k regex:conv_tc2_kernel -c 3 -o $O/conv_tc2_3x3 -f python tools/bench_conv.py --only "l3 3x3" --iters 1 > /dev/null 2>&1
$NCU -k regex:wgrad_tc2_kernel -c 2 -o $O/wgrad_tc2_3x3 -f python tools/bench_conv.py --only "l3 3x3" --iters 1 > /dev/null 2>&1
$NCU -k regex:conv_tc2_kernel -c 6 -o $O/conv_tc2_1x1 -f python tools/bench_conv.py --only "l3 1x1 256->1024" --iters 1 > /dev/null 2>&1
$NCU -k regex:wgrad_tc_kernel -c 2 -o $O/wgrad_tc_rows -f python tools/bench_conv.py --only "stem 3x3 64->128" --iters 1 > /dev/null 2>&1
$NCU -k regex:conv_tc_kernel -c 6 -o $O/conv_tc_64 -f python tools/bench_conv.py --only "stem 3x3 64->64" --iters 1 > /dev/null 2>&1
$NCU -k regex:"bn_act_fwd|bn_bwd" -c 66 -o $O/bn -f python tools/bench_bn.py --iters 1 > /dev/null 2>&1
$NCU -k regex:"ocr_attn|region_softmax|attn_softmax|wgrad_tc_kernel" -c 10 -o $O/ocr -f python bench.py --model ocr --profile-run --steps 1 --warmup 0 > /dev/null 2>&1
$NCU -k regex:"ppm_|tcb_|nll_|logsoftmax|sgd_momentum" -c 16 -o $O/tail -f python bench.py --profile-run --steps 1 --warmup 0 > /dev/null 2>&1
for r in conv_tc2_3x3 wgrad_tc2_3x3 conv_tc2_1x1 wgrad_tc_rows conv_tc_64 bn ocr tail; do
  python tools/ncu_summary.py full $O/$r.ncu-rep > $O/ncu_$r.md 2>&1
done
python tools/ncu_traffic.py $O/conv_tc2_3x3.ncu-rep --kernel conv_tc2_kernel --pick median --launch "layer3 3x3 d2 256->256 fwd (bf16x3), 10x60x107" --algorithmic-bytes 134.0e6 --out $O/r2_traffic.json > $O/traffic.log 2>&1
ls -la $O/*.ncu-rep > $O/ncu_reports.txt 2>&1
rm -f $O/bn.ncu-rep $O/ocr.ncu-rep $O/tail.ncu-rep $O/conv_tc_64.ncu-rep $O/wgrad_tc_rows.ncu-rep $O/conv_tc2_1x1.ncu-rep
du -sh $O
