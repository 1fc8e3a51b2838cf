#!/bin/bash
# Round-2 measurement evidence, one call on the GPU box:  bash tools/capture_profiles.sh   (writes gpurun_out/r2/)
# ncu launch lists are taken on one stream without the CUDA graph so that ncu sees every launch (cold caches, serialised:
# compare shares, not absolutes); numbers printed by runs under ncu are never bench values.
set -u
O=gpurun_out/r2
mkdir -p $O
COMMON="--no-cpu --no-also --no-census --no-gpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_train_step.csv \
    python bench.py --profile-step --warmup 1 --no-graph --no-multistream $COMMON > $O/ncu_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_infer_step.csv \
    python bench.py --mode infer --profile-step --warmup 1 --no-graph --no-multistream $COMMON > $O/ncu_infer.log 2>&1
if [ -z "${QUICK:-}" ]; then   # QUICK=1: skip the two --set full captures (kernels unchanged since the last capture)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tma_kernel -c 1 --launch-skip 3 -f -o $O/ncu_conv_tma_b4 \
    python tools/bench_layers.py b4_img > $O/ncu_full_conv.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:wgrad_tma2_kernel -c 1 --launch-skip 3 -f -o $O/ncu_wgrad_tma2_b4 \
    python tools/bench_layers.py wg_b4_img > $O/ncu_full_wgrad.log 2>&1
fi
timeout 600 python tools/bench_layers.py > $O/layer_timings.txt 2>&1
for b in 1 8 32 128; do
  timeout 300 python bench.py --mode infer --batch $b $COMMON 2>/dev/null | tail -1 > $O/bench_infer_b$b.json
done
for b in 16 32; do
  timeout 300 python bench.py --mode train --batch $b $COMMON 2>/dev/null | tail -1 > $O/bench_train_b$b.json
done
timeout 300 python tools/timeline.py train 8 > $O/timeline_train_step.txt 2>&1
timeout 300 python tools/timeline.py infer 8 > $O/timeline_infer_step.txt 2>&1
timeout 300 python tools/bench_packbatch.py > $O/packbatch.txt 2>&1
timeout 300 python tools/bench_bn_bwd.py > $O/bn_bwd.txt 2>&1
ls -la $O
