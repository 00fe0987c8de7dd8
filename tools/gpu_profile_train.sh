#!/bin/bash
# ncu evidence for the training kernels (B=32 per GPU) + BASELINE config 5 inference numbers.
mkdir -p gpurun_out
CMD="python bench_train.py --steps 1 --warmup 3 --batch 32"
HK_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/train_plain.log 2>&1 &&
HK_TRAIN_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_kernel" -s 105 -c 8 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo "wgrad full rc=$?"
HK_TRAIN_NO_GRAPH=1 $CMD > gpurun_out/train_plain2.log 2>&1 &&
HK_TRAIN_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_apply|bn_bwd_partial|bn_apply_fwd|bn_stats_partial|maxpool_bwd|maxpool_argmax|stem_wgrad_partial|bce_fwd_bwd" -s 330 -c 14 -o gpurun_out/prof_train_small $CMD > gpurun_out/ncu_train_small.log 2>&1
echo "train small full rc=$?"
for K in 4 16 32; do
  timeout 300 python bench.py --height 960 --width 1280 --batch 16 --keypoints $K --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg5_k$K.log 2>&1; echo "cfg5 K=$K rc=$?"; tail -1 gpurun_out/bench_cfg5_k$K.log | cut -c1-260
done
