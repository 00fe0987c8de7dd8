#!/bin/bash
# Informational lines for profiles/: the same-GPU torch/cuDNN competitor (with clocks) and BASELINE configs[4] (960x1280, K = 4/16/32).
mkdir -p gpurun_out
OUT=gpurun_out/r02_competitor_and_config5.txt
: > $OUT
echo "# $(date -u +%FT%TZ)  $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader)" >> $OUT
for dt in bf16 fp32; do
  steps=30; [ $dt = fp32 ] && steps=5
  echo "## python bench.py --impl torch_gpu --torch-dtype $dt --steps $steps --warmup 3   (same network through torch/cuDNN on this GPU)" >> $OUT
  ( nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 200 > gpurun_out/clk_$dt.log & echo $! > gpurun_out/clk.pid )
  timeout 600 python bench.py --impl torch_gpu --torch-dtype $dt --steps $steps --warmup 3 2>/dev/null | grep '^{' >> $OUT
  kill $(cat gpurun_out/clk.pid) 2>/dev/null
  echo "clocks under load (sm MHz, max, W, reasons), median-ish sample: $(sort gpurun_out/clk_$dt.log | sed -n "$(( $(wc -l < gpurun_out/clk_$dt.log) / 2 + 1 ))p")" >> $OUT
done
for K in 4 16 32; do
  echo "## python bench.py --height 960 --width 1280 --batch 16 --keypoints $K --steps 30 --warmup 3 --no-cpu-baseline --no-train-step   (BASELINE configs[4])" >> $OUT
  timeout 600 python bench.py --height 960 --width 1280 --batch 16 --keypoints $K --steps 30 --warmup 3 --no-cpu-baseline --no-train-step 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
keep={k:d[k] for k in ('metric','value','unit','ms_per_step','dtype','config','e2e','gpu_launches_per_step','clocks','frac_of_bf16_peak_whole_step') if k in d}
keep['roofline']={k:d['roofline'][k] for k in ('achieved','peak','frac','ms_in_step','launches')}
print(json.dumps(keep))" >> $OUT
done
cat $OUT | cut -c1-400
