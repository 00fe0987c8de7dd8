#!/bin/bash
# Round-2 ncu evidence for the inference path (GPU box).  Each ncu command runs only after the identical plain command exited 0.
# Launch offsets are derived from the plain run's own launches-per-step count (warmup 3 -> skip 3 steps).  Reports larger than
# 20 MB are dumped to their raw-page CSV on the box and dropped (gpurun brings back at most 64 MiB).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --profile-mode"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
LPS=$(python - <<'PY'
import json
for l in open("gpurun_out/plain.log"):
    if l.startswith("{"):
        print(json.loads(l)["gpu_launches_per_step"])
PY
)
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
KR='regex:conv_tc|stem_tc|stem_pool|head_|maxpool|argmax'
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k "$KR" -s $((3 * LPS)) -c $LPS --csv --log-file gpurun_out/step_dram.csv $CMD > gpurun_out/ncu_dram.log 2>&1
echo "per-launch dram rc=$?"
python tools/step_dram_to_json.py gpurun_out/step_dram.csv gpurun_out/step_per_launch_dram.json
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none -k "$KR" -s $((3 * LPS)) -c $LPS -o gpurun_out/prof_step_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full step capture rc=$?"
ncu -i gpurun_out/prof_step_full.ncu-rep --page raw --csv > gpurun_out/prof_step_full_raw.csv 2> gpurun_out/ncu_export.log; echo "raw export rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:stem_pool|head_|argmax' -s 12 -c 4 -o gpurun_out/prof_small_src $CMD > gpurun_out/ncu_small.log 2>&1
echo "stem_pool/head/decode capture with source rc=$?"
for f in gpurun_out/*.ncu-rep; do
  sz=$(stat -c %s "$f"); if [ "$sz" -gt 20000000 ]; then echo "dropping $f ($sz bytes; raw CSV kept)"; rm -f "$f"; fi
done
du -sh gpurun_out
