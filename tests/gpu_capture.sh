#!/bin/bash
# ncu evidence for profiles/ (one B200, run under gpurun):  bash tests/gpu_capture.sh <tag>
# launch lists (gpu__time_duration, --clock-control none) and one `--set full` capture of the step kernel for C2 and C5,
# at the launch size the default bench uses (16 steps per launch).  Numbers printed under ncu are never bench values.
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
for cfg in C2 C5; do
  cmd="python bench.py --config $cfg --steps 32 --warmup 3 --profile-only --no-c5"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_$cfg.csv $cmd > $out/${tag}_launches_$cfg.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:pgbart_step -s 3 -c 2 -f -o $out/${tag}_full_$cfg $cmd > $out/${tag}_full_$cfg.log 2>&1
  ncu -i $out/${tag}_full_$cfg.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw_$cfg.csv 2>/dev/null
  ncu -i $out/${tag}_full_$cfg.ncu-rep --page source --csv --print-source sass > $out/${tag}_ncu_source_$cfg.csv 2>/dev/null
  ls -la $out/${tag}_full_$cfg.ncu-rep
  if [ $(stat -c %s $out/${tag}_full_$cfg.ncu-rep) -gt 20000000 ]; then rm -f $out/${tag}_full_$cfg.ncu-rep; fi
done
