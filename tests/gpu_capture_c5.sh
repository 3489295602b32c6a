#!/bin/bash
# ncu evidence for the C5 workload at 8 chains per GPU (one B200, under gpurun):  bash tests/gpu_capture_c5.sh <tag>
# (1) targeted counters of two 16-step launches (few replay passes), (2) one `--set full` capture of one launch,
# (3) the launch list.  Numbers printed under ncu are never bench values.
tag=${1:-r2h}
out=gpurun_out
mkdir -p $out
cmd="python bench.py --config C5 --steps 32 --warmup 3 --profile-only --clock-ms 0"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
timeout 170 ncu --metrics $M --clock-control none -k regex:pgbart_step -s 3 -c 2 --csv --log-file $out/${tag}_counters_C5.csv $cmd > $out/${tag}_counters_C5.log 2>&1
echo "counters rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_C5.csv $cmd > $out/${tag}_launches_C5.log 2>&1
echo "launches rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:pgbart_step -s 3 -c 1 -f -o $out/${tag}_full_C5 $cmd > $out/${tag}_full_C5.log 2>&1
echo "full rc=$?"
if [ -f $out/${tag}_full_C5.ncu-rep ]; then
  ncu -i $out/${tag}_full_C5.ncu-rep --page raw --csv > $out/${tag}_ncu_full_raw_C5.csv 2>/dev/null
  ncu -i $out/${tag}_full_C5.ncu-rep --page source --csv --print-source sass > $out/${tag}_ncu_source_C5.csv 2>/dev/null
  ls -la $out/${tag}_full_C5.ncu-rep
  if [ $(stat -c %s $out/${tag}_full_C5.ncu-rep) -gt 20000000 ]; then rm -f $out/${tag}_full_C5.ncu-rep; fi
fi
tail -3 $out/${tag}_counters_C5.csv | cut -c1-400
