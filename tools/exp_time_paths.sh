#!/bin/bash
# time-axis transforms per record length: default path, and each implementation forced (P3D_TIME_PATH)
out=gpurun_out/time_paths.txt; : > $out
for cfg in ${CFGS:-"2048 1000 1000" "2048 1000 1000 0" "512 1000 1000" "1000 1000 1000" "1024 1000 1000" "2000 1000 1000" "4096 500 500" "2500 600 600" "4000 500 500"}; do
  echo "== $cfg default" >> $out; python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  for path in ${PATHS:-tma direct pipeline}; do
    echo "== $cfg P3D_TIME_PATH=$path" >> $out; P3D_TIME_PATH=$path python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  done
  for path in tma pipeline; do
    echo "== $cfg envelope P3D_TIME_PATH=$path" >> $out; P3D_BENCH_ENVELOPE=1 P3D_TIME_PATH=$path python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  done
done
cat $out
