#!/bin/bash
# variants of the direct (one-pass) time-axis kernels against the transposing pipeline
out=gpurun_out/time_variants.txt; : > $out
cfg="${CFG:-2048 1000 1000}"
echo "== $cfg pipeline" >> $out; python tools/bench_time_axis.py $cfg >> $out 2>&1
for v in ${VARS:-0 1 2 4 5 6}; do
  echo "== $cfg direct variant=$v" >> $out; P3D_TIME_DIRECT=1 P3D_TIME_VARIANT=$v python tools/bench_time_axis.py $cfg >> $out 2>&1
done
cat $out
