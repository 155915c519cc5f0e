#!/bin/bash
# radix-2 split one-pass kernels for the long records (default) with the tables in shared memory / in global memory
out=gpurun_out/time_split.txt; : > $out
python -m pytest tests/test_gpu_parity.py tests/test_envelope.py -q -k "time_axis or envelope" 2>&1 | tail -4 >> $out
for cfg in ${CFGS:-"2048 1000 1000" "2048 1000 1000 0" "4096 500 500" "2000 1000 1000" "4000 500 500"}; do
  echo "== $cfg default" >> $out; python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  echo "== $cfg P3D_TIME_NO_SMEM_TABLES=1" >> $out; P3D_TIME_NO_SMEM_TABLES=1 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
done
cat $out
