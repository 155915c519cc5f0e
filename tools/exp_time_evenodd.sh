out=gpurun_out/time_evenodd.txt; : > $out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_envelope.py -q -x -k "time_axis or envelope" 2>&1 | tail -4 >> $out
for cfg in "2048 1000 1000" "2048 1000 1000 0" "4096 500 500" "2000 1000 1000" "4000 500 500"; do
  echo "== $cfg default" >> $out; timeout 120 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  echo "== $cfg envelope" >> $out; P3D_BENCH_ENVELOPE=1 timeout 120 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
done
cat $out
