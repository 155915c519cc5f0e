out=gpurun_out/time_env.txt; : > $out
python -m pytest tests/test_envelope.py -q 2>&1 | tail -4 >> $out
for cfg in "2048 1000 1000" "4096 500 500" "2000 1000 1000" "4000 500 500"; do
  echo "== $cfg envelope default" >> $out; P3D_BENCH_ENVELOPE=1 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  echo "== $cfg envelope P3D_TIME_SPLIT=0 P3D_TIME_PATH=tma" >> $out; P3D_TIME_SPLIT=0 P3D_TIME_PATH=tma P3D_BENCH_ENVELOPE=1 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  echo "== $cfg envelope P3D_TIME_PATH=pipeline" >> $out; P3D_TIME_PATH=pipeline P3D_BENCH_ENVELOPE=1 python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
done
cat $out
