out=gpurun_out/time_stride.txt; : > $out
for cfg in "2048 128 128" "2048 256 256" "2048 512 512" "2048 1000 1000" "2048 1024 1024"; do
  for v in 0 2; do
  echo "== $cfg direct variant=$v" >> $out; P3D_TIME_DIRECT=1 P3D_TIME_VARIANT=$v python tools/bench_time_axis.py $cfg 2>&1 | grep -v "round trip" >> $out
  done
done
cat $out
