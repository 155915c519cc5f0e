# quick device-resident numbers of every BASELINE config in every precision mode (development helper)
for cfg in 1 3 4 5; do for prec in auto 32 64; do
  python bench.py --config $cfg --precision $prec --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-diag 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('C$cfg $prec: %.0f slice-it/s  f64 share %.2f  shares %s' % (d['value'], r['complex128_share_of_slice_iterations'], {k: round(v, 3) for k, v in r['kernel_share_of_step'].items()}))
print('   ', d['config']['plan'].split('cols_iter=')[1][:260])"
done; done
