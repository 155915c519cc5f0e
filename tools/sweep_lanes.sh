for l in 1 2 3 4 6 8; do python bench.py --lanes $l --steps 2 --warmup 2 --no-cpu-baseline --no-diag 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('lanes $l: device %.0f  e2e %.0f  ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value']))"; done
