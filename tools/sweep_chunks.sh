for lc in "4 0" "4 96" "4 64" "4 48" "6 64" "8 48" "8 32"; do set -- $lc; python bench.py --lanes $1 --chunk $2 --steps 2 --warmup 2 --no-cpu-baseline --no-diag 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('lanes $1 chunk $2: device %.0f  e2e %.0f  ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value']))"; done
