# end-to-end rate (host pinned in / out) against the number of lanes and the chunk size; LC="lanes chunk;..." overrides
IFS=';' read -ra COMBOS <<< "${LC:-4 0;4 64;6 40;8 32;8 24;12 24;16 16}"
for lc in "${COMBOS[@]}"; do set -- $lc; python bench.py --lanes $1 --chunk $2 --steps 2 --warmup 2 --no-cpu-baseline --no-diag 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('lanes $1 chunk $2: device %.0f  e2e %.0f  ratio %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value']))"; done
