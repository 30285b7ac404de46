for c in 1 2 3 4 6; do
  PT_TC_CHUNKS=$c python bench.py --instances 12500 --steps 40 --warmup 5 --no-lca --no-weak --no-config4 --no-config5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('chunks', $c, 'resident ms', round(d['ms_per_step'], 4), 'e2e M/s', round(d['e2e']['value'] / 1e6, 2), 'e2e ms', round(12500 / d['e2e']['value'] * 1e3, 4))
"
done
