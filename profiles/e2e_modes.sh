# resident / double-buffered / serial end-to-end rates at the full batch and at one rank's shard of the N = 8 run
for n in 100000 12500; do
  python bench.py --instances $n --steps 40 --warmup 5 --no-lca --no-config4 --no-config5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
n = d['config']['instances']
print('instances', n, 'resident ms %.4f' % d['ms_per_step'], 'e2e double-buffered ms %.4f' % (n / d['e2e']['value'] * 1e3), 'serial ms %.4f' % (n / d['e2e']['serial_value'] * 1e3), 'displayed', d['e2e']['displayed'])"
done
