# the 8-GPU records of the round: the full bench line, then a longer run of the headline alone (40 steps) with both exchange paths
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
run 29530 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?; python -c "
import json
d = json.loads(open('gpurun_out/bench_n8.json').readline())
print('N=8 value %.2f M/s, ms %.4f, kernel %.4f, e2e %.2f (serial %.2f), weak %.2f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'] / 1e6, d['e2e']['serial_value'] / 1e6, d['weak_scaling']['value'] / 1e6))
print(d['clocks'], d['gpu_launches'], d['counters']['work_items'], d['config4']['ms'], d['config5']['seconds'], d['lca']['queries_per_s'])"
for p in 1 0; do
  PT_COMM_P2P=$p run 2951$p --steps 40 --warmup 5 --no-lca --no-config4 --no-config5 --no-weak 2>gpurun_out/n8_ab_$p.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('p2p', $p, 'value %.2f M/s' % (d['value'] / 1e6), 'ms %.4f' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms'], 'e2e %.2f' % (d['e2e']['value'] / 1e6), 'serial %.2f' % (d['e2e']['serial_value'] / 1e6))"
done
