# the 8-GPU record of the round: the full bench line (add PT_COMM_P2P=0 in the environment for the NCCL exchange)
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
run 29530 --steps 40 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?; python -c "
import json
d = json.loads(open('gpurun_out/bench_n8.json').readline())
print('N=8 value %.2f M/s, ms %.4f, kernel %.4f, e2e %.2f (serial %.2f), weak %.2f' % (d['value'] / 1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'] / 1e6, d['e2e']['serial_value'] / 1e6, d['weak_scaling']['value'] / 1e6))
print([round(x[0], 4) for x in d['roofline']['kernel_ms_per_rank_avg_max']], [round(x[1], 4) for x in d['roofline']['kernel_ms_per_rank_avg_max']])
print(d['clocks'], d['gpu_launches'], d['counters']['work_items'], d['config4']['ms'], d['config5']['seconds'], d['lca']['queries_per_s'])"
