# the round's closing run on one B200: GPU tests, smoke, the default bench line, a small-shard bench (19 rotating copies, the N = 8 shard size),
# and the launch list of a short bench under ncu (metrics-only pass)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/final_tests.log | head -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
python bench.py --instances 12512 --no-lca --no-config4 --no-config5 --no-cpu-baseline > gpurun_out/final_bench_shard.json 2> gpurun_out/final_bench_shard.err; echo "shard bench rc=$?"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/launch_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/launches_r02.csv
