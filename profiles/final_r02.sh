cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/final_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:dp_who_flow_kernel" -s 2 -c 2 -f -o gpurun_out/r02_who_flow python profiles/r02_driver.py who > gpurun_out/cap_who_flow.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -2
