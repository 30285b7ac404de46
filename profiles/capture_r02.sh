set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:tc_persistent -s 3 -c 1 -f -o gpurun_out/r02_tc python profiles/r02_driver.py tc > gpurun_out/cap_tc.log 2>&1
$NCU -k "regex:lca_query_kernel|lca_adjacent_kernel|lca_sweeps_kernel" -s 2 -c 4 -f -o gpurun_out/r02_lca python profiles/r02_driver.py lca > gpurun_out/cap_lca.log 2>&1
$NCU -k "regex:dp_who_flow_kernel" -s 2 -c 2 -f -o gpurun_out/r02_who python profiles/r02_driver.py who > gpurun_out/cap_who.log 2>&1
$NCU -k "regex:enum_kernel" -s 1 -c 1 -f -o gpurun_out/r02_enum python profiles/r02_driver.py enum > gpurun_out/cap_enum.log 2>&1
$NCU -k "regex:iso_kernel|nw_parse_kernel" -s 2 -c 2 -f -o gpurun_out/r02_iso python profiles/r02_driver.py iso > gpurun_out/cap_iso.log 2>&1
$NCU -k "regex:br_cover_kernel|br_mark_kernel|lca_sweeps_kernel" -s 3 -c 3 -f -o gpurun_out/r02_bridges python profiles/r02_driver.py bridges > gpurun_out/cap_br.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config4 > gpurun_out/launch_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
