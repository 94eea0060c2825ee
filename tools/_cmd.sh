timeout 1200 python -m pytest tests -x -q -m gpu --timeout 900 --timeout-method thread > gpurun_out/r3g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3g_tests.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r3g_ref.json 2> gpurun_out/r3g_ref.err; tail -c 400 gpurun_out/r3g_ref.json
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r3g_c3.json 2> gpurun_out/r3g_c3.err || tail -5 gpurun_out/r3g_c3.err
python profiles/bench_summary.py gpurun_out/r3g_c3.json 2>/dev/null | head -1 | cut -c1-200
timeout 300 python bench.py --workload c3_groups --steps 50 --warmup 5 --no-cpu-baseline --no-collider-rows --no-config4 --no-steady-state > gpurun_out/r3g_groups.json 2> gpurun_out/r3g_groups.err; python profiles/bench_summary.py gpurun_out/r3g_groups.json 2>/dev/null | head -1 | cut -c1-200
NOX="--no-cpu-baseline --no-collider-rows --no-parity --no-config4 --no-steady-state --repeats 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 400 --csv --log-file gpurun_out/launches_r3g.csv python bench.py --steps 5 --warmup 5 $NOX > gpurun_out/r3g_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
