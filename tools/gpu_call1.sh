#!/bin/bash
# one gpurun call: bulk parity tests, bench A/B of the record loads, e2e breakdown
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_bulk_gpu.py -m gpu -x -q ) > gpurun_out/c1_tests.log 2>&1
CB200_LD256=1 timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/c1_bench_ld256.json 2> gpurun_out/c1_bench_ld256.err
CB200_LD256=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench_ld128.json 2> gpurun_out/c1_bench_ld128.err
timeout 200 python profiles/e2e_breakdown.py > gpurun_out/c1_e2e_breakdown.txt 2>&1
tail -3 gpurun_out/c1_tests.log
cat gpurun_out/c1_e2e_breakdown.txt
