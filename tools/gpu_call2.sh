#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_tests.log 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
timeout 200 python profiles/e2e_breakdown.py > gpurun_out/c2_e2e_breakdown.txt 2>&1
tail -5 gpurun_out/c2_tests.log
cat gpurun_out/c2_e2e_breakdown.txt
