mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 900 python -m pytest tests/test_sharded_multiprocess_gpu.py -x -q -m gpu --timeout 600 --timeout-method thread > gpurun_out/r2n_mp_tests.log 2>&1; echo "mp tests rc=$?"; tail -5 gpurun_out/r2n_mp_tests.log
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2n_g$N.json 2> gpurun_out/r2n_g$N.err || { echo "bench g$N FAILED"; tail -c 2500 gpurun_out/r2n_g$N.err; }
python profiles/bench_summary.py gpurun_out/r2n_g$N.json
done
