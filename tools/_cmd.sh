for g in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2952$g bench.py --gpus $g --steps 50 --warmup 5 > gpurun_out/r3h_g$g.json 2> gpurun_out/r3h_g$g.err || tail -5 gpurun_out/r3h_g$g.err
python profiles/bench_summary.py gpurun_out/r3h_g$g.json 2>/dev/null | head -1 | cut -c1-330
done
