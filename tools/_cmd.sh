mkdir -p gpurun_out
bash tools/gpu_kickoff.sh r2m tests
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2m_c3.json 2> gpurun_out/r2m_c3.err || { echo "c3 FAILED"; tail -c 2500 gpurun_out/r2m_c3.err; }
python profiles/bench_summary.py gpurun_out/r2m_c3.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m_c3.json"))
print("value", d["value"], "e2e", d["e2e"], "\nsteady", d.get("steady_state_window_drain"))
PY
