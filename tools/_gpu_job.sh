mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_1gpu.json"))
print("value %.4g ms %.3f e2e %.4g (%.2f ms; in-order %.2f) frac %.3f pack %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["in_order_ms_per_step"], d["roofline"]["frac"], {k:d["roofline_pack"][k] for k in ("achieved","frac","ms")}))
PY
