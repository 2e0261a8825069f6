mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipelined.py -x -q > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pipe.log
timeout 900 python tools/run_configs.py --only 1,2,3,5 --cpu-budget 4 > gpurun_out/configs_v2.jsonl 2> gpurun_out/configs_v2.err; echo "configs rc=$?"; tail -3 gpurun_out/configs_v2.err
python - <<PY
import json
for l in open("gpurun_out/configs_v2.jsonl"):
    d=json.loads(l)
    if "kernel_only" in d:
        print(d["config"][:60], "| kernel %.2f ms (fp4 frac %.2f) | e2e %.1f ms | pipelined %.1f ms | cpu %.3g" % (d["kernel_only"]["ms"], d["kernel_only"]["frac_of_measured_fp4_peak"], d["e2e"]["ms"], d["e2e_pipelined"]["ms"], d["cpu_baseline"]["value"]))
    else:
        print(d)
PY
