mkdir -p gpurun_out
tools/ubench_pcie 512 > gpurun_out/ubench_pcie.json 2>&1; cat gpurun_out/ubench_pcie.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
DG_TRACE=1 timeout 300 python tools/e2e_trace.py > gpurun_out/e2e_trace.log 2>&1; echo "trace rc=$?"
DG_TRACE=1 timeout 300 python tools/e2e_trace.py --panels 12 --chunk-mb 32 > gpurun_out/e2e_trace_p12.log 2>&1; echo "trace rc=$?"
timeout 600 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/e2e_sweep.log
