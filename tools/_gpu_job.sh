mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; cat gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
