mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; cat gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or two_gpu or device_memory" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2gpu.log
