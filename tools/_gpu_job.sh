mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for i in 1 2; do
timeout 120 $TR --nproc-per-node 2 --master-port 2954$i bench.py --gpus 2 --steps 10 --warmup 3 --watchdog 90 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "run $i rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('N=2 value %.4g (%.2f ms) e2e %.4g (%.2f ms)'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))" 2>&1 | tail -1
done
