mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipelined.py -x -q -k "repack or sars" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pipe.log
for pb in 134217728 67108864 33554432; do
timeout 600 python bench.py --steps 20 --no-cpu-baseline --panel-bytes $pb > gpurun_out/bench_pb$pb.json 2> gpurun_out/bench_pb$pb.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_pb$pb.json"))
print($pb, "value %.4g ms %.3f e2e %.4g (%.2f ms) frac %.3f pack_ms %.3f serial %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline_pack"]["ms"], d["roofline_pack"]["note"][:60]))
PY
done
timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-repack-overlap > gpurun_out/bench_noov.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_noov.json')); print('no overlap: value %.4g ms %.3f'%(d['value'], d['ms_per_step']))"
