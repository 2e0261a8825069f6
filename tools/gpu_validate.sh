#!/bin/bash
# One GPU-box pass: parity tests, smoke, both bench arms, the ncu launch list and one --set full capture.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>&1; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
if [ -n "$NCU_FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_kernel' \
   --launch-skip 6 -c 1 -o $OUT/prof_tc_fp4_$TAG -f python tools/profile_step.py --n 20000 --u16 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT
