mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pack_ops_kernel' --launch-skip 20 -c 1 -o gpurun_out/prof_pack_fp4_v7 -f python tools/profile_step.py --n 20000 --u16 > gpurun_out/ncu_pack_v7.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pack_v7.log
