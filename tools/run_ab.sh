python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -m gpu -x -q 2>&1 | tail -4
python tools/ab_kernels.py --warmup 5 --frames 10 --variants 22 > gpurun_out/ab_early.log 2>&1
python tools/profile_step.py --workload config4 --frames 6 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_build_slots" -s 12 -c 3 -o gpurun_out/r2_k6 python tools/profile_step.py --workload config4 --frames 6 > gpurun_out/ncu_k6.log 2>&1
