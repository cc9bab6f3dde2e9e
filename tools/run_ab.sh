python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/ab_kernels.py --warmup 5 --frames 10 --variants 22 > gpurun_out/ab_early.log 2>&1
python tools/ab_kernels.py --warmup 300 --frames 4 --variants 22 > gpurun_out/ab_steady.log 2>&1
python tools/ab_kernels.py --workload config3 --warmup 300 --frames 4 --variants 22 > gpurun_out/ab_steady3.log 2>&1
python tools/ab_kernels.py --workload config3 --warmup 60 --frames 50 --variants 22 > gpurun_out/ab_mid3.log 2>&1
