python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/ab_kernels.py --warmup 5 --frames 10 --variants 22 > gpurun_out/ab_early.log 2>&1
