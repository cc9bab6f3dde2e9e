python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json
