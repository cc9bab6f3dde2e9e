python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/ab_kernels.py --frames 6 --variants 22,11 > gpurun_out/ab_prod.log 2>&1
grep -h K6_ms gpurun_out/ab_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'], d['variant'], 'K4', d['span_ms']['K4'], 'K6/sweep', d['K6_ms_per_sweep'], 'prep', d['span_ms']['k_build_slots+prep'], 'sum', d['sum_ms'], d['state_hash'], d['rows_hash'])
"
tail -1 gpurun_out/ab_prod.log
