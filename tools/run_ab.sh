python tools/ab_kernels.py --frames 6 --variants 22 > gpurun_out/ab_prod.log 2>&1
for t in b5 b6 b8; do python tools/ab_kernels.py --frames 6 --variants 22 --lib $t > gpurun_out/ab_$t.log 2>&1; done
grep -h K6_ms gpurun_out/ab_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['lib'], 'build+prep', d['span_ms']['k_build_slots+prep'], 'sum', d['sum_ms'], d['state_hash'], d['rows_hash'])
"
