# quick kernel timing on cfg2 and cfg4 (developer tool)
for w in cfg2 cfg4; do
python bench.py --workload $w --steps 3 --warmup 3 --e2e-steps 1 --cpu-sample 50 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$w', 'value %.2fM q/s' % (d['value']/1e6), 'probe %.2f ms score %.2f ms' % (k['probe_ms'], k['score_ms']), 'e2e %.2fM' % (d['e2e']['value']/1e6))"
done
