mkdir -p gpurun_out; T=gpurun_out/r2_t51
for c in c1 c2 c3 c4; do timeout 200 python bench.py --config $c --no-cpu-baseline > ${T}_bench_$c.log 2>&1; tail -1 ${T}_bench_$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c', round(d['value']), d['unit'], round(d['ms_per_step'],3), 'ms', d.get('gpu_launches_per_step'), round(d['e2e']['value']))" || tail -5 ${T}_bench_$c.log; done
