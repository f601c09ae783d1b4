mkdir -p gpurun_out; T=gpurun_out/r2_t20
timeout 1500 python -m pytest tests -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > ${T}_bench.log 2>&1
tail -3 ${T}_tests.log
tail -1 ${T}_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_ms'], d['phase_ms']); print(d['roofline_k1'])"
