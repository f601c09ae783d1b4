mkdir -p gpurun_out; T=gpurun_out/r2_t23
timeout 1500 python -m pytest tests -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench.log 2>&1
tail -3 ${T}_tests.log
for f in ${T}_bench.log; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['phase_ms'], d['gpu_launches_per_step'])"; done
