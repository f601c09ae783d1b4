mkdir -p gpurun_out; T=gpurun_out/r2_t47
timeout 900 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -4 ${T}_tests.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['phase_ms'])"
