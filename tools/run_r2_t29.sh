mkdir -p gpurun_out; T=gpurun_out/r2_t29
DMF_FORK=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1_nofork.log 2>&1
tail -1 ${T}_bench_n1_nofork.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['phase_ms'], d['roofline']['avg_ms'], d['clocks'])"
