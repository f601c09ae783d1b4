mkdir -p gpurun_out; T=gpurun_out/r2_t52
timeout 600 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -3 ${T}_tests.log
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline_k2_bwd']['what'][:80])"
