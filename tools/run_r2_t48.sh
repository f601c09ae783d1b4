# same-box A/B: stored-probability backward (default) against the recompute backward (DMF_STORE_E=0)
mkdir -p gpurun_out; T=gpurun_out/r2_t48
for v in 1 0 1 0; do
DMF_STORE_E=$v timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_store$v.log 2>&1
tail -1 ${T}_bench_store$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('DMF_STORE_E=$v', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['roofline']['kernel'], round(d['roofline']['frac'],3), {k: round(v,2) for k,v in d['phase_ms'].items() if k in ('rowlse_x4','infonce_bwd','graph_total')}, d['loss_check']['loss'], d['loss_check']['grad_norm'])" | tee -a ${T}_summary.log
done
