mkdir -p gpurun_out; T=gpurun_out/r2_t39
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > ${T}_bench_n8.log 2>&1
echo "rc=$?"
tail -1 ${T}_bench_n8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['loss_check'], d['phase_ms'])" || tail -30 ${T}_bench_n8.log
