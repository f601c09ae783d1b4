mkdir -p gpurun_out; T=gpurun_out/r2_t17
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-diag --no-cpu-baseline > ${T}_bench_n8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-diag --no-cpu-baseline --h2d-at-step-start > ${T}_bench_n8_start.log 2>&1
for f in ${T}_bench_n8.log ${T}_bench_n8_start.log; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_diag'], d['phase_ms'])"; done
