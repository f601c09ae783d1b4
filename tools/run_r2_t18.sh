mkdir -p gpurun_out; T=gpurun_out/r2_t18
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > ${T}_bench_n8_high.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --nccl-priority normal > ${T}_bench_n8_normal.log 2>&1
for f in ${T}_bench_n8_high.log ${T}_bench_n8_normal.log; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['phase_ms'])"; done
