mkdir -p gpurun_out; T=gpurun_out/r2_t25
timeout 400 python -m pytest tests -x -q -m gpu -k "data_parallel" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > ${T}_bench_n8.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > ${T}_bench_n4.log 2>&1
tail -2 ${T}_tests.log
for f in ${T}_bench_n8.log ${T}_bench_n4.log ${T}_bench_n1.log; do tail -1 $f | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['loss_check'], d['phase_ms'])"; done
