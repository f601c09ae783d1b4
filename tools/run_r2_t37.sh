mkdir -p gpurun_out; T=gpurun_out/r2_t37
timeout 500 python -m pytest tests -x -q -m gpu -k "data_parallel" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -8 ${T}_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > ${T}_bench_n2.log 2>&1
tail -1 ${T}_bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['loss_check'], d['phase_ms'])"
