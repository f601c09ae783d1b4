mkdir -p gpurun_out; T=gpurun_out/r2_t15
timeout 300 python -m pytest tests -x -q -m gpu -k "dssl_variants or data_parallel" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-diag > ${T}_bench_n8.log 2>&1
tail -3 ${T}_tests.log; tail -c 3000 ${T}_bench_n8.log
