mkdir -p gpurun_out; T=gpurun_out/r2_t16
nvidia-smi topo -m > ${T}_topo.log 2>&1; lscpu | grep -i "numa\|socket\|model name" >> ${T}_topo.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --e2e-diag --no-cpu-baseline > ${T}_bench_n8.log 2>&1
tail -c 2500 ${T}_bench_n8.log; cat ${T}_topo.log | head -30
