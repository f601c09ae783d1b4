mkdir -p gpurun_out; T=gpurun_out/r2_t13
timeout 1500 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${T}_bench.log 2>&1
DMF_WGRAD_T=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-kernel-rooflines > ${T}_bench_T.log 2>&1
tail -3 ${T}_tests.log
