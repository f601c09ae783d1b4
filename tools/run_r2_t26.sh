mkdir -p gpurun_out; T=gpurun_out/r2_t26
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stored or rowcol_sums_kernel" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -15 ${T}_tests.log
timeout 200 python tools/kernel_bench.py --what stored --B 65536 > ${T}_kb.log 2>&1; cat ${T}_kb.log | tail -8
