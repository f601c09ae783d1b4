mkdir -p gpurun_out; T=gpurun_out/r2_t40
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "head_gemm" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -25 ${T}_tests.log
