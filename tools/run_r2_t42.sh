mkdir -p gpurun_out; T=gpurun_out/r2_t42
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_head" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -25 ${T}_tests.log
