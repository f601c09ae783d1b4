mkdir -p gpurun_out; T=gpurun_out/r2_t54
timeout 600 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -3 ${T}_tests.log
