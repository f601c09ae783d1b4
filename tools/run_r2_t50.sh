mkdir -p gpurun_out; T=gpurun_out/r2_t50
timeout 300 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke rc=$?" >> ${T}_smoke.log; tail -2 ${T}_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -3 ${T}_tests.log
