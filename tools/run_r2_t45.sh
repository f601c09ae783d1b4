# final single-GPU evidence of round 2: smoke, full GPU test-suite, default bench line, steady-state kernel bench, ncu capture of the forward
mkdir -p gpurun_out; T=gpurun_out/r2_t45
timeout 300 python __graft_entry__.py smoke > ${T}_smoke.log 2>&1; echo "smoke rc=$?" >> ${T}_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 900 python bench.py > ${T}_bench_default.log 2>&1
timeout 300 python tools/kernel_bench.py --what stored --B 65536 --iters 60 > ${T}_kb_stored.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rowcol_sum_tc4 -s 3 -c 1 -o ${T}_fwd_e -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_fwd.log 2>&1
tail -4 ${T}_smoke.log; tail -3 ${T}_tests.log; tail -4 ${T}_kb_stored.log; tail -1 ${T}_bench_default.log | cut -c1-400
