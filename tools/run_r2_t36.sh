# single-GPU evidence set after the stored-probability backward: full GPU test-suite, default bench line, kernel micro-bench
# at steady-state clocks, ncu --set full captures of the two K2 kernels, launch list of two eager steps
mkdir -p gpurun_out; T=gpurun_out/r2_t36
timeout 900 python -m pytest tests -x -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 900 python bench.py > ${T}_bench_default.log 2>&1
timeout 300 python tools/kernel_bench.py --what stored --B 65536 --iters 60 > ${T}_kb_stored.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rowcol_sum_tc4 -s 3 -c 1 -o ${T}_fwd_e -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e -s 1 -c 1 -o ${T}_bwd_e0 -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_bwd0.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e -s 3 -c 1 -o ${T}_bwd_e1 -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_bwd1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-kernel-rooflines --no-loss-check --no-cpu-baseline > ${T}_ncu_list.log 2>&1
ls -la ${T}_*; tail -3 ${T}_tests.log; tail -4 ${T}_kb_stored.log; tail -1 ${T}_bench_default.log | cut -c1-600
