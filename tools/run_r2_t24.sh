# final single-GPU evidence set of round 2: default bench line, small configs, kernel micro-bench, ncu captures
mkdir -p gpurun_out; T=gpurun_out/r2_t24
timeout 900 python bench.py > ${T}_bench_default.log 2>&1
for c in c1 c2 c3 c4; do timeout 300 python bench.py --config $c > ${T}_bench_$c.log 2>&1; done
timeout 600 python tools/kernel_bench.py --B 65536 > ${T}_kb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_tc6 -s 2 -c 1 -o ${T}_tc6 -f python tools/kernel_bench.py --what bwd --B 65536 > ${T}_ncu_tc6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc2 -s 8 -c 1 -o ${T}_gemm_k512 -f python tools/kernel_bench.py --what gemm --B 65536 > ${T}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rowcol_sum_tc4 -s 2 -c 1 -o ${T}_fwd -f python tools/kernel_bench.py --what fwdfused --B 65536 > ${T}_ncu_fwd.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-kernel-rooflines --no-loss-check --no-cpu-baseline > ${T}_ncu_list.log 2>&1
ls -la ${T}_*; tail -2 ${T}_kb.log
