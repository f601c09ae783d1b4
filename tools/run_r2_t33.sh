mkdir -p gpurun_out; T=gpurun_out/r2_t33
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rowcol_sum_tc4 -s 3 -c 1 -o ${T}_fwd_e -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:infonce_bwd_e -s 1 -c 2 -o ${T}_bwd_e -f python tools/kernel_bench.py --what stored --B 65536 --iters 1 > ${T}_ncu_bwd.log 2>&1
ls -la ${T}_*; tail -3 ${T}_ncu_fwd.log
