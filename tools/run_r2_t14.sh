mkdir -p gpurun_out; T=gpurun_out/r2_t14
timeout 1500 python -m pytest tests -q -m gpu > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 3 --graph off --no-kernel-rooflines --no-loss-check > ${T}_ncu.log 2>&1
tail -3 ${T}_tests.log
