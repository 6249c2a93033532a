mkdir -p gpurun_out
timeout 200 python tools/measure_t5.py --shapes 1000:1000000 --reps 1 --oracle-sample 0 > gpurun_out/t5_warm.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_t5_count|k_t5_write" -s 2 -c 2 -o gpurun_out/t5_full python tools/measure_t5.py --shapes 1000:1000000 --reps 1 --oracle-sample 0 > gpurun_out/t5_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_t2_count|k_t2_plan" -s 2 -c 2 -o gpurun_out/t3_full python tools/measure_t2.py --t3 --shapes 1000:1000000 --reps 1 --oracle-sample 0 > gpurun_out/t3_ncu.log 2>&1
tail -2 gpurun_out/t5_warm.log; tail -2 gpurun_out/t5_ncu.log; tail -2 gpurun_out/t3_ncu.log
