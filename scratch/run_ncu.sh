mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffae_tc -s 3 -c 1 -o gpurun_out/prof_tc_final python bench.py --steps 2 --warmup 3 --e2e-steps 1 --machines 300 > gpurun_out/ncu_full_final.log 2>&1; tail -1 gpurun_out/ncu_full_final.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ffae|minmax|roll|anomaly" -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/ncu_launch_final.log 2>&1; tail -2 gpurun_out/launches_final.csv | cut -c1-200
