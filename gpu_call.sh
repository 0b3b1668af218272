mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_step_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2>&1; tail -2 gpurun_out/b.log | cut -c1-200; wc -l gpurun_out/r02_launches_step_final.csv
