mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "poisson or step or sim or harness" 2>&1 | tail -3
timeout 300 python tools_sor_pass_cost.py 2>&1 | tail -3
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sor_blocked_tma --csv --log-file gpurun_out/sor_solve_dram.csv python tools_one_solve.py > /dev/null 2>&1; tail -5 gpurun_out/sor_solve_dram.csv
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_sched.json 2> gpurun_out/bench_n1_sched.err; tail -c 300 gpurun_out/bench_n1_sched.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_sched.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'], d['roofline']['launches_per_solve'], d['gpu_launches'])
print(d['extra']['graph_step']['ms_per_step'], d['extra']['ensemble']['results'])
"
