set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_now.json 2> gpurun_out/bench_n1_now.err; tail -c 600 gpurun_out/bench_n1_now.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_now.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'], d['extra']['ensemble'], d['cpu_baseline'])
"
