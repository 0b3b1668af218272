mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "advect" 2>&1 | tail -3
timeout 300 python bench_kernels.py 2>&1 | head -5
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_lb.json 2> gpurun_out/bench_n1_lb.err; tail -c 300 gpurun_out/bench_n1_lb.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_lb.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['ms'])
for k,v in d['roofline_advect']['kernels'].items(): print(k, v['ms'], v['frac'])
"
