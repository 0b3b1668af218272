mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -m gpu -k "advect or step or dist or native or sim or ensemble" 2>&1 | tail -4
timeout 200 python bench_kernels.py --ensemble 16384 --ens-shape 61x81 --ens-variant 0,7 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_advtrim.json 2> gpurun_out/bench_n1_advtrim.err; tail -c 300 gpurun_out/bench_n1_advtrim.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_advtrim.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'])
for k,v in d['roofline_advect']['kernels'].items(): print(k, v['ms'], v['frac'])
print(d['extra']['graph_step']['ms_per_step'])
"
