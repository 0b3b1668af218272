mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -c 300 gpurun_out/bench_n1_final.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['cpu_baseline'], d['clocks'])
for k,v in d['roofline_advect']['kernels'].items(): print(k, v['ms'], v['frac'])
print(d['extra']['graph_step']['ms_per_step'], d['extra']['ensemble']['results'], d['extra']['e2e_frames_only']['ms_per_step'])
"
