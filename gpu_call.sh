mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n${N}_side.json 2> gpurun_out/bench_n${N}_side.err; tail -c 300 gpurun_out/bench_n${N}_side.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n${N}_side.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('parity'))
print(d.get('phases_last_step',{}).get('max_over_ranks_ms'))
for k,v in d.get('extra',{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','efficiency_vs_single_gpu_same_rectangle','single_gpu_same_rectangle_ms','error')} if isinstance(v,dict) else v)
"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box_as_n${N}_side.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_same_box_as_n${N}_side.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
"
