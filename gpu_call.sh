mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu -k "decomposed_step" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_fold.json 2> gpurun_out/bench_n2_fold.err; tail -c 400 gpurun_out/bench_n2_fold.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n2_fold.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('parity'), d.get('phases_last_step'))
print(d['config'])
"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_same_box_as_n2_fold.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_n1_same_box_as_n2_fold.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
"
