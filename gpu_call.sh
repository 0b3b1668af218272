mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -x -q -m gpu -k "divergence or step or dist or native or tile" 2>&1 | tail -3
timeout 300 python bench_kernels.py --out gpurun_out/kernels_div_rolling.json 2>&1 | head -3
