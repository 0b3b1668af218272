set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ensemble" 2>&1 | tail -5
timeout 300 python bench_kernels.py --ensemble 65536 --ens-shape 80x60 --ens-variant 0,7,12 --out gpurun_out/r02_ens_sweep3.json 2>&1 | tail -20
timeout 200 python bench_kernels.py --ensemble 16384 --ens-shape 61x81 --ens-variant 0,7 --out gpurun_out/r02_ens_sweep3_61x81.json 2>&1 | tail -10
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ensemble_reg --launch-skip 4 -c 1 -f -o gpurun_out/ens_reg_r2_16steps_b python bench_kernels.py --ensemble 8192 --ens-variant 0 2>&1 | tail -3
