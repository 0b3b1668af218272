set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ensemble" 2>&1 | tail -15
timeout 300 python bench_kernels.py --ensemble 65536 --ens-shape 80x60 --ens-variant 5,6,7,8,9,14,12 --out gpurun_out/r02_ens_sweep.json 2>&1 | tail -20
timeout 200 python bench_kernels.py --ensemble 16384 --ens-shape 61x81 --ens-variant 5,7,8,9 --out gpurun_out/r02_ens_sweep_61x81.json 2>&1 | tail -10
timeout 300 compute-sanitizer --tool racecheck python tools_sanitize_ens.py > gpurun_out/r02_sanitizer_ens_racecheck.log 2>&1; tail -5 gpurun_out/r02_sanitizer_ens_racecheck.log
timeout 300 compute-sanitizer --tool memcheck python tools_sanitize_ens.py > gpurun_out/r02_sanitizer_ens_memcheck.log 2>&1; tail -5 gpurun_out/r02_sanitizer_ens_memcheck.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ensemble_reg -c 1 -f -o gpurun_out/ens_reg_r4 python bench_kernels.py --ensemble 8192 --ens-variant 7 2>&1 | tail -3
