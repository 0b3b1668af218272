mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ensemble" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool racecheck python tools_sanitize_ens.py > gpurun_out/r02_sanitizer_ens_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_ens_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python tools_sanitize_ens.py > gpurun_out/r02_sanitizer_ens_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_ens_memcheck.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ensemble_reg -c 1 -f -o gpurun_out/ens_reg_pipe_1step python bench_kernels.py --ensemble 8192 --ens-variant 0 --ens-steps 1 2>&1 | tail -2
