mkdir -p gpurun_out
( time timeout 220 compute-sanitizer --tool memcheck python tools_sanitize.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1 ) 2>&1 | tail -3; tail -4 gpurun_out/r02_sanitizer_memcheck.log
( time timeout 220 compute-sanitizer --tool racecheck python tools_sanitize.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1 ) 2>&1 | tail -3; tail -4 gpurun_out/r02_sanitizer_racecheck.log
