import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import esp32_fluid_simulation_b200 as fb
from esp32_fluid_simulation_b200 import synth
from bench_kernels import timeit
stream = torch.cuda.Stream(); ctx = fb.Context(0, stream)
n = 4096
with torch.cuda.stream(stream):
    d = torch.randn(n, n, device="cuda"); p = torch.empty(n, n, device="cuda")
stream.synchronize()
out = {}
for T in (6, 7, 8):
    ctx.set_option("sor_t", T)
    for iters in range(1, T + 1):
        ms = timeit(stream, lambda: ctx.poisson_solve(p, d, n, n, 1.0, iters, 1.96), reps=20, warm=3)
        out[f"T{T}_iters{iters}"] = round(ms, 5)
        print(T, iters, round(ms, 5), flush=True)
ctx.set_option("sor_t", 6)
for iters in (12, 50):
    ms = timeit(stream, lambda: ctx.poisson_solve(p, d, n, n, 1.0, iters, 1.96), reps=10, warm=3)
    print("T6", iters, round(ms, 5))
    out[f"T6_iters{iters}"] = round(ms, 5)
json.dump(out, open("gpurun_out/sor_pass_cost.json", "w"), indent=1)
