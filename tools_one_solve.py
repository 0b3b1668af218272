import sys, os
sys.path.insert(0, os.getcwd())
import torch
import esp32_fluid_simulation_b200 as fb
stream = torch.cuda.Stream(); ctx = fb.Context(0, stream)
n = 4096
with torch.cuda.stream(stream):
    d = torch.randn(n, n, device="cuda") * 10; p = torch.empty(n, n, device="cuda")
stream.synchronize()
for _ in range(2):
    ctx.poisson_solve(p, d, n, n, 1.0, 50, 1.96)
stream.synchronize()
