"""Parity tests proper: the CUDA path, called through the C ABI (ctypes binding of
include/fluid_b200.h), against the oracle on the same seeded inputs — BIT-EXACT
for every field, floats included (tolerance 0 ulp; north_star allows 1e-5
relative per step, the build is stricter)."""
import numpy as np
import pytest
import torch

from conftest import assert_bit_equal
from gpu_util import rand_drags, rand_fields, to_dev, to_host

pytestmark = pytest.mark.gpu

DT = np.float32(1 / 30.0)
SHAPES = [(2, 2), (3, 2), (2, 5), (7, 3), (5, 4), (33, 17), (61, 81), (80, 60), (257, 129),
          (130, 70), (512, 300), (1024, 512)]
SOR_VARIANTS = [0, 1]


@pytest.fixture(params=SOR_VARIANTS, ids=["sor=half-sweeps", "sor=blocked"])
def sor_variant(request, ctx):
    ctx.set_option("sor", request.param)
    yield request.param
    ctx.set_option("sor", 1)


@pytest.fixture(params=[0, 1], ids=["advect=gather", "advect=tma"])
def advect_variant(request, ctx):
    ctx.set_option("advect", request.param)
    yield request.param
    ctx.set_option("advect", 1)


# ---- per-operator parity -------------------------------------------------------------

@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("no_slip", [1, 0])
@pytest.mark.parametrize("vmax", [45.0, 400.0])
def test_advect_velocity(ctx, oracle, advect_variant, shape, no_slip, vmax):
    dim_x, dim_y = shape
    v, _ = rand_fields(1, dim_x, dim_y, vmax)
    dv, out = to_dev(v), torch.empty_like(to_dev(v))
    ctx.advect(out, dv, dv, dim_x, dim_y, DT, no_slip)       # p aliases vel, as at ino:253
    assert_bit_equal(to_host(out), oracle.advect_vec2f(v, v, DT, no_slip), "advect v")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("no_slip", [0, 1])
@pytest.mark.parametrize("vmax", [45.0, 400.0])
def test_advect_dye(ctx, oracle, advect_variant, shape, no_slip, vmax):
    dim_x, dim_y = shape
    v, c = rand_fields(2, dim_x, dim_y, vmax)
    dv, dc = to_dev(v), to_dev(c)
    out = torch.empty_like(dc)
    ctx.advect(out, dc, dv, dim_x, dim_y, DT, no_slip)
    assert_bit_equal(to_host(out, np.uint32), oracle.advect_rgb_uq32(c, v, DT, no_slip), "advect dye")


def test_advect_dye_saturation(ctx, oracle, advect_variant):
    dim_x, dim_y = 96, 64
    v, _ = rand_fields(3, dim_x, dim_y, 80.0)
    c = np.full((dim_y, dim_x, 3), 0xFFFFFFFF, np.uint32)
    c[::3, ::2] = 0xFFFFFF80
    dc, dv = to_dev(c), to_dev(v)
    out = torch.empty_like(dc)
    ctx.advect(out, dc, dv, dim_x, dim_y, DT, 0)
    got = to_host(out, np.uint32)
    assert_bit_equal(got, oracle.advect_rgb_uq32(c, v, DT, 0), "saturating dye")
    assert got.max() == 0xFFFFFFFF


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dx", [1.0, 0.5, 3.0])
def test_divergence_and_gradient(ctx, oracle, shape, dx):
    dim_x, dim_y = shape
    v, _ = rand_fields(4, dim_x, dim_y, 100.0)
    dv = to_dev(v)
    div = torch.empty(dim_y, dim_x, dtype=torch.float32, device="cuda")
    ctx.calculate_divergence(div, dv, dim_x, dim_y, dx)
    want_div = oracle.calculate_divergence(v, dx)
    assert_bit_equal(to_host(div), want_div, "divergence")
    p = np.random.default_rng(5).normal(0, 10, (dim_y, dim_x)).astype(np.float32)
    ctx.subtract_gradient(dv, to_dev(p), dim_x, dim_y, dx)   # in place
    assert_bit_equal(to_host(dv), oracle.subtract_gradient(v.copy(), p, dx), "gradient")


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("iters,omega,dx", [(10, 1.96, 1.0), (1, 1.0, 1.0), (7, 1.5, 2.0), (0, 1.96, 1.0),
                                            (50, 1.96, 1.0)])
def test_poisson_solve(ctx, oracle, sor_variant, shape, iters, omega, dx):
    dim_x, dim_y = shape
    d = np.random.default_rng(6).normal(0, 20, (dim_y, dim_x)).astype(np.float32)
    p = torch.full((dim_y, dim_x), 7.0, dtype=torch.float32, device="cuda")  # must be ignored
    ctx.poisson_solve(p, to_dev(d), dim_x, dim_y, dx, iters, omega)
    assert_bit_equal(to_host(p), oracle.poisson_solve(d, dx, iters, omega), "pressure")


@pytest.mark.parametrize("shape", [0, 1, 2, 3, 4, 5, 6, 7],
                         ids=["direct-96", "direct-192", "tma-96", "tma-192", "tma-144", "tma-160", "tma-192b", "tma-176"])
@pytest.mark.parametrize("one_launch", [1, 0], ids=["one-launch", "launch-per-pass"])
@pytest.mark.parametrize("t_block", [1, 2, 3, 4, 6, 8])
def test_poisson_solve_every_blocking_depth(ctx, oracle, t_block, shape, one_launch):
    if one_launch and shape not in (2, 3, 5):
        pytest.skip("single-launch solve exists for shapes 2, 3, 5")
    ctx.set_option("sor", 1)
    ctx.set_option("sor_t", t_block)
    ctx.set_option("sor_shape", shape)
    ctx.set_option("sor_one_launch", one_launch)
    try:
        for dim_x, dim_y, iters in [(300, 200, 13), (61, 81, 10), (1000, 40, 9), (1024, 1100, 17)]:
            d = np.random.default_rng(7).normal(0, 20, (dim_y, dim_x)).astype(np.float32)
            p = torch.empty(dim_y, dim_x, dtype=torch.float32, device="cuda")
            ctx.poisson_solve(p, to_dev(d), dim_x, dim_y, 1.0, iters, 1.96)
            assert_bit_equal(to_host(p), oracle.poisson_solve(d, 1.0, iters, 1.96), f"T={t_block}")
    finally:
        ctx.set_option("sor_t", 6)
        ctx.set_option("sor_shape", 7)
        ctx.set_option("sor_one_launch", 0)


@pytest.mark.parametrize("shape", [(2, 2), (4, 3), (61, 81), (128, 192), (129, 193), (260, 200), (1023, 517), (1024, 520)])
@pytest.mark.parametrize("field", ["signed-zeros", "sparse", "denormal"])
def test_poisson_walls_and_zeros(ctx, oracle, shape, field):
    """The blocked kernel treats walls without branches: cells outside the domain hold +0 and the
    interior sum is used for wall nodes (poisson.cpp:63-99 vs :101-112).  That is bit-identical only if
    signs of zero and tiny values behave exactly like the reference's running sum — probe exactly those."""
    dim_x, dim_y = shape
    rng = np.random.default_rng(dim_x * 1000 + dim_y)
    d = rng.normal(0, 20, (dim_y, dim_x)).astype(np.float32)
    if field == "signed-zeros":
        d[rng.random(d.shape) < 0.5] = 0.0
        d[rng.random(d.shape) < 0.3] = -0.0
    elif field == "sparse":
        d[:] = 0.0
        d[dim_y // 2, dim_x // 2] = -3.0
        d[0, 0] = 1.0
        d[-1, -1] = -0.0
    else:
        d = (d * np.float32(1e-41)).astype(np.float32)     # denormal divergence -> denormal pressures
    for iters, omega in ((9, 1.96), (3, 1.0), (17, 0.7)):
        p = torch.full((dim_y, dim_x), -7.0, dtype=torch.float32, device="cuda")
        ctx.poisson_solve(p, to_dev(d), dim_x, dim_y, 1.0, iters, omega)
        assert_bit_equal(to_host(p), oracle.poisson_solve(d, 1.0, iters, omega), f"{field} K={iters} w={omega}")


def test_half_sweep_colours(ctx, oracle):
    dim_x, dim_y = 61, 81
    rng = np.random.default_rng(8)
    d = rng.normal(0, 5, (dim_y, dim_x)).astype(np.float32)
    p = rng.normal(0, 5, (dim_y, dim_x)).astype(np.float32)
    for parity in (0, 1):
        dp = to_dev(p)
        ctx.sor_half_sweep(dp, to_dev(d), dim_x, dim_y, 1.0, 1.96, parity)
        assert_bit_equal(to_host(dp), oracle.sor_half_sweep(p.copy(), d, 1.0, 1.96, parity), f"colour {parity}")


def test_fact1_sweep_order_on_device(ctx, sor_variant):
    d = torch.ones(4, 5, dtype=torch.float32, device="cuda")
    p = torch.full((4, 5), 7.0, dtype=torch.float32, device="cuda")
    ctx.poisson_solve(p, d, 5, 4, 1.0, 1, 1.0)
    p = to_host(p)
    assert p[0, 0] == np.float32(-0.5) and p[1, 1] == np.float32(-0.25)
    assert abs(p[0, 1] + 0.6944) < 1e-4 and abs(p[1, 2] + 0.5208) < 1e-4


def test_apply_drags(ctx, oracle):
    dim_x, dim_y = 61, 81
    v, _ = rand_fields(9, dim_x, dim_y, 10.0)
    dr = rand_drags(10, dim_x, dim_y, 300, oob=True)          # > one parameter chunk, some out of range
    dr[17] = dr[3]                                            # duplicate node: later record wins
    dr[17]["vx"] = 123.0
    dv = to_dev(v)
    ctx.apply_drags(dv, dr, dim_x, dim_y)
    assert_bit_equal(to_host(dv), oracle.apply_drags(v.copy(), dr), "drags")


@pytest.mark.parametrize("shape", [(2, 2), (5, 4), (61, 81), (33, 100), (200, 67)])
def test_upscale4_rgb565(ctx, oracle, shape):
    dim_x, dim_y = shape
    _, c = rand_fields(11, dim_x, dim_y, 1.0)
    out = torch.zeros((dim_x - 1) * 4, (dim_y - 1) * 4, dtype=torch.int16, device="cuda")
    ctx.upscale4_rgb565(out, to_dev(c), dim_x, dim_y)
    assert_bit_equal(to_host(out, np.uint16), oracle.upscale4_rgb565(c), "rgb565 frame")


# ---- whole steps ------------------------------------------------------------------------

def run_steps(ctx, v, c, drags_per_step, iters, want_fields=True):
    dim_y, dim_x = v.shape[:2]
    dv, dc = to_dev(v), to_dev(c)
    dp = torch.empty(dim_y, dim_x, dtype=torch.float32, device="cuda")
    dd = torch.empty_like(dp)
    for dr in drags_per_step:
        ctx.step(dv, dc, dr, dim_x, dim_y, DT, 1.0, iters, 1.96, dp, dd)
    return to_host(dv), to_host(dc, np.uint32), to_host(dp), to_host(dd)


@pytest.mark.parametrize("fuse", [0, 1, 2, 3], ids=["unfused", "fuse-div", "fuse-grad", "fused"])
@pytest.mark.parametrize("shape,iters,steps", [((61, 81), 10, 100), ((80, 60), 10, 20), ((7, 3), 10, 5),
                                               ((2, 2), 3, 3), ((300, 260), 50, 3), ((1024, 1024), 50, 2)])
def test_step_matches_oracle(ctx, oracle, shape, iters, steps, fuse):
    """north_star: <=1e-5/step and bounded drift over 100 steps; measured drift here is 0 bits."""
    from esp32_fluid_simulation_b200 import synth
    ctx.set_option("fuse", fuse)
    try:
        dim_x, dim_y = shape
        v = synth.velocity(dim_x, dim_y, vmax=60.0)
        c = synth.dye(dim_x, dim_y)
        drs = [synth.drags(dim_x, dim_y, s, n=16) for s in range(steps)]
        got = run_steps(ctx, v, c, drs, iters)
        ov, oc = v.copy(), c.copy()
        for dr in drs:
            ov, oc, op, od = oracle.step(ov, oc, dr, DT, 1.0, iters, 1.96, want_fields=True)
        for name, g, w in zip("vcpd", got, (ov, oc, op, od)):
            assert_bit_equal(g, w, f"{name} after {steps} steps")
    finally:
        ctx.set_option("fuse", 5)


@pytest.mark.parametrize("shape", [(61, 81), (256, 192), (1000, 333), (1024, 512)])
def test_advect_drags_divergence(ctx, oracle, shape):
    """The fused first half of loop() as its own entry point == advect, drags, divergence one after the other."""
    dim_x, dim_y = shape
    v, _ = rand_fields(21, dim_x, dim_y, 120.0)
    dr = rand_drags(22, dim_x, dim_y, 40)
    out, div = torch.empty(dim_y, dim_x, 2, device="cuda"), torch.empty(dim_y, dim_x, device="cuda")
    ctx.advect_drags_divergence(out, div, to_dev(v), dr, dim_x, dim_y, DT, 1.0)
    want = oracle.apply_drags(oracle.advect_vec2f(v, v, DT, True), dr)
    assert_bit_equal(to_host(out), want, "forced velocity")
    assert_bit_equal(to_host(div), oracle.calculate_divergence(want, 1.0), "divergence")


def test_step_pingpong_equals_step(ctx, oracle):
    """fs_step_pingpong (dye c_in -> c_out, the caller swaps like ino:286) == fs_step == the reference."""
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y, iters, steps = 320, 200, 20, 4
    v, c = synth.velocity(dim_x, dim_y, vmax=80.0), synth.dye(dim_x, dim_y)
    drs = [synth.drags(dim_x, dim_y, s, n=12) for s in range(steps)]
    dv, ca, cb = to_dev(v), to_dev(c), torch.empty(dim_y, dim_x, 3, dtype=torch.int32, device="cuda")
    for dr in drs:
        ctx.step_pingpong(dv, ca, cb, dr, dim_x, dim_y, DT, 1.0, iters, 1.96)
        ca, cb = cb, ca
    ov, oc = v.copy(), c.copy()
    for dr in drs:
        ov, oc = oracle.step(ov, oc, dr, DT, 1.0, iters, 1.96)
    assert_bit_equal(to_host(dv), ov, "v")
    assert_bit_equal(to_host(ca, np.uint32), oc, "c")
    with pytest.raises(Exception):
        ctx.step_pingpong(dv, ca, ca, None, dim_x, dim_y, DT, 1.0, iters, 1.96)   # c_in must differ from c_out


def test_step_golden_regression(ctx, golden):
    """The committed 20-step fixture generated from the reference itself."""
    g = golden("regress20_61x81.npz")
    got = run_steps(ctx, g["v0"], g["c0"], [None] * 20, 10)
    for name, a in zip("vcpd", got):
        assert_bit_equal(a, g[name], name)


def test_step_golden_with_drags(ctx, golden):
    g = golden("steps_drags.npz")
    for k in ("33x17", "80x60"):
        got = run_steps(ctx, g[k + "_v0"], g[k + "_c0"], [g[k + "_drags"]] * 5, 10)
        for name, a in zip("vcpd", got):
            assert_bit_equal(a, g[k + "_" + name], f"{k} {name}")


def test_host_pointer_dropins(ctx, oracle):
    """fsh_*: numpy in, numpy out — the literal drop-in for a reference call."""
    import esp32_fluid_simulation_b200 as fb
    dim_x, dim_y = 61, 81
    v, c = rand_fields(12, dim_x, dim_y, 70.0)
    out = np.empty_like(v)
    fb.advect(out, v, v, dim_x, dim_y, DT, True, ctx=ctx)
    assert_bit_equal(out, oracle.advect_vec2f(v, v, DT, 1), "fsh advect v")
    outc = np.empty_like(c)
    fb.advect(outc, c, v, dim_x, dim_y, DT, False, ctx=ctx)
    assert_bit_equal(outc, oracle.advect_rgb_uq32(c, v, DT, 0), "fsh advect dye")
    div = np.empty((dim_y, dim_x), np.float32)
    fb.calculate_divergence(div, v, dim_x, dim_y, 1.0, ctx=ctx)
    assert_bit_equal(div, oracle.calculate_divergence(v, 1.0), "fsh div")
    p = np.empty_like(div)
    fb.poisson_solve(p, div, dim_x, dim_y, 1.0, 10, 1.96, ctx=ctx)
    assert_bit_equal(p, oracle.poisson_solve(div, 1.0, 10, 1.96), "fsh poisson")
    v2 = v.copy()
    fb.subtract_gradient(v2, p, dim_x, dim_y, 1.0, ctx=ctx)
    assert_bit_equal(v2, oracle.subtract_gradient(v.copy(), p, 1.0), "fsh grad")
    dr = rand_drags(13, dim_x, dim_y, 5)
    hv, hc = v.copy(), c.copy()
    hp, hd = np.empty_like(div), np.empty_like(div)
    ctx.step(hv, hc, dr, dim_x, dim_y, DT, 1.0, 10, 1.96, hp, hd)
    ov, oc, op, od = oracle.step(v.copy(), c.copy(), dr, DT, 1.0, 10, 1.96, want_fields=True)
    for name, a, b in zip("vcpd", (hv, hc, hp, hd), (ov, oc, op, od)):
        assert_bit_equal(a, b, f"fsh step {name}")
    img = np.empty(((dim_x - 1) * 4, (dim_y - 1) * 4), np.uint16)
    ctx.upscale4_rgb565(img, c, dim_x, dim_y)
    assert_bit_equal(img, oracle.upscale4_rgb565(c), "fsh upscale")


@pytest.mark.parametrize("bands,vx,expect_redo", [(8, None, False), (3, None, False), (1, None, False),
                                                   (8, -30000.0, True), (8, 30000.0, None)])
def test_host_pointer_step_dye_bands(ctx, oracle, bands, vx, expect_redo):
    """fsh_step sends the dye up and down in row bands (band b is advected once bands 0..b+1 arrived).  A drag
    fast enough to backtrace DOWN across more than a band (30000 nodes/s * dt = 1000 rows) must trip the
    valid-rectangle flag and be redone in one piece; the same drag pointing up reaches bands that are
    already there (what the projection does to its neighbours is not pinned down).  Same bits every way."""
    dim_x, dim_y = 128, 1024
    v, c = rand_fields(31, dim_x, dim_y, 40.0)
    dr = rand_drags(32, dim_x, dim_y, 6)
    if vx is not None:
        dr["cx"] = np.minimum(dr["cx"], 300)                  # in the first bands, far from the bottom wall
        dr["vx"] = vx                                         # .vx drives the row (j) axis, ino:266
    ctx.set_option("e2e_bands", bands)
    redos = ctx.get_option("e2e_redos")
    try:
        hv, hc = v.copy(), c.copy()
        ctx.step(hv, hc, dr, dim_x, dim_y, DT, 1.0, 8, 1.96)
    finally:
        ctx.set_option("e2e_bands", 8)
    ov, oc = oracle.step(v.copy(), c.copy(), dr, DT, 1.0, 8, 1.96)
    assert_bit_equal(hv, ov, f"fsh step v ({bands} bands)")
    assert_bit_equal(hc, oc, f"fsh step dye ({bands} bands)")
    if expect_redo is not None:
        assert ctx.get_option("e2e_redos") - redos == int(expect_redo)


def test_invalid_arguments(ctx):
    import esp32_fluid_simulation_b200 as fb
    v = torch.zeros(4, 4, 2, device="cuda")
    with pytest.raises(fb.FluidError):
        ctx.advect(v, v, v, 4, 4, DT, True)                  # next_p aliases p
    with pytest.raises(fb.FluidError):
        ctx.calculate_divergence(torch.zeros(4, 1, device="cuda"), torch.zeros(4, 1, 2, device="cuda"), 1, 4, 1.0)
    p = torch.zeros(4, 4, device="cuda")
    with pytest.raises(fb.FluidError):
        ctx.poisson_solve(p, p, 4, 4, 1.0, 1, 1.0)           # p aliases div
    with pytest.raises(ValueError):
        ctx.poisson_solve(p, np.zeros((4, 4), np.float32), 4, 4, 1.0, 1, 1.0)


# ---- BASELINE.json full sizes: oracle where it finishes in seconds, properties beyond ------

def test_4096_one_step_vs_oracle(ctx, oracle):
    """configs[2]: 4096x4096, 50 SOR iterations — one full step against the oracle (~6 s of CPU)."""
    from esp32_fluid_simulation_b200 import synth
    n = 4096
    v, c = synth.velocity(n, n), synth.dye(n, n)
    dr = synth.drags(n, n, 0, n=16)
    got = run_steps(ctx, v, c, [dr], 50)
    ov, oc, op, od = oracle.step(v, c, dr, DT, 1.0, 50, 1.96, want_fields=True)
    for name, a, b in zip("vcpd", got, (ov, oc, op, od)):
        assert_bit_equal(a, b, f"4096^2 {name}")


def test_4096_size_independent_properties(ctx):
    """(a) SOR is linear in d and power-of-two scaling is exact in binary floating point, so
    solve(4*d) == 4*solve(d) bit for bit; (b) the initial contents of p are ignored; (c) zero
    velocity makes velocity advection the identity and rounds the dye to 24 significant bits."""
    n = 4096
    g = torch.Generator(device="cuda").manual_seed(1)
    d = torch.randn(n, n, device="cuda", generator=g) * 10
    p1 = torch.full((n, n), 3.0, device="cuda")
    p2 = torch.full((n, n), -9.0, device="cuda")
    ctx.poisson_solve(p1, d, n, n, 1.0, 50, 1.96)
    ctx.poisson_solve(p2, d * 4, n, n, 1.0, 50, 1.96)
    assert torch.equal(p1 * 4, p2)
    v = torch.randn(n, n, 2, device="cuda", generator=g)
    z = torch.zeros_like(v)
    out = torch.empty_like(v)
    ctx.advect(out, v, z, n, n, DT, True)
    assert torch.equal(out, v)
    c = torch.randint(2 ** 24, 2 ** 30, (n, n, 3), device="cuda", dtype=torch.int32, generator=g)
    outc = torch.empty_like(c)
    ctx.advect(outc, c, z, n, n, DT, False)
    want = c.to(torch.float32).to(torch.int64).clamp(max=2 ** 31 - 1).to(torch.int32)  # RNE to 24 bits
    want[-1, -1] = c[-1, -1]                                  # corner-copy node is exact
    assert torch.equal(outc, want)


# ---- batched ensemble: one CTA per grid, state resident in shared memory ------------------

@pytest.mark.parametrize("shape,iters,batch,n_steps", [((61, 81), 10, 5, 1), ((80, 60), 10, 300, 3), ((60, 80), 7, 4, 2),
                                                       ((7, 3), 4, 3, 2), ((2, 2), 3, 2, 1), ((64, 90), 10, 2, 1)])
@pytest.mark.parametrize("variant", [0, 5, 6, 7, 8, 9, 14, 20, 21])
def test_ensemble_step(ctx, oracle, shape, iters, batch, n_steps, variant):
    """variant = option "ensemble": 0 automatic (register-tiled projection, ensemble_reg.cuh), 5 the
    first-generation kernel, 6-9 register-tiled with 2/4/6/8 rows per thread, 14 = 4 rows, dye streamed, 20 / 21 =
    automatic with the bulk-copy pipelined flow forced on / off."""
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y = shape
    if variant not in (0, 5) and batch > 8:
        batch = 149 + variant          # (more grids than CTAs for every variant without repeating the big case 7 times)
    max_drags = 4
    v = np.stack([synth.velocity(dim_x, dim_y, seed=100 + b, vmax=90.0) for b in range(batch)])
    c = np.stack([synth.dye(dim_x, dim_y, seed=200 + b, n_splats=6) for b in range(batch)])
    drags = np.zeros((n_steps, batch, max_drags), synth.DRAG_DTYPE)
    counts = np.zeros((n_steps, batch), np.int32)
    for s in range(n_steps):
        for b in range(batch):
            k = (b + s) % (max_drags + 1)
            counts[s, b] = k
            drags[s, b, :k] = synth.drags(dim_x, dim_y, s * 1000 + b, n=max_drags, vmax=300.0)[:k]
    dv, dc = to_dev(v), to_dev(c)
    ctx.set_option("ensemble", variant)
    try:
        ctx.ensemble_step(dv, dc, batch, dim_x, dim_y, DT, 1.0, iters, 1.96, n_steps, drags, counts, max_drags)
    finally:
        ctx.set_option("ensemble", 0)
    gv, gc = to_host(dv), to_host(dc, np.uint32)
    check = range(batch) if batch <= 8 else [0, 1, 147, 148, 149, batch - 1]
    for b in check:
        ov, oc = v[b].copy(), c[b].copy()
        for s in range(n_steps):
            ov, oc = oracle.step(ov, oc, drags[s, b, :counts[s, b]], DT, 1.0, iters, 1.96)
        assert_bit_equal(gv[b], ov, f"grid {b} velocity")
        assert_bit_equal(gc[b], oc, f"grid {b} dye")


def test_ensemble_too_large_is_unsupported(ctx):
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200._lib import FS_ERR_UNSUPPORTED
    v = torch.zeros(1, 128, 128, 2, device="cuda")
    c = torch.zeros(1, 128, 128, 3, device="cuda", dtype=torch.int32)
    with pytest.raises(fb.FluidError) as e:
        ctx.ensemble_step(v, c, 1, 128, 128, DT, 1.0, 10, 1.96)
    assert e.value.code == FS_ERR_UNSUPPORTED


@pytest.mark.parametrize("shape", [(5, 4), (61, 81), (512, 300)])
def test_poisson_residual(ctx, oracle, shape):
    """Residual probe (warp-shuffle reduction): max |gs(p) - p| is exact, the 2-norm to 1e-6."""
    dim_x, dim_y = shape
    rng = np.random.default_rng(21)
    d = rng.normal(0, 5, (dim_y, dim_x)).astype(np.float32)
    for iters in (0, 5, 40):
        p = oracle.poisson_solve(d, 1.0, iters, 1.96)
        q = p.copy()                                           # one Gauss-Seidel value per node from the SAME p:
        ge = oracle.sor_half_sweep(p.copy(), d, 1.0, 1.0, 0)   # omega = 1 turns the SOR update into gs itself
        go = oracle.sor_half_sweep(p.copy(), d, 1.0, 1.0, 1)
        jj, ii = np.indices(p.shape)
        gs = np.where((ii + jj) % 2 == 0, ge, go)
        r = (gs - q).astype(np.float32)
        m, l2 = ctx.poisson_residual(to_dev(p), to_dev(d), dim_x, dim_y, 1.0)
        assert np.float32(m) == np.abs(r).max()
        assert abs(l2 - np.sqrt((r.astype(np.float64) ** 2).sum())) <= 1e-6 * max(l2, 1e-30)
        assert np.isfinite(l2)


# ---- the sketch's own initial condition and input path on the device ------------------

@pytest.mark.parametrize("shape,batch", [((61, 81), 1), ((80, 60), 3), ((5, 4), 2), ((2, 2), 1), ((16, 9), 1), ((512, 300), 1),
                                         ((1024, 768), 1)])
def test_init_color_wheel_on_device(ctx, oracle, shape, batch):
    """fs_init_color_wheel == setup() of the sketch (ino:196-241; the compiled sketch when oracle/_ref travelled)."""
    dim_x, dim_y = shape
    v = torch.full((batch, dim_y, dim_x, 2), 7.0, device="cuda")
    c = torch.full((batch, dim_y, dim_x, 3), 7, dtype=torch.int32, device="cuda")
    ctx.init_color_wheel(v, c, batch, dim_x, dim_y)
    ov, oc = oracle.init_color_wheel(dim_x, dim_y)
    for b in range(batch):
        assert_bit_equal(to_host(v[b]), ov, f"velocity, grid {b}")
        assert_bit_equal(to_host(c[b], np.uint32), oc, f"dye, grid {b}")


def test_touch_to_drags_on_device(ctx, oracle):
    """fs_touch_to_drags == touch_routine() of the sketch (ino:63-96), incl. the queue's depth-10 drop."""
    from esp32_fluid_simulation_b200 import synth
    rng = np.random.default_rng(11)
    batch, n_samples, max_drags = 37, 70, 10
    samples = np.zeros((batch, n_samples, 3), np.int32)
    samples[..., 0] = rng.random((batch, n_samples)) < 0.8
    samples[..., 1:] = rng.integers(100, 4000, (batch, n_samples, 2))
    samples[0, :, 0] = 0                                   # nobody touches grid 0
    samples[1, :, 0] = 1                                   # a continuous swipe: 69 records, 10 kept
    d_drags = torch.zeros(batch, max_drags, 3, dtype=torch.int32, device="cuda")
    d_counts = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ctx.touch_to_drags(d_drags, d_counts, to_dev(samples), n_samples, batch, max_drags, 61, 81)
    got = d_drags.cpu().numpy().view(synth.DRAG_DTYPE).reshape(batch, max_drags)
    counts = d_counts.cpu().numpy()
    for b in range(batch):
        script = [tuple(int(x) for x in r) for r in samples[b]]
        if getattr(oracle, "ref", None) is not None and oracle.ref.has_ino:
            want = oracle.ref.ino_touch(script)            # the sketch itself: its queue holds 10
        else:
            want = synth.touch_drags(script, 61, 81)[:max_drags]
        assert counts[b] == len(want), f"grid {b}: {counts[b]} records, want {len(want)}"
        assert_bit_equal(got[b, :len(want)].view(np.uint8), want.view(np.uint8), f"grid {b}")
    assert counts[0] == 0 and counts[1] == 10


def test_ensemble_seeded_and_driven_on_device(ctx, oracle):
    """setup() -> touch_routine() -> loop() x 3 without leaving the device == the same chain on the CPU checker."""
    from esp32_fluid_simulation_b200 import synth
    rng = np.random.default_rng(12)
    batch, n_samples, max_drags, dim_x, dim_y, n_steps = 9, 12, 10, 61, 81, 3
    v = torch.empty(batch, dim_y, dim_x, 2, device="cuda")
    c = torch.empty(batch, dim_y, dim_x, 3, dtype=torch.int32, device="cuda")
    ctx.init_color_wheel(v, c, batch, dim_x, dim_y)
    samples = np.zeros((n_steps, batch, n_samples, 3), np.int32)
    samples[..., 0] = rng.random((n_steps, batch, n_samples)) < 0.7
    samples[..., 1] = rng.integers(300, 3600, (n_steps, batch, n_samples))       # inside the calibrated range: on the grid
    samples[..., 2] = rng.integers(300, 3700, (n_steps, batch, n_samples))
    d_drags = torch.zeros(n_steps, batch, max_drags, 3, dtype=torch.int32, device="cuda")
    d_counts = torch.zeros(n_steps, batch, dtype=torch.int32, device="cuda")
    d_samples = to_dev(samples)
    for s in range(n_steps):
        ctx.touch_to_drags(d_drags[s], d_counts[s], d_samples[s], n_samples, batch, max_drags, dim_x, dim_y)
    ctx.ensemble_step_dev(v, c, d_drags, d_counts, max_drags, batch, dim_x, dim_y, DT, 1.0, 10, 1.96, n_steps)
    gv, gc = to_host(v), to_host(c, np.uint32)
    for b in range(batch):
        ov, oc = oracle.init_color_wheel(dim_x, dim_y)
        for s in range(n_steps):
            dr = synth.touch_drags([tuple(int(x) for x in r) for r in samples[s, b]], dim_x, dim_y)[:max_drags]
            ov, oc = oracle.step(ov, oc, dr, DT, 1.0, 10, 1.96)
        assert_bit_equal(gv[b], ov, f"grid {b} velocity")
        assert_bit_equal(gc[b], oc, f"grid {b} dye")


# ---- f1: the RGB565 frame rendered inside the dye advect (ino:282 + ino:116-177) -----------------------

@pytest.mark.parametrize("fuse", [5, 1], ids=["frame-in-advect", "advect-then-upscale"])
@pytest.mark.parametrize("shape", [(64, 32), (61, 81), (130, 70), (200, 67), (512, 300), (1024, 512)])
def test_advect_rgb_frame(ctx, oracle, shape, fuse):
    dim_x, dim_y = shape
    v, c = rand_fields(31, dim_x, dim_y, 150.0)
    c[0, 0] = 0xFFFFFFFF
    ctx.set_option("fuse", fuse)
    try:
        out = torch.empty(dim_y, dim_x, 3, dtype=torch.int32, device="cuda")
        frame = torch.zeros((dim_x - 1) * 4, (dim_y - 1) * 4, dtype=torch.int16, device="cuda")
        ctx.advect_rgb_frame(out, frame, to_dev(c), to_dev(v), dim_x, dim_y, DT, False)
    finally:
        ctx.set_option("fuse", 5)
    want = oracle.advect_rgb_uq32(c, v, DT, False)
    assert_bit_equal(to_host(out, np.uint32), want, "advected dye")
    assert_bit_equal(to_host(frame, np.uint16), oracle.upscale4_rgb565(want), "RGB565 frame")


def test_step_frame(ctx, oracle):
    """loop() + draw_routine(): fs_step_frame == the reference step followed by the reference's frame."""
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y, iters, steps = 320, 200, 20, 3
    v, c = synth.velocity(dim_x, dim_y, vmax=80.0), synth.dye(dim_x, dim_y)
    dv, ca, cb = to_dev(v), to_dev(c), torch.empty(dim_y, dim_x, 3, dtype=torch.int32, device="cuda")
    frame = torch.zeros((dim_x - 1) * 4, (dim_y - 1) * 4, dtype=torch.int16, device="cuda")
    ov, oc = v.copy(), c.copy()
    for s in range(steps):
        dr = synth.drags(dim_x, dim_y, s, n=12)
        ctx.step_frame(dv, ca, cb, frame, dr, dim_x, dim_y, DT, 1.0, iters, 1.96)
        ca, cb = cb, ca
        ov, oc = oracle.step(ov, oc, dr, DT, 1.0, iters, 1.96)
        assert_bit_equal(to_host(frame, np.uint16), oracle.upscale4_rgb565(oc), f"frame after step {s}")
    assert_bit_equal(to_host(dv), ov, "v")
    assert_bit_equal(to_host(ca, np.uint32), oc, "c")


def test_upscale_into_an_unaligned_frame(ctx, oracle):
    """fs_upscale4_rgb565 into a frame that is only 2-byte aligned (a uint16_t* inside a larger buffer)."""
    dim_x, dim_y = 33, 17
    _, c = rand_fields(32, dim_x, dim_y, 1.0)
    n = 16 * (dim_x - 1) * (dim_y - 1)
    buf = torch.zeros(n + 8, dtype=torch.int16, device="cuda")
    for off in (1, 2, 3):
        ctx.upscale4_rgb565(buf[off:off + n], to_dev(c), dim_x, dim_y)
        assert_bit_equal(to_host(buf[off:off + n], np.uint16).reshape((dim_x - 1) * 4, (dim_y - 1) * 4),
                         oracle.upscale4_rgb565(c), f"offset {off}")


# ---- f2: one CUDA-graph launch per step + the asynchronous frame stream (ino:285-288) -------------------

def test_sim_graph_steps_match_oracle(oracle):
    """fs_sim: after two eager steps the step is ONE graph launch whose drag records are re-armed per step."""
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y, iters, steps = 256, 192, 20, 9
    ctx = fb.Context(0, torch.cuda.Stream())
    sim = fb.Sim(ctx, dim_x, dim_y, DT, 1.0, iters, 1.96)
    v, c = synth.velocity(dim_x, dim_y, vmax=90.0), synth.dye(dim_x, dim_y)
    sim.upload(v, c)
    launches0 = ctx.launch_count
    ov, oc = v.copy(), c.copy()
    for s in range(steps):
        dr = synth.drags(dim_x, dim_y, s, n=(s * 5) % 17)          # a different number of records every step, 0 included
        assert sim.step(dr)
        ov, oc, op, od = oracle.step(ov, oc, dr, DT, 1.0, iters, 1.96, want_fields=True)
    gv, gc, gp, gd = sim.download()
    assert_bit_equal(gv, ov, "v")
    assert_bit_equal(gc, oc, "c")
    assert_bit_equal(gp, op, "p")
    assert_bit_equal(gd, od, "d")
    st = sim.stats
    assert st["steps"] == steps and st["graph_launches"] == steps - 2 and st["eager_steps"] == 2, st
    per_step = (ctx.launch_count - launches0) / steps
    assert per_step == int(per_step) and per_step >= 4, per_step     # kernels per step are still counted
    sim.close()


def test_sim_small_grid_steps_without_a_graph(oracle):
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200 import synth
    ctx = fb.Context(0)
    sim = fb.Sim(ctx, 61, 81, DT, 1.0, 10, 1.96)
    v, c = synth.velocity(61, 81, vmax=90.0), synth.dye(61, 81)
    sim.upload(v, c)
    ov, oc = v.copy(), c.copy()
    for s in range(5):
        dr = synth.drags(61, 81, s, n=4)
        assert sim.step(dr)
        ov, oc = oracle.step(ov, oc, dr, DT, 1.0, 10, 1.96)
    gv, gc, _, _ = sim.download()
    assert_bit_equal(gv, ov, "v")
    assert_bit_equal(gc, oc, "c")
    assert sim.stats["graph_launches"] == 0
    sim.close()


def test_sim_frame_stream_is_a_double_buffer(oracle):
    """The colour hand-off of ino:285-288: the producer may be two frames ahead, then it must wait for the consumer;
    frames arrive in order in pinned host memory and equal draw_routine() of the reference on each step's dye."""
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y, iters = 192, 128, 12
    ctx = fb.Context(0, torch.cuda.Stream())
    sim = fb.Sim(ctx, dim_x, dim_y, DT, 1.0, iters, 1.96, frame=True)
    v, c = synth.velocity(dim_x, dim_y, vmax=90.0), synth.dye(dim_x, dim_y)
    sim.upload(v, c)
    assert sim.acquire_frame() is None                               # nothing produced yet
    ov, oc = v.copy(), c.copy()
    want = []
    produced = consumed = 0
    for s in range(8):
        dr = synth.drags(dim_x, dim_y, s, n=6)
        if not sim.step(dr):                                         # both slots full: consume one, then the step goes through
            assert produced - consumed == 2
            f = sim.acquire_frame()
            assert_bit_equal(f, want[consumed], f"frame {consumed}")
            sim.release_frame()
            consumed += 1
            assert sim.step(dr)
        produced += 1
        ov, oc = oracle.step(ov, oc, dr, DT, 1.0, iters, 1.96)
        want.append(oracle.upscale4_rgb565(oc))
    while consumed < produced:
        f = sim.acquire_frame()
        assert_bit_equal(f, want[consumed], f"frame {consumed}")
        sim.release_frame()
        consumed += 1
    assert sim.acquire_frame() is None
    gv, gc, _, _ = sim.download()
    assert_bit_equal(gv, ov, "v")
    assert_bit_equal(gc, oc, "c")
    assert sim.stats["graph_launches"] >= 4
    sim.close()
