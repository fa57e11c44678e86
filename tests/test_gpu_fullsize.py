"""Full-size (64^4, BASELINE.json configs[3]) checks that do not need the CPU oracle to finish in seconds:
size-independent properties and agreement of the alternative kernels with each other, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RES = 64
DIMS = (RES, RES, RES, RES)


@pytest.fixture(scope="module")
def env():
    from ofblend_b200 import capi, synth
    ctx = capi.Context(0)
    api = capi.HostAPI(ctx)
    i0 = ctx.to_device(synth.post_process(synth.two_drop_phi(DIMS, 0), api))
    i1 = ctx.to_device(synth.post_process(synth.two_drop_phi(DIMS, 1), api))
    yield ctx, i0, i1
    ctx.close()


def test_extrapolation_kernels_agree_bit_for_bit(env):
    """Dense kernel, Vec4 work list and component-plane work list on the real marker of the 64^4 pair:
    the same 7 sweeps must give identical bits (each keeps the reference's 81-tap order)."""
    ctx, i0, i1 = env
    vel = ctx.grid(DIMS, 4)
    ctx.optical_flow4d(vel, i0, i1, None, 1e-3, 1e-4, 4., 1e-2, 0.1)
    dst = ctx.grid(DIMS, 4)
    mk = ctx.grid(DIMS, 1)
    ctx.project_cells(dst, vel, i0, i1, mk, 4., 40)
    start = dst.download()
    m = mk.download()
    assert 0.05 < float((m == 0).mean()) < 0.6           # the workload really is sparse-but-scattered
    out = {}
    default_mode = ctx.get_option("expol_mode")
    for mode in (2, 1, 0, 3, 4, 5, 6):
        ctx.set_option("expol_mode", mode)
        dst.upload(start)
        ctx.cv_expol_blur4d(dst, mk, 7)
        out[mode] = dst.download()
    ctx.set_option("expol_mode", default_mode)
    assert np.array_equal(out[1], out[2])
    assert np.array_equal(out[0], out[2])
    assert np.array_equal(out[3], out[2])                 # 4y x 2z work-list items
    assert np.array_equal(out[4], out[2])                 # 4y x 4z work-list items
    assert np.array_equal(out[5], out[2]) and np.array_equal(out[6], out[2])   # the same with lane shuffles
    # marked cells and the outer shell never change (ref knCvExpolBlur4d :613-626, bnd = 1)
    keep = (m != 0)
    keep[0], keep[-1], keep[:, 0], keep[:, -1] = True, True, True, True
    keep[:, :, 0], keep[:, :, -1], keep[:, :, :, 0], keep[:, :, :, -1] = True, True, True, True
    s4 = start.reshape(m.shape + (4,))
    o4 = out[1].reshape(m.shape + (4,))
    assert np.array_equal(o4[keep], s4[keep])
    assert not np.array_equal(o4[~keep], s4[~keep])
    for g in (vel, dst, mk):
        g.free()


def test_cg_apply_variants_same_iterations_and_bits(env):
    """The CG apply variants only change cache hints / occupancy: stopping iteration and solution bits are equal."""
    ctx, i0, i1 = env
    res = {}
    for v in (0, 1, 7, 11):
        ctx.set_option("apply_variant", v)
        vel = ctx.grid(DIMS, 4)
        it = ctx.optical_flow4d(vel, i0, i1, None, 1e-3, 1e-4, 0., 1e-2, -1.)
        res[v] = (it, vel.download())
        vel.free()
    ctx.set_option("apply_variant", 11)
    assert res[0][0] == res[1][0] == res[7][0] == res[11][0]
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][1], res[7][1]) and np.array_equal(res[0][1], res[11][1])


def test_identical_inputs_give_zero_deformation(env):
    """i0 == i1: zero right-hand side, the reference returns before the first CG iteration (:287-291).  (Without the
    final projection: projectCell's +-step walk leaves a non-zero d even for identical SDFs, in the reference too.)"""
    ctx, i0, _ = env
    from ofblend_b200 import capi, synth
    vel = ctx.grid(DIMS, 4)
    p = dict(synth.MODE1_PARAMS)
    p["doFinalProject"] = False
    err, tr = ctx.optical_flow_multiscale4d(vel, i0, i0, capi.make_params(**p), want_trace=True)
    assert not vel.download().any()
    assert all(int(tr.cg_iters[q]) == 0 for q in range(min(tr.n_solves, 64)))
    assert err == 0.0
    vel.free()


def test_zero_velocity_advection_is_identity_inside_the_shell(env):
    """advect4d with vel = 0: interior cells keep their value exactly, the 1-cell shell is zero (ref :1275-1290)."""
    ctx, i0, _ = env
    vel = ctx.grid(DIMS, 4)
    g = ctx.grid(DIMS, 1)
    src = i0.download()
    g.upload(src)
    ctx.advect4d(vel, g, 1.0)
    a = g.download().reshape(DIMS[::-1])
    s = src.reshape(DIMS[::-1])
    assert np.array_equal(a[1:-1, 1:-1, 1:-1, 1:-1], s[1:-1, 1:-1, 1:-1, 1:-1])
    shell = np.ones(a.shape, bool)
    shell[1:-1, 1:-1, 1:-1, 1:-1] = False
    assert not a[shell].any()
    vel.free()
    g.free()


def test_gaussian_blur_is_linear_at_full_size(env):
    """blur(a + b) == blur(a) + blur(b) up to fp32 rounding of the 625-tap sums, and a constant field stays constant
    in the cells whose window lies inside the grid (normalised weights)."""
    ctx, i0, i1 = env
    rng = np.random.default_rng(5)
    shape = DIMS[::-1] + (4,)
    A = rng.standard_normal(shape).astype(np.float32)
    B = rng.standard_normal(shape).astype(np.float32)
    out = []
    for x in (A, B, A + B, np.full(shape, 1.25, np.float32)):
        g = ctx.grid(DIMS, 4)
        g.upload(x)
        ctx.gaussian_blur4d(g, 2.0, 1)
        out.append(g.download().reshape(shape))
        g.free()
    inner = (slice(4, -4),) * 4
    assert np.abs(out[2][inner] - (out[0][inner] + out[1][inner])).max() < 2e-5
    assert np.abs(out[3][inner] - 1.25).max() < 1e-5


def test_mode1_at_128_matches_the_record_and_mode2_matches_the_oracle():
    """BASELINE.json configs[4] (128^4, the north-star size) on one GPU.  A reference run of mode 1 at this size takes CPU
    hours, so the solve itself is held to the committed single-GPU record (CG iterations of all eleven solves, final error,
    integer checksum of the deformation's bit patterns -- the record every multi-GPU bench line is compared with), and
    what CAN be checked against the CPU at full size is: the deformation applied to the 4D SDF (mode 2, advect4d) and the
    error metric, both bit for bit against the oracle (pinned to the reference by tests/test_oracle_vs_ref.py)."""
    import json
    import os
    from ofblend_b200 import capi, synth
    from oracle import port
    res = 128
    dims = (res,) * 4
    ctx = capi.Context(0)
    try:
        api = capi.HostAPI(ctx)
        h0 = synth.post_process(synth.two_drop_phi(dims, 0), api)
        h1 = synth.post_process(synth.two_drop_phi(dims, 1), api)
        i0, i1 = ctx.to_device(h0), ctx.to_device(h1)
        vel = ctx.grid(dims, 4)
        err, tr = ctx.optical_flow_multiscale4d(vel, i0, i1, capi.make_params(**synth.MODE1_PARAMS), want_trace=True)
        rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_n1_record.json")))["128"]
        assert list(tr.cg_iters[:tr.n_solves]) == rec["cg_iters"]
        assert float(err) == rec["final_error"]
        v = vel.download()
        u = np.ascontiguousarray(v).view(np.uint32).ravel()
        chk = "%016x-%08x" % (int(u.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(u)))
        assert chk == rec["deformation_checksum"], chk
        # mode 2 at full size: GPU advect4d of the GPU deformation == oracle advect4d of the same deformation
        g = ctx.to_device(h0)
        ctx.advect4d(vel, g, 1.0)
        adv = g.download().reshape(h0.shape)
        port.set_threads(os.cpu_count() or 1)
        want = port.advect4d(v, h0)
        assert np.array_equal(adv, want)
        # (fp64 sums of 2.4e8 terms in a different grouping: equal to the last float32 digit, not bit-guaranteed)
        np.testing.assert_allclose(ctx.calc_ls_diff4d(g, i1, None, 0.005, 13), port.calc_ls_diff4d(want, h1, 0.005, 13), rtol=1e-6)
        for x in (i0, i1, vel, g):
            x.free()
    finally:
        ctx.close()
