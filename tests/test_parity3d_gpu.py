"""GPU parity of the 3D Poisson path: every operator, the V-cycle and FMG of libmg_b200.so against the
CPU oracle on the same seeded inputs.  Bar: bit-exact (float and double), both residual modes.
All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest

from util import assert_bits_equal, oracles, random_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
RANGES = [(0, 1, 0, 1, 0, 1), (0.25, 1.75, -0.5, 0.7, 0.1, 3.3)]  # the second has h not a power of two


def _pair(mg, n, dtype, corrected, rng_range, seed=12345):
    """engine + oracles with identical random v, f on the finest level"""
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
    eng = mg.MultiGrid3D(n, rng_range, dtype=dtype, residual_mode=mode)
    orcs = oracles(3, dtype, corrected, n, range=rng_range)
    rng = np.random.default_rng(seed)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    eng.set_v(0, v0)
    eng.set_f(0, f0)
    for o in orcs:
        o.v(0)[...] = v0
        o.f(0)[...] = f0
    return eng, orcs, v0, f0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 9, 17, 33, 65])
def test_init_problem(mg, n, dtype):
    eng = mg.MultiGrid3D(n, RANGES[0], dtype=dtype)
    for o in oracles(3, dtype, False, n, range=RANGES[0]):
        for l in range(eng.numGrids):
            assert eng.level_size(l) == o.sizes[l]
            assert_bits_equal(eng.get_f(l), o.f(l), "InitF level %d" % l)
            assert_bits_equal(eng.get_v(l), o.v(l), "InitV level %d" % l)
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_init_problem_general_range(mg, dtype):
    eng = mg.MultiGrid3D(33, RANGES[1], dtype=dtype)
    for o in oracles(3, dtype, False, 33, range=RANGES[1]):
        for l in range(eng.numGrids):
            assert_bits_equal(eng.get_f(l), o.f(l), "InitF level %d" % l)
    eng.close()


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 9, 17, 33, 65])
def test_relax(mg, n, dtype, rng_range):
    eng, orcs, _, _ = _pair(mg, n, dtype, False, rng_range)
    eng.Relax(0, 3)
    got = eng.get_v(0)
    for o in orcs:
        o.relax(0, 3)
        assert_bits_equal(got, o.v(0), "Relax n=%d" % n)
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 17, 33, 65])
def test_residual_and_norm(mg, n, dtype, rng_range, corrected):
    eng, orcs, _, _ = _pair(mg, n, dtype, corrected, rng_range)
    r = eng.CalculateResidual(0)
    l2, linf = eng.residual_norm(0)
    for o in orcs:
        ro = o.residual(0)
        assert_bits_equal(r, ro, "CalculateResidual")
        ol2, olinf = o.residual_norms(0)
        assert abs(l2 - ol2) <= 1e-12 * max(ol2, 1e-300)
        assert linf == olinf
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [5, 9, 17, 33, 65, 129])
def test_restrict_interpolate_correct_set_host_ops(mg, n, dtype):
    """The NOCUDA-signature operators on caller-owned host arrays."""
    eng = mg.MultiGrid3D(5, dtype=dtype)
    orcs = oracles(3, dtype, False, 5)
    rng = np.random.default_rng(7)
    fine = random_field(rng, (n,) * 3, dtype)
    coarse = eng.Restrict(fine)
    fine2 = random_field(rng, (n,) * 3, dtype)
    got_i = eng.Interpolate(fine2.copy(), coarse)
    got_c = eng.ApplyCorrection(fine.copy(), fine2)
    got_s0 = eng.setToValue(fine.copy(), 2.5, False)
    got_s1 = eng.setToValue(fine.copy(), -1.0, True)
    for o in orcs:
        assert_bits_equal(coarse, o.restrict(fine), "Restrict")
        assert_bits_equal(got_i, o.interpolate(fine2.copy(), coarse), "Interpolate")
        assert_bits_equal(got_c, o.apply_correction(fine.copy(), fine2), "ApplyCorrection")
        assert_bits_equal(got_s0, o.set_to_value(fine.copy(), 2.5, False), "setToValue interior")
        assert_bits_equal(got_s1, o.set_to_value(fine.copy(), -1.0, True), "setToValue all")
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [5, 9, 17, 33, 65, 129, 257])
def test_fused_residual_restrict_and_interpolate_correct(mg, n, dtype, corrected):
    eng, orcs, v0, f0 = _pair(mg, n, dtype, corrected, RANGES[1])
    rng = np.random.default_rng(99)
    cn = (n - 1) // 2 + 1
    cv = random_field(rng, (cn,) * 3, dtype)
    eng.set_v(1, cv)  # must be overwritten with zeros by the fused kernel
    eng.residual_restrict(0)
    got_cf, got_cv = eng.get_f(1), eng.get_v(1)
    eng.set_v(1, cv)
    eng.interpolate_correct(0)
    got_v = eng.get_v(0)
    eng.interpolate_level(0)
    got_vi = eng.get_v(0)
    eng.restrict_level(0, mg.MG_FIELD_F)
    got_rf = eng.get_f(1)
    for o in orcs:
        assert_bits_equal(got_cf, o.restrict(o.residual(0)), "fused residual+restrict")
        assert not got_cv.any(), "coarse v must be zeroed including the boundary"
        e = np.zeros((n,) * 3, dtype)
        o.interpolate(e, cv)
        assert_bits_equal(got_v, o.apply_correction(v0.copy(), e), "fused interpolate+correct")
        assert_bits_equal(got_vi, o.interpolate(got_v.copy(), cv), "Interpolate into v")
        assert_bits_equal(got_rf, o.restrict(f0), "Restrict(f)")
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu", [(17, 2), (33, 2), (65, 2), (33, 1), (17, 50)])
def test_vcycle_history(mg, n, nu, dtype, corrected):
    """Reference problem from v = 0: residual-norm history per V-cycle and the final solution."""
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
    eng = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mode)
    orcs = oracles(3, dtype, corrected, n)
    hist = [eng.residual_norm(0)]
    for _ in range(3):
        eng.VCycle(0, nu, nu)
        hist.append(eng.residual_norm(0))
    for o in orcs:
        ohist = [o.residual_norms(0)]
        for _ in range(3):
            o.vcycle(0, nu, nu)
            ohist.append(o.residual_norms(0))
        for l in range(eng.numGrids):
            assert_bits_equal(eng.get_v(l), o.v(l), "v level %d after 3 V-cycles" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "f level %d after 3 V-cycles" % l)
        for (a, am), (b, bm) in zip(hist, ohist):
            assert abs(a - b) <= 1e-10 * abs(b), (hist, ohist)  # north-star tolerance; v itself is bitwise
            assert am == bm
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
def test_fmg(mg, dtype, corrected):
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
    eng = mg.MultiGrid3D(33, dtype=dtype, residual_mode=mode)
    eng.FullMultiGridVCycle(0, 2, 3, 3)
    for o in oracles(3, dtype, corrected, 33):
        o.fmg(0, 2, 3, 3)
        for l in range(eng.numGrids):
            assert_bits_equal(eng.get_v(l), o.v(l), "FMG v level %d" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "FMG f level %d" % l)
    eng.close()


def test_vcycle_host_roundtrip(mg):
    n, dtype = 33, np.float64
    eng = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    o = oracles(3, dtype, True, n)[0]
    v = np.zeros((n,) * 3, dtype)
    f = o.f(0).copy()
    eng.vcycle_host(v, f, 2, 2, cycles=2)
    o.vcycle(0, 2, 2)
    o.vcycle(0, 2, 2)
    assert_bits_equal(v, o.v(0), "vcycle_host")
    eng.close()


def test_argument_errors(mg):
    with pytest.raises(mg.MGError):
        mg.MultiGrid3D([17, 17, 9])
    with pytest.raises(mg.MGError):
        mg.MultiGrid3D(18)
    with pytest.raises(mg.MGError):
        mg.MultiGrid3D(17, (0, 1, 1, 0, 0, 1))
    eng = mg.MultiGrid3D(9)
    with pytest.raises(mg.MGError):
        eng.Relax(7, 1)
    with pytest.raises(mg.MGError):
        eng.Interpolate(np.zeros((9,) * 3, np.float32), np.zeros((4,) * 3, np.float32))
    eng.close()


# ---- large levels: the TMA-staged z-marching smoother (n >= 257) --------------------------------

@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
def test_relax_tma_path_vs_oracle(mg, dtype, rng_range):
    n = 257
    eng, orcs, _, _ = _pair(mg, n, dtype, False, rng_range)
    eng.Relax(0, 2)
    got = eng.get_v(0)
    o = orcs[0]  # the C restatement (the compiled reference is ~10x slower at this size)
    o.relax(0, 2)
    assert_bits_equal(got, o.v(0), "Relax n=257 (TMA smoother)")
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_vcycle_257_vs_oracle(mg, dtype):
    n = 257
    eng = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    o = oracles(3, dtype, True, n)[0]
    eng.VCycle(0, 2, 2)
    o.vcycle(0, 2, 2)
    for l in range(eng.numGrids):
        assert_bits_equal(eng.get_v(l), o.v(l), "v level %d after V(2,2) at 257^3" % l)
    l2, linf = eng.residual_norm(0)
    ol2, olinf = o.residual_norms(0)
    assert abs(l2 - ol2) <= 1e-10 * ol2 and linf == olinf
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [257, 513])
def test_smoother_variants_are_bit_identical(mg, n, dtype):
    """MG_SMOOTHER_COLOUR (plain kernel) and the default (TMA-staged) must agree bit for bit, from random data."""
    rng = np.random.default_rng(2024)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    outs = []
    for smoother in (mg.MG_SMOOTHER_COLOUR, mg.MG_SMOOTHER_AUTO):
        eng = mg.MultiGrid3D(n, RANGES[1], dtype=dtype, residual_mode=mg.MG_CORRECTED)
        eng.set_smoother(smoother, 1)
        eng.set_v(0, v0)
        eng.set_f(0, f0)
        eng.VCycle(0, 2, 1)
        outs.append(eng.get_v(0))
        eng.close()
    assert_bits_equal(outs[0], outs[1], "smoother variants")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu2", [(65, 2), (257, 1), (257, 2)])
def test_colour1_only_correction_is_exact(mg, monkeypatch, n, nu2, dtype):
    """Inside a V-cycle the engine corrects only the colour-1 points after the prolongation (the red half-sweep that
    follows overwrites every interior colour-0 point unread).  MG_B200_FULL_CORRECTION=1 switches that off: both
    engines must leave the same bits on every level."""
    rng = np.random.default_rng(4242)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    outs = []
    for full in (False, True):
        if full:
            monkeypatch.setenv("MG_B200_FULL_CORRECTION", "1")
        else:
            monkeypatch.delenv("MG_B200_FULL_CORRECTION", raising=False)
        eng = mg.MultiGrid3D(n, RANGES[1], dtype=dtype, residual_mode=mg.MG_CORRECTED)
        eng.set_v(0, v0)
        eng.set_f(0, f0)
        for _ in range(3):  # eager, captured, replayed
            eng.VCycle(0, 2, nu2)
        outs.append([eng.get_v(l) for l in range(eng.numGrids)])
        eng.close()
    for l, (a, b) in enumerate(zip(*outs)):
        assert_bits_equal(a, b, "level %d" % l)


# ---- temporally blocked smoother (two sweeps per HBM pass) ----------------------------------------

@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu", [(257, 2), (257, 3), (513, 4)])
def test_fused_smoother_is_bit_identical(mg, n, nu, dtype, rng_range):
    """MG_SMOOTHER_FUSED (two RB sweeps in one pass, out of place) against the one-colour-per-launch kernel,
    from random data; nu = 3 exercises the fused pass followed by a two-pass remainder sweep."""
    rng = np.random.default_rng(77)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    outs = []
    for smoother in (mg.MG_SMOOTHER_COLOUR, mg.MG_SMOOTHER_FUSED):
        eng = mg.MultiGrid3D(n, rng_range, dtype=dtype)
        eng.set_smoother(smoother)
        eng.set_v(0, v0)
        eng.set_f(0, f0)
        eng.Relax(0, nu)
        outs.append(eng.get_v(0))
        eng.close()
    assert_bits_equal(outs[0], outs[1], "fused smoother")


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_smoother_vcycle_vs_oracle(mg, dtype):
    n = 257
    eng = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    eng.set_smoother(mg.MG_SMOOTHER_FUSED, 2)
    o = oracles(3, dtype, True, n)[0]
    for _ in range(3):  # the third call replays the CUDA graph captured by the second
        eng.VCycle(0, 2, 2)
        o.vcycle(0, 2, 2)
    for l in range(eng.numGrids):
        assert_bits_equal(eng.get_v(l), o.v(l), "v level %d after 3 V(2,2) with the fused smoother" % l)
    eng.close()
