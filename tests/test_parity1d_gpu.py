"""GPU parity of the 1D path (MultiGrid1D) against the CPU oracle, bit-exact, via the C ABI.
BASELINE.json configs[0] is this program at N = 1025 (the reference's own CPU-runnable case)."""
import numpy as np
import pytest

from util import assert_bits_equal, oracles, random_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
RANGES = [(0, 1), (0.25, 1.75)]


def _pair(mg, n, dtype, corrected, rng_range, seed=77):
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
    eng = mg.MultiGrid1D(n, rng_range, dtype=dtype, residual_mode=mode)
    orcs = oracles(1, dtype, corrected, n, range=rng_range)
    rng = np.random.default_rng(seed)
    v0 = random_field(rng, (n,), dtype)
    f0 = random_field(rng, (n,), dtype)
    eng.set_v(0, v0)
    eng.set_f(0, f0)
    for o in orcs:
        o.v(0)[...] = v0
        o.f(0)[...] = f0
    return eng, orcs, v0, f0


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 17, 1025, 8193])
def test_init_problem(mg, n, dtype, rng_range):
    eng = mg.MultiGrid1D(n, rng_range, dtype=dtype)
    for o in oracles(1, dtype, False, n, range=rng_range):
        for l in range(eng.numGrids):
            assert eng.level_size(l) == o.sizes[l]
            assert_bits_equal(eng.get_v(l), o.v(l), "InitV level %d" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "InitF level %d" % l)
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 9, 129, 1025, 4097])
def test_relax_residual(mg, n, dtype, rng_range, corrected):
    eng, orcs, _, _ = _pair(mg, n, dtype, corrected, rng_range)
    eng.Relax(0, 7)
    got_v = eng.get_v(0)
    got_r = eng.CalculateResidual(0)
    l2, linf = eng.residual_norm(0)
    for o in orcs:
        o.relax(0, 7)
        assert_bits_equal(got_v, o.v(0), "Relax")
        assert_bits_equal(got_r, o.residual(0), "CalculateResidual")
        ol2, olinf = o.residual_norms(0)
        assert abs(l2 - ol2) <= 1e-12 * max(ol2, 1e-300) and linf == olinf
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [5, 9, 1025, 65537])
def test_host_ops(mg, n, dtype):
    eng = mg.MultiGrid1D(5, dtype=dtype)
    orcs = oracles(1, dtype, False, 5)
    rng = np.random.default_rng(3)
    fine = random_field(rng, (n,), dtype)
    fine2 = random_field(rng, (n,), dtype)
    coarse = eng.Restrict(fine)
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
@pytest.mark.parametrize("n", [5, 129, 1025])
def test_fused_level_ops(mg, n, dtype, corrected):
    eng, orcs, v0, f0 = _pair(mg, n, dtype, corrected, RANGES[1])
    rng = np.random.default_rng(8)
    cn = (n - 1) // 2 + 1
    cv = random_field(rng, (cn,), dtype)
    eng.set_v(1, cv)
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
        assert not got_cv.any()
        e = np.zeros((n,), dtype)
        o.interpolate(e, cv)
        assert_bits_equal(got_v, o.apply_correction(v0.copy(), e), "fused interpolate+correct")
        assert_bits_equal(got_vi, o.interpolate(got_v.copy(), cv), "Interpolate into v")
        assert_bits_equal(got_rf, o.restrict(f0), "Restrict(f)")
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu", [(1025, 1000), (1025, 2), (129, 100)])
def test_vcycle_history(mg, n, nu, dtype, corrected):
    """configs[0]: N = 1025 with the thesis' saturating nu = 1000 (and the non-converging V(2,2), App. B13)."""
    mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
    eng = mg.MultiGrid1D(n, dtype=dtype, residual_mode=mode)
    orcs = oracles(1, dtype, corrected, n)
    hist = [eng.residual_norm(0)]
    for _ in range(2):
        eng.VCycle(0, nu, nu)
        hist.append(eng.residual_norm(0))
    for o in orcs:
        ohist = [o.residual_norms(0)]
        for _ in range(2):
            o.vcycle(0, nu, nu)
            ohist.append(o.residual_norms(0))
        for l in range(eng.numGrids):
            assert_bits_equal(eng.get_v(l), o.v(l), "v level %d" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "f level %d" % l)
        for (a, am), (b, bm) in zip(hist, ohist):
            assert abs(a - b) <= 1e-10 * abs(b) and am == bm
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_fmg_thesis_parameters(mg, dtype):
    """FMG(2,1000,1000) at n = 1025 as in N1/Poisson1DSolver.cpp:13-25 (n reduced to configs[0])."""
    n = 1025
    eng = mg.MultiGrid1D(n, dtype=dtype)
    eng.FullMultiGridVCycle(0, 2, 1000, 1000)
    got = eng.get_v(0)
    for o in oracles(1, dtype, False, n):
        o.fmg(0, 2, 1000, 1000)
        assert_bits_equal(got, o.v(0), "FMG v")
    x = np.linspace(0, 1, n)
    exact = (np.exp(x) + x - 3) / (1 + np.exp(-x))
    assert np.max(np.abs(got - exact)) < 1.2e-3  # discretisation error 9.16e-4
    eng.close()


def test_vcycle_host(mg):
    n, dtype = 1025, np.float64
    eng = mg.MultiGrid1D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    o = oracles(1, dtype, True, n)[0]
    v = o.v(0).copy()
    f = o.f(0).copy()
    eng.vcycle_host(v, f, 100, 100, cycles=2)
    o.vcycle(0, 100, 100)
    o.vcycle(0, 100, 100)
    assert_bits_equal(v, o.v(0), "vcycle_host")
    eng.close()


def test_argument_errors(mg):
    with pytest.raises(mg.MGError):
        mg.MultiGrid1D(1000)
    with pytest.raises(mg.MGError):
        mg.MultiGrid1D(17, (1, 0))
    eng = mg.MultiGrid1D(17)
    with pytest.raises(mg.MGError):
        eng.Relax(4, 1)
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_abs_error_is_printdiffapproxreal_reduced(mg, dtype):
    """mg1d_abs_error against Grid1D::PrintDiffApproxReal's arithmetic (N1/Grid1D.cpp:46-60) restated with the same libm;
    SURVEY.md 4 quotes max|v-u| = 9.164e-4 for the float FMG(2,1000,1000) solve at n = 1025."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype, libm.expf.argtypes = ctypes.c_float, [ctypes.c_float]
    libm.exp.restype, libm.exp.argtypes = ctypes.c_double, [ctypes.c_double]
    n = 1025
    eng = mg.MultiGrid1D(n, dtype=dtype)
    eng.FullMultiGridVCycle(0, 2, 1000, 1000)
    v = eng.get_v(0)
    t = dtype
    h = t(1.0) / t(n - 1)
    real = np.empty(n, dtype=dtype)
    for j in range(n):
        xj = t(0.0) + t(j) * h
        if dtype == np.float32:
            real[j] = (t(libm.expf(xj)) + xj - t(3)) / (t(1) + t(libm.expf(-xj)))
        else:
            real[j] = (libm.exp(xj) + xj - 3) / (1 + libm.exp(-xj))
    diff = np.abs((v - real).astype(np.float64))
    mean, mx = eng.abs_error(0)
    assert mx == diff.max()
    assert abs(mean - diff.mean()) <= 1e-12 * diff.mean()
    assert abs(mx - 9.164e-4) < 5e-6
    eng.close()
