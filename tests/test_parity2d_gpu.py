"""GPU parity of the 2D Lyapunov path (MultiGrid2D) against the CPU oracle, bit-exact, via the C ABI.
K1, K2 and the denominator vary per point, so this path is FMA-sensitive: bit equality here proves the
kernels never contract."""
import numpy as np
import pytest

from util import assert_bits_equal, oracles, random_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
RANGES = [(0, 1, 0, 1), (0, 20, 0, 20), (0.3, 1.9, -0.4, 2.2)]
A_SHIPPED = (-1.0, -2.0, 0.0, -3.0)  # N2/LyapunovSolver.cpp:19-22
A_OTHER = (-0.7, 1.3, 0.45, -2.1)


def _pair(mg, n, dtype, rng_range, A=A_SHIPPED, alfa=2, seed=4321):
    eng = mg.MultiGrid2D(n, rng_range, A=A, alfa=alfa, dtype=dtype)
    orcs = oracles(2, dtype, False, n, range=rng_range, A=A, alfa=alfa)
    rng = np.random.default_rng(seed)
    v0 = random_field(rng, (n, n), dtype)
    f0 = random_field(rng, (n, n), dtype)
    eng.set_v(0, v0)
    eng.set_f(0, f0)
    for o in orcs:
        o.v(0)[...] = v0
        o.f(0)[...] = f0
    return eng, orcs, v0, f0


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 17, 65, 257])
def test_init_problem(mg, n, dtype, rng_range):
    eng = mg.MultiGrid2D(n, rng_range, dtype=dtype)
    for o in oracles(2, dtype, False, n, range=rng_range):
        for l in range(eng.numGrids):
            assert eng.level_size(l) == o.sizes[l]
            assert_bits_equal(eng.get_v(l), o.v(l), "InitV level %d" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "InitF level %d" % l)
    eng.close()


@pytest.mark.parametrize("A,alfa", [(A_SHIPPED, 2), (A_OTHER, 3)])
@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 9, 33, 129, 513])
def test_relax_residual(mg, n, dtype, rng_range, A, alfa):
    eng, orcs, _, _ = _pair(mg, n, dtype, rng_range, A, alfa)
    eng.Relax(0, 3)
    got_v = eng.get_v(0)
    got_r = eng.CalculateResidual(0)
    l2, linf = eng.residual_norm(0)
    for o in orcs:
        o.relax(0, 3)
        assert_bits_equal(got_v, o.v(0), "Relax")
        ro = o.residual(0)
        assert_bits_equal(got_r, ro, "CalculateResidual")
        ol2, olinf = o.residual_norms(0)
        assert abs(l2 - ol2) <= 1e-12 * max(ol2, 1e-300) and linf == olinf
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [5, 9, 65, 257, 1025])
def test_host_ops(mg, n, dtype):
    eng = mg.MultiGrid2D(5, dtype=dtype)
    orcs = oracles(2, dtype, False, 5)
    rng = np.random.default_rng(11)
    fine = random_field(rng, (n, n), dtype)
    fine2 = random_field(rng, (n, n), dtype)
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


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [5, 17, 129, 513])
def test_fused_level_ops(mg, n, dtype):
    eng, orcs, v0, f0 = _pair(mg, n, dtype, RANGES[2], A_OTHER, 3)
    rng = np.random.default_rng(5)
    cn = (n - 1) // 2 + 1
    cv = random_field(rng, (cn, cn), dtype)
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
        e = np.zeros((n, n), dtype)
        o.interpolate(e, cv)
        assert_bits_equal(got_v, o.apply_correction(v0.copy(), e), "fused interpolate+correct")
        assert_bits_equal(got_vi, o.interpolate(got_v.copy(), cv), "Interpolate into v")
        assert_bits_equal(got_rf, o.restrict(f0), "Restrict(f)")
    eng.close()


@pytest.mark.parametrize("rng_range", RANGES[:2])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu", [(65, 2), (257, 2), (65, 50)])
def test_vcycle_history(mg, n, nu, dtype, rng_range):
    eng = mg.MultiGrid2D(n, rng_range, dtype=dtype)
    orcs = oracles(2, dtype, False, n, range=rng_range)
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
            assert_bits_equal(eng.get_v(l), o.v(l), "v level %d" % l)
            assert_bits_equal(eng.get_f(l), o.f(l), "f level %d" % l)
        for (a, am), (b, bm) in zip(hist, ohist):
            assert abs(a - b) <= 1e-5 * abs(b) if dtype == np.float32 else abs(a - b) <= 1e-10 * abs(b)
            assert am == bm
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_vcycle_full_size_config(mg, dtype):
    """BASELINE.json configs[1] at its full size, 1025 x 1025, V(2,2): four cycles (eager, captured, two graph
    replays) bit for bit against the oracle on every level, and the SURVEY.md 8c history of the reference
    (9.637410504e+04 -> 2.626493724e+04, 1.122402051e+04, 8.076382287e+03 in float)."""
    n = 1025
    eng = mg.MultiGrid2D(n, dtype=dtype)
    orcs = oracles(2, dtype, False, n)
    hist = [eng.residual_norm(0)[0]]
    for _ in range(4):
        eng.VCycle(0, 2, 2)
        hist.append(eng.residual_norm(0)[0])
    for o in orcs:
        for _ in range(4):
            o.vcycle(0, 2, 2)
        for l in range(eng.numGrids):
            assert_bits_equal(eng.get_v(l), o.v(l), "v level %d" % l)
    want = [9.637410504e+04, 2.626493724e+04, 1.122402051e+04, 8.076382287e+03]
    for a, b in zip(hist, want):
        assert abs(a - b) <= 2e-5 * b, (hist, want)
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_fmg_and_mean_abs_error(mg, dtype):
    """FMG with the thesis parameters at n = 65 on [0,20]^2: the known answer of thesis Fig. 4.3
    (mean absolute error 5.32 at n = 65) and bit equality with the oracle."""
    n = 65
    eng = mg.MultiGrid2D(n, (0, 20, 0, 20), dtype=dtype)
    eng.FullMultiGridVCycle(0, 2, 500, 500)
    got = eng.get_v(0)
    mae = eng.mean_abs_error()
    for o in oracles(2, dtype, False, n, range=(0, 20, 0, 20)):
        o.fmg(0, 2, 500, 500)
        assert_bits_equal(got, o.v(0), "FMG(2,500,500) v")
    assert abs(mae - 5.32) < 0.01, mae
    eng.close()


def test_vcycle_host(mg):
    n, dtype = 129, np.float32
    eng = mg.MultiGrid2D(n, dtype=dtype)
    o = oracles(2, dtype, False, n)[0]
    v = o.v(0).copy()
    f = o.f(0).copy()
    eng.vcycle_host(v, f, 2, 2, cycles=2)
    o.vcycle(0, 2, 2)
    o.vcycle(0, 2, 2)
    assert_bits_equal(v, o.v(0), "vcycle_host")
    eng.close()


def test_argument_errors(mg):
    with pytest.raises(mg.MGError):
        mg.MultiGrid2D([17, 9])
    with pytest.raises(mg.MGError):
        mg.MultiGrid2D(20)
    eng = mg.MultiGrid2D(9)
    with pytest.raises(mg.MGError):
        eng.Relax(5, 1)
    eng.close()
