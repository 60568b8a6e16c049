"""GPU parity of the 3D Poisson path on NON-CUBIC grids (mg3b_* of the C ABI).

Checkers: the reference itself compiled with its assertions off (oracle/_ref, ref3d_*x: the only thing between the reference
and such grids is the pair of asserts at N3/Grid3D.cpp:10-11) and the plain-C restatement orc3b_* (pinned to it bit for bit by
tests/test_oracle.py).  Bar: bit-exact (float and double, both residual modes, power-of-two and general ranges); residual norms
within 1e-12 relative (device reduction order).  A cubic shape must give the bits of the tuned cubic engine."""
import numpy as np
import pytest

from oracle import port, ref
from util import assert_bits_equal, random_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
SHAPES = [(33, 17, 9), (17, 33, 65), (65, 33, 33), (9, 9, 33), (5, 3, 9), (129, 65, 33)]  # (sizeX, sizeY, sizeZ)
RANGES = [(0, 1, 0, 1, 0, 1), (0.25, 1.75, -0.5, 0.7, 0.1, 3.3)]


def checkers(dtype, corrected, shape, rng_range):
    out = [port.PortBox3D(dtype, corrected, shape=shape, range=rng_range)]
    if ref.available():
        out.append(ref.RefMG(3, dtype, corrected, shape=shape, range=rng_range))
    return out


def pair(mg, shape, dtype, corrected, rng_range, seed=4711):
    eng = mg.MultiGrid3DBox(shape, rng_range, dtype=dtype, residual_mode=mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT)
    orcs = checkers(dtype, corrected, shape, rng_range)
    rng = np.random.default_rng(seed)
    v0 = random_field(rng, eng.shape(0), dtype)
    f0 = random_field(rng, eng.shape(0), dtype)
    eng.set_v(0, v0)
    eng.set_f(0, f0)
    for o in orcs:
        o.v(0)[...] = v0
        o.f(0)[...] = f0
    return eng, orcs


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
def test_hierarchy_and_init(mg, shape, dtype, rng_range):
    eng = mg.MultiGrid3DBox(shape, rng_range, dtype=dtype)
    for o in checkers(dtype, False, shape, rng_range):
        assert eng.numGrids == o.num_levels
        for l in range(eng.numGrids):
            assert eng.shape(l) == tuple(o.shape(l))
            assert_bits_equal(eng.get_f(l), o.f(l), "InitF level %d" % l)
            assert_bits_equal(eng.get_v(l), o.v(l), "InitV level %d" % l)
    eng.close()


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES[:4])
def test_relax_and_residual(mg, shape, dtype, rng_range):
    for corrected in (False, True):
        eng, orcs = pair(mg, shape, dtype, corrected, rng_range)
        eng.Relax(0, 3)
        got_v, got_r = eng.get_v(0), eng.CalculateResidual(0)
        l2, linf = eng.residual_norm(0)
        for o in orcs:
            o.relax(0, 3)
            assert_bits_equal(got_v, o.v(0), "Relax")
            r = o.residual(0)
            assert_bits_equal(got_r, r, "CalculateResidual")
            ol2, olinf = o.residual_norms(0)
            assert abs(l2 - ol2) <= 1e-12 * abs(ol2) and linf == olinf
        eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES[:4])
def test_level_operators(mg, shape, dtype):
    """Restrict(f), residual+restrict (+ zeroed coarse v), Interpolate, Interpolate + ApplyCorrection on the level arrays"""
    eng, orcs = pair(mg, shape, dtype, True, RANGES[0])
    rng = np.random.default_rng(99)
    cv = random_field(rng, eng.shape(1), dtype)
    eng.restrict_level(0)
    got_cf = eng.get_f(1)
    eng.set_v(1, cv)
    eng.interpolate_correct(0)
    got_v = eng.get_v(0)
    eng.residual_restrict(0)
    got_cf2, got_cv2 = eng.get_f(1), eng.get_v(1)
    eng.set_v(1, cv)
    eng.interpolate_level(0)
    got_vi = eng.get_v(0)
    o = orcs[0]  # the restatement exposes the operators on free arrays
    assert_bits_equal(got_cf, o.restrict(o.f(0)), "Restrict(f)")
    tmp = np.zeros_like(o.v(0))
    o.interpolate(tmp, cv)
    v1 = o.v(0).copy()
    v1[1:-1, 1:-1, 1:-1] += tmp[1:-1, 1:-1, 1:-1]
    assert_bits_equal(got_v, v1, "Interpolate + ApplyCorrection")
    o.v(0)[...] = v1
    assert_bits_equal(got_cf2, o.restrict(o.residual(0)), "Restrict(CalculateResidual)")
    assert not got_cv2.any()
    vi = v1.copy()
    o.interpolate(vi, cv)
    assert_bits_equal(got_vi, vi, "Interpolate (interior only)")
    eng.close()


@pytest.mark.parametrize("corrected", [False, True])
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES)
def test_vcycles(mg, shape, dtype, corrected):
    """V(2,2), V(2,1), V(0,3) from the reference problem: v and f of every level, bit for bit"""
    eng = mg.MultiGrid3DBox(shape, RANGES[0], dtype=dtype, residual_mode=mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT)
    orcs = checkers(dtype, corrected, shape, RANGES[0])
    for (v1, v2) in ((2, 2), (2, 1), (0, 3)):
        eng.VCycle(0, v1, v2)
        l2, _ = eng.residual_norm(0)
        for o in orcs:
            o.vcycle(0, v1, v2)
            for l in range(eng.numGrids):
                assert_bits_equal(eng.get_v(l), o.v(l), "v level %d after V(%d,%d)" % (l, v1, v2))
                assert_bits_equal(eng.get_f(l), o.f(l), "f level %d after V(%d,%d)" % (l, v1, v2))
            ol2 = o.residual_norms(0)[0]
            assert abs(l2 - ol2) <= 1e-12 * abs(ol2) or not np.isfinite(ol2)
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", SHAPES[:5])
def test_fmg_and_general_range(mg, shape, dtype):
    for rng_range in RANGES:
        eng = mg.MultiGrid3DBox(shape, rng_range, dtype=dtype, residual_mode=mg.MG_CORRECTED)
        eng.FullMultiGridVCycle(0, 2, 2, 2)
        for o in checkers(dtype, True, shape, rng_range):
            o.fmg(0, 2, 2, 2)
            for l in range(eng.numGrids):
                assert_bits_equal(eng.get_v(l), o.v(l), "v level %d after FMG" % l)
        eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_cubic_shape_gives_the_cubic_engines_bits(mg, dtype):
    a = mg.MultiGrid3DBox((33, 33, 33), dtype=dtype, residual_mode=mg.MG_CORRECTED)
    b = mg.MultiGrid3D(33, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    for _ in range(2):
        a.VCycle(0, 2, 2)
        b.VCycle(0, 2, 2)
    assert a.numGrids == b.numGrids
    for l in range(a.numGrids):
        assert_bits_equal(a.get_v(l), b.get_v(l), "level %d" % l)
    a.close()
    b.close()


def test_convergence_and_host_call(mg):
    """an anisotropic box converges (slower than the cube: its coarsest level keeps more than one unknown), and the host-array
    call returns what the resident cycle computes"""
    shape = (129, 65, 33)
    eng = mg.MultiGrid3DBox(shape, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    r0 = eng.residual_norm(0)[0]
    hv = np.zeros(eng.shape(0))
    hf = eng.get_f(0)
    eng.vcycle_host(hv, hf, 2, 2, 3)
    r3 = eng.residual_norm(0)[0]
    o = port.PortBox3D(np.float64, True, shape=shape)
    for _ in range(3):
        o.vcycle(0, 2, 2)
    want = o.residual_norms(0)[0]
    assert abs(r3 - want) <= 1e-12 * want and r3 < 0.25 * r0  # (the cube contracts by 0.12 per cycle, this box by 0.57)
    assert_bits_equal(hv, eng.get_v(0), "vcycle_host")
    eng2 = mg.MultiGrid3DBox(shape, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    for _ in range(3):
        eng2.VCycle(0, 2, 2)
    assert_bits_equal(hv, eng2.get_v(0), "host call vs resident cycles")
    eng.close()
    eng2.close()


def test_bad_arguments(mg):
    with pytest.raises(mg.MGError):
        mg.MultiGrid3DBox((33, 18, 9))      # not 2^k + 1
    with pytest.raises(mg.MGError):
        mg.MultiGrid3DBox((33, 17, 2))
    eng = mg.MultiGrid3DBox((9, 5, 17))
    with pytest.raises(mg.MGError):
        eng.Relax(7, 1)
    with pytest.raises(mg.MGError):
        eng.residual_restrict(eng.numGrids - 1)
    eng.close()
