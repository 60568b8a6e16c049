"""Weighted-Jacobi smoother (MG_SMOOTHER_JACOBI, 3D).  The reference has no Jacobi smoother, so
there is no reference parity to claim: the C restatement oracle/mg_oracle_impl.h::orc3d_relax_jacobi DEFINES it,
is checked here against an independent numpy statement of the same arithmetic, and the CUDA path is checked
bit-for-bit against it."""
import numpy as np
import pytest

from oracle import port
from util import assert_bits_equal, random_field

RANGES = [(0.0, 1.0, 0.0, 1.0, 0.0, 1.0), (0.0, 1.5, -0.25, 0.5, 1.0, 3.0)]  # the second has h not a power of two
DTYPES = [np.float32, np.float64]


def numpy_jacobi(v, f, rng_range, omega, sweeps):
    """One array statement of the sweep; same operation order as the C loop, so it must agree to the bit."""
    T = v.dtype.type
    n = v.shape[0]
    r = rng_range
    hx, hy, hz = [(T(r[2 * a + 1]) - T(r[2 * a])) / T(n - 1) for a in range(3)]
    hx2, hy2, hz2 = hx * hx, hy * hy, hz * hz
    w = T(omega)
    v = v.copy()
    for _ in range(sweeps):
        c = v[1:-1, 1:-1, 1:-1]  # arrays are [z, y, x]
        O, E = v[1:-1, 1:-1, :-2], v[1:-1, 1:-1, 2:]
        N, S = v[1:-1, :-2, 1:-1], v[1:-1, 2:, 1:-1]
        D, U = v[:-2, 1:-1, 1:-1], v[2:, 1:-1, 1:-1]
        num = (O * (hy2 * hz2) + E * (hy2 * hz2) + N * (hx2 * hz2) + S * (hx2 * hz2) + D * (hx2 * hy2) + U * (hx2 * hy2)
               - f[1:-1, 1:-1, 1:-1] * hx2 * hy2 * hz2)
        gs = num / (T(2) * (hy2 * hz2 + hx2 * hz2 + hx2 * hy2))
        nv = v.copy()
        nv[1:-1, 1:-1, 1:-1] = c + w * (gs - c)
        v = nv
    return v


@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [3, 5, 9, 17])
def test_port_jacobi_equals_numpy_statement(n, dtype, rng_range):
    rng = np.random.default_rng(12345)
    o = port.PortMG(3, dtype, True, n=max(n, 3), range=rng_range)
    o._v[0][...] = random_field(rng, o.shape(0), dtype)
    o._f[0][...] = random_field(rng, o.shape(0), dtype)
    want = numpy_jacobi(o.v(0), o.f(0), rng_range, 6.0 / 7.0, 3)
    o.relax_jacobi(0, 3)
    assert_bits_equal(o.v(0), want, "jacobi n=%d" % n)


def test_port_jacobi_vcycle_converges():
    o = port.PortMG(3, np.float64, True, n=33)
    r0 = o.residual_norms(0)[0]
    hist = []
    for _ in range(5):
        o.vcycle_jacobi(0, 2, 2)
        hist.append(o.residual_norms(0)[0])
    rates = [b / a for a, b in zip([r0] + hist[:-1], hist)]
    assert max(rates) < 0.35, rates  # ~0.27 per V(2,2) with omega = 6/7 (red-black Gauss-Seidel: ~0.12)


# ----------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("rng_range", RANGES)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,sweeps", [(3, 1), (5, 2), (9, 3), (33, 1), (65, 4), (129, 3)])
def test_gpu_jacobi_relax_bit_exact(mg, n, sweeps, dtype, rng_range):
    rng = np.random.default_rng(777 + n)
    o = port.PortMG(3, dtype, True, n=n, range=rng_range)
    e = mg.MultiGrid3D(n, dtype=dtype, range=rng_range, residual_mode=mg.MG_CORRECTED)
    e.set_smoother(mg.MG_SMOOTHER_JACOBI)
    v0, f0 = random_field(rng, o.shape(0), dtype), random_field(rng, o.shape(0), dtype)
    o._v[0][...] = v0
    o._f[0][...] = f0
    e.set_v(0, v0)
    e.set_f(0, f0)
    for omega in (6.0 / 7.0, 0.5):
        e.set_jacobi_weight(omega)
        o.relax_jacobi(0, sweeps, omega)
        e.Relax(0, sweeps)
        assert_bits_equal(e.get_v(0), o.v(0), "jacobi relax n=%d omega=%g" % (n, omega))
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu1,nu2", [(17, 2, 2), (65, 2, 1), (129, 3, 3)])
def test_gpu_jacobi_vcycle_bit_exact(mg, n, nu1, nu2, dtype):
    o = port.PortMG(3, dtype, True, n=n)
    e = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    e.set_smoother(mg.MG_SMOOTHER_JACOBI)
    for cyc in range(3):  # the third cycle replays the captured graph
        o.vcycle_jacobi(0, nu1, nu2)
        e.VCycle(0, nu1, nu2)
        for l in range(e.numGrids):
            assert_bits_equal(e.get_v(l), o.v(l), "jacobi V-cycle %d level %d" % (cyc, l))
    l2 = e.residual_norm(0)[0]
    assert abs(l2 - o.residual_norms(0)[0]) <= 1e-10 * l2  # device reduction order != sequential sum
    e.close()


@pytest.mark.gpu
def test_gpu_jacobi_large_level_and_weight_check(mg):
    """257^3: the levels that otherwise take the TMA kernels; and the argument check of the weight."""
    n = 257
    o = port.PortMG(3, np.float64, True, n=n)
    e = mg.MultiGrid3D(n, dtype=np.float64, residual_mode=mg.MG_CORRECTED)
    e.set_smoother(mg.MG_SMOOTHER_JACOBI)
    o.vcycle_jacobi(0, 2, 2)
    e.VCycle(0, 2, 2)
    assert_bits_equal(e.get_v(0), o.v(0), "jacobi V(2,2) 257^3")
    with pytest.raises(mg.MGError):
        e.set_jacobi_weight(0.0)
    with pytest.raises(mg.MGError):
        e.set_jacobi_weight(1.5)
    e.close()
