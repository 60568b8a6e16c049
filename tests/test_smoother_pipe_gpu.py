"""The register-tiled temporally blocked smoother (mg3d_smooth_pipe.cu, MG_SMOOTHER_PIPE, the default on large levels)
and the hashes that pin the BASELINE sizes to the reference itself.

  * bit-identity of two sweeps per HBM pass with the plain one-colour-per-launch kernel from random data (float and
    double), odd sweep counts, several passes, coarse TMA levels;
  * the exactness guard: fields holding tiny, huge, negative-zero and non-finite values take the literal-arithmetic
    fallback and still equal the plain kernel bit for bit;
  * MG_ARITH_FAST within the north-star tolerance (1e-10 relative fp64, 1e-5 fp32) of the exact result;
  * `tests/golden/hashes3d.json` (SHA-256 and additive checksum of v on every level after V(2,2) cycles, produced by the
    compiled reference, tests/golden/make_hash.py): the engine's bits at 33^3 ... 513^3 (1025^3: test_full_size_gpu.py).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, field_checksum
from util import assert_bits_equal, oracles, random_field

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]
UNIT = (0, 1, 0, 1, 0, 1)


def _relaxed(mg, smoother, n, dtype, v0, f0, nu, level=0, arith=None, rng_range=UNIT):
    eng = mg.MultiGrid3D(n, rng_range, dtype=dtype)
    eng.set_smoother(smoother)
    if arith is not None:
        eng.set_arith(arith)
    eng.set_v(level, v0)
    eng.set_f(level, f0)
    eng.Relax(level, nu)
    out = eng.get_v(level)
    eng.close()
    return out


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n,nu", [(257, 2), (257, 3), (257, 4), (513, 2), (513, 5)])
def test_pipe_smoother_is_bit_identical(mg, n, nu, dtype):
    rng = np.random.default_rng(1000 + n + nu)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    want = _relaxed(mg, mg.MG_SMOOTHER_COLOUR, n, dtype, v0, f0, nu)
    got = _relaxed(mg, mg.MG_SMOOTHER_PIPE, n, dtype, v0, f0, nu)
    assert_bits_equal(got, want, "pipelined smoother, n=%d nu=%d" % (n, nu))
    auto = _relaxed(mg, mg.MG_SMOOTHER_AUTO, n, dtype, v0, f0, nu)
    assert_bits_equal(auto, want, "default smoother, n=%d nu=%d" % (n, nu))


@pytest.mark.parametrize("dtype", DTYPES)
def test_pipe_smoother_on_a_coarse_level_with_nonzero_boundary(mg, dtype):
    """Level 1 of a 513^3 hierarchy (257^3), Dirichlet values that are not zero (the colour-0 boundary points are the only
    colour-0 values the pass reads)."""
    n, level, nl = 513, 1, 257
    rng = np.random.default_rng(5)
    v0 = random_field(rng, (nl,) * 3, dtype)
    f0 = random_field(rng, (nl,) * 3, dtype)
    want = _relaxed(mg, mg.MG_SMOOTHER_COLOUR, n, dtype, v0, f0, 2, level=level)
    got = _relaxed(mg, mg.MG_SMOOTHER_PIPE, n, dtype, v0, f0, 2, level=level)
    assert_bits_equal(got, want, "pipelined smoother on level 1")
    assert_bits_equal(got[0], v0[0], "Dirichlet plane kept")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("poison", ["tiny", "huge", "negzero", "inf", "subnormal"])
def test_pipe_guard_falls_back_exactly(mg, dtype, poison):
    """Values outside the range in which the scaled formula is provably the reference's: the pass must notice and the
    literal-arithmetic pass behind it must produce the plain kernel's bits (also the answer to `div_by_const` near
    underflow: the plain kernels divide IEEE-exactly there)."""
    n = 257
    rng = np.random.default_rng(99)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    fi = np.finfo(dtype)
    val = {"tiny": fi.tiny * 64, "huge": fi.max / 4, "negzero": -0.0, "inf": np.inf, "subnormal": fi.tiny / 1024}[poison]
    idx = rng.integers(1, n - 1, size=(200, 3))
    for k, (z, y, x) in enumerate(idx):
        (v0 if k % 2 else f0)[z, y, x] = val if k % 3 else -val
    v0[n // 2, n // 2, 1:20] = val   # a run of neighbours: sums of nothing but poisoned values
    with np.errstate(all="ignore"):
        want = _relaxed(mg, mg.MG_SMOOTHER_COLOUR, n, dtype, v0, f0, 2)
        got = _relaxed(mg, mg.MG_SMOOTHER_PIPE, n, dtype, v0, f0, 2)
    assert_bits_equal(got, want, "guarded pass with %s values" % poison)


@pytest.mark.parametrize("dtype", DTYPES)
def test_relax_with_tiny_values_vs_reference(mg, dtype):
    """|numerator| near the underflow threshold through Relax against the compiled reference."""
    n = 33
    rng = np.random.default_rng(3)
    fi = np.finfo(dtype)
    scale = fi.tiny * 2.0 ** 8
    v0 = (random_field(rng, (n,) * 3, dtype) * scale).astype(dtype)
    f0 = (random_field(rng, (n,) * 3, dtype) * scale).astype(dtype)
    v0[5, 5, 5] = -0.0
    eng = mg.MultiGrid3D(n, UNIT, dtype=dtype)
    eng.set_v(0, v0)
    eng.set_f(0, f0)
    eng.Relax(0, 2)
    got = eng.get_v(0)
    eng.close()
    for o in oracles(3, dtype, False, n):
        o.v(0)[...] = v0
        o.f(0)[...] = f0
        o.relax(0, 2)
        assert_bits_equal(got, o.v(0), "Relax on near-underflow data vs %s" % type(o).__name__)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_fast_arithmetic_within_tolerance(mg, dtype, tol):
    n = 257
    exact = mg.MultiGrid3D(n, UNIT, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    fast = mg.MultiGrid3D(n, UNIT, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    fast.set_arith(mg.MG_ARITH_FAST)
    for cyc in range(3):
        exact.VCycle(0, 2, 2)
        fast.VCycle(0, 2, 2)
        a, b = exact.residual_norm(0)[0], fast.residual_norm(0)[0]
        # the residual amplifies rounding differences of v by 6/h^2 = 4e5: 1e-16 relative in v shows as ~1e-10 .. 1e-9 relative
        # in ||r||.  In float32 the residual sits on its round-off floor from the second cycle on (pure rounding noise in both
        # runs): only the first cycle is compared there.
        if dtype == np.float64:
            assert abs(a - b) <= 1e-8 * a, (cyc, a, b)
        elif cyc == 0:
            assert abs(a - b) <= 2e-2 * a, (cyc, a, b)
    va, vb = exact.get_v(0), fast.get_v(0)
    assert np.max(np.abs(va.astype(np.float64) - vb)) <= tol * np.max(np.abs(va))
    assert not np.array_equal(va, vb) or dtype == np.float32  # it IS a different rounding sequence
    exact.close()
    fast.close()


def test_odd_pass_counts_and_graph_replay_keep_track_of_the_buffers(mg):
    """a captured V-cycle bakes in which v buffer each level uses.  Interleave replays with calls that flip
    the buffer an odd number of times (Relax(0, 2), V(2,1)) and compare with an engine that never captures."""
    n, dtype = 257, np.float64
    rng = np.random.default_rng(8)
    v0 = random_field(rng, (n,) * 3, dtype)
    f0 = random_field(rng, (n,) * 3, dtype)
    outs = []
    for smoother in (mg.MG_SMOOTHER_COLOUR, mg.MG_SMOOTHER_AUTO):
        eng = mg.MultiGrid3D(n, UNIT, dtype=dtype, residual_mode=mg.MG_CORRECTED)
        eng.set_smoother(smoother)
        eng.set_v(0, v0)
        eng.set_f(0, f0)
        for _ in range(3):
            eng.VCycle(0, 2, 2)
        eng.Relax(0, 2)          # one pass: v now lives in the other buffer
        for _ in range(3):
            eng.VCycle(0, 2, 2)  # must not replay the graph captured on the first buffer
        for _ in range(4):
            eng.VCycle(0, 2, 1)  # three passes per cycle on level 0: every replay ends on the other buffer
        eng.FullMultiGridVCycle(0, 1, 2, 1)
        outs.append([eng.get_v(l) for l in range(eng.numGrids)])
        eng.close()
    for l, (a, b) in enumerate(zip(*outs)):
        assert_bits_equal(b, a, "level %d" % l)


# ---- hashes of the reference's own runs ----------------------------------------------------------------------------

with open(os.path.join(GOLDEN_DIR, "hashes3d.json")) as _fh:
    HASHES = {k: v for k, v in json.load(_fh).items() if not k.startswith("_")}


def _sha(a):
    return hashlib.sha256(memoryview(np.ascontiguousarray(a).reshape(-1).view(np.uint8))).hexdigest()


@pytest.mark.parametrize("key", sorted(k for k in HASHES if HASHES[k]["n"] <= 513))
def test_vcycle_bits_equal_the_reference_run(mg, key):
    rec = HASHES[key]
    n = rec["n"]
    dtype = np.float64 if rec["dtype"] == "f64" else np.float32
    mode = mg.MG_CORRECTED if rec["mode"] == "corrected" else mg.MG_REF_COMPAT
    eng = mg.MultiGrid3D(n, UNIT, dtype=dtype, residual_mode=mode)
    for c, ent in enumerate(rec["cycles"]):
        eng.VCycle(0, rec["v1"], rec["v2"])
        for l in range(eng.numGrids):
            assert "%016x" % eng.field_checksum(l) == ent["checksum_v"][l], "checksum, cycle %d level %d" % (c + 1, l)
        if n <= 257 or c == len(rec["cycles"]) - 1:
            for l in range(eng.numGrids):
                assert _sha(eng.get_v(l)) == ent["sha256_v"][l], "sha256, cycle %d level %d" % (c + 1, l)
        if "l2" in ent:
            l2, linf = eng.residual_norm(0)
            assert abs(l2 - float(ent["l2"])) <= 1e-10 * float(ent["l2"]) and linf == float(ent["linf"])
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_device_checksum_is_the_numpy_checksum(mg, dtype):
    n = 65
    rng = np.random.default_rng(11)
    eng = mg.MultiGrid3D(n, UNIT, dtype=dtype)
    v0 = random_field(rng, (n,) * 3, dtype)
    v0[3, 4, 5] = -0.0
    eng.set_v(0, v0)
    assert eng.field_checksum(0) == field_checksum(v0)
    assert eng.field_checksum(0, mg.MG_FIELD_F) == field_checksum(eng.get_f(0))
    v0[10, 11, 12], v0[10, 11, 13] = v0[10, 11, 13], v0[10, 11, 12]  # a swap must change it
    assert eng.field_checksum(0) != field_checksum(v0)
    eng.close()


@pytest.mark.parametrize("dtype", DTYPES)
def test_abs_error_is_printdiff_reduced(mg, dtype):
    """mg3d_abs_error against Grid3D::PrintDiff's arithmetic (N3/Grid3D.cpp:146-152) restated in numpy."""
    n = 65
    eng = mg.MultiGrid3D(n, UNIT, dtype=dtype, residual_mode=mg.MG_CORRECTED)
    for _ in range(4):
        eng.VCycle(0, 2, 2)
    v = eng.get_v(0)
    h = dtype(1.0) / dtype(n - 1)
    x = (dtype(0.0) + np.arange(n).astype(dtype) * h).astype(dtype)  # float x = x_a + posX*h_x
    s = np.sin(np.float64(3.141592653589793) * x.astype(np.float64))
    real = (s[None, None, :] * s[None, :, None] * s[:, None, None]).astype(dtype)
    diff = np.abs((real - v).astype(np.float64))
    mean, mx = eng.abs_error(0)
    assert mx == diff.max()
    assert abs(mean - diff.mean()) <= 1e-12 * diff.mean()
    assert mx < 2.5e-4  # discretisation error of the 65^3 grid (2.0e-4)
    eng.close()
