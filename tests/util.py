"""Shared helpers of the test-suite (checker side only: the oracle never runs in the product)."""
import numpy as np

from oracle import port, ref


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, what
    if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        diff = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
        first = tuple(diff[0]) if len(diff) else None
        da = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError("%s: %d of %d values differ (first at %s: %r vs %r, max |diff| %.3e)" % (
            what, len(diff), a.size, first, a[first] if first else None, b[first] if first else None,
            np.nanmax(da) if da.size else 0.0))


def oracles(dim, dtype, corrected, n, range=None, **kw):
    """The CPU checkers for one configuration: the C restatement always, the compiled reference
    when oracle/_ref is present."""
    out = [port.PortMG(dim, dtype, corrected, n=n, range=range, **kw)]
    if ref.available():
        out.append(ref.RefMG(dim, dtype, corrected, n=n, range=range, **kw))
    return out


def random_field(rng, shape, dtype):
    return rng.uniform(-1.0, 1.0, size=shape).astype(dtype)
