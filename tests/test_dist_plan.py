"""CPU: the z-slab plan of the multi-GPU 3D driver (pure arithmetic, no GPU): slabs tile the grid exactly,
nest over the levels, carry the ghost planes the kernels need, and switch to agglomerated levels below the
threshold."""
import pytest


@pytest.mark.parametrize("n,P", [(1025, 2), (1025, 4), (1025, 8), (2049, 8), (513, 2), (257, 8), (129, 8), (65, 8)])
def test_slabs_tile_and_nest(mg, n, P):
    plan = mg.MultiGrid3D.plan_level
    levels = []
    s = n
    while s >= 3:
        levels.append(s)
        if s == 3:
            break
        s = (s - 1) // 2 + 1
    prev = None
    for nl in levels:
        plans = [plan(nl, P, r) for r in range(P)]
        dist = plans[0]["dist"]
        assert all(p["dist"] == dist for p in plans)
        if not dist:
            assert all(p["z0"] == 0 and p["nzl"] == nl and p["own_lo"] == 0 and p["own_hi"] == nl for p in plans)
            assert (nl - 1) // P < 8 or nl < 257
        else:
            owned = []
            for r, p in enumerate(plans):
                a, b = p["z0"] + p["own_lo"], p["z0"] + p["own_hi"]
                owned.append((a, b))
                assert p["own_lo"] == (4 if r > 0 else 0)              # four ghost planes below ...
                assert p["nzl"] - p["own_hi"] == (4 if r < P - 1 else 0)  # ... and above (two RB sweeps per smoother pass)
                assert b - a >= 8
            assert owned[0][0] == 0 and owned[-1][1] == nl
            assert all(owned[r][1] == owned[r + 1][0] for r in range(P - 1))
            if prev is not None:  # nesting: the coarse slab starts at half the fine slab's start
                assert all(prev[r][0] == 2 * owned[r][0] for r in range(P))
        prev = owned if dist else None


def test_single_gpu_plan(mg):
    p = mg.MultiGrid3D.plan_level(1025, 1, 0)
    assert p == {"dist": 0, "z0": 0, "nzl": 1025, "own_lo": 0, "own_hi": 1025}
