"""CPU, world_size 2, gloo: the multi-GPU z-slab SCHEDULE of the 3D driver (pde_multigrid_b200/csrc/mg3d_host.c).

Two processes hold the slabs that mg3d_plan_level (the product's own partition arithmetic, called through the
C ABI) assigns them; everything outside a slab is NaN.  They run V(2,2) cycles with numpy stand-ins for the
kernels, restricted to the planes a rank owns, and exchange planes over gloo exactly where the C driver does:
  * after every half-sweep: the just-updated colour's top plane up / bottom plane down (depth 1) -- ONLY the points
    of that colour travel, the other colour of the ghost plane keeps whatever it held,
  * before the fused residual+restrict: the two top planes up (the kernel reads v two planes below the slab),
  * after it: coarse f planes (distributed coarse level) or an all-gather (first agglomerated level),
  * after prolongation+correction: fine v, depth 1 both ways -- and, as in the engine's V-cycle, the correction is
    applied to the colour-1 points only and only colour 1 travels: the red half-sweep that follows overwrites every
    interior colour-0 point without reading it.
The weighted-Jacobi option runs through the same test with its own slab schedule (new colour 0 into a scratch array whose
roles alternate with v, colour 1 in place, per-colour exchanges of whichever array holds the new values, copy back of the
owned planes and the nearest ghosts after an odd number of sweeps) against the whole-grid definition of the sweep.
The temporally blocked smoother (smoother "pipe", the default on the large levels) has its own schedule: ONE exchange of four
planes of colour 1 in each direction in front of every two-sweep pass, the halo planes swept redundantly (R1 on [a-3, b+3),
B1 on [a-2, b+2), R2 on [a-1, b+1), B2 on [a, b)), only the owned planes written (to the other v buffer, whose remaining planes
are NaN here), ghost planes of v refreshed lazily by whoever reads them next (residual+restrict: two planes up, one down;
prolongation: the coarse level), coarse f exchanged four planes deep.
If a ghost depth or an exchange were missing, a NaN (or a stale value) would reach an owned plane.  The owned planes
must equal the same cycles run sequentially on the whole grid WITH THE FULL CORRECTION, bit for bit (RB Gauss-Seidel
is partition-invariant, and the colour-0 half of the correction is dead)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N = 65
NU = 2
NU_J = 3  # odd: every Jacobi smoothing call ends on the scratch array and copies back
NU_P = 3  # temporally blocked smoother: one two-sweep pass plus a colour-by-colour remainder sweep
PIPE_MIN_N = 33  # levels the emulated temporally blocked smoother takes (the engine: levels with tensor maps, n >= 257)
CYCLES = 2


def sizes(n):
    out = [n]
    while out[-1] > 3:
        out.append((out[-1] - 1) // 2 + 1)
    return out


# ---- numpy stand-ins for the kernels (z, y, x axes), acting on planes [lo, hi) only ----------------

def relax_colour(v, f, colour, lo, hi):
    n = v.shape[1]
    h2 = (1.0 / (n - 1)) ** 2
    z, y, x = np.meshgrid(np.arange(lo, hi), np.arange(1, n - 1), np.arange(1, n - 1), indexing="ij")
    mask = ((x + y + z) % 2) == colour
    s = (v[lo:hi, 1:-1, :-2] + v[lo:hi, 1:-1, 2:] + v[lo:hi, :-2, 1:-1] + v[lo:hi, 2:, 1:-1] +
         v[lo - 1:hi - 1, 1:-1, 1:-1] + v[lo + 1:hi + 1, 1:-1, 1:-1] - f[lo:hi, 1:-1, 1:-1] * h2) / 6.0
    blk = v[lo:hi, 1:-1, 1:-1]
    blk[mask] = s[mask]


def residual_plane(v, f, z):
    n = v.shape[1]
    r = np.zeros((n, n))
    if z < 1 or z > n - 2:
        return r
    h2 = (1.0 / (n - 1)) ** 2
    c = v[z, 1:-1, 1:-1]
    r[1:-1, 1:-1] = f[z, 1:-1, 1:-1] - (v[z, 1:-1, :-2] - 2 * c + v[z, 1:-1, 2:]) / h2 \
        - (v[z, :-2, 1:-1] - 2 * c + v[z, 2:, 1:-1]) / h2 - (v[z - 1, 1:-1, 1:-1] - 2 * c + v[z + 1, 1:-1, 1:-1]) / h2
    return r


def residual_restrict(v, f, cf, cv, clo, chi):
    cn = cf.shape[1]
    w1 = np.array([0.25, 0.5, 0.25])
    for cz in range(clo, chi):
        cv[cz] = 0.0
        cf[cz] = 0.0
        if cz == 0 or cz == cn - 1:
            continue
        acc = np.zeros((cn - 2, cn - 2))
        for dz in (-1, 0, 1):
            r = residual_plane(v, f, 2 * cz + dz)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    acc += w1[dz + 1] * w1[dy + 1] * w1[dx + 1] * r[2 + dy:2 * cn - 2 + dy:2, 2 + dx:2 * cn - 2 + dx:2]
        cf[cz, 1:-1, 1:-1] = acc


def colour_mask(n, z, colour):
    """points (y, x) of plane z whose colour (x + y + z) & 1 equals `colour`"""
    y, x = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return ((x + y + z) % 2) == colour


def interpolate_add(v, cv, lo, hi, colour=None):
    n = v.shape[1]
    for z in range(lo, hi):
        cz = z // 2
        pz = cv[cz] if z % 2 == 0 else 0.5 * (cv[cz] + cv[cz + 1])
        py = np.empty((n, pz.shape[1]))
        py[0::2] = pz
        py[1::2] = 0.5 * (pz[:-1] + pz[1:])
        e = np.empty((n, n))
        e[:, 0::2] = py
        e[:, 1::2] = 0.5 * (py[:, :-1] + py[:, 1:])
        if colour is None:
            v[z, 1:-1, 1:-1] += e[1:-1, 1:-1]
        else:
            m = colour_mask(n, z, colour)[1:-1, 1:-1]
            blk = v[z, 1:-1, 1:-1]
            blk[m] += e[1:-1, 1:-1][m]


# ---- the driver's schedule ---------------------------------------------------------------------------

class Level:
    def __init__(self, n, plan):
        self.n = n
        self.dist = plan["dist"]
        self.z0, self.nzl = plan["z0"], plan["nzl"]
        self.a, self.b = plan["z0"] + plan["own_lo"], plan["z0"] + plan["own_hi"]  # owned global planes
        self.v = np.full((n, n, n), np.nan)
        self.f = np.full((n, n, n), np.nan)
        self.vg_valid = True
        self.vg_deep = True
        self.v_other = np.full((n, n, n), np.nan)  # second v buffer of the temporally blocked smoother

    def mirror_v_ghosts(self):
        """mg3d_host.c::mirror_v_ghosts: the ghost planes of the other buffer get the values of the current one"""
        for z in list(range(max(self.z0, 0), self.a)) + list(range(self.b, min(self.z0 + self.nzl, self.n))):
            self.v_other[z] = self.v[z]

    def interior(self):
        return max(self.a, 1), min(self.b, self.n - 1)


def exchange(L, arr, rank, world, depth_up, down, colour=None):
    """depth_up planes travel up, `down` planes travel down.  colour None: whole planes; otherwise only that colour's points of
    the ghost planes are overwritten."""
    if not L.dist:
        return
    down = int(down)
    reqs = []
    if rank + 1 < world:
        if depth_up:
            reqs.append(dist.isend(torch.from_numpy(arr[L.b - depth_up:L.b].copy()), rank + 1))
        if down:
            up_ghost = torch.empty((down, L.n, L.n), dtype=torch.float64)
            reqs.append(dist.irecv(up_ghost, rank + 1))
    if rank > 0:
        if down:
            reqs.append(dist.isend(torch.from_numpy(arr[L.a:L.a + down].copy()), rank - 1))
        if depth_up:
            lo_ghost = torch.empty((depth_up, L.n, L.n), dtype=torch.float64)
            reqs.append(dist.irecv(lo_ghost, rank - 1))
    for r in reqs:
        r.wait()

    def put(z, plane):
        if colour is None:
            arr[z] = plane
        else:
            m = colour_mask(L.n, z, colour)
            arr[z][m] = plane[m]

    if rank + 1 < world and down:
        for k in range(down):
            put(L.b + k, up_ghost.numpy()[k])
    if rank > 0 and depth_up:
        for k in range(depth_up):
            put(L.a - depth_up + k, lo_ghost.numpy()[k])


def ensure_v_ghosts(L, rank, world):
    """mg3d_host.c::ensure_v_ghosts"""
    if L.dist and not L.vg_valid:
        exchange(L, L.v, rank, world, 2, 1)
        L.vg_valid = True


def relax_pipe(L, rank, world, nu, correct_from=None):
    """mg3d_host.c::relax_level_ex on a level the temporally blocked smoother takes: pairs of sweeps per pass, a remaining
    single sweep colour by colour.  correct_from: the coarse level whose prolongation + correction (colour 1 only) the FIRST
    pass applies on the fly, on the owned planes AND on its four halo planes per side (the coarse level was exchanged two planes
    up and three down for that), without ever writing the corrected input back."""
    n = L.n
    while nu >= 2:
        if not L.vg_deep:  # InitV / the zeroed coarse v left valid ghosts of full depth
            exchange(L, L.v, rank, world, 4, 4, 1)
        w = L.v.copy()
        if correct_from is not None:
            C = correct_from
            if C.dist:
                exchange(C, C.v, rank, world, 2, 3)
                C.vg_valid = True
            interpolate_add(w, C.v, max(L.a - 4, 1), min(L.b + 4, n - 1), 1)
            correct_from = None
        for colour, e in ((0, 3), (1, 2), (0, 1), (1, 0)):  # R1, B1, R2, B2 on the owned planes grown by e
            lo, hi = max(L.a - e, 1), min(L.b + e, n - 1)
            relax_colour(w, L.f, colour, lo, hi)
        out = L.v_other                  # the other buffer: only the owned planes are written; of its ghost planes the
        out[L.a:L.b] = w[L.a:L.b]        # next pass will read the Dirichlet points (mirror_v_ghosts)
        L.v_other = L.v
        L.v = out
        L.vg_valid = not L.dist
        L.vg_deep = False
        nu -= 2
    if nu:
        ensure_v_ghosts(L, rank, world)
        L.vg_deep = False
        relax(L, rank, world, nu)


def relax(L, rank, world, nu):
    lo, hi = L.interior()
    for _ in range(nu):
        for colour in (0, 1):
            relax_colour(L.v, L.f, colour, lo, hi)
            exchange(L, L.v, rank, world, 1, 1, colour)


OMEGA = 6.0 / 7.0


def jacobi_colour(dst, own, oth, f, colour, lo, hi):
    """k_jacobi_colour on planes [lo, hi): the points of `colour` become old + omega*(gs - old), gs from the values
    `oth` holds at the six neighbours; non-interior points are copied when dst is not own."""
    n = dst.shape[1]
    h2 = (1.0 / (n - 1)) ** 2
    for z in range(lo, hi):
        m = colour_mask(n, z, colour)
        if dst is not own:
            dst[z][m] = own[z][m]
        if z < 1 or z > n - 2:
            continue
        gs = (oth[z, 1:-1, :-2] + oth[z, 1:-1, 2:] + oth[z, :-2, 1:-1] + oth[z, 2:, 1:-1] + oth[z - 1, 1:-1, 1:-1] +
              oth[z + 1, 1:-1, 1:-1] - f[z, 1:-1, 1:-1] * h2) / 6.0
        old = own[z, 1:-1, 1:-1]
        new = old + OMEGA * (gs - old)
        mi = m[1:-1, 1:-1]
        blk = dst[z, 1:-1, 1:-1]
        blk[mi] = new[mi]


def relax_jacobi(L, rank, world, nu):
    """mg3d_host.c::relax_jacobi_level: new colour 0 into a scratch array, colour 1 in place, scratch and v swap roles
    every sweep; per-colour halo exchanges; an odd count ends with a copy back of the owned planes and the nearest ghost
    on each side only."""
    if not hasattr(L, "scratch"):
        L.scratch = np.full((L.n,) * 3, np.nan)
    cur = L.v
    for _ in range(nu):
        nxt = L.scratch if cur is L.v else L.v
        jacobi_colour(nxt, cur, L.v, L.f, 0, L.a, L.b)    # colour-1 neighbours always live in v
        jacobi_colour(L.v, L.v, cur, L.f, 1, L.a, L.b)    # in place; colour-0 neighbours = the OLD ones
        cur = nxt
        exchange(L, cur, rank, world, 1, 1, 0)
        exchange(L, L.v, rank, world, 1, 1, 1)
    if cur is not L.v:
        for z in range(max(L.a - 1, max(L.z0, 0)), min(L.b + 1, L.z0 + L.nzl, L.n)):
            m = colour_mask(L.n, z, 0)
            L.v[z][m] = cur[z][m]


def relax_jacobi_whole(L, nu):
    """the definition: every interior point from the old values (oracle/mg_oracle_impl.h::orc3d_relax_jacobi)"""
    n = L.n
    h2 = (1.0 / (n - 1)) ** 2
    for _ in range(nu):
        v = L.v
        gs = (v[1:-1, 1:-1, :-2] + v[1:-1, 1:-1, 2:] + v[1:-1, :-2, 1:-1] + v[1:-1, 2:, 1:-1] + v[:-2, 1:-1, 1:-1] +
              v[2:, 1:-1, 1:-1] - L.f[1:-1, 1:-1, 1:-1] * h2) / 6.0
        nv = v.copy()
        nv[1:-1, 1:-1, 1:-1] = v[1:-1, 1:-1, 1:-1] + OMEGA * (gs - v[1:-1, 1:-1, 1:-1])
        L.v = nv


def vcycle(levels, l, rank, world, engine_schedule=True, smoother="gs"):
    """engine_schedule: colour-1-only correction (and exchange) as in mg3d_host.c::vcycle_rec; False = the reference's
    full ApplyCorrection, used for the sequential run the slabs are compared with.  smoother "jacobi": the weighted
    Jacobi option (full correction, both colours exchanged); the sequential run uses the whole-grid definition."""
    L = levels[l]
    pipe = smoother == "pipe" and engine_schedule and L.n >= PIPE_MIN_N
    if smoother == "jacobi":
        smooth = (lambda: relax_jacobi(L, rank, world, NU_J)) if engine_schedule else (lambda: relax_jacobi_whole(L, NU_J))
        engine_colour = None
    elif pipe:
        smooth = lambda: relax_pipe(L, rank, world, NU_P)
        engine_colour = 1
    elif smoother == "pipe":
        smooth = lambda: (ensure_v_ghosts(L, rank, world), setattr(L, "vg_deep", False), relax(L, rank, world, NU_P))
        engine_colour = 1 if engine_schedule else None
    else:
        smooth = lambda: relax(L, rank, world, NU)
        engine_colour = 1 if engine_schedule else None
    smooth()
    if l + 1 < len(levels):
        C = levels[l + 1]
        if pipe and not L.vg_deep:                                # one launch: colour 0 two up / one down (if stale), colour 1 four
            exchange(L, L.v, rank, world, 2, 0 if L.vg_valid else 1, 0)   # each way -- what the first pass of the post-smoothing
            exchange(L, L.v, rank, world, 4, 4, 1)                # reads of its neighbours (v does not change until then)
            L.vg_valid = L.vg_deep = True
        elif smoother == "pipe":                                  # two planes up; one down as well if the ghosts are stale
            exchange(L, L.v, rank, world, 2, 0 if L.vg_valid else 1)
            L.vg_valid = True
        else:
            exchange(L, L.v, rank, world, 2, 0)                   # plane a-2 for the fused residual+restrict
        if C.dist or not L.dist:
            clo, chi = C.a, C.b
        else:                                                     # first agglomerated level: planes under my slab
            clo, chi = L.a // 2, (L.b // 2 if rank < world - 1 else C.n)
        residual_restrict(L.v, L.f, C.f, C.v, clo, chi)
        if L.dist:
            C.v[max(C.z0, 0):C.z0 + C.nzl] = 0.0                  # coarse v = 0 wherever this rank stores it
            C.mirror_v_ghosts()
        C.vg_valid = C.vg_deep = True
        if C.dist:
            exchange(C, C.f, rank, world, 4, 4)
        elif L.dist:                                              # all-gather of the equal shares + top plane from the last rank
            m = (C.n - 1) // world
            parts = [torch.empty((m, C.n, C.n), dtype=torch.float64) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(C.f[rank * m:(rank + 1) * m].copy()))
            for r, p in enumerate(parts):
                C.f[r * m:(r + 1) * m] = p.numpy()
            if rank != world - 1:                                 # the restricted residual is +0 on the Dirichlet plane n-1:
                C.f[C.n - 1] = 0.0                                # written locally (mg3d_host.c::gather_level, MG_TOP_ZERO)
        vcycle(levels, l + 1, rank, world, engine_schedule, smoother)
        if pipe and NU_P >= 2:  # prolongation + correction ride on the first pass of the post-smoothing
            relax_pipe(L, rank, world, NU_P, correct_from=C)
            return
        lo, hi = L.interior()
        if smoother == "pipe":
            ensure_v_ghosts(C, rank, world)
        interpolate_add(L.v, C.v, lo, hi, engine_colour)
        L.vg_deep = False
        if smoother == "pipe" and L.dist and (pipe or not L.vg_valid):
            L.vg_valid = False                                    # lazily: the next pass (or reader) fetches what it needs
        else:
            exchange(L, L.v, rank, world, 1, 1, engine_colour)
    smooth()


def problem(n):
    x = np.linspace(0.0, 1.0, n)
    f = -3 * np.pi ** 2 * np.sin(np.pi * x)[:, None, None] * np.sin(np.pi * x)[None, :, None] * np.sin(np.pi * x)[None, None, :]
    return np.zeros((n, n, n)), f


def worker(rank, world, port, plans, out_queue, smoother="gs", N=N):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    levels = [Level(n, plans[(n, rank)]) for n in sizes(N)]
    v0, f0 = problem(N)
    L0 = levels[0]
    sl = slice(L0.z0, L0.z0 + L0.nzl)
    L0.v[sl], L0.f[sl] = v0[sl], f0[sl]
    L0.mirror_v_ghosts()
    for _ in range(CYCLES):
        vcycle(levels, 0, rank, world, True, smoother)
    out_queue.put((rank, L0.a, L0.b, L0.v[L0.a:L0.b].copy()))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("smoother,world", [("gs", 2), ("jacobi", 2), ("pipe", 2), ("pipe", 4), ("gs", 4)])
def test_slab_schedule_gloo(mg, monkeypatch, smoother, world):
    """world 2: every rank has one neighbour; world 4: ranks 1 and 2 have ghosts on both sides and the gather has four shares"""
    N = 129 if smoother == "pipe" else 65  # pipe: two distributed levels (129, 65), the coarse one fed by exchanged f ghosts
    monkeypatch.setenv("MG_B200_DIST_MIN_N", "65")  # distribute the small test grids too (default threshold: n >= 257)
    plans = {(n, r): mg.MultiGrid3D.plan_level(n, world, r) for n in sizes(N) for r in range(world)}
    assert plans[(65, 0)]["dist"] == 1 and plans[(33, 0)]["dist"] == 0
    # sequential reference: the same stand-in kernels on the whole grid in one process
    one = {(n, 0): mg.MultiGrid3D.plan_level(n, 1, 0) for n in sizes(N)}
    ref_levels = [Level(n, one[(n, 0)]) for n in sizes(N)]
    ref_levels[0].v, ref_levels[0].f = problem(N)
    for L in ref_levels[1:]:
        L.v[...] = 0.0
        L.f[...] = 0.0
    for _ in range(CYCLES):
        vcycle(ref_levels, 0, 0, 1, engine_schedule=False, smoother=smoother)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, plans, q, smoother, N)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = 0
    for rank, a, b, v in got:
        assert not np.isnan(v).any(), "a NaN reached the owned planes of rank %d: a ghost plane was missing" % rank
        assert np.array_equal(v, ref_levels[0].v[a:b]), "rank %d differs from the sequential run" % rank
        covered += b - a
    assert covered == N
    r = ref_levels[0]
    assert np.isfinite(r.v).all() and np.abs(r.v).max() > 0.1  # the cycles did something (solution ~ sin sin sin)
