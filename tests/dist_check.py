"""Multi-GPU equivalence check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

RB Gauss-Seidel is partition-invariant, so the z-slab engine must reproduce the single-GPU engine BIT FOR BIT
(SURVEY.md 8e).  Every rank runs the distributed engine and, on its own GPU, the single-GPU engine, and compares
the planes it owns: V-cycles (both residual modes, float and double, with and without the TMA kernels), FMG,
the fused level operators and the residual norms (allreduce)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pde_multigrid_b200 as mg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = []

    def new_uid():
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = (ctypes.c_ubyte * 128)()
            mg._lib.check(mg.lib().mg_comm_unique_id(raw))
            buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().tolist())

    def same(a, b):
        return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))

    # default plan (levels n >= 257 distributed) at 513^3, and a deep-nesting plan (threshold lowered to 65) at 129/257
    runs = [(513, None)] + ([(129, "65"), (257, "65")] if world <= 4 else [(257, "65")])
    if os.environ.get("DIST_CHECK_RUNS"):  # e.g. "257:65,513:" to repeat a subset while debugging
        runs = [(int(a), b or None) for a, b in (r.split(":") for r in os.environ["DIST_CHECK_RUNS"].split(","))]
    sizes = [r[0] for r in runs]
    for n, min_n in runs:
        if min_n is None:
            os.environ.pop("MG_B200_DIST_MIN_N", None)
        else:
            os.environ["MG_B200_DIST_MIN_N"] = min_n
        for dtype in (np.float64, np.float32):
            for corrected in (True, False):
                mode = mg.MG_CORRECTED if corrected else mg.MG_REF_COMPAT
                d = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mode, rank=rank, nranks=world, nccl_unique_id=new_uid())
                s = mg.MultiGrid3D(n, dtype=dtype, residual_mode=mode)
                tag = "n=%d %s %s" % (n, np.dtype(dtype).name, "corrected" if corrected else "ref_compat")
                # random finest fields through set_field (owned planes + ghost refresh)
                rng = np.random.default_rng(100 + n)
                v0 = rng.uniform(-1, 1, (n, n, n)).astype(dtype)
                f0 = rng.uniform(-1, 1, (n, n, n)).astype(dtype)
                zb, zc = d.owned_range(0)
                d.set_v(0, v0[zb:zb + zc])
                d.set_f(0, f0[zb:zb + zc])
                s.set_v(0, v0)
                s.set_f(0, f0)
                for step in range(2):
                    d.VCycle(0, 2, 1)
                    s.VCycle(0, 2, 1)
                    for l in range(d.numGrids):
                        zb, zc = d.owned_range(l)
                        if not same(d.get_v(l), s.get_v(l)[zb:zb + zc]):
                            failures.append("%s: v level %d differs after V-cycle %d on rank %d" % (tag, l, step, rank))
                        if not same(d.get_f(l), s.get_f(l)[zb:zb + zc]):
                            failures.append("%s: f level %d differs after V-cycle %d on rank %d" % (tag, l, step, rank))
                    dn, sn = d.residual_norm(0), s.residual_norm(0)
                    if not (abs(dn[0] - sn[0]) <= 1e-12 * abs(sn[0]) and dn[1] == sn[1]):
                        failures.append("%s: norms %r vs %r" % (tag, dn, sn))
                    if d.field_checksum(0) != s.field_checksum(0) or d.field_checksum(1, mg.MG_FIELD_F) != s.field_checksum(1, mg.MG_FIELD_F):
                        failures.append("%s: field checksums (summed over the ranks) differ from the single-GPU ones" % tag)
                # pure two-sweep passes (temporally blocked smoother on the slabs), eager, captured, replayed
                for step in range(3):
                    d.VCycle(0, 2, 2)
                    s.VCycle(0, 2, 2)
                for l in range(d.numGrids):
                    zb, zc = d.owned_range(l)
                    if not same(d.get_v(l), s.get_v(l)[zb:zb + zc]):
                        failures.append("%s: v level %d differs after three V(2,2) on rank %d" % (tag, l, rank))
                d.Relax(0, 4)
                s.Relax(0, 4)
                zb, zc = d.owned_range(0)
                if not same(d.get_v(0), s.get_v(0)[zb:zb + zc]):
                    failures.append("%s: Relax(0, 4) differs on rank %d" % (tag, rank))
                zb, zc = d.owned_range(0)
                if not same(d.CalculateResidual(0), s.CalculateResidual(0)[zb:zb + zc]):
                    failures.append("%s: CalculateResidual differs on rank %d" % (tag, rank))
                # FMG on the reference problem
                d.init_problem()
                s.init_problem()
                d.FullMultiGridVCycle(0, 1, 2, 2)
                s.FullMultiGridVCycle(0, 1, 2, 2)
                for l in range(d.numGrids):
                    zb, zc = d.owned_range(l)
                    if not same(d.get_v(l), s.get_v(l)[zb:zb + zc]):
                        failures.append("%s: FMG v level %d differs on rank %d" % (tag, l, rank))
                if n == sizes[-1] and dtype == np.float64 and corrected:
                    # the plain (non-TMA) kernels through the same slab logic
                    d.init_problem()
                    s.init_problem()
                    d.set_smoother(mg.MG_SMOOTHER_COLOUR, 1)
                    d.VCycle(0, 1, 1)
                    s.VCycle(0, 1, 1)
                    zb, zc = d.owned_range(0)
                    if not same(d.get_v(0), s.get_v(0)[zb:zb + zc]):
                        failures.append("%s: plain-kernel V-cycle differs on rank %d" % (tag, rank))
                if n == sizes[-1] and corrected:
                    # weighted Jacobi (also partition-invariant): odd and even sweep counts, third cycle = graph replay
                    d.init_problem()
                    s.init_problem()
                    d.set_smoother(mg.MG_SMOOTHER_JACOBI, 1)
                    s.set_smoother(mg.MG_SMOOTHER_JACOBI, 1)
                    for step in range(3):
                        d.VCycle(0, 3, 2)
                        s.VCycle(0, 3, 2)
                    for l in range(d.numGrids):
                        zb, zc = d.owned_range(l)
                        if not same(d.get_v(l), s.get_v(l)[zb:zb + zc]):
                            failures.append("%s: Jacobi V-cycle v level %d differs on rank %d" % (tag, l, rank))
                d.close()
                s.close()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures[:10]:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("DIST_CHECK %s world=%d failures=%d" % ("OK" if flag.item() == 0 else "FAILED", world, int(flag.item())), flush=True)
    dist.destroy_process_group()
    return 1 if flag.item() else 0


if __name__ == "__main__":
    sys.exit(main())
