#!/usr/bin/env python
"""Multi-GPU check of the sharded ensemble (run by hand under torchrun on >= 2 GPUs; pytest -m gpu runs on one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Every rank holds a contiguous slice of the recordings (host.SamplingEnsemble(rank=, world=)); rank 0 also runs the whole
ensemble on its own GPU.  Per-chain paths and accept decisions of the slices must equal the unsharded run BITWISE, and
fetch_ll / accept counts through the native NCCL allreduce (dmt_allreduce_stats) must equal the unsharded sums.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmt_b200  # noqa: E402
from dmt_b200 import configs  # noqa: E402
from dmt_b200 import host as H  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M, K, nit = 96 + 5, 8, 4
    layouts = [([(0, 2), (3, 5), (6, 7)], 0.8), ([(0, 3), (4, 7)], 0.7)]
    prob = configs.make_problem("lorenz", M, K=K, dt=0.01, seed=3, layouts=layouts)
    rec = dict(theta=prob.theta, L=prob.L, Sigma=prob.Sigma, v=prob.v, x0=prob.x0, xbar=prob.xbar)

    def run(se):
        se.init_paths()
        bes = [H.BlockEnsemble(se, r, rho, nit) for r, rho in layouts]
        for be in bes:
            H.enable_guiding_cache(be)
        out = []
        for i in range(nit):
            for be in bes:
                H.blocking_sweep(be, i)
                H.accept_reject_proposal_path(be, i)
                out.append((H.fetch_ll(be), H.fetch_ll_o(be), se.ctx.get_last_accept(be.layout).copy()))
        return se.ctx.get_X(0), out

    se = H.SamplingEnsemble(prob.model, rec, (prob.n_pts, prob.tt), device=local, seed=77, two_sided_laws=False, rank=rank, world=world)
    uid = torch.from_numpy(se.ctx.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    dist.broadcast(uid, 0)
    se.comm_init(uid.cpu().numpy())
    X, out = run(se)
    lo, hi = se.chain_lo, se.chain_hi
    ok = True
    if rank == 0:
        full = H.SamplingEnsemble(prob.model, rec, (prob.n_pts, prob.tt), device=local, seed=77, two_sided_laws=False)
        Xf, outf = run(full)
        ok &= bool(np.array_equal(X, Xf[:, :, lo:hi]))
        for (ll, llo, acc), (llf, llof, accf) in zip(out, outf):
            ok &= bool(np.array_equal(acc, accf[:, lo:hi]))
            ok &= abs(ll - llf) <= 1e-11 * abs(llf) and abs(llo - llof) <= 1e-11 * abs(llof)
        ref = torch.from_numpy(Xf).cuda()
    else:
        ref = torch.empty((prob.n_pts.sum(), prob.d, M), dtype=torch.float64, device="cuda")
    dist.broadcast(ref, 0)
    ok &= bool(np.array_equal(X, ref.cpu().numpy()[:, :, lo:hi]))
    # the same run with the library's own peer-memory all-reduce instead of NCCL: same sums (to rounding: NCCL's order is its own),
    # bit-identical on every rank
    se2 = H.SamplingEnsemble(prob.model, rec, (prob.n_pts, prob.tt), device=local, seed=77, two_sided_laws=False, rank=rank, world=world)
    se2.p2p_init()
    X2, out2 = run(se2)
    ok &= bool(np.array_equal(X2, X))
    for (ll, llo, _), (ll2, llo2, _) in zip(out, out2):
        ok &= abs(ll - ll2) <= 1e-12 * abs(ll) and abs(llo - llo2) <= 1e-12 * abs(llo)
    mine = torch.tensor([o[0] for o in out2] + [o[1] for o in out2], dtype=torch.float64, device="cuda")
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    ok &= all(bool(torch.equal(every[0], e)) for e in every)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi_gpu_check world=%d: %s (chains %d, slices of rank0 %d..%d)" % (world, "PASS" if flag.item() == 1.0 else "FAIL", M, lo, hi))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
