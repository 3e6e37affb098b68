#!/usr/bin/env python
"""bench.py — guided path updates/s (chains x EM steps / s, FP64) of the blocking path update, BASELINE.json config C3:
Lorenz 3-D, 4096 chains PER GPU (weak scaling), 200 observation intervals x 100 EM steps, two staggered block layouts
(10 / 11 blocks) alternated, pCN rho = 0.9, every chain with its own data and guiding term (P = M).

One "step" = one blocking sweep over one layout, exactly the tutorial loop body
(/root/reference/docs/src/tutorials/block_collection/inference_with_blocking.md:52-58):
    set_obs! -> recompute_guiding_term!(P only) -> find_W_for_X! + loglikhd! -> draw_proposal_path! ->
    accept_reject_proposal_path! -> (N > 1: NCCL allreduce of ll sums / accept counts)
and it advances every chain by S = 20,000 guided Euler–Maruyama steps: units per step = M x S per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c2|c4|c5] [--chains M] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {  # per chain per EM step of draw_proposal_path! = 8(2 dw + d) [+ 8(d(d+1)/2 + d) when P = M]
    "c2": (48, 88), "c3": (72, 144), "c4": (96, 208), "c5": (64, 280), "c1": (32, 72),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c3")
    ap.add_argument("--chains", type=int, default=None, help="chains PER GPU (default: the config's M)")
    ap.add_argument("--psets", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-chains", type=int, default=None, help="chains in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--one-call", action="store_true", help="blocking sweep through dmt_blocking_sweep (same launches; per-kernel timing is then not available)")
    ap.add_argument("--no-cache", action="store_true", help="blocking sweep with the full backward filter every sweep (no guiding cache)")
    ap.add_argument("--separate", action="store_true", help="blocking sweep with the three separate passes instead of the fused one")
    ap.add_argument("--sweep-mode", type=int, default=0, help="fused pass: 0 auto (software-pipelined where eligible), 1 register-tile kernel, 2 pipelined")
    ap.add_argument("--eager-noise", action="store_true", help="blocking sweep stores W_acc / W° every sweep (default: lazy noise, rebuilt on demand)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_arm(cfg_name, n_chains, sweeps, warm, threads=None):
    """Times the C restatement of the reference algorithm (oracle/, OpenMP over recordings) on a bounded sample of the
    same workload.  kind = "port": the Julia reference and its un-vendored numerical dependencies cannot run here."""
    import dmt_b200
    from dmt_b200 import configs
    from oracle import orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from harness import OracleEnsemble
    import ctypes as C
    threads = threads or os.cpu_count() or 1
    prob = configs.named_config(cfg_name, M=n_chains, seed=123)
    lib = orc.load(omp=True)
    ora = OracleEnsemble(orc, lib, prob, seed=123)
    blocking = len(prob.layouts) > 1
    # initial paths: whole-path guiding term, fresh noise, forced accept
    for P in ora.pairs:
        bb = P.biblock(0, prob.K - 1, True, 0.0)
        P.recompute_guiding_term(bb, 0)
        tries = 0
        while not P.draw_proposal_path(bb, seed=123, chain=tries, it=9999):
            tries += 1
        P.accept_reject(bb, float("inf"))
        P.loglikhd(bb, 0)
        if not blocking:
            for lay in ora.layouts:
                lay[ora.pairs.index(P)][0].ll[0] = bb.ll[0]
    arr_t = C.c_void_p * len(ora.pairs)
    handles = arr_t(*[P.h for P in ora.pairs])
    gt0 = orc.gtile0_of(prob.n_pts)
    flat = []
    for l, lay in enumerate(ora.layouts):
        nb = len(lay[0])
        blk = (orc.BiBlock * (nb * prob.M))()
        for c in range(prob.M):
            for b in range(nb):
                blk[c * nb + b] = lay[c][b]
        flat.append((blk, nb))
    nacc = C.c_int()

    def sweep(it):
        l = it % len(flat)
        blk, nb = flat[l]
        lib.orc_sweep_many(handles, blk, prob.M, nb, 123, 0, it, l, gt0.ctypes.data_as(C.POINTER(C.c_int)), int(blocking), threads,
                           C.byref(nacc))
    for it in range(warm):
        sweep(it)
    t0 = time.perf_counter()
    for it in range(warm, warm + sweeps):
        sweep(it)
    dt = time.perf_counter() - t0
    units = prob.M * prob.steps_per_chain * sweeps
    return units / dt, dt / sweeps, threads, "%d chains x %d sweeps of config %s (%d steps/chain), %d OpenMP threads" % (
        prob.M, sweeps, cfg_name, prob.steps_per_chain, threads)


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=0.0, t1=1e300):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if not (t0 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(a):
    import torch
    import torch.distributed as dist
    import dmt_b200
    from dmt_b200 import _lib, configs

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    base = configs.named_config(a.config, M=1, seed=0)  # only to learn the default M cheaply
    M = a.chains or {"c1": 1, "c2": 1024, "c3": 4096, "c4": 16384, "c5": 8192}[a.config]
    prob = configs.named_config(a.config, M=M, seed=0, chain_offset=rank * M, P=a.psets)
    del base
    blocking = len(prob.layouts) > 1
    nlay = len(prob.layouts)
    ctx = dmt_b200.Ctx(prob.model, prob.n_pts, prob.tt, prob.M, prob.P, obs_dim=prob.m, device=local, n_layouts=nlay + 1,
                       chain_offset=rank * M, seed=2026, pset_of_chain=prob.pset_of_chain)
    configs.upload(prob, ctx)
    ctx.set_sweep_mode(a.sweep_mode)
    lazy = blocking and not a.eager_noise and not a.separate and a.sweep_mode != 1
    whole = nlay
    ctx.set_blocks(whole, [(0, prob.K - 1)], 0.0)
    ctx.recompute_guiding_term(whole, _lib.P_ONLY)
    nfail = ctx.init_paths(whole, iter0=1 << 20, max_tries=100)
    assert nfail == 0, "init_paths left %d failing chains" % nfail
    if not blocking:
        ctx.recompute_guiding_term(0, _lib.P_ONLY)
        ctx.loglikhd(0, 0, 0)
    if lazy:
        ctx.set_lazy_noise(True)
    if blocking and not a.no_cache:  # smoothing: the laws stay fixed, only the blocks' frozen end points move => K1 through the guiding cache
        for l in range(nlay):
            ctx.enable_guiding_cache(l)
    allreduce_kind = "none"
    if world > 1:
        # the small stats all-reduce as the library's own one-shot kernel over NVLink peer memory (dmt_p2p_init); if any rank cannot
        # map its peers (IPC not permitted), every rank falls back to the NCCL communicator inside libdmt
        ok = torch.ones(1, device=dev)
        if os.environ.get("DMT_NO_P2P"):
            ok.zero_()
        else:
            try:
                mine = torch.from_numpy(ctx.p2p_export().copy()).to(dev)
            except Exception:
                mine = torch.zeros(64, dtype=torch.uint8, device=dev)
                ok.zero_()
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            if ok.item():
                try:
                    ctx.p2p_init(world, rank, np.stack([h.cpu().numpy() for h in allh]))
                except Exception:
                    ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item():
            allreduce_kind = "p2p kernel (NVLink peer memory)"
        else:
            ctx.p2p_disable()
            uid = torch.from_numpy(ctx.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
            dist.broadcast(uid, 0)
            ctx.comm_init(world, rank, uid.cpu().numpy())
            allreduce_kind = "ncclAllReduce"
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    fused = blocking and not a.separate  # find_W_for_X! + loglikhd! + draw_proposal_path! in one pass (dmt_find_W_loglikhd_draw)
    k1_each_step = (a.config == "c5")    # BASELINE C5: "each sweep = set_params -> K1 (P = M) -> K2" (backward-filter dominated)
    theta_dev_host = np.repeat(prob.theta[:, None], prob.P, axis=1).copy()
    onecall = fused and a.one_call  # dmt_blocking_sweep: set_obs! .. draw_proposal_path! in one call (same launches, one timing bracket)
    if onecall:
        names = ["sweep_fused", "accept", "stats"]
    elif fused:
        names = ["set_obs", "bwd_filter", "sweep_fused", "accept", "stats"]
    elif k1_each_step:
        names = ["set_params_aux", "bwd_filter", "draw", "accept", "stats"]
    else:
        names = (["set_obs", "bwd_filter", "invsolve_ll"] if blocking else []) + ["draw", "accept", "stats"]
    # stats = reduce + finish kernels; c5: put_record + aux_linearise; one-call sweep: set_obs gather + the fused pass
    launches_per_step = len(names) + 1 + (1 if k1_each_step else 0) + (3 if onecall else 0)  # + set_obs, 2 apply + 1 apply_c
    ev = {n: [] for n in names}

    def sweep(it, timed, E=None):
        l = it % nlay
        marks = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append(e)
        mark()
        if onecall:
            ctx.blocking_sweep(l, it); mark()   # set_obs!, recompute_guiding_term!(P only), find_W_for_X!, loglikhd!, draw_proposal_path!
        elif blocking:
            ctx.set_artificial_obs(l); mark()
            ctx.recompute_guiding_term(l, _lib.P_ONLY); mark()
            if fused:
                ctx.find_W_loglikhd_draw(l, it); mark()
            else:
                ctx.find_W_and_loglikhd(l); mark()
        if k1_each_step:  # new parameters (theta jitter keeps the data valid), aux laws re-linearised on the device, K1, then the path update
            ctx.set_params(theta_dev_host, side=0, stores=1)
            ctx.set_aux_linearised(None, side=0, store=_lib.STORE_PP); mark()   # same points, new theta: re-linearised on the device
            ctx.recompute_guiding_term(l, _lib.P_ONLY); mark()
        if not fused:
            ctx.draw_proposal_path(l, it); mark()
        ctx.accept_reject_path(l, it, E); mark()   # E: host-drawn Exp(1) (the reference's rand(Exponential(1.0)), src/biblock.jl:122) or device Philox
        stats = ctx.allreduce_stats(l); mark()        # [sum ll, sum ll°, accept counts...] (NCCL allreduce when N > 1)
        if timed:
            for n, e0, e1 in zip(names, marks[:-1], marks[1:]):
                ev[n].append((e0, e1))
        return stats

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync(); torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi takes a moment to start: launch it before the warm-up, keep only the timed window
    for it in range(a.warmup):
        sweep(it, False)
    barrier()
    e_start = torch.cuda.Event(enable_timing=True); e_stop = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e_start.record(stream)
    last = None
    for it in range(a.warmup, a.warmup + a.steps):
        last = sweep(it, True)
    e_stop.record(stream)
    barrier()
    clocks = sampler.stop(t_wall0, time.time())
    ms = torch.tensor([e_start.elapsed_time(e_stop)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    units_per_step = prob.M * prob.steps_per_chain * world
    value = units_per_step * a.steps / (ms_total * 1e-3)
    kern_ms = {n: float(np.mean([e0.elapsed_time(e1) for e0, e1 in ev[n]])) for n in names}
    its = list(range(a.warmup, a.warmup + a.steps))
    kern_ms_by_layout = {n: [float(np.mean([e0.elapsed_time(e1) for (e0, e1), it in zip(ev[n], its) if it % nlay == l] or [0.0]))
                             for l in range(nlay)] for n in names if n in ("bwd_filter", "sweep_fused", "draw", "invsolve_ll")}

    # ---- roofline of the dominant kernel of the unit of work: fwd_kernel<Lorenz, OP_DRAW> (pCN + guided EM + ll)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bsh, bper = ALGO_BYTES[a.config]
    bpu = bper if prob.P == prob.M else bsh
    kname, kop = "draw", "OP_DRAW"
    if fused:
        # the fused pass reads X_acc and H,F and writes W_acc, W°, X°; the accepted noise never leaves registers, so it moves
        # 8(2d + 2dw) + 8(d(d+1)/2 + d) = 168 B per step for Lorenz — LESS than SURVEY §8(d)'s draw (144) + K5/K4 (48) figures
        bpu = 8 * (2 * prob.d + (0 if lazy else 2 * prob.dw)) + (8 * (prob.d * (prob.d + 1) // 2 + prob.d) if prob.P == prob.M else 0)
        kname, kop = "sweep_fused", "OP_SWEEP"
    algo_bytes = bpu * prob.M * prob.steps_per_chain
    achieved = algo_bytes / (kern_ms[kname] * 1e-3) / 1e9
    roofline = {"kernel": "fwd_kernel<%s, %s>" % (_lib.MODEL_NAMES[prob.model], kop), "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_unit": bpu, "units_per_launch": prob.M * prob.steps_per_chain, "launch_ms": kern_ms[kname]}
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel, from the committed `ncu --set full` capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get(roofline["kernel"])
        if ent and ent.get("units_per_launch") == roofline["units_per_launch"]:
            roofline["traffic"] = ent["dram_bytes_per_launch"]
            roofline["traffic_source"] = ent["source"]
    except Exception:
        pass
    if blocking:  # the whole sweep also moves the K1 write and the K5+K4 pass (reported separately)
        d, dw = prob.d, prob.dw
        nh = d * (d + 1) // 2
        # K1 writes H,F; then either the fused pass (bpu above) or the K5+K4 pass (X, H,F in; W out) followed by the draw
        k1_bytes = 8 * (nh + d) if a.no_cache else 8 * (d + d * d) + 8 * d  # full K1 writes H,F; cached K1 reads F0,Psi and writes F
        per_step = k1_bytes + (bpu if fused else (8 * (d + dw) + 8 * (nh + d)) + bpu)
        sweep_bytes = per_step * prob.M * prob.steps_per_chain
        roofline["sweep"] = {"algorithmic_bytes_per_unit": sweep_bytes // (prob.M * prob.steps_per_chain),
                             "achieved": sweep_bytes / (ms_total / a.steps * 1e-3) / 1e9, "frac": sweep_bytes / (ms_total / a.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the C ABI with HOST buffers: what crosses the boundary every step of the reference loop is the
    # accept step's Exp(1) draws (host RNG, as in the reference: H2D from pinned memory) and the per-(block, chain) ll and
    # accept flags that the user's loop reads back (D2H); paths stay on the device (the tutorials read them every 400th step)
    e2e = None
    if not a.no_e2e:
        nb_l = [len(r) for r, _ in prob.layouts]
        rng = np.random.default_rng(7)
        E_pinned = [torch.empty((nb, prob.M), dtype=torch.float64, pin_memory=True) for nb in nb_l]
        for t in E_pinned:
            t.copy_(torch.from_numpy(rng.exponential(size=tuple(t.shape))))
        E_np = [t.numpy() for t in E_pinned]
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, a.steps // 2)
        for it in range(a.warmup + a.steps, a.warmup + a.steps + n_e2e):
            l = it % nlay
            sweep(it, False, E_np[l])                                 # H2D of E inside dmt_accept_reject_path
            ll_host = ctx.get_ll(l, 0)                                # D2H
            acc_host = ctx.get_last_accept(l)                         # D2H
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": units_per_step * n_e2e / float(dt.item()), "unit": "guided EM steps/s",
               "h2d_bytes_per_step": int(np.mean(nb_l) * prob.M * 8), "d2h_bytes_per_step": int(np.mean(nb_l) * prob.M * 9 + 8 * (2 + np.mean(nb_l))),
               "steps": n_e2e, "note": "sweep with host-drawn E (pinned) + dmt_get_ll + dmt_get_last_accept + stats, wall clock, max over ranks"}

    out = {
        "metric": "guided path updates/sec (chains x EM steps/s, FP64)", "value": value, "unit": "guided EM steps/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[2] (C3): Lorenz 3-D, %d chains per GPU, %d obs intervals x %d EM steps, "
                               "BlockCollection %s blocks alternated, pCN rho=0.9, P=%d parameter/data sets per GPU"
                               % (prob.M, prob.K, int(prob.n_pts[0] - 1), "/".join(str(len(r)) for r, _ in prob.layouts), prob.P)
                   if a.config == "c3" else "config %s: model %s, %d chains per GPU, K=%d" % (a.config, _lib.MODEL_NAMES[prob.model], prob.M, prob.K),
                   "config_id": a.config, "chains_per_gpu": prob.M, "steps_per_chain": prob.steps_per_chain,
                   "l2": "inputs (paths %.1f GB + guiding term %.1f GB per GPU) are far larger than the 126 MB L2"
                         % (2 * 8 * ctx.S * (prob.d + prob.dw) * prob.M / 1e9, 8 * ctx.S * (prob.d * (prob.d + 1) // 2 + prob.d) * prob.P / 1e9),
                   "step": "one blocking sweep over one layout" if blocking else "draw + accept", "stats_allreduce": allreduce_kind},
        "roofline": roofline, "kernel_ms": kern_ms, "kernel_ms_by_layout": kern_ms_by_layout, "gpu_launches": launches_per_step * a.steps, "clocks": clocks,
        "last_stats": {"sum_ll": float(last[0]), "sum_ll_prop": float(last[1]), "accept_frac": float(np.sum(last[2:]) / (len(last[2:]) * prob.M * world))},
    }
    if e2e:
        out["e2e"] = e2e
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        nch = a.cpu_chains or min(512, 16 * (os.cpu_count() or 1))
        v, spt, thr, sample = cpu_arm(a.config, nch, sweeps=4, warm=1)
        out["cpu_baseline"] = {"value": v, "unit": "guided EM steps/s", "cores": thr, "kind": "port", "sample": sample,
                               "note": "C restatement of the reference algorithm (oracle/), not the Julia package: Julia is not installed"}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def reference_arm(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    nch = a.cpu_chains or min(512, 16 * (os.cpu_count() or 1))
    M = a.chains or {"c1": 1, "c2": 1024, "c3": 4096, "c4": 16384, "c5": 8192}[a.config]
    v, spt, thr, sample = cpu_arm(a.config, nch, sweeps=max(1, min(a.steps, 4)), warm=max(0, min(a.warmup, 1)))
    out = {"impl": "reference", "metric": "guided path updates/sec (chains x EM steps/s, FP64)", "value": v, "unit": "guided EM steps/s",
           "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": spt * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "bounded sample of config %s (%d of %d chains per step)" % (a.config, nch, M), "config_id": a.config},
           "cpu_baseline": {"value": v, "unit": "guided EM steps/s", "cores": thr, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "guided EM steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference arm = C restatement of the reference algorithm (oracle/, OpenMP); the Julia reference cannot run in this image"}
    print(json.dumps(out))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)
