// sweep_spec_kernel.cuh — the blocking sweep's fused forward pass for SMALL ensembles: warp-specialised.
//
// Same arithmetic and reference calls as sweep_pipe_kernel / fwd_kernel<MD, OP_SWEEP>
//   find_W_for_X!(b); loglikhd!(b); draw_proposal_path!(bb)   (/root/reference/docs/src/tutorials/block_collection/inference_with_blocking.md:55-57,
//   /root/reference/src/block.jl:120-152, /root/reference/src/biblock.jl:80-106).
// When the ensemble is split over 8 GPUs (BASELINE.json north_star: 512 chains x 10 blocks per GPU = 168 warps of (chain, block)
// units for 592 warp schedulers) the sweep is bound by the latency of ONE warp walking through a 4-step tile: generator (~970
// instructions), inverse solve + accepted-path likelihood (~330) and the proposal recursion (~400), one after the other — 6,700 cycles
// per tile measured (profiles/r02_tuning.md).  Only the proposal recursion is sequential in time.  Here a CTA of four warps shares one
// group of 32 chains and the three pieces run side by side, one tile apart, with ONE block barrier per tile:
//   warp 0 "A":  tile t+1:  X (register double buffer) and H,F (TMA ring) -> K5: the accepted noise dW of the tile's four steps, which
//                           do not depend on each other, + the accepted path's log-likelihood  -> dW into shared memory
//   warps 1,2 "R": tile t+1: the tile's normals, half of the Philox / Box-Muller calls each                   -> xi into shared memory
//   warp 3 "P":  tile t:    pCN refresh + guided Euler-Maruyama recursion + proposal log-likelihood from dW, xi, H,F  -> X° (and W°)
// Random stream and results are those of the other two kernels (bit-identical normals; FP64 rounding of a different FMA contraction).
#pragma once
#include "sweep_kernel.cuh"

namespace dmt {

constexpr int SPEC_STAGES = 4; // H,F ring: tile t (P), t+1 (A), t+2 and t+3 in flight

template <class MD> constexpr size_t sweep_spec_smem() {
    constexpr int D = MD::D, DW = MD::DW, NG = D * (D + 1) / 2 + D;
    return (size_t)SPEC_STAGES * (NG * 128 + 8) * 8 + (size_t)2 * 2 * (4 * DW) * 32 * 8 + SPEC_STAGES * 8;
}

template <class MD, bool LAZYW>
__global__ void __launch_bounds__(128, 4) sweep_spec_kernel(const DevCtx cx, const LayoutDev ly, const FwdArgs fa) {
    constexpr int D = MD::D, DW = MD::DW, NPAR = MD::NPAR, NH = D * (D + 1) / 2, NG = NH + D, NAUX = D * D + D + NH;
    constexpr int STAGE = NG * 128 + 8, NZ = 4 * DW;
    extern __shared__ __align__(128) unsigned char sp_smem[];
    double *ring = reinterpret_cast<double *>(sp_smem);          // [SPEC_STAGES][STAGE]
    double *dwb = ring + SPEC_STAGES * STAGE;                      // [2][NZ][32]  accepted noise of a tile (A -> P)
    double *zb = dwb + 2 * NZ * 32;                                // [2][NZ][32]  standard normals of a tile (R -> P)
    uint64_t *bars = reinterpret_cast<uint64_t *>(zb + 2 * NZ * 32);

    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;    // 0: A, 1-2: R, 3: P
    const int c0 = blockIdx.x * 32, b = blockIdx.y;
    if (c0 >= cx.M) return;
    const int c_raw = c0 + lane;
    const int c = min(c_raw, cx.M - 1);
    const bool live = c_raw < cx.M;
    const size_t M = cx.M, P = cx.P;
    const int ps = c;
    const int i0 = ly.i0[b], i1 = ly.i1[b];
    const bool last = ly.last[b] != 0;
    const uint32_t chunk_bytes = 32u * (uint32_t)min(32, cx.M - c0);
    const size_t gstr = P * 4;

    if (threadIdx.x == 0)
        for (int s = 0; s < SPEC_STAGES; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    int T = 0; // tiles of this block
    for (int k = i0; k <= i1; k++) T += (cx.nsteps[k] + 3) >> 2;

    // a cursor over the block's tiles: (interval, tile inside it)
    struct Cur { int k, q, ntl, t0; };
    auto cur_init = [&]() { Cur u; u.k = i0; u.q = 0; u.ntl = (cx.nsteps[i0] + 3) >> 2; u.t0 = cx.tile0[i0]; return u; };
    auto cur_next = [&](Cur &u) {
        if (++u.q == u.ntl) {
            u.q = 0;
            if (++u.k <= i1) { u.ntl = (cx.nsteps[u.k] + 3) >> 2; u.t0 = cx.tile0[u.k]; }
        }
    };

    if (role == 0) {
        // ================================================================== A: TMA producer, K5, accepted-path log-likelihood
        Cur pf = cur_init(); // prefetch cursor of the H,F ring
        int n_pf = 0;
        auto issue = [&]() {
            if (n_pf >= T) return;
            if (lane == 0) {
                const double *gp = g_tile_of<NG>(cx, ly, pf.k, i1, last, 0, c0).base;
                uint64_t *bar = &bars[n_pf % SPEC_STAGES];
                double *dst = ring + (size_t)(n_pf % SPEC_STAGES) * STAGE;
                mbar_expect_tx(bar, NG * chunk_bytes + 64u);
#pragma unroll
                for (int a = 0; a < NG; a++) bulk_g2s(dst + a * 128, gp + ((size_t)pf.q * NG + a) * gstr, chunk_bytes, bar);
                bulk_g2s(dst + NG * 128, cx.dt + (size_t)(pf.t0 + pf.q) * 4, 32u, bar);
                bulk_g2s(dst + NG * 128 + 4, cx.sqdt + (size_t)(pf.t0 + pf.q) * 4, 32u, bar);
            }
            n_pf++;
            cur_next(pf);
        };
        Cur xc = cur_init(); // cursor of the X register prefetch
        int n_x = 0;
        double xa[D][4], xb[D][4];
        auto load_x = [&](double (&dst)[D][4]) {
            if (n_x >= T) return;
            const double *xin = cx.X + (size_t)cx.parX[(size_t)xc.k * M + c] * cx.Xbuf + ((size_t)xc.t0 * D * M + c) * 4;
#pragma unroll
            for (int i = 0; i < D; i++) ld256(xin + ((size_t)xc.q * D + i) * M * 4, dst[i]);
            n_x++;
            cur_next(xc);
        };
        issue(); issue(); issue();        // tiles 0, 1, 2 (tile 3 follows after the first barrier)
        load_x(xa); load_x(xb);           // tiles 0, 1
        Cur u = cur_init();
        double x[D], ll = 0.0;
        {
            const int sl = cx.parX[(size_t)i0 * M + c];
#pragma unroll
            for (int i = 0; i < D; i++) x[i] = cx.X0[sl * cx.X0buf + ((size_t)i0 * D + i) * M + c];
        }
        double th[NPAR], Bm[D * D], beta[D], at[NH];
        int k_loaded = -1, nst = 0;
        uint8_t pw = 0;
        for (int t = 0; t < T; t++) { // phase t-1 of the block: the accepted side of tile t
            if (u.k != k_loaded) { // a new interval: its law record, its start point
                const GTile<NG> gt = g_tile_of<NG>(cx, ly, u.k, i1, last, 0, ps);
                const double *tp = cx.theta[gt.slot][gt.store] + (size_t)u.k * NPAR * P + ps;
#pragma unroll
                for (int i = 0; i < NPAR; i++) th[i] = tp[(size_t)i * P];
                const double *ap = cx.aux[gt.slot][gt.store] + (size_t)u.k * NAUX * P + ps;
#pragma unroll
                for (int i = 0; i < D * D; i++) Bm[i] = ap[(size_t)i * P];
#pragma unroll
                for (int i = 0; i < D; i++) beta[i] = ap[(size_t)(D * D + i) * P];
                if (!MD::CONSTDIFF) {
#pragma unroll
                    for (int i = 0; i < NH; i++) at[i] = ap[(size_t)(D * D + D + i) * P];
                }
                nst = cx.nsteps[u.k];
                pw = cx.parW[(size_t)u.k * M + c];
                if (u.k > i0) {
                    const uint8_t px = cx.parX[(size_t)u.k * M + c];
#pragma unroll
                    for (int i = 0; i < D; i++) x[i] = cx.X0[(size_t)px * cx.X0buf + ((size_t)u.k * D + i) * M + c];
                }
                k_loaded = u.k;
            }
            const typename MD::Par par(th);
            mbar_wait(&bars[t % SPEC_STAGES], (uint32_t)(t / SPEC_STAGES) & 1u);
            const double *st = ring + (size_t)(t % SPEC_STAGES) * STAGE;
            const double *sg = st + lane * 4;
            if (t == 0) { // loglikhd_obs(PP[1], y1) = -c - y'Hy/2 + F'y  (src/block.jl:178)
                const GTile<NG> gt = g_tile_of<NG>(cx, ly, i0, i1, last, 0, ps);
                double s0 = -*gt.c0;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    double hx = 0.0;
#pragma unroll
                    for (int j = 0; j < D; j++) hx = fma(sg[sidx<D>(i, j) * 128], x[j], hx);
                    s0 += x[i] * (sg[(NH + i) * 128] - 0.5 * hx);
                }
                ll = s0;
            }
            double (&xt)[D][4] = (t & 1) ? xb : xa;
            double *dwo = dwb + (size_t)(t & 1) * NZ * 32;
            double w[LAZYW ? 1 : DW][4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int i = 4 * u.q + s;
                if (i < nst) {
                    double Hs[NH], F[D], gd[D], G = 0.0, xn[D], res[D], dwv[DW];
#pragma unroll
                    for (int a = 0; a < NH; a++) Hs[a] = sg[a * 128 + s];
#pragma unroll
                    for (int a = 0; a < D; a++) F[a] = sg[(NH + a) * 128 + s];
                    const double dt = st[NG * 128 + s];
                    const typename MD::Diff df(par, x);
                    guided_terms<MD, true>(par, df, Bm, beta, at, Hs, F, x, gd, G);
                    ll = fma(G, dt, ll);
#pragma unroll
                    for (int a = 0; a < D; a++) xn[a] = xt[a][s];
#pragma unroll
                    for (int a = 0; a < D; a++) res[a] = xn[a] - x[a] - gd[a] * dt; // K5: dW = sigma^+ (x' - x - (b + a r) dt)   (A.5)
                    df.inv_sig(res, dwv);
#pragma unroll
                    for (int a = 0; a < D; a++) x[a] = xn[a];
#pragma unroll
                    for (int j = 0; j < DW; j++) {
                        dwo[(s * DW + j) * 32 + lane] = dwv[j];
                        if (!LAZYW) w[j][s] = dwv[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < DW; j++) {
                        dwo[(s * DW + j) * 32 + lane] = 0.0;
                        if (!LAZYW) w[j][s] = 0.0;
                    }
                }
            }
            if (!LAZYW && live) {
                double *Wacc = cx.W + (size_t)pw * cx.Wbuf + ((size_t)u.t0 * DW * M + c) * 4;
#pragma unroll
                for (int j = 0; j < DW; j++) st256(Wacc + ((size_t)u.q * DW + j) * M * 4, w[j]);
            }
            if (t & 1) load_x(xb); else load_x(xa); // tile t+2 into the buffer just consumed
            cur_next(u);
            __syncthreads();                          // ---- end of phase t-1: dW(t) is visible; stage (t-1) % 4 is free
            issue();                                  // tile t+3
        }
        __syncthreads();                              // the last phase (P works on tile T-1)
        if (live) ly.ll[(size_t)b * M + c] = ll;
    } else if (role <= 2) {
        // ================================================================== R: the tile's normals, half of the generator calls per warp
        Cur u = cur_init();
        for (int t = 0; t < T; t++) {
            double *zo = zb + (size_t)(t & 1) * NZ * 32;
            const uint32_t k0 = (uint32_t)cx.seed, k1 = (uint32_t)(cx.seed >> 32);
#pragma unroll
            for (int cc = 0; cc < DW; cc++) {
                const int call = (role - 1) * DW + cc;
                u32x4 ctr = {cx.chain_offset + (uint32_t)c, (uint32_t)(u.t0 + u.q), fa.iter, ctr_word3(STREAM_PCN, (uint32_t)ly.id, (uint32_t)call)};
                double z0, z1;
                box_muller(philox4x32_10(ctr, k0, k1), z0, z1);
                zo[(2 * call) * 32 + lane] = z0;
                zo[(2 * call + 1) * 32 + lane] = z1;
            }
            cur_next(u);
            __syncthreads();
        }
        __syncthreads();
    } else {
        // ================================================================== P: pCN refresh + guided EM recursion + proposal log-likelihood
        Cur u = cur_init();
        double xo[D], llo = 0.0;
        bool ok = true;
        {
            const int sl = cx.parX[(size_t)i0 * M + c];
#pragma unroll
            for (int i = 0; i < D; i++) xo[i] = cx.X0[sl * cx.X0buf + ((size_t)i0 * D + i) * M + c];
        }
        const double rho = ly.rho[b], crho = sqrt(1.0 - rho * rho);
        double th[NPAR], Bm[D * D], beta[D], at[NH];
        int k_loaded = -1, nst = 0;
        uint8_t pw = 0, px = 0;
        __syncthreads(); // phase -1: A and R prepare tile 0
        for (int t = 0; t < T; t++) {
            if (u.k != k_loaded) {
                const GTile<NG> gt = g_tile_of<NG>(cx, ly, u.k, i1, last, 0, ps);
                const double *tp = cx.theta[gt.slot][gt.store] + (size_t)u.k * NPAR * P + ps;
#pragma unroll
                for (int i = 0; i < NPAR; i++) th[i] = tp[(size_t)i * P];
                const double *ap = cx.aux[gt.slot][gt.store] + (size_t)u.k * NAUX * P + ps;
#pragma unroll
                for (int i = 0; i < D * D; i++) Bm[i] = ap[(size_t)i * P];
#pragma unroll
                for (int i = 0; i < D; i++) beta[i] = ap[(size_t)(D * D + i) * P];
                if (!MD::CONSTDIFF) {
#pragma unroll
                    for (int i = 0; i < NH; i++) at[i] = ap[(size_t)(D * D + D + i) * P];
                }
                nst = cx.nsteps[u.k];
                pw = cx.parW[(size_t)u.k * M + c];
                px = cx.parX[(size_t)u.k * M + c];
                if (live) { // XX°[k].x[1] = y1
                    double *x0p = cx.X0 + (size_t)(1 ^ px) * cx.X0buf + (size_t)u.k * D * M + c;
#pragma unroll
                    for (int i = 0; i < D; i++) x0p[(size_t)i * M] = xo[i];
                }
                k_loaded = u.k;
            }
            const typename MD::Par par(th);
            mbar_wait(&bars[t % SPEC_STAGES], (uint32_t)(t / SPEC_STAGES) & 1u); // (landed a phase ago: A waited for it too)
            const double *st = ring + (size_t)(t % SPEC_STAGES) * STAGE;
            const double *sg = st + lane * 4;
            const double *dwi = dwb + (size_t)(t & 1) * NZ * 32, *zi = zb + (size_t)(t & 1) * NZ * 32;
            if (t == 0) { // the same start term as the accepted path: same law, same start point
                const GTile<NG> gt = g_tile_of<NG>(cx, ly, i0, i1, last, 0, ps);
                double s0 = -*gt.c0;
#pragma unroll
                for (int i = 0; i < D; i++) {
                    double hx = 0.0;
#pragma unroll
                    for (int j = 0; j < D; j++) hx = fma(sg[sidx<D>(i, j) * 128], xo[j], hx);
                    s0 += xo[i] * (sg[(NH + i) * 128] - 0.5 * hx);
                }
                llo = s0;
            }
            double wo[LAZYW ? 1 : DW][4], xot[D][4];
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int i = 4 * u.q + s;
                if (i < nst) {
                    double Hs[NH], F[D], dwo[DW], swo[D], gdo[D], Go = 0.0, xon[D];
#pragma unroll
                    for (int a = 0; a < NH; a++) Hs[a] = sg[a * 128 + s];
#pragma unroll
                    for (int a = 0; a < D; a++) F[a] = sg[(NH + a) * 128 + s];
                    const double dt = st[NG * 128 + s], sq = st[NG * 128 + 4 + s];
#pragma unroll
                    for (int j = 0; j < DW; j++) { // K3: dW° = rho dW + sqrt(1-rho^2) sqrt(dt) xi   (A.2)
                        dwo[j] = rho * dwi[(s * DW + j) * 32 + lane] + crho * sq * zi[(s * DW + j) * 32 + lane];
                        if (!LAZYW) wo[j][s] = dwo[j];
                    }
                    const typename MD::Diff dfo(par, xo);
                    guided_terms<MD, true>(par, dfo, Bm, beta, at, Hs, F, xo, gdo, Go);
                    llo = fma(Go, dt, llo);
                    dfo.sig_mul(dwo, swo);
#pragma unroll
                    for (int a = 0; a < D; a++) xon[a] = fma(gdo[a], dt, xo[a]) + swo[a];
                    bool fin = dfo.ok();
#pragma unroll
                    for (int a = 0; a < D; a++) fin = fin && isfinite(xon[a]);
                    ok = ok && fin && MD::bound_ok(par, xon); // src/block.jl:181 (ll° := -Inf once, after the loop)
#pragma unroll
                    for (int a = 0; a < D; a++) { xot[a][s] = xon[a]; xo[a] = xon[a]; }
                } else {
#pragma unroll
                    for (int a = 0; a < D; a++) xot[a][s] = 0.0;
                    if (!LAZYW) {
#pragma unroll
                        for (int j = 0; j < DW; j++) wo[j][s] = 0.0;
                    }
                }
            }
            if (live) {
                if (!LAZYW) {
                    double *Wprop = cx.W + (size_t)(1 ^ pw) * cx.Wbuf + ((size_t)u.t0 * DW * M + c) * 4;
#pragma unroll
                    for (int j = 0; j < DW; j++) st256(Wprop + ((size_t)u.q * DW + j) * M * 4, wo[j]);
                }
                double *Xout = cx.X + (size_t)(1 ^ px) * cx.Xbuf + ((size_t)u.t0 * D * M + c) * 4;
#pragma unroll
                for (int i = 0; i < D; i++) st256(Xout + ((size_t)u.q * D + i) * M * 4, xot[i]);
            }
            cur_next(u);
            __syncthreads(); // ---- end of phase t
        }
        if (live) {
            ly.ll[((size_t)ly.nb + b) * M + c] = ok ? llo : -INFINITY;
            ly.ok[(size_t)b * M + c] = ok ? 1 : 0;
        }
    }
}

} // namespace dmt
