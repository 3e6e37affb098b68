// tsit5_kernel.cuh — K1 with UPSTREAM's solver: adaptive Tsitouras 5(4), OrdinaryDiffEq's default step-size control, dense output
// on the path grid (optional mode, dmt_set_bwd_solver(ctx, DMT_K1_TSIT5, reltol, abstol)).
//
// GuidedProposals 0.1.0 solves the (H,F,c) system with OrdinaryDiffEq's Tsit5() (OrdinaryDiffEq 5.41.0,
// /root/reference/Manifest.toml:352-356; call sites /root/reference/src/sampling_unit.jl:60-66 and /root/reference/src/block.jl:104-110)
// and saves it on the imputation grid.  The default of this library is classical RK4 ON that grid (kernels.cuh, bwd_kernel), which is
// ~4 orders of magnitude closer to the ODE than a 5(4) pair at reltol 1e-3; this kernel exists so that a user who wants upstream's
// NUMBERS (its O(tolerance) guiding term) rather than the ODE's can have them without Julia-side H,F,c uploads.  It restates the
// published method (Tsitouras 2011) and the controller OrdinaryDiffEq 5.x documents as its default (the tests compare it with a
// line-by-line CPU twin in the test infrastructure) and cannot be claimed bit-equal to upstream.
//
// One thread = one (parameter set, block, side), like bwd_kernel.  The state is (H packed symmetric, F, c); off-diagonal entries of H
// count twice in the error norm so that it equals the norm over the full d x d matrix the reference integrates.  Exact-observation
// (blocking) intervals integrate (H,F,c) from H = I/eps like upstream does (the covariance form belongs to the RK4 mode).
// Not tuned: seven stage vectors live in local memory for d >= 3; the mode is for parity work, not for the hot loop.
#pragma once
#include "kernels.cuh"

namespace dmt {

struct Tsit5Args {
    int side_mask;
    double reltol, abstol;
    int *steps; // [2] accepted / rejected steps, summed over threads (diagnostics), or null
};

__constant__ double TS_A[7][6] = {
    {0},
    {0.161},
    {-0.008480655492356989, 0.335480655492357},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
__constant__ double TS_BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                                0.5823571654525552, -0.45808210592918697, 0.015151515151515152};
__constant__ double TS_R[7][4] = {
    {1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
    {0.0, 0.13169999999999998, -0.2234, 0.1017},
    {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
    {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
    {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
    {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
    {0.0, 1.5, -4.0, 2.5}};

template <class MD>
__global__ void __launch_bounds__(32) bwd_tsit5_kernel(const DevCtx cx, const LayoutDev ly, const Tsit5Args ta) {
    constexpr int D = MD::D, NH = D * (D + 1) / 2, NG = NH + D, NAUX = D * D + D + NH, N = NH + D + 1;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y, side = blockIdx.z;
    if (ps >= cx.P || !((ta.side_mask >> side) & 1)) return;
    const size_t P = cx.P;
    const int i0 = ly.i0[b], i1 = ly.i1[b];
    const bool last = ly.last[b] != 0;
    const double reltol = ta.reltol, abstol = ta.abstol;
    double y[N]; // H (packed), F, c — carried from one interval into the jump of the previous one
#pragma unroll
    for (int i = 0; i < N; i++) y[i] = 0.0;
    int n_acc = 0, n_rej = 0;

    // weighted RMS norm over the FULL matrix H (off-diagonal entries twice), F and c
    auto norm = [&](const double *e, const double *u0, const double *u1) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
            for (int j = i; j < D; j++) {
                const int q = sidx<D>(i, j);
                const double sc = abstol + fmax(fabs(u0[q]), fabs(u1[q])) * reltol, r = e[q] / sc;
                s += (i == j ? 1.0 : 2.0) * r * r;
            }
#pragma unroll
        for (int q = NH; q < N; q++) {
            const double sc = abstol + fmax(fabs(u0[q]), fabs(u1[q])) * reltol, r = e[q] / sc;
            s += r * r;
        }
        return sqrt(s / (double)(D * D + D + 1));
    };

    for (int k = i1; k >= i0; --k) {
        const int store = (k == i1 && !last) ? 1 : 0;
        const int slot = side ^ cx.parP[store][(size_t)k * P + ps];
        double Bm[D * D], beta[D], at[NH];
        {
            const double *ap = cx.aux[slot][store] + (size_t)k * NAUX * P + ps;
#pragma unroll
            for (int i = 0; i < D * D; i++) Bm[i] = ap[(size_t)i * P];
#pragma unroll
            for (int i = 0; i < D; i++) beta[i] = ap[(size_t)(D * D + i) * P];
#pragma unroll
            for (int i = 0; i < NH; i++) at[i] = ap[(size_t)(D * D + D + i) * P];
        }
        const int nst = cx.nsteps[k], t0 = cx.tile0[k];
        const int gt0 = store ? cx.ppb_tile0[k] : t0;
        double *Gp = cx.G[slot][store] + ((size_t)gt0 * NG * P + ps) * 4;
        const size_t gstr = P * 4;
        const double *dtp = cx.dt + (size_t)t0 * 4;
        // dy/ds, s = T - t (the filter runs backward in t)
        auto rhs = [&](const double *u, double *du) {
            double dc;
            hfc_rhs<D, MD::ATIL_DIAG>(Bm, beta, at, u, u + NH, du, du + NH, dc);
#pragma unroll
            for (int i = 0; i < NG; i++) du[i] = -du[i];
            du[NG] = -dc;
        };
        // ---- values at the interval end
        if (store) { // exact artificial observation: H = I/eps, F = v/eps, c = (d log 2pi + d log eps + v'v/eps)/2
            double vv = 0.0;
#pragma unroll
            for (int i = 0; i < NH; i++) y[i] = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) {
                const double v = cx.vart[slot][((size_t)k * D + i) * P + ps];
                y[sidx<D>(i, i)] = 1.0 / cx.eps;
                y[NH + i] = v / cx.eps;
                vv = fma(v, v, vv);
            }
            y[NG] = 0.5 * (D * 1.8378770664093453 + D * log(cx.eps) + vv / cx.eps);
        } else {
            obs_jump<D>(cx.m, cx.obs[slot] + (size_t)k * (cx.m * D + cx.m * cx.m + cx.m) * P + ps, P, y, y + NH, y[NG]);
        }
        // grid distances from the interval end: sgrid(j) = T - t[j] = sum_{i >= j} dt[i]
        double S = 0.0;
        for (int j = 0; j < nst; j++) S += dtp[j];
        double K[7][N], ynew[N], tmp[N], err[N];
        rhs(y, K[0]);
        double dt;
        {   // Hairer's initial step (OrdinaryDiffEq ode_determine_initdt)
            const double d0 = norm(y, y, y), d1 = norm(K[0], y, y);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
            dt0 = fmin(dt0, S);
#pragma unroll
            for (int i = 0; i < N; i++) tmp[i] = fma(dt0, K[0][i], y[i]);
            rhs(tmp, K[1]);
#pragma unroll
            for (int i = 0; i < N; i++) err[i] = K[1][i] - K[0][i];
            const double d2 = norm(err, y, y) / dt0, dm = fmax(d1, d2);
            const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
            dt = fmin(fmin(100.0 * dt0, dt1), S);
        }
        const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qsmin = 1.0, qsmax = 1.2;
        double s = 0.0, qold = 1e-4;
        int next = nst - 1;        // next grid point to save; its distance from the interval end:
        double snext = dtp[nst - 1];
        int guard = 0;
        while (next >= 0 && guard++ < 2000000) {
            bool lastst = false;
            if (s + dt >= S * (1.0 - 1e-14)) { dt = S - s; lastst = true; }
#pragma unroll
            for (int st = 1; st < 7; st++) {
#pragma unroll
                for (int i = 0; i < N; i++) {
                    double a = 0.0;
#pragma unroll
                    for (int j = 0; j < st; j++) a += TS_A[st][j] * K[j][i];
                    tmp[i] = y[i] + dt * a;
                }
                if (st == 6) {
#pragma unroll
                    for (int i = 0; i < N; i++) ynew[i] = tmp[i];
                }
                rhs(tmp, K[st]);
            }
#pragma unroll
            for (int i = 0; i < N; i++) {
                double e = 0.0;
#pragma unroll
                for (int j = 0; j < 7; j++) e += TS_BT[j] * K[j][i];
                err[i] = dt * e;
            }
            double EEst = norm(err, y, ynew), q, q11 = 0.0;
            if (!(EEst == EEst)) EEst = 1e300;
            if (EEst == 0.0) q = 1.0 / qmax;
            else {
                q11 = pow(EEst, beta1);
                q = q11 / pow(qold, beta2);
                q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / gamma));
            }
            if (EEst <= 1.0) {
                while (next >= 0 && snext <= s + dt + 1e-15 * S) { // dense output on every grid point inside (s, s + dt]
                    double th = (lastst && next == 0) ? 1.0 : (snext - s) / dt;
                    th = fmin(th, 1.0);
                    double bth[7];
#pragma unroll
                    for (int j = 0; j < 7; j++) bth[j] = th * (TS_R[j][0] + th * (TS_R[j][1] + th * (TS_R[j][2] + th * TS_R[j][3])));
                    double *gp = Gp + (size_t)(next >> 2) * NG * gstr + (next & 3);
                    double cval = 0.0;
#pragma unroll
                    for (int i = 0; i < N; i++) {
                        double a = 0.0;
#pragma unroll
                        for (int j = 0; j < 7; j++) a += bth[j] * K[j][i];
                        const double v = (th == 1.0) ? ynew[i] : y[i] + dt * a;
                        if (i < NG) gp[(size_t)i * gstr] = v;
                        else cval = v;
                    }
                    if (next == 0) cx.c0[slot][store][(size_t)k * P + ps] = cval;
                    next--;
                    if (next >= 0) snext += dtp[next];
                }
                s += dt;
#pragma unroll
                for (int i = 0; i < N; i++) { y[i] = ynew[i]; K[0][i] = K[6][i]; } // FSAL
                if (q >= qsmin && q <= qsmax) q = 1.0;
                dt = dt / q;
                qold = fmax(EEst, 1e-4);
                n_acc++;
            } else {
                dt = dt / fmin(1.0 / qmin, q11 / gamma);
                n_rej++;
            }
        }
    }
    if (ta.steps) { atomicAdd(&ta.steps[0], n_acc); atomicAdd(&ta.steps[1], n_rej); }
}

} // namespace dmt
