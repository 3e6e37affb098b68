// kernels.cuh — hand-written sm_100a kernels of the guided-proposal path update.
//
// DATA LAYOUT IN HBM.  Time is cut into TILES of 4 Euler–Maruyama steps that never straddle an
// observation interval.  Every streamed array is  [tile][component][chain or pset][4 steps]  so that ONE LANE OWNS
// ONE 32-BYTE DRAM SECTOR per component per tile and moves it with a single 256-bit LDG/STG.  This is the
// [time][dim][chain] structure-of-arrays of the north star with the time axis blocked by 4: loads/stores of a warp
// are still contiguous (32 lanes x 32 B = 1 KiB), but accepted/proposal buffer selection (per-chain parity bits,
// the SoA form of the reference's pointer swaps, /root/reference/src/biblock.jl:148-173) and pset gathers become
// sector-exact: no over-fetch on reads, no read-modify-write on writes.
//   X  [2][NT][D ][M][4]   X[..][s] = state AFTER step 4*tile+s  (XX[k].x[2:end] of the reference)
//   X0 [2][K ][D ][M]      first point of every interval          (XX[k].x[1])
//   W  [2][NT][DW][M][4]   Wiener increments                      (WW[k])
//   G  [slot][store] [NT or NTb][NH+D][P][4]   guiding term H (packed symmetric), F at the LEFT point of each step
//   c0 [slot][store] [K][P]                     c at the interval start
// parX/parW [K][M], parP [store][K][P]: which physical buffer/slot currently holds the ACCEPTED object.
#pragma once
#include "models.cuh"
#include "philox.cuh"
#include <stdint.h>

namespace dmt {

// tuning knobs (defaults = the measured best, profiles/; overridable with -D for experiments)
#ifndef DMT_FWD_TPB
#define DMT_FWD_TPB 64
#endif
#ifndef DMT_FWD_MINB
#define DMT_FWD_MINB 0 // min resident CTAs per SM promised to ptxas (caps registers); 0 = per-model default (fwd_minb)
#endif
#ifndef DMT_EVICT_FIRST
#define DMT_EVICT_FIRST 1 // 1: streamed sectors are marked L2::evict_first (spill lines and per-interval records stay in L2)
#endif
#ifndef DMT_BWD_MINB
#define DMT_BWD_MINB 1
#endif
#ifndef DMT_BWD_TPB
#define DMT_BWD_TPB 32
#endif
constexpr int FWD_TPB = DMT_FWD_TPB;
constexpr int BWD_TPB = DMT_BWD_TPB;

struct DevCtx {
    int M, P, K, NT, NTb, m, two_sided;
    const int *tile0;     // [K+1] first tile of interval k
    const int *step0;     // [K+1] first (unpadded) step of interval k
    const int *pt0;       // [K+1] first (unpadded) point of interval k
    const int *nsteps;    // [K]
    const int *ppb_tile0; // [K] first tile in the PPb store, -1 if interval k never ends a non-terminal block
    const double *dt, *sqdt; // [NT*4] (padded like the tiles)
    const int *pset;      // [M]
    const int *k_of_tile; // [NT] interval that owns tile t of the PP store
    double *X, *W, *X0;
    size_t Xbuf, Wbuf, X0buf; // buffer strides in doubles
    uint8_t *parX, *parW;     // [K][M]
    uint8_t *parP[2];         // [store] [K][P]
    double *G[2][2];          // [slot][store]
    double *c0[2][2];         // [slot][store] [K][P]
    double *theta[2][2];      // [slot][store] [K][NPAR][P]
    double *aux[2][2];        // [slot][store] [K][NAUX][P]   B (row-major d*d), beta (d), atilde (packed NH)
    double *obs[2];           // [slot] [K][NOBS][P]          L (m*d), Sigma (m*m), v (m)
    double *vart[2];          // [slot] [K][D][P]             artificial exact observation of the PPb law
    double eps;
    uint64_t seed;
    uint32_t chain_offset;
};

struct LayoutDev {
    int nb, id;
    const int *i0, *i1;
    const uint8_t *last;
    const double *rho;
    double *ll;        // [2][nb][M]
    uint8_t *ok;       // [nb][M] success of the last forward op
    uint8_t *last_acc; // [nb][M]
    uint8_t *acc_hist; // [hist_len][nb][M] or null
    double *ll_hist;   // [hist_len][2][nb][M] or null
    int hist_len;
    // layout-private guiding term of the ACCEPTED laws (dmt_enable_guiding_cache): null => the shared store cx.G / cx.c0
    double *Gl[2];     // [store] tiles like cx.G[slot][store]
    double *c0l[2];    // [store] [K][P]
    // guiding cache: (F, c) of a block are exactly affine / quadratic in the block's artificial end-point observation v
    double *FP[2];     // [store] [tiles / FPG][D + D*D][P][FPG][4]: F0 (v = 0) then Psi[i][m] = dF_i/dv_m  (fp_off below)
    double *cq;        // [nb][1 + D + NH][P]: c = c0 + q.v + v'Qv/2 at the block start
    double *v_last;    // [nb][D][P]: the artificial observation the private F, c were last materialised for
    const int *blk_of_k; // [K] block of this layout that contains interval k
};

struct BwdArgs {
    int side_mask;
    int use_override; // 1: every non-terminal block uses v[] as its artificial observation (cache build probes)
    double v[6];
};

enum { OP_DRAW = 0, OP_RECOMPUTE = 1, OP_LOGLIK = 2, OP_INVSOLVE = 3, OP_INVSOLVE_LL = 4, OP_INIT = 5, OP_SWEEP = 6 };

struct FwdArgs {
    uint32_t iter;
    int law_side, w_side, skip;
    const double *Z; // device, [S][DW][M] standard normals or null
    int lazy_w;      // OP_SWEEP: do not store W_acc / W° (dmt_set_lazy_noise; set by the launcher, never by callers)
};

// ------------------------------------------------------------------------------------------- 256-bit sector access
__device__ __forceinline__ void ld256(const double *p, double *v) { // streaming read-only sector
#if defined(DMT_LD256_SPLIT) // (experiment: the sector as two 128-bit loads)
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "l"(p + 2));
#elif DMT_EVICT_FIRST
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#endif
}
__device__ __forceinline__ void ld256u(const double *p, double *v) { // chain-uniform data (dt, sqrt dt): keep in L1
#if defined(DMT_DT_SCALAR)
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
#else
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#endif
}
__device__ __forceinline__ void st256(double *p, const double *v) {
#if DMT_EVICT_FIRST
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory");
#endif
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------- one EM step's terms
// r = F - H x ; guided drift gd = b + a r (A.3) ; integrand G = (b - btilde).r [- tr((a-at)H)/2 + r'(a-at)r/2] (A.4)
template <class MD, bool WANT_G>
__device__ __forceinline__ void guided_terms(const typename MD::Par &par, const typename MD::Diff &df, const double *Bm,
                                             const double *beta, const double *at, const double *Hs, const double *F,
                                             const double *x, double *gd, double &G) {
    constexpr int D = MD::D, NH = D * (D + 1) / 2;
    double r[D], b[D], ar[D];
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = F[i];
#pragma unroll
        for (int j = 0; j < D; j++) s = fma(-Hs[sidx<D>(i, j)], x[j], s);
        r[i] = s;
    }
    MD::drift(par, x, b);
    df.a_mul(r, ar);
#pragma unroll
    for (int i = 0; i < D; i++) gd[i] = b[i] + ar[i];
    if (WANT_G) {
        double g = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) {
            double bt = beta[i];
#pragma unroll
            for (int j = 0; j < D; j++) bt = fma(Bm[i * D + j], x[j], bt);
            g = fma(b[i] - bt, r[i], g);
        }
        if (!MD::CONSTDIFF) {
            double a[NH];
            df.a_sym(a);
            double tr = 0.0, q = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++)
#pragma unroll
                for (int j = 0; j < D; j++) {
                    double da = a[sidx<D>(i, j)] - at[sidx<D>(i, j)];
                    tr = fma(da, Hs[sidx<D>(i, j)], tr);
                    q = fma(r[i] * da, r[j], q);
                }
            g += -0.5 * tr + 0.5 * q;
        }
        G = g;
    }
}

// (the forward kernel K2/K3/K4/K5 lives in fwd_kernel.cuh)

// =========================================================================================== K1 backward filter
template <int D, bool DIAG>
__device__ __forceinline__ void hfc_rhs(const double *Bm, const double *beta, const double *at, const double *H, const double *F,
                                        double *dH, double *dF, double &dc) {
    // dH = -B'H - HB + H at H ; dF = -B'F + H at F + H beta ; dc = beta'F + F' at F/2 - tr(H at)/2   (A.1), arranged as
    //   N = at H,  C = B - N/2,  C2 = B - N :   dH = -(HC + (HC)'),   dF = H beta - C2'F      (H, at symmetric)
    // which needs one d^3 product instead of three; DIAG: at is diagonal, N is a row scaling of H.
    double C[D][D], C2[D][D], tr = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            double n;
            if (DIAG) n = at[sidx<D>(k, k)] * H[sidx<D>(k, j)];
            else {
                n = 0.0;
#pragma unroll
                for (int l = 0; l < D; l++) n = fma(at[sidx<D>(k, l)], H[sidx<D>(l, j)], n);
            }
            C2[k][j] = Bm[k * D + j] - n;
            C[k][j] = fma(-0.5, n, Bm[k * D + j]);
            if (k == j) tr += n;
        }
    double Mx[D][D];
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s = fma(H[sidx<D>(i, k)], C[k][j], s);
            Mx[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = i; j < D; j++) dH[sidx<D>(i, j)] = -(Mx[i][j] + Mx[j][i]);
    double bF = 0.0, FaF = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            s = fma(H[sidx<D>(i, k)], beta[k], s);
            s = fma(-C2[k][i], F[k], s);
        }
        dF[i] = s;
        bF = fma(beta[i], F[i], bF);
        if (DIAG) FaF = fma(at[sidx<D>(i, i)] * F[i], F[i], FaF);
        else {
            double aF = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) aF = fma(at[sidx<D>(i, k)], F[k], aF);
            FaF = fma(F[i], aF, FaF);
        }
    }
    dc = bF + 0.5 * FaF - 0.5 * tr;
}

template <int D>
__device__ __forceinline__ void pnu_rhs(const double *Bm, const double *beta, const double *at, const double *Pm, const double *nu,
                                        double *dP, double *dnu) {
    // dP = B P + P B' - at ; dnu = B nu + beta   (covariance form used on exact-observation intervals)
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = i; j < D; j++) {
            double s = -at[sidx<D>(i, j)];
#pragma unroll
            for (int k = 0; k < D; k++) {
                s = fma(Bm[i * D + k], Pm[sidx<D>(k, j)], s);
                s = fma(Bm[j * D + k], Pm[sidx<D>(k, i)], s);
            }
            dP[sidx<D>(i, j)] = s;
        }
        double s = beta[i];
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(Bm[i * D + k], nu[k], s);
        dnu[i] = s;
    }
}

// inverse and log-determinant of a packed SPD matrix via Cholesky
template <int D> __device__ __forceinline__ bool spd_inverse(const double *A, double *Ai, double &logdet) {
    double L[D][D], Li[D][D];
    bool good = true;
    double ld = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        double s = A[sidx<D>(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
        good = good && (s > 0.0);
        const double l = sqrt(s), il = 1.0 / l;
        L[j][j] = l;
        ld += log(l);
#pragma unroll
        for (int i = j + 1; i < D; i++) {
            double t = A[sidx<D>(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
            L[i][j] = t * il;
        }
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
        Li[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < D; i++) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; k++) s -= L[i][k] * Li[k][j];
            Li[i][j] = s / L[i][i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = i; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < D; k++) s = fma(Li[k][i], Li[k][j], s);
            Ai[sidx<D>(i, j)] = s;
        }
    logdet = 2.0 * ld;
    return good;
}

// jump at a (partial, Gaussian) observation: H += L' S^-1 L ; F += L' S^-1 v ; c += (m log 2pi + log det S + v' S^-1 v)/2
template <int D>
__device__ __forceinline__ void obs_jump(int m, const double *op, size_t P, double *H, double *F, double &c) {
    // op -> [NOBS][P] record of this (k, pset): L (m*D), Sigma (m*m), v (m)
    double Lm[D][D], Sg[D][D], v[D], Lc[D][D];
#pragma unroll
    for (int a = 0; a < D; a++) {
#pragma unroll
        for (int j = 0; j < D; j++) Lm[a][j] = (a < m) ? op[(size_t)(a * D + j) * P] : 0.0;
#pragma unroll
        for (int bq = 0; bq < D; bq++) Sg[a][bq] = (a < m && bq < m) ? op[(size_t)(m * D + a * m + bq) * P] : (a == bq ? 1.0 : 0.0);
        v[a] = (a < m) ? op[(size_t)(m * D + m * m + a) * P] : 0.0;
    }
    double ld = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) { // Cholesky of Sigma (identity-padded beyond m)
        double s = Sg[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) s -= Lc[j][k] * Lc[j][k];
        const double l = sqrt(s);
        Lc[j][j] = l;
        if (j < m) ld += log(l);
#pragma unroll
        for (int i = j + 1; i < D; i++) {
            double t = Sg[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= Lc[i][k] * Lc[j][k];
            Lc[i][j] = t / l;
        }
    }
    // Y = Lc^-1 L, y = Lc^-1 v (forward substitution, row by row)
    double Y[D][D], y[D];
#pragma unroll
    for (int a = 0; a < D; a++) {
        const double il = 1.0 / Lc[a][a];
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = Lm[a][j];
#pragma unroll
            for (int k = 0; k < a; k++) s -= Lc[a][k] * Y[k][j];
            Y[a][j] = s * il;
        }
        double s = v[a];
#pragma unroll
        for (int k = 0; k < a; k++) s -= Lc[a][k] * y[k];
        y[a] = s * il;
    }
    double yy = 0.0;
#pragma unroll
    for (int a = 0; a < D; a++) yy = fma(y[a], y[a], yy);
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = i; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < D; a++) s = fma(Y[a][i], Y[a][j], s);
            H[sidx<D>(i, j)] += s;
        }
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++) s = fma(Y[a][i], y[a], s);
        F[i] += s;
    }
    c += 0.5 * (m * 1.8378770664093453 /* log 2pi */ + 2.0 * ld + yy);
}

// recompute_guiding_term!(b::Block)  src/block.jl:104-110.  One thread = one (pset, block, side).
// grid = (ceil(P/TPB), n_blocks, 2 sides).  Regular intervals: classical RK4 on the path grid for (H,F,c), backward,
// with the observation jump fused at the interval end.  The last interval of a non-terminal block (the PPb law with
// its exact artificial observation, Sigma = eps I) is integrated in covariance form (P = H^-1, nu = P F), which is a
// linear non-stiff ODE, and converted to (H,F) per grid point; c has a closed form there.
template <class MD>
__global__ void __launch_bounds__(BWD_TPB, DMT_BWD_MINB) bwd_kernel(const DevCtx cx, const LayoutDev ly, const BwdArgs ba) {
    const int side_mask = ba.side_mask;
    constexpr int D = MD::D, NH = D * (D + 1) / 2, NG = NH + D, NAUX = D * D + D + NH;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y, side = blockIdx.z;
    if (ps >= cx.P || !((side_mask >> side) & 1)) return;
    const size_t P = cx.P;
    const int i0 = ly.i0[b], i1 = ly.i1[b];
    const bool last = ly.last[b] != 0;
    double H[NH], F[D], cc = 0.0;
#pragma unroll
    for (int i = 0; i < NH; i++) H[i] = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) F[i] = 0.0;

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
        const bool priv = (side == 0) && (ly.Gl[store] != nullptr); // layout-private accepted-law store (guiding cache)
        double *Gp = (priv ? ly.Gl[store] : cx.G[slot][store]) + ((size_t)gt0 * NG * P + ps) * 4;
        const size_t gstr = P * 4;
        const double *dtp = cx.dt + (size_t)t0 * 4;

        // one grid step back of each formulation, as lambdas so that the time loop below can be written twice (see BUF)
        double Pm[NH], nu[D], trB = 0.0, Tt = 0.0, logdet = 0.0;
        auto step_pnu = [&](double h) { // covariance form on an exact-observation interval: RK4 on (P, nu), then H = P^-1, F = H nu
            Tt += h;
            double k1P[NH], k1n[D], kP[NH], kn[D], Ps[NH], ns[D], aP[NH], an[D];
            pnu_rhs<D>(Bm, beta, at, Pm, nu, k1P, k1n);
#pragma unroll
            for (int i = 0; i < NH; i++) { Ps[i] = fma(-0.5 * h, k1P[i], Pm[i]); aP[i] = k1P[i]; }
#pragma unroll
            for (int i = 0; i < D; i++) { ns[i] = fma(-0.5 * h, k1n[i], nu[i]); an[i] = k1n[i]; }
            pnu_rhs<D>(Bm, beta, at, Ps, ns, kP, kn);
#pragma unroll
            for (int i = 0; i < NH; i++) { Ps[i] = fma(-0.5 * h, kP[i], Pm[i]); aP[i] = fma(2.0, kP[i], aP[i]); }
#pragma unroll
            for (int i = 0; i < D; i++) { ns[i] = fma(-0.5 * h, kn[i], nu[i]); an[i] = fma(2.0, kn[i], an[i]); }
            pnu_rhs<D>(Bm, beta, at, Ps, ns, kP, kn);
#pragma unroll
            for (int i = 0; i < NH; i++) { Ps[i] = fma(-h, kP[i], Pm[i]); aP[i] = fma(2.0, kP[i], aP[i]); }
#pragma unroll
            for (int i = 0; i < D; i++) { ns[i] = fma(-h, kn[i], nu[i]); an[i] = fma(2.0, kn[i], an[i]); }
            pnu_rhs<D>(Bm, beta, at, Ps, ns, kP, kn);
#pragma unroll
            for (int i = 0; i < NH; i++) Pm[i] = fma(-h / 6.0, aP[i] + kP[i], Pm[i]);
#pragma unroll
            for (int i = 0; i < D; i++) nu[i] = fma(-h / 6.0, an[i] + kn[i], nu[i]);
            spd_inverse<D>(Pm, H, logdet);
#pragma unroll
            for (int i = 0; i < D; i++) {
                double sacc = 0.0;
#pragma unroll
                for (int q = 0; q < D; q++) sacc = fma(H[sidx<D>(i, q)], nu[q], sacc);
                F[i] = sacc;
            }
        };
        auto step_hfc = [&](double h) { // classical RK4 on (H, F, c) from t[j+1] back to t[j]
            double kH[NH], kF[D], kc, aH[NH], aF[D], ac, Hs[NH], Fs[D];
            hfc_rhs<D, MD::ATIL_DIAG>(Bm, beta, at, H, F, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < NH; i++) { Hs[i] = fma(-0.5 * h, kH[i], H[i]); aH[i] = kH[i]; }
#pragma unroll
            for (int i = 0; i < D; i++) { Fs[i] = fma(-0.5 * h, kF[i], F[i]); aF[i] = kF[i]; }
            ac = kc;
            hfc_rhs<D, MD::ATIL_DIAG>(Bm, beta, at, Hs, Fs, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < NH; i++) { Hs[i] = fma(-0.5 * h, kH[i], H[i]); aH[i] = fma(2.0, kH[i], aH[i]); }
#pragma unroll
            for (int i = 0; i < D; i++) { Fs[i] = fma(-0.5 * h, kF[i], F[i]); aF[i] = fma(2.0, kF[i], aF[i]); }
            ac = fma(2.0, kc, ac);
            hfc_rhs<D, MD::ATIL_DIAG>(Bm, beta, at, Hs, Fs, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < NH; i++) { Hs[i] = fma(-h, kH[i], H[i]); aH[i] = fma(2.0, kH[i], aH[i]); }
#pragma unroll
            for (int i = 0; i < D; i++) { Fs[i] = fma(-h, kF[i], F[i]); aF[i] = fma(2.0, kF[i], aF[i]); }
            ac = fma(2.0, kc, ac);
            hfc_rhs<D, MD::ATIL_DIAG>(Bm, beta, at, Hs, Fs, kH, kF, kc);
            const double h6 = h / 6.0;
#pragma unroll
            for (int i = 0; i < NH; i++) H[i] = fma(-h6, aH[i] + kH[i], H[i]);
#pragma unroll
            for (int i = 0; i < D; i++) F[i] = fma(-h6, aF[i] + kF[i], F[i]);
            cc = fma(-h6, ac + kc, cc);
        };
        if (store) { // exact artificial observation (guid_prop_for_blocking, src/sampling_unit.jl:61-66)
#pragma unroll
            for (int i = 0; i < D; i++)
#pragma unroll
                for (int j = i; j < D; j++) Pm[sidx<D>(i, j)] = (i == j) ? cx.eps : 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) nu[i] = ba.use_override ? ba.v[i] : cx.vart[slot][((size_t)k * D + i) * P + ps];
#pragma unroll
            for (int i = 0; i < D; i++) trB += Bm[i * D + i];
        } else {
            obs_jump<D>(cx.m, cx.obs[slot] + (size_t)k * (cx.m * D + cx.m * cx.m + cx.m) * P + ps, P, H, F, cc);
        }
        // BUF: collect the 4 grid points of a tile in registers and write whole 32-byte sectors (one STG.256 per component)
        // instead of 4 partial-sector stores; needs the time loop unrolled by 4, affordable for the small models only.
        constexpr bool BUF = (NG <= 14);
        if (BUF) {
            for (int q = ((nst + 3) >> 2) - 1; q >= 0; --q) {
                double tb[NG][4], dt4[4];
                ld256u(dtp + (size_t)q * 4, dt4);
#pragma unroll
                for (int sl = 3; sl >= 0; --sl) {
                    if (4 * q + sl < nst) {
                        if (store) step_pnu(dt4[sl]); else step_hfc(dt4[sl]);
#pragma unroll
                        for (int i = 0; i < NH; i++) tb[i][sl] = H[i];
#pragma unroll
                        for (int i = 0; i < D; i++) tb[NH + i][sl] = F[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < NG; i++) tb[i][sl] = 0.0;
                    }
                }
#pragma unroll
                for (int i = 0; i < NG; i++) st256(Gp + ((size_t)q * NG + i) * gstr, tb[i]);
            }
        } else {
            for (int j = nst - 1; j >= 0; --j) {
                if (store) step_pnu(dtp[j]); else step_hfc(dtp[j]);
                double *gp = Gp + (size_t)(j >> 2) * NG * gstr + (j & 3);
#pragma unroll
                for (int i = 0; i < NH; i++) gp[(size_t)i * gstr] = H[i];
#pragma unroll
                for (int i = 0; i < D; i++) gp[(size_t)(NH + i) * gstr] = F[i];
            }
        }
        if (store) { // c = d/2 log 2pi + log det P / 2 + tr(B)(T-t) + nu'H nu / 2 at the interval start
            double nF = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) nF = fma(nu[i], F[i], nF);
            cc = 0.5 * D * 1.8378770664093453 + 0.5 * logdet + trB * Tt + 0.5 * nF;
        }
        (priv ? ly.c0l[store] : cx.c0[slot][store])[(size_t)k * P + ps] = cc;
    }
}

// =========================================================================================== K1, row-cooperative variant
// For wide states and few (pset, block) pairs (BASELINE C5: Jansen–Rit d = 6, one terminal block) one thread per pset is the
// wrong mapping: a 28-component Riccati system per thread spills heavily and only P threads exist.  Here D consecutive lanes
// form a group that owns ONE pset: lane r keeps row r of the symmetric H and F_r in registers; one gather of H through
// warp shuffles per RK4 stage feeds both products of  dH = -(HC + (HC)'),  C = B - a H / 2  (a diagonal):
//     (HC)_{rj} = sum_k H_rk C_kj          and, by symmetry of H,          (HC)_{jr} = sum_k H_kj C_kr
// so no transpose exchange is needed.  32/D psets per warp, 32x more threads, same arithmetic per component as bwd_kernel
// (results agree to FP64 rounding).  Terminal blocks of diagonal-a models only (what C5 needs); others use bwd_kernel.
// SPARSE: B carries the structural zeros of the model's Jacobian (JacMask): the product skips them at compile time.
template <class MD, bool SPARSE = false>
__global__ void __launch_bounds__(128) bwd_coop_kernel(const DevCtx cx, const LayoutDev ly, const BwdArgs ba) {
    constexpr int D = MD::D, NH = D * (D + 1) / 2, NG = NH + D, NAUX = D * D + D + NH, GPW = 32 / D;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane / D, r = lane % D;
    const int base = g * D;
    const int ps_raw = (blockIdx.x * (blockDim.x >> 5) + wid) * GPW + g;
    const int b = blockIdx.y, side = blockIdx.z;
    if (!((ba.side_mask >> side) & 1)) return;
    const bool active = (g < GPW) && (ps_raw < cx.P);
    const int ps = min(ps_raw, cx.P - 1); // idle lanes shadow a valid pset (they take part in the shuffles, never store)
    const size_t P = cx.P;
    const int i0 = ly.i0[b], i1 = ly.i1[b];
    double Hrow[D], Fr = 0.0, cc = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) Hrow[j] = 0.0;
    // exchange buffer of the RHS: every lane publishes its row of H and its F, then reads the whole (H, F) of its parameter
    // set back (broadcast reads, conflict-free across the groups of a warp).  Two buffers alternate, so one __syncwarp per
    // evaluation is enough.  (Replaces D^2 + D double shuffles per evaluation: the kernel was bound by the shuffle pipe.)
    constexpr int GSZ = D * D + D;
    __shared__ __align__(16) double xch[4][2][(32 / D + 1) * GSZ];   // rows of C = B - aH/2 and F
    __shared__ __align__(16) double mch[4][2][(32 / D + 1) * D * D]; // rows of M = HC
    int xbuf = 0;

    for (int k = i1; k >= i0; --k) {
        const int slot = side ^ cx.parP[0][(size_t)k * P + ps];
        double Brow[D], Bcol[D], beta[D], ad[D], adr, trB = 0.0;
        {
            const double *ap = cx.aux[slot][0] + (size_t)k * NAUX * P + ps;
#pragma unroll
            for (int i = 0; i < D; i++) {
                Brow[i] = ap[(size_t)(r * D + i) * P]; // row / column r of B: lane-specific VALUES, static register indices
                Bcol[i] = ap[(size_t)(i * D + r) * P];
                beta[i] = ap[(size_t)(D * D + i) * P];
                ad[i] = ap[(size_t)(D * D + D + sidx<D>(i, i)) * P];
                trB += ap[(size_t)(i * D + i) * P];
            }
            adr = ap[(size_t)(D * D + D + r * D - r * (r - 1) / 2) * P]; // a_rr
        }
        const int nst = cx.nsteps[k], t0 = cx.tile0[k];
        const bool priv = (side == 0) && (ly.Gl[0] != nullptr);
        double *Gp = (priv ? ly.Gl[0] : cx.G[slot][0]) + ((size_t)t0 * NG * P + ps) * 4;
        const size_t gstr = P * 4;
        const double *dtp = cx.dt + (size_t)t0 * 4;

        {   // jump at the observation: H += L'S^-1 L, F += L'S^-1 v, c += (m log 2pi + log det S + v'S^-1 v)/2; lane r does row r
            const int m = cx.m;
            const double *op = cx.obs[slot] + (size_t)k * (m * D + m * m + m) * P + ps;
            double Sg[D][D], Lc[D][D], Lcol[D], v[D], Ycol[D], y[D];
#pragma unroll
            for (int a = 0; a < D; a++) {
                Lcol[a] = (a < m) ? op[(size_t)(a * D + r) * P] : 0.0;
#pragma unroll
                for (int bq = 0; bq < D; bq++) Sg[a][bq] = (a < m && bq < m) ? op[(size_t)(m * D + a * m + bq) * P] : (a == bq ? 1.0 : 0.0);
                v[a] = (a < m) ? op[(size_t)(m * D + m * m + a) * P] : 0.0;
            }
            double ld = 0.0;
#pragma unroll
            for (int j = 0; j < D; j++) {
                double s = Sg[j][j];
#pragma unroll
                for (int q = 0; q < j; q++) s -= Lc[j][q] * Lc[j][q];
                const double l = sqrt(s);
                Lc[j][j] = l;
                if (j < m) ld += log(l);
#pragma unroll
                for (int i = j + 1; i < D; i++) {
                    double t = Sg[i][j];
#pragma unroll
                    for (int q = 0; q < j; q++) t -= Lc[i][q] * Lc[j][q];
                    Lc[i][j] = t / l;
                }
            }
            double yy = 0.0;
#pragma unroll
            for (int a = 0; a < D; a++) { // forward substitution on column r of L and on v
                const double il = 1.0 / Lc[a][a];
                double s = Lcol[a], sv = v[a];
#pragma unroll
                for (int q = 0; q < a; q++) { s -= Lc[a][q] * Ycol[q]; sv -= Lc[a][q] * y[q]; }
                Ycol[a] = s * il;
                y[a] = sv * il;
                yy = fma(y[a], y[a], yy);
            }
#pragma unroll
            for (int j = 0; j < D; j++) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < D; a++) s = fma(Ycol[a], __shfl_sync(FULL, Ycol[a], base + j), s);
                Hrow[j] += s;
            }
#pragma unroll
            for (int a = 0; a < D; a++) Fr = fma(Ycol[a], y[a], Fr);
            cc += 0.5 * (m * 1.8378770664093453 + 2.0 * ld + yy);
        }

        // RHS with the symmetry of Hdot = -(M + M'), M = HC, C = B - aH/2: lane r publishes ITS row of C, reads all of C,
        // forms its row of M, publishes it, and reads column r of M back (two exchanges, ~40 % fewer DFMAs than forming both
        // HC and (HC)' locally); tr(aH) = 2 (tr B - tr C).
        auto rhs = [&](const double *Hr, double F_r, double *dHr, double &dFr, double &dc) {
            double Mx[D], Fall[D], trC = 0.0;
            double *blk = &xch[wid][xbuf][g * GSZ], *mb = &mch[wid][xbuf][g * D * D];
            xbuf ^= 1;
#pragma unroll
            for (int j = 0; j < D; j++) blk[r * D + j] = fma(-0.5 * adr, Hr[j], Brow[j]); // C_rj
            blk[D * D + r] = F_r;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < D; j++) { Mx[j] = 0.0; Fall[j] = blk[D * D + j]; }
#pragma unroll
            for (int q = 0; q < D; q++) {
#pragma unroll
                for (int j = 0; j < D; j++) {
                    if (SPARSE && !JacMask<MD>::nz(q, j)) continue; // a structural zero of C (compile-time: q, j are unrolled)
                    const double cqj = blk[q * D + j];
                    Mx[j] = fma(Hr[q], cqj, Mx[j]); // (HC)_{rj}
                    if (j == q) trC += cqj;
                }
            }
#pragma unroll
            for (int j = 0; j < D; j++) mb[r * D + j] = Mx[j];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < D; j++) dHr[j] = -(Mx[j] + mb[j * D + r]); // (HC)_{rj} + (HC)_{jr}
            double s = 0.0, bF = 0.0, FaF = 0.0;
#pragma unroll
            for (int q = 0; q < D; q++) {
                s = fma(Hr[q], beta[q], s);
                s = fma(-fma(-ad[q], Hr[q], Bcol[q]), Fall[q], s); // - (B - aH)_{qr} F_q
                bF = fma(beta[q], Fall[q], bF);
                FaF = fma(ad[q] * Fall[q], Fall[q], FaF);
            }
            dFr = s;
            dc = bF + 0.5 * FaF - (trB - trC);
        };

        double tb[D + 1][4] = {}; // this lane's components of the current 4-step tile: (r, r..D-1) of H and F_r -> 256-bit stores
        for (int j = nst - 1; j >= 0; --j) { // classical RK4 from t[j+1] back to t[j], row r of H, F_r, c (replicated)
            const double h = dtp[j];
            double kH[D], kF, kc, aH[D], aF, ac, Hs[D], Fs;
            rhs(Hrow, Fr, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < D; i++) { Hs[i] = fma(-0.5 * h, kH[i], Hrow[i]); aH[i] = kH[i]; }
            Fs = fma(-0.5 * h, kF, Fr); aF = kF; ac = kc;
            rhs(Hs, Fs, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < D; i++) { Hs[i] = fma(-0.5 * h, kH[i], Hrow[i]); aH[i] = fma(2.0, kH[i], aH[i]); }
            Fs = fma(-0.5 * h, kF, Fr); aF = fma(2.0, kF, aF); ac = fma(2.0, kc, ac);
            rhs(Hs, Fs, kH, kF, kc);
#pragma unroll
            for (int i = 0; i < D; i++) { Hs[i] = fma(-h, kH[i], Hrow[i]); aH[i] = fma(2.0, kH[i], aH[i]); }
            Fs = fma(-h, kF, Fr); aF = fma(2.0, kF, aF); ac = fma(2.0, kc, ac);
            rhs(Hs, Fs, kH, kF, kc);
            const double h6 = h / 6.0;
#pragma unroll
            for (int i = 0; i < D; i++) Hrow[i] = fma(-h6, aH[i] + kH[i], Hrow[i]);
            Fr = fma(-h6, aF + kF, Fr);
            cc = fma(-h6, ac + kc, cc);
#pragma unroll
            for (int sl = 0; sl < 4; sl++) {
                if ((j & 3) == sl) {
#pragma unroll
                    for (int jj = 0; jj < D; jj++) tb[jj][sl] = Hrow[jj];
                    tb[D][sl] = Fr;
                }
            }
            if ((j & 3) == 0 && active) { // the tile is complete (slots past the interval end hold don't-care values)
                double *gp = Gp + (size_t)(j >> 2) * NG * gstr;
#pragma unroll
                for (int jj = 0; jj < D; jj++)
                    if (jj >= r) st256(gp + (size_t)(r * D - r * (r - 1) / 2 + (jj - r)) * gstr, tb[jj]); // packed upper (r, jj)
                st256(gp + (size_t)(NH + r) * gstr, tb[D]);
            }
        }
        if (active && r == 0) (priv ? ly.c0l[0] : cx.c0[slot][0])[(size_t)k * P + ps] = cc;
    }
}

// =========================================================================================== guiding cache (K1 fast path)
// In a smoothing sweep with blocking only the frozen end point v of each non-terminal block changes between two calls of
// recompute_guiding_term!(be, Val(:P_only)) (src/biblock.jl:275-278 then src/block.jl:104-110).  H does not depend on v, and the
// RK4-discretised F and c are EXACTLY affine resp. quadratic in v (the F equation is linear given the H stages, the jump
// and the covariance-form start are affine in v).  So per layout we keep F0 = F(v=0), Psi = dF/dv and (c0, q, Q) of
// c = c0 + q.v + v'Qv/2 — obtained by probing the unchanged bwd_kernel with v = 0, +-s e_m, s(e_m+e_n) — and a sweep's K1
// becomes the streaming update F = F0 + Psi v (120 B per step instead of ~320 FP64 instructions per step).
// Layout of F0, Psi.  cache_apply_kernel SKIPS whole (parameter set, block) pairs, so neighbouring parameter sets are streamed
// independently of each other — and DRAM / L2 move 64-byte sector pairs: with one 32-byte sector per (tile, component, parameter set)
// next to its NEIGHBOUR's sector, a moved block next to an unmoved one fetched both (measured, profiles/r02ag_summary.md: 7.1 GB read
// where ~3.2 GB were needed).  So FPG consecutive tiles of ONE parameter set are stored next to each other: a lane owns a whole
// 128-byte line per component and tile group, every fetched byte belongs to a block that moved.
#ifndef DMT_FPG
#define DMT_FPG 4
#endif
constexpr int FPG = DMT_FPG;
__host__ __device__ __forceinline__ size_t fp_tiles_padded(size_t nt) { return (nt + FPG - 1) / FPG * FPG; }
__device__ __forceinline__ size_t fp_off(int t, int comp, int NF, size_t P, int ps) { // doubles; the sector of (tile t, component, pset)
    return ((((size_t)(t / FPG) * NF + comp) * P + ps) * FPG + (t % FPG)) * 4;
}
__device__ __forceinline__ bool cache_tile_of(const DevCtx &cx, const LayoutDev &ly, int store, int k, int &b) {
    b = ly.blk_of_k[k];
    const bool is_plast = (k == ly.i1[b]) && !ly.last[b]; // this interval's law comes from PPb in this layout
    return store ? is_plast : !is_plast;
}
// after a probe run: mode 0: F0 := F ; mode 1: Psi[:, m] := (F - F0) * inv_scale.   grid (ceil(P/128), tiles of `store`)
template <int D>
__global__ void cache_extract_kernel(const DevCtx cx, const LayoutDev ly, int store, const int *k_of_t, int mode, int m, double inv_scale) {
    constexpr int NH = D * (D + 1) / 2, NG = NH + D, NF = D + D * D;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (ps >= cx.P) return;
    int b;
    if (!cache_tile_of(cx, ly, store, k_of_t[t], b)) return;
    const size_t P = cx.P;
    const double *gp = ly.Gl[store] + (((size_t)t * NG + NH) * P + ps) * 4;
    double *fp = ly.FP[store];
#pragma unroll
    for (int i = 0; i < D; i++) {
        double f[4], f0[4];
        ld256(gp + (size_t)i * P * 4, f);
        if (mode == 0) {
            st256(fp + fp_off(t, i, NF, P, ps), f);
        } else {
            ld256(fp + fp_off(t, i, NF, P, ps), f0);
#pragma unroll
            for (int s = 0; s < 4; s++) f[s] = (f[s] - f0[s]) * inv_scale;
            st256(fp + fp_off(t, D + i * D + m, NF, P, ps), f);
        }
    }
}
// the per-sweep K1 of a cached layout: F = F0 + Psi v for every tile of every non-terminal block whose end point moved.
// A CTA owns (128 / FPG parameter sets, block, slice z of the block's tile groups).  Its first warp makes the test "did this block's end
// point move?" once per parameter set and COMPACTS the ones that did (ballot + popc) into shared memory together with their v; the CTA's
// threads then walk the flattened items (tile group, moved parameter set, tile of the group), so every warp runs with all lanes active
// whatever fraction moved (ncu, profiles/r02ak_summary.md: with one static lane group per parameter set the pass reached the DRAM roof only
// when ~85 % had moved — the bytes in flight scaled with the active lanes) and reads whole 128-byte lines of parameter sets that moved.
// The block's intervals i0..i1-1 live in store 0, its last interval (the PPb law) in store 1.    grid (ceil(P * FPG / 128), nb, Z)
template <int D>
__global__ void __launch_bounds__(128) cache_apply_kernel(const DevCtx cx, const LayoutDev ly) {
    constexpr int NH = D * (D + 1) / 2, NG = NH + D, NF = D + D * D;
    constexpr int NPS = 128 / FPG; // parameter sets per CTA
    static_assert(NPS <= 32 && NPS * FPG == 128, "one warp compacts the CTA's parameter sets");
    __shared__ double vs[NPS][D];
    __shared__ int psl[NPS];
    __shared__ int n_moved;
    const int b = blockIdx.y;
    if (ly.last[b]) return; // terminal block: no artificial observation
    const size_t P = cx.P;
    const int kend = ly.i1[b];
    if (threadIdx.x < 32) {
        const int ps = blockIdx.x * NPS + threadIdx.x;
        bool moved = false; // false: the end point did not move (the block's last proposal in the other layout was rejected), F is current
        double v[D];
        if (threadIdx.x < NPS && ps < cx.P) {
            const int sv = cx.parP[1][(size_t)kend * P + ps];
#pragma unroll
            for (int mm = 0; mm < D; mm++) {
                v[mm] = cx.vart[sv][((size_t)kend * D + mm) * P + ps];
                moved = moved || !(v[mm] == ly.v_last[((size_t)b * D + mm) * P + ps]);
            }
        }
        const unsigned mask = __ballot_sync(0xffffffffu, moved);
        if (moved) {
            const int j = __popc(mask & ((1u << threadIdx.x) - 1u));
            psl[j] = ps;
#pragma unroll
            for (int mm = 0; mm < D; mm++) vs[j][mm] = v[mm];
        }
        if (threadIdx.x == 0) n_moved = __popc(mask);
    }
    __syncthreads();
    const int nm = n_moved;
    if (nm == 0) return;
    const int per_g = nm * FPG; // items per tile group
    constexpr size_t cstr = (size_t)FPG * 4; // component stride / P of the F0, Psi store
    for (int store = 0; store < 2; store++) {
        int ta, tb;
        if (store == 0) { ta = cx.tile0[ly.i0[b]]; tb = cx.tile0[kend]; }
        else {
            ta = cx.ppb_tile0[kend];
            if (ta < 0) continue;
            tb = ta + ((cx.nsteps[kend] + 3) >> 2);
        }
        if (tb <= ta) continue;
        const int g0 = ta / FPG, n_items = ((tb - 1) / FPG - g0 + 1) * per_g;
        const double *fpb = ly.FP[store];
        double *gpb = ly.Gl[store] + (size_t)NH * P * 4;
        for (int id = (int)(blockIdx.z * 128 + threadIdx.x); id < n_items; id += 128 * (int)gridDim.z) {
            const int g = id / per_g, r = id - g * per_g, j = r / FPG;
            const int t = (g0 + g) * FPG + (r % FPG);
            if (t < ta || t >= tb) continue;
            const int ps = psl[j];
            double v[D];
#pragma unroll
            for (int mm = 0; mm < D; mm++) v[mm] = vs[j][mm];
            const double *fp = fpb + fp_off(t, 0, NF, P, ps);
            double *gp = gpb + ((size_t)t * NG * P + ps) * 4;
            constexpr int CB = D <= 3 ? D : 1; // components per batch: every load of a batch is in flight before the first FMA
#pragma unroll
            for (int i0 = 0; i0 < D; i0 += CB) {
                double f[CB][4], p[CB][D][4];
#pragma unroll
                for (int i = 0; i < CB; i++) {
                    ld256(fp + (size_t)(i0 + i) * P * cstr, f[i]);
#pragma unroll
                    for (int mm = 0; mm < D; mm++) ld256(fp + (size_t)(D + (i0 + i) * D + mm) * P * cstr, p[i][mm]);
                }
#pragma unroll
                for (int i = 0; i < CB; i++) {
#pragma unroll
                    for (int mm = 0; mm < D; mm++)
#pragma unroll
                        for (int s = 0; s < 4; s++) f[i][s] = fma(p[i][mm][s], v[mm], f[i][s]);
                    st256(gp + (size_t)(i0 + i) * P * 4, f[i]);
                }
            }
        }
    }
}
// c at the block start of probe run `run` -> Crun[run][nb][P]
__global__ void cache_collect_c_kernel(const DevCtx cx, const LayoutDev ly, int run, double *Crun) {
    const int ps = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (ps >= cx.P) return;
    Crun[((size_t)run * ly.nb + b) * cx.P + ps] = ly.c0l[0][(size_t)ly.i0[b] * cx.P + ps];
}
// (c0, q, Q) from the probes: run 0: v = 0; 1..D: +s e_m; D+1..2D: -s e_m; then s(e_m + e_n), m < n
template <int D>
__global__ void cache_solve_cq_kernel(const DevCtx cx, const LayoutDev ly, const double *Crun, double s) {
    constexpr int NH = D * (D + 1) / 2, NC = 1 + D + NH;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (ps >= cx.P) return;
    const size_t P = cx.P, str = (size_t)ly.nb * P, o = (size_t)b * P + ps;
    const double c0 = Crun[o];
    double q[D], Q[NH];
#pragma unroll
    for (int m = 0; m < D; m++) {
        const double cp = Crun[(size_t)(1 + m) * str + o], cm = Crun[(size_t)(1 + D + m) * str + o];
        q[m] = (cp - cm) / (2.0 * s);
        Q[sidx<D>(m, m)] = (cp + cm - 2.0 * c0) / (s * s);
    }
    int run = 1 + 2 * D;
#pragma unroll
    for (int m = 0; m < D; m++)
#pragma unroll
        for (int n = m + 1; n < D; n++) {
            const double cpair = Crun[(size_t)run * str + o];
            Q[sidx<D>(m, n)] = (cpair - c0 - s * (q[m] + q[n])) / (s * s) - 0.5 * (Q[sidx<D>(m, m)] + Q[sidx<D>(n, n)]);
            run++;
        }
    double *cq = ly.cq + (size_t)b * NC * P + ps;
    cq[0] = c0;
#pragma unroll
    for (int m = 0; m < D; m++) cq[(size_t)(1 + m) * P] = q[m];
#pragma unroll
    for (int a = 0; a < NH; a++) cq[(size_t)(1 + D + a) * P] = Q[a];
}
template <int D> __global__ void cache_apply_c_kernel(const DevCtx cx, const LayoutDev ly) {
    constexpr int NH = D * (D + 1) / 2, NC = 1 + D + NH;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (ps >= cx.P || ly.last[b]) return;
    const size_t P = cx.P;
    const int kend = ly.i1[b];
    const int sv = cx.parP[1][(size_t)kend * P + ps];
    double v[D];
#pragma unroll
    for (int m = 0; m < D; m++) v[m] = cx.vart[sv][((size_t)kend * D + m) * P + ps];
#pragma unroll
    for (int m = 0; m < D; m++) ly.v_last[((size_t)b * D + m) * P + ps] = v[m]; // (after both cache_apply_kernel launches, in stream order)
    const double *cq = ly.cq + (size_t)b * NC * P + ps;
    double c = cq[0];
#pragma unroll
    for (int m = 0; m < D; m++) {
        double Qv = 0.0;
#pragma unroll
        for (int n = 0; n < D; n++) Qv = fma(cq[(size_t)(1 + D + sidx<D>(m, n)) * P], v[n], Qv);
        c += v[m] * (cq[(size_t)(1 + m) * P] + 0.5 * Qv);
    }
    ly.c0l[0][(size_t)ly.i0[b] * P + ps] = c;
}

// auxiliary law := Jacobian linearisation of the target at xbar (last paragraph)
template <class MD>
__global__ void aux_linearise_kernel(const DevCtx cx, int slot_side, int store, int k0, int k1, const double *xbar /*[k1-k0+1][D][P]*/) {
    constexpr int D = MD::D, NPAR = MD::NPAR, NH = D * (D + 1) / 2, NAUX = D * D + D + NH;
    const int ps = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = k0 + blockIdx.y;
    if (ps >= cx.P || k > k1) return;
    const size_t P = cx.P;
    const int slot = slot_side ^ cx.parP[store][(size_t)k * P + ps];
    double th[NPAR], xb[D], J[D * D], b[D], a[NH];
#pragma unroll
    for (int i = 0; i < NPAR; i++) th[i] = cx.theta[slot][store][((size_t)k * NPAR + i) * P + ps];
#pragma unroll
    for (int i = 0; i < D; i++) xb[i] = xbar[((size_t)(k - k0) * D + i) * P + ps];
    const typename MD::Par par(th);
    MD::jac(par, xb, J);
    MD::drift(par, xb, b);
    const typename MD::Diff df(par, xb);
    df.a_sym(a);
    double *ap = cx.aux[slot][store] + (size_t)k * NAUX * P + ps;
#pragma unroll
    for (int i = 0; i < D * D; i++) ap[(size_t)i * P] = J[i];
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = b[i];
#pragma unroll
        for (int j = 0; j < D; j++) s -= J[i * D + j] * xb[j];
        ap[(size_t)(D * D + i) * P] = s;
    }
#pragma unroll
    for (int i = 0; i < NH; i++) ap[(size_t)(D * D + D + i) * P] = a[i];
}

// =========================================================================================== model-independent kernels

// K7  GP.set_obs!(bb)  src/biblock.jl:275-278: artificial obs of b.P_last AND b°.P_last := b.XX[end].x[end]
__global__ void set_artificial_obs_kernel(const DevCtx cx, const LayoutDev ly, int D) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= cx.M || ly.last[b]) return;
    const size_t M = cx.M, P = cx.P;
    const int k = ly.i1[b], ps = cx.pset[c];
    const int j = cx.nsteps[k] - 1;
    const int sl = cx.parX[(size_t)k * M + c];
    const double *xp = cx.X + (size_t)sl * cx.Xbuf + (((size_t)(cx.tile0[k] + (j >> 2)) * D) * M + c) * 4 + (j & 3);
    for (int i = 0; i < D; i++) {
        const double v = xp[(size_t)i * M * 4];
        cx.vart[0][((size_t)k * D + i) * P + ps] = v;
        if (cx.two_sided) cx.vart[1][((size_t)k * D + i) * P + ps] = v;
    }
}

// K6  accept_reject_proposal_path!(bb, i)  src/biblock.jl:121-127
__global__ void accept_kernel(const DevCtx cx, const LayoutDev ly, uint32_t iter, const double *E /*[nb][M] or null*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= cx.M) return;
    const size_t M = cx.M, idx = (size_t)b * M + c;
    const double ll = ly.ll[idx], llo = ly.ll[(size_t)ly.nb * M + idx];
    const double e = E ? E[idx] : accept_exponential(cx.seed, cx.chain_offset + (uint32_t)c, (uint32_t)b, iter, (uint32_t)ly.id);
    const bool acc = e > -(llo - ll); // IEEE: NaN => false => reject
    if (acc) {                        // swap_paths!: XX and WW of every interval of the block
        for (int k = ly.i0[b]; k <= ly.i1[b]; ++k) {
            cx.parX[(size_t)k * M + c] ^= 1;
            cx.parW[(size_t)k * M + c] ^= 1;
        }
    }
    ly.last_acc[idx] = acc;
    if (ly.acc_hist && iter < (uint32_t)ly.hist_len) { // set_accepted!, save_ll! BEFORE swap_ll!
        ly.acc_hist[(size_t)iter * ly.nb * M + idx] = acc;
        ly.ll_hist[((size_t)iter * 2 + 0) * ly.nb * M + idx] = ll;
        ly.ll_hist[((size_t)iter * 2 + 1) * ly.nb * M + idx] = llo;
    }
    if (acc) { ly.ll[idx] = llo; ly.ll[(size_t)ly.nb * M + idx] = ll; }
}

__global__ void save_ll_kernel(const DevCtx cx, const LayoutDev ly, uint32_t iter) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= cx.M || !ly.ll_hist || iter >= (uint32_t)ly.hist_len) return;
    const size_t M = cx.M, idx = (size_t)b * M + c;
    ly.ll_hist[((size_t)iter * 2 + 0) * ly.nb * M + idx] = ly.ll[idx];
    ly.ll_hist[((size_t)iter * 2 + 1) * ly.nb * M + idx] = ly.ll[(size_t)ly.nb * M + idx];
}

// swap_XX!/swap_WW!/swap_ll! (src/biblock.jl:158-173,206-208) per (chain, block)
// mask: null = every (chain, block); per_block == 0: [M], one flag per chain; per_block == 1: [nb][M], one flag per (block, chain) —
// the BiBlock-level swaps of one block of one recording (src/biblock.jl:148-209)
__global__ void swap_paths_kernel(const DevCtx cx, const LayoutDev ly, int what, const uint8_t *mask, int per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= cx.M || (mask && !mask[(per_block ? (size_t)b * cx.M : 0) + c])) return;
    const size_t M = cx.M, idx = (size_t)b * M + c;
    for (int k = ly.i0[b]; k <= ly.i1[b]; ++k) {
        if (what & 1) cx.parX[(size_t)k * M + c] ^= 1;
        if (what & 2) cx.parW[(size_t)k * M + c] ^= 1;
    }
    if (what & 8) {
        const double a = ly.ll[idx], o = ly.ll[(size_t)ly.nb * M + idx];
        ly.ll[idx] = o;
        ly.ll[(size_t)ly.nb * M + idx] = a;
    }
}
// swap_PP! (src/biblock.jl:182-199) per (pset, block): terminal: PP[i0..i1]; non-terminal: PP[i0..i1] (PP + P_excl) and
// PPb[i0..i1] (Pb_excl + P_last)
__global__ void swap_laws_kernel(const DevCtx cx, const LayoutDev ly, const uint8_t *mask, int per_block) {
    const int ps = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (ps >= cx.P || (mask && !mask[(per_block ? (size_t)b * cx.P : 0) + ps])) return;
    const size_t P = cx.P;
    for (int k = ly.i0[b]; k <= ly.i1[b]; ++k) {
        cx.parP[0][(size_t)k * P + ps] ^= 1;
        if (!ly.last[b]) cx.parP[1][(size_t)k * P + ps] ^= 1;
    }
}

// natural [point][dim][chain] <-> tiled, parity-resolved.  dir 0: device tiles -> out, 1: in -> device tiles
__global__ void xfer_X_kernel(const DevCtx cx, int side, int D, double *nat, int dir) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (c >= cx.M) return;
    const size_t M = cx.M;
    const int sl = side ^ cx.parX[(size_t)k * M + c];
    const int nst = cx.nsteps[k], p0 = cx.pt0[k], t0 = cx.tile0[k];
    for (int i = 0; i < D; i++) {
        double *x0 = cx.X0 + (size_t)sl * cx.X0buf + ((size_t)k * D + i) * M + c;
        double *np0 = nat + ((size_t)p0 * D + i) * M + c;
        if (dir) *x0 = *np0; else *np0 = *x0;
        for (int j = 0; j < nst; j++) {
            double *xp = cx.X + (size_t)sl * cx.Xbuf + (((size_t)(t0 + (j >> 2)) * D + i) * M + c) * 4 + (j & 3);
            double *np = nat + ((size_t)(p0 + j + 1) * D + i) * M + c;
            if (dir) *xp = *np; else *np = *xp;
        }
    }
}
// paths of a FEW chains (thinned path saving): device tiles -> out[point][dim][n_sel], parity-resolved like xfer_X_kernel
__global__ void gather_X_kernel(const DevCtx cx, int side, int D, const int *sel, int n_sel, double *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (s >= n_sel) return;
    const size_t M = cx.M;
    const int c = sel[s];
    const int sl = side ^ cx.parX[(size_t)k * M + c];
    const int nst = cx.nsteps[k], p0 = cx.pt0[k], t0 = cx.tile0[k];
    for (int i = 0; i < D; i++) {
        out[((size_t)p0 * D + i) * n_sel + s] = cx.X0[(size_t)sl * cx.X0buf + ((size_t)k * D + i) * M + c];
        for (int j = 0; j < nst; j++)
            out[((size_t)(p0 + j + 1) * D + i) * n_sel + s] =
                cx.X[(size_t)sl * cx.Xbuf + (((size_t)(t0 + (j >> 2)) * D + i) * M + c) * 4 + (j & 3)];
    }
}
// noise of a FEW chains: device tiles -> out[step][dw][n_sel], parity-resolved like xfer_W_kernel
__global__ void gather_W_kernel(const DevCtx cx, int side, int DW, const int *sel, int n_sel, double *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (s >= n_sel) return;
    const size_t M = cx.M;
    const int c = sel[s];
    const int sl = side ^ cx.parW[(size_t)k * M + c];
    const int nst = cx.nsteps[k], s0 = cx.step0[k], t0 = cx.tile0[k];
    for (int i = 0; i < DW; i++)
        for (int j = 0; j < nst; j++)
            out[((size_t)(s0 + j) * DW + i) * n_sel + s] = cx.W[(size_t)sl * cx.Wbuf + (((size_t)(t0 + (j >> 2)) * DW + i) * M + c) * 4 + (j & 3)];
}
__global__ void xfer_W_kernel(const DevCtx cx, int side, int DW, double *nat, int dir) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (c >= cx.M) return;
    const size_t M = cx.M;
    const int sl = side ^ cx.parW[(size_t)k * M + c];
    const int nst = cx.nsteps[k], s0 = cx.step0[k], t0 = cx.tile0[k];
    for (int i = 0; i < DW; i++)
        for (int j = 0; j < nst; j++) {
            double *wp = cx.W + (size_t)sl * cx.Wbuf + (((size_t)(t0 + (j >> 2)) * DW + i) * M + c) * 4 + (j & 3);
            double *np = nat + ((size_t)(s0 + j) * DW + i) * M + c;
            if (dir) *wp = *np; else *np = *wp;
        }
}
// u° = deepcopy(u)  (src/sampling_pair.jl:51): accepted X, X0, W -> proposal buffers, sector by sector
__global__ void copy_acc_to_prop_kernel(const DevCtx cx, int D, int DW) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (c >= cx.M) return;
    const size_t M = cx.M;
    const int sx = cx.parX[(size_t)k * M + c], sw = cx.parW[(size_t)k * M + c];
    const int t0 = cx.tile0[k], ntl = (cx.nsteps[k] + 3) >> 2;
    for (int i = 0; i < D; i++) {
        cx.X0[(size_t)(sx ^ 1) * cx.X0buf + ((size_t)k * D + i) * M + c] = cx.X0[(size_t)sx * cx.X0buf + ((size_t)k * D + i) * M + c];
        for (int q = 0; q < ntl; q++) {
            double v[4];
            const size_t o = (((size_t)(t0 + q) * D + i) * M + c) * 4;
            ld256(cx.X + (size_t)sx * cx.Xbuf + o, v);
            st256(cx.X + (size_t)(sx ^ 1) * cx.Xbuf + o, v);
        }
    }
    for (int i = 0; i < DW; i++)
        for (int q = 0; q < ntl; q++) {
            double v[4];
            const size_t o = (((size_t)(t0 + q) * DW + i) * M + c) * 4;
            ld256(cx.W + (size_t)sw * cx.Wbuf + o, v);
            st256(cx.W + (size_t)(sw ^ 1) * cx.Wbuf + o, v);
        }
}

// guiding term natural [n_k][d*d][P], [n_k][d][P], [n_k][P] <-> packed tiles of interval k.  dir 0: get, 1: upload
__global__ void xfer_guiding_kernel(const DevCtx cx, int side, int store, int k, int D, double *Hn, double *Fn, double *cn, int dir,
                                    double *G_priv, double *c0_priv) { // *_priv: a layout-private store (guiding cache) or null
    const int ps = blockIdx.x * blockDim.x + threadIdx.x;
    if (ps >= cx.P) return;
    const size_t P = cx.P;
    const int NH = D * (D + 1) / 2, NG = NH + D;
    const int slot = side ^ cx.parP[store][(size_t)k * P + ps];
    const int nst = cx.nsteps[k];
    const int gt0 = store ? cx.ppb_tile0[k] : cx.tile0[k];
    double *Gp = (G_priv ? G_priv : cx.G[slot][store]) + ((size_t)gt0 * NG * P + ps) * 4;
    for (int j = 0; j < nst; j++) {
        double *gp = Gp + (size_t)(j >> 2) * NG * P * 4 + (j & 3);
        for (int a = 0; a < D; a++) {
            for (int bq = 0; bq < D; bq++) {
                const int lo = a < bq ? a : bq, hi = a < bq ? bq : a;
                const int si = lo * D - lo * (lo - 1) / 2 + (hi - lo);
                double *hn = Hn + ((size_t)j * D * D + a * D + bq) * P + ps;
                if (dir) { if (a <= bq) gp[(size_t)si * P * 4] = *hn; } else *hn = gp[(size_t)si * P * 4];
            }
            double *fn = Fn + ((size_t)j * D + a) * P + ps;
            if (dir) gp[(size_t)(NH + a) * P * 4] = *fn; else *fn = gp[(size_t)(NH + a) * P * 4];
        }
    }
    double *c0 = &(c0_priv ? c0_priv : cx.c0[slot][store])[(size_t)k * P + ps];
    if (dir) *c0 = cn[ps];
    else {
        cn[ps] = *c0;
        for (int j = 1; j <= nst; j++) cn[(size_t)j * P + ps] = NAN; // only c at the interval start is kept
        for (int q = 0; q < D * D; q++) Hn[((size_t)nst * D * D + q) * P + ps] = NAN;
        for (int q = 0; q < D; q++) Fn[((size_t)nst * D + q) * P + ps] = NAN;
    }
}

// test hook: the device's Philox -> N(0,1) / Exp(1) streams for given counters (checked against the oracle's libm versions)
template <int DW>
__global__ void debug_normals_kernel(uint64_t seed, uint32_t chain0, uint32_t tile0, uint32_t iter, uint32_t layout, int n_chains, int n_tiles, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains * n_tiles) return;
    const int c = i / n_tiles, q = i % n_tiles;
    double z[4 * DW];
    tile_normals<DW>(seed, chain0 + (uint32_t)c, tile0 + (uint32_t)q, iter, layout, z);
    for (int k = 0; k < 4 * DW; k++) out[(size_t)i * 4 * DW + k] = z[k];
}
__global__ void debug_exponentials_kernel(uint64_t seed, uint32_t chain0, uint32_t iter, uint32_t layout, int n_chains, int n_blocks, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains * n_blocks) return;
    out[i] = accept_exponential(seed, chain0 + (uint32_t)(i / n_blocks), (uint32_t)(i % n_blocks), iter, layout);
}

// fetch_ll / fetch_ll° / accept counts (src/block_ensemble.jl:140,152,175-179): fixed-order tree reduction
// (warp shuffle + shared memory) over chains -> partial[3][nb][ncta] = (sum ll, sum ll°, #accepted of the last decision)
__global__ void __launch_bounds__(256) reduce_stats_kernel(const double *ll, const uint8_t *last_acc, int M, int nb, double *partial) {
    __shared__ double sm[3][8];
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    if (c < M) {
        v[0] = ll[(size_t)b * M + c];
        v[1] = ll[((size_t)nb + b) * M + c];
        v[2] = (double)last_acc[(size_t)b * M + c];
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) sm[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            double t = (threadIdx.x < (blockDim.x >> 5)) ? sm[q][threadIdx.x] : 0.0;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) partial[((size_t)q * nb + b) * gridDim.x + blockIdx.x] = t;
        }
    }
}
// second stage (one CTA): per-block sums of the CTA partials in index order, then totals over blocks in index order.
// out[0]=sum ll, out[1]=sum ll°, out[2..2+nb)=accept counts, out[2+nb..2+2nb)=per-block ll, out[2+2nb..2+3nb)=per-block ll°
__global__ void finish_stats_kernel(const double *partial, int ncta, int nb, double *out) {
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        double s[3] = {0.0, 0.0, 0.0};
        for (int q = 0; q < 3; q++)
            for (int i = 0; i < ncta; i++) s[q] += partial[((size_t)q * nb + b) * ncta + i];
        out[2 + b] = s[2];
        out[2 + nb + b] = s[0];
        out[2 + 2 * nb + b] = s[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, o = 0.0;
        for (int i = 0; i < nb; i++) { a += out[2 + nb + i]; o += out[2 + 2 * nb + i]; }
        out[0] = a; out[1] = o;
    }
}
__global__ void accept_counts_kernel(const uint8_t *acc_hist, int nb, int M, uint32_t it0, uint32_t it1, unsigned long long *counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    unsigned long long n = 0;
    if (c < M)
        for (uint32_t it = it0; it <= it1; ++it) n += acc_hist[((size_t)it * nb + b) * M + c];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&counts[b], n);
}


// ------------------------------------------------------------------------------------------- C1 over NVLink peer memory
// One-shot all-reduce of the few statistics doubles ([sum ll, sum ll°, accept counts per block]) written as ONE
// single-CTA kernel over peer-mapped memory instead of a NCCL call: every rank stores its values into slot `rank` of every
// peer's exchange buffer, publishes a sequence number with a system-scope release, waits for all peers' sequence numbers with
// acquire loads, and sums the slots IN RANK ORDER (bitwise reproducible for a given world size, identical on every rank).
// Two data parities alternate: a rank can be at most one reduction ahead of any other.
constexpr int P2P_MAX_RANKS = 16, P2P_MAX_VALS = 256;
struct P2PBuf {
    double data[2][P2P_MAX_RANKS][P2P_MAX_VALS];
    unsigned long long flag[P2P_MAX_RANKS];
    int error;
};
struct P2PArgs {
    P2PBuf *peer[P2P_MAX_RANKS]; // peer[r]: rank r's buffer as mapped into this process (peer[rank] = the local one)
    int rank, world, nval;
    unsigned long long seq;      // 1, 2, 3, ... the same on every rank
};
__global__ void __launch_bounds__(P2P_MAX_VALS) p2p_allreduce_kernel(const P2PArgs a, double *vals) {
    const int t = threadIdx.x, par = (int)(a.seq & 1ull);
    if (t < a.nval) {
        const double v = vals[t];
        for (int r = 0; r < a.world; r++) a.peer[r]->data[par][a.rank][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < a.world) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&a.peer[t]->flag[a.rank]), "l"(a.seq) : "memory");
        const unsigned long long *mine = &a.peer[a.rank]->flag[t];
        unsigned long long seen = 0, t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            if (seen >= a.seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 20000000000ull) { a.peer[a.rank]->error = 1; break; } // 20 s: a peer never arrived
        } while (true);
    }
    __syncthreads();
    if (t < a.nval) {
        double s = 0.0;
        for (int r = 0; r < a.world; r++) s += *(volatile const double *)&a.peer[a.rank]->data[par][r][t];
        vals[t] = s;
    }
}

} // namespace dmt
