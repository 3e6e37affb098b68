// models.cuh — drift / diffusion of the target laws as compile-time device code (sm_100a).
// Replaces, on this path, the DiffusionDefinition.jl `@load_diffusion` examples the reference's tutorials use
// (/root/reference/docs/src/tutorials/preamble.md:27-35: b, σ, nonhypo, nonhypo_σ).  Definitions: SURVEY.md Appendix B.
// One struct per model; everything is unrolled over the tiny static dimensions so state lives in registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace dmt {

enum { M_FHN = 0, M_LV = 1, M_LORENZ = 2, M_PROK = 3, M_JR = 4, M_OU2 = 5 };

// packed upper-triangular index of a symmetric DxD matrix
template <int D> __host__ __device__ constexpr int sidx(int i, int j) {
    return (i <= j) ? (i * D - i * (i - 1) / 2 + (j - i)) : (j * D - j * (j - 1) / 2 + (i - j));
}
template <int D> struct Dim { static constexpr int NH = D * (D + 1) / 2; };

template <int MODEL> struct Model;

// ------------------------------------------------------------------------------------------- FitzHugh–Nagumo
template <> struct Model<M_FHN> {
    static constexpr int D = 2, DW = 1, NPAR = 5;
    static constexpr bool CONSTDIFF = true;
    static constexpr bool ATIL_DIAG = true; // a = sigma sigma' of this law is diagonal => so is the auxiliary atilde (K1 exploits it)
    struct Par {
        double ieps, s, gam, bet, sig, sig2, isig;
        __device__ explicit Par(const double *th) : ieps(1.0 / th[0]), s(th[1]), gam(th[2]), bet(th[3]), sig(th[4]), sig2(th[4] * th[4]), isig(1.0 / th[4]) {}
    };
    struct Diff {
        const Par &p;
        __device__ Diff(const Par &p_, const double *) : p(p_) {}
        __device__ bool ok() const { return true; }
        __device__ void a_mul(const double *r, double *ar) const { ar[0] = 0.0; ar[1] = p.sig2 * r[1]; }
        __device__ void sig_mul(const double *dw, double *o) const { o[0] = 0.0; o[1] = p.sig * dw[0]; }
        __device__ void inv_sig(const double *res, double *dw) const { dw[0] = res[1] * p.isig; }
        __device__ void a_sym(double *a) const { a[0] = 0; a[1] = 0; a[2] = p.sig2; }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        b[0] = (x[0] - x[1] - x[0] * x[0] * x[0] + p.s) * p.ieps;
        b[1] = p.gam * x[0] - x[1] + p.bet;
    }
    static __device__ void jac(const Par &p, const double *x, double *J) {
        J[0] = (1.0 - 3.0 * x[0] * x[0]) * p.ieps; J[1] = -p.ieps;
        J[2] = p.gam;                              J[3] = -1.0;
    }
    static __device__ bool bound_ok(const Par &, const double *) { return true; }
};

// ------------------------------------------------------------------------------------------- Lotka–Volterra
template <> struct Model<M_LV> {
    static constexpr int D = 2, DW = 2, NPAR = 6;
    static constexpr bool CONSTDIFF = true;
    static constexpr bool ATIL_DIAG = true; // a = sigma sigma' of this law is diagonal => so is the auxiliary atilde (K1 exploits it)
    struct Par {
        double al, be, ga, de, s1, s2;
        __device__ explicit Par(const double *th) : al(th[0]), be(th[1]), ga(th[2]), de(th[3]), s1(th[4]), s2(th[5]) {}
    };
    struct Diff {
        const Par &p;
        __device__ Diff(const Par &p_, const double *) : p(p_) {}
        __device__ bool ok() const { return true; }
        __device__ void a_mul(const double *r, double *ar) const { ar[0] = p.s1 * p.s1 * r[0]; ar[1] = p.s2 * p.s2 * r[1]; }
        __device__ void sig_mul(const double *dw, double *o) const { o[0] = p.s1 * dw[0]; o[1] = p.s2 * dw[1]; }
        __device__ void inv_sig(const double *res, double *dw) const { dw[0] = res[0] / p.s1; dw[1] = res[1] / p.s2; }
        __device__ void a_sym(double *a) const { a[0] = p.s1 * p.s1; a[1] = 0; a[2] = p.s2 * p.s2; }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        b[0] = p.al * x[0] - p.be * x[0] * x[1];
        b[1] = p.de * x[0] * x[1] - p.ga * x[1];
    }
    static __device__ void jac(const Par &p, const double *x, double *J) {
        J[0] = p.al - p.be * x[1]; J[1] = -p.be * x[0];
        J[2] = p.de * x[1];        J[3] = p.de * x[0] - p.ga;
    }
    static __device__ bool bound_ok(const Par &, const double *x) { return x[0] > 0.0 && x[1] > 0.0; }
};

// ------------------------------------------------------------------------------------------- Lorenz
template <> struct Model<M_LORENZ> {
    static constexpr int D = 3, DW = 3, NPAR = 4;
    static constexpr bool CONSTDIFF = true;
    static constexpr bool ATIL_DIAG = true; // a = sigma sigma' of this law is diagonal => so is the auxiliary atilde (K1 exploits it)
    struct Par {
        double t1, t2, t3, s, s2, is;
        __device__ explicit Par(const double *th) : t1(th[0]), t2(th[1]), t3(th[2]), s(th[3]), s2(th[3] * th[3]), is(1.0 / th[3]) {}
    };
    struct Diff {
        const Par &p;
        __device__ Diff(const Par &p_, const double *) : p(p_) {}
        __device__ bool ok() const { return true; }
        __device__ void a_mul(const double *r, double *ar) const { ar[0] = p.s2 * r[0]; ar[1] = p.s2 * r[1]; ar[2] = p.s2 * r[2]; }
        __device__ void sig_mul(const double *dw, double *o) const { o[0] = p.s * dw[0]; o[1] = p.s * dw[1]; o[2] = p.s * dw[2]; }
        __device__ void inv_sig(const double *res, double *dw) const { dw[0] = res[0] * p.is; dw[1] = res[1] * p.is; dw[2] = res[2] * p.is; }
        __device__ void a_sym(double *a) const { a[0] = p.s2; a[1] = 0; a[2] = 0; a[3] = p.s2; a[4] = 0; a[5] = p.s2; }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        b[0] = p.t1 * (x[1] - x[0]);
        b[1] = p.t2 * x[0] - x[1] - x[0] * x[2];
        b[2] = x[0] * x[1] - p.t3 * x[2];
    }
    static __device__ void jac(const Par &p, const double *x, double *J) {
        J[0] = -p.t1;        J[1] = p.t1; J[2] = 0.0;
        J[3] = p.t2 - x[2];  J[4] = -1.0; J[5] = -x[0];
        J[6] = x[1];         J[7] = x[0]; J[8] = -p.t3;
    }
    static __device__ bool bound_ok(const Par &, const double *) { return true; }
};

// ------------------------------------------------------------------------------------------- Prokaryotic autoregulation
// state (RNA, P, P2, DNA); 8 reactions, hazards h, b = S h, a = S diag(h) S^T, sigma(x) = chol(a(x)) (lower)
template <> struct Model<M_PROK> {
    static constexpr int D = 4, DW = 4, NPAR = 9;
    static constexpr bool CONSTDIFF = false;
    static constexpr bool ATIL_DIAG = false;
    struct Par {
        double c[8], K;
        __device__ explicit Par(const double *th) {
#pragma unroll
            for (int i = 0; i < 8; i++) c[i] = th[i];
            K = th[8];
        }
    };
    static __device__ void hazards(const Par &p, const double *x, double *h) {
        h[0] = p.c[0] * x[3] * x[2];
        h[1] = p.c[1] * (p.K - x[3]);
        h[2] = p.c[2] * x[3];
        h[3] = p.c[3] * x[0];
        h[4] = p.c[4] * x[1] * (x[1] - 1.0) * 0.5;
        h[5] = p.c[5] * x[2];
        h[6] = p.c[6] * x[0];
        h[7] = p.c[7] * x[1];
    }
    struct Diff {
        double a[10]; // packed upper: (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)
        double l[10]; // lower Cholesky factor, l[sidx(i,j)] = L_{max,min}
        bool good;
        __device__ Diff(const Par &p, const double *x) {
            double h[8];
            hazards(p, x, h);
            // S rows: RNA [0,0,1,0,0,0,-1,0]; P [0,0,0,1,-2,2,0,-1]; P2 [-1,1,0,0,1,-1,0,0]; DNA [-1,1,0,0,0,0,0,0]
            double h01 = h[0] + h[1];
            a[0] = h[2] + h[6];                       // RNA,RNA
            a[1] = 0.0;                               // RNA,P
            a[2] = 0.0;                               // RNA,P2
            a[3] = 0.0;                               // RNA,DNA
            a[4] = h[3] + 4.0 * h[4] + 4.0 * h[5] + h[7]; // P,P
            a[5] = -2.0 * h[4] - 2.0 * h[5];          // P,P2
            a[6] = 0.0;                               // P,DNA
            a[7] = h01 + h[4] + h[5];                 // P2,P2
            a[8] = h01;                               // P2,DNA
            a[9] = h01;                               // DNA,DNA
            good = chol();
        }
        __device__ bool chol() {
            constexpr int D = 4;
            bool g = true;
#pragma unroll
            for (int j = 0; j < D; j++) {
                double s = a[sidx<D>(j, j)];
#pragma unroll
                for (int k = 0; k < j; k++) s -= l[sidx<D>(j, k)] * l[sidx<D>(j, k)];
                g = g && (s > 0.0);
                double lj = sqrt(s);
                l[sidx<D>(j, j)] = lj;
                double ilj = 1.0 / lj;
#pragma unroll
                for (int i = j + 1; i < D; i++) {
                    double t = a[sidx<D>(i, j)];
#pragma unroll
                    for (int k = 0; k < j; k++) t -= l[sidx<D>(i, k)] * l[sidx<D>(j, k)];
                    l[sidx<D>(i, j)] = t * ilj;
                }
            }
            return g;
        }
        __device__ bool ok() const { return good; }
        __device__ void a_mul(const double *r, double *ar) const {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double s = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) s += a[sidx<4>(i, j)] * r[j];
                ar[i] = s;
            }
        }
        __device__ void sig_mul(const double *dw, double *o) const { // lower-triangular L dw
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double s = 0;
#pragma unroll
                for (int j = 0; j <= i; j++) s += l[sidx<4>(i, j)] * dw[j];
                o[i] = s;
            }
        }
        __device__ void inv_sig(const double *res, double *dw) const { // forward substitution L dw = res
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double s = res[i];
#pragma unroll
                for (int j = 0; j < i; j++) s -= l[sidx<4>(i, j)] * dw[j];
                dw[i] = s / l[sidx<4>(i, i)];
            }
        }
        __device__ void a_sym(double *o) const {
#pragma unroll
            for (int i = 0; i < 10; i++) o[i] = a[i];
        }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        double h[8];
        hazards(p, x, h);
        b[0] = h[2] - h[6];
        b[1] = h[3] - 2.0 * h[4] + 2.0 * h[5] - h[7];
        b[2] = -h[0] + h[1] + h[4] - h[5];
        b[3] = -h[0] + h[1];
    }
    static __device__ void jac(const Par &p, const double *x, double *J) {
        // dh/dx (8x4), J = S dh/dx
        double d0_3 = p.c[0] * x[2], d0_2 = p.c[0] * x[3], d1_3 = -p.c[1], d2_3 = p.c[2], d3_0 = p.c[3];
        double d4_1 = p.c[4] * (2.0 * x[1] - 1.0) * 0.5, d5_2 = p.c[5], d6_0 = p.c[6], d7_1 = p.c[7];
        J[0] = -d6_0;  J[1] = 0.0;                   J[2] = 0.0;            J[3] = d2_3;
        J[4] = d3_0;   J[5] = -2.0 * d4_1 - d7_1;    J[6] = 2.0 * d5_2;     J[7] = 0.0;
        J[8] = 0.0;    J[9] = d4_1;                  J[10] = -d0_2 - d5_2;  J[11] = -d0_3 + d1_3;
        J[12] = 0.0;   J[13] = 0.0;                  J[14] = -d0_2;         J[15] = -d0_3 + d1_3;
    }
    static __device__ bool bound_ok(const Par &p, const double *x) {
        return x[0] > 0.0 && x[1] > 1.0 && x[2] > 0.0 && x[3] > 0.0 && x[3] < p.K;
    }
};

// ------------------------------------------------------------------------------------------- Jansen–Rit
template <> struct Model<M_JR> {
    static constexpr int D = 6, DW = 1, NPAR = 10;
    static constexpr bool CONSTDIFF = true;
    static constexpr bool ATIL_DIAG = true; // a = sigma sigma' of this law is diagonal => so is the auxiliary atilde (K1 exploits it)
    struct Par {
        double A, a, B, b, C1, C2, C3, C4, numax, v0, r, mu, sigy, sigy2, isigy;
        __device__ explicit Par(const double *th)
            : A(th[0]), a(th[1]), B(th[2]), b(th[3]), C1(th[4]), C2(0.8 * th[4]), C3(0.25 * th[4]), C4(0.25 * th[4]),
              numax(th[5]), v0(th[6]), r(th[7]), mu(th[8]), sigy(th[9]), sigy2(th[9] * th[9]), isigy(1.0 / th[9]) {}
        __device__ double sigm(double v) const { return numax / (1.0 + exp(r * (v0 - v))); }
        __device__ double dsigm(double v) const { double s = sigm(v); return r * s * (1.0 - s / numax); }
    };
    struct Diff {
        const Par &p;
        __device__ Diff(const Par &p_, const double *) : p(p_) {}
        __device__ bool ok() const { return true; }
        __device__ void a_mul(const double *r, double *ar) const {
            ar[0] = ar[1] = ar[2] = ar[3] = ar[5] = 0.0;
            ar[4] = p.sigy2 * r[4];
        }
        __device__ void sig_mul(const double *dw, double *o) const {
            o[0] = o[1] = o[2] = o[3] = o[5] = 0.0;
            o[4] = p.sigy * dw[0];
        }
        __device__ void inv_sig(const double *res, double *dw) const { dw[0] = res[4] * p.isigy; }
        __device__ void a_sym(double *a) const {
#pragma unroll
            for (int i = 0; i < 21; i++) a[i] = 0.0;
            a[sidx<6>(4, 4)] = p.sigy2;
        }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        b[0] = x[3];
        b[1] = x[4];
        b[2] = x[5];
        b[3] = p.A * p.a * p.sigm(x[1] - x[2]) - 2.0 * p.a * x[3] - p.a * p.a * x[0];
        b[4] = p.A * p.a * (p.mu + p.C2 * p.sigm(p.C1 * x[0])) - 2.0 * p.a * x[4] - p.a * p.a * x[1];
        b[5] = p.B * p.b * p.C4 * p.sigm(p.C3 * x[0]) - 2.0 * p.b * x[5] - p.b * p.b * x[2];
    }
    static __device__ void jac(const Par &p, const double *x, double *J) {
#pragma unroll
        for (int i = 0; i < 36; i++) J[i] = 0.0;
        J[0 * 6 + 3] = 1.0; J[1 * 6 + 4] = 1.0; J[2 * 6 + 5] = 1.0;
        double s12 = p.dsigm(x[1] - x[2]);
        J[3 * 6 + 0] = -p.a * p.a; J[3 * 6 + 1] = p.A * p.a * s12; J[3 * 6 + 2] = -p.A * p.a * s12; J[3 * 6 + 3] = -2.0 * p.a;
        J[4 * 6 + 0] = p.A * p.a * p.C2 * p.C1 * p.dsigm(p.C1 * x[0]); J[4 * 6 + 1] = -p.a * p.a; J[4 * 6 + 4] = -2.0 * p.a;
        J[5 * 6 + 0] = p.B * p.b * p.C4 * p.C3 * p.dsigm(p.C3 * x[0]); J[5 * 6 + 2] = -p.b * p.b; J[5 * 6 + 5] = -2.0 * p.b;
    }
    static __device__ bool bound_ok(const Par &, const double *) { return true; }
};

// ------------------------------------------------------------------------------------------- 2-D linear (OU) test model
template <> struct Model<M_OU2> {
    static constexpr int D = 2, DW = 2, NPAR = 8;
    static constexpr bool CONSTDIFF = true;
    static constexpr bool ATIL_DIAG = true; // a = sigma sigma' of this law is diagonal => so is the auxiliary atilde (K1 exploits it)
    struct Par {
        double B[4], be[2], s1, s2;
        __device__ explicit Par(const double *th) : s1(th[6]), s2(th[7]) {
            B[0] = th[0]; B[1] = th[1]; B[2] = th[2]; B[3] = th[3]; be[0] = th[4]; be[1] = th[5];
        }
    };
    struct Diff {
        const Par &p;
        __device__ Diff(const Par &p_, const double *) : p(p_) {}
        __device__ bool ok() const { return true; }
        __device__ void a_mul(const double *r, double *ar) const { ar[0] = p.s1 * p.s1 * r[0]; ar[1] = p.s2 * p.s2 * r[1]; }
        __device__ void sig_mul(const double *dw, double *o) const { o[0] = p.s1 * dw[0]; o[1] = p.s2 * dw[1]; }
        __device__ void inv_sig(const double *res, double *dw) const { dw[0] = res[0] / p.s1; dw[1] = res[1] / p.s2; }
        __device__ void a_sym(double *a) const { a[0] = p.s1 * p.s1; a[1] = 0; a[2] = p.s2 * p.s2; }
    };
    static __device__ void drift(const Par &p, const double *x, double *b) {
        b[0] = p.B[0] * x[0] + p.B[1] * x[1] + p.be[0];
        b[1] = p.B[2] * x[0] + p.B[3] * x[1] + p.be[1];
    }
    static __device__ void jac(const Par &p, const double *, double *J) { J[0] = p.B[0]; J[1] = p.B[1]; J[2] = p.B[2]; J[3] = p.B[3]; }
    static __device__ bool bound_ok(const Par &, const double *) { return true; }
};

// ------------------------------------------------------------------------------------------- structural zeros of the Jacobian
// When the auxiliary law is the device-side Jacobian linearisation of the target (dmt_set_aux_linearised), B has the Jacobian's sparsity
// pattern, and the backward filter's d^3 product H (B - a H / 2) only needs the entries that can be non-zero.  nz(q, j): entry (q, j) of
// C = B - atilde H / 2 may be non-zero (rows of C that carry noise are full: atilde H fills them).  Default: dense.
template <class MD> struct JacMask {
    static constexpr bool SPARSE = false;
    static __host__ __device__ constexpr bool nz(int, int) { return true; }
};
template <> struct JacMask<Model<M_JR>> { // Jansen-Rit: x' = (x3, x4, x5, ...), three second-order blocks coupled through sigmoids; noise on row 4
    static constexpr bool SPARSE = true;
    static __host__ __device__ constexpr bool nz(int q, int j) {
        return q == 4 || (q < 3 && j == q + 3) || (q == 3 && j <= 3) || (q == 5 && (j == 0 || j == 2 || j == 5));
    }
};

} // namespace dmt
