// fastmath.cuh — FP64 log / sqrt / sincospi specialised for the Box–Muller transform of the pCN refresh (K3).
//
// Why not libdevice: ncu's source page of the first draw kernel (profiles/r01a) showed that only ~31 % of issued
// instructions were FP64 math; libdevice's log/sincospi spend two UMOVs per polynomial coefficient (64-bit immediates
// cannot be DFMA operands) plus special-case handling for arguments that cannot occur here.  These versions
//   * take their coefficients from the constant bank (DFMA Rd, Ra, c[bank][off], Rc — no UMOV),
//   * assume the argument ranges Box–Muller produces (u1 in [2^-53, 1], angle in [0, 2)), so no denormal / NaN / Inf paths,
//   * are accurate to ~1 ulp (checked against the oracle's libm transforms to < 1e-14 on the GPU, tests/test_gpu_rng.py).
#pragma once
#include <cuda_runtime.h>

namespace dmt {

__constant__ double FM_LOG[9] = {0x1.5555555555555p-1, 0x1.999999999999ap-2, 0x1.2492492492492p-2, 0x1.c71c71c71c71cp-3, 0x1.745d1745d1746p-3,
                                 0x1.3b13b13b13b14p-3, 0x1.1111111111111p-3, 0x1.e1e1e1e1e1e1ep-4, 0x1.af286bca1af28p-4}; // 2/(2k+3)
__constant__ double FM_SIN[8] = {0x1.921fb54442d18p+1, -0x1.4abbce625be53p+2, 0x1.466bc6775aae2p+1, -0x1.32d2cce62bd86p-1,
                                 0x1.50783487ee782p-4, -0x1.e3074fde8871fp-8, 0x1.e8f434d018d63p-12, -0x1.6fadb9f155744p-16};
__constant__ double FM_COS[8] = {-0x1.3bd3cc9be45dep+2, 0x1.03c1f081b5ac4p+2, -0x1.55d3c7e3cbffap+0, 0x1.e1f506891babbp-3,
                                 -0x1.a6d1f2a204a8cp-6, 0x1.f9d38a3763cc3p-10, -0x1.b6e24f44b128fp-14, 0x1.20c62c2f2d7f5p-18};
__constant__ double FM_LN2[2] = {0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33}; // ln 2 = hi + lo, hi has 21 trailing zero bits

// natural log of a NORMAL positive double (no zero / denormal / inf / nan handling)
__device__ __forceinline__ double log_pos_normal(double u) {
    int hi = __double2hiint(u), lo = __double2loint(u);
    int e = (hi >> 20) - 1023;
    int him = (hi & 0x000fffff) | 0x3ff00000;          // mantissa in [1, 2)
    if (him >= 0x3ff6a09f) { him -= 0x00100000; e += 1; } // -> [sqrt(1/2), sqrt(2))
    const double m = __hiloint2double(him, lo);
    const double f = m - 1.0;                            // exact
    const double d = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); // ~20 bits
    double t = fma(-d, r, 1.0);
    r = fma(r, t, r);
    t = fma(-d, r, 1.0);
    r = fma(r, t, r);                                    // 1/d to ~1 ulp
    double s = f * r;
    s = fma(fma(-s, d, f), r, s);                        // s = f / (2 + f), residual-corrected
    const double s2 = s * s;
    double p = FM_LOG[8];
#pragma unroll
    for (int k = 7; k >= 0; k--) p = fma(p, s2, FM_LOG[k]);
    const double lm = fma(s * s2, p, 2.0 * s);           // log m = 2 atanh(s)
    const double ed = (double)e;
    return fma(ed, FM_LN2[0], fma(ed, FM_LN2[1], lm));
}

// sqrt of a non-negative finite double that is either 0 or >= ~1e-300 (here: -2 log u in [0, 73.5])
__device__ __forceinline__ double sqrt_nonneg(double x) {
    const double xs = fmax(x, 1e-290);                   // x == 0 (u1 == 1, probability 2^-53) -> ~1e-145 ~ 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    double g = xs * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    const double dres = fma(-g, g, xs);
    return fma(dres, h, g);
}

// sin(pi t), cos(pi t) for t in [0, 2]
__device__ __forceinline__ void sincospi_02(double t, double &s, double &c) {
    const double kd = rint(t + t);                       // quarter-period index 0..4
    const int k = (int)kd;
    const double f = fma(-0.5, kd, t);                   // [-1/4, 1/4], exact
    const double f2 = f * f;
    double ps = FM_SIN[7], pc = FM_COS[7];
#pragma unroll
    for (int i = 6; i >= 0; i--) { ps = fma(ps, f2, FM_SIN[i]); pc = fma(pc, f2, FM_COS[i]); }
    const double sf = ps * f, cf = fma(pc, f2, 1.0);
    const double a = (k & 1) ? cf : sf, b = (k & 1) ? sf : cf;
    s = (k & 2) ? -a : a;
    c = ((k + 1) & 2) ? -b : b;
}

} // namespace dmt
