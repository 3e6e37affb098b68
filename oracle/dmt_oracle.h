/*
 * dmt_oracle.h — CPU ORACLE for the guided-proposal path update.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libdmt.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference (DiffusionMCMCTools.jl) ships an empty test suite
 * (/root/reference/test/runtests.jl:4-6), no golden vectors, and its arithmetic lives in
 * un-vendored Julia packages (GuidedProposals 0.1.0 tree 65cd150e…, DiffusionDefinition 0.1.0
 * tree 0ed61dbd…, ObservationSchemes 0.1.0 tree 06a23cff…; /root/reference/Manifest.toml:131-135,
 * 200-206,324-328) that cannot run here (no Julia).  This file restates
 *   (i)  the container / ordering semantics that ARE in /root/reference/src (cited per function),
 *   (ii) the published guided-proposal equations for the arithmetic.
 * It is pinned only by analytic known-answer tests (tests/test_oracle_*.py).
 *
 * Storage is deliberately the reference's: one law + one X array + one W array PER INTERVAL,
 * accept = swap of per-interval pointers (src/biblock.jl:158-173) — nothing like the device's SoA.
 */
#ifndef DMT_ORACLE_H
#define DMT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXD 6
#define ORC_MAXPAR 12

enum { ORC_FHN = 0, ORC_LV = 1, ORC_LORENZ = 2, ORC_PROK = 3, ORC_JR = 4, ORC_OU2 = 5 };

/* ---- models ---- */
int  orc_model_dims(int model, int *d, int *dw, int *npar, int *constdiff);
void orc_drift(int model, const double *th, const double *x, double *b);
void orc_sigma(int model, const double *th, const double *x, double *S /* d x dw row-major */);
void orc_jacobian(int model, const double *th, const double *x, double *J /* d x d row-major */);
int  orc_bound_ok(int model, const double *th, const double *x);
/* Jacobian-at-reference-point auxiliary law: B=J(xbar), beta=b(xbar)-J xbar, atilde=a(xbar) */
void orc_linearise(int model, const double *th, const double *xbar, double *B, double *beta, double *at);

/* ---- Philox4x32-10 (Salmon et al. 2011, Random123) and the normal/exponential transforms ---- */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* 4*dw standard normals of one 4-step tile: normal index n = slot*dw + j, call n/2, cos for even n */
void   orc_tile_normals(uint64_t seed, uint32_t chain, uint32_t gtile, uint32_t iter, uint32_t layout, int dw, double *z);
double orc_accept_exponential(uint64_t seed, uint32_t chain, uint32_t block, uint32_t iter, uint32_t layout);

/* ---- one recording = SamplingPair (src/sampling_pair.jl:36-55) ---- */
typedef struct orc_pair orc_pair;

orc_pair *orc_pair_create(int model, int K, const int *n /*[K] points per interval*/,
                          const double *t /* concatenated, sum n */, int m, double artificial_noise);
void orc_pair_destroy(orc_pair *p);

/* side: 0 = u (accepted), 1 = u° (proposal).  store: 0 = PP, 1 = PPb (blocking laws). */
void orc_set_theta(orc_pair *p, int side, int store, int k, const double *theta);
void orc_set_aux(orc_pair *p, int side, int store, int k, const double *B, const double *beta, const double *at);
void orc_set_obs(orc_pair *p, int side, int k, const double *L, const double *Sig, const double *v);
void orc_set_start(orc_pair *p, const double *x0); /* XX[1].x[1] of both u and u° */
void orc_set_W(orc_pair *p, int side, int k, const double *dW /* (n_k-1) x dw increments */);
void orc_set_X(orc_pair *p, int side, int k, const double *X /* n_k x d */);
void orc_get_W(const orc_pair *p, int side, int k, double *dW);
void orc_get_X(const orc_pair *p, int side, int k, double *X);
void orc_get_HFc(const orc_pair *p, int side, int store, int k, double *H /*n x d x d*/, double *F /*n x d*/, double *c /*n*/);
void orc_set_HFc(orc_pair *p, int side, int store, int k, const double *H, const double *F, const double *c);

/* ---- a BiBlock = index range over a pair (src/block.jl:60-78, src/biblock.jl:43-63) ---- */
typedef struct {
    int i0, i1;   /* 0-based inclusive interval range */
    int last;     /* L flag: terminal block */
    double rho;
    double ll[2]; /* b.ll, b°.ll */
} orc_biblock;

/* GP.set_obs!(bb)                      src/biblock.jl:275-280 */
void orc_set_artificial_obs(orc_pair *p, const orc_biblock *bb);
/* recompute_guiding_term!(b::Block)    src/block.jl:104-110; side selects bb.b or bb.b° */
void orc_recompute_guiding_term(orc_pair *p, const orc_biblock *bb, int side);
/* the same with upstream's ODE solver: adaptive Tsit5 (OrdinaryDiffEq defaults reltol 1e-3, abstol 1e-6), dense output on the path
 * grid; returns the number of accepted steps.  See the header comment in dmt_oracle.c for what can be claimed of it. */
int orc_recompute_guiding_term_tsit5(orc_pair *p, const orc_biblock *bb, int side, double reltol, double abstol);
void orc_tsit5_tableau(double *c7, double *a7x6, double *btilde7, double *r7x4);
/* find_W_for_X!(bb)                    src/biblock.jl:300 -> src/block.jl:120-131 */
void orc_find_W_for_X(orc_pair *p, const orc_biblock *bb);
/* loglikhd!(bb) / loglikhd°!(bb)       src/biblock.jl:240,248 -> src/block.jl:140-152 */
double orc_loglikhd(orc_pair *p, orc_biblock *bb, int side, int skip);
/* draw_proposal_path!(bb)              src/biblock.jl:80-106.  Z: standard normals, block-local layout
 * [step][dw] over the block's intervals in order (NULL => Philox with (seed,chain,iter,layout)). Returns success. */
int orc_draw_proposal_path(orc_pair *p, orc_biblock *bb, const double *Z, uint64_t seed, uint32_t chain,
                           uint32_t iter, uint32_t layout, const int *gtile0 /*[K] global tile offset per interval*/);
/* recompute_path!(b°, b.WW; skip)      src/block.jl:161-187 with law side `law_side`, noise side `w_side`,
 * output into X of `law_side`; sets bb->ll[law_side]. Returns success. */
int orc_recompute_path(orc_pair *p, orc_biblock *bb, int law_side, int w_side, int skip);
/* accept_reject_proposal_path!(bb,i)   src/biblock.jl:121-127.  Returns accepted; ll_hist_out[2] = saved (b.ll, b°.ll) */
int orc_accept_reject(orc_pair *p, orc_biblock *bb, double E, double *ll_hist_out);
/* swaps                                src/biblock.jl:148-209 */
void orc_swap_XX(orc_pair *p, const orc_biblock *bb);
void orc_swap_WW(orc_pair *p, const orc_biblock *bb);
void orc_swap_PP(orc_pair *p, const orc_biblock *bb);
void orc_swap_ll(orc_biblock *bb);

/* ---- bulk CPU baseline: draw+accept over many independent recordings, OpenMP over recordings ---- */
double orc_sweep_many(orc_pair **pairs, orc_biblock *blocks /*[M][nb]*/, int M, int nb, uint64_t seed,
                      uint32_t chain0, uint32_t iter, uint32_t layout, const int *gtile0, int blocking, int nthreads,
                      int *n_accept);

#ifdef __cplusplus
}
#endif
#endif
