/* abi_smoke.c — the C ABI exercised WITHOUT ctypes or Python: a plain C program compiled with gcc against include/dmt.h that builds a
 * small Lorenz ensemble, runs init_paths! + one blocking sweep + accept_reject_proposal_path! on the device and checks the results for
 * basic sanity.  Proves that the header compiles as C (not C++) and that the entry points can be bound by any FFI.
 *   gcc -std=c99 -Iinclude tests/abi_smoke.c -o abi_smoke -ldl && ./abi_smoke diffusionmcmctools.jl_b200/libdmt.so
 * (run from tests/test_gpu_abi_c.py under `pytest -m gpu`; the library is dlopen'ed so that the program links without CUDA). */
#include "dmt.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LOAD(name) do { *(void **)(&p_##name) = dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; } } while (0)
#define CHECK(call) do { int32_t rc_ = (call); if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)rc_, p_dmt_last_error(ctx)); return 3; } } while (0)

static int32_t (*p_dmt_create)(const dmt_config *, const int32_t *, const double *, const int32_t *, dmt_ctx **);
static int32_t (*p_dmt_destroy)(dmt_ctx *);
static const char *(*p_dmt_last_error)(const dmt_ctx *);
static int32_t (*p_dmt_model_dims)(int32_t, int32_t *, int32_t *, int32_t *, int32_t *);
static int32_t (*p_dmt_set_params)(dmt_ctx *, int32_t, int32_t, int32_t, int32_t, const double *);
static int32_t (*p_dmt_set_obs)(dmt_ctx *, int32_t, int32_t, int32_t, const double *, const double *, const double *);
static int32_t (*p_dmt_set_aux_linearised)(dmt_ctx *, int32_t, int32_t, int32_t, int32_t, const double *);
static int32_t (*p_dmt_set_start)(dmt_ctx *, const double *);
static int32_t (*p_dmt_set_blocks)(dmt_ctx *, int32_t, int32_t, const int32_t *, const int32_t *, const double *, const uint8_t *, int32_t);
static int32_t (*p_dmt_recompute_guiding_term)(dmt_ctx *, int32_t, int32_t);
static int32_t (*p_dmt_init_paths)(dmt_ctx *, int32_t, uint32_t, int32_t, int32_t *);
static int32_t (*p_dmt_blocking_sweep)(dmt_ctx *, int32_t, uint32_t);
static int32_t (*p_dmt_accept_reject_path)(dmt_ctx *, int32_t, uint32_t, const double *);
static int32_t (*p_dmt_get_ll)(dmt_ctx *, int32_t, int32_t, double *);
static int32_t (*p_dmt_get_last_accept)(dmt_ctx *, int32_t, uint8_t *);
static int32_t (*p_dmt_get_X)(dmt_ctx *, int32_t, double *);

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "libdmt.so";
    void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen(%s): %s\n", path, dlerror()); return 2; }
    LOAD(dmt_create); LOAD(dmt_destroy); LOAD(dmt_last_error); LOAD(dmt_model_dims); LOAD(dmt_set_params); LOAD(dmt_set_obs);
    LOAD(dmt_set_aux_linearised); LOAD(dmt_set_start); LOAD(dmt_set_blocks); LOAD(dmt_recompute_guiding_term); LOAD(dmt_init_paths);
    LOAD(dmt_blocking_sweep); LOAD(dmt_accept_reject_path); LOAD(dmt_get_ll); LOAD(dmt_get_last_accept); LOAD(dmt_get_X);

    enum { M = 40, K = 6, NST = 20, NPT = NST + 1, MOBS = 2, NB = 2 };
    int32_t d = 0, dw = 0, npar = 0, cd = 0;
    dmt_ctx *ctx = NULL;
    if (p_dmt_model_dims(DMT_LORENZ, &d, &dw, &npar, &cd) || d != 3 || dw != 3 || npar != 4) { fprintf(stderr, "model dims\n"); return 3; }

    int32_t n_pts[K];
    double tt[K * NPT];
    for (int k = 0; k < K; k++) {
        n_pts[k] = NPT;
        for (int i = 0; i < NPT; i++) tt[k * NPT + i] = 0.2 * k + 0.2 * i / NST;
    }
    dmt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.model = DMT_LORENZ; cfg.n_chains = M; cfg.n_psets = M; cfg.n_intervals = K; cfg.obs_dim = MOBS; cfg.device = 0;
    cfg.two_sided_laws = 0; cfg.ll_hist_len = 4; cfg.n_layouts = 2; cfg.chain_offset = 0; cfg.seed = 7; cfg.artificial_noise = 1e-11;
    CHECK(p_dmt_create(&cfg, n_pts, tt, NULL, &ctx));

    /* theta [npar][P]; L [K][m*d][P], Sigma [K][m*m][P], v [K][m][P]; xbar [K][d][P]; x0 [d][M]: parameter set / chain index fastest */
    static double theta[4 * M], L[K * MOBS * 3 * M], Sig[K * MOBS * MOBS * M], v[K * MOBS * M], xbar[K * 3 * M], x0[3 * M];
    const double th[4] = {10.0, 28.0, 8.0 / 3.0, 2.0}, start[3] = {1.5, -1.5, 25.0};
    for (int p = 0; p < M; p++) {
        for (int i = 0; i < 4; i++) theta[i * M + p] = th[i];
        for (int i = 0; i < 3; i++) x0[i * M + p] = start[i] * (1.0 + 0.01 * p / M);
        for (int k = 0; k < K; k++) {
            const double Lm[MOBS][3] = {{1, 0, 0}, {0, 0, 1}};
            for (int a = 0; a < MOBS; a++) {
                for (int j = 0; j < 3; j++) L[((k * MOBS + a) * 3 + j) * M + p] = Lm[a][j];
                for (int b = 0; b < MOBS; b++) Sig[((k * MOBS + a) * MOBS + b) * M + p] = a == b ? 1.0 : 0.0;
            }
            v[(k * MOBS + 0) * M + p] = start[0] + 0.3 * sin(1.0 + k + 0.1 * p);
            v[(k * MOBS + 1) * M + p] = start[2] + 0.3 * cos(2.0 + k + 0.1 * p);
            for (int i = 0; i < 3; i++) xbar[(k * 3 + i) * M + p] = start[i];
        }
    }
    CHECK(p_dmt_set_params(ctx, DMT_ACCEPTED, 3, 0, K - 1, theta));
    CHECK(p_dmt_set_obs(ctx, DMT_ACCEPTED, 0, K - 1, L, Sig, v));
    CHECK(p_dmt_set_aux_linearised(ctx, DMT_ACCEPTED, DMT_STORE_PP, 0, K - 1, xbar));
    CHECK(p_dmt_set_aux_linearised(ctx, DMT_ACCEPTED, DMT_STORE_PPB, 0, K - 1, xbar));
    CHECK(p_dmt_set_start(ctx, x0));

    const int32_t whole_i0[1] = {0}, whole_i1[1] = {K - 1}, i0[NB] = {0, 3}, i1[NB] = {2, K - 1};
    const double rho1[1] = {0.0}, rho[NB] = {0.7, 0.7};
    CHECK(p_dmt_set_blocks(ctx, 1, 1, whole_i0, whole_i1, rho1, NULL, 0));
    CHECK(p_dmt_set_blocks(ctx, 0, NB, i0, i1, rho, NULL, 4));
    CHECK(p_dmt_recompute_guiding_term(ctx, 1, DMT_P_ONLY));
    int32_t n_failed = -1;
    CHECK(p_dmt_init_paths(ctx, 1, 1000u, 50, &n_failed));
    if (n_failed != 0) { fprintf(stderr, "init_paths left %d failing chains\n", (int)n_failed); return 4; }

    CHECK(p_dmt_blocking_sweep(ctx, 0, 0u));
    CHECK(p_dmt_accept_reject_path(ctx, 0, 0u, NULL));
    static double ll[NB * M], llo[NB * M], X[K * NPT * 3 * M];
    static uint8_t acc[NB * M];
    CHECK(p_dmt_get_ll(ctx, 0, DMT_ACCEPTED, ll));
    CHECK(p_dmt_get_ll(ctx, 0, DMT_PROPOSAL, llo));
    CHECK(p_dmt_get_last_accept(ctx, 0, acc));
    CHECK(p_dmt_get_X(ctx, DMT_ACCEPTED, X));
    int n_acc = 0, bad = 0;
    for (int i = 0; i < NB * M; i++) {
        n_acc += acc[i] != 0;
        if (!isfinite(ll[i])) bad++;
    }
    for (int i = 0; i < K * NPT * 3 * M; i++) if (!isfinite(X[i])) bad++;
    /* every interval starts where the previous one ended (XX[k].x[1] == XX[k-1].x[end]): exactly inside a block; where two blocks meet
     * (k == 3) the left block's proposal was steered towards the frozen end point and ends within one Euler-Maruyama step's noise
     * (sigma sqrt(dt) ~ 0.2 on this program's crude uniform grid; the reference's tau-transformed grids make that step tiny) of it */
    for (int k = 1; k < K; k++)
        for (int j = 0; j < 3 * M; j++) {
            const double a = X[(size_t)(k * NPT) * 3 * M + j], b = X[(size_t)(k * NPT - 1) * 3 * M + j];
            if (k == i0[1] ? fabs(a - b) > 1.0 : a != b) bad++;
        }
    printf("abi_smoke: %d of %d (block, chain) proposals accepted, ll[0] = %.6f, bad = %d\n", n_acc, NB * M, ll[0], bad);
    CHECK(p_dmt_destroy(ctx));
    return (bad == 0 && n_acc > 0 && n_acc < NB * M) ? 0 : 5;
}
