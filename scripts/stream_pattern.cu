// stream_pattern.cu — micro-benchmark of the draw kernel's MEMORY pattern only (no arithmetic): what bandwidth can the
// [tile][component][chain][4] sector-per-lane layout reach on this GPU, as a function of how the loads are issued and of
// how many threads are in flight?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int NG = 9, DW = 3, D = 3;

__device__ __forceinline__ void ld256(const double *p, double *v) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double *p, const double *v) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "l"(p) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t s, const void *g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// fully parallel: one thread per (chain, tile)
__global__ void k_par(const double *G, const double *Win, double *Wout, double *Xout, int M, int NT) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
    if (c >= M) return;
    double acc[4] = {0, 0, 0, 0}, v[4];
    for (int a = 0; a < NG; a++) { ld256(G + (((size_t)t * NG + a) * M + c) * 4, v); for (int i = 0; i < 4; i++) acc[i] += v[i]; }
    for (int j = 0; j < DW; j++) {
        ld256(Win + (((size_t)t * DW + j) * M + c) * 4, v);
        for (int i = 0; i < 4; i++) v[i] += acc[i];
        st256(Wout + (((size_t)t * DW + j) * M + c) * 4, v);
        st256(Xout + (((size_t)t * D + j) * M + c) * 4, v);
    }
}
// sequential over `len` tiles per thread (one thread per (chain, block)), LDG.256 into registers, optional L2 prefetch distance
template <int PF>
__global__ void k_seq(const double *G, const double *Win, double *Wout, double *Xout, int M, int len) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= M) return;
    double carry = 0;
    for (int q = 0; q < len; q++) {
        const size_t t = (size_t)b * len + q;
        double g[NG][4], w[DW][4];
        for (int a = 0; a < NG; a++) ld256(G + ((t * NG + a) * M + c) * 4, g[a]);
        for (int j = 0; j < DW; j++) ld256(Win + ((t * DW + j) * M + c) * 4, w[j]);
        if (PF > 0 && q + PF < len) {
            for (int a = 0; a < NG; a++) asm volatile("prefetch.global.L2 [%0];" ::"l"(G + (((t + PF) * NG + a) * M + c) * 4));
            for (int j = 0; j < DW; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(Win + (((t + PF) * DW + j) * M + c) * 4));
        }
        double acc[4] = {carry, carry, carry, carry};
        for (int a = 0; a < NG; a++) for (int i = 0; i < 4; i++) acc[i] += g[a][i];
        for (int j = 0; j < DW; j++) {
            for (int i = 0; i < 4; i++) w[j][i] += acc[i];
            st256(Wout + ((t * DW + j) * M + c) * 4, w[j]);
            st256(Xout + ((t * D + j) * M + c) * 4, w[j]);
        }
        carry = acc[3] * 1e-30;
    }
}
// ---- the two candidates WITH a dependent arithmetic chain of `work` DFMAs per step (stands for the Euler-Maruyama recursion):
// (a) the shipped pattern: whole 4-step tile loaded, 4 x work, tile stored;
// (b) layout [step][component][chain] (8 bytes per lane, LDG.64/STG.64) with a rolling register pipeline PF steps deep, so
//     loads stay in flight during the recursion.  Same bytes, same work, same thread count.
__device__ __forceinline__ double ld64(const double *p) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ void st64(double *p, double v) { asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__global__ void __launch_bounds__(64, 6) k_tile_work(const double *G, const double *Win, double *Wout, double *Xout, int M, int len, int work) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= M) return;
    double carry = 0;
    for (int q = 0; q < len; q++) {
        const size_t t = (size_t)b * len + q;
        double g[NG][4], w[DW][4];
        for (int a = 0; a < NG; a++) ld256(G + ((t * NG + a) * M + c) * 4, g[a]);
        for (int j = 0; j < DW; j++) ld256(Win + ((t * DW + j) * M + c) * 4, w[j]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double acc = carry;
            for (int a = 0; a < NG; a++) acc += g[a][i];
            for (int k = 0; k < work; k++) acc = fma(acc, 0.999999, 1e-9);
            for (int j = 0; j < DW; j++) w[j][i] += acc;
            carry = acc * 1e-30;
        }
        for (int j = 0; j < DW; j++) {
            st256(Wout + ((t * DW + j) * M + c) * 4, w[j]);
            st256(Xout + ((t * D + j) * M + c) * 4, w[j]);
        }
    }
}
template <int PF>
__global__ void __launch_bounds__(64, 6) k_step_work(const double *G, const double *Win, double *Wout, double *Xout, int M, int len, int work) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (c >= M) return;
    const size_t s0 = (size_t)b * len;
    double g[PF][NG], w[PF][DW], carry = 0;
#pragma unroll
    for (int p = 0; p < PF; p++) {
        for (int a = 0; a < NG; a++) g[p][a] = ld64(G + ((s0 + p) * NG + a) * M + c);
        for (int j = 0; j < DW; j++) w[p][j] = ld64(Win + ((s0 + p) * DW + j) * M + c);
    }
    for (int s = 0; s < len; s += PF) {
#pragma unroll
        for (int p = 0; p < PF; p++) {
            const size_t t = s0 + s + p;
            double acc = carry;
            for (int a = 0; a < NG; a++) acc += g[p][a];
            double wv[DW];
            for (int j = 0; j < DW; j++) wv[j] = w[p][j];
            if (s + p + PF < len) { // slot p is free again: request step s + p + PF
                for (int a = 0; a < NG; a++) g[p][a] = ld64(G + ((t + PF) * NG + a) * M + c);
                for (int j = 0; j < DW; j++) w[p][j] = ld64(Win + ((t + PF) * DW + j) * M + c);
            }
            for (int k = 0; k < work; k++) acc = fma(acc, 0.999999, 1e-9);
            for (int j = 0; j < DW; j++) {
                st64(Wout + (t * DW + j) * M + c, wv[j] + acc);
                st64(Xout + (t * D + j) * M + c, wv[j] + acc);
            }
            carry = acc * 1e-30;
        }
    }
}
// sequential, everything read through an S-stage cp.async ring in shared memory
template <int S, int TPB>
__global__ void k_seq_cp(const double *G, const double *Win, double *Wout, double *Xout, int M, int len) {
    extern __shared__ double2 sm[];
    constexpr int NC = NG + DW;
    const int tid = threadIdx.x, c = blockIdx.x * TPB + tid, b = blockIdx.y;
    if (c >= M) return;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm) + tid * 16u;
    auto issue = [&](size_t t, int st) {
        for (int a = 0; a < NC; a++) {
            const double *src = (a < NG) ? G + ((t * NG + a) * M + c) * 4 : Win + ((t * DW + (a - NG)) * M + c) * 4;
            cp_async16(s0 + (uint32_t)(((st * NC + a) * 2 + 0) * TPB) * 16u, src);
            cp_async16(s0 + (uint32_t)(((st * NC + a) * 2 + 1) * TPB) * 16u, src + 2);
        }
    };
    for (int p = 0; p < S - 1; p++) { if (p < len) issue((size_t)b * len + p, p); cp_commit(); }
    double carry = 0;
    for (int q = 0; q < len; q++) {
        const size_t t = (size_t)b * len + q;
        if (q + S - 1 < len) issue(t + S - 1, (q + S - 1) % S);
        cp_commit();
        cp_wait<S - 1>();
        const double2 *sg = sm + (size_t)(q % S) * NC * 2 * TPB + tid;
        double acc[4] = {carry, carry, carry, carry};
        for (int a = 0; a < NG; a++) { double2 u = sg[(a * 2) * TPB], v = sg[(a * 2 + 1) * TPB]; acc[0] += u.x; acc[1] += u.y; acc[2] += v.x; acc[3] += v.y; }
        for (int j = 0; j < DW; j++) {
            double2 u = sg[((NG + j) * 2) * TPB], v = sg[((NG + j) * 2 + 1) * TPB];
            double w[4] = {u.x + acc[0], u.y + acc[1], v.x + acc[2], v.y + acc[3]};
            st256(Wout + ((t * DW + j) * M + c) * 4, w);
            st256(Xout + ((t * D + j) * M + c) * 4, w);
        }
        carry = acc[3] * 1e-30;
    }
    cp_wait<0>();
}

// ---- TMA variant: the warp's H,F (and W) chunks are contiguous (32 lanes x 32 B = 1 KiB per component): one elected lane
// copies them global -> shared with cp.async.bulk (full-line L2 access, no registers), S tiles ahead, per-warp mbarriers.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)),
                 "r"(parity) : "memory");
}
template <int S, int WPB, bool W_TMA>
__global__ void __launch_bounds__(WPB * 32) k_seq_tma(const double *G, const double *Win, double *Wout, double *Xout, int M, int len) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int NC = W_TMA ? NG + DW : NG;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *ring = reinterpret_cast<double *>(smraw) + (size_t)wid * S * NC * 128; // [stage][comp][lane][4]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + (size_t)WPB * S * NC * 1024) + wid * S;
    const int c0 = (blockIdx.x * WPB + wid) * 32, c = c0 + lane, b = blockIdx.y;
    if (c0 >= M) return;
    if (lane == 0) for (int s = 0; s < S; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](size_t t, int st) {
        if (lane == 0) {
            mbar_expect_tx(&bars[st], NC * 1024);
            for (int a = 0; a < NG; a++) bulk_g2s(ring + ((size_t)st * NC + a) * 128, G + ((t * NG + a) * M + c0) * 4, 1024, &bars[st]);
            if (W_TMA) for (int j = 0; j < DW; j++) bulk_g2s(ring + ((size_t)st * NC + NG + j) * 128, Win + ((t * DW + j) * M + c0) * 4, 1024, &bars[st]);
        }
    };
    for (int p = 0; p < S - 1 && p < len; p++) issue((size_t)b * len + p, p);
    double carry = 0;
    for (int q = 0; q < len; q++) {
        const size_t t = (size_t)b * len + q;
        if (q + S - 1 < len) issue(t + S - 1, (q + S - 1) % S); // that stage was drained in iteration q-1 (syncwarp below)
        double w[DW][4];
        if (!W_TMA) for (int j = 0; j < DW; j++) ld256(Win + ((t * DW + j) * M + c) * 4, w[j]);
        mbar_wait(&bars[q % S], (q / S) & 1);
        const double *sg = ring + (size_t)(q % S) * NC * 128 + lane * 4;
        double acc[4] = {carry, carry, carry, carry};
        for (int a = 0; a < NG; a++) {
            const double2 u = *reinterpret_cast<const double2 *>(sg + a * 128), v = *reinterpret_cast<const double2 *>(sg + a * 128 + 2);
            acc[0] += u.x; acc[1] += u.y; acc[2] += v.x; acc[3] += v.y;
        }
        if (W_TMA) for (int j = 0; j < DW; j++) {
            const double2 u = *reinterpret_cast<const double2 *>(sg + (NG + j) * 128), v = *reinterpret_cast<const double2 *>(sg + (NG + j) * 128 + 2);
            w[j][0] = u.x; w[j][1] = u.y; w[j][2] = v.x; w[j][3] = v.y;
        }
        __syncwarp(); // every lane has read its sectors of this stage: it may be refilled
        for (int j = 0; j < DW; j++) {
            for (int i = 0; i < 4; i++) w[j][i] += acc[i];
            st256(Wout + ((t * DW + j) * M + c) * 4, w[j]);
            st256(Xout + ((t * D + j) * M + c) * 4, w[j]);
        }
        carry = acc[3] * 1e-30;
    }
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char **argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 4096, NT = 5000;
    size_t nG = (size_t)NT * NG * M * 4, nW = (size_t)NT * DW * M * 4;
    double *G, *Wi, *Wo, *Xo;
    CK(cudaMalloc(&G, nG * 8)); CK(cudaMalloc(&Wi, nW * 8)); CK(cudaMalloc(&Wo, nW * 8)); CK(cudaMalloc(&Xo, nW * 8));
    CK(cudaMemset(G, 0, nG * 8)); CK(cudaMemset(Wi, 0, nW * 8));
    const double gb = (nG + 3 * nW) * 8 / 1e9;
    printf("M=%d tiles=%d bytes moved %.2f GB (read %.2f, write %.2f)\n", M, NT, gb, (nG + nW) * 8 / 1e9, 2 * nW * 8 / 1e9);
    auto rep = [&](const char *n, float ms) { printf("%-44s %8.3f ms  %7.1f GB/s\n", n, ms, gb / ms * 1e3); };
    rep("copy (cudaMemcpy D2D, same bytes r+w)", timeit([&] { cudaMemcpyAsync(G, G + nG / 2, (size_t)(gb * 1e9 / 2), cudaMemcpyDeviceToDevice); }));
    rep("par: thread per (chain,tile), TPB 128", timeit([&] { k_par<<<dim3((M + 127) / 128, NT), 128>>>(G, Wi, Wo, Xo, M, NT); }));
    if (argc > 2) { // second argument: only the tile-vs-rolling comparison
        for (int nb : {10, 20}) for (int work : {0, 64, 128, 256, 512}) {
            const int len = NT / nb; char nm[96];
            snprintf(nm, 96, "tile   LDG.256      blocks=%2d work=%3d/step", nb, work);
            rep(nm, timeit([&] { k_tile_work<<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len, work); }));
            snprintf(nm, 96, "rolling LDG.64 PF=4 blocks=%2d work=%3d/step", nb, work);
            rep(nm, timeit([&] { k_step_work<4><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len * 4, work); }));
            snprintf(nm, 96, "rolling LDG.64 PF=8 blocks=%2d work=%3d/step", nb, work);
            rep(nm, timeit([&] { k_step_work<8><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len * 4, work); }));
        }
        return 0;
    }
    for (int nb : {10, 20, 50, 100, 500}) {
        const int len = NT / nb; char nm[96];
        snprintf(nm, 96, "seq LDG.256 PF0   blocks=%3d (threads %d)", nb, M * nb);
        rep(nm, timeit([&] { k_seq<0><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len); }));
        snprintf(nm, 96, "seq LDG.256 PF2   blocks=%3d", nb);
        rep(nm, timeit([&] { k_seq<2><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len); }));
        snprintf(nm, 96, "seq cp.async S=2  blocks=%3d", nb);
        CK(cudaFuncSetAttribute(k_seq_cp<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 12 * 2 * 64 * 16));
        rep(nm, timeit([&] { k_seq_cp<2, 64><<<dim3((M + 63) / 64, nb), 64, 2 * 12 * 2 * 64 * 16>>>(G, Wi, Wo, Xo, M, len); }));
#define RUN_TMA(S_, W_)                                                                                                   \
    {                                                                                                                    \
        constexpr int NCc = (W_) ? NG + DW : NG;                                                                         \
        const int smem = 2 * (S_) * NCc * 1024 + 2 * (S_) * 8;                                                           \
        CK(cudaFuncSetAttribute(k_seq_tma<S_, 2, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));               \
        snprintf(nm, 96, "seq TMA ring S=%d %s blocks=%3d", S_, (W_) ? "G+W" : "G  ", nb);                                \
        rep(nm, timeit([&] { k_seq_tma<S_, 2, W_><<<dim3((M + 63) / 64, nb), 64, smem>>>(G, Wi, Wo, Xo, M, len); }));     \
    }
        RUN_TMA(2, false) RUN_TMA(3, false) RUN_TMA(4, false) RUN_TMA(3, true) RUN_TMA(4, true)
        snprintf(nm, 96, "seq cp.async S=4  blocks=%3d", nb);
        CK(cudaFuncSetAttribute(k_seq_cp<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 12 * 2 * 64 * 16));
        rep(nm, timeit([&] { k_seq_cp<4, 64><<<dim3((M + 63) / 64, nb), 64, 4 * 12 * 2 * 64 * 16>>>(G, Wi, Wo, Xo, M, len); }));
    }
    return 0;
}
