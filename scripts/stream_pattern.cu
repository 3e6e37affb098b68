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
    for (int nb : {10, 20, 50, 100, 500}) {
        const int len = NT / nb; char nm[96];
        snprintf(nm, 96, "seq LDG.256 PF0   blocks=%3d (threads %d)", nb, M * nb);
        rep(nm, timeit([&] { k_seq<0><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len); }));
        snprintf(nm, 96, "seq LDG.256 PF2   blocks=%3d", nb);
        rep(nm, timeit([&] { k_seq<2><<<dim3((M + 63) / 64, nb), 64>>>(G, Wi, Wo, Xo, M, len); }));
        snprintf(nm, 96, "seq cp.async S=2  blocks=%3d", nb);
        CK(cudaFuncSetAttribute(k_seq_cp<2, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 12 * 2 * 64 * 16));
        rep(nm, timeit([&] { k_seq_cp<2, 64><<<dim3((M + 63) / 64, nb), 64, 2 * 12 * 2 * 64 * 16>>>(G, Wi, Wo, Xo, M, len); }));
        snprintf(nm, 96, "seq cp.async S=4  blocks=%3d", nb);
        CK(cudaFuncSetAttribute(k_seq_cp<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 12 * 2 * 64 * 16));
        rep(nm, timeit([&] { k_seq_cp<4, 64><<<dim3((M + 63) / 64, nb), 64, 4 * 12 * 2 * 64 * 16>>>(G, Wi, Wo, Xo, M, len); }));
    }
    return 0;
}
