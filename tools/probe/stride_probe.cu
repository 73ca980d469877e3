// stride_probe.cu -- what bandwidth can a tile-shaped copy reach on B200?
//
// The strided-axis FFT kernels move tiles of N rows x (P*16) contiguous bytes
// whose rows are `inner*16` bytes apart.  This probe measures the ceiling of
// that access pattern with the arithmetic removed, for
//   reg : the register path of fft_pow2_kernel (each thread LDG.128 x E, STG.128 x E)
//   tma : persistent CTAs, cp.async.bulk.tensor box loads into a ring of shared
//         memory stages and bulk tensor stores out of them
// over (outer, N, inner) views of one big buffer.  Build: see tools/probe/build.sh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct __align__(16) c128 { double x, y; };

// ---------------------------------------------------------------------------
template <int P, int E>
__global__ void reg_copy(const c128* __restrict__ in, c128* __restrict__ out, long long N, long long inner,
                         long long tiles_per_outer) {
    const int tid = threadIdx.x;
    const int p = tid % P, q = tid / P;
    const int TP = blockDim.x / P;          // == N / E
    const long long bid = blockIdx.x;
    const long long o = bid / tiles_per_outer;
    const long long i = (bid - o * tiles_per_outer) * P + p;
    const c128* gin = in + o * N * inner + i;
    c128* gout = out + o * N * inner + i;
    c128 v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = gin[(long long)(q + e * TP) * inner];
#pragma unroll
    for (int e = 0; e < E; ++e) gout[(long long)(q + e * TP) * inner] = v[e];
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}

// tile = N rows x P elements; box rows = BR (<= 256), N/BR boxes per tile
template <int STAGES>
__global__ void tma_copy(const __grid_constant__ CUtensorMap min, const __grid_constant__ CUtensorMap mout,
                         int N, int P, int BR, long long tiles_per_outer, long long ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES];
    const uint32_t tile_bytes = (uint32_t)N * P * 16;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long first = blockIdx.x, step = gridDim.x;
    const long long mine = first < ntiles ? (ntiles - first + step - 1) / step : 0;
    auto issue = [&](long long k) {
        const long long t = first + k * step;
        const int s = (int)(k % STAGES);
        const long long o = t / tiles_per_outer;
        const long long i = (t - o * tiles_per_outer) * P;
        mbar_expect_tx(&full[s], tile_bytes);
        for (int r = 0; r < N; r += BR)
            tma_load_3d(smem + (size_t)s * tile_bytes + (size_t)r * P * 16, &min, (int)(2 * i), r, (int)o, &full[s]);
    };
    for (long long k = 0; k < STAGES - 1 && k < mine; ++k) issue(k);
    for (long long k = 0; k < mine; ++k) {
        const int s = (int)(k % STAGES);
        mbar_wait(&full[s], (uint32_t)((k / STAGES) & 1));
        const long long t = first + k * step;
        const long long o = t / tiles_per_outer;
        const long long i = (t - o * tiles_per_outer) * P;
        for (int r = 0; r < N; r += BR)
            tma_store_3d(&mout, (int)(2 * i), r, (int)o, smem + (size_t)s * tile_bytes + (size_t)r * P * 16);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the stage that tile k+STAGES-1 will land in was stored by tile k-1
        if (k + STAGES - 1 < mine) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue(k + STAGES - 1);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
    return (EncodeFn)fn;
}
static CUtensorMap make_map(void* base, long long outer, long long N, long long inner, int P, int BR, int l2promo) {
    static EncodeFn enc = get_encode();
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)(2 * inner), (cuuint64_t)N, (cuuint64_t)outer};
    cuuint64_t strides[2] = {(cuuint64_t)(inner * 16), (cuuint64_t)(N * inner * 16)};
    cuuint32_t box[3] = {(cuuint32_t)(2 * P), (cuuint32_t)BR, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

static const char* verify(const c128* b, long long total) {
    static const long long pos[5] = {0, 1, 2, 3, 4};
    for (int k = 0; k < 5; ++k) {
        long long i = k == 0 ? 0 : (k == 4 ? total - 1 : (total / 4) * pos[k] + 12345 * k);
        c128 h;
        CK(cudaMemcpy(&h, b + i, 16, cudaMemcpyDeviceToHost));
        if (((unsigned char*)&h)[5] != 1) return "WRONG";
    }
    return "ok";
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

template <class F>
static double bench(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return time_ms(a, b) / reps;
}

template <int P, int E>
static void run_reg(const char* tag, const c128* in, c128* out, long long outer, long long N, long long inner, int smem_pad) {
    const long long tpo = inner / P;
    const long long grid = outer * tpo;
    const int threads = (int)(N / E) * P;
    if (threads > 1024 || threads < 32) return;
    CK(cudaFuncSetAttribute(reg_copy<P, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaMemset(out, 0, outer * N * inner * 16));
    double ms = bench([&] { reg_copy<P, E><<<(unsigned)grid, threads, smem_pad>>>(in, out, N, inner, tpo); });
    const double bytes = 2.0 * outer * N * inner * 16;
    printf("%-8s reg P=%2d E=%2d thr=%4d smem=%3dK  %8.3f ms %7.1f GB/s %s\n", tag, P, E, threads, smem_pad / 1024, ms, bytes / ms / 1e6, verify(out, outer * N * inner));
    fflush(stdout);
}

template <int STAGES>
static void run_tma(const char* tag, c128* in, c128* out, long long outer, long long N, long long inner, int P, int ctas_per_sm, int l2promo) {
    const int BR = N > 256 ? 256 : (int)N;
    CUtensorMap mi = make_map(in, outer, N, inner, P, BR, l2promo), mo = make_map(out, outer, N, inner, P, BR, l2promo);
    const long long tpo = inner / P, ntiles = outer * tpo;
    const size_t smem = (size_t)STAGES * N * P * 16;
    if (smem * ctas_per_sm > 225 * 1024) return;
    CK(cudaFuncSetAttribute(tma_copy<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = 148LL * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    CK(cudaMemset(out, 0, outer * N * inner * 16));
    double ms = bench([&] { tma_copy<STAGES><<<(unsigned)grid, 32, smem>>>(mi, mo, (int)N, P, BR, tpo, ntiles); });
    const double bytes = 2.0 * outer * N * inner * 16;
    printf("%-8s tma P=%2d stages=%d cta/sm=%d l2p=%d tile=%3zuK  %8.3f ms %7.1f GB/s %s\n", tag, P, STAGES, ctas_per_sm, l2promo,
           (size_t)N * P * 16 / 1024, ms, bytes / ms / 1e6, verify(out, outer * N * inner));
    fflush(stdout);
}

int main(int argc, char** argv) {
    const long long S = argc > 1 ? atoll(argv[1]) : 512;
    const long long pad = argc > 2 ? atoll(argv[2]) : 0;      // extra elements per axis-0 row (breaks the power-of-two stride)
    const int quick = argc > 3 ? atoi(argv[3]) : 0;
    const long long total = S * (S * S + pad);
    c128 *a, *b;
    CK(cudaMalloc(&a, total * 16));
    CK(cudaMalloc(&b, total * 16));
    CK(cudaMemset(a, 1, total * 16));
    CK(cudaMemset(b, 0, total * 16));
    double ms = bench([&] { CK(cudaMemcpyAsync(b, a, total * 16, cudaMemcpyDeviceToDevice)); });
    printf("S=%lld memcpy D2D %8.3f ms %7.1f GB/s\n", S, ms, 2.0 * total * 16 / ms / 1e6);
    struct View { const char* tag; long long outer, N, inner; } views[2] = {{"axis1", S, S, S}, {"axis0", 1, S, S * S + pad}};
    for (auto& v : views) {
        if (quick) {
            if (v.outer != 1) continue;
            run_reg<4, 16>(v.tag, a, b, v.outer, v.N, v.inner, 0);
            run_reg<8, 16>(v.tag, a, b, v.outer, v.N, v.inner, 0);
            run_reg<16, 16>(v.tag, a, b, v.outer, v.N, v.inner, 0);
            run_tma<2>(v.tag, a, b, v.outer, v.N, v.inner, 4, 1, 0);
            run_tma<2>(v.tag, a, b, v.outer, v.N, v.inner, 8, 1, 0);
            continue;
        }
        // register path: P x E grid, unconstrained occupancy and 2/1 CTAs per SM via shared memory ballast
        for (int pad : {0, 100 * 1024}) {
            run_reg<2, 8>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<4, 8>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<8, 8>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<16, 8>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<4, 16>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<8, 16>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<16, 16>(v.tag, a, b, v.outer, v.N, v.inner, pad);
            run_reg<32, 16>(v.tag, a, b, v.outer, v.N, v.inner, pad);
        }
        for (int l2p : {0, 2}) {
            for (int P : {2, 4, 8, 16}) {
                run_tma<2>(v.tag, a, b, v.outer, v.N, v.inner, P, 1, l2p);
                run_tma<3>(v.tag, a, b, v.outer, v.N, v.inner, P, 1, l2p);
                run_tma<4>(v.tag, a, b, v.outer, v.N, v.inner, P, 1, l2p);
                run_tma<2>(v.tag, a, b, v.outer, v.N, v.inner, P, 2, l2p);
                run_tma<3>(v.tag, a, b, v.outer, v.N, v.inner, P, 2, l2p);
                run_tma<2>(v.tag, a, b, v.outer, v.N, v.inner, P, 4, l2p);
            }
        }
    }
    return 0;
}
