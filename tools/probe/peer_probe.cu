// peer_probe.cu -- how fast can ONE GPU push bytes into a peer's HBM over NVLink, per SM and
// in total, by access path?  (single process, two devices, cudaDeviceEnablePeerAccess)
//
//   lsu   : every thread loads 16 B from local HBM and stores 16 B to the peer (st.global),
//           a warp = 512 contiguous bytes; ROW > 0 scatters 16*ROW-byte runs at a 16 KiB pitch
//           (the store pattern of a strided FFT stage with P = ROW pencils per tile)
//   bulk  : local HBM -> shared memory (cp.async.bulk + mbarrier) -> peer (cp.async.bulk
//           shared -> global, bulk groups), CHUNK bytes per request, two buffers per CTA
//   ce    : cudaMemcpyPeerAsync (copy engines) for reference
// for grids of 8 .. 148 CTAs (one per SM).  Sizes the producer side of the fused
// FFT + redistribution kernels.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_probe.bin peer_probe.cu && ./peer_probe.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROW>
__global__ void __launch_bounds__(256) lsu_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
    const size_t step = (size_t)gridDim.x * blockDim.x * 4;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x) * 4 + threadIdx.x; i < n16; i += step) {
        uint4 a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = (i + k * 256 < n16) ? src[i + k * 256] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            size_t j = i + k * 256;
            if (j >= n16) continue;
            if constexpr (ROW > 0) {
                // runs of ROW units at a pitch of 1024 units: unit j -> row j / ROW, column j % ROW;
                // rows are visited pitch-major inside blocks of 1024 rows so the footprint stays n16 units
                const size_t blk = j / (1024 * (size_t)ROW), r = j % (1024 * (size_t)ROW);
                const size_t row = r / ROW, col = r % ROW;
                const size_t tile = blk % (1024 / ROW), big = blk / (1024 / ROW);
                j = big * 1024 * 1024 + row * 1024 + tile * ROW + col;
                if (j >= n16) continue;
            }
            dst[j] = a[k];
        }
    }
}

template <int CHUNK>
__global__ void __launch_bounds__(128) bulk_kernel(const char* __restrict__ src, char* __restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[2];
    const size_t nchunks = bytes / CHUNK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t par[2] = {0, 0};
        int s = 0;
        size_t c = blockIdx.x;
        // prologue: first load
        if (c < nchunks) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[0])), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(sm)), "l"(src + c * CHUNK), "r"(CHUNK), "r"(s32(&full[0])) : "memory");
        }
        for (; c < nchunks; c += gridDim.x) {
            const size_t cn = c + gridDim.x;
            // the other buffer must have been READ by its store before it is refilled
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (cn < nchunks) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s ^ 1])), "r"(CHUNK) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sm + (s ^ 1) * CHUNK)), "l"(src + cn * CHUNK), "r"(CHUNK), "r"(s32(&full[s ^ 1])) : "memory");
            }
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                         ::"r"(s32(&full[s])), "r"(par[s]) : "memory");
            par[s] ^= 1;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CHUNK), "r"(s32(sm + s * CHUNK)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            s ^= 1;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F>
static float timeit(F&& launch, cudaStream_t st, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t bytes = (size_t)1 << 30;
    const int reps = 5;
    char *src, *dst_local, *dst_peer = nullptr;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&src, bytes));
    CK(cudaMalloc(&dst_local, bytes));
    CK(cudaMemset(src, 1, bytes));
    if (ndev > 1) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, 0, 1));
        if (!can) { printf("no peer access 0 -> 1\n"); ndev = 1; }
    }
    if (ndev > 1) {
        CK(cudaSetDevice(1));
        CK(cudaMalloc(&dst_peer, bytes));
        CK(cudaMemset(dst_peer, 0, bytes));
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(0));
        CK(cudaDeviceEnablePeerAccess(1, 0));
    }
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    if (argc > 1 && ndev > 1) {
        // "./peer_probe.bin ncu": one launch of each store pattern at 148 CTAs into the peer, for
        // ncu --metrics nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,gpu__time_duration.sum
        CK(cudaFuncSetAttribute(bulk_kernel<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768));
        lsu_kernel<0><<<148, 256, 0, st>>>((const uint4*)src, (uint4*)dst_peer, bytes / 16);
        lsu_kernel<8><<<148, 256, 0, st>>>((const uint4*)src, (uint4*)dst_peer, bytes / 16);
        lsu_kernel<16><<<148, 256, 0, st>>>((const uint4*)src, (uint4*)dst_peer, bytes / 16);
        bulk_kernel<32768><<<148, 128, 2 * 32768, st>>>(src, dst_peer, bytes);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        printf("ncu mode: 4 launches done\n");
        return 0;
    }
    const int grids[] = {8, 16, 32, 64, 96, 148, 296};
    for (int target = 0; target < (ndev > 1 ? 2 : 1); ++target) {
        char* dst = target ? dst_peer : dst_local;
        const char* name = target ? "peer (NVLink)" : "local HBM";
        printf("== destination: %s, %zu MiB per launch\n", name, bytes >> 20);
        if (target) {
            float ms = timeit([&] { CK(cudaMemcpyPeerAsync(dst, 1, src, 0, bytes, st)); }, st, reps);
            printf("copy engine                      %8.3f ms %8.1f GB/s\n", ms, bytes / ms / 1e6);
        }
        for (int g : grids) {
            float ms = timeit([&] { lsu_kernel<0><<<g, 256, 0, st>>>((const uint4*)src, (uint4*)dst, bytes / 16); }, st, reps);
            printf("lsu contiguous     grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / ms / 1e6, bytes / ms / 1e6 / g);
        }
        for (int g : grids) {
            float ms = timeit([&] { lsu_kernel<8><<<g, 256, 0, st>>>((const uint4*)src, (uint4*)dst, bytes / 16); }, st, reps);
            printf("lsu 128-B rows     grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / ms / 1e6, bytes / ms / 1e6 / g);
        }
        for (int g : grids) {
            float ms = timeit([&] { lsu_kernel<16><<<g, 256, 0, st>>>((const uint4*)src, (uint4*)dst, bytes / 16); }, st, reps);
            printf("lsu 256-B rows     grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / ms / 1e6, bytes / ms / 1e6 / g);
        }
        CK(cudaFuncSetAttribute(bulk_kernel<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768));
        CK(cudaFuncSetAttribute(bulk_kernel<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096));
        CK(cudaFuncSetAttribute(bulk_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 512));
        CK(cudaFuncSetAttribute(bulk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128));
        for (int g : grids) {
            float ms = timeit([&] { bulk_kernel<32768><<<g, 128, 2 * 32768, st>>>(src, dst, bytes); }, st, reps);
            printf("bulk 32 KiB chunks grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / ms / 1e6, bytes / ms / 1e6 / g);
        }
        for (int g : grids) {
            float ms = timeit([&] { bulk_kernel<4096><<<g, 128, 2 * 4096, st>>>(src, dst, bytes); }, st, reps);
            printf("bulk 4 KiB chunks  grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / ms / 1e6, bytes / ms / 1e6 / g);
        }
        for (int g : grids) {
            if (g < 64) continue;
            float ms = timeit([&] { bulk_kernel<512><<<g, 128, 2 * 512, st>>>(src, dst, bytes / 8); }, st, reps);
            printf("bulk 512 B chunks  grid %4d      %8.3f ms %8.1f GB/s  (%.2f GB/s per CTA)\n", g, ms, bytes / 8 / ms / 1e6, bytes / 8 / ms / 1e6 / g);
        }
    }
    return 0;
}
