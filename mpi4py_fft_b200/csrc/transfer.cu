// transfer.cu -- the global transpose: pack kernel -> NCCL all-to-all(v) ->
// unpack kernel, all enqueued on the caller's stream.
//
// Replaces MPI_Alltoallw over subarray datatypes
// (/root/reference/mpi4py_fft/pencil.py:12-29 builds one datatype per peer that
// selects block i of the balanced distribution along an axis; pencil.py:182-183
// / 200-201 exchange them).  Equivalent formulation used here:
//   forward : peer i is sent   A[.., sA_i:sA_i+nA_i (axisA), ..]   (C order)
//             and its block lands in B[.., sB_i:sB_i+nB_i (axisB), ..]
//   backward: the same with A and B swapped.
// A block is contiguous in its array exactly when nothing but unit extents
// precede the split axis (outer == 1); then the pack (or unpack) pass is skipped
// and NCCL reads (writes) the array directly.
//
// NCCL is bound at run time with dlopen (the copy torch already loaded when
// there is one), so the library has no link-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>
#include "b200fft.h"
#include "internal.h"
#include "transfer_put.h"

namespace b2f {

// ---- NCCL binding ----------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
    std::string why;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy already in the process (torch's)
        if (h) break;
    }
    if (!h) {
        const char* env = getenv("B2F_NCCL_LIB");
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    for (const char* n : names) {
        if (h) break;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) {
        g_nccl.why = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return;
    }
    g_nccl.handle = h;
#define B2F_SYM(field, name)                                                  \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));  \
    if (!g_nccl.field) {                                                      \
        g_nccl.why = std::string("libnccl lacks symbol ") + name;             \
        return;                                                               \
    }
    B2F_SYM(GetUniqueId, "ncclGetUniqueId")
    B2F_SYM(CommInitRank, "ncclCommInitRank")
    B2F_SYM(CommDestroy, "ncclCommDestroy")
    B2F_SYM(Send, "ncclSend")
    B2F_SYM(Recv, "ncclRecv")
    B2F_SYM(GroupStart, "ncclGroupStart")
    B2F_SYM(GroupEnd, "ncclGroupEnd")
    B2F_SYM(AllReduce, "ncclAllReduce")
    B2F_SYM(GetErrorString, "ncclGetErrorString")
    B2F_SYM(GetVersion, "ncclGetVersion")
#undef B2F_SYM
    g_nccl.ok = true;
}

static int need_nccl() {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) {
        set_error(g_nccl.why);
        return B2F_ENCCL;
    }
    return B2F_OK;
}
static int nccl_fail(ncclResult_t r, const char* what) {
    set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
    return B2F_ENCCL;
}

// ---- balanced block distribution (== reference pencil.py:5-9) -----------------
struct BlockDist {
    long long n, p, q, r;   // n items over p owners: q = n / p, first r owners get q + 1
    __host__ __device__ long long start(long long i) const { return i * q + (i < r ? i : r); }
    __host__ __device__ long long len(long long i) const { return q + (i < r ? 1 : 0); }
    __host__ __device__ long long owner(long long a) const {
        const long long big = r * (q + 1);
        return a < big ? a / (q + 1) : r + (a - big) / q;
    }
};
static BlockDist make_dist(long long n, long long p) { return BlockDist{n, p, n / p, n % p}; }

// ---- pack / unpack kernel -----------------------------------------------------
// The array is (outer, N, U) in units of V bytes (U = inner * itemsize / V); the
// packed buffer holds, for owner i = 0..p-1 in turn, the block
// (outer, len(i), U) in C order, i.e. segment i starts at unit outer*U*start(i).
// One thread moves one unit; consecutive threads walk a row of the array, so
// both sides are accessed in contiguous runs of len(i)*U units.
template <class V, bool PACK>
__global__ void __launch_bounds__(256) copy_blocks_kernel(const V* __restrict__ src, V* __restrict__ dst,
                                                         long long outer, long long U, BlockDist bd) {
    const long long row = bd.n * U;
    const long long total = outer * row;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += step) {
        const long long o = idx / row;
        const long long u = idx - o * row;
        const long long a = u / U;
        const long long i = bd.owner(a);
        const long long s = bd.start(i), ln = bd.len(i);
        const long long packed = outer * U * s + o * ln * U + (u - s * U);
        if (PACK) dst[packed] = src[idx];
        else dst[idx] = src[packed];
    }
}

template <bool PACK>
static cudaError_t launch_copy_blocks(const void* src, void* dst, long long outer, long long n, long long inner,
                                      int itemsize, long long p, cudaStream_t st) {
    const long long row_bytes_per_a = inner * itemsize;
    int v = 16;
    while (v > 1 && (row_bytes_per_a % v || ((uintptr_t)src % v) || ((uintptr_t)dst % v))) v >>= 1;
    const long long U = row_bytes_per_a / v;
    const long long total = outer * n * U;
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;   // persistent-ish grid: 16 CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    const BlockDist bd = make_dist(n, p);
    switch (v) {
        case 16: copy_blocks_kernel<uint4, PACK><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)src, (uint4*)dst, outer, U, bd); break;
        case 8: copy_blocks_kernel<uint2, PACK><<<(unsigned)blocks, 256, 0, st>>>((const uint2*)src, (uint2*)dst, outer, U, bd); break;
        case 4: copy_blocks_kernel<uint32_t, PACK><<<(unsigned)blocks, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, outer, U, bd); break;
        case 2: copy_blocks_kernel<uint16_t, PACK><<<(unsigned)blocks, 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst, outer, U, bd); break;
        default: copy_blocks_kernel<uint8_t, PACK><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)src, (uint8_t*)dst, outer, U, bd); break;
    }
    count_launch();
    return cudaGetLastError();
}

// ---- peer-memory put: the whole redistribution as ONE kernel ------------------
// Every rank stores the block each peer needs straight into that peer's array
// (peer pointers are CUDA-IPC mappings of the peers' work buffers, reached over
// NVLink/NVSwitch), so the pack pass, the staging buffers and the unpack pass of
// the NCCL formulation disappear: the block is read once from local HBM and
// written once into its final place.
//
// Geometry (direction-neutral: S = source side, D = destination side).  The
// group-local shape is collapsed to five extents around the two split axes,
// ax1 < ax2:  (P, e1, M, e2, Q).  The block for peer i has extents
// (P, b1, M, b2, Q) on both sides and is a set of P*b1*M rows, each b2*Q
// elements long and contiguous in the source AND in the destination.
//   source      S[p, o1S + i1, m, o2S + i2, q]     extents (P, e1S, M, e2S, Q)
//   destination D[p, o1D + i1, m, o2D + i2, q]     extents (P, e1D, M, e2D, Q)
// All row quantities are kept in units of V bytes (V = widest vector that
// divides every row length and offset).
#if defined(__CUDACC__)
template <class V>
__global__ void __launch_bounds__(256) put_blocks_kernel(const PutParams prm) {
    const PutPeer& pr = prm.peer[blockIdx.y];
    const V* __restrict__ src = reinterpret_cast<const V*>(prm.src);
    V* __restrict__ dst = reinterpret_cast<V*>(pr.dst);
    const long long step = (long long)gridDim.x * blockDim.x;
    long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent units per thread in flight
    for (; u + 3 * step < pr.units; u += 4 * step) {
        long long s0, d0, s1, d1, s2, d2, s3, d3;
        put_locate(prm, pr, u, &s0, &d0);
        put_locate(prm, pr, u + step, &s1, &d1);
        put_locate(prm, pr, u + 2 * step, &s2, &d2);
        put_locate(prm, pr, u + 3 * step, &s3, &d3);
        const V a = src[s0], b = src[s1], c = src[s2], d = src[s3];
        dst[d0] = a;
        dst[d1] = b;
        dst[d2] = c;
        dst[d3] = d;
    }
    for (; u < pr.units; u += step) {
        long long s0, d0;
        put_locate(prm, pr, u, &s0, &d0);
        dst[d0] = src[s0];
    }
    __threadfence_system();   // peer stores are performed before the kernel retires
}
#endif

}  // namespace b2f

using namespace b2f;

struct b2f_comm_s {
    ncclComm_t comm;
    int nranks, rank;
    int* d_flag;   // two ints for the stream-ordered group barrier of the peer-memory path
};

struct b2f_transfer_s {
    b2f_comm comm;
    // flag barrier over peer memory (b2f_transfer_set_flags): flags[j] points at the
    // nranks 64-bit arrival counters of group rank j (my own array for j == rank), as
    // mapped in this process; epoch counts the barriers this transfer has enqueued
    unsigned long long* flags[B2F_MAX_PEERS] = {};
    unsigned long long epoch = 0;
    bool has_flags = false;
    // staging buffers of the pack -> NCCL -> unpack path: owned by the transfer (two transfers in
    // flight on different streams must not share them), sized on first use, freed with the handle
    void* ws[2] = {nullptr, nullptr};
    size_t ws_cap[2] = {0, 0};
    int nranks, rank, ndims, itemsize;
    std::vector<long long> shape, subA, subB;
    int axisA, axisB;
    long long outerA, innerA, NA, outerB, innerB, NB;
    long long volA, volB;   // elements
};

extern "C" {

int b2f_comm_unique_id(void* id128) {
    if (!id128) return B2F_EINVAL;
    int rc = need_nccl();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id128, &id, 128);
    return B2F_OK;
}

int b2f_comm_create(b2f_comm* comm, const void* id128, int nranks, int rank) {
    if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("b2f_comm_create: bad arguments");
        return B2F_EINVAL;
    }
    int rc = need_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c;
    ncclResult_t r = g_nccl.CommInitRank(&c, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
    int* flag = nullptr;
    cudaError_t e = cudaMalloc((void**)&flag, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(flag, 0, 2 * sizeof(int));
    if (e != cudaSuccess) {
        g_nccl.CommDestroy(c);
        return cuda_fail(e, "cudaMalloc(barrier flag)");
    }
    *comm = new b2f_comm_s{c, nranks, rank, flag};
    return B2F_OK;
}

int b2f_comm_destroy(b2f_comm comm) {
    if (!comm) return B2F_OK;
    if (g_nccl.ok && comm->comm) g_nccl.CommDestroy(comm->comm);
    if (comm->d_flag) cudaFree(comm->d_flag);
    delete comm;
    return B2F_OK;
}

int b2f_transfer_create(b2f_transfer* out, b2f_comm comm, int nranks, int rank, int ndims,
                        const int64_t* shape, int itemsize, const int64_t* subshapeA, int axisA,
                        const int64_t* subshapeB, int axisB) {
    if (!out || !shape || !subshapeA || !subshapeB || ndims < 2 || nranks < 1 || rank < 0 || rank >= nranks ||
        axisA < 0 || axisA >= ndims || axisB < 0 || axisB >= ndims || axisA == axisB || itemsize < 1) {
        set_error("b2f_transfer_create: bad arguments");
        return B2F_EINVAL;
    }
    // comm may be NULL: the handle then serves geometry / pack / unpack only
    if (comm && (comm->nranks != nranks || comm->rank != rank)) {
        set_error("b2f_transfer_create: communicator does not match nranks/rank");
        return B2F_EINVAL;
    }
    b2f_transfer_s* t = new b2f_transfer_s;
    t->comm = comm;
    t->nranks = nranks;
    t->rank = rank;
    t->ndims = ndims;
    t->itemsize = itemsize;
    t->shape.assign(shape, shape + ndims);
    t->subA.assign(subshapeA, subshapeA + ndims);
    t->subB.assign(subshapeB, subshapeB + ndims);
    t->axisA = axisA;
    t->axisB = axisB;
    // consistency with the balanced distribution: A is full along axisA and holds
    // my block of axisB, B the other way round, every other extent is shared
    const BlockDist dA = make_dist(shape[axisA], nranks), dB = make_dist(shape[axisB], nranks);
    bool ok = shape[axisA] >= nranks && shape[axisB] >= nranks;
    for (int i = 0; i < ndims && ok; ++i) {
        if (i == axisA) ok = subshapeA[i] == shape[i] && subshapeB[i] == dA.len(rank);
        else if (i == axisB) ok = subshapeB[i] == shape[i] && subshapeA[i] == dB.len(rank);
        else ok = subshapeA[i] == shape[i] && subshapeB[i] == shape[i];
    }
    if (!ok) {
        delete t;
        set_error("b2f_transfer_create: subshapes are not the balanced blocks of shape for this rank");
        return B2F_EINVAL;
    }
    auto prod = [](const std::vector<long long>& v, int a, int b) {
        long long p = 1;
        for (int i = a; i < b; ++i) p *= v[i];
        return p;
    };
    t->NA = shape[axisA];
    t->NB = shape[axisB];
    t->outerA = prod(t->subA, 0, axisA);
    t->innerA = prod(t->subA, axisA + 1, ndims);
    t->outerB = prod(t->subB, 0, axisB);
    t->innerB = prod(t->subB, axisB + 1, ndims);
    t->volA = prod(t->subA, 0, ndims);
    t->volB = prod(t->subB, 0, ndims);
    *out = t;
    return B2F_OK;
}

int b2f_transfer_destroy(b2f_transfer t) {
    if (t) {
        for (int k = 0; k < 2; ++k)
            if (t->ws[k]) cudaFree(t->ws[k]);
    }
    delete t;
    return B2F_OK;
}

static int workspace(b2f_transfer t, int slot, size_t bytes, void** out) {
    if (t->ws_cap[slot] < bytes) {
        if (t->ws[slot]) {
            cudaError_t e = cudaFree(t->ws[slot]);
            t->ws[slot] = nullptr;
            t->ws_cap[slot] = 0;
            if (e != cudaSuccess) return cuda_fail(e, "cudaFree(transfer workspace)");
        }
        cudaError_t e = cudaMalloc(&t->ws[slot], bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(transfer workspace)");
        t->ws_cap[slot] = bytes;
    }
    *out = t->ws[slot];
    return B2F_OK;
}

int b2f_transfer_geometry(b2f_transfer t, int64_t* send_counts, int64_t* send_offsets, int64_t* recv_counts,
                          int64_t* recv_offsets) {
    if (!t) return B2F_EINVAL;
    const BlockDist dA = make_dist(t->NA, t->nranks), dB = make_dist(t->NB, t->nranks);
    const long long rowA = t->outerA * t->innerA, rowB = t->outerB * t->innerB;
    for (int i = 0; i < t->nranks; ++i) {
        if (send_counts) send_counts[i] = dA.len(i) * rowA;
        if (send_offsets) send_offsets[i] = dA.start(i) * rowA;
        if (recv_counts) recv_counts[i] = dB.len(i) * rowB;
        if (recv_offsets) recv_offsets[i] = dB.start(i) * rowB;
    }
    return B2F_OK;
}

// direction 0: pack A by axisA blocks; 1: pack B by axisB blocks
int b2f_transfer_pack(b2f_transfer t, int direction, const void* d_src, void* d_packed, void* stream) {
    if (!t || !d_src || !d_packed) return B2F_EINVAL;
    cudaError_t e = direction == 0
        ? launch_copy_blocks<true>(d_src, d_packed, t->outerA, t->NA, t->innerA, t->itemsize, t->nranks, (cudaStream_t)stream)
        : launch_copy_blocks<true>(d_src, d_packed, t->outerB, t->NB, t->innerB, t->itemsize, t->nranks, (cudaStream_t)stream);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "pack kernel");
}

// direction 0: scatter packed segments into B by axisB blocks; 1: into A by axisA blocks
int b2f_transfer_unpack(b2f_transfer t, int direction, const void* d_packed, void* d_dst, void* stream) {
    if (!t || !d_packed || !d_dst) return B2F_EINVAL;
    cudaError_t e = direction == 0
        ? launch_copy_blocks<false>(d_packed, d_dst, t->outerB, t->NB, t->innerB, t->itemsize, t->nranks, (cudaStream_t)stream)
        : launch_copy_blocks<false>(d_packed, d_dst, t->outerA, t->NA, t->innerA, t->itemsize, t->nranks, (cudaStream_t)stream);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "unpack kernel");
}

static int run_transfer(b2f_transfer t, int direction, const void* d_src, void* d_dst, cudaStream_t st) {
    // geometry of the source side (S) and destination side (D) for this direction
    const long long outerS = direction == 0 ? t->outerA : t->outerB;
    const long long innerS = direction == 0 ? t->innerA : t->innerB;
    const long long NS = direction == 0 ? t->NA : t->NB;
    const long long outerD = direction == 0 ? t->outerB : t->outerA;
    const long long innerD = direction == 0 ? t->innerB : t->innerA;
    const long long ND = direction == 0 ? t->NB : t->NA;
    const size_t bytesS = (size_t)(direction == 0 ? t->volA : t->volB) * t->itemsize;
    const size_t bytesD = (size_t)(direction == 0 ? t->volB : t->volA) * t->itemsize;
    const int p = t->nranks;
    if (p == 1) {
        // both pencils are the whole group-local block: plain copy
        if (d_src != d_dst) {
            cudaError_t e = cudaMemcpyAsync(d_dst, d_src, bytesS, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(transfer, single rank)");
        }
        return B2F_OK;
    }
    if (!t->comm) {
        set_error("transfer was created without a communicator (geometry/pack-only handle)");
        return B2F_EINVAL;
    }
    int rc = need_nccl();
    if (rc) return rc;
    const char* sendbase = (const char*)d_src;
    char* recvbase = (char*)d_dst;
    if (outerS != 1) {
        void* ws;
        if ((rc = workspace(t, 0, bytesS, &ws))) return rc;
        if ((rc = b2f_transfer_pack(t, direction, d_src, ws, st))) return rc;
        sendbase = (const char*)ws;
    }
    if (outerD != 1) {
        void* ws;
        if ((rc = workspace(t, 1, bytesD, &ws))) return rc;
        recvbase = (char*)ws;
    }
    const BlockDist dS = make_dist(NS, p), dD = make_dist(ND, p);
    const long long rowS = outerS * innerS * t->itemsize, rowD = outerD * innerD * t->itemsize;   // bytes per unit of the split axis
    ncclResult_t r = g_nccl.GroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int k = 0; k < p; ++k) {
        // start with my right-hand neighbour so that the p ranks do not all hit peer 0 first
        const int i = (t->rank + k) % p;
        const size_t sbytes = (size_t)(dS.len(i) * rowS), soff = (size_t)(dS.start(i) * rowS);
        const size_t rbytes = (size_t)(dD.len(i) * rowD), roff = (size_t)(dD.start(i) * rowD);
        if (i == t->rank) {
            cudaError_t e = cudaMemcpyAsync(recvbase + roff, sendbase + soff, sbytes, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) {
                g_nccl.GroupEnd();
                return cuda_fail(e, "cudaMemcpyAsync(self segment)");
            }
            continue;
        }
        r = g_nccl.Send(sendbase + soff, sbytes, ncclUint8, i, t->comm->comm, st);
        if (r == ncclSuccess) r = g_nccl.Recv(recvbase + roff, rbytes, ncclUint8, i, t->comm->comm, st);
        if (r != ncclSuccess) {
            g_nccl.GroupEnd();
            return nccl_fail(r, "ncclSend/ncclRecv");
        }
    }
    r = g_nccl.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    if (outerD != 1) {
        if ((rc = b2f_transfer_unpack(t, direction, recvbase, d_dst, st))) return rc;
    }
    return B2F_OK;
}

int b2f_transfer_forward(b2f_transfer t, const void* d_A, void* d_B, void* stream) {
    if (!t || !d_A || !d_B) {
        set_error("b2f_transfer_forward: null argument");
        return B2F_EINVAL;
    }
    return run_transfer(t, 0, d_A, d_B, (cudaStream_t)stream);
}

int b2f_transfer_backward(b2f_transfer t, const void* d_B, void* d_A, void* stream) {
    if (!t || !d_A || !d_B) {
        set_error("b2f_transfer_backward: null argument");
        return B2F_EINVAL;
    }
    return run_transfer(t, 1, d_B, d_A, (cudaStream_t)stream);
}

}  // extern "C"

// ---- peer-memory path -----------------------------------------------------------
static cudaError_t launch_put(const PutParams& prm, cudaStream_t st) {
    long long most = 0;
    for (int k = 0; k < prm.npeers; ++k) most = prm.peer[k].units > most ? prm.peer[k].units : most;
    if (most == 0) return cudaSuccess;
    long long blocks = (most + 4 * 256 - 1) / (4 * 256);
    const long long nsm = sm_count();
    const long long cap = (long long)option("put_ctas_per_peer", (nsm * 8) / prm.npeers > nsm ? (nsm * 8) / prm.npeers : nsm);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks, (unsigned)prm.npeers);
    switch (prm.vec) {
        case 16: put_blocks_kernel<uint4><<<grid, 256, 0, st>>>(prm); break;
        case 8: put_blocks_kernel<uint2><<<grid, 256, 0, st>>>(prm); break;
        case 4: put_blocks_kernel<uint32_t><<<grid, 256, 0, st>>>(prm); break;
        case 2: put_blocks_kernel<uint16_t><<<grid, 256, 0, st>>>(prm); break;
        default: put_blocks_kernel<uint8_t><<<grid, 256, 0, st>>>(prm); break;
    }
    count_launch();
    return cudaGetLastError();
}

static int build_put(b2f_transfer t, int direction, const void* d_src, void* const* peer_dst, PutParams* prm) {
    if (t->nranks > B2F_PUT_MAX_PEERS) {
        set_error("peer-memory transfer supports groups of up to " + std::to_string(B2F_PUT_MAX_PEERS) + " ranks");
        return B2F_EUNSUPPORTED;
    }
    const int axisS = direction == 0 ? t->axisA : t->axisB, axisD = direction == 0 ? t->axisB : t->axisA;
    if (put_build(prm, t->ndims, t->shape.data(), t->itemsize, axisS, axisD, t->nranks, t->rank, d_src, peer_dst)) {
        set_error("put_build failed");
        return B2F_EINVAL;
    }
    return B2F_OK;
}

// Stream-ordered barrier over the group.  When it completes on this rank's stream every
// peer has reached it on its own stream, i.e. everything the peers enqueued before it
// (their kernels reading or writing the windows) is done.
//
// Flag form (default once the peers' flag arrays are mapped): one tiny kernel; thread j
// publishes this rank's arrival count in peer j's array (st.release.sys over NVLink) and
// spins on the count peer j left here (ld.acquire.sys).  No NCCL kernel, no proxy thread:
// a few microseconds instead of an all-reduce launch, which matters once a redistribution
// is cut into chunks with a barrier each.  Counts only grow (epoch = number of barriers of
// this transfer so far, the same on every rank because barriers are enqueued collectively).
struct FlagBarrier {
    unsigned long long* flags[B2F_MAX_PEERS];
    unsigned long long epoch;
    int p, rank;
};
__global__ void __launch_bounds__(32) flag_barrier_kernel(const FlagBarrier fb) {
    const int j = threadIdx.x;
    __threadfence_system();
    if (j < fb.p && j != fb.rank) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fb.flags[j] + fb.rank), "l"(fb.epoch) : "memory");
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fb.flags[fb.rank] + j) : "memory");
        } while (seen < fb.epoch);
    }
    __syncthreads();
    __threadfence_system();
}

// Fallback: a 1-int all-reduce on the group's NCCL communicator.
static int nccl_barrier(b2f_comm c, cudaStream_t st) {
    ncclResult_t r = g_nccl.AllReduce(c->d_flag, c->d_flag + 1, 1, ncclInt32, ncclSum, c->comm, st);
    return r == ncclSuccess ? B2F_OK : nccl_fail(r, "ncclAllReduce(barrier)");
}

static int group_barrier(b2f_transfer t, cudaStream_t st) {
    if (t->has_flags && option("flag_barrier", 1)) {
        FlagBarrier fb;
        for (int j = 0; j < t->nranks; ++j) fb.flags[j] = t->flags[j];
        fb.epoch = ++t->epoch;
        fb.p = t->nranks;
        fb.rank = t->rank;
        flag_barrier_kernel<<<1, 32, 0, st>>>(fb);
        count_launch();
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? B2F_OK : cuda_fail(e, "flag barrier kernel");
    }
    if (!t->comm) {
        set_error("transfer has neither peer flags nor a communicator for its group barrier");
        return B2F_EINVAL;
    }
    int rc = need_nccl();
    if (rc) return rc;
    return nccl_barrier(t->comm, st);
}

extern "C" {

int b2f_malloc(void** d_ptr, size_t bytes) {
    if (!d_ptr) return B2F_EINVAL;
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "cudaMalloc");
}

int b2f_free(void* d_ptr) {
    cudaError_t e = cudaFree(d_ptr);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "cudaFree");
}

int b2f_ipc_export(const void* d_ptr, void* handle64) {
    if (!d_ptr || !handle64) return B2F_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr));
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
    memcpy(handle64, &h, 64);
    return B2F_OK;
}

int b2f_ipc_open(const void* handle64, void** d_peer) {
    if (!handle64 || !d_peer) return B2F_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(d_peer, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "cudaIpcOpenMemHandle");
}

int b2f_ipc_close(void* d_peer) {
    cudaError_t e = cudaIpcCloseMemHandle(d_peer);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int b2f_transfer_put(b2f_transfer t, int direction, const void* d_src, void* const* peer_dst, void* stream) {
    if (!t || !d_src || !peer_dst || (direction != 0 && direction != 1)) {
        set_error("b2f_transfer_put: bad arguments");
        return B2F_EINVAL;
    }
    PutParams prm;
    int rc = build_put(t, direction, d_src, peer_dst, &prm);
    if (rc) return rc;
    cudaError_t e = launch_put(prm, (cudaStream_t)stream);
    return e == cudaSuccess ? B2F_OK : cuda_fail(e, "put kernel");
}

// PeerStore of the stage that feeds transfer `t` in `direction`: the stage's last
// step transformed the axis the transfer splits (axisS), its output block is the
// transfer's source array.
static int build_peer_store(b2f_transfer t, int direction, b2f_plan plan, void* const* peer_dst, PeerStore* ps) {
    int axis, precision, ndims;
    long long n;
    const long long* oshape;
    if (!plan_scatter_info(plan, &axis, &n, &precision, &oshape, &ndims)) {
        set_error("fused redistribution: the stage does not end in a Stockham step");
        return B2F_EUNSUPPORTED;
    }
    const int axisS = direction == 0 ? t->axisA : t->axisB, axisD = direction == 0 ? t->axisB : t->axisA;
    const std::vector<long long>& subS = direction == 0 ? t->subA : t->subB;
    if (ndims != t->ndims || axis != axisS || t->itemsize != 2 * precision || t->nranks > B2F_MAX_PEERS) {
        set_error("fused redistribution: the stage's last axis is not the axis the transfer splits");
        return B2F_EUNSUPPORTED;
    }
    for (int i = 0; i < ndims; ++i)
        if (oshape[i] != subS[i]) {
            set_error("fused redistribution: stage output shape differs from the transfer's source block");
            return B2F_EINVAL;
        }
    if (peer_store_build(ps, ndims, t->shape.data(), axisS, axisD, t->nranks, t->rank, peer_dst)) {
        set_error("fused redistribution: bad geometry");
        return B2F_EINVAL;
    }
    return B2F_OK;
}

static int barrier_hook(void* ctx, cudaStream_t st) { return group_barrier((b2f_transfer)ctx, st); }

static int execute_scatter_impl(b2f_plan plan, const void* d_in, void* d_work, double scale, b2f_transfer t,
                                int direction, void* const* peer_dst, int sync_flags, const ChunkSpec* chunk,
                                void* stream) {
    if (!plan || !d_in || !t || !peer_dst || (direction != 0 && direction != 1)) {
        set_error("b2f_execute_scatter: bad arguments");
        return B2F_EINVAL;
    }
    PeerStore ps;
    int rc = build_peer_store(t, direction, plan, peer_dst, &ps);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool multi = t->nranks > 1;
    const bool enter = multi && (sync_flags & 1), leave = multi && (sync_flags & 2);
    if ((enter || leave) && !t->has_flags && !t->comm) {
        set_error("transfer was created without a communicator");
        return B2F_EINVAL;
    }
    // d_work receives the intermediate of a multi-axis stage; a single-step stage never touches it
    rc = run_plan(plan, d_in, d_work ? d_work : const_cast<void*>(d_in), scale, st, &ps,
                  enter ? barrier_hook : nullptr, enter ? (void*)t : nullptr, chunk);
    if (rc) return rc;
    return leave ? group_barrier(t, st) : B2F_OK;
}

int b2f_execute_scatter(b2f_plan plan, const void* d_in, void* d_work, double scale, b2f_transfer t, int direction,
                        void* const* peer_dst, int sync, void* stream) {
    return execute_scatter_impl(plan, d_in, d_work, scale, t, direction, peer_dst, sync ? 3 : 0, nullptr, stream);
}

int b2f_execute_scatter_chunk(b2f_plan plan, const void* d_in, double scale, b2f_transfer t, int direction,
                              void* const* peer_dst, int sync_flags, int mode, int64_t begin, int64_t count,
                              int64_t view_outer, int64_t view_ostride, int grid_cap, void* stream) {
    ChunkSpec ch{mode, begin, count, view_outer, view_ostride, grid_cap};
    return execute_scatter_impl(plan, d_in, nullptr, scale, t, direction, peer_dst, sync_flags, &ch, stream);
}

int b2f_transfer_set_flags(b2f_transfer t, void* const* peer_flags) {
    if (!t) return B2F_EINVAL;
    if (!peer_flags) {
        t->has_flags = false;
        return B2F_OK;
    }
    if (t->nranks > B2F_MAX_PEERS) {
        set_error("flag barrier supports groups of up to " + std::to_string(B2F_MAX_PEERS) + " ranks");
        return B2F_EUNSUPPORTED;
    }
    for (int j = 0; j < t->nranks; ++j) {
        if (!peer_flags[j] || ((uintptr_t)peer_flags[j] & 7)) {
            set_error("b2f_transfer_set_flags: null or misaligned flag array");
            return B2F_EINVAL;
        }
        t->flags[j] = reinterpret_cast<unsigned long long*>(peer_flags[j]);
    }
    t->epoch = 0;
    t->has_flags = true;
    return B2F_OK;
}

int b2f_transfer_barrier(b2f_transfer t, void* stream) {
    if (!t) return B2F_EINVAL;
    if (t->nranks == 1) return B2F_OK;
    return group_barrier(t, (cudaStream_t)stream);
}

int b2f_plan_can_scatter(b2f_plan plan, b2f_transfer t, int direction) {
    if (!plan || !t) return 0;
    PeerStore ps;
    void* dummy[B2F_MAX_PEERS] = {};
    return build_peer_store(t, direction, plan, dummy, &ps) == B2F_OK ? 1 : 0;
}

int b2f_transfer_exchange_p2p(b2f_transfer t, int direction, const void* d_src, void* const* peer_dst, void* stream) {
    if (!t || !d_src || !peer_dst || (direction != 0 && direction != 1)) {
        set_error("b2f_transfer_exchange_p2p: bad arguments");
        return B2F_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (t->nranks == 1) return b2f_transfer_put(t, direction, d_src, peer_dst, stream);
    if (!t->has_flags && !t->comm) {
        set_error("transfer was created without a communicator");
        return B2F_EINVAL;
    }
    int rc;
    PutParams prm;
    if ((rc = build_put(t, direction, d_src, peer_dst, &prm))) return rc;
    if ((rc = group_barrier(t, st))) return rc;      // the peers are done with their windows
    cudaError_t e = launch_put(prm, st);
    if (e != cudaSuccess) return cuda_fail(e, "put kernel");
    return group_barrier(t, st);                      // every block has landed in my window
}

}  // extern "C"
