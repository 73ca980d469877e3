// transfer_put.h -- geometry of the peer-memory redistribution (transfer.cu).
//
// Host + device, free of CUDA built-ins, so that tests/emu can replay the put
// kernel's index map on the CPU for every (virtual) rank of a group and compare
// it with the reference's Alltoallw semantics
// (/root/reference/mpi4py_fft/pencil.py:12-29,182-183,200-201).
#pragma once
#include <stdint.h>
#include "fft_core.cuh"   // B2F_HD

namespace b2f {

#define B2F_PUT_MAX_PEERS 16

struct PutPeer {
    char* dst;
    long long b1, rowu;        // rows per (p, m) and units per row
    long long e1D, e2Du;       // destination extents (row axis in units)
    long long o1D, o2Du;       // destination offsets
    long long o1S, o2Su;       // source offsets
    long long units;           // P * b1 * M * rowu
};
struct PutParams {
    const char* src;
    long long M, e1S, e2Su;
    int npeers;
    int vec;                   // bytes per unit
    PutPeer peer[B2F_PUT_MAX_PEERS];
};

// unit u of the block for one peer -> unit index in the source array and in the
// destination array
B2F_HD void put_locate(const PutParams& prm, const PutPeer& pr, long long u, long long* su, long long* du) {
    const long long row = u / pr.rowu;
    const long long c = u - row * pr.rowu;
    const long long t = row / prm.M;
    const long long m = row - t * prm.M;
    const long long p = t / pr.b1;
    const long long i1 = t - p * pr.b1;
    *su = ((p * prm.e1S + i1 + pr.o1S) * prm.M + m) * prm.e2Su + pr.o2Su + c;
    *du = ((p * pr.e1D + i1 + pr.o1D) * prm.M + m) * pr.e2Du + pr.o2Du + c;
}

inline long long put_gcd(long long a, long long b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) {
        const long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// balanced block distribution (== reference pencil.py:5-9)
inline void put_blockdist(long long n, long long p, long long i, long long* len, long long* start) {
    const long long q = n / p, r = n % p;
    *len = q + (i < r ? 1 : 0);
    *start = i * q + (i < r ? i : r);
}

// Fill PutParams for rank `rank` of a group of `p`: the source array is split
// along axisS (it is full along axisS and holds this rank's block of axisD), the
// destination arrays are full along axisD.  `shape` is the group-local shape.
// peer_dst[i] = base address of peer i's destination array (slot k of the
// kernel serves peer (rank + k) % p, so k = 0 is the local copy).
// Returns 0, or -1 when the group is larger than B2F_PUT_MAX_PEERS.
inline int put_build(PutParams* out, int ndims, const long long* shape, int itemsize, int axisS, int axisD,
                     int p, int rank, const void* src, void* const* peer_dst) {
    if (p > B2F_PUT_MAX_PEERS) return -1;
    const int ax1 = axisS < axisD ? axisS : axisD, ax2 = axisS < axisD ? axisD : axisS;
    long long P = 1, M = 1, Q = 1;
    for (int i = 0; i < ax1; ++i) P *= shape[i];
    for (int i = ax1 + 1; i < ax2; ++i) M *= shape[i];
    for (int i = ax2 + 1; i < ndims; ++i) Q *= shape[i];
    const long long NS = shape[axisS], ND = shape[axisD];
    long long nD_me, sD_me;
    put_blockdist(ND, p, rank, &nD_me, &sD_me);
    const bool s_first = (axisS == ax1);
    // source extents along ax1 / ax2
    const long long e1S = s_first ? NS : nD_me, e2S = s_first ? nD_me : NS;
    // widest vector: divides every row length / extent / offset (bytes) and the bases
    long long g = 16;
    g = put_gcd(g, (long long)((uintptr_t)src % 16));
    const long long qb = Q * itemsize;
    g = put_gcd(g, e2S * qb);
    for (int i = 0; i < p; ++i) {
        long long nS_i, sS_i;
        put_blockdist(NS, p, i, &nS_i, &sS_i);
        const long long b2 = s_first ? nD_me : nS_i;
        const long long e2D = s_first ? ND : nS_i;
        const long long o2S = s_first ? 0 : sS_i, o2D = s_first ? sD_me : 0;
        g = put_gcd(g, b2 * qb);
        g = put_gcd(g, e2D * qb);
        g = put_gcd(g, o2S * qb);
        g = put_gcd(g, o2D * qb);
        g = put_gcd(g, (long long)((uintptr_t)peer_dst[i] % 16));
    }
    int v = 16;
    while (v > 1 && g % v) v >>= 1;
    out->src = (const char*)src;
    out->M = M;
    out->e1S = e1S;
    out->e2Su = e2S * qb / v;
    out->npeers = p;
    out->vec = v;
    for (int k = 0; k < p; ++k) {
        const int i = (rank + k) % p;
        long long nS_i, sS_i;
        put_blockdist(NS, p, i, &nS_i, &sS_i);
        PutPeer& pr = out->peer[k];
        pr.dst = (char*)peer_dst[i];
        const long long b1 = s_first ? nS_i : nD_me, b2 = s_first ? nD_me : nS_i;
        pr.b1 = b1;
        pr.rowu = b2 * qb / v;
        // destination array of peer i: extent nS_i along axisS, ND along axisD
        pr.e1D = s_first ? nS_i : ND;
        pr.e2Du = (s_first ? ND : nS_i) * qb / v;
        pr.o1D = s_first ? 0 : sD_me;
        pr.o2Du = (s_first ? sD_me : 0) * qb / v;
        pr.o1S = s_first ? sS_i : 0;
        pr.o2Su = (s_first ? 0 : sS_i) * qb / v;
        pr.units = P * b1 * M * pr.rowu;
    }
    return 0;
}

// PeerStore (fft_core.cuh) of a stage whose last step transformed axisS of the
// group-local `shape` and whose output feeds a transfer axisS -> axisD inside a
// group of p ranks: the fused form of the same redistribution.
inline int peer_store_build(PeerStore* ps, int ndims, const long long* shape, int axisS, int axisD, int p, int rank,
                            void* const* peer_dst) {
    if (p > B2F_MAX_PEERS || axisS == axisD) return -1;
    PeerStore z = {};
    *ps = z;
    const long long NS = shape[axisS], ND = shape[axisD];
    long long nD, sD;
    put_blockdist(ND, p, rank, &nD, &sD);
    ps->p = p;
    ps->q = (int)(NS / p);
    ps->r = (int)(NS % p);
    ps->nD = nD;
    ps->ND = ND;
    ps->sD = sD;
    for (int i = 0; i < p; ++i) ps->base[i] = peer_dst[i];
    long long M = 1, Q = 1;
    if (axisD < axisS) {
        for (int i = axisD + 1; i < axisS; ++i) M *= shape[i];
        for (int i = axisS + 1; i < ndims; ++i) Q *= shape[i];
        ps->mode = 0;
        ps->stride = Q;
    } else {
        for (int i = axisS + 1; i < axisD; ++i) M *= shape[i];
        for (int i = axisD + 1; i < ndims; ++i) Q *= shape[i];
        ps->mode = 1;
        ps->stride = M * ND * Q;
    }
    ps->M = M;
    ps->Q = Q;
    return 0;
}

}  // namespace b2f
