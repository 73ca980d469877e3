/*
 * b200fft.h -- C ABI of libb200fft.so: the B200-native replacement for the two
 * native boundaries under mpi4py-fft's PFFT.forward/backward hot path.
 *
 *   (1) serial transforms:  fftw_planxfftn()            /root/reference/mpi4py_fft/fftw/fftw_planxfftn.h:9-17
 *                           fftw_execute_dft|_r2c|_c2r|_r2r (new-array execute)
 *                                                       /root/reference/mpi4py_fft/fftw/fftw_xfftn.pyx:29-48,291-292
 *                           fftw_destroy_plan           /root/reference/mpi4py_fft/fftw/fftw_xfftn.pyx:162-163
 *   (2) global transpose:   MPI_Alltoallw over subarray datatypes
 *                                                       /root/reference/mpi4py_fft/pencil.py:12-29,182-183,200-201
 *
 * Conventions
 *   - plain C linkage, opaque handles, caller-owned *device* buffers, sizes as
 *     int64_t; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success or a negative B2F_E* code;
 *     b2f_last_error() gives a thread-local human readable message.
 *   - enqueue calls take a cudaStream_t (passed as void*) and never
 *     synchronise the host; one forward/backward chain is one stream.
 *   - transform kinds are FFTW's integers, as the reference passes them
 *     (/root/reference/mpi4py_fft/fftw/utilities.pyx:7-26).
 *   - not thread-safe per handle; distinct handles may be used concurrently.
 */
#ifndef B200FFT_H
#define B200FFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_VERSION 100

#if defined(__GNUC__)
#define B2F_API __attribute__((visibility("default")))
#else
#define B2F_API
#endif

/* ---- error codes ------------------------------------------------------- */
#define B2F_OK            0
#define B2F_EINVAL       -1   /* bad argument (shape/axes/kind/handle)        */
#define B2F_EUNSUPPORTED -2   /* valid request this build cannot run          */
#define B2F_ECUDA        -3   /* CUDA runtime error (message has the detail)  */
#define B2F_ENCCL        -4   /* NCCL error or libnccl not loadable           */
#define B2F_ENOMEM       -5

/* ---- transform kinds (== FFTW / reference encoding) -------------------- */
#define B2F_FORWARD  (-1)     /* c2c, exponent sign -1                        */
#define B2F_BACKWARD (+1)     /* c2c, exponent sign +1                        */
#define B2F_R2C      (-2)     /* reference private kind, fftw_planxfftn.c:3-8 */
#define B2F_C2R      (+2)
#define B2F_REDFT00   3       /* DCT-I   */
#define B2F_REDFT01   4       /* DCT-III */
#define B2F_REDFT10   5       /* DCT-II  */
#define B2F_REDFT11   6       /* DCT-IV  */
#define B2F_RODFT00   7       /* DST-I   */
#define B2F_RODFT01   8       /* DST-III */
#define B2F_RODFT10   9       /* DST-II  */
#define B2F_RODFT11  10       /* DST-IV  */

typedef struct b2f_plan_s     *b2f_plan;
typedef struct b2f_comm_s     *b2f_comm;
typedef struct b2f_transfer_s *b2f_transfer;

/* ---- library ----------------------------------------------------------- */
B2F_API int         b2f_version(void);
B2F_API const char *b2f_last_error(void);
/* number of kernels this library has launched since load (bench evidence)  */
B2F_API int64_t     b2f_launch_count(void);
/* tuning knobs, e.g. ("variant", v): pick an alternative kernel instance    */
B2F_API int         b2f_set_option(const char *key, int64_t value);
B2F_API int64_t     b2f_get_option(const char *key);

/* ---- (1) serial transforms -------------------------------------------- */
/*
 * Plan a batched transform over `naxes` axes of a C-contiguous `ndims`-D block.
 * Replaces fftw_planxfftn(ndims, sizes_in, in, sizes_out, out, naxes, axes,
 * kind, flags): same meaning of sizes/axes/kind; `precision` (4 = float,
 * 8 = double) replaces the fftwf_/fftw_ module split (factory.py:22-40);
 * buffers are NOT bound at plan time (every execute is a "new-array execute").
 *   c2c   : kind[0] = B2F_FORWARD|B2F_BACKWARD, sizes_out == sizes_in
 *   r2c   : kind[0] = B2F_R2C, sizes_out[axes[naxes-1]] = n/2+1
 *   c2r   : kind[0] = B2F_C2R, logical sizes are sizes_out (fftw_planxfftn.c:23)
 *   r2r   : kind[i] in 3..10 per axis
 * Results are unnormalised, exactly as FFTW's.
 */
B2F_API int b2f_planxfftn(b2f_plan *plan, int ndims,
                  const int64_t *sizes_in, const int64_t *sizes_out,
                  int naxes, const int *axes, const int *kind,
                  int precision, unsigned flags);

/*
 * Enqueue the transform on `stream`: out = scale * T(in).  `scale` is fused
 * into the last butterfly pass (replaces the separate `output *= M` sweep,
 * /root/reference/mpi4py_fft/libfft.py:412-413).  in == out is allowed for
 * c2c and r2r.  Multi-axis C2R overwrites `d_in` (FFTW's c2r destroys its
 * input as well).
 */
B2F_API int b2f_execute(b2f_plan plan, const void *d_in, void *d_out, double scale, void *stream);

B2F_API int b2f_destroy_plan(b2f_plan plan);

/* one line per stage: kernel family, N, batch geometry (print_plan analogue) */
B2F_API int b2f_plan_describe(b2f_plan plan, char *buf, size_t buflen);

/*
 * Dealiasing step of a padded transform along one axis of an (outer, n, inner)
 * complex block: mode 0 truncates a padded spectrum (n_src = padded, n_dst = kept
 * modes), mode 1 zero-pads it back; half_spectrum != 0 for the r2c layout.
 * Same mode selection and Nyquist treatment as FFTBase._truncation_forward /
 * _padding_backward (/root/reference/mpi4py_fft/libfft.py:263-311); `scale` is
 * applied on the way (the reference's `*= M`, libfft.py:412-413).
 */
B2F_API int b2f_pad_truncate(int mode, int half_spectrum, int precision,
                     const void *d_src, void *d_dst,
                     int64_t outer, int64_t n_src, int64_t n_dst, int64_t inner,
                     double scale, void *stream);

/*
 * Fold that dealiasing step into the transform itself: after this call the SPECTRUM side of a
 * one-axis plan (output of B2F_FORWARD / B2F_R2C, input of B2F_BACKWARD / B2F_C2R) is an
 * (outer, n_keep, inner) block -- the forward kernel's last pass writes only the kept modes
 * (Nyquist rule included), the backward kernel's first pass reads them and takes zeros for the
 * rest, so the padded spectrum never exists in memory and the extra pass of b2f_pad_truncate
 * goes.  sizes_in / sizes_out given at plan time stay the padded ones.  Returns
 * B2F_EUNSUPPORTED when the plan's kernel family has no such flavour (every single-tile Stockham
 * r2c / c2r length and the 2^k and 3 * 2^k c2c lengths -- what a 3/2-rule or factor-2 padded solver
 * transforms -- have one; 5 * 2^k / 7 * 2^k c2c, chirp-z, dense and four-step stages do not); the caller then runs b2f_execute + b2f_pad_truncate.  n_keep = 0 switches it off.
 * Replaces libfft.py:263-311 + 408-422 (truncate / pad, then scale) inside the FFT launch.
 */
B2F_API int b2f_plan_set_truncation(b2f_plan plan, int64_t n_keep);

/* ---- (2) global transpose --------------------------------------------- */
/* NCCL communicator for one 1-D process group (replaces the MPI
 * sub-communicator of MPI_Cart_sub, pencil.py:84-88).  The 128-byte id is
 * produced on one rank and shipped to the others by the host side.         */
B2F_API int b2f_comm_unique_id(void *id128);
B2F_API int b2f_comm_create(b2f_comm *comm, const void *id128, int nranks, int rank);
B2F_API int b2f_comm_destroy(b2f_comm comm);

/*
 * Redistribution plan between pencil A (aligned on axisA) and pencil B
 * (aligned on axisB) inside a group of `nranks` (== Transfer.__init__,
 * pencil.py:154-166).  `shape` is the group-local shape: full along both
 * axisA and axisB.  `comm` may be NULL when nranks == 1, or to get a handle
 * that only serves b2f_transfer_geometry / _pack / _unpack.
 * Creating a transfer does not touch the GPU (geometry only).
 */
B2F_API int b2f_transfer_create(b2f_transfer *t, b2f_comm comm, int nranks, int rank,
                        int ndims, const int64_t *shape, int itemsize,
                        const int64_t *subshapeA, int axisA,
                        const int64_t *subshapeB, int axisB);

/* pack -> all-to-all(v) -> unpack, all on `stream` (== Alltoallw, pencil.py:182,200) */
B2F_API int b2f_transfer_forward (b2f_transfer t, const void *d_A, void *d_B, void *stream);
B2F_API int b2f_transfer_backward(b2f_transfer t, const void *d_B, void *d_A, void *stream);
B2F_API int b2f_transfer_destroy (b2f_transfer t);

/* per-peer element counts / offsets of the packed exchange, arrays of nranks
 * (host only; pins the index maps bit-exactly against the reference)        */
B2F_API int b2f_transfer_geometry(b2f_transfer t,
                          int64_t *send_counts, int64_t *send_offsets,
                          int64_t *recv_counts, int64_t *recv_offsets);

/* the two halves of a transfer as separate device ops (direction 0: A->B,
 * 1: B->A); `packed` holds nranks segments in peer order                    */
B2F_API int b2f_transfer_pack  (b2f_transfer t, int direction, const void *d_src, void *d_packed, void *stream);
B2F_API int b2f_transfer_unpack(b2f_transfer t, int direction, const void *d_packed, void *d_dst, void *stream);

/* ---- (2b) the same redistribution over peer memory (NVLink) ---------------
 * One kernel per transfer: every rank stores the block each peer needs straight
 * into that peer's array -- no pack, no staging, no unpack.  The destination
 * arrays ("windows") are cudaMalloc'ed by the library, exported once with CUDA
 * IPC and mapped by the peers; the 64-byte handles travel through the host side
 * (as the NCCL unique id does).  Replaces the same Alltoallw (pencil.py:182,200). */
B2F_API int b2f_malloc(void **d_ptr, size_t bytes);          /* cudaMalloc: IPC-exportable memory */
B2F_API int b2f_free(void *d_ptr);
B2F_API int b2f_ipc_export(const void *d_ptr, void *handle64);
B2F_API int b2f_ipc_open(const void *handle64, void **d_peer);
B2F_API int b2f_ipc_close(void *d_peer);
/* direction 0: A -> B, 1: B -> A.  peer_dst[i] = base of the destination array on
 * group rank i as mapped in THIS process (peer_dst[rank] = the local one).
 * _put enqueues the kernel only (the caller orders it against the peers);
 * _exchange_p2p = group barrier, put, group barrier, all on `stream`.          */
B2F_API int b2f_transfer_put(b2f_transfer t, int direction, const void *d_src, void *const *peer_dst, void *stream);
B2F_API int b2f_transfer_exchange_p2p(b2f_transfer t, int direction, const void *d_src, void *const *peer_dst, void *stream);

/* Group barrier of the peer-memory path without NCCL: peer_flags[j] = the array of nranks
 * 64-bit arrival counters owned by group rank j (zeroed, in IPC-exported memory such as
 * b2f_malloc's) as mapped in THIS process, peer_flags[rank] = the local one.  Once set,
 * b2f_transfer_exchange_p2p / b2f_execute_scatter(_chunk) order the ranks with one tiny
 * kernel (st.release.sys / ld.acquire.sys on the counters over NVLink) instead of a 1-int
 * ncclAllReduce; NULL returns to NCCL.  b2f_transfer_barrier enqueues the barrier alone.
 * Replaces the synchronisation implied by the blocking MPI_Alltoallw (pencil.py:182,200). */
B2F_API int b2f_transfer_set_flags(b2f_transfer t, void *const *peer_flags);
B2F_API int b2f_transfer_barrier(b2f_transfer t, void *stream);

/* ---- (1)+(2) fused: the stage's last pass stores into the owners' windows ----
 * out = scale * T(in), but the last butterfly pass of the stage writes every
 * point straight into the array of the rank that owns it after transfer `t`
 * (direction 0: A -> B): the stage output never exists locally and the transfer
 * costs no kernel of its own -- NVLink traffic overlaps the transform tile by
 * tile.  Needs the stage's last step to be a Stockham step along the axis `t`
 * splits (b2f_plan_can_scatter says so); otherwise run b2f_execute followed by
 * b2f_transfer_exchange_p2p.  `d_work` holds the intermediate of a multi-axis
 * stage (may be NULL for one axis).  sync != 0 wraps the last step in the two
 * group barriers, as b2f_transfer_exchange_p2p does.
 * Replaces libfft.py:412-413 + pencil.py:182-183 in one launch.                */
B2F_API int b2f_plan_can_scatter(b2f_plan plan, b2f_transfer t, int direction);
B2F_API int b2f_execute_scatter(b2f_plan plan, const void *d_in, void *d_work, double scale,
                        b2f_transfer t, int direction, void *const *peer_dst, int sync, void *stream);

/* ---- pipelining a redistribution with the stage that consumes it ------------
 * Partial launches of a one-axis Stockham stage, so that the producing stage can
 * store chunk c into the peers' windows while the consuming stage already
 * transforms chunk c-1 (two streams, an exit barrier per chunk):
 *   mode 1: inner indices [begin, begin+count) of every pencil row; view_outer > 0
 *           re-views the block as view_outer rows view_ostride elements apart (a
 *           range of the last array axis when other axes follow the transformed one)
 *   mode 2: outer indices [begin, begin+count)
 * grid_cap limits the persistent grid (SMs) so that two stages share the GPU.
 * sync_flags of the scatter form: bit 0 = group barrier before, bit 1 = after.   */
B2F_API int b2f_execute_chunk(b2f_plan plan, const void *d_in, void *d_out, double scale,
                      int mode, int64_t begin, int64_t count, int64_t view_outer, int64_t view_ostride,
                      int grid_cap, void *stream);
B2F_API int b2f_execute_scatter_chunk(b2f_plan plan, const void *d_in, double scale,
                      b2f_transfer t, int direction, void *const *peer_dst, int sync_flags,
                      int mode, int64_t begin, int64_t count, int64_t view_outer, int64_t view_ostride,
                      int grid_cap, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FFT_H */
