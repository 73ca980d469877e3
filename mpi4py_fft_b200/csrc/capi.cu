// capi.cu -- the C ABI of libb200fft.so (include/b200fft.h): library state,
// plan construction (which kernel runs which axis) and plan execution.
//
// Plan layout mirrors what fftw_planxfftn builds for FFTW's guru interface
// (/root/reference/mpi4py_fft/fftw/fftw_planxfftn.c:10-77): the transformed axes
// become one batched 1-D step each, everything else is batch.  Every step views
// the C-contiguous block as (outer, n, inner) and runs either
//   - the power-of-two Stockham kernel (c2c, n = 2^k <= 8192), contiguous
//     (inner == 1) or strided flavour, or
//   - the dense-matrix kernel (any kind, any n <= B2F_GENERIC_MAX_N).
// The first step goes in -> out, the rest run in place on out (c2r: in place on
// in, then in -> out), so a d-axis stage costs exactly one read and one write
// of the block per axis and no scratch.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "b200fft.h"
#include "dft_generic.h"
#include "internal.h"
#include "fft_configs.h"
#include "chirpz_host.h"
#include "rot_plan.h"
#include "lengths.h"

#define B2F_GENERIC_MAX_N 4096

namespace b2f {

// ---- library state ---------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};
static std::mutex g_mu;
static std::map<std::string, int64_t> g_opts;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return B2F_ECUDA;
}
int64_t option(const char* key, int64_t dflt) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_opts.find(key);
    if (it != g_opts.end()) return it->second;
    std::string env = std::string("B2F_") + key;
    for (auto& c : env) c = (char)toupper((unsigned char)c);
    const char* v = getenv(env.c_str());
    if (v && *v) return atoll(v);
    return dflt;
}

// ---- device facts -----------------------------------------------------------
int sm_count() {
    static int cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int& c = cache[dev & 63];
    if (!c) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
        c = v;
    }
    return c;
}

typedef void* (*AnyFn)();
void* driver_entry_point(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return q == cudaDriverEntryPointSuccess ? fn : nullptr;
}

}  // namespace b2f
#include <cuda.h>
namespace b2f {
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = (TensorMapEncodeFn)driver_entry_point("cuTensorMapEncodeTiled");
    return fn;
}

// ---- caches ----------------------------------------------------------------
struct DevKey {
    int dev, a;
    long long n;
    bool operator<(const DevKey& o) const {
        if (dev != o.dev) return dev < o.dev;
        if (a != o.a) return a < o.a;
        return n < o.n;
    }
};
struct MatEntry { double* d; int rows, cols; };
static std::map<DevKey, MatEntry> g_mat;

const double* generic_matrix(int kind, long long n, int* rows, int* cols) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevKey key{dev, kind, n};
    auto it = g_mat.find(key);
    if (it != g_mat.end()) {
        *rows = it->second.rows;
        *cols = it->second.cols;
        return it->second.d;
    }
    std::vector<double> M;
    long long r = 0, c = 0;
    if (build_matrix(kind, n, M, r, c) != 0) return nullptr;
    double* d = nullptr;
    if (cudaMalloc((void**)&d, M.size() * 8) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, M.data(), M.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    g_mat[key] = MatEntry{d, (int)r, (int)c};
    *rows = (int)r;
    *cols = (int)c;
    return d;
}

// chirp-z tables on the device, one set per (kind, n, precision, device)
struct ChirpEntry {
    void *pre, *filt, *post;
    long long n_in, n_out;
    int in_mode, out_real, M;
};
static std::map<DevKey, ChirpEntry> g_chirp;

template <class T>
static void* upload_ld(const std::vector<long double>& v) {
    std::vector<T> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = (T)v[i];
    void* d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return d;
}

static const ChirpEntry* chirp_tables(int kind, long long n, int precision) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DevKey key{dev, kind * 16 + precision, n};
    auto it = g_chirp.find(key);
    if (it != g_chirp.end()) return &it->second;
    ChirpSpec sp;
    if (chirp_build(kind, n, &sp)) return nullptr;
    ChirpEntry e;
    e.n_in = sp.n_in;
    e.n_out = sp.n_out;
    e.in_mode = sp.in_mode;
    e.out_real = sp.out_real;
    e.M = sp.M;
    if (precision == 8) {
        e.pre = upload_ld<double>(sp.pre);
        e.filt = upload_ld<double>(sp.filt);
        e.post = upload_ld<double>(sp.post);
    } else {
        e.pre = upload_ld<float>(sp.pre);
        e.filt = upload_ld<float>(sp.filt);
        e.post = upload_ld<float>(sp.post);
    }
    if (!e.pre || !e.filt || !e.post) return nullptr;
    return &(g_chirp[key] = e);
}

// ---- plan ------------------------------------------------------------------
enum StepType { STEP_POW2 = 0, STEP_GENERIC = 1, STEP_REAL = 2, STEP_CHIRP = 3, STEP_FOURSTEP = 4 };
enum Buf { BUF_IN = 0, BUF_OUT = 1 };

struct Step {
    StepType type;
    int kind;
    int axis;
    long long outer, inner;
    long long n_in, n_out;   // elements along the axis in source / destination
    int in_c, out_c;         // reals per element
    Buf src, dst;
    // pow2 (twiddle tables belong to the kernel instance, fft_pow2_inst.cuh)
    bool swap;
    // generic
    const double* M;
    int rows, cols;
    // chirp-z
    const void* chirp;
    // four-step split (lengths.h): n = n1 * n2, `sub` = the plan of the n2-point transforms along j2
    long long n1, n2;
    ::b2f_plan_s* sub;
};

// Default variant of the staged strided kernels (fft_configs.h B2F_TMA_TABLE /
// 100 + B2F_CPA_TABLE) for a length and the distance between consecutive points
// of a pencil; -1 = use the register-path kernel.  "hostile" rows sit in
// different 2 MiB pages: every row of a tile costs an address translation, so
// the widest tile that fits wins (stride probe).
static int staged_default(int n, long long row_bytes) {
    const bool hostile = row_bytes >= (2LL << 20);
    switch (n) {
        case 64: return 0;
        case 128: return hostile ? 0 : 1;
        case 256: return hostile ? 103 : 102;    // cp.async, 256-B rows, twiddles in shared memory (+ L2 prefetch)
        case 384: return hostile ? 101 : -1;     // register path is as fast on friendly strides
        case 768: return 100;
        case 512: return hostile ? 105 : 2;      // profiles/r1d_sweep_opt.txt
        case 1024: return hostile ? 100 : 102;
        case 2048: return 0;
        default: return -1;
    }
}
static int staged_fallback(int n, int precision) {
    // n = 2048 complex64 (the C4 stages, rows at an odd 8-byte pitch): 32 points per thread and
    // 512 threads keep the float kernel out of the 64-register cap that made var 100 spill
    // (stage 1 of 2048^3 r2c: 30.1 -> 25.4 ms, profiles/r2_c4_variants.txt)
    if (n == 2048 && precision == 4) return 102;
    switch (n) {
        case 384: return 101;
        case 768: return 100;
        case 256: return 102;
        case 512: return 105;
        case 1024: return 102;
        case 2048: return 100;
        default: return -1;
    }
}

// does the chirp-z convolution of this transform fit the largest tile (M <= 8192)?
static bool chirp_fits(int kind, long long n) {
    const long long n_out = kind == B2F_R2C ? n / 2 + 1 : n;
    return n >= 1 && n + n_out - 1 <= B2F_POW2_MAX_N && !(kind == B2F_REDFT00 && n < 2);
}

}  // namespace b2f

using namespace b2f;

struct b2f_plan_s {
    int precision;
    int ndims;
    int kind0;
    std::vector<long long> sizes_in, sizes_out;
    std::vector<Step> steps;
    // alternative schedule for out-of-place execution (empty: none)
    std::vector<RotPlanStep> rot;
    bool rot_swap = false;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // dealiasing folded into a one-axis stage (b2f_plan_set_truncation): modes kept on the spectrum side
    long long trunc_keep = 0;
};

static int add_step(b2f_plan_s* pl, int kind, int axis, const std::vector<long long>& shp_in,
                    const std::vector<long long>& shp_out, int in_c, int out_c, Buf src, Buf dst) {
    Step s;
    memset(&s, 0, sizeof(s));
    s.kind = kind;
    s.axis = axis;
    s.outer = 1;
    s.inner = 1;
    for (int i = 0; i < axis; ++i) s.outer *= shp_in[i];
    for (int i = axis + 1; i < pl->ndims; ++i) s.inner *= shp_in[i];
    s.n_in = shp_in[axis];
    s.n_out = shp_out[axis];
    s.in_c = in_c;
    s.out_c = out_c;
    s.src = src;
    s.dst = dst;
    // logical transform length: the real side for r2c / c2r
    const long long n = (kind == B2F_C2R) ? s.n_out : s.n_in;
    FourStep fs;
    if ((kind == B2F_FORWARD || kind == B2F_BACKWARD) && is_stockham(n) && option("stockham", 1)) {
        s.type = STEP_POW2;
        s.swap = (kind == B2F_BACKWARD);
    } else if ((kind == B2F_FORWARD || kind == B2F_BACKWARD) && !is_stockham(n) && !chirp_fits(kind, n) &&
               option("stockham", 1) && fourstep_split(n, &fs)) {
        // beyond one tile (2^k > 8192, other lengths > 4096): two transforms of lengths that have kernels
        s.type = STEP_FOURSTEP;
        s.swap = (kind == B2F_BACKWARD);
        s.n1 = fs.n1;
        s.n2 = fs.n2;
        // step 1 as a plan of its own: the block viewed (outer, n2, n1 * inner), transform along the middle axis
        const int64_t v[3] = {(int64_t)s.outer, (int64_t)fs.n2, (int64_t)(fs.n1 * s.inner)};
        const int ax = 1;
        b2f_plan sub = nullptr;
        const int rc = b2f_planxfftn(&sub, 3, v, v, 1, &ax, &kind, pl->precision, 0);
        if (rc) return rc;
        s.sub = sub;
    } else if ((kind == B2F_R2C || kind == B2F_C2R) && n >= 4 && n % 2 == 0 && is_stockham(n / 2) &&
               option("real_engine", 0) != 1 && option("stockham", 1)) {
        // even-length real transform = n/2-point complex Stockham + split/merge pass
        s.type = STEP_REAL;
        if ((kind == B2F_R2C ? s.n_out : s.n_in) != n / 2 + 1) {
            set_error("sizes_in/sizes_out do not match the transform kind along axis " + std::to_string(axis));
            return B2F_EINVAL;
        }
    } else if ((kind == B2F_REDFT00 || kind == B2F_RODFT00) && is_stockham(kind == B2F_REDFT00 ? n - 1 : n + 1) &&
               (kind == B2F_REDFT00 ? n - 1 : n + 1) >= 2 && option("real_engine", 0) != 1 && option("r2r_engine", 0) != 1 &&
               option("generic_engine", 0) == 0 && option("stockham", 1)) {
        // DCT-I of 2^k + 1 (Chebyshev grids) and DST-I of 2^k - 1 points (and the other Stockham families):
        // the real transform of the even / odd extension of length 2(n -+ 1), read through an index map
        s.type = STEP_REAL;
        if (s.n_out != s.n_in || in_c != 1 || out_c != 1) {
            set_error("sizes_in/sizes_out do not match the transform kind along axis " + std::to_string(axis));
            return B2F_EINVAL;
        }
    } else if ((kind == B2F_REDFT10 || kind == B2F_REDFT01 || kind == B2F_RODFT10 || kind == B2F_RODFT01 ||
                kind == B2F_REDFT11 || kind == B2F_RODFT11) && n >= 4 &&
               n % 2 == 0 && is_stockham(n / 2) && option("real_engine", 0) != 1 && option("r2r_engine", 0) != 1 &&
               option("generic_engine", 0) == 0 && option("stockham", 1)) {
        // DCT / DST of kinds II and III of even length: the n/2-point complex Stockham transform of the
        // permuted (Makhoul) sequence plus a quarter-wave twiddle pass; kinds IV: the n/2-point transform of
        // pre-twiddled pairs, post-twiddled (fft_core.cuh r2r_*, r2r4_*)
        s.type = STEP_REAL;
        if (s.n_out != s.n_in || in_c != 1 || out_c != 1) {
            set_error("sizes_in/sizes_out do not match the transform kind along axis " + std::to_string(axis));
            return B2F_EINVAL;
        }
    } else if (option("generic_engine", 0) != 1 && (n > option("dense_max", 32) || option("generic_engine", 0) == 2) &&
               chirp_fits(kind, n)) {
        // any other length / kind: chirp-z convolution on two power-of-two FFTs
        const ChirpEntry* ce = chirp_tables(kind, n, pl->precision);
        if (!ce) {
            set_error("cannot build chirp-z tables (kind " + std::to_string(kind) + ", n " + std::to_string(n) + ")");
            return B2F_EINVAL;
        }
        const long long stored_in = ce->in_mode == 2 ? n / 2 + 1 : ce->n_in;
        if (s.n_in != stored_in || s.n_out != ce->n_out || in_c != (ce->in_mode == 1 ? 1 : 2) ||
            out_c != (ce->out_real ? 1 : 2)) {
            set_error("sizes_in/sizes_out do not match the transform kind along axis " + std::to_string(axis));
            return B2F_EINVAL;
        }
        s.type = STEP_CHIRP;
        s.chirp = ce;
    } else {
        if (n > B2F_GENERIC_MAX_N) {
            set_error("transform length " + std::to_string(n) + " of kind " + std::to_string(kind) +
                      " is not supported by this build (c2c: every length with a Stockham kernel, products n1 * n2 of a 2^k "
                      "in 64..2048 and such a length, any other length up to 4096; r2c/c2r: twice a Stockham length or <= 4096)");
            return B2F_EUNSUPPORTED;
        }
        s.type = STEP_GENERIC;
        s.M = generic_matrix(kind, n, &s.rows, &s.cols);
        if (!s.M) {
            set_error("cannot build transform matrix (kind " + std::to_string(kind) + ", n " + std::to_string(n) + ")");
            return B2F_EINVAL;
        }
        if (s.rows != s.n_out * out_c || s.cols != s.n_in * in_c) {
            set_error("sizes_in/sizes_out do not match the transform kind along axis " + std::to_string(axis));
            return B2F_EINVAL;
        }
    }
    pl->steps.push_back(s);
    return B2F_OK;
}

extern "C" {

int b2f_version(void) { return B2F_VERSION; }
const char* b2f_last_error(void) { return g_err.c_str(); }
int64_t b2f_launch_count(void) { return g_launches.load(); }

int b2f_set_option(const char* key, int64_t value) {
    if (!key) return B2F_EINVAL;
    std::lock_guard<std::mutex> lk(g_mu);
    g_opts[key] = value;
    return B2F_OK;
}
int64_t b2f_get_option(const char* key) {
    if (!key) return 0;
    if (!strcmp(key, "sm_count")) return sm_count();   // device fact, read-only
    return option(key, 0);
}

int b2f_planxfftn(b2f_plan* plan, int ndims, const int64_t* sizes_in, const int64_t* sizes_out,
                  int naxes, const int* axes, const int* kind, int precision, unsigned flags) {
    (void)flags;
    g_err.clear();   // the message reported with a failure is this call's, not an earlier one's
    if (!plan || !sizes_in || !sizes_out || !axes || !kind || ndims < 1 || naxes < 1 || naxes > ndims) {
        set_error("b2f_planxfftn: bad arguments");
        return B2F_EINVAL;
    }
    if (precision != 4 && precision != 8) {
        set_error("precision must be 4 (float) or 8 (double); long double has no device type");
        return B2F_EUNSUPPORTED;
    }
    std::vector<int> ax(axes, axes + naxes);
    std::vector<char> seen(ndims, 0);
    for (int i = 0; i < naxes; ++i) {
        if (ax[i] < 0) ax[i] += ndims;
        if (ax[i] < 0 || ax[i] >= ndims || seen[ax[i]]) {
            set_error("b2f_planxfftn: bad or repeated axis");
            return B2F_EINVAL;
        }
        seen[ax[i]] = 1;
    }
    b2f_plan_s* pl = new b2f_plan_s;
    pl->precision = precision;
    pl->ndims = ndims;
    pl->kind0 = kind[0];
    pl->sizes_in.assign(sizes_in, sizes_in + ndims);
    pl->sizes_out.assign(sizes_out, sizes_out + ndims);
    for (int i = 0; i < ndims; ++i)
        if (sizes_in[i] < 1 || sizes_out[i] < 1) {
            delete pl;
            set_error("b2f_planxfftn: sizes must be positive");
            return B2F_EINVAL;
        }
    const int k0 = kind[0];
    const int last = ax[naxes - 1];
    int rc = B2F_OK;
    auto same_except = [&](int except) {
        for (int i = 0; i < ndims; ++i)
            if (i != except && sizes_in[i] != sizes_out[i]) return false;
        return true;
    };
    if (k0 == B2F_FORWARD || k0 == B2F_BACKWARD) {
        if (!same_except(-1)) rc = B2F_EINVAL;
        // c2c axes commute: the LAST listed axis runs last, because it is the aligned
        // axis of the pencil (mpifft.py:311,321) -- the one the following transfer
        // splits -- and only the last step can store into the peers' windows
        for (int i = 0; i < naxes && rc == B2F_OK; ++i)
            rc = add_step(pl, k0, ax[i], pl->sizes_in, pl->sizes_out, 2, 2, i == 0 ? BUF_IN : BUF_OUT, BUF_OUT);
        if (rc == B2F_OK) {
            long long elems = 0;
            if (build_rotation(ndims, pl->sizes_in.data(), ax.data(), naxes, &pl->rot, &elems))
                pl->scratch_bytes = (size_t)elems * 2 * precision;
            pl->rot_swap = (k0 == B2F_BACKWARD);
        }
    } else if (k0 == B2F_R2C) {
        if (!same_except(last) || sizes_out[last] != sizes_in[last] / 2 + 1) rc = B2F_EINVAL;
        if (rc == B2F_OK) rc = add_step(pl, B2F_R2C, last, pl->sizes_in, pl->sizes_out, 1, 2, BUF_IN, BUF_OUT);
        for (int i = naxes - 2; i >= 0 && rc == B2F_OK; --i)
            rc = add_step(pl, B2F_FORWARD, ax[i], pl->sizes_out, pl->sizes_out, 2, 2, BUF_OUT, BUF_OUT);
    } else if (k0 == B2F_C2R) {
        if (!same_except(last) || sizes_in[last] != sizes_out[last] / 2 + 1) rc = B2F_EINVAL;
        for (int i = 0; i < naxes - 1 && rc == B2F_OK; ++i)
            rc = add_step(pl, B2F_BACKWARD, ax[i], pl->sizes_in, pl->sizes_in, 2, 2, BUF_IN, BUF_IN);
        if (rc == B2F_OK) rc = add_step(pl, B2F_C2R, last, pl->sizes_in, pl->sizes_out, 2, 1, BUF_IN, BUF_OUT);
    } else {
        if (!same_except(-1)) rc = B2F_EINVAL;
        for (int i = naxes - 1; i >= 0 && rc == B2F_OK; --i) {
            if (kind[i] < B2F_REDFT00 || kind[i] > B2F_RODFT11) {
                rc = B2F_EINVAL;
                break;
            }
            rc = add_step(pl, kind[i], ax[i], pl->sizes_in, pl->sizes_out, 1, 1,
                          i == naxes - 1 ? BUF_IN : BUF_OUT, BUF_OUT);
        }
    }
    if (rc == B2F_OK)
        for (const Step& st : pl->steps)
            if (st.type == STEP_FOURSTEP) {
                long long elems = 1;
                for (int i = 0; i < ndims; ++i) elems *= sizes_in[i];
                const size_t need = (size_t)elems * 2 * precision;
                if (need > pl->scratch_bytes) pl->scratch_bytes = need;
            }
    if (rc != B2F_OK) {
        if (rc == B2F_EINVAL && g_err.empty()) set_error("b2f_planxfftn: sizes_in/sizes_out/kind are inconsistent");
        for (Step& st : pl->steps)
            if (st.sub) b2f_destroy_plan(st.sub);
        delete pl;
        return rc;
    }
    *plan = pl;
    return B2F_OK;
}

int b2f_plan_set_truncation(b2f_plan pl, int64_t n_keep) {
    if (!pl) return B2F_EINVAL;
    if (n_keep == 0) {
        pl->trunc_keep = 0;
        return B2F_OK;
    }
    if (pl->steps.size() != 1) {
        set_error("dealiasing can be folded into one-axis stages only");
        return B2F_EUNSUPPORTED;
    }
    const Step& s = pl->steps[0];
    const long long n = (s.kind == B2F_C2R) ? s.n_out : s.n_in;        // logical (padded) length
    const long long full = (s.kind == B2F_R2C || s.kind == B2F_C2R) ? n / 2 + 1 : n;
    if (n_keep < 1 || n_keep > full) {
        set_error("b2f_plan_set_truncation: kept modes must be in [1, padded modes]");
        return B2F_EINVAL;
    }
    // the 5 * 2^k / 7 * 2^k c2c kernels are built without the dealiasing flavour (build time)
    const bool ok = (s.type == STEP_POW2 && !is_mixed57(n)) || (s.type == STEP_REAL && s.kind < B2F_REDFT00);
    if (!ok) {
        set_error("this stage's kernel family has no dealiasing flavour (only single-tile Stockham c2c / r2c / c2r "
                  "stages have one); use b2f_pad_truncate");
        return B2F_EUNSUPPORTED;
    }
    pl->trunc_keep = n_keep;
    pl->rot.clear();
    return B2F_OK;
}

int b2f_execute(b2f_plan pl, const void* d_in, void* d_out, double scale, void* stream) {
    if (!pl || !d_in || !d_out) {
        set_error("b2f_execute: null plan or buffer");
        return B2F_EINVAL;
    }
    return run_plan(pl, d_in, d_out, scale, (cudaStream_t)stream, nullptr, nullptr, nullptr);
}

int b2f_execute_chunk(b2f_plan pl, const void* d_in, void* d_out, double scale, int mode, int64_t begin, int64_t count,
                      int64_t view_outer, int64_t view_ostride, int grid_cap, void* stream) {
    if (!pl || !d_in || !d_out) {
        set_error("b2f_execute_chunk: null plan or buffer");
        return B2F_EINVAL;
    }
    ChunkSpec ch{mode, begin, count, view_outer, view_ostride, grid_cap};
    return run_plan(pl, d_in, d_out, scale, (cudaStream_t)stream, nullptr, nullptr, nullptr, &ch);
}

}  // extern "C"

namespace b2f {

bool plan_scatter_info(b2f_plan pl, int* axis, long long* n, int* precision, const long long** out_shape, int* ndims) {
    if (!pl || pl->steps.empty()) return false;
    const Step& s = pl->steps.back();
    if (s.type != STEP_POW2) return false;
    *axis = s.axis;
    *n = s.n_out;
    *precision = pl->precision;
    *out_shape = pl->sizes_out.data();
    *ndims = pl->ndims;
    return true;
}

// Runs the steps of a plan.  With `peer` the last step stores into the owners'
// arrays (fused redistribution) and `before_last(ctx)` is enqueued right before
// it (the group barrier that says the peers' windows may be overwritten).
int run_plan(b2f_plan pl, const void* d_in, void* d_out, double scale, cudaStream_t st, const PeerStore* peer_last,
             int (*before_last)(void*, cudaStream_t), void* ctx, const ChunkSpec* chunk) {
    if (chunk && chunk->mode != 0) {
        const Step& s0 = pl->steps.back();
        const bool ok = pl->steps.size() == 1 && s0.type == STEP_POW2 && chunk->begin >= 0 && chunk->count >= 0 &&
                        ((chunk->mode == 1 && s0.inner > 1 && chunk->begin + chunk->count <= s0.inner) ||
                         (chunk->mode == 2 && chunk->begin + chunk->count <= s0.outer && chunk->view_outer == 0));
        if (!ok) {
            set_error("partial execution needs a one-axis Stockham stage and a range inside the block");
            return B2F_EUNSUPPORTED;
        }
    }
    if (!pl->rot.empty() && !peer_last && !before_last && !(chunk && chunk->mode != 0) && option("rotate", 0)) {
        // out of place and not overlapping: the first step writes d_out while d_in is still being read
        const char* a = (const char*)d_in;
        const char* b = (const char*)d_out;
        const bool apart = (a + pl->scratch_bytes <= b) || (b + pl->scratch_bytes <= a);
        if (apart) {
            if (!pl->scratch) {
                cudaError_t e = cudaMalloc(&pl->scratch, pl->scratch_bytes);
                if (e != cudaSuccess) {
                    pl->scratch = nullptr;
                    (void)cudaGetLastError();
                }
            }
            if (pl->scratch) {
                const int vr = (int)option("variant_rot", -1);
                // bulk copies need 16-byte aligned pencils (checked before the first launch: the
                // schedule cannot be abandoned half way)
                const bool ok = !(((uintptr_t)d_in | (uintptr_t)d_out) & 15);
                const int mask = (int)option("rot_step_mask", 7);   // profiling: run a subset of the steps
                for (size_t si = 0; si < pl->rot.size() && ok; ++si) {
                    const RotPlanStep& r = pl->rot[si];
                    if (!((mask >> si) & 1)) continue;
                    void* bufs[3] = {const_cast<void*>(d_in), d_out, pl->scratch};
                    RotStep rs{bufs[r.src], bufs[r.dst], r.batches, r.I, r.O, r.in_i, r.in_o, r.in_b,
                               r.out_o, r.out_n, r.out_b, si + 1 == pl->rot.size() ? scale : 1.0,
                               pl->rot_swap ? 1 : 0, 0};
                    cudaError_t e;
                    if (vr >= 1000) {
                        // register-path engine: the strided kernels with whole pencils in, rotated rows out
                        FftParams prm;
                        memset(&prm, 0, sizeof(prm));
                        prm.scale = rs.scale;
                        prm.swap = rs.swap;
                        prm.in_ostride = r.in_o;
                        prm.out_ostride = r.out_o;
                        prm.in_nstride = 1;
                        prm.out_nstride = r.out_n;
                        prm.in_istride = r.in_i;
                        prm.inner = r.I;
                        e = cudaSuccess;
                        for (long long b = 0; b < r.batches && e == cudaSuccess; ++b) {
                            prm.in = (const char*)rs.in + b * r.in_b * 2 * pl->precision;
                            prm.out = (char*)rs.out + b * r.out_b * 2 * pl->precision;
                            const int v = vr - 1000;
                            if (pl->precision == 8)
                                e = r.n <= 256 ? launch_pow2_small_f64(r.n, v, true, prm, r.O, st)
                                  : r.n <= 1024 ? launch_pow2_mid_f64(r.n, v, true, prm, r.O, st)
                                                : launch_pow2_large_f64(r.n, v, true, prm, r.O, st);
                            else
                                e = r.n <= 256 ? launch_pow2_small_f32(r.n, v, true, prm, r.O, st)
                                  : r.n <= 1024 ? launch_pow2_mid_f32(r.n, v, true, prm, r.O, st)
                                                : launch_pow2_large_f32(r.n, v, true, prm, r.O, st);
                        }
                    } else {
                        e = launch_rot(pl->precision, r.n, vr >= 0 ? vr : rot_default(r.n), rs, st);
                        if (e == cudaErrorInvalidValue && vr >= 0) e = launch_rot(pl->precision, r.n, rot_default(r.n), rs, st);
                    }
                    if (e != cudaSuccess) return cuda_fail(e, "b2f_execute: rotating kernel launch");
                }
                if (ok) return B2F_OK;
            }
        }
    }
    const int variant = (int)option("variant", 0);
    const int variant_c = (int)option("variant_contig", variant);
    const int variant_s = (int)option("variant_strided", variant);
    const int variant_t = (int)option("variant_tma", -1);   // -1: chosen per step (staged_default)
    const int engine = (int)option("strided_engine", 0);
    const bool strict = option("variant_strict", 0) != 0;
    const size_t nsteps = pl->steps.size();
    for (size_t si = 0; si < nsteps; ++si) {
        const Step& s = pl->steps[si];
        const void* src = (s.src == BUF_IN) ? d_in : d_out;
        void* dst = (s.dst == BUF_IN) ? const_cast<void*>(d_in) : d_out;
        const double sc = (si + 1 == nsteps) ? scale : 1.0;
        const PeerStore* peer = (si + 1 == nsteps) ? peer_last : nullptr;
        if (si + 1 == nsteps && before_last) {
            const int rc = before_last(ctx, st);
            if (rc) return rc;
        }
        cudaError_t e;
        if (s.type == STEP_POW2) {
            FftParams prm;
            memset(&prm, 0, sizeof(prm));
            if (peer) prm.peer = *peer;
            prm.in = src;
            prm.out = dst;
            prm.scale = sc;
            prm.swap = s.swap ? 1 : 0;
            const bool strided = s.inner > 1;
            const int n = (int)s.n_in;
            long long outer = s.outer;
            if (strided) {
                prm.in_ostride = prm.out_ostride = s.n_in * s.inner;
                prm.in_nstride = prm.out_nstride = s.inner;
                prm.inner = s.inner;
            } else {
                prm.in_ostride = prm.out_ostride = s.n_in;
                prm.npencils = s.outer;
            }
            if (pl->trunc_keep > 0) {
                // padded transform: the spectrum side (output of a forward, input of a backward
                // transform) holds trunc_keep modes per pencil instead of n
                prm.trunc.n = (int)pl->trunc_keep;
                prm.trunc.np = n;
                long long& side = s.swap ? prm.in_ostride : prm.out_ostride;
                side = strided ? pl->trunc_keep * s.inner : pl->trunc_keep;
            }
            const bool part = chunk && chunk->mode != 0;
            if (part && pl->trunc_keep > 0) {
                set_error("partial execution of a truncating stage is not supported");
                return B2F_EUNSUPPORTED;
            }
            if (part) {
                const long long esz = 2LL * pl->precision;
                long long off;
                if (chunk->mode == 2) {
                    off = chunk->begin * prm.in_ostride;
                    outer = chunk->count;
                    prm.npencils = outer;
                    prm.peer.ooff = chunk->begin;
                } else {
                    off = chunk->begin;
                    prm.inner = chunk->count;
                    prm.peer.ioff = chunk->begin;
                    if (chunk->view_outer > 0) {
                        if (s.outer != 1) {
                            set_error("a re-viewed partial launch needs the transformed axis to be the first one");
                            return B2F_EUNSUPPORTED;
                        }
                        outer = chunk->view_outer;
                        prm.in_ostride = prm.out_ostride = chunk->view_ostride;
                        prm.peer.vstride = chunk->view_ostride;
                    }
                }
                prm.in = src = (const char*)src + off * esz;
                prm.out = dst = (char*)dst + off * esz;
                if (outer == 0 || (strided && prm.inner == 0)) continue;
            }
            int var = strided ? variant_s : variant_c;
            auto launch = [&](int v) -> cudaError_t {
                if (is_mixed(n))
                    return pl->precision == 8 ? launch_pow2_mixed_f64(n, v, strided, prm, outer, st)
                                              : launch_pow2_mixed_f32(n, v, strided, prm, outer, st);
                if (is_mixed57(n))
                    return pl->precision == 8 ? launch_pow2_mixed57_f64(n, v, strided, prm, outer, st)
                                              : launch_pow2_mixed57_f32(n, v, strided, prm, outer, st);
                if (pl->precision == 8) {
                    if (n <= 256) return launch_pow2_small_f64(n, v, strided, prm, outer, st);
                    if (n <= 1024) return launch_pow2_mid_f64(n, v, strided, prm, outer, st);
                    return launch_pow2_large_f64(n, v, strided, prm, outer, st);
                }
                if (n <= 256) return launch_pow2_small_f32(n, v, strided, prm, outer, st);
                if (n <= 1024) return launch_pow2_mid_f32(n, v, strided, prm, outer, st);
                return launch_pow2_large_f32(n, v, strided, prm, outer, st);
            };
            // strided axes: TMA-staged persistent kernel where one is built for n and
            // the layout can be described to TMA, else the register-path kernel
            //   strided_engine: 0 = auto, 1 = register path only, 2 = TMA only (sweeps)
            e = cudaErrorInvalidValue;
            bool done = false;
            if (strided && engine != 1 && pl->trunc_keep == 0) {
                PeerStore peer_part;
                const PeerStore* peer_arg = peer;
                if (peer && part) {
                    peer_part = prm.peer;
                    peer_arg = &peer_part;
                }
                TmaStep ts{src, dst, outer, s.n_in, prm.inner, sc, s.swap ? 1 : 0, peer_arg,
                           part ? prm.in_nstride : 0, part ? prm.in_ostride : 0, part ? chunk->grid_cap : 0};
                auto staged = [&](int v) {
                    return pl->precision == 8 ? launch_tma_f64(n, v, ts, st) : launch_tma_f32(n, v, ts, st);
                };
                if (variant_t >= 0) {
                    e = staged(variant_t);
                } else {
                    // measured defaults (profiles/r1_sweep*): first choice, then the
                    // cp.async flavour for layouts a TMA descriptor cannot express
                    const long long row_bytes = s.inner * 2LL * pl->precision;
                    const int first = staged_default(n, row_bytes);
                    e = first >= 0 ? staged(first) : cudaErrorInvalidValue;
                    if (e == cudaErrorInvalidValue && first >= 0 && first < 100) e = staged(staged_fallback(n, pl->precision));
                }
                done = (e != cudaErrorInvalidValue) || engine == 2;
            }
            if (!done) {
                e = launch(var);
                if (e == cudaErrorInvalidValue && var != 0 && !strict) e = launch(0);   // variant not built for this n
            }
        } else if (s.type == STEP_FOURSTEP) {
            if (peer || (chunk && chunk->mode != 0)) {
                set_error("a four-step stage can neither scatter nor run in pieces");
                return B2F_EUNSUPPORTED;
            }
            if (!pl->scratch) {
                cudaError_t ce = cudaMalloc(&pl->scratch, pl->scratch_bytes);
                if (ce != cudaSuccess) {
                    pl->scratch = nullptr;
                    return cuda_fail(ce, "cudaMalloc(four-step scratch)");
                }
            }
            const long long n1 = s.n1, n2 = s.n2, nn = s.n_in;
            const size_t esz = 2 * (size_t)pl->precision;
            // 1: n2-point transforms along j2 (stride n1 * inner): src -> scratch
            int rc = run_plan(s.sub, src, pl->scratch, 1.0, st, nullptr, nullptr, nullptr);
            if (rc) return rc;
            // 2: twiddle W_n^(j1 k2)
            e = launch_fourstep_twiddle(pl->precision, pl->scratch, s.outer, n2, n1, s.inner, s.swap ? 1 : 0, st);
            if (e != cudaSuccess) return cuda_fail(e, "four-step twiddle kernel");
            // 3: n1-point transforms along j1 for every k2, stored k1-major: X[n2 k1 + k2]
            if (s.inner == 1) {
                // rows of n1 contiguous points in, transposed out: the rotating kernel with I = n2, O = 1
                RotStep rs{pl->scratch, dst, s.outer, n2, 1, n1, n1, nn, nn, n2, nn, sc, s.swap ? 1 : 0, 0};
                e = launch_rot(pl->precision, (int)n1, rot_default((int)n1), rs, st);
            } else {
                FftParams prm;
                memset(&prm, 0, sizeof(prm));
                prm.scale = sc;
                prm.swap = s.swap ? 1 : 0;
                prm.in_ostride = n1 * s.inner;     // k2 -> k2 + 1 in the scratch
                prm.in_nstride = s.inner;          // j1 -> j1 + 1
                prm.out_ostride = s.inner;         // k2 -> k2 + 1 in the result
                prm.out_nstride = n2 * s.inner;    // k1 -> k1 + 1
                prm.inner = s.inner;
                e = cudaSuccess;
                for (long long o = 0; o < s.outer && e == cudaSuccess; ++o) {
                    prm.in = (const char*)pl->scratch + (size_t)o * nn * s.inner * esz;
                    prm.out = (char*)dst + (size_t)o * nn * s.inner * esz;
                    const int n1i = (int)n1;
                    if (pl->precision == 8)
                        e = n1i <= 256 ? launch_pow2_small_f64(n1i, 0, true, prm, n2, st)
                          : n1i <= 1024 ? launch_pow2_mid_f64(n1i, 0, true, prm, n2, st)
                                        : launch_pow2_large_f64(n1i, 0, true, prm, n2, st);
                    else
                        e = n1i <= 256 ? launch_pow2_small_f32(n1i, 0, true, prm, n2, st)
                          : n1i <= 1024 ? launch_pow2_mid_f32(n1i, 0, true, prm, n2, st)
                                        : launch_pow2_large_f32(n1i, 0, true, prm, n2, st);
                }
            }
        } else if (s.type == STEP_REAL) {
            if (peer) {
                set_error("fused redistribution needs a power-of-two c2c Stockham step last");
                return B2F_EUNSUPPORTED;
            }
            FftParams prm;
            memset(&prm, 0, sizeof(prm));
            prm.in = src;
            prm.out = dst;
            prm.scale = sc;
            const bool strided = s.inner > 1;
            const bool r2r = s.kind >= B2F_REDFT00;
            const int mode = s.kind == B2F_R2C ? 1 : s.kind == B2F_C2R ? 2
                           : (s.kind == B2F_REDFT00 || s.kind == B2F_RODFT00) ? 5
                           : (s.kind == B2F_REDFT11 || s.kind == B2F_RODFT11) ? 6
                           : (s.kind == B2F_REDFT10 || s.kind == B2F_RODFT10) ? 3 : 4;
            const long long nreal = (mode == 2) ? s.n_out : s.n_in;
            // complex points of the schedule: half the real length; kinds I: half the extension's length
            const long long nc = s.kind == B2F_REDFT00 ? s.n_in - 1 : s.kind == B2F_RODFT00 ? s.n_in + 1 : nreal / 2;
            prm.flip = (s.kind == B2F_RODFT10 || s.kind == B2F_RODFT01 || s.kind == B2F_RODFT00 || s.kind == B2F_RODFT11) ? 1 : 0;
            if (r2r && !strided) {
                // real rows of n values on both sides
                prm.in_ostride = prm.out_ostride = s.n_in;
                prm.npencils = s.outer;
            } else if (strided) {
                prm.in_ostride = s.n_in * s.inner;
                prm.out_ostride = s.n_out * s.inner;
                prm.in_nstride = prm.out_nstride = s.inner;
                prm.inner = s.inner;
            } else {
                // rows in complex units: 2N reals = N complex on the real side, N+1 on the spectral side
                prm.in_ostride = mode == 1 ? nc : s.n_in;
                prm.out_ostride = mode == 1 ? s.n_out : nc;
                prm.npencils = s.outer;
            }
            if (pl->trunc_keep > 0 && !r2r) {
                prm.trunc.n = (int)pl->trunc_keep;
                prm.trunc.np = (int)nc + 1;
                long long& side = mode == 1 ? prm.out_ostride : prm.in_ostride;
                side = strided ? pl->trunc_keep * s.inner : pl->trunc_keep;
            }
            e = pl->precision == 8 ? launch_real_f64((int)nc, mode, strided, prm, s.outer, st)
                                   : launch_real_f32((int)nc, mode, strided, prm, s.outer, st);
        } else if (s.type == STEP_CHIRP) {
            if (peer) {
                set_error("fused redistribution needs a power-of-two c2c Stockham step last");
                return B2F_EUNSUPPORTED;
            }
            const ChirpEntry* ce = reinterpret_cast<const ChirpEntry*>(s.chirp);
            ChirpParams prm;
            memset(&prm, 0, sizeof(prm));
            prm.in = src;
            prm.out = dst;
            prm.pre = ce->pre;
            prm.filt = ce->filt;
            prm.post = ce->post;
            prm.n_in = (int)ce->n_in;
            prm.n_out = (int)ce->n_out;
            prm.in_mode = ce->in_mode;
            prm.out_real = ce->out_real;
            prm.scale = sc;
            const bool strided = s.inner > 1;
            prm.in_ostride = s.n_in * s.inner;
            prm.out_ostride = s.n_out * s.inner;
            prm.in_nstride = prm.out_nstride = s.inner;
            prm.inner = s.inner;
            prm.npencils = s.outer;
            e = pl->precision == 8 ? launch_chirp_f64(ce->M, strided, prm, s.outer, st)
                                   : launch_chirp_f32(ce->M, strided, prm, s.outer, st);
        } else {
            if (peer) {
                set_error("fused redistribution needs a power-of-two Stockham step last");
                return B2F_EUNSUPPORTED;
            }
            GenericParams g;
            memset(&g, 0, sizeof(g));
            g.in = src;
            g.out = dst;
            g.M = s.M;
            g.npencils = s.outer * s.inner;
            g.inner = s.inner;
            g.in_n = s.n_in;
            g.out_n = s.n_out;
            g.in_c = s.in_c;
            g.out_c = s.out_c;
            g.rows = s.rows;
            g.cols = s.cols;
            g.scale = sc;
            e = launch_generic(pl->precision, g, st);
        }
        if (e != cudaSuccess) return cuda_fail(e, "b2f_execute: kernel launch");
    }
    return B2F_OK;
}

}  // namespace b2f

extern "C" {

int b2f_destroy_plan(b2f_plan pl) {
    if (pl && pl->scratch) cudaFree(pl->scratch);
    if (pl)
        for (Step& st : pl->steps)
            if (st.sub) b2f_destroy_plan(st.sub);
    delete pl;   // tables are cached library-wide
    return B2F_OK;
}

int b2f_plan_describe(b2f_plan pl, char* buf, size_t buflen) {
    if (!pl || !buf || !buflen) return B2F_EINVAL;
    std::string s;
    for (const Step& st : pl->steps) {
        char line[256];
        snprintf(line, sizeof(line), "%s kind=%d axis=%d n_in=%lld n_out=%lld outer=%lld inner=%lld %s->%s\n",
                 st.type == STEP_FOURSTEP ? "stockham-fourstep"
                 : st.type == STEP_POW2 ? (st.inner > 1 ? "stockham-strided" : "stockham-contig")
                 : st.type == STEP_REAL ? (st.kind >= B2F_REDFT00 ? (st.inner > 1 ? "stockham-r2r-strided" : "stockham-r2r-contig")
                                                                   : (st.inner > 1 ? "stockham-real-strided" : "stockham-real-contig"))
                 : st.type == STEP_CHIRP ? (st.inner > 1 ? "chirpz-strided" : "chirpz-contig") : "dense-matrix",
                 st.kind, st.axis, st.n_in, st.n_out, st.outer, st.inner,
                 st.src == BUF_IN ? "in" : "out", st.dst == BUF_IN ? "in" : "out");
        s += line;
    }
    for (const RotPlanStep& r : pl->rot) {
        char line[256];
        static const char* nm[3] = {"in", "out", "scratch"};
        snprintf(line, sizeof(line), "out-of-place alternative: stockham-rotating n=%d batches=%lld I=%lld O=%lld %s->%s\n",
                 r.n, r.batches, r.I, r.O, nm[r.src], nm[r.dst]);
        s += line;
    }
    strncpy(buf, s.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
    return B2F_OK;
}

}  // extern "C"
