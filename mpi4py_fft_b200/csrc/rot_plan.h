// rot_plan.h -- host-side schedule of the rotating kernels (fft_rot.cuh): which
// strides each of the three out-of-place steps of a 3-axis c2c stage uses.
// Pure host code, shared by capi.cu and the CPU emulator (tests/emu/emu_fft.cpp).
#pragma once
#include <vector>

namespace b2f {

// One step: the contiguous axis (length n) of src is transformed and the block is
// stored into dst with its last three axes rotated, [a][b][c] -> [b][c][a]:
//   pencil (i, o) of batch b starts at  b*in_b + i*in_i + o*in_o   (i = a, o = b)
//   result k of that pencil lands at    b*out_b + o*out_o + k*out_n + i
struct RotPlanStep {
    int n;
    long long batches, I, O;
    long long in_i, in_o, in_b, out_o, out_n, out_b;
    int src, dst;   // 0 = caller's input, 1 = caller's output, 2 = plan scratch
};

// default variant (fft_configs.h B2F_ROT_TABLE) of the rotating kernel for a length; -1 = none built
static inline int rot_default(int n) {
    switch (n) {
        case 64: case 128: case 256: case 512: case 1024: case 2048: return 0;
        default: return -1;
    }
}

// [x][y][z] -z-> [y][z][x] -x-> [z][x][y] -y-> [x][y][z]: three out-of-place steps,
// in -> out -> scratch -> out, each reading whole pencils and writing page-local
// rows.  `axes` must be the last three axes of the block (any order: c2c axes
// commute); leading axes are batch.  Returns false when the schedule does not apply.
static inline bool build_rotation(int ndims, const long long* sizes, const int* axes, int naxes,
                                  std::vector<RotPlanStep>* steps, long long* block_elems) {
    if (ndims < 3 || naxes != 3) return false;
    bool seen[3] = {false, false, false};
    for (int k = 0; k < 3; ++k) {
        if (axes[k] < ndims - 3 || axes[k] >= ndims) return false;
        seen[axes[k] - (ndims - 3)] = true;
    }
    if (!(seen[0] && seen[1] && seen[2])) return false;
    const long long X = sizes[ndims - 3], Y = sizes[ndims - 2], Z = sizes[ndims - 1];
    if (rot_default((int)X) < 0 || rot_default((int)Y) < 0 || rot_default((int)Z) < 0) return false;
    if (X > 2048 || Y > 2048 || Z > 2048) return false;
    long long B = 1;
    for (int i = 0; i < ndims - 3; ++i) B *= sizes[i];
    const long long vol = X * Y * Z;
    steps->clear();
    steps->push_back(RotPlanStep{(int)Z, B, X, Y, Y * Z, Z, vol, Z * X, X, vol, 0, 1});
    steps->push_back(RotPlanStep{(int)X, B, Y, Z, Z * X, X, vol, X * Y, Y, vol, 1, 2});
    steps->push_back(RotPlanStep{(int)Y, B, Z, X, X * Y, Y, vol, Y * Z, Z, vol, 2, 1});
    *block_elems = B * vol;
    return true;
}

}  // namespace b2f
