// Declarations shared by the forward conv translation units.
#pragma once
#include "da_geometry.cuh"

namespace sky {

constexpr int BLOCK_M = 128;  // output pixels per tile (UMMA M)
constexpr int BLOCK_K = 32;   // tf32 values per 128-byte swizzled row
constexpr int UMMA_K = 8;     // k per tcgen05.mma, kind::tf32

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int f_pad_of(int F) { return round_up(F < 16 ? 16 : F, 16); }

struct FwdArgs {
    const float *x;
    const float *offsets;       // device [h][k2][2]
    const float *offsets_host;  // host copy of the same table (may be NULL: disables the band-staged kernel)
    const float *packed;
    const float *bias;
    const float *residual;
    float *y;
    double *stats;              // [B][F][2] running (sum, sum of squares) of y per sample and filter, or NULL
    int B, h, w, C, F, k;
    int flags;
    float slope;
    int math_mode;
    int plain_stride;           // 0: distortion-aware sampling; 1 or 2: plain SAME conv with that stride (direct kernel only)
    cudaStream_t stream;
};

// direct-gather kernel (any C): corners read straight from global/L2
int launch_fwd_direct(const FwdArgs &a);
// band-staged persistent kernel (C % 32 == 0): input band in shared memory via TMA.  Returns SKY_ERR_UNSUPPORTED
// (without setting the error text) when no tiling fits, so the caller can take the direct kernel.
int launch_fwd_band(const FwdArgs &a);

}  // namespace sky
