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
    const float *aux = nullptr; // SKY_EPI_SUN_BLEND: the sky prediction in the log domain, [M][3]
    float threshold = 0.f;      // SKY_EPI_SUN_BLEND: the alpha ramp width (inference.py:36)
    float *y;
    double *stats;              // [B][F][2] running (sum, sum of squares) of y per sample and filter, or NULL
    int B, h, w, C, F, k;
    int nslices = 0;            // > 0: this launch covers nslices equal slices of F filters each (blockIdx.z), ldF = nslices * F
    int ldF = 0;                // filters of the whole layer when this launch covers a slice of them (0: == F); y, residual,
                                // stats are pre-offset to the slice and strided by ldF
    int flags;
    float slope;
    int math_mode;
    int plain_stride;           // 0: distortion-aware sampling; 1 or 2: plain SAME conv with that stride (direct kernel only)
    // transposed != 0 (direct kernel only): the data gradient of a plain conv of stride `plain_stride` run as a forward pass over dy
    // (x = dy [B,h,w,C]) with the flipped, transposed kernel: output pixel (i, j) of the out_h x out_w gradient map reads tap (a, b) at
    // ((i + a - tp_ph0) / stride, (j + b - tp_pw0) / stride) when both divisions are exact and land inside dy
    int transposed = 0, out_h = 0, out_w = 0, tp_ph0 = 0, tp_pw0 = 0;
    cudaStream_t stream;
};

#ifdef __CUDACC__
// SKY_EPI_SUN_BLEND — the tail of inference.generator_in_step (inference.py:90-92, 106-110) for one pixel.  v: the sun
// prediction in the log domain (what sun_decode returns, generator.py:153-156); sky: the sky prediction in the log domain.
//   alpha = min(1, max(0, max_c(logDecompress(sky_c)) - 1 + T) / T);  v_c = (1 - alpha) * sky_c + alpha * v_c
// (hdr_logDecompression of the blend is the SKY_EPI_LOG_DECOMPRESS step that follows).
__device__ __forceinline__ void sun_blend3(float v[3], const float *__restrict__ sky, float threshold)
{
    const float s0 = __ldg(sky), s1 = __ldg(sky + 1), s2 = __ldg(sky + 2);
    const float gmax = fmaxf(fmaxf(s0, s1), s2);                       // exp is monotonic: max of the linear values
    const float lin = __fdiv_rn(__fsub_rn(expf(__fmul_rn(gmax, 2.3978953f)), 1.f), 10.f);
    const float alpha = fminf(1.f, __fdiv_rn(fmaxf(0.f, __fadd_rn(__fsub_rn(lin, 1.f), threshold)), threshold));
    const float na = __fsub_rn(1.f, alpha);
    v[0] = __fadd_rn(__fmul_rn(na, s0), __fmul_rn(alpha, v[0]));
    v[1] = __fadd_rn(__fmul_rn(na, s1), __fmul_rn(alpha, v[1]));
    v[2] = __fadd_rn(__fmul_rn(na, s2), __fmul_rn(alpha, v[2]));
}
#endif

// direct-gather kernel (any C): corners read straight from global/L2
int launch_fwd_direct(const FwdArgs &a);
// the dispatch behind sky_conv2d_fwd (plain SAME conv: small-filter / band-staged / direct kernel, filter slices)
int conv2d_plain_entry(FwdArgs a, int stride);
// halo of input pixels (relative to an output pixel) the taps of a distortion-aware layer touch (da_conv_fwd_band.cu)
void compute_halo(const float *off, int h, int w, int k, int *hy_lo, int *hy_hi, int *hx_lo, int *hx_hi);
// fp32 CUDA-core kernel for plain stride-1 layers with F <= 4 filters (conv_smallc.cu); SKY_ERR_UNSUPPORTED when it does not apply
int launch_fwd_smallf(const FwdArgs &a);
// band-staged persistent kernel (C % 32 == 0): input band in shared memory via TMA.  Returns SKY_ERR_UNSUPPORTED
// (without setting the error text) when no tiling fits, so the caller can take the direct kernel.
int launch_fwd_band(const FwdArgs &a);

}  // namespace sky
