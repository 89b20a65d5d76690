// Plain SAME stride-1 convolution for image-like inputs (C <= 4), e.g. the generator's first layer conv1_d
// (generator.py:60: 7x7, 3 -> 32 on the LDR panorama).  K = k*k*C is tiny (147) and the layer is HBM/FMA-bound, so it
// does not go through the tensor-core pipeline: one thread per output pixel keeps F <= 32 accumulators in registers,
// the input patch (k rows x (128 + k - 1) pixels) and the whole kernel variable are staged in shared memory, fp32 FMA
// in the reference's (tap, c) order.  Epilogue: bias, optional LeakyReLU, coalesced stores, per-(sample, filter)
// moments for the instance norm that follows.
#include "sky_common.cuh"

namespace sky {

constexpr int SC_THREADS = 128;   // pixels per block (one row segment)
constexpr int SC_FMAX = 32;

template <int C>
__global__ void __launch_bounds__(SC_THREADS)
conv2d_smallc_kernel(const float *__restrict__ x, const float *__restrict__ kernel, const float *__restrict__ bias,
                     float *__restrict__ y, double *__restrict__ stats, int h, int w, int F, int k, int flags, float slope)
{
    extern __shared__ float sm[];
    const int r = k / 2, pw = SC_THREADS + k - 1;
    float *wts = sm;                          // [k*k*C][SC_FMAX]
    float *patch = sm + k * k * C * SC_FMAX;  // [k][pw][C]
    float *red = patch + k * pw * C;          // [4 warps][SC_FMAX][2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * SC_THREADS, i = blockIdx.y, b = blockIdx.z;

    for (int e = tid; e < k * k * C * SC_FMAX; e += SC_THREADS) {
        const int f = e % SC_FMAX, row = e / SC_FMAX;
        wts[e] = f < F ? kernel[(size_t)row * F + f] : 0.f;
    }
    for (int e = tid; e < k * pw * C; e += SC_THREADS) {
        const int c = e % C, px = (e / C) % pw, a = e / (C * pw);
        const int yy = i + a - r, xx = x0 + px - r;
        patch[e] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? x[(((size_t)b * h + yy) * w + xx) * C + c] : 0.f;
    }
    __syncthreads();

    float acc[SC_FMAX];
#pragma unroll
    for (int f = 0; f < SC_FMAX; ++f) acc[f] = 0.f;
    for (int a = 0; a < k; ++a)
        for (int bb = 0; bb < k; ++bb) {
            const float *pv = patch + (a * pw + tid + bb) * C;
            const float4 *wrow = reinterpret_cast<const float4 *>(wts + (a * k + bb) * C * SC_FMAX);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float v = pv[c];
#pragma unroll
                for (int f4 = 0; f4 < SC_FMAX / 4; ++f4) {
                    const float4 wv = wrow[c * (SC_FMAX / 4) + f4];
                    acc[4 * f4 + 0] = fmaf(v, wv.x, acc[4 * f4 + 0]);
                    acc[4 * f4 + 1] = fmaf(v, wv.y, acc[4 * f4 + 1]);
                    acc[4 * f4 + 2] = fmaf(v, wv.z, acc[4 * f4 + 2]);
                    acc[4 * f4 + 3] = fmaf(v, wv.w, acc[4 * f4 + 3]);
                }
            }
        }
    const int j = x0 + tid;
    const bool ok = j < w;
#pragma unroll
    for (int f = 0; f < SC_FMAX; ++f) {
        float v = acc[f] + (f < F ? __ldg(bias + f) : 0.f);
        if (flags & SKY_EPI_LEAKY_RELU) v = v > 0.f ? v : v * slope;
        acc[f] = v;
    }
    if (ok) {
        float *dst = y + (((size_t)b * h + i) * w + j) * F;
        if (F == SC_FMAX) {
#pragma unroll
            for (int f4 = 0; f4 < SC_FMAX / 4; ++f4)
                reinterpret_cast<float4 *>(dst)[f4] = make_float4(acc[4 * f4], acc[4 * f4 + 1], acc[4 * f4 + 2], acc[4 * f4 + 3]);
        } else {
#pragma unroll
            for (int f = 0; f < SC_FMAX; ++f)
                if (f < F) dst[f] = acc[f];
        }
    }
    if (stats) {
#pragma unroll
        for (int f = 0; f < SC_FMAX; ++f) {
            float s1 = ok ? acc[f] : 0.f, s2 = s1 * s1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) { red[(warp * SC_FMAX + f) * 2] = s1; red[(warp * SC_FMAX + f) * 2 + 1] = s2; }
        }
        __syncthreads();
        if (tid < 2 * F) {
            const int f = tid >> 1, which = tid & 1;
            double s = 0.0;
            for (int wq = 0; wq < SC_THREADS / 32; ++wq) s += (double)red[(wq * SC_FMAX + f) * 2 + which];
            atomicAdd(stats + ((size_t)b * F + f) * 2 + which, s);
        }
    }
}

}  // namespace sky

using namespace sky;

extern "C" int sky_conv2d_smallc_fwd(const float *x, const float *kernel, const float *bias, float *y, double *stats, int B, int h,
                                     int w, int C, int F, int k, int epilogue_flags, float slope, void *stream)
{
    SKY_REQUIRE(x && kernel && bias && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(C >= 1 && C <= 4 && F >= 1 && F <= SC_FMAX && k % 2 == 1 && k <= 11, SKY_ERR_UNSUPPORTED,
                "small-C conv covers C <= 4, F <= 32, odd k <= 11 (got C=%d F=%d k=%d)", C, F, k);
    SKY_REQUIRE(!(epilogue_flags & ~(SKY_EPI_LEAKY_RELU)), SKY_ERR_UNSUPPORTED, "small-C conv supports only the LeakyReLU epilogue");
    SKY_REQUIRE(F != SC_FMAX || ((uintptr_t)y & 15) == 0, SKY_ERR_INVALID, "y must be 16-byte aligned");
    const int pw = SC_THREADS + k - 1;
    const size_t smem = ((size_t)k * k * C * SC_FMAX + (size_t)k * pw * C + 4 * SC_FMAX * 2) * sizeof(float);
    dim3 grid((w + SC_THREADS - 1) / SC_THREADS, h, B);
    cudaStream_t st = (cudaStream_t)stream;
#define SKY_LAUNCH_SC(CC)                                                                                                     \
    do {                                                                                                                      \
        static bool configured = false;                                                                                       \
        if (!configured) {                                                                                                    \
            SKY_CHECK_CUDA(cudaFuncSetAttribute(conv2d_smallc_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
            configured = true;                                                                                                \
        }                                                                                                                     \
        conv2d_smallc_kernel<CC><<<grid, SC_THREADS, smem, st>>>(x, kernel, bias, y, stats, h, w, F, k, epilogue_flags, slope); \
    } while (0)
    switch (C) {
        case 1: SKY_LAUNCH_SC(1); break;
        case 2: SKY_LAUNCH_SC(2); break;
        case 3: SKY_LAUNCH_SC(3); break;
        default: SKY_LAUNCH_SC(4); break;
    }
#undef SKY_LAUNCH_SC
    SKY_CHECK_CUDA(cudaGetLastError());
    return SKY_OK;
}
