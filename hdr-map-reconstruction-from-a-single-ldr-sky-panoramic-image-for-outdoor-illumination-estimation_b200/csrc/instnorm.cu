// Instance normalisation applied from fused moments (tfa.layers.InstanceNormalization as used at generator.py:15,19,
// 27-33 and 61-85: per (sample, channel) over H x W, biased variance, epsilon 1e-3, affine gamma/beta), with the
// LeakyReLU (generator.py:28) or the residual add (generator.py:35) that follows it folded into the same pass.
//
// The moments (sum, sum of squares, fp64) come from the conv epilogue (sky_da_conv2d_fwd `stats`), so a res-block costs
// one read and one write per activation on top of its two convolutions.
#include "sky_common.cuh"

namespace sky {

constexpr int IN_THREADS = 256;

// grid: (pixel chunks, B).  Each CTA first derives scale/shift for every channel of its sample, then streams pixels.
__global__ void __launch_bounds__(IN_THREADS)
instnorm_apply_kernel(const float *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ gamma,
                      const float *__restrict__ beta, const float *__restrict__ residual, float *__restrict__ y, int hw, int F,
                      float eps, int flags, float slope, int pix_per_cta)
{
    extern __shared__ float ss[];   // scale[F], shift[F]
    float *scale = ss, *shift = ss + F;
    const int b = blockIdx.y;
    for (int f = threadIdx.x; f < F; f += IN_THREADS) {
        const double s1 = stats[((size_t)b * F + f) * 2 + 0], s2 = stats[((size_t)b * F + f) * 2 + 1];
        const double mean = s1 / hw;
        double var = s2 / hw - mean * mean;     // tf.nn.moments: biased variance
        var = var < 0.0 ? 0.0 : var;
        // tf.nn.batch_normalization: inv = rsqrt(var + eps) * gamma ; y = x * inv + (beta - mean * inv)
        const float inv = rsqrtf((float)var + eps) * gamma[f];
        scale[f] = inv;
        shift[f] = beta[f] - (float)mean * inv;
    }
    __syncthreads();
    const size_t base = ((size_t)b * hw + (size_t)blockIdx.x * pix_per_cta) * F;
    const int npix = min(pix_per_cta, hw - blockIdx.x * pix_per_cta);
    const int total4 = npix * F / 4;            // F % 4 == 0 (checked by the entry point)
    const float4 *xv = reinterpret_cast<const float4 *>(x + base);
    const float4 *rv = residual ? reinterpret_cast<const float4 *>(residual + base) : nullptr;
    float4 *yv = reinterpret_cast<float4 *>(y + base);
    const int f4n = F / 4;
    for (int e = threadIdx.x; e < total4; e += IN_THREADS) {
        const int f = (e % f4n) * 4;
        float4 v = __ldg(xv + e);
        v.x = fmaf(v.x, scale[f + 0], shift[f + 0]);
        v.y = fmaf(v.y, scale[f + 1], shift[f + 1]);
        v.z = fmaf(v.z, scale[f + 2], shift[f + 2]);
        v.w = fmaf(v.w, scale[f + 3], shift[f + 3]);
        if (flags & SKY_EPI_LEAKY_RELU) {
            v.x = v.x > 0.f ? v.x : v.x * slope;
            v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope;
            v.w = v.w > 0.f ? v.w : v.w * slope;
        }
        if (rv) {
            const float4 r = __ldg(rv + e);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        yv[e] = v;
    }
}

}  // namespace sky

using namespace sky;

extern "C" int sky_instnorm_apply(const float *x, const double *stats, const float *gamma, const float *beta, const float *residual,
                                  float *y, int B, int h, int w, int F, float eps, int flags, float slope, void *stream)
{
    SKY_REQUIRE(x && stats && gamma && beta && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(F % 4 == 0, SKY_ERR_UNSUPPORTED, "instance norm kernel needs filters %% 4 == 0 (got %d)", F);
    SKY_REQUIRE(!(flags & SKY_EPI_RESIDUAL) || residual, SKY_ERR_INVALID, "SKY_EPI_RESIDUAL without a residual pointer");
    const int hw = h * w;
    // aim for >= 2 waves of 148 SMs, at least 16 pixels per CTA
    int chunks = (2 * 148 + B - 1) / B;
    int pix = (hw + chunks - 1) / chunks;
    if (pix < 16) pix = 16;
    chunks = (hw + pix - 1) / pix;
    dim3 grid(chunks, B);
    instnorm_apply_kernel<<<grid, IN_THREADS, 2 * F * sizeof(float), (cudaStream_t)stream>>>(
        x, stats, gamma, beta, (flags & SKY_EPI_RESIDUAL) ? residual : nullptr, y, hw, F, eps, flags, slope, pix);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
