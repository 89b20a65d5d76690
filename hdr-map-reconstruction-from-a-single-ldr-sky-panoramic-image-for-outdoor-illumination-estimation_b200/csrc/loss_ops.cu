// Kernels around the generator in the train / test step (train.py): everything that is elementwise or a reduction.
//
//   ldr_synth          train._preprocessing (train.py:54-94): exposure, shot + read noise, relu, clip, per-sample camera response
//                      (tf_utils.apply_rf / interp_1d / sample_1d, tf_utils.py:191-255: DoRF LUT lerp), 8-bit quantisation — one pass
//   hdr_log_codec      tf_utils.hdr_logCompression / hdr_logDecompression (tf_utils.py:263-280)
//   loss_reduce        sum |a-b| (L1, train.py:322), sum (a-1)^2 and sum a^2 (LSGAN terms, train.py:235-237)
//   kl_divergence      tf.keras.losses.KLDivergence (train.py:303): clip to [1e-7, 1], sum y_t log(y_t / y_p)
//   dog_base / dog_l1  tf_utils.DoG (tf_utils.py:61-73): bilinear x2, 3x3 Gaussian (REFLECT), five more 3x3 Gaussians, four differences,
//                      and the sum |DoG_l(a) - DoG_l(b)| of train.py:314-319 without materialising the eight blurred images
//   adam_step          Keras Adam (train_sun.py: optimizer_sun), flat buffers like the RMSprop kernel
// All HBM-bound, fp32 like the reference; reductions end in one fp64 atomic per warp.
#include "sky_common.cuh"

namespace sky {

__device__ __forceinline__ void warp_sum_to(double *out, float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, (double)v);
}

// ---- LDR synthesis ---------------------------------------------------------------------------------------------------
__global__ void ldr_synth_kernel(const float *__restrict__ hdr, const float *__restrict__ t, const float *__restrict__ crf,
                                 const float *__restrict__ sigma_s, const float *__restrict__ sigma_c,
                                 const float *__restrict__ noise_s, const float *__restrict__ noise_c, float *__restrict__ hdr_t,
                                 float *__restrict__ ldr, int B, int hw, int C, int K, int quantize)
{
    const long total = (long)B * hw * C;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C), b = (int)(e / ((long)hw * C));
        float v = __fmul_rn(__ldg(hdr + e), __ldg(t + b));                                    // :64
        if (noise_s) v = __fadd_rn(v, __fmul_rn(__ldg(noise_s + e), __fmul_rn(__ldg(sigma_s + b * C + c), v)));   // :70-72
        if (noise_c) v = __fadd_rn(v, __fmul_rn(__ldg(sigma_c + b * C + c), __ldg(noise_c + e)));                 // :73-74
        v = fmaxf(v, 0.f);                                                                    // :75
        hdr_t[e] = v;
        const float x = fminf(v, 1.f);                                                        // :78
        // apply_rf: pos = (K-1) x; lerp between the two clamped neighbours of the sample's response curve
        const float pos = __fmul_rn((float)(K - 1), x);
        const float y0 = floorf(pos), y1 = __fadd_rn(y0, 1.f);
        const float *rf = crf + (size_t)b * K;
        const float v0 = __ldg(rf + min(max((int)y0, 0), K - 1)), v1 = __ldg(rf + min(max((int)y1, 0), K - 1));
        float out = __fadd_rn(__fmul_rn(__fsub_rn(y1, pos), v0), __fmul_rn(__fsub_rn(pos, y0), v1));
        if (quantize) out = __fdiv_rn(rintf(__fmul_rn(out, 255.f)), 255.f);                   // :84-92 (JPEG round trip omitted)
        ldr[e] = out;
    }
}

__global__ void hdr_log_codec_kernel(const float *__restrict__ x, float *__restrict__ y, long n, int decompress)
{
    const float log11 = logf(11.f);
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float v = x[e];
        y[e] = decompress ? __fdiv_rn(__fsub_rn(expf(__fmul_rn(v, log11)), 1.f), 10.f)
                          : __fdiv_rn(logf(__fadd_rn(1.f, __fmul_rn(10.f, v))), log11);
    }
}

// ---- reductions ------------------------------------------------------------------------------------------------------
__global__ void loss_reduce_kernel(int kind, const float *__restrict__ a, const float *__restrict__ b, long n, double *__restrict__ out)
{
    float acc = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float x = a[e];
        if (kind == 0) acc += fabsf(x - b[e]);
        else if (kind == 1) acc += (x - 1.f) * (x - 1.f);
        else acc += x * x;
    }
    warp_sum_to(out, acc);
}

__global__ void kl_divergence_kernel(const float *__restrict__ yt, const float *__restrict__ yp, long n, double *__restrict__ out)
{
    float acc = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float t = fminf(fmaxf(yt[e], 1e-7f), 1.f), p = fminf(fmaxf(yp[e], 1e-7f), 1.f);
        acc += t * logf(t / p);
    }
    warp_sum_to(out, acc);
}

// ---- Difference of Gaussians -----------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }   // tf.pad REFLECT

// one sample of the x2 bilinear upsampling (tf.image.resize, half-pixel centres) of x [h, w, C] at (oy, ox, c)
__device__ __forceinline__ float up2(const float *__restrict__ img, int h, int w, int C, int c, int oy, int ox)
{
    const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), 0.5f), 0.5f);
    const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), 0.5f), 0.5f);
    const float fly = floorf(fy), flx = floorf(fx);
    const int ylo = max((int)fly, 0), yhi = min((int)ceilf(fy), h - 1);
    const int xlo = max((int)flx, 0), xhi = min((int)ceilf(fx), w - 1);
    const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
    const float tl = __ldg(img + ((size_t)ylo * w + xlo) * C + c), tr = __ldg(img + ((size_t)ylo * w + xhi) * C + c);
    const float bl = __ldg(img + ((size_t)yhi * w + xlo) * C + c), br = __ldg(img + ((size_t)yhi * w + xhi) * C + c);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
}

// 1-D taps of tfa.image.gaussian_filter2d(filter_shape 3): softmax(-x^2 / (2 sigma^2)), x in {-1, 0, 1} -> (edge, centre)
__device__ __forceinline__ void gauss_taps(float sigma, float *edge, float *centre)
{
    const float q = expf(-1.f / (2.f * sigma * sigma));
    const float s = 1.f + 2.f * q;
    *edge = q / s;
    *centre = 1.f / s;
}

// base [B, 2h, 2w, C] = Gaussian_{sigma0}(resize_x2(x)), REFLECT padding
__global__ void dog_base_kernel(const float *__restrict__ x, float *__restrict__ base, int B, int h, int w, int C, float sigma0)
{
    const int H2 = 2 * h, W2 = 2 * w;
    float ke, kc;
    gauss_taps(sigma0, &ke, &kc);
    const long total = (long)B * H2 * W2 * C;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % C), ox = (int)((o / C) % W2), oy = (int)((o / ((long)C * W2)) % H2), b = (int)(o / ((long)C * W2 * H2));
        const float *img = x + (size_t)b * h * w * C;
        float acc = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = reflect(oy + dy, H2);
            const float wy = dy == 0 ? kc : ke;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = reflect(ox + dx, W2);
                acc = fmaf(wy * (dx == 0 ? kc : ke), up2(img, h, w, C, c, yy, xx), acc);
            }
        }
        base[o] = acc;
    }
}

// out4[l] += sum | DoG_l(base_a) - DoG_l(base_b) |,  DoG_l = G_{s[l+1]} - G_{s[l]} applied to the base image (REFLECT)
__global__ void dog_l1_kernel(const float *__restrict__ base_a, const float *__restrict__ base_b, int B, int H2, int W2, int C,
                              double *__restrict__ out4)
{
    // gaussian_kernels1 / gaussian_kernels2 of tf_utils.py:67-68 overlap: five distinct sigmas
    const float sig[5] = { 1.2262735f, 1.5450078f, 1.9465878f, 2.452547f, 3.0900156f };
    float kcc[5], kec[5], kee[5];      // centre*centre, edge*centre, edge*edge
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        float e, c;
        gauss_taps(sig[l], &e, &c);
        kcc[l] = c * c; kec[l] = e * c; kee[l] = e * e;
    }
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    const long total = (long)B * H2 * W2 * C;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % C), ox = (int)((o / C) % W2), oy = (int)((o / ((long)C * W2)) % H2), b = (int)(o / ((long)C * W2 * H2));
        float ctr[2], edge[2] = { 0.f, 0.f }, corner[2] = { 0.f, 0.f };
        const float *src[2] = { base_a + (size_t)b * H2 * W2 * C + c, base_b + (size_t)b * H2 * W2 * C + c };
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = reflect(oy + dy, H2);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = reflect(ox + dx, W2);
                    const float v = __ldg(src[s] + ((size_t)yy * W2 + xx) * C);
                    if (dy == 0 && dx == 0) ctr[s] = v;
                    else if (dy == 0 || dx == 0) edge[s] += v;
                    else corner[s] += v;
                }
            }
        }
        float g[2][5];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int l = 0; l < 5; ++l) g[s][l] = kcc[l] * ctr[s] + kec[l] * edge[s] + kee[l] * corner[s];
#pragma unroll
        for (int l = 0; l < 4; ++l) acc[l] += fabsf((g[0][l + 1] - g[0][l]) - (g[1][l + 1] - g[1][l]));
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) warp_sum_to(out4 + l, acc[l]);
}

// ---- Adam ------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float *__restrict__ w, float *__restrict__ m, float *__restrict__ v, const float *__restrict__ g, long n,
                            float lr_t, float b1, float b2, float eps, float grad_scale)
{
    // 7 streams of n floats (4 reads, 3 writes): 128-bit accesses; the tail (n % 4) is handled by the first threads
    const long n4 = n / 4;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
        const float4 gv = __ldg(reinterpret_cast<const float4 *>(g) + e);
        float4 mv = reinterpret_cast<float4 *>(m)[e], vv = reinterpret_cast<float4 *>(v)[e], wv = reinterpret_cast<float4 *>(w)[e];
        const float ga[4] = { gv.x * grad_scale, gv.y * grad_scale, gv.z * grad_scale, gv.w * grad_scale };
        float *mp = &mv.x, *vp = &vv.x, *wp = &wv.x;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mp[u] = b1 * mp[u] + (1.f - b1) * ga[u];
            vp[u] = b2 * vp[u] + (1.f - b2) * ga[u] * ga[u];
            wp[u] -= lr_t * mp[u] / (sqrtf(vp[u]) + eps);
        }
        reinterpret_cast<float4 *>(m)[e] = mv; reinterpret_cast<float4 *>(v)[e] = vv; reinterpret_cast<float4 *>(w)[e] = wv;
    }
    for (long e = 4 * n4 + blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float gv = g[e] * grad_scale;
        const float mm = b1 * m[e] + (1.f - b1) * gv, vv = b2 * v[e] + (1.f - b2) * gv * gv;
        m[e] = mm; v[e] = vv;
        w[e] -= lr_t * mm / (sqrtf(vv) + eps);
    }
}

// out [n, Cp] = concat(a [n, Ca], b [n, Cb]) zero-padded to Cp channels (discriminator input, discriminator.py:42: tf.concat of the LDR
// and HDR images; 6 -> 8 channels so the 4x4/2 conv gathers 16-byte chunks)
__global__ void concat2_pad_kernel(const float *__restrict__ a, int Ca, const float *__restrict__ b, int Cb, float *__restrict__ out,
                                   int Cp, long n)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n * Cp; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % Cp);
        const long p = e / Cp;
        out[e] = c < Ca ? a[p * Ca + c] : (c < Ca + Cb ? b[p * Cb + (c - Ca)] : 0.f);
    }
}

// Vgg16.call preprocessing (vgg16.py:136-144): out [n, 4] = (255 x_c - mean_c, 0) — 3 -> 4 channels for the 16-byte gather
__global__ void vgg_preprocess_kernel(const float *__restrict__ x, float *__restrict__ out, long n, float m0, float m1, float m2)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < n; p += (long)gridDim.x * blockDim.x) {
        const float *s = x + p * 3;
        reinterpret_cast<float4 *>(out)[p] = make_float4(__fsub_rn(__fmul_rn(255.f, s[0]), m0), __fsub_rn(__fmul_rn(255.f, s[1]), m1),
                                                         __fsub_rn(__fmul_rn(255.f, s[2]), m2), 0.f);
    }
}

static int blocks_for(long total, int cap = 148 * 8)
{
    long b = (total + 255) / 256;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace sky

using namespace sky;

extern "C" int sky_ldr_synth(const float *hdr, const float *t, const float *crf, const float *sigma_s, const float *sigma_c,
                             const float *noise_s, const float *noise_c, float *hdr_t, float *ldr, int B, int hw, int C, int K,
                             int quantize, void *stream)
{
    SKY_REQUIRE(hdr && t && crf && hdr_t && ldr && B > 0 && hw > 0 && C > 0 && K >= 2, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE((!noise_s || sigma_s) && (!noise_c || sigma_c), SKY_ERR_INVALID, "noise without its sigma");
    ldr_synth_kernel<<<blocks_for((long)B * hw * C), 256, 0, (cudaStream_t)stream>>>(hdr, t, crf, sigma_s, sigma_c, noise_s, noise_c, hdr_t,
                                                                                     ldr, B, hw, C, K, quantize);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_hdr_log_codec(const float *x, float *y, long n, int decompress, void *stream)
{
    SKY_REQUIRE(x && y && n > 0, SKY_ERR_INVALID, "bad arguments");
    hdr_log_codec_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, decompress);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_loss_reduce(int kind, const float *a, const float *b, long n, double *out, void *stream)
{
    SKY_REQUIRE(a && out && n > 0 && kind >= 0 && kind <= 2 && (kind != 0 || b), SKY_ERR_INVALID, "bad arguments");
    loss_reduce_kernel<<<blocks_for(n, 148 * 4), 256, 0, (cudaStream_t)stream>>>(kind, a, b, n, out);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_kl_divergence(const float *y_true, const float *y_pred, long n, double *out, void *stream)
{
    SKY_REQUIRE(y_true && y_pred && out && n > 0, SKY_ERR_INVALID, "bad arguments");
    kl_divergence_kernel<<<blocks_for(n, 148 * 4), 256, 0, (cudaStream_t)stream>>>(y_true, y_pred, n, out);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dog_base(const float *x, float *base, int B, int h, int w, int C, void *stream)
{
    SKY_REQUIRE(x && base && B > 0 && h > 1 && w > 1 && C > 0, SKY_ERR_INVALID, "bad arguments");
    dog_base_kernel<<<blocks_for((long)B * 4 * h * w * C), 256, 0, (cudaStream_t)stream>>>(x, base, B, h, w, C, 1.2489996f);   // tf_utils.py:61
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dog_l1(const float *base_a, const float *base_b, int B, int H2, int W2, int C, double *out4, void *stream)
{
    SKY_REQUIRE(base_a && base_b && out4 && B > 0 && H2 > 1 && W2 > 1 && C > 0, SKY_ERR_INVALID, "bad arguments");
    dog_l1_kernel<<<blocks_for((long)B * H2 * W2 * C, 148 * 4), 256, 0, (cudaStream_t)stream>>>(base_a, base_b, B, H2, W2, C, out4);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_adam_step(float *w, float *m, float *v, const float *g, long n, float lr, float beta1, float beta2, float eps,
                             long step, float grad_scale, void *stream)
{
    SKY_REQUIRE(w && m && v && g && n > 0 && step >= 1, SKY_ERR_INVALID, "bad arguments");
    // Keras: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
    SKY_REQUIRE((((uintptr_t)w | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) == 0, SKY_ERR_INVALID, "w, m, v, g must be 16-byte aligned");
    adam_kernel<<<blocks_for((n + 3) / 4, 148 * 16), 256, 0, (cudaStream_t)stream>>>(w, m, v, g, n, lr_t, beta1, beta2, eps, grad_scale);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_concat2_pad(const float *a, int Ca, const float *b, int Cb, float *out, int Cp, long n, void *stream)
{
    SKY_REQUIRE(a && b && out && Ca > 0 && Cb > 0 && Cp >= Ca + Cb && n > 0, SKY_ERR_INVALID, "bad arguments");
    concat2_pad_kernel<<<blocks_for(n * Cp), 256, 0, (cudaStream_t)stream>>>(a, Ca, b, Cb, out, Cp, n);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_vgg_preprocess(const float *x, float *out4, long n, float mean0, float mean1, float mean2, void *stream)
{
    SKY_REQUIRE(x && out4 && n > 0 && ((uintptr_t)out4 & 15) == 0, SKY_ERR_INVALID, "bad arguments");
    vgg_preprocess_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(x, out4, n, mean0, mean1, mean2);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

// ---- Radiance RGBE (inference.py:156 / utils.py:83-84: cv2.imwrite("*.hdr")) -----------------------------------------------------
// Greg Ward's shared-exponent encoding of one pixel: v = max(r, g, b) = m * 2^e with m in [0.5, 1) -> bytes (r, g, b) * 256 m / v, e + 128.
// Done on the device so the D2H copy of a prediction is 4 bytes per pixel instead of 12.
__global__ void rgbe_encode_kernel(const float *__restrict__ rgb, uint8_t *__restrict__ out, long npix, int bgr)
{
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
        const float c0 = rgb[3 * p], c1 = rgb[3 * p + 1], c2 = rgb[3 * p + 2];
        const float r = bgr ? c2 : c0, g = c1, b = bgr ? c0 : c2;
        const float v = fmaxf(r, fmaxf(g, b));
        uchar4 o = make_uchar4(0, 0, 0, 0);
        if (v >= 1e-32f) {
            int e;
            const float m = frexpf(v, &e);
            const float s = m * 256.f / v;
            o = make_uchar4((unsigned char)(int)(fmaxf(r, 0.f) * s), (unsigned char)(int)(fmaxf(g, 0.f) * s),
                            (unsigned char)(int)(fmaxf(b, 0.f) * s), (unsigned char)(e + 128));
        }
        reinterpret_cast<uchar4 *>(out)[p] = o;
    }
}

extern "C" int sky_rgbe_encode(const float *rgb, uint8_t *rgbe, long npix, int bgr, void *stream)
{
    SKY_REQUIRE(rgb && rgbe && npix > 0 && ((uintptr_t)rgbe & 3) == 0, SKY_ERR_INVALID, "bad arguments");
    rgbe_encode_kernel<<<blocks_for(npix), 256, 0, (cudaStream_t)stream>>>(rgb, rgbe, npix, bgr);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
