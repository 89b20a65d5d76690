// Sun branch of generator inference (inference.py:88-110): everything between the sun-position softmax and the sun
// decoder that is not a convolution.
//
//   softmax_max_bwd      d(max_j softmax_j)/d(logits) with the ReLU mask of sunpose_net.py:68 — the seed of Grad-CAM
//                        (inference.py:98: y_c = reduce_max(sunpose_cmf); grad_cam.py:31: tf.gradients(y_c, A_k))
//   transpose / dense_bwd_data   dX = dY . W^T of a Keras Dense, as the weight-streaming forward kernel over a cached W^T
//   maxpool2x2_bwd       tf.nn.max_pool gradient (MaxPoolGrad: the first maximum in scan order receives the gradient)
//   gradcam              grad_cam.layer (grad_cam.py:29-45): channel weights = spatial mean of the gradient, then
//                        relu(sum_c w_c A_c)
//   sunrad_input         generator.sun_rad_estimation (generator.py:160-164): resize CAM2 / CAM3, concat with the LDR image
//   bn_fold              Keras BatchNormalization in inference mode folded into the bias-free conv in front of it
//                        (sunrad_net.py:11-26)
//   sun_radiance         sunRadNet.call tail (sunrad_net.py:56-71) + hdr_logCompression (tf_utils.py:263-271), tiled x3
// All HBM/latency-bound elementwise or reduction kernels; fp32 throughout like the reference.
#include "sky_common.cuh"

namespace sky {

// ---- Grad-CAM seed ------------------------------------------------------------------------------------------------
// One CTA per row.  y_c = max_i sm_i; TF's reduce_max gradient spreads 1 evenly over ties; the softmax backward gives
// g_a[i] = sm_i * (ind_i / cnt - sum_j ind_j sm_j / cnt) = sm_i * (ind_i / cnt - y_c); the ReLU in front of the softmax
// (sunpose_net.py:68) passes it where its output is positive (ReluGrad tests the output).
// With `pick` (one class index per row: tf.gather_nd at argmax(sunpose_gt), train.py:263-265) the score is sm[pick] instead of the
// maximum: g_a[i] = sm_i * ((i == pick) - sm[pick]).
__global__ void __launch_bounds__(256) softmax_max_bwd_kernel(const float *__restrict__ sm, const float *__restrict__ act,
                                                              const int *__restrict__ pick, float *__restrict__ yc,
                                                              float *__restrict__ g, int N)
{
    __shared__ float redf[8];
    __shared__ int redi[8];
    const float *row = sm + (size_t)blockIdx.x * N;
    const float *arow = act + (size_t)blockIdx.x * N;
    float *out = g + (size_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (pick) {
        const int j = pick[blockIdx.x];
        const float sj = row[j];
        if (threadIdx.x == 0 && yc) yc[blockIdx.x] = sj;
        for (int i = threadIdx.x; i < N; i += 256) {
            const float ga = row[i] * ((i == j ? 1.f : 0.f) - sj);
            out[i] = arow[i] > 0.f ? ga : 0.f;
        }
        return;
    }
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += 256) m = fmaxf(m, row[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) redf[warp] = m;
    __syncthreads();
    m = redf[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, redf[i]);
    int cnt = 0;
    for (int i = threadIdx.x; i < N; i += 256) cnt += row[i] == m;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) redi[warp] = cnt;
    __syncthreads();
    cnt = 0;
    for (int i = 0; i < 8; ++i) cnt += redi[i];
    const float share = 1.f / (float)cnt;
    if (threadIdx.x == 0 && yc) yc[blockIdx.x] = m;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float s = row[i];
        const float ga = s * ((s == m ? share : 0.f) - m);
        out[i] = arow[i] > 0.f ? ga : 0.f;
    }
}

// ---- Dense backward (data) -----------------------------------------------------------------------------------------
// out [N, K] = in [K, N]^T, 32x32 tiles through shared memory.
__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int K, int N)
{
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8)
        if (k0 + r < K && n0 + threadIdx.x < N) tile[r][threadIdx.x] = in[(size_t)(k0 + r) * N + n0 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8)
        if (n0 + r < N && k0 + threadIdx.x < K) out[(size_t)(n0 + r) * K + k0 + threadIdx.x] = tile[threadIdx.x][r];
}

// dx *= (act > 0)   (ReluGrad of the activation that produced the Dense input; act may be NULL)
__global__ void relu_mask_kernel(float *__restrict__ dx, const float *__restrict__ act, long total)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
        if (!(act[e] > 0.f)) dx[e] = 0.f;
}

// ---- max-pool backward ---------------------------------------------------------------------------------------------
// 2x2 / 2 SAME windows do not overlap: each input element belongs to exactly one window, so dx is written once, without
// atomics or a memset.  The gradient goes to the first maximum in (row, column) scan order (TensorFlow's MaxPoolGrad on
// both devices selects it with a strict '>' scan).
__global__ void maxpool2x2_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dx, int B,
                                      int h, int w, int C, int oh, int ow)
{
    const int cv = C / 4;
    const long total = (long)B * oh * ow * cv;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % cv) * 4;
        const int ox = (int)((o / cv) % ow), oy = (int)((o / ((long)cv * ow)) % oh), b = (int)(o / ((long)cv * ow * oh));
        const float4 g = __ldg(reinterpret_cast<const float4 *>(dy + (((size_t)b * oh + oy) * ow + ox) * C + c));
        float4 v[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int yy = 2 * oy + (q >> 1), xx = 2 * ox + (q & 1);
            ok[q] = yy < h && xx < w;
            v[q] = ok[q] ? __ldg(reinterpret_cast<const float4 *>(x + (((size_t)b * h + yy) * w + xx) * C + c))
                         : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        int ax = 0, ay = 0, az = 0, aw = 0;
        float mx = v[0].x, my = v[0].y, mz = v[0].z, mw = v[0].w;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            if (v[q].x > mx) { mx = v[q].x; ax = q; }
            if (v[q].y > my) { my = v[q].y; ay = q; }
            if (v[q].z > mz) { mz = v[q].z; az = q; }
            if (v[q].w > mw) { mw = v[q].w; aw = q; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!ok[q]) continue;
            const int yy = 2 * oy + (q >> 1), xx = 2 * ox + (q & 1);
            *reinterpret_cast<float4 *>(dx + (((size_t)b * h + yy) * w + xx) * C + c) =
                make_float4(ax == q ? g.x : 0.f, ay == q ? g.y : 0.f, az == q ? g.z : 0.f, aw == q ? g.w : 0.f);
        }
    }
}

// ---- Grad-CAM ------------------------------------------------------------------------------------------------------
// wsum[b][c] += sum over this CTA's pixels of grad[b, p, c]   (wsum zeroed by the caller); grid (chunks, B)
__global__ void __launch_bounds__(256) gradcam_reduce_kernel(const float *__restrict__ grad, float *__restrict__ wsum, int hw,
                                                             int C, int pix_per_cta)
{
    extern __shared__ float part[];   // [256][4]
    const int b = blockIdx.y, c4n = C / 4;
    const int c4 = threadIdx.x % c4n, prow = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * pix_per_cta, npix = min(pix_per_cta, hw - p0);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = prow; p < npix; p += pstep) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(grad + ((size_t)b * hw + p0 + p) * C) + c4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4 *>(part)[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < C) {
        const int cc4 = threadIdx.x >> 2, u = threadIdx.x & 3;
        float s = 0.f;
        for (int r = 0; r < pstep; ++r) s += part[(r * c4n + cc4) * 4 + u];
        atomicAdd(wsum + (size_t)b * C + threadIdx.x, s);
    }
}

// cam[b, p] = relu(sum_c (wsum[b][c] / hw) * A[b, p, c]); C / 4 lanes cooperate on one pixel
__global__ void __launch_bounds__(256) gradcam_apply_kernel(const float *__restrict__ A, const float *__restrict__ wsum,
                                                            float *__restrict__ cam, int B, int hw, int C)
{
    const int lpp = C / 4;                               // lanes per pixel: 8, 16 or 32
    const long gthread = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const long pix = gthread / lpp;
    const int c4 = (int)(gthread % lpp);
    const bool ok = pix < (long)B * hw;
    float s = 0.f;
    if (ok) {
        const int b = (int)(pix / hw);
        const float4 a = __ldg(reinterpret_cast<const float4 *>(A + (size_t)pix * C) + c4);
        const float4 wv = __ldg(reinterpret_cast<const float4 *>(wsum + (size_t)b * C) + c4);
        const float inv = 1.f / (float)hw;
        s = a.x * (wv.x * inv) + a.y * (wv.y * inv) + a.z * (wv.z * inv) + a.w * (wv.w * inv);
    }
    for (int o = lpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (ok && c4 == 0) cam[pix] = fmaxf(s, 0.f);
}

// ---- sunRadNet input -------------------------------------------------------------------------------------------------
// One sample of tf.image.resize(BILINEAR, half-pixel centres) of a single-channel map — the arithmetic of
// resize_bilinear_kernel (da_conv_fwd.cu), one explicitly rounded fp32 op per TensorFlow op.
__device__ __forceinline__ float resize_tap(const float *__restrict__ img, int h, int w, int oy, int ox, float sy, float sx)
{
    const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), sy), 0.5f);
    const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), sx), 0.5f);
    const float fly = floorf(fy), flx = floorf(fx);
    const int ylo = max((int)fly, 0), yhi = min((int)ceilf(fy), h - 1);
    const int xlo = max((int)flx, 0), xhi = min((int)ceilf(fx), w - 1);
    const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
    const float tl = __ldg(img + ylo * w + xlo), tr = __ldg(img + ylo * w + xhi);
    const float bl = __ldg(img + yhi * w + xlo), br = __ldg(img + yhi * w + xhi);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
}

// out[b, y, x, 0:3] = ldr, [3] = cam1, [4] = resize(cam2), [5] = resize(cam3), [6:Cp] = 0 (channel padding so that the
// 4x4 / 2 conv in front of sunRadNet gathers whole 16-byte chunks; the padded kernel rows are zero)
__global__ void sunrad_input_kernel(const float *__restrict__ ldr, const float *__restrict__ cam1, const float *__restrict__ cam2,
                                    const float *__restrict__ cam3, float *__restrict__ out, int B, int H, int W, int h2, int w2,
                                    int h3, int w3, int Cp)
{
    const float sy2 = __fdiv_rn((float)h2, (float)H), sx2 = __fdiv_rn((float)w2, (float)W);
    const float sy3 = __fdiv_rn((float)h3, (float)H), sx3 = __fdiv_rn((float)w3, (float)W);
    const long total = (long)B * H * W;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int x = (int)(o % W), y = (int)((o / W) % H), b = (int)(o / ((long)W * H));
        const float *l = ldr + o * 3;
        const float c2 = resize_tap(cam2 + (size_t)b * h2 * w2, h2, w2, y, x, sy2, sx2);
        const float c3 = resize_tap(cam3 + (size_t)b * h3 * w3, h3, w3, y, x, sy3, sx3);
        float *dst = out + o * Cp;
        dst[0] = __ldg(l); dst[1] = __ldg(l + 1); dst[2] = __ldg(l + 2); dst[3] = __ldg(cam1 + o);
        dst[4] = c2; dst[5] = c3;
        for (int c = 6; c < Cp; ++c) dst[c] = 0.f;
    }
}

// ---- BatchNormalization (inference) folded into the preceding bias-free conv ------------------------------------------
// y = (conv(x) - mean) * gamma * rsqrt(var + eps) + beta  ==  conv'(x) + bias' with kernel'[:, f] = kernel[:, f] * s_f,
// bias'_f = beta_f - mean_f * s_f, s_f = gamma_f * rsqrt(var_f + eps)
__global__ void bn_fold_kernel(const float *__restrict__ kernel, const float *__restrict__ gamma, const float *__restrict__ beta,
                               const float *__restrict__ mean, const float *__restrict__ var, float eps,
                               float *__restrict__ kernel_out, float *__restrict__ bias_out, long K, int F)
{
    const long total = K * F;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int f = (int)(e % F);
        const float s = gamma[f] * rsqrtf(var[f] + eps);
        kernel_out[e] = kernel[e] * s;
        if (e < F) bias_out[f] = beta[f] - mean[f] * s;
    }
}

// ---- sun radiance ----------------------------------------------------------------------------------------------------
// sunRadNet.call (sunrad_net.py:56-71) on x = sunpose_pred / reduce_max(sunpose_pred) (generator.py:160), then
// hdr_logCompression (inference.py:105) and tf.tile(., 3) (generator.py:167):
//   v = exp(-(1-x)^2 / (beta+eps)) * gamma / (beta*sqrt(pi) + eps), clamped at 30000;  out = log(1 + 10 v) / log(11)
// gb [B][2]: the raw Dense(1) outputs (gamma, beta) before the sigmoid; gmax: the global maximum of sm (device scalar).
__global__ void sun_radiance_kernel(const float *__restrict__ sm, const float *__restrict__ gmax, const float *__restrict__ gb,
                                    float *__restrict__ out3, float *__restrict__ lin1, int B, int hw, float eps, float sqrt_pi)
{
    const long total = (long)B * hw;
    const float mx = __ldg(gmax);
    const float log11 = logf(11.f);
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int b = (int)(e / hw);
        const float gamma = 1.f / (1.f + expf(-__ldg(gb + 2 * b))), beta = 1.f / (1.f + expf(-__ldg(gb + 2 * b + 1)));
        const float x = __fdiv_rn(__ldg(sm + e), mx);
        const float d = __fsub_rn(1.f, x);
        float v = -__fmul_rn(d, d);
        v = __fdiv_rn(v, __fadd_rn(beta, eps));
        v = expf(v);
        v = __fmul_rn(v, gamma);
        v = __fdiv_rn(v, __fadd_rn(__fmul_rn(beta, sqrt_pi), eps));
        v = v > 30000.f ? 30000.f : v;
        if (lin1) lin1[e] = v;
        const float c = __fdiv_rn(logf(__fadd_rn(1.f, __fmul_rn(10.f, v))), log11);
        out3[3 * e] = c; out3[3 * e + 1] = c; out3[3 * e + 2] = c;
    }
}

// max over a whole tensor of non-negative floats (softmax outputs): the integer order of their bit patterns is the float order
__global__ void max_nonneg_kernel(const float *__restrict__ x, float *__restrict__ out, long total)
{
    float m = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) m = fmaxf(m, x[e]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));
}

static int ew_blocks(long total, int per = 256, int cap = 148 * 8)
{
    long b = (total + per - 1) / per;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace sky

using namespace sky;

extern "C" int sky_softmax_max_bwd(const float *sm, const float *act, float *yc, float *g, int rows, int N, void *stream)
{
    SKY_REQUIRE(sm && act && g && rows > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    softmax_max_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(sm, act, nullptr, yc, g, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_transpose(const float *in, float *out, int K, int N, void *stream)
{
    SKY_REQUIRE(in && out && K > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    transpose_kernel<<<dim3((N + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, K, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dense_bwd_data(const float *dy, const float *Wt, const float *act, float *dx, int B, int K, int N, void *stream)
{
    SKY_REQUIRE(dy && Wt && dx && B > 0 && K > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    // dx [B, K] = dy [B, N] . Wt [N, K]: the forward weight-streaming kernel with the roles of K and N exchanged
    int rc = sky_dense_fwd(dy, Wt, nullptr, dx, B, N, K, 0, stream);
    if (rc != SKY_OK) return rc;
    if (act) {
        const long total = (long)B * K;
        relu_mask_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(dx, act, total);
        SKY_CHECK_LAUNCH();
    }
    return SKY_OK;
}

extern "C" int sky_maxpool2x2_bwd(const float *x, const float *dy, float *dx, int B, int h, int w, int C, void *stream)
{
    SKY_REQUIRE(x && dy && dx && B > 0 && h > 0 && w > 0 && C > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(C % 4 == 0, SKY_ERR_UNSUPPORTED, "max-pool kernel needs C %% 4 == 0 (got %d)", C);
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const long total = (long)B * oh * ow * (C / 4);
    maxpool2x2_bwd_kernel<<<ew_blocks(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, B, h, w, C, oh, ow);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_gradcam(const float *grad, const float *A, float *wsum, float *cam, int B, int h, int w, int C, void *stream)
{
    SKY_REQUIRE(grad && A && wsum && cam && B > 0 && h > 0 && w > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(C == 32 || C == 64 || C == 128, SKY_ERR_UNSUPPORTED, "Grad-CAM kernel covers the 32/64/128-channel sun-position maps (got %d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w;
    SKY_CHECK_CUDA(cudaMemsetAsync(wsum, 0, (size_t)B * C * sizeof(float), st));
    int chunks = (2 * 148 + B - 1) / B;
    int pix = (hw + chunks - 1) / chunks;
    if (pix < 16) pix = 16;
    chunks = (hw + pix - 1) / pix;
    gradcam_reduce_kernel<<<dim3(chunks, B), 256, 256 * 4 * sizeof(float), st>>>(grad, wsum, hw, C, pix);
    SKY_CHECK_LAUNCH();
    const long threads = (long)B * hw * (C / 4);
    gradcam_apply_kernel<<<(int)((threads + 255) / 256), 256, 0, st>>>(A, wsum, cam, B, hw, C);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_sunrad_input(const float *ldr, const float *cam1, const float *cam2, const float *cam3, float *out, int B,
                                int H, int W, int h2, int w2, int h3, int w3, int Cp, void *stream)
{
    SKY_REQUIRE(ldr && cam1 && cam2 && cam3 && out && B > 0 && H > 0 && W > 0 && h2 > 0 && w2 > 0 && h3 > 0 && w3 > 0, SKY_ERR_INVALID,
                "bad arguments");
    SKY_REQUIRE(Cp >= 6, SKY_ERR_INVALID, "the sunRadNet input has 6 channels (got Cp=%d)", Cp);
    sunrad_input_kernel<<<ew_blocks((long)B * H * W), 256, 0, (cudaStream_t)stream>>>(ldr, cam1, cam2, cam3, out, B, H, W, h2, w2, h3, w3, Cp);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_bn_fold(const float *kernel, const float *gamma, const float *beta, const float *mean, const float *var,
                           float eps, float *kernel_out, float *bias_out, long K, int F, void *stream)
{
    SKY_REQUIRE(kernel && gamma && beta && mean && var && kernel_out && bias_out && K > 0 && F > 0, SKY_ERR_INVALID, "bad arguments");
    bn_fold_kernel<<<ew_blocks(K * F), 256, 0, (cudaStream_t)stream>>>(kernel, gamma, beta, mean, var, eps, kernel_out, bias_out, K, F);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_max_nonneg(const float *x, float *out, long n, void *stream)
{
    SKY_REQUIRE(x && out && n > 0, SKY_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    max_nonneg_kernel<<<ew_blocks(n, 256, 148 * 4), 256, 0, st>>>(x, out, n);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_sun_radiance(const float *sm, const float *gmax, const float *gb, float *out3, float *lin1, int B, int hw,
                                float eps, void *stream)
{
    SKY_REQUIRE(sm && gmax && gb && out3 && B > 0 && hw > 0, SKY_ERR_INVALID, "bad arguments");
    // deltafunc_const = tf.sqrt(pi) (sunrad_net.py:35): fp32 sqrt of fp32(pi)
    const float sqrt_pi = sqrtf(3.14159265358979323846f);
    sun_radiance_kernel<<<ew_blocks((long)B * hw), 256, 0, (cudaStream_t)stream>>>(sm, gmax, gb, out3, lin1, B, hw, eps, sqrt_pi);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

// first index of the row maximum (tf.math.argmax)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float *__restrict__ x, int *__restrict__ idx, int N)
{
    __shared__ float vals[256];
    __shared__ int ids[256];
    const float *row = x + (size_t)blockIdx.x * N;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int i = threadIdx.x; i < N; i += 256)
        if (row[i] > m) { m = row[i]; mi = i; }
    vals[threadIdx.x] = m; ids[threadIdx.x] = mi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float o = vals[threadIdx.x + s];
            const int oi = ids[threadIdx.x + s];
            if (o > vals[threadIdx.x] || (o == vals[threadIdx.x] && oi < ids[threadIdx.x])) { vals[threadIdx.x] = o; ids[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) idx[blockIdx.x] = ids[0];
}

// The un-fused tail of train.generator_in_step (train.py:256-259, 289-298): alpha from the sky prediction, the two scaled branches,
// their sum and the three decompressed images.  One pixel (3 channels) per thread.
__global__ void blend_split_kernel(const float *__restrict__ sky_gamma, const float *__restrict__ sun_gamma, float thr,
                                   float *__restrict__ y_gamma, float *__restrict__ y_lin, float *__restrict__ sky_lin,
                                   float *__restrict__ sun_lin, float *__restrict__ alpha_out, long npix)
{
    const float l11 = 2.3978953f;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
        const float *s = sky_gamma + 3 * p, *u = sun_gamma + 3 * p;
        const float s0 = s[0], s1 = s[1], s2 = s[2];
        const float gmax = fmaxf(fmaxf(s0, s1), s2);
        const float lin = __fdiv_rn(__fsub_rn(expf(__fmul_rn(gmax, l11)), 1.f), 10.f);
        const float a = fminf(1.f, __fdiv_rn(fmaxf(0.f, __fadd_rn(__fsub_rn(lin, 1.f), thr)), thr));
        const float na = __fsub_rn(1.f, a);
        if (alpha_out) alpha_out[p] = a;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float sk = __fmul_rn(na, s[c]), sn = __fmul_rn(a, u[c]);
            const float yg = __fadd_rn(sk, sn);
            y_gamma[3 * p + c] = yg;
            y_lin[3 * p + c] = __fdiv_rn(__fsub_rn(expf(__fmul_rn(yg, l11)), 1.f), 10.f);
            if (sky_lin) sky_lin[3 * p + c] = __fdiv_rn(__fsub_rn(expf(__fmul_rn(sk, l11)), 1.f), 10.f);
            if (sun_lin) sun_lin[3 * p + c] = __fdiv_rn(__fsub_rn(expf(__fmul_rn(sn, l11)), 1.f), 10.f);
        }
    }
}

extern "C" int sky_softmax_pick_bwd(const float *sm, const float *act, const int *pick, float *yc, float *g, int rows, int N, void *stream)
{
    SKY_REQUIRE(sm && act && pick && g && rows > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    sky::softmax_max_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(sm, act, pick, yc, g, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_argmax_rows(const float *x, int *idx, int rows, int N, void *stream)
{
    SKY_REQUIRE(x && idx && rows > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    argmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, idx, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_blend_split(const float *sky_gamma, const float *sun_gamma, float threshold, float *y_gamma, float *y_lin,
                               float *sky_lin, float *sun_lin, float *alpha, long npix, void *stream)
{
    SKY_REQUIRE(sky_gamma && sun_gamma && y_gamma && y_lin && npix > 0 && threshold > 0.f, SKY_ERR_INVALID, "bad arguments");
    long blocks = (npix + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    blend_split_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(sky_gamma, sun_gamma, threshold, y_gamma, y_lin, sky_lin, sun_lin, alpha,
                                                                      npix);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
