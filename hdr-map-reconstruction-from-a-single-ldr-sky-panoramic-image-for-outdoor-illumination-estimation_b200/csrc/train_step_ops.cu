// Elementwise / reduction kernels of the full GAN train step (train.train_step, train.py:382-415) that neither the inference path
// nor the sun pre-train step needed:
//   Keras BatchNormalization with batch statistics, forward and backward (sunrad_net.py:17,25; discriminator.py:16,24);
//   adjoint of tf.image.resize (the decoders' resize-deconv, ops.py:122);
//   adjoint of the sun-radiance head (sunrad_net.py:56-71 with the max-normalisation of generator.py:160, hdr_logCompression and the
//     x3 tile of generator.py:167) and of the two Dense(1) layers in front of it;
//   the tail of train.generator_in_step (train.py:256-259, 289-298) in one pass forward and one backward, with the L1 term of :324;
//   adjoints of the LSGAN, L1 / perceptual terms; max-pool gradient fused with the ReLU mask of the layer below (VGG16).
// All HBM-bound: each distinct tensor is read or written once per kernel.
#include "sky_common.cuh"

namespace sky {

constexpr int TS_THREADS = 256;

static int ts_blocks(long total, int per = TS_THREADS, int cap = 148 * 8)
{
    long b = (total + per - 1) / per;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// batch normalisation, training mode
// ---------------------------------------------------------------------------------------------------------------------
// Column sums of a [rows x F] matrix per batch group: sums[g][f][2] += (sum x, sum x^2).  grid (F/32 column slabs, row chunks, groups);
// block 32 x 8: lane = column, 8 row lanes.
__global__ void __launch_bounds__(TS_THREADS) bn_stats_kernel(const float *__restrict__ x, double *__restrict__ sums, int rows_per_group,
                                                                int F, int rows_per_cta)
{
    __shared__ float red[8][32][2];
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + lane, g = blockIdx.z;
    const int lo = blockIdx.y * rows_per_cta, hi = min(rows_per_group, lo + rows_per_cta);
    float s1 = 0.f, s2 = 0.f;
    if (f < F)
        for (int r = lo + ry; r < hi; r += 8) {
            const float v = __ldg(x + ((size_t)g * rows_per_group + r) * F + f);
            s1 += v; s2 = fmaf(v, v, s2);
        }
    red[ry][lane][0] = s1; red[ry][lane][1] = s2;
    __syncthreads();
    if (ry < 2 && f < F) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += (double)red[q][lane][ry];
        atomicAdd(sums + ((size_t)g * F + f) * 2 + ry, s);
    }
}

// mean_var[g][f] = (mean, biased variance); moving statistics (if given) take one momentum step per group, in group order, with the
// Bessel-corrected variance (TensorFlow FusedBatchNorm's running-average output).
__global__ void bn_finalize_kernel(const double *__restrict__ sums, float *__restrict__ mean_var, float *__restrict__ moving_mean,
                                   float *__restrict__ moving_var, int F, int groups, int n, float momentum)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    for (int g = 0; g < groups; ++g) {
        const double s1 = sums[((size_t)g * F + f) * 2], s2 = sums[((size_t)g * F + f) * 2 + 1];
        const double mean = s1 / n;
        double var = s2 / n - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        mean_var[((size_t)g * F + f) * 2] = (float)mean;
        mean_var[((size_t)g * F + f) * 2 + 1] = (float)var;
        if (moving_mean) {
            const double unbiased = var * ((double)n / (double)(n > 1 ? n - 1 : 1));
            moving_mean[f] = moving_mean[f] * momentum + (float)mean * (1.f - momentum);
            moving_var[f] = moving_var[f] * momentum + (float)unbiased * (1.f - momentum);
        }
    }
}

// y = lrelu(x * inv + (beta - mean * inv)), inv = rsqrt(var + eps) * gamma  (tf.nn.batch_normalization's association)
__global__ void __launch_bounds__(TS_THREADS) bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ mean_var,
                                                               const float *__restrict__ gamma, const float *__restrict__ beta,
                                                               float *__restrict__ y, long rows_per_group, int F, int groups, float eps,
                                                               int flags, float slope)
{
    const int f4n = F / 4;
    const long total4 = rows_per_group * groups * f4n;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4; e += (long)gridDim.x * blockDim.x) {
        const int f = (int)(e % f4n) * 4;
        const int g = (int)((e / f4n) / rows_per_group);
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + e);
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w };
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 mv = __ldg(reinterpret_cast<const float2 *>(mean_var) + (size_t)g * F + f + u);
            const float inv = rsqrtf(mv.y + eps) * __ldg(gamma + f + u);
            float v = fmaf(xs[u], inv, __ldg(beta + f + u) - mv.x * inv);
            if (flags & SKY_EPI_LEAKY_RELU) v = v > 0.f ? v : v * slope;
            o[u] = v;
        }
        reinterpret_cast<float4 *>(y)[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// sums[g][f][2] += (sum dy', sum dy' * xhat), dy' = dy * lrelu'(act)
__global__ void __launch_bounds__(TS_THREADS) bn_bwd_reduce_kernel(const float *__restrict__ x, const float *__restrict__ mean_var,
                                                                    const float *__restrict__ dy, const float *__restrict__ act,
                                                                    double *__restrict__ sums, int rows_per_group, int F, int rows_per_cta,
                                                                    float eps, float slope)
{
    __shared__ float red[8][32][2];
    const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + lane, g = blockIdx.z;
    const int lo = blockIdx.y * rows_per_cta, hi = min(rows_per_group, lo + rows_per_cta);
    float s1 = 0.f, s2 = 0.f;
    if (f < F) {
        const float2 mv = __ldg(reinterpret_cast<const float2 *>(mean_var) + (size_t)g * F + f);
        const float rstd = rsqrtf(mv.y + eps);
        for (int r = lo + ry; r < hi; r += 8) {
            const size_t o = ((size_t)g * rows_per_group + r) * F + f;
            float gdy = __ldg(dy + o);
            if (act) gdy *= __ldg(act + o) > 0.f ? 1.f : slope;
            s1 += gdy;
            s2 = fmaf(gdy, (__ldg(x + o) - mv.x) * rstd, s2);
        }
    }
    red[ry][lane][0] = s1; red[ry][lane][1] = s2;
    __syncthreads();
    if (ry < 2 && f < F) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += (double)red[q][lane][ry];
        atomicAdd(sums + ((size_t)g * F + f) * 2 + ry, s);
    }
}

// dx = gamma * rstd * (dy' - mean(dy') - xhat * mean(dy' * xhat));  block 0 adds dgamma[f] += sum_g sum(dy' xhat), dbeta[f] += sum_g sum(dy')
__global__ void __launch_bounds__(TS_THREADS) bn_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ mean_var,
                                                                   const float *__restrict__ gamma, const float *__restrict__ dy,
                                                                   const float *__restrict__ act, const double *__restrict__ sums,
                                                                   float *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                                   long rows_per_group, int F, int groups, float eps, float slope)
{
    if (blockIdx.x == 0)
        for (int f = threadIdx.x; f < F; f += TS_THREADS) {
            double a1 = 0.0, a2 = 0.0;
            for (int g = 0; g < groups; ++g) { a1 += sums[((size_t)g * F + f) * 2]; a2 += sums[((size_t)g * F + f) * 2 + 1]; }
            dbeta[f] += (float)a1;
            dgamma[f] += (float)a2;
        }
    const int f4n = F / 4;
    const long total4 = rows_per_group * groups * f4n;
    const float inv_n = 1.f / (float)rows_per_group;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4; e += (long)gridDim.x * blockDim.x) {
        const int f = (int)(e % f4n) * 4;
        const int g = (int)((e / f4n) / rows_per_group);
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x) + e);
        float4 gv = __ldg(reinterpret_cast<const float4 *>(dy) + e);
        if (act) {
            const float4 av = __ldg(reinterpret_cast<const float4 *>(act) + e);
            gv.x *= av.x > 0.f ? 1.f : slope; gv.y *= av.y > 0.f ? 1.f : slope;
            gv.z *= av.z > 0.f ? 1.f : slope; gv.w *= av.w > 0.f ? 1.f : slope;
        }
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w }, gs[4] = { gv.x, gv.y, gv.z, gv.w };
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 mv = __ldg(reinterpret_cast<const float2 *>(mean_var) + (size_t)g * F + f + u);
            const float rstd = rsqrtf(mv.y + eps);
            const float m1 = (float)(sums[((size_t)g * F + f + u) * 2] * (double)inv_n);
            const float m2 = (float)(sums[((size_t)g * F + f + u) * 2 + 1] * (double)inv_n);
            const float xh = (xs[u] - mv.x) * rstd;
            o[u] = __ldg(gamma + f + u) * rstd * (gs[u] - m1 - xh * m2);
        }
        reinterpret_cast<float4 *>(dx)[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// adjoint of tf.image.resize(BILINEAR, half-pixel centres): every dy element is scattered to its four source pixels with the
// forward pass's own index / lerp arithmetic (resize_bilinear_kernel in da_conv_fwd.cu)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void resize_bilinear_bwd_kernel(const float *__restrict__ dy, float *__restrict__ dx, int B, int h, int w, int C, int oh, int ow)
{
    const float sy = __fdiv_rn((float)h, (float)oh), sx = __fdiv_rn((float)w, (float)ow);
    const int cv = C / 4;
    const long total = (long)B * oh * ow * cv;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % cv) * 4;
        const int ox = (int)((o / cv) % ow), oy = (int)((o / ((long)cv * ow)) % oh), b = (int)(o / ((long)cv * ow * oh));
        const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), sy), 0.5f);
        const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), sx), 0.5f);
        const float fly = floorf(fy), flx = floorf(fx);
        const int ylo = max((int)fly, 0), yhi = min((int)ceilf(fy), h - 1);
        const int xlo = max((int)flx, 0), xhi = min((int)ceilf(fx), w - 1);
        const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
        const float4 g = __ldg(reinterpret_cast<const float4 *>(dy) + o);
        // out = top + (bot - top) * ly, top = tl + (tr - tl) * lx  =>  weights (1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly lx
        const float wy[2] = { 1.f - ly, ly }, wx[2] = { 1.f - lx, lx };
        const int ys[2] = { ylo, yhi }, xs[2] = { xlo, xhi };
        float *img = dx + (size_t)b * h * w * C + c;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const float wgt = wy[a] * wx[d];
                if (wgt == 0.f) continue;
                float *dst = img + ((size_t)ys[a] * w + xs[d]) * C;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(wgt * g.x), "f"(wgt * g.y), "f"(wgt * g.z),
                             "f"(wgt * g.w)
                             : "memory");
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// sun-radiance head, backward
// ---------------------------------------------------------------------------------------------------------------------
// Forward (sun_radiance_kernel, sun_ops.cu): x = sm / max; gam = sigmoid(g); bet = sigmoid(b); e = exp(-(1-x)^2 / (bet+eps));
// v = e * gam / (bet*sqrt(pi) + eps), clamped at 30000; out_c = log(1 + 10 v) / log 11, c = 0..2.
// d_out3 [B,hw,3] -> dnorm [B,hw] = dL/dx; dgb64 [B][2] += dL/d(g, b) (pre-sigmoid); red2 += (sum dnorm * x, number of elements equal to max).
__global__ void __launch_bounds__(TS_THREADS) sun_radiance_bwd_kernel(const float *__restrict__ sm, const float *__restrict__ gmax,
                                                                       const float *__restrict__ gb, const float *__restrict__ d_out3,
                                                                       float *__restrict__ dnorm, double *__restrict__ dgb64,
                                                                       double *__restrict__ red2, int hw, float eps, float sqrt_pi)
{
    __shared__ float red[TS_THREADS / 32][4];
    const int b = blockIdx.y;
    const float mx = __ldg(gmax);
    const float gam = 1.f / (1.f + expf(-__ldg(gb + 2 * b))), bet = 1.f / (1.f + expf(-__ldg(gb + 2 * b + 1)));
    const float den1 = bet + eps, den2 = bet * sqrt_pi + eps;
    const float log11 = logf(11.f);
    float a_g = 0.f, a_b = 0.f, a_dot = 0.f, a_cnt = 0.f;
    for (int i = blockIdx.x * TS_THREADS + threadIdx.x; i < hw; i += gridDim.x * TS_THREADS) {
        const size_t e = (size_t)b * hw + i;
        const float p = __ldg(sm + e);
        const float x = p / mx, d = 1.f - x;
        const float ex = expf(-(d * d) / den1);
        const float v = ex * gam / den2;
        const float vc = v > 30000.f ? 30000.f : v;
        const float dsum = __ldg(d_out3 + 3 * e) + __ldg(d_out3 + 3 * e + 1) + __ldg(d_out3 + 3 * e + 2);
        float dv = dsum * 10.f / ((1.f + 10.f * vc) * log11);
        if (v > 30000.f) dv = 0.f;                                        // tf.where(x > 30000., 30000., x): no gradient through the clamp
        const float du = dv * v;                                          // d/du of exp(u) * gam / den2
        const float dx = du * 2.f * d / den1;
        a_g += dv * ex / den2;                                            // dL/d gam
        a_b += du * (d * d) / (den1 * den1) - dv * v * sqrt_pi / den2;    // dL/d bet: through u and through the normaliser
        a_dot = fmaf(dx, x, a_dot);
        a_cnt += (p == mx) ? 1.f : 0.f;
        dnorm[e] = dx;
    }
    a_g = warp_sum(a_g); a_b = warp_sum(a_b); a_dot = warp_sum(a_dot); a_cnt = warp_sum(a_cnt);
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    if (lane == 0) { red[wq][0] = a_g; red[wq][1] = a_b; red[wq][2] = a_dot; red[wq][3] = a_cnt; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int q = 0; q < TS_THREADS / 32; ++q) s += (double)red[q][threadIdx.x];
        if (threadIdx.x == 0) atomicAdd(dgb64 + 2 * b, s * (double)(gam * (1.f - gam)));
        else if (threadIdx.x == 1) atomicAdd(dgb64 + 2 * b + 1, s * (double)(bet * (1.f - bet)));
        else atomicAdd(red2 + (threadIdx.x - 2), s);
    }
}

// x = p / max(p) over the whole batch (generator.py:160): dp_i (+)= dnorm_i / m - [p_i == m] * (sum_j dnorm_j x_j) / (m * ties)
__global__ void maxnorm_bwd_kernel(const float *__restrict__ sm, const float *__restrict__ gmax, const float *__restrict__ dnorm,
                                   const double *__restrict__ red2, float *__restrict__ dsm, long n, int accumulate)
{
    const float mx = __ldg(gmax);
    const float corr = (float)(red2[0] / ((double)mx * (red2[1] > 0.0 ? red2[1] : 1.0)));
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        float g = __ldg(dnorm + e) / mx;
        if (__ldg(sm + e) == mx) g -= corr;
        dsm[e] = accumulate ? dsm[e] + g : g;
    }
}

// Dense(1) x 2 (sunrad_net.py:43-44, merged [K,2] kernel): dflat[b,k] = sum_j dgb[b,j] W[k,j]; dW[k,j] = sum_b flat[b,k] dgb[b,j];
// dbias[j] = sum_b dgb[b,j].  One thread per k.
__global__ void sunrad_heads_bwd_kernel(const float *__restrict__ flat, const float *__restrict__ W, const double *__restrict__ dgb64,
                                        float *__restrict__ dW, float *__restrict__ dbias, float *__restrict__ dflat, int B, int K)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int b = 0; b < B; ++b) { s0 += dgb64[2 * b]; s1 += dgb64[2 * b + 1]; }
        dbias[0] = (float)s0; dbias[1] = (float)s1;
    }
    if (k >= K) return;
    const float w0 = __ldg(W + 2 * (size_t)k), w1 = __ldg(W + 2 * (size_t)k + 1);
    float a0 = 0.f, a1 = 0.f;
    for (int b = 0; b < B; ++b) {
        const float g0 = (float)dgb64[2 * b], g1 = (float)dgb64[2 * b + 1];
        const float fv = __ldg(flat + (size_t)b * K + k);
        a0 = fmaf(fv, g0, a0); a1 = fmaf(fv, g1, a1);
        dflat[(size_t)b * K + k] = fmaf(g0, w0, g1 * w1);
    }
    dW[2 * (size_t)k] = a0; dW[2 * (size_t)k + 1] = a1;
}

// ---------------------------------------------------------------------------------------------------------------------
// tail of train.generator_in_step
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float log_decompress(float v) { return __fdiv_rn(__fsub_rn(expf(__fmul_rn(v, 2.3978953f)), 1.f), 10.f); }

// c_sky / c_sun: conv1_f / conv1_u outputs (bias included).  generator.py:120-124, 150-156; train.py:251, 256-259, 289-298, 324.
__global__ void __launch_bounds__(TS_THREADS) train_tail_fwd_kernel(const float *__restrict__ c_sky, const float *__restrict__ c_sun,
                                                                     const float *__restrict__ ldr, const float *__restrict__ sun_rad_gamma,
                                                                     const float *__restrict__ hdr_t, float thr, float slope,
                                                                     float *__restrict__ y_gamma, float *__restrict__ y_lin,
                                                                     float *__restrict__ sky_lin, float *__restrict__ sun_lin,
                                                                     float *__restrict__ alpha, double *__restrict__ l1_acc, long npix)
{
    float acc = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < npix; e += (long)gridDim.x * blockDim.x) {
        float s[3], u[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s[c] = fmaxf(__fadd_rn(__ldg(ldr + 3 * e + c), lrelu(__ldg(c_sky + 3 * e + c), slope)), 0.f);
            u[c] = fmaxf(__fadd_rn(__ldg(sun_rad_gamma + 3 * e + c), lrelu(__ldg(c_sun + 3 * e + c), slope)), 0.f);
        }
        const float gmax = fmaxf(fmaxf(s[0], s[1]), s[2]);                // exp is monotonic: the max of the linear values
        const float a = fminf(1.f, __fdiv_rn(fmaxf(0.f, __fadd_rn(__fsub_rn(log_decompress(gmax), 1.f), thr)), thr));
        const float na = __fsub_rn(1.f, a);
        alpha[e] = a;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ss = __fmul_rn(na, s[c]), su = __fmul_rn(a, u[c]);
            const float yg = __fadd_rn(ss, su);
            const float yl = log_decompress(yg);
            y_gamma[3 * e + c] = yg;
            y_lin[3 * e + c] = yl;
            sky_lin[3 * e + c] = log_decompress(ss);
            sun_lin[3 * e + c] = log_decompress(su);
            acc += fabsf(yl - __ldg(hdr_t + 3 * e + c));
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(l1_acc, (double)acc);
}

// g_lin = g_dog + g_dis[3:6] + w_l1 sign(y_lin - hdr_t);  g_y = g_lin * log(11) (10 y_lin + 1) / 10 + w_vgg g_vgg;
// dc_sky = (1 - alpha) g_y [s > 0] lrelu'(c_sky);  dc_sun = alpha g_y [u > 0] lrelu'(c_sun);  d_sun_rad_gamma = alpha g_y [u > 0]
__global__ void __launch_bounds__(TS_THREADS) train_tail_bwd_kernel(const float *__restrict__ c_sky, const float *__restrict__ c_sun,
                                                                     const float *__restrict__ ldr, const float *__restrict__ sun_rad_gamma,
                                                                     const float *__restrict__ alpha, const float *__restrict__ y_lin,
                                                                     const float *__restrict__ hdr_t, const float *__restrict__ g_dog,
                                                                     const float *__restrict__ g_dis8, const float *__restrict__ g_vgg4,
                                                                     float w_l1, float w_vgg, float slope, float *__restrict__ dc_sky,
                                                                     float *__restrict__ dc_sun, float *__restrict__ d_srg, long npix)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < npix; e += (long)gridDim.x * blockDim.x) {
        const float a = __ldg(alpha + e);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float yl = __ldg(y_lin + 3 * e + c), dif = yl - __ldg(hdr_t + 3 * e + c);
            float gl = w_l1 * (dif > 0.f ? 1.f : (dif < 0.f ? -1.f : 0.f));
            if (g_dog) gl += __ldg(g_dog + 3 * e + c);
            if (g_dis8) gl += __ldg(g_dis8 + 8 * e + 3 + c);
            float gy = gl * (2.3978953f * (10.f * yl + 1.f) / 10.f);
            if (g_vgg4) gy = fmaf(w_vgg, __ldg(g_vgg4 + 4 * e + c), gy);
            const float cs = __ldg(c_sky + 3 * e + c), cu = __ldg(c_sun + 3 * e + c);
            const float s = __fadd_rn(__ldg(ldr + 3 * e + c), lrelu(cs, slope));
            const float u = __fadd_rn(__ldg(sun_rad_gamma + 3 * e + c), lrelu(cu, slope));
            const float gs = s > 0.f ? (1.f - a) * gy : 0.f;
            const float gu = u > 0.f ? a * gy : 0.f;
            dc_sky[3 * e + c] = gs * (cs > 0.f ? 1.f : slope);
            dc_sun[3 * e + c] = gu * (cu > 0.f ? 1.f : slope);
            d_srg[3 * e + c] = gu;
        }
    }
}

// LSGAN terms on the discriminator's VALID output, held as the SAME-padded map [B, hh, ww] (discriminator.py:48):
// inside the window rows r0..r1-1 / columns c0..c1-1: g = scale * 2 (d - target), acc += (d - target)^2; outside: g = 0.
__global__ void lsgan_bwd_kernel(const float *__restrict__ d, float *__restrict__ g, double *__restrict__ acc, int B, int hh, int ww, int r0,
                                 int r1, int c0, int c1, float target, float scale)
{
    float a = 0.f;
    const long total = (long)B * hh * ww;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int x = (int)(e % ww), y = (int)((e / ww) % hh);
        float gv = 0.f;
        if (y >= r0 && y < r1 && x >= c0 && x < c1) {
            const float df = __ldg(d + e) - target;
            gv = scale * 2.f * df;
            a = fmaf(df, df, a);
        }
        if (g) g[e] = gv;
    }
    a = warp_sum(a);
    if (acc && (threadIdx.x & 31) == 0) atomicAdd(acc, (double)a);
}

// g (+)= scale * sign(a - b); acc += sum |a - b|   (tf.reduce_mean(tf.abs(a - b)) and its adjoint, train.py:307-309)
__global__ void l1_bwd_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ g, double *__restrict__ acc, long n,
                              float scale, int accumulate)
{
    float s = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float d = __ldg(a + e) - __ldg(b + e);
        s += fabsf(d);
        const float gv = scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        if (g) g[e] = accumulate ? g[e] + gv : gv;
    }
    s = warp_sum(s);
    if (acc && (threadIdx.x & 31) == 0) atomicAdd(acc, (double)s);
}

// tf.nn.max_pool 2x2/2 SAME gradient (first maximum in scan order) fused with the ReLU gradient of the layer that produced x
// (x = relu(...): the routed gradient is dropped where x <= 0) and with an optional additive gradient for the same tensor.
__global__ void maxpool2x2_bwd_relu_kernel(const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ dx, int B, int h,
                                           int w, int C, int oh, int ow, int relu_mask)
{
    const int cv = C / 4;
    const long total = (long)B * oh * ow * cv;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % cv) * 4;
        const int ox = (int)((o / cv) % ow), oy = (int)((o / ((long)cv * ow)) % oh), b = (int)(o / ((long)cv * ow * oh));
        const float4 g = __ldg(reinterpret_cast<const float4 *>(dy + (((size_t)b * oh + oy) * ow + ox) * C + c));
        const float gs[4] = { g.x, g.y, g.z, g.w };
        float v[4][4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int yy = 2 * oy + (q >> 1), xx = 2 * ox + (q & 1);
            ok[q] = yy < h && xx < w;
            float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            if (ok[q]) t = __ldg(reinterpret_cast<const float4 *>(x + (((size_t)b * h + yy) * w + xx) * C + c));
            v[q][0] = t.x; v[q][1] = t.y; v[q][2] = t.z; v[q][3] = t.w;
        }
        int am[4];
        float mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            am[u] = 0; mv[u] = v[0][u];
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (v[q][u] > mv[u]) { mv[u] = v[q][u]; am[u] = q; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!ok[q]) continue;
            const int yy = 2 * oy + (q >> 1), xx = 2 * ox + (q & 1);
            float o4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) o4[u] = (am[u] == q && (!relu_mask || mv[u] > 0.f)) ? gs[u] : 0.f;
            *reinterpret_cast<float4 *>(dx + (((size_t)b * h + yy) * w + xx) * C + c) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

}  // namespace sky

using namespace sky;

extern "C" int sky_zero(void *p, size_t bytes, void *stream)
{
    SKY_REQUIRE(p != nullptr || bytes == 0, SKY_ERR_INVALID, "NULL pointer");
    if (bytes) SKY_CHECK_CUDA(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
    return SKY_OK;
}

extern "C" int sky_bn_train_stats(const float *x, double *sums, float *mean_var, float *moving_mean, float *moving_var, int B, int hw, int F,
                                  int groups, float momentum, void *stream)
{
    SKY_REQUIRE(x && sums && mean_var, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && hw > 0 && F > 0 && groups > 0 && B % groups == 0, SKY_ERR_INVALID, "bad dimension (the batch must split evenly into the groups)");
    SKY_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), SKY_ERR_INVALID, "moving mean and variance come together");
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = (B / groups) * hw;
    SKY_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)groups * F * 2 * sizeof(double), st));
    const int slabs = (F + 31) / 32;
    int chunks = (4 * 148) / (slabs * groups);
    if (chunks < 1) chunks = 1;
    int rows_per_cta = (rows + chunks - 1) / chunks;
    if (rows_per_cta < 64) rows_per_cta = 64;
    chunks = (rows + rows_per_cta - 1) / rows_per_cta;
    bn_stats_kernel<<<dim3(slabs, chunks, groups), TS_THREADS, 0, st>>>(x, sums, rows, F, rows_per_cta);
    SKY_CHECK_LAUNCH();
    bn_finalize_kernel<<<(F + 127) / 128, 128, 0, st>>>(sums, mean_var, moving_mean, moving_var, F, groups, rows, momentum);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_bn_train_apply(const float *x, const float *mean_var, const float *gamma, const float *beta, float *y, int B, int hw,
                                  int F, int groups, float eps, int epilogue_flags, float slope, void *stream)
{
    SKY_REQUIRE(x && mean_var && gamma && beta && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && hw > 0 && F > 0 && groups > 0 && B % groups == 0 && F % 4 == 0, SKY_ERR_INVALID, "bad dimension");
    const long rows = (long)(B / groups) * hw;
    bn_apply_kernel<<<ts_blocks(rows * groups * (F / 4)), TS_THREADS, 0, (cudaStream_t)stream>>>(x, mean_var, gamma, beta, y, rows, F, groups, eps,
                                                                                                  epilogue_flags, slope);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_bn_train_bwd(const float *x, const float *mean_var, const float *gamma, const float *dy, const float *act, double *sums,
                                float *dx, float *dgamma, float *dbeta, int B, int hw, int F, int groups, float eps, float slope, void *stream)
{
    SKY_REQUIRE(x && mean_var && gamma && dy && sums && dx && dgamma && dbeta, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && hw > 0 && F > 0 && groups > 0 && B % groups == 0 && F % 4 == 0, SKY_ERR_INVALID, "bad dimension");
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = (B / groups) * hw;
    SKY_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)groups * F * 2 * sizeof(double), st));
    const int slabs = (F + 31) / 32;
    int chunks = (4 * 148) / (slabs * groups);
    if (chunks < 1) chunks = 1;
    int rows_per_cta = (rows + chunks - 1) / chunks;
    if (rows_per_cta < 64) rows_per_cta = 64;
    chunks = (rows + rows_per_cta - 1) / rows_per_cta;
    bn_bwd_reduce_kernel<<<dim3(slabs, chunks, groups), TS_THREADS, 0, st>>>(x, mean_var, dy, act, sums, rows, F, rows_per_cta, eps, slope);
    SKY_CHECK_LAUNCH();
    bn_bwd_apply_kernel<<<ts_blocks((long)rows * groups * (F / 4)), TS_THREADS, 0, st>>>(x, mean_var, gamma, dy, act, sums, dx, dgamma, dbeta,
                                                                                        (long)rows, F, groups, eps, slope);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_resize_bilinear_bwd(const float *dy, float *dx, int B, int h, int w, int C, int oh, int ow, int accumulate, void *stream)
{
    SKY_REQUIRE(dy && dx, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && oh > 0 && ow > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(C % 4 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0, SKY_ERR_UNSUPPORTED, "resize adjoint needs C %% 4 == 0 and 16-byte aligned tensors");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * h * w * C * sizeof(float), st));
    resize_bilinear_bwd_kernel<<<ts_blocks((long)B * oh * ow * (C / 4), TS_THREADS, 148 * 16), TS_THREADS, 0, st>>>(dy, dx, B, h, w, C, oh, ow);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_sun_radiance_bwd(const float *sm, const float *gmax, const float *gb, const float *d_out3, float *dnorm, double *dgb64,
                                    double *red2, int B, int hw, float eps, void *stream)
{
    SKY_REQUIRE(sm && gmax && gb && d_out3 && dnorm && dgb64 && red2 && B > 0 && hw > 0, SKY_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(dgb64, 0, (size_t)B * 2 * sizeof(double), st));
    SKY_CHECK_CUDA(cudaMemsetAsync(red2, 0, 2 * sizeof(double), st));
    const float sqrt_pi = sqrtf(3.14159265358979323846f);
    int chunks = (hw + 4 * TS_THREADS - 1) / (4 * TS_THREADS);
    if (chunks > 16) chunks = 16;
    sun_radiance_bwd_kernel<<<dim3(chunks, B), TS_THREADS, 0, st>>>(sm, gmax, gb, d_out3, dnorm, dgb64, red2, hw, eps, sqrt_pi);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_maxnorm_bwd(const float *sm, const float *gmax, const float *dnorm, const double *red2, float *dsm, long n, int accumulate,
                               void *stream)
{
    SKY_REQUIRE(sm && gmax && dnorm && red2 && dsm && n > 0, SKY_ERR_INVALID, "bad arguments");
    maxnorm_bwd_kernel<<<ts_blocks(n), TS_THREADS, 0, (cudaStream_t)stream>>>(sm, gmax, dnorm, red2, dsm, n, accumulate);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_sunrad_heads_bwd(const float *flat, const float *W, const double *dgb64, float *dW, float *dbias, float *dflat, int B, int K,
                                    void *stream)
{
    SKY_REQUIRE(flat && W && dgb64 && dW && dbias && dflat && B > 0 && K > 0, SKY_ERR_INVALID, "bad arguments");
    sunrad_heads_bwd_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flat, W, dgb64, dW, dbias, dflat, B, K);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_train_tail_fwd(const float *c_sky, const float *c_sun, const float *ldr, const float *sun_rad_gamma, const float *hdr_t,
                                  float threshold, float slope, float *y_gamma, float *y_lin, float *sky_lin, float *sun_lin, float *alpha,
                                  double *l1_acc, long npix, void *stream)
{
    SKY_REQUIRE(c_sky && c_sun && ldr && sun_rad_gamma && hdr_t && y_gamma && y_lin && sky_lin && sun_lin && alpha && l1_acc && npix > 0 && threshold > 0.f,
                SKY_ERR_INVALID, "bad arguments");
    train_tail_fwd_kernel<<<ts_blocks(npix), TS_THREADS, 0, (cudaStream_t)stream>>>(c_sky, c_sun, ldr, sun_rad_gamma, hdr_t, threshold, slope, y_gamma,
                                                                                     y_lin, sky_lin, sun_lin, alpha, l1_acc, npix);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_train_tail_bwd(const float *c_sky, const float *c_sun, const float *ldr, const float *sun_rad_gamma, const float *alpha,
                                  const float *y_lin, const float *hdr_t, const float *g_dog, const float *g_dis8, const float *g_vgg4, float w_l1,
                                  float w_vgg, float slope, float *dc_sky, float *dc_sun, float *d_sun_rad_gamma, long npix, void *stream)
{
    SKY_REQUIRE(c_sky && c_sun && ldr && sun_rad_gamma && alpha && y_lin && hdr_t && dc_sky && dc_sun && d_sun_rad_gamma && npix > 0, SKY_ERR_INVALID,
                "bad arguments");
    train_tail_bwd_kernel<<<ts_blocks(npix), TS_THREADS, 0, (cudaStream_t)stream>>>(c_sky, c_sun, ldr, sun_rad_gamma, alpha, y_lin, hdr_t, g_dog, g_dis8,
                                                                                     g_vgg4, w_l1, w_vgg, slope, dc_sky, dc_sun, d_sun_rad_gamma, npix);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_lsgan_bwd(const float *d_same, float *g_same, double *acc, int B, int hh, int ww, int r0, int r1, int c0, int c1, float target,
                             float scale, void *stream)
{
    SKY_REQUIRE(d_same && B > 0 && hh > 0 && ww > 0 && r0 >= 0 && r1 <= hh && c0 >= 0 && c1 <= ww, SKY_ERR_INVALID, "bad arguments");
    lsgan_bwd_kernel<<<ts_blocks((long)B * hh * ww), TS_THREADS, 0, (cudaStream_t)stream>>>(d_same, g_same, acc, B, hh, ww, r0, r1, c0, c1, target, scale);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_l1_bwd(const float *a, const float *b, float *g, double *acc, long n, float scale, int accumulate, void *stream)
{
    SKY_REQUIRE(a && b && n > 0, SKY_ERR_INVALID, "bad arguments");
    l1_bwd_kernel<<<ts_blocks(n), TS_THREADS, 0, (cudaStream_t)stream>>>(a, b, g, acc, n, scale, accumulate);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_maxpool2x2_bwd_relu(const float *x, const float *dy, float *dx, int B, int h, int w, int C, int relu_mask, void *stream)
{
    SKY_REQUIRE(x && dy && dx && B > 0 && h > 0 && w > 0 && C > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(C % 4 == 0, SKY_ERR_UNSUPPORTED, "max-pool kernel needs C %% 4 == 0 (got %d)", C);
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    maxpool2x2_bwd_relu_kernel<<<ts_blocks((long)B * oh * ow * (C / 4), TS_THREADS, 148 * 16), TS_THREADS, 0, (cudaStream_t)stream>>>(x, dy, dx, B, h, w,
                                                                                                                                    C, oh, ow, relu_mask);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
