// Backward kernels of train_sun.sun_train_step (train_sun.py:220-264) that the inference path does not already have:
//
//   kl_divergence_bwd        d KLDivergence / d y_pred (clip gradient included)
//   dog_l1_bwd, dog_base_bwd the adjoint of tf_utils.DoG under the L1 of train_sun.py:247-253: sign of the four differences pushed
//                            back through the five Gaussians, the base Gaussian (REFLECT padding folds the border taps back) and
//                            the x2 bilinear resize
//   softmax_bwd_rows         softmax backward for an arbitrary upstream gradient, fused with the ReLU mask of sunpose_net.py:68
//   dense_bwd_filter         dW = x^T . dy, db = column sums of dy (Keras Dense, sunpose_net.py:49-52)
//   da_conv2d_smallc_bwd_filter   weight / bias gradient of the distortion-aware layer on the 3-channel panorama (sunlayer1.conv1)
// fp32 like the reference; scatter-type adjoints use float atomics (summation order is not deterministic).
#include "da_conv.cuh"

namespace sky {

// ---- KL -----------------------------------------------------------------------------------------------------------------
// L = scale * sum t log(t / p), t = clip(y_true), p = clip(y_pred): dL/dy_pred = -scale * t / p inside the clip range, 0 outside.
__global__ void kl_bwd_kernel(const float *__restrict__ yt, const float *__restrict__ yp, float *__restrict__ g, long n, float scale,
                              int accumulate)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float t = fminf(fmaxf(yt[e], 1e-7f), 1.f), p = yp[e];
        const float d = (p >= 1e-7f && p <= 1.f) ? -scale * t / p : 0.f;
        g[e] = accumulate ? g[e] + d : d;
    }
}

// ---- DoG adjoint ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_i(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ void gauss_taps2(float sigma, float *edge, float *centre)
{
    const float q = expf(-1.f / (2.f * sigma * sigma));
    const float s = 1.f + 2.f * q;
    *edge = q / s;
    *centre = 1.f / s;
}

// dbase_a += scale * sum_l (K_{l+1} - K_l)^T sign(DoG_l(base_a) - DoG_l(base_b))      (dbase_a zeroed by the launcher)
__global__ void dog_l1_bwd_kernel(const float *__restrict__ base_a, const float *__restrict__ base_b, float *__restrict__ dbase, int B,
                                  int H2, int W2, int C, float scale)
{
    const float sig[5] = { 1.2262735f, 1.5450078f, 1.9465878f, 2.452547f, 3.0900156f };
    float kcc[5], kec[5], kee[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        float e, c;
        gauss_taps2(sig[l], &e, &c);
        kcc[l] = c * c; kec[l] = e * c; kee[l] = e * e;
    }
    const long total = (long)B * H2 * W2 * C;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % C), ox = (int)((o / C) % W2), oy = (int)((o / ((long)C * W2)) % H2), b = (int)(o / ((long)C * W2 * H2));
        const size_t img = (size_t)b * H2 * W2 * C + c;
        float ctr[2], edge[2] = { 0.f, 0.f }, corner[2] = { 0.f, 0.f };
        const float *src[2] = { base_a + img, base_b + img };
        int ys[3], xs[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) { ys[d] = reflect_i(oy + d - 1, H2); xs[d] = reflect_i(ox + d - 1, W2); }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float v = __ldg(src[s] + ((size_t)ys[dy] * W2 + xs[dx]) * C);
                    if (dy == 1 && dx == 1) ctr[s] = v;
                    else if (dy == 1 || dx == 1) edge[s] += v;
                    else corner[s] += v;
                }
        float wc = 0.f, we = 0.f, wk = 0.f;     // coefficients of the centre / edge / corner taps after summing the four levels
        float gprev[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) gprev[s] = kcc[0] * ctr[s] + kec[0] * edge[s] + kee[0] * corner[s];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            float gn[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) gn[s] = kcc[l + 1] * ctr[s] + kec[l + 1] * edge[s] + kee[l + 1] * corner[s];
            const float delta = (gn[0] - gprev[0]) - (gn[1] - gprev[1]);
            const float sgn = delta > 0.f ? 1.f : (delta < 0.f ? -1.f : 0.f);
            wc += sgn * (kcc[l + 1] - kcc[l]); we += sgn * (kec[l + 1] - kec[l]); wk += sgn * (kee[l + 1] - kee[l]);
            gprev[0] = gn[0]; gprev[1] = gn[1];
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float coef = (dy == 1 && dx == 1) ? wc : ((dy == 1 || dx == 1) ? we : wk);
                if (coef != 0.f) atomicAdd(dbase + img + ((size_t)ys[dy] * W2 + xs[dx]) * C, scale * coef);
            }
    }
}

// dx (+)= Up^T G0^T dbase : adjoint of dog_base_kernel.  dx zeroed by the launcher unless accumulate.
__global__ void dog_base_bwd_kernel(const float *__restrict__ dbase, float *__restrict__ dx, int B, int h, int w, int C, float sigma0)
{
    const int H2 = 2 * h, W2 = 2 * w;
    float ke, kc;
    gauss_taps2(sigma0, &ke, &kc);
    const long total = (long)B * H2 * W2 * C;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const float g = dbase[o];
        if (g == 0.f) continue;
        const int c = (int)(o % C), ox = (int)((o / C) % W2), oy = (int)((o / ((long)C * W2)) % H2), b = (int)(o / ((long)C * W2 * H2));
        float *img = dx + (size_t)b * h * w * C + c;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = reflect_i(oy + dy, H2);
            const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)yy, 0.5f), 0.5f), 0.5f);
            const float fly = floorf(fy), ly = __fsub_rn(fy, fly);
            const int ylo = max((int)fly, 0), yhi = min((int)ceilf(fy), h - 1);
#pragma unroll
            for (int dxx = -1; dxx <= 1; ++dxx) {
                const int xx = reflect_i(ox + dxx, W2);
                const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)xx, 0.5f), 0.5f), 0.5f);
                const float flx = floorf(fx), lx = __fsub_rn(fx, flx);
                const int xlo = max((int)flx, 0), xhi = min((int)ceilf(fx), w - 1);
                const float gq = g * (dy == 0 ? kc : ke) * (dxx == 0 ? kc : ke);
                atomicAdd(img + ((size_t)ylo * w + xlo) * C, gq * (1.f - lx) * (1.f - ly));
                atomicAdd(img + ((size_t)ylo * w + xhi) * C, gq * lx * (1.f - ly));
                atomicAdd(img + ((size_t)yhi * w + xlo) * C, gq * (1.f - lx) * ly);
                atomicAdd(img + ((size_t)yhi * w + xhi) * C, gq * lx * ly);
            }
        }
    }
}

// ---- softmax backward -------------------------------------------------------------------------------------------------------
// g_z[i] = (act[i] > 0) * sm[i] * (g[i] - sum_j g[j] sm[j]);  one CTA per row
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float *__restrict__ sm, const float *__restrict__ g,
                                                               const float *__restrict__ act, float *__restrict__ gz, int N)
{
    __shared__ float red[8];
    const size_t base = (size_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dot = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) dot = fmaf(g[base + i], sm[base + i], dot);
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    dot = 0.f;
    for (int i = 0; i < 8; ++i) dot += red[i];
    for (int i = threadIdx.x; i < N; i += 256) {
        const float v = sm[base + i] * (g[base + i] - dot);
        gz[base + i] = (!act || act[base + i] > 0.f) ? v : 0.f;
    }
}

// ---- Dense weight gradient --------------------------------------------------------------------------------------------------
// dW[k, n] = sum_b x[b, k] dy[b, n];  CTA tile: 64 k-rows x 256 columns, batch staged 32 rows at a time.  HBM-bound on the dW write.
constexpr int DW_BN = 256, DW_BK = 64, DW_THREADS = 256;
__global__ void __launch_bounds__(DW_THREADS) dense_bwd_filter_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                                      float *__restrict__ dW, float *__restrict__ db, int B, int K, int N)
{
    __shared__ __align__(16) float dys[32][DW_BN];
    __shared__ __align__(16) float xs[32][DW_BK];
    const int n0 = blockIdx.x * DW_BN, k0 = blockIdx.y * DW_BK;
    const int n4 = threadIdx.x & 63, kq = threadIdx.x >> 6;          // 64 float4 columns x 4 row groups of 16 k-rows
    float4 acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b0 = 0; b0 < B; b0 += 32) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * DW_BN; e += DW_THREADS) {
            const int bb = e / DW_BN, nn = e % DW_BN;
            dys[bb][nn] = (b0 + bb < B && n0 + nn < N) ? __ldg(dy + (size_t)(b0 + bb) * N + n0 + nn) : 0.f;
        }
        for (int e = threadIdx.x; e < 32 * DW_BK; e += DW_THREADS) {
            const int bb = e / DW_BK, kk = e % DW_BK;
            xs[bb][kk] = (b0 + bb < B && k0 + kk < K) ? __ldg(x + (size_t)(b0 + bb) * K + k0 + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int bb = 0; bb < 32; ++bb) {
            const float4 d = *reinterpret_cast<const float4 *>(&dys[bb][4 * n4]);
            if (kq == 0) { bsum.x += d.x; bsum.y += d.y; bsum.z += d.z; bsum.w += d.w; }
            const float4 *xr = reinterpret_cast<const float4 *>(&xs[bb][16 * kq]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 xv = xr[q];
                const float xa[4] = { xv.x, xv.y, xv.z, xv.w };
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float4 &a = acc[4 * q + u];
                    a.x = fmaf(xa[u], d.x, a.x); a.y = fmaf(xa[u], d.y, a.y); a.z = fmaf(xa[u], d.z, a.z); a.w = fmaf(xa[u], d.w, a.w);
                }
            }
        }
    }
    const int n = n0 + 4 * n4;
    if (n < N) {      // N % 4 == 0
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int k = k0 + 16 * kq + r;
            if (k < K) *reinterpret_cast<float4 *>(dW + (size_t)k * N + n) = acc[r];
        }
        if (db && blockIdx.y == 0 && kq == 0) *reinterpret_cast<float4 *>(db + n) = bsum;
    }
}

// ---- weight gradient of the small-C distortion-aware layer ------------------------------------------------------------------
// dK[t*C + c, f] = sum_{pixels} pix(t, c) dy(f), dbias[f] = sum dy(f).  One CTA per (128-pixel row segment, row group, sample):
// phase 1 (thread = pixel) evaluates the reference geometry and the four-corner blend of every tap into shared memory, phase 2
// (thread = filter x tap group) contracts over the 128 pixels; partial sums go to global memory with atomics.
constexpr int SW_THREADS = 128, SW_F = 32, SW_ROWS = 4;
template <int C>
__global__ void __launch_bounds__(SW_THREADS)
da_smallc_wgrad_kernel(const float *__restrict__ x, const float *__restrict__ offsets, const float *__restrict__ dy,
                       float *__restrict__ dkernel, float *__restrict__ dbias, int h, int w, int F, int k, int in_h, int in_w, int ph0,
                       int pw0)
{
    extern __shared__ float sw[];
    const int k2 = k * k, KC = k2 * C;
    float *pixs = sw;                       // [KC][128]
    float *dys = sw + (size_t)KC * SW_THREADS;   // [128][32]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SW_THREADS, b = blockIdx.z;
    const int f = tid & 31, kq = tid >> 5;  // phase 2: filter, tap group (kc = kq, kq + 4, ...)
    constexpr int MAXO = 48;                // ceil(11*11*4 / 4 / ... ) bound checked on the host: KC <= 4 * MAXO
    float acc[MAXO];
#pragma unroll
    for (int q = 0; q < MAXO; ++q) acc[q] = 0.f;
    float bacc = 0.f;
    const float *img = x + (size_t)b * h * w * C;
    for (int ii = 0; ii < SW_ROWS; ++ii) {
        const int i = blockIdx.y * SW_ROWS + ii;
        if (i >= h) break;
        const int j = x0 + tid;
        __syncthreads();
        for (int t = 0; t < k2; ++t) {
            float pix[C];
#pragma unroll
            for (int c = 0; c < C; ++c) pix[c] = 0.f;
            if (j < w) {
                const float2 o = __ldg(reinterpret_cast<const float2 *>(offsets) + (size_t)i * k2 + t);
                const Sample s = da_sample(i, j, t / k, t % k, o.x, o.y, in_h, in_w);
                const int ys[4] = { s.y0 - ph0, s.y0 - ph0, s.y1 - ph0, s.y1 - ph0 };
                const int xs[4] = { s.x0 - pw0, s.x1 - pw0, s.x0 - pw0, s.x1 - pw0 };
                const float wq[4] = { s.w0, s.w1, s.w2, s.w3 };
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool inside = ys[q] >= 0 && ys[q] < h && xs[q] >= 0 && xs[q] < w;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float px = inside ? __ldg(img + ((size_t)ys[q] * w + xs[q]) * C + c) : 0.f;
                        pix[c] = (q == 0) ? __fmul_rn(wq[0], px) : __fadd_rn(pix[c], __fmul_rn(wq[q], px));
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) pixs[(size_t)(t * C + c) * SW_THREADS + tid] = pix[c];
        }
        for (int e = tid; e < SW_THREADS * SW_F; e += SW_THREADS) {
            const int px = e / SW_F, ff = e % SW_F;
            dys[e] = (x0 + px < w && ff < F) ? __ldg(dy + (((size_t)b * h + i) * w + x0 + px) * F + ff) : 0.f;
        }
        __syncthreads();
        for (int px = 0; px < SW_THREADS; px += 4) {            // four pixels per broadcast 128-bit read of the blended samples
            const float d0 = dys[px * SW_F + f], d1 = dys[(px + 1) * SW_F + f], d2 = dys[(px + 2) * SW_F + f], d3 = dys[(px + 3) * SW_F + f];
            if (kq == 0) bacc += (d0 + d1) + (d2 + d3);
#pragma unroll
            for (int q = 0; q < MAXO; ++q) {
                const int kc = kq + 4 * q;
                if (kc < KC) {
                    const float4 pv = *reinterpret_cast<const float4 *>(pixs + (size_t)kc * SW_THREADS + px);
                    acc[q] = fmaf(pv.w, d3, fmaf(pv.z, d2, fmaf(pv.y, d1, fmaf(pv.x, d0, acc[q]))));
                }
            }
        }
    }
    if (f < F) {
#pragma unroll
        for (int q = 0; q < MAXO; ++q) {
            const int kc = kq + 4 * q;
            if (kc < KC) atomicAdd(dkernel + (size_t)kc * F + f, acc[q]);
        }
        if (dbias && kq == 0) atomicAdd(dbias + f, bacc);
    }
}

static int ew_grid(long total, int cap = 148 * 8)
{
    long b = (total + 255) / 256;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace sky

using namespace sky;

extern "C" int sky_kl_divergence_bwd(const float *y_true, const float *y_pred, float *g, long n, float scale, int accumulate, void *stream)
{
    SKY_REQUIRE(y_true && y_pred && g && n > 0, SKY_ERR_INVALID, "bad arguments");
    kl_bwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(y_true, y_pred, g, n, scale, accumulate);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dog_l1_bwd(const float *base_a, const float *base_b, float *dbase_a, int B, int H2, int W2, int C, float scale,
                              void *stream)
{
    SKY_REQUIRE(base_a && base_b && dbase_a && B > 0 && H2 > 1 && W2 > 1 && C > 0, SKY_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long total = (long)B * H2 * W2 * C;
    SKY_CHECK_CUDA(cudaMemsetAsync(dbase_a, 0, total * sizeof(float), st));
    dog_l1_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(base_a, base_b, dbase_a, B, H2, W2, C, scale);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dog_base_bwd(const float *dbase, float *dx, int B, int h, int w, int C, int accumulate, void *stream)
{
    SKY_REQUIRE(dbase && dx && B > 0 && h > 1 && w > 1 && C > 0, SKY_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * h * w * C * sizeof(float), st));
    dog_base_bwd_kernel<<<ew_grid((long)B * 4 * h * w * C), 256, 0, st>>>(dbase, dx, B, h, w, C, 1.2489996f);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_softmax_bwd_rows(const float *sm, const float *g, const float *act, float *gz, int rows, int N, void *stream)
{
    SKY_REQUIRE(sm && g && gz && rows > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    softmax_bwd_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(sm, g, act, gz, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_dense_bwd_filter(const float *x, const float *dy, float *dW, float *db, int B, int K, int N, void *stream)
{
    SKY_REQUIRE(x && dy && dW && B > 0 && K > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(N % 4 == 0 && ((uintptr_t)dW & 15) == 0 && (!db || ((uintptr_t)db & 15) == 0), SKY_ERR_UNSUPPORTED,
                "Dense weight gradient needs units %% 4 == 0 and 16-byte aligned outputs (N=%d)", N);
    dim3 grid((N + DW_BN - 1) / DW_BN, (K + DW_BK - 1) / DW_BK);
    dense_bwd_filter_kernel<<<grid, DW_THREADS, 0, (cudaStream_t)stream>>>(x, dy, dW, db, B, K, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_conv2d_smallc_bwd_filter(const float *x, const float *dy, const float *offsets, float *dkernel, float *dbias, int B,
                                               int h, int w, int C, int F, int k, void *stream)
{
    SKY_REQUIRE(x && dy && offsets && dkernel && B > 0 && h > 0 && w > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(C >= 1 && C <= 4 && F >= 1 && F <= SW_F && k >= 3 && k * k * C <= 4 * 48, SKY_ERR_UNSUPPORTED,
                "small-C weight gradient covers C <= 4, F <= 32, k*k*C <= 192 (got C=%d F=%d k=%d)", C, F, k);
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(dkernel, 0, (size_t)k * k * C * F * sizeof(float), st));
    if (dbias) SKY_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)F * sizeof(float), st));
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    const size_t smem = ((size_t)k * k * C * SW_THREADS + SW_THREADS * SW_F) * sizeof(float);
    dim3 grid((w + SW_THREADS - 1) / SW_THREADS, (h + SW_ROWS - 1) / SW_ROWS, B);
#define SKY_LAUNCH_SW(CC)                                                                                                      \
    do {                                                                                                                       \
        SKY_ENSURE_DYN_SMEM(da_smallc_wgrad_kernel<CC>, 200 * 1024);                    \
                                                                                                                              \
        da_smallc_wgrad_kernel<CC><<<grid, SW_THREADS, smem, st>>>(x, offsets, dy, dkernel, dbias, h, w, F, k, h + pht, w + pwt, ph0, pw0); \
    } while (0)
    SKY_REQUIRE(smem <= 200 * 1024, SKY_ERR_UNSUPPORTED, "kernel too large for the staged blend (k=%d C=%d)", k, C);
    switch (C) {
        case 1: SKY_LAUNCH_SW(1); break;
        case 2: SKY_LAUNCH_SW(2); break;
        case 3: SKY_LAUNCH_SW(3); break;
        default: SKY_LAUNCH_SW(4); break;
    }
#undef SKY_LAUNCH_SW
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
