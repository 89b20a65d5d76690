// Plain SAME stride-1 convolution for image-like inputs (C <= 4), e.g. the generator's first layer conv1_d
// (generator.py:60: 7x7, 3 -> 32 on the LDR panorama).  K = k*k*C is tiny (147) and the layer is HBM/FMA-bound, so it
// does not go through the tensor-core pipeline: one thread per output pixel keeps F <= 32 accumulators in registers,
// the input patch (k rows x (128 + k - 1) pixels) and the whole kernel variable are staged in shared memory, fp32 FMA
// in the reference's (tap, c) order.  Epilogue: bias, optional LeakyReLU, coalesced stores, per-(sample, filter)
// moments for the instance norm that follows.
#include "da_conv.cuh"

namespace sky {

constexpr int SC_THREADS = 128;   // pixels per block (one row segment)
constexpr int SC_FMAX = 32;

// bias, optional LeakyReLU, coalesced stores, per-(sample, filter) moments for the instance norm that follows
__device__ __forceinline__ void smallc_epilogue(float (&acc)[SC_FMAX], const float *__restrict__ bias, float *__restrict__ y,
                                                double *__restrict__ stats, float *red, int b, int i, int j, int h, int w, int F,
                                                int flags, float slope, int ldF = 0)
{
    // F = filters of this launch (<= 32); ldF = filters of the layer when the launch covers a slice of them (bias, y, stats pre-offset)
    if (ldF == 0) ldF = F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ok = j < w;
#pragma unroll
    for (int f = 0; f < SC_FMAX; ++f) {
        float v = acc[f] + ((bias && f < F) ? __ldg(bias + f) : 0.f);
        if (flags & SKY_EPI_LEAKY_RELU) v = v > 0.f ? v : v * slope;
        acc[f] = v;
    }
    if (ok) {
        float *dst = y + (((size_t)b * h + i) * w + j) * ldF;
        if (F == SC_FMAX && (ldF & 3) == 0) {
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
            atomicAdd(stats + ((size_t)b * ldF + f) * 2 + which, s);
        }
    }
}

// blockIdx.x = column tile * nslices + filter slice (slices of 32 filters: VGG conv1_1 has 64)
template <int C>
__global__ void __launch_bounds__(SC_THREADS)
conv2d_smallc_kernel(const float *__restrict__ x, const float *__restrict__ kernel, const float *__restrict__ bias,
                     float *__restrict__ y, double *__restrict__ stats, int h, int w, int Ftot, int nslices, int k, int flags, float slope)
{
    extern __shared__ float sm[];
    const int r = k / 2, pw = SC_THREADS + k - 1;
    float *wts = sm;                          // [k*k*C][SC_FMAX]
    float *patch = sm + k * k * C * SC_FMAX;  // [k][pw][C]
    float *red = patch + k * pw * C;          // [4 warps][SC_FMAX][2]
    const int tid = threadIdx.x;
    const int slice = blockIdx.x % nslices, f0 = slice * SC_FMAX, F = min(SC_FMAX, Ftot - f0);
    const int x0 = (blockIdx.x / nslices) * SC_THREADS, i = blockIdx.y, b = blockIdx.z;

    for (int e = tid; e < k * k * C * SC_FMAX; e += SC_THREADS) {
        const int f = e % SC_FMAX, row = e / SC_FMAX;
        wts[e] = f < F ? kernel[(size_t)row * Ftot + f0 + f] : 0.f;
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
    smallc_epilogue(acc, bias ? bias + f0 : nullptr, y + f0, stats ? stats + 2 * f0 : nullptr, red, b, i, x0 + tid, h, w, F, flags, slope, Ftot);
}


// The distortion-aware layer on an image-like input (sunlayer1.conv1 with the wiring of sunpose_net.py:11: 7x7, 3 -> 32 on the LDR
// panorama).  Same structure as above, but every (pixel, tap) is sampled with the reference geometry (da_sample) and blended from
// four corners.  The rows [i + hy_lo, i + hy_hi] (halo from the offset table) are staged at FULL width, so the 360-degree wrap
// stays inside the staged rows; corners in other rows (zenith-row taps on tall maps) are read from global memory.
template <int C>
__global__ void __launch_bounds__(SC_THREADS)
da_conv2d_smallc_kernel(const float *__restrict__ x, const float *__restrict__ offsets, const float *__restrict__ kernel,
                        const float *__restrict__ bias, float *__restrict__ y, double *__restrict__ stats, int h, int w, int F, int k,
                        int hy_lo, int nrows, int in_h, int in_w, int ph0, int pw0, int flags, float slope)
{
    extern __shared__ float sm[];
    const int k2 = k * k;
    float *wts = sm;                           // [k*k*C][SC_FMAX]
    float2 *offs = reinterpret_cast<float2 *>(sm + k2 * C * SC_FMAX);   // [k2] (y, x) offsets of this output row
    float *rows = sm + k2 * C * SC_FMAX + 2 * k2;   // [nrows][w][C]
    float *red = rows + nrows * w * C;              // [4 warps][SC_FMAX][2]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * SC_THREADS, i = blockIdx.y, b = blockIdx.z;
    const int r0 = i + hy_lo;                  // first staged input row (may be negative: zero halo)

    for (int e = tid; e < k2 * C * SC_FMAX; e += SC_THREADS) {
        const int f = e % SC_FMAX, row = e / SC_FMAX;
        wts[e] = f < F ? kernel[(size_t)row * F + f] : 0.f;
    }
    for (int e = tid; e < nrows * w * C; e += SC_THREADS) {
        const int yy = r0 + e / (w * C);
        rows[e] = (yy >= 0 && yy < h) ? x[((size_t)b * h + yy) * w * C + e % (w * C)] : 0.f;
    }
    for (int t = tid; t < k2; t += SC_THREADS) offs[t] = reinterpret_cast<const float2 *>(offsets)[(size_t)i * k2 + t];
    __syncthreads();

    const int j = x0 + tid;
    float acc[SC_FMAX];
#pragma unroll
    for (int f = 0; f < SC_FMAX; ++f) acc[f] = 0.f;
    if (j < w) {
        const float *img = x + (size_t)b * h * w * C;
        for (int t = 0; t < k2; ++t) {
            const float2 o = offs[t];
            const Sample s = da_sample(i, j, t / k, t % k, o.x, o.y, in_h, in_w);
            const int ys[4] = { s.y0 - ph0, s.y0 - ph0, s.y1 - ph0, s.y1 - ph0 };
            const int xs[4] = { s.x0 - pw0, s.x1 - pw0, s.x0 - pw0, s.x1 - pw0 };
            const float wq[4] = { s.w0, s.w1, s.w2, s.w3 };
            float pix[C];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool inside = ys[q] >= 0 && ys[q] < h && xs[q] >= 0 && xs[q] < w;     // else: zero halo (:125-150)
                const int rr = ys[q] - r0;
                const float *src = (rr >= 0 && rr < nrows) ? rows + (rr * w + xs[q]) * C : img + ((size_t)ys[q] * w + xs[q]) * C;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float px = inside ? src[c] : 0.f;
                    pix[c] = (q == 0) ? __fmul_rn(wq[0], px) : __fadd_rn(pix[c], __fmul_rn(wq[q], px));   // add_n order (:112-113)
                }
            }
            const float4 *wrow = reinterpret_cast<const float4 *>(wts + t * C * SC_FMAX);
#pragma unroll
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int f4 = 0; f4 < SC_FMAX / 4; ++f4) {
                    const float4 wv = wrow[c * (SC_FMAX / 4) + f4];
                    acc[4 * f4 + 0] = fmaf(pix[c], wv.x, acc[4 * f4 + 0]);
                    acc[4 * f4 + 1] = fmaf(pix[c], wv.y, acc[4 * f4 + 1]);
                    acc[4 * f4 + 2] = fmaf(pix[c], wv.z, acc[4 * f4 + 2]);
                    acc[4 * f4 + 3] = fmaf(pix[c], wv.w, acc[4 * f4 + 3]);
                }
            }
        }
    }
    smallc_epilogue(acc, bias, y, stats, red, b, i, j, h, w, F, flags, slope);
}

// ---------------------------------------------------------------------------------------------------------------------
// Plain SAME stride-1 convolution with very few filters (F <= 4): conv1_f / conv1_u of the decoders (generator.py:76,85: 7x7, 32 -> 3).
// 1.2 GFLOP at B = 32 — the layer is all input reuse, and on the tensor-core path its cost is the 49 producer / MMA hand-shakes per
// tile (170 us).  Here a CTA stages the (8 + k - 1) x (32 + k - 1) input patch of an 8 x 32 output tile in shared memory (pixel stride
// C + 4 floats: the 128-bit reads of 32 adjacent pixels spread over all banks) and the kernel variable as [tap][c][4]; a thread owns two
// output pixels (rows ty and ty + 4), so one broadcast weight read feeds six FMAs.  fp32 FMA on the TF32-rounded packed weights.
// Epilogue: bias, LeakyReLU, residual, ReLU, alpha blend (SKY_EPI_SUN_BLEND), log decompression — the whole tail of generator inference.
constexpr int SF_TH = 8, SF_TW = 32, SF_THREADS = 128;

__global__ void __launch_bounds__(SF_THREADS)
conv2d_smallf_kernel(const float *__restrict__ x, const float *__restrict__ packed, const float *__restrict__ bias,
                     const float *__restrict__ residual, const float *__restrict__ aux, float *__restrict__ y, int B, int h, int w, int C,
                     int F, int Fp, int k, int flags, float slope, float threshold)
{
    extern __shared__ __align__(16) float sf[];
    const int k2 = k * k, r = k / 2, PW = SF_TW + k - 1, PH = SF_TH + k - 1, PS = C + 4;
    float *wts = sf;                              // [k2][C][4]
    float *patch = sf + (size_t)k2 * C * 4;       // [PH][PW][PS]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int tiles_x = (w + SF_TW - 1) / SF_TW, tiles_y = (h + SF_TH - 1) / SF_TH, ntiles = tiles_x * tiles_y * B;

    // kernel variable from the packed tensor-core image: k-block (c / 32) * k2 + tap, row f, element c % 32 (SWIZZLE_128B)
    for (int e = tid; e < k2 * C; e += SF_THREADS) {
        const int t = e / C, c = e % C;
        int kb = (c / 32) * k2 + t, kk = c & 31;                       // chunk-major k-blocks when C % 32 == 0 (da_pack_weights_kernel)
        if (C % 32 != 0) { kb = (t * C + c) >> 5; kk = (t * C + c) & 31; }
        const float *tile = packed + (size_t)kb * Fp * 32;
        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
        float *wp = &wv.x;
        for (int f = 0; f < F; ++f) wp[f] = __ldg(tile + (sw128_offset((uint32_t)f, (uint32_t)(kk >> 2)) >> 2) + (kk & 3));
        reinterpret_cast<float4 *>(wts)[e] = wv;
    }
    const int c4n = C / 4;
    // persistent over tiles: the unpacked kernel variable is staged once per CTA
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / (tiles_x * tiles_y), trem = tile % (tiles_x * tiles_y);
    const int i0 = (trem / tiles_x) * SF_TH, j0 = (trem % tiles_x) * SF_TW;
    __syncthreads();                              // the previous tile's patch is no longer being read
    for (int e = tid; e < PH * PW * c4n; e += SF_THREADS) {
        const int c4 = e % c4n, px = (e / c4n) % PW, py = e / (c4n * PW);
        const int yy = i0 + py - r, xx = j0 + px - r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(reinterpret_cast<const float4 *>(x + (((size_t)b * h + yy) * w + xx) * C) + c4);
        *reinterpret_cast<float4 *>(patch + (size_t)(py * PW + px) * PS + 4 * c4) = v;
    }
    __syncthreads();

    float acc[2][4] = { { 0.f, 0.f, 0.f, 0.f }, { 0.f, 0.f, 0.f, 0.f } };
    for (int a = 0; a < k; ++a)
        for (int bb = 0; bb < k; ++bb) {
            const float *p0 = patch + (size_t)((ty + a) * PW + tx + bb) * PS;
            const float *p1 = p0 + (size_t)4 * PW * PS;                      // output row ty + 4
            const float4 *wt = reinterpret_cast<const float4 *>(wts) + (size_t)(a * k + bb) * C;
            for (int c4 = 0; c4 < c4n; ++c4) {
                const float4 v0 = *reinterpret_cast<const float4 *>(p0 + 4 * c4), v1 = *reinterpret_cast<const float4 *>(p1 + 4 * c4);
                const float a0[4] = { v0.x, v0.y, v0.z, v0.w }, a1[4] = { v1.x, v1.y, v1.z, v1.w };
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 wv = wt[4 * c4 + u];
                    acc[0][0] = fmaf(a0[u], wv.x, acc[0][0]); acc[0][1] = fmaf(a0[u], wv.y, acc[0][1]);
                    acc[0][2] = fmaf(a0[u], wv.z, acc[0][2]); acc[0][3] = fmaf(a0[u], wv.w, acc[0][3]);
                    acc[1][0] = fmaf(a1[u], wv.x, acc[1][0]); acc[1][1] = fmaf(a1[u], wv.y, acc[1][1]);
                    acc[1][2] = fmaf(a1[u], wv.z, acc[1][2]); acc[1][3] = fmaf(a1[u], wv.w, acc[1][3]);
                }
            }
        }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int i = i0 + ty + 4 * half, j = j0 + tx;
        if (i >= h || j >= w) continue;
        const size_t m = ((size_t)b * h + i) * w + j;
        float v[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float val = acc[half][f];
            if (f < F) {
                if (bias) val += __ldg(bias + f);
                if (flags & SKY_EPI_LEAKY_RELU) val = val > 0.f ? val : val * slope;
                if (flags & SKY_EPI_RESIDUAL) val += __ldg(residual + m * F + f);
                if (flags & SKY_EPI_RELU) val = fmaxf(val, 0.f);
            }
            v[f] = val;
        }
        if (flags & SKY_EPI_SUN_BLEND) sun_blend3(v, aux + m * 3, threshold);            // F == 3
#pragma unroll
        for (int f = 0; f < 4; ++f)
            if (f < F) {
                if (flags & SKY_EPI_LOG_DECOMPRESS) v[f] = (expf(v[f] * 2.3978953f) - 1.f) / 10.f;
                y[m * F + f] = v[f];
            }
    }
    }   // tiles
}

// Returns SKY_ERR_UNSUPPORTED (without an error text) when the layer is outside what this kernel covers.
int launch_fwd_smallf(const FwdArgs &a)
{
    if (a.F > 4 || a.C % 4 != 0 || a.k % 2 == 0 || a.plain_stride != 1 || a.stats != nullptr || a.math_mode != SKY_MATH_TF32 ||
        (a.flags & (SKY_EPI_FORCE_DIRECT | SKY_EPI_MASK)) || a.ldF > 0 || a.transposed)
        return SKY_ERR_UNSUPPORTED;
    const size_t smem = ((size_t)a.k * a.k * a.C * 4 + (size_t)(SF_TH + a.k - 1) * (SF_TW + a.k - 1) * (a.C + 4)) * sizeof(float);
    if (smem > 110 * 1024) return SKY_ERR_UNSUPPORTED;
    SKY_ENSURE_DYN_SMEM(conv2d_smallf_kernel, 110 * 1024);
    const int ntiles = ((a.w + SF_TW - 1) / SF_TW) * ((a.h + SF_TH - 1) / SF_TH) * a.B;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;          // two resident CTAs per SM
    conv2d_smallf_kernel<<<grid, SF_THREADS, smem, a.stream>>>(a.x, a.packed, a.bias, a.residual, a.aux, a.y, a.B, a.h, a.w, a.C, a.F, f_pad_of(a.F),
                                                               a.k, a.flags, a.slope, a.threshold);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

static int launch_fwd_smallc_da(const FwdArgs &a, const float *kernel)
{
    if (a.C > 4 || a.F > SC_FMAX || a.k > 11 || a.offsets_host == nullptr || (a.flags & ~SKY_EPI_LEAKY_RELU) ||
        (a.F == SC_FMAX && ((uintptr_t)a.y & 15) != 0))
        return SKY_ERR_UNSUPPORTED;
    int hy_lo, hy_hi, hx_lo, hx_hi, ph0, pht, pw0, pwt;
    compute_halo(a.offsets_host, a.h, a.w, a.k, &hy_lo, &hy_hi, &hx_lo, &hx_hi);
    pad_axis(a.h, a.k, &ph0, &pht);
    pad_axis(a.w, a.k, &pw0, &pwt);
    const int nrows = hy_hi - hy_lo + 1;
    const size_t smem = ((size_t)a.k * a.k * a.C * SC_FMAX + (size_t)nrows * a.w * a.C + 4 * SC_FMAX * 2 + 2 * a.k * a.k) * sizeof(float);
    if (smem > 200 * 1024) return SKY_ERR_UNSUPPORTED;
    dim3 grid((a.w + SC_THREADS - 1) / SC_THREADS, a.h, a.B);
#define SKY_LAUNCH_SCDA(CC)                                                                                                   \
    do {                                                                                                                      \
        SKY_ENSURE_DYN_SMEM(da_conv2d_smallc_kernel<CC>, 200 * 1024);                    \
                                                                                                                             \
        da_conv2d_smallc_kernel<CC><<<grid, SC_THREADS, smem, a.stream>>>(a.x, a.offsets, kernel, a.bias, a.y, a.stats, a.h, \
                                                                          a.w, a.F, a.k, hy_lo, nrows, a.h + pht, a.w + pwt, ph0, pw0,   \
                                                                          a.flags, a.slope);                                            \
    } while (0)
    switch (a.C) {
        case 1: SKY_LAUNCH_SCDA(1); break;
        case 2: SKY_LAUNCH_SCDA(2); break;
        case 3: SKY_LAUNCH_SCDA(3); break;
        default: SKY_LAUNCH_SCDA(4); break;
    }
#undef SKY_LAUNCH_SCDA
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_conv2d_smallc_fwd(const float *x, const float *kernel, const float *bias, float *y, double *stats, int B, int h,
                                     int w, int C, int F, int k, int epilogue_flags, float slope, void *stream)
{
    SKY_REQUIRE(x && kernel && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(C >= 1 && C <= 4 && F >= 1 && F <= 256 && k % 2 == 1 && k <= 11, SKY_ERR_UNSUPPORTED,
                "small-C conv covers C <= 4, F <= 256, odd k <= 11 (got C=%d F=%d k=%d)", C, F, k);
    SKY_REQUIRE(!(epilogue_flags & ~(SKY_EPI_LEAKY_RELU | SKY_EPI_RELU)) && (epilogue_flags & (SKY_EPI_LEAKY_RELU | SKY_EPI_RELU)) != (SKY_EPI_LEAKY_RELU | SKY_EPI_RELU),
                SKY_ERR_UNSUPPORTED, "small-C conv supports the LeakyReLU or the ReLU epilogue");
    SKY_REQUIRE(((uintptr_t)y & 15) == 0, SKY_ERR_INVALID, "y must be 16-byte aligned");
    if (epilogue_flags & SKY_EPI_RELU) { epilogue_flags = SKY_EPI_LEAKY_RELU; slope = 0.f; }      // ReLU = LeakyReLU of slope 0
    const int pw = SC_THREADS + k - 1;
    const int nslices = (F + SC_FMAX - 1) / SC_FMAX;
    const size_t smem = ((size_t)k * k * C * SC_FMAX + (size_t)k * pw * C + 4 * SC_FMAX * 2) * sizeof(float);
    dim3 grid(((w + SC_THREADS - 1) / SC_THREADS) * nslices, h, B);
    cudaStream_t st = (cudaStream_t)stream;
#define SKY_LAUNCH_SC(CC)                                                                                                     \
    do {                                                                                                                      \
        SKY_ENSURE_DYN_SMEM(conv2d_smallc_kernel<CC>, 200 * 1024);                    \
                                                                                                                             \
        conv2d_smallc_kernel<CC><<<grid, SC_THREADS, smem, st>>>(x, kernel, bias, y, stats, h, w, F, nslices, k, epilogue_flags, slope); \
    } while (0)
    switch (C) {
        case 1: SKY_LAUNCH_SC(1); break;
        case 2: SKY_LAUNCH_SC(2); break;
        case 3: SKY_LAUNCH_SC(3); break;
        default: SKY_LAUNCH_SC(4); break;
    }
#undef SKY_LAUNCH_SC
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_conv2d_smallc_fwd(const float *x, const float *offsets, const float *offsets_host, const float *kernel,
                                        const float *bias, float *y, double *stats, int B, int h, int w, int C, int F, int k,
                                        int epilogue_flags, float slope, void *stream)
{
    SKY_REQUIRE(x && offsets && offsets_host && kernel && bias && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(C >= 1 && C <= 4 && F >= 1 && F <= SC_FMAX && k >= 3 && k <= 11, SKY_ERR_UNSUPPORTED,
                "small-C distortion-aware conv covers C <= 4, F <= 32, odd k in 3..11 (got C=%d F=%d k=%d)", C, F, k);
    SKY_REQUIRE(!(epilogue_flags & ~(SKY_EPI_LEAKY_RELU)), SKY_ERR_UNSUPPORTED, "small-C conv supports only the LeakyReLU epilogue");
    FwdArgs a;
    a.x = x; a.offsets = offsets; a.offsets_host = offsets_host; a.packed = nullptr; a.bias = bias; a.residual = nullptr; a.y = y;
    a.stats = stats; a.B = B; a.h = h; a.w = w; a.C = C; a.F = F; a.k = k; a.flags = epilogue_flags; a.slope = slope;
    a.math_mode = 0; a.plain_stride = 0; a.stream = (cudaStream_t)stream;
    int rc = launch_fwd_smallc_da(a, kernel);
    SKY_REQUIRE(rc != SKY_ERR_UNSUPPORTED, rc, "small-C distortion-aware conv: panorama too wide for the staged rows (w=%d)", w);
    return rc;
}
