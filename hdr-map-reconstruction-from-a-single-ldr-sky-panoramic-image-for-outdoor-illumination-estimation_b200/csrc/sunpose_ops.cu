// Small kernels of the sun-position network (sunpose_net.py): 2x2/2 SAME max-pool (ops.maxpool2d, ops.py:287-300),
// Keras Dense as a weight-streaming skinny GEMM (fc1: [B, 8192] x [8192, 4096], fc2: [B, 4096] x [4096, 4096]; B = 32 rows
// against 134 MB / 67 MB of fp32 weights => HBM-bound, fp32 FMA on the CUDA cores, split-K over the SMs), the bias + ReLU
// finalisation of the split-K partial sums, and the row softmax (tf.nn.softmax, sunpose_net.py:70).
#include "sky_common.cuh"

namespace sky {

// tf.nn.max_pool(ksize 2, strides 2, SAME): out = ceil(n/2), window clipped at the bottom / right edge.
__global__ void maxpool2x2_kernel(const float *__restrict__ x, float *__restrict__ y, int B, int h, int w, int C, int oh, int ow)
{
    const int cv = C / 4;
    const long total = (long)B * oh * ow * cv;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % cv) * 4;
        const int ox = (int)((o / cv) % ow), oy = (int)((o / ((long)cv * ow)) % oh), b = (int)(o / ((long)cv * ow * oh));
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int yy = 2 * oy + dy, xx = 2 * ox + dx;
                if (yy < h && xx < w) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(x + (((size_t)b * h + yy) * w + xx) * C + c));
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            }
        *reinterpret_cast<float4 *>(y + (((size_t)b * oh + oy) * ow + ox) * C + c) = m;
    }
}

constexpr int DN_THREADS = 128;  // output columns per CTA
constexpr int DN_BM = 32;        // batch rows per pass (accumulators per thread)
constexpr int DN_BK = 64;        // k values staged per step

// y[b, n] += sum_{k in this CTA's K range} x[b, k] * W[k, n]      (y zeroed by the caller; fp32)
// grid: (N / 128, ksplit, ceil(B / 32)).  Thread t owns column n0 + t: one coalesced weight load feeds 32 FMAs.
__global__ void __launch_bounds__(DN_THREADS)
dense_splitk_kernel(const float *__restrict__ x, const float *__restrict__ W, float *__restrict__ y, int B, int K, int N, int k_per)
{
    __shared__ __align__(16) float xs[DN_BK][DN_BM];      // transposed activation tile: xs[k][b]
    const int n = blockIdx.x * DN_THREADS + threadIdx.x;
    const int b0 = blockIdx.z * DN_BM;
    const int k_lo = blockIdx.y * k_per, k_hi = min(K, k_lo + k_per);
    float acc[DN_BM];
#pragma unroll
    for (int b = 0; b < DN_BM; ++b) acc[b] = 0.f;
    for (int k0 = k_lo; k0 < k_hi; k0 += DN_BK) {
        __syncthreads();
        for (int e = threadIdx.x; e < DN_BK * DN_BM; e += DN_THREADS) {
            const int kk = e % DN_BK, b = e / DN_BK;           // consecutive threads -> consecutive k: coalesced rows of x
            xs[kk][b] = (b0 + b < B && k0 + kk < k_hi) ? __ldg(x + (size_t)(b0 + b) * K + k0 + kk) : 0.f;
        }
        __syncthreads();
        if (n < N) {
            const int kn = min(DN_BK, k_hi - k0);
#pragma unroll 4
            for (int kk = 0; kk < kn; ++kk) {
                const float wv = __ldg(W + (size_t)(k0 + kk) * N + n);
                const float4 *xr = reinterpret_cast<const float4 *>(xs[kk]);
#pragma unroll
                for (int q = 0; q < DN_BM / 4; ++q) {
                    const float4 xv = xr[q];
                    acc[4 * q + 0] = fmaf(xv.x, wv, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(xv.y, wv, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(xv.z, wv, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(xv.w, wv, acc[4 * q + 3]);
                }
            }
        }
    }
    if (n < N) {
#pragma unroll
        for (int b = 0; b < DN_BM; ++b)
            if (b0 + b < B) atomicAdd(y + (size_t)(b0 + b) * N + n, acc[b]);
    }
}

// ---- weight-streaming Dense for wide layers (N % 4 == 0, K % 4 == 0, N >= 256) ---------------------------------------------------
// The layer is HBM-bound on W (fc1: 134 MB, fc2: 67 MB per pass against 32 rows of activations), so the kernel is built around
// keeping bytes in flight: one producer lane streams [32 k][256 n] weight tiles and the matching [32 b][32 k] activation tiles
// through a 4-stage shared-memory ring with two 2-D TMA tensor loads per stage (out-of-bounds rows / columns arrive as zeros, so
// ragged K, N and B need no special cases); eight consumer warps — two groups that each take half of a stage's k-rows — own two
// columns per thread and 2 x 32 fp32 accumulators: per 4 k-rows a thread issues 8 LDS.32 of weights and 32 broadcast LDS.128 of
// activations for 256 FMAs.  Split-K over blockIdx.y; partial sums meet in y with atomics.
constexpr int DS_BN = 256, DS_BK = 32, DS_STAGES = 4, DS_CONSUMERS = 256, DS_THREADS = DS_CONSUMERS + 32;
constexpr int DS_W_STAGE = DS_BK * DS_BN * 4, DS_X_STAGE = DN_BM * DS_BK * 4;
constexpr int DS_SMEM = DS_STAGES * (DS_W_STAGE + DS_X_STAGE) + 2 * DS_STAGES * 8 + 1024;

__global__ void __launch_bounds__(DS_THREADS, 1)
dense_stream_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, float *__restrict__ y, int B,
                    int K, int N, int k_per)
{
    extern __shared__ uint8_t ds_raw[];
    uint8_t *ds_smem = ds_raw + ((1024u - (smem_u32(ds_raw) & 1023u)) & 1023u);
    float *Ws = reinterpret_cast<float *>(ds_smem);                                   // [STAGES][BK][BN]
    float *xs = reinterpret_cast<float *>(ds_smem + DS_STAGES * DS_W_STAGE);          // [STAGES][32 b][BK]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ds_smem + DS_STAGES * (DS_W_STAGE + DS_X_STAGE));
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + DS_STAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * DS_BN, b0 = blockIdx.z * DN_BM;
    const int k_lo = blockIdx.y * k_per, k_hi = min(K, k_lo + k_per);
    const int nsteps = (k_hi - k_lo + DS_BK - 1) / DS_BK;
    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);                        // the producer's arrive.expect_tx
            mbar_init(empty0 + 8 * s, DS_CONSUMERS / 32);       // one arrival per consumer warp
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == DS_CONSUMERS / 32) {
        if (lane == 0) {
            prefetch_tmap(&tmap_w);
            prefetch_tmap(&tmap_x);
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % DS_STAGES, k0 = k_lo + it * DS_BK;
                mbar_wait(empty0 + 8 * s, ((it / DS_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(DS_W_STAGE + DS_X_STAGE));
                tma_load_2d(smem_u32(Ws + s * DS_BK * DS_BN), &tmap_w, n0, k0, full0 + 8 * s);
                tma_load_2d(smem_u32(xs + s * DN_BM * DS_BK), &tmap_x, k0, b0, full0 + 8 * s);
            }
        }
    } else {
        // ---------------- consumers: columns n0 + ct and n0 + 128 + ct, k-rows [16 * khalf, 16 * khalf + 16) of every stage ----------------
        const int ct = tid & 127, khalf = tid >> 7;
        float acc0[DN_BM], acc1[DN_BM];
#pragma unroll
        for (int b = 0; b < DN_BM; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % DS_STAGES;
            mbar_wait(full0 + 8 * s, (it / DS_STAGES) & 1);
            const float *wt = Ws + s * DS_BK * DS_BN + ct, *xt = xs + s * DN_BM * DS_BK;
#pragma unroll 1
            for (int k4 = khalf * (DS_BK / 2); k4 < (khalf + 1) * (DS_BK / 2); k4 += 4) {
                float w0[4], w1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { w0[u] = wt[(k4 + u) * DS_BN]; w1[u] = wt[(k4 + u) * DS_BN + 128]; }
#pragma unroll
                for (int b = 0; b < DN_BM; ++b) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xt + b * DS_BK + k4);
                    acc0[b] = fmaf(xv.x, w0[0], acc0[b]); acc1[b] = fmaf(xv.x, w1[0], acc1[b]);
                    acc0[b] = fmaf(xv.y, w0[1], acc0[b]); acc1[b] = fmaf(xv.y, w1[1], acc1[b]);
                    acc0[b] = fmaf(xv.z, w0[2], acc0[b]); acc1[b] = fmaf(xv.z, w1[2], acc1[b]);
                    acc0[b] = fmaf(xv.w, w0[3], acc0[b]); acc1[b] = fmaf(xv.w, w1[3], acc1[b]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }
        const int na = n0 + ct, nb = n0 + 128 + ct;
#pragma unroll
        for (int b = 0; b < DN_BM; ++b)
            if (b0 + b < B) {
                if (na < N) atomicAdd(y + (size_t)(b0 + b) * N + na, acc0[b]);
                if (nb < N) atomicAdd(y + (size_t)(b0 + b) * N + nb, acc1[b]);
            }
    }
}

// The data gradient of the same layers without a transposed copy of W: dx [B, K] = dy [B, N] . W^T with W read in its own [K, N]
// layout.  Same ring and the same FMA structure with the roles exchanged — a thread owns two k-ROWS of the [256 k][32 n] weight tile
// (TMA, 128-byte swizzle: the 128-bit reads of 32 threads on 32 different rows spread over all banks), the dy tile [32 b][32 n] is
// broadcast; the contraction runs over n (split over blockIdx.y).
constexpr int DT_BKR = 256, DT_BN = 32;
constexpr int DT_W_STAGE = DT_BKR * DT_BN * 4, DT_X_STAGE = DN_BM * DT_BN * 4;
constexpr int DT_SMEM = DS_STAGES * (DT_W_STAGE + DT_X_STAGE) + 2 * DS_STAGES * 8 + 1024;

__global__ void __launch_bounds__(DS_THREADS, 1)
dense_stream_t_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_dy, float *__restrict__ dx, int B,
                      int K, int N, int n_per)
{
    extern __shared__ uint8_t ds_raw[];
    uint8_t *ds_smem = ds_raw + ((1024u - (smem_u32(ds_raw) & 1023u)) & 1023u);
    uint8_t *Ws = ds_smem;                                                            // [STAGES][256 k][32 n], swizzled
    float *xs = reinterpret_cast<float *>(ds_smem + DS_STAGES * DT_W_STAGE);          // [STAGES][32 b][32 n]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ds_smem + DS_STAGES * (DT_W_STAGE + DT_X_STAGE));
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + DS_STAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k0 = blockIdx.x * DT_BKR, b0 = blockIdx.z * DN_BM;
    const int n_lo = blockIdx.y * n_per, n_hi = min(N, n_lo + n_per);
    const int nsteps = (n_hi - n_lo + DT_BN - 1) / DT_BN;
    if (tid == 0) {
        for (int s = 0; s < DS_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, DS_CONSUMERS / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == DS_CONSUMERS / 32) {
        if (lane == 0) {
            prefetch_tmap(&tmap_w);
            prefetch_tmap(&tmap_dy);
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % DS_STAGES, n0 = n_lo + it * DT_BN;
                mbar_wait(empty0 + 8 * s, ((it / DS_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(full0 + 8 * s, (uint32_t)(DT_W_STAGE + DT_X_STAGE));
                tma_load_2d(smem_u32(Ws + s * DT_W_STAGE), &tmap_w, n0, k0, full0 + 8 * s);
                tma_load_2d(smem_u32(xs + s * DN_BM * DT_BN), &tmap_dy, n0, b0, full0 + 8 * s);
            }
        }
    } else {
        // consumers: k-rows k0 + ct and k0 + 128 + ct, n-columns [16 * nhalf, 16 * nhalf + 16) of every stage
        const int ct = tid & 127, nhalf = tid >> 7;
        float acc0[DN_BM], acc1[DN_BM];
#pragma unroll
        for (int b = 0; b < DN_BM; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % DS_STAGES;
            mbar_wait(full0 + 8 * s, (it / DS_STAGES) & 1);
            const uint8_t *wt = Ws + s * DT_W_STAGE;
            const float *xt = xs + s * DN_BM * DT_BN;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const int c = 4 * nhalf + q;                                          // 16-byte chunk of the 128-byte tile rows
                const float4 w0 = *reinterpret_cast<const float4 *>(wt + ct * 128 + ((c ^ (ct & 7)) << 4));
                const float4 w1 = *reinterpret_cast<const float4 *>(wt + (ct + 128) * 128 + ((c ^ (ct & 7)) << 4));
#pragma unroll
                for (int b = 0; b < DN_BM; ++b) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xt + b * DT_BN + 4 * c);
                    acc0[b] = fmaf(xv.x, w0.x, acc0[b]); acc1[b] = fmaf(xv.x, w1.x, acc1[b]);
                    acc0[b] = fmaf(xv.y, w0.y, acc0[b]); acc1[b] = fmaf(xv.y, w1.y, acc1[b]);
                    acc0[b] = fmaf(xv.z, w0.z, acc0[b]); acc1[b] = fmaf(xv.z, w1.z, acc1[b]);
                    acc0[b] = fmaf(xv.w, w0.w, acc0[b]); acc1[b] = fmaf(xv.w, w1.w, acc1[b]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }
        const int ka = k0 + ct, kb = k0 + 128 + ct;
#pragma unroll
        for (int b = 0; b < DN_BM; ++b)
            if (b0 + b < B) {
                if (ka < K) atomicAdd(dx + (size_t)(b0 + b) * K + ka, acc0[b]);
                if (kb < K) atomicAdd(dx + (size_t)(b0 + b) * K + kb, acc1[b]);
            }
    }
}

// y = act(y + bias)
__global__ void dense_finalize_kernel(float *__restrict__ y, const float *__restrict__ bias, long total, int N, int relu)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        float v = y[e] + (bias ? __ldg(bias + (int)(e % N)) : 0.f);
        y[e] = relu ? fmaxf(v, 0.f) : v;
    }
}

// row softmax; one CTA per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float *__restrict__ x, float *__restrict__ y, int N)
{
    __shared__ float red[8];
    const float *row = x + (size_t)blockIdx.x * N;
    float *out = y + (size_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += 256) m = fmaxf(m, row[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += 256) s += expf(row[i] - m);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    const float inv = 1.f / s;
    for (int i = threadIdx.x; i < N; i += 256) out[i] = expf(row[i] - m) * inv;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_maxpool2x2_fwd(const float *x, float *y, int B, int h, int w, int C, void *stream)
{
    SKY_REQUIRE(x && y && B > 0 && h > 0 && w > 0 && C > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(C % 4 == 0, SKY_ERR_UNSUPPORTED, "max-pool kernel needs C %% 4 == 0 (got %d)", C);
    const int oh = (h + 1) / 2, ow = (w + 1) / 2;
    const long total = (long)B * oh * ow * (C / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool2x2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, B, h, w, C, oh, ow);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

static int dense_finish(float *y, const float *bias, int B, int N, int relu, cudaStream_t st)
{
    if (bias || relu) {      // bias == NULL: the plain product (sky_dense_bwd_data)
        const long total = (long)B * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        dense_finalize_kernel<<<blocks, 256, 0, st>>>(y, bias, total, N, relu);
        SKY_CHECK_LAUNCH();
    }
    return SKY_OK;
}

extern "C" int sky_dense_fwd(const float *x, const float *W, const float *bias, float *y, int B, int K, int N, int relu, void *stream)
{
    SKY_REQUIRE(x && W && y && B > 0 && K > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(y, 0, (size_t)B * N * sizeof(float), st));
    const int bcta = (B + DN_BM - 1) / DN_BM;
    if (N % 4 == 0 && K % 4 == 0 && N >= DS_BN && ((uintptr_t)W & 15) == 0 && ((uintptr_t)x & 15) == 0) {
        SKY_ENSURE_DYN_SMEM(dense_stream_kernel, DS_SMEM);
        CUtensorMap tmap_w, tmap_x;
        int rc = encode_2d_tensor_map(&tmap_w, W, K, N, DS_BN, DS_BK);
        if (rc != SKY_OK) return rc;
        rc = encode_2d_tensor_map(&tmap_x, x, B, K, DS_BK, DN_BM);
        if (rc != SKY_OK) return rc;
        const int ncta = (N + DS_BN - 1) / DS_BN;
        int ksplit = 148 / (ncta * bcta);                                  // one wave of CTAs, one per SM
        if (ksplit < 1) ksplit = 1;
        int k_per = (K + ksplit - 1) / ksplit;
        k_per = (k_per + DS_BK - 1) / DS_BK * DS_BK;
        ksplit = (K + k_per - 1) / k_per;
        dense_stream_kernel<<<dim3(ncta, ksplit, bcta), DS_THREADS, DS_SMEM, st>>>(tmap_w, tmap_x, y, B, K, N, k_per);
        SKY_CHECK_LAUNCH();
        return dense_finish(y, bias, B, N, relu, st);
    }
    const int ncta = (N + DN_THREADS - 1) / DN_THREADS;
    int ksplit = (4 * 148 + ncta * bcta - 1) / (ncta * bcta);          // about four waves of CTAs
    int k_per = (K + ksplit - 1) / ksplit;
    k_per = (k_per + DN_BK - 1) / DN_BK * DN_BK;
    ksplit = (K + k_per - 1) / k_per;
    dense_splitk_kernel<<<dim3(ncta, ksplit, bcta), DN_THREADS, 0, st>>>(x, W, y, B, K, N, k_per);
    SKY_CHECK_LAUNCH();
    return dense_finish(y, bias, B, N, relu, st);
}

namespace sky {
// dx = 0 where the ReLU output that fed the layer is not positive
__global__ void dense_relu_mask_kernel(float *__restrict__ dx, const float *__restrict__ act, long total)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
        if (!(act[e] > 0.f)) dx[e] = 0.f;
}
}  // namespace sky

// dx [B, K] = dy [B, N] . W^T from W in its own [K, N] layout (Keras Dense kernel): no transposed copy.  dx is overwritten and zeroed
// where `act` (the ReLU output that fed the layer, may be NULL) is not positive.
// Needs N % 4 == 0 and K >= 256; SKY_ERR_UNSUPPORTED otherwise (the caller then streams a transposed copy through sky_dense_bwd_data).
extern "C" int sky_dense_bwd_data_nt(const float *dy, const float *W, const float *act, float *dx, int B, int K, int N, void *stream)
{
    SKY_REQUIRE(dy && W && dx && B > 0 && K > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    SKY_REQUIRE(N % 4 == 0 && K >= DT_BKR && ((uintptr_t)W & 15) == 0 && ((uintptr_t)dy & 15) == 0, SKY_ERR_UNSUPPORTED,
                "the transposed weight-streaming kernel needs N %% 4 == 0, K >= 256 and 16-byte aligned operands (K=%d N=%d)", K, N);
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * K * sizeof(float), st));
    SKY_ENSURE_DYN_SMEM(dense_stream_t_kernel, DT_SMEM);
    CUtensorMap tmap_w, tmap_dy;
    int rc = encode_2d_tensor_map_sw(&tmap_w, W, K, N, DT_BN, DT_BKR, 1);
    if (rc != SKY_OK) return rc;
    rc = encode_2d_tensor_map(&tmap_dy, dy, B, N, DT_BN, DN_BM);
    if (rc != SKY_OK) return rc;
    const int kcta = (K + DT_BKR - 1) / DT_BKR, bcta = (B + DN_BM - 1) / DN_BM;
    int nsplit = 148 / (kcta * bcta);                                   // one wave of CTAs, one per SM
    if (nsplit < 1) nsplit = 1;
    int n_per = (N + nsplit - 1) / nsplit;
    n_per = (n_per + DT_BN - 1) / DT_BN * DT_BN;
    nsplit = (N + n_per - 1) / n_per;
    dense_stream_t_kernel<<<dim3(kcta, nsplit, bcta), DS_THREADS, DT_SMEM, st>>>(tmap_w, tmap_dy, dx, B, K, N, n_per);
    SKY_CHECK_LAUNCH();
    if (act) {
        const long total = (long)B * K;
        long blocks = (total + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        dense_relu_mask_kernel<<<(int)blocks, 256, 0, st>>>(dx, act, total);
        SKY_CHECK_LAUNCH();
    }
    return SKY_OK;
}

extern "C" int sky_softmax_rows(const float *x, float *y, int rows, int N, void *stream)
{
    SKY_REQUIRE(x && y && rows > 0 && N > 0, SKY_ERR_INVALID, "bad arguments");
    softmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, y, N);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
