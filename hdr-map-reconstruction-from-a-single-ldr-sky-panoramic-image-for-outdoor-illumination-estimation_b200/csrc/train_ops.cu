// Training-side elementwise / reduction kernels of the residual trunk:
//   instance-norm backward (tfa.layers.InstanceNormalization, generator.py:15,19) fused with the LeakyReLU backward
//     (generator.py:30) that precedes it in the backward pass and with the residual gradient add (generator.py:35),
//   the synthetic L2 objective used by the trunk train bench, and Keras RMSprop (train.py:201-202).
// All HBM-bound: each distinct tensor is read or written once per kernel.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "sky_common.cuh"

namespace sky {

constexpr int TR_THREADS = 256;

__device__ __forceinline__ void in_scale_shift(const double *stats, int b, int f, int F, int hw, float eps, float *mean, float *rstd)
{
    const double s1 = stats[((size_t)b * F + f) * 2 + 0], s2 = stats[((size_t)b * F + f) * 2 + 1];
    const double m = s1 / hw;
    double var = s2 / hw - m * m;
    var = var < 0.0 ? 0.0 : var;
    *mean = (float)m;
    *rstd = rsqrtf((float)var + eps);
}

// Pass 1: per (sample, channel) sums of dy' and dy' * xhat over the plane, where dy' = dy * lrelu'(act) if `act` is
// given (act = the activation tensor that followed the norm: its sign equals the sign of the normalised value).
// grid (pixel chunks, B); sums[b][f][2] (fp64) must be zeroed by the caller.
__global__ void __launch_bounds__(TR_THREADS)
instnorm_bwd_reduce_kernel(const float *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ dy,
                           const float *__restrict__ act, double *__restrict__ sums, int hw, int F, float eps, float slope,
                           int pix_per_cta)
{
    extern __shared__ float sm[];       // mean[F], rstd[F], then partial[TR_THREADS/?]
    float *mean = sm, *rstd = sm + F;
    const int b = blockIdx.y;
    for (int f = threadIdx.x; f < F; f += TR_THREADS) in_scale_shift(stats, b, f, F, hw, eps, &mean[f], &rstd[f]);
    __syncthreads();
    const int f4n = F / 4;
    const int npix = min(pix_per_cta, hw - blockIdx.x * pix_per_cta);
    const size_t base = ((size_t)b * hw + (size_t)blockIdx.x * pix_per_cta) * F;
    // thread -> fixed float4 channel group (requires TR_THREADS % f4n == 0), strided over pixels
    const int c4 = threadIdx.x % f4n, prow = threadIdx.x / f4n, pstep = TR_THREADS / f4n;
    float a1[4] = { 0.f, 0.f, 0.f, 0.f }, a2[4] = { 0.f, 0.f, 0.f, 0.f };
    for (int p = prow; p < npix; p += pstep) {
        const size_t o = base + (size_t)p * F + 4 * c4;
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + o));
        float4 g = __ldg(reinterpret_cast<const float4 *>(dy + o));
        if (act) {
            const float4 av = __ldg(reinterpret_cast<const float4 *>(act + o));
            g.x *= av.x > 0.f ? 1.f : slope; g.y *= av.y > 0.f ? 1.f : slope;
            g.z *= av.z > 0.f ? 1.f : slope; g.w *= av.w > 0.f ? 1.f : slope;
        }
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w }, gs[4] = { g.x, g.y, g.z, g.w };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - mean[4 * c4 + u]) * rstd[4 * c4 + u];
            a1[u] += gs[u];
            a2[u] = fmaf(gs[u], xh, a2[u]);
        }
    }
    // combine the pstep threads that share c4 through shared memory, one fp64 atomic per (block, channel, sum)
    float *part = sm + 2 * F;           // [TR_THREADS][8]
#pragma unroll
    for (int u = 0; u < 4; ++u) { part[threadIdx.x * 8 + u] = a1[u]; part[threadIdx.x * 8 + 4 + u] = a2[u]; }
    __syncthreads();
    if (threadIdx.x < 2 * F) {
        const int f = threadIdx.x >> 1, which = threadIdx.x & 1;
        const int cc4 = f >> 2, u = f & 3;
        double s = 0.0;
        for (int r = 0; r < pstep; ++r) s += (double)part[(r * f4n + cc4) * 8 + which * 4 + u];
        atomicAdd(sums + ((size_t)b * F + f) * 2 + which, s);
    }
}

// Pass 2: dx = gamma * rstd * (dy' - mean(dy') - xhat * mean(dy' * xhat)) (+ extra), and the parameter gradients
// dgamma[f] += sum_b sum(dy' * xhat), dbeta[f] += sum_b sum(dy') (added once per sample by chunk 0).
__global__ void __launch_bounds__(TR_THREADS)
instnorm_bwd_apply_kernel(const float *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ gamma,
                          const float *__restrict__ dy, const float *__restrict__ act, const double *__restrict__ sums,
                          const float *__restrict__ extra, float *__restrict__ dx, float *__restrict__ dgamma,
                          float *__restrict__ dbeta, int hw, int F, float eps, float slope, int pix_per_cta)
{
    extern __shared__ float sm[];       // mean, rstd, k1 = gamma*rstd, m1 = mean(dy'), m2 = mean(dy'*xhat)
    float *mean = sm, *rstd = sm + F, *k1 = sm + 2 * F, *m1 = sm + 3 * F, *m2 = sm + 4 * F;
    const int b = blockIdx.y;
    for (int f = threadIdx.x; f < F; f += TR_THREADS) {
        in_scale_shift(stats, b, f, F, hw, eps, &mean[f], &rstd[f]);
        k1[f] = gamma[f] * rstd[f];
        const double s1 = sums[((size_t)b * F + f) * 2 + 0], s2 = sums[((size_t)b * F + f) * 2 + 1];
        m1[f] = (float)(s1 / hw);
        m2[f] = (float)(s2 / hw);
        if (blockIdx.x == 0) {
            atomicAdd(dbeta + f, (float)s1);
            atomicAdd(dgamma + f, (float)s2);
        }
    }
    __syncthreads();
    const size_t base = ((size_t)b * hw + (size_t)blockIdx.x * pix_per_cta) * F;
    const int npix = min(pix_per_cta, hw - blockIdx.x * pix_per_cta);
    const int total4 = npix * F / 4, f4n = F / 4;
    for (int e = threadIdx.x; e < total4; e += TR_THREADS) {
        const int f = (e % f4n) * 4;
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + base) + e);
        float4 g = __ldg(reinterpret_cast<const float4 *>(dy + base) + e);
        if (act) {
            const float4 av = __ldg(reinterpret_cast<const float4 *>(act + base) + e);
            g.x *= av.x > 0.f ? 1.f : slope; g.y *= av.y > 0.f ? 1.f : slope;
            g.z *= av.z > 0.f ? 1.f : slope; g.w *= av.w > 0.f ? 1.f : slope;
        }
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w }, gs[4] = { g.x, g.y, g.z, g.w };
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - mean[f + u]) * rstd[f + u];
            o[u] = k1[f + u] * (gs[u] - m1[f + u] - xh * m2[f + u]);
        }
        if (extra) {
            const float4 ev = __ldg(reinterpret_cast<const float4 *>(extra + base) + e);
            o[0] += ev.x; o[1] += ev.y; o[2] += ev.z; o[3] += ev.w;
        }
        reinterpret_cast<float4 *>(dx + base)[e] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// One-launch version: the CTAs that share a sample form a thread-block cluster; each reduces its pixel chunk, the per-CTA partial sums
// meet through distributed shared memory (no global atomics, no memset, no second launch), then every CTA applies the gradient to its
// chunk — from registers when the chunk is small enough to have been kept there (the 8x32 trunk planes), else re-read (L2).
// grid (CL, B), cluster (CL, 1, 1); CL * pix_per_cta >= hw.  Small chunks: 256 threads with the chunk in registers; large chunks: 1024
// threads streaming twice (the stage is bound by the loads in flight).
constexpr int IN_CACHE = 4;             // pixel iterations per thread kept in registers
template <bool CACHE, int THREADS>
__global__ void __launch_bounds__(THREADS)
instnorm_bwd_cluster_kernel(const float *__restrict__ x, const double *__restrict__ stats, const float *__restrict__ gamma,
                            const float *__restrict__ dy, const float *__restrict__ act, const float *__restrict__ extra,
                            float *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta, int hw, int F, float eps,
                            float slope, int pix_per_cta)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float sm[];       // mean[F], rstd[F], k1[F], m1[F], m2[F], csum[2F] (this CTA's partial sums), part[THREADS][8]
    float *mean = sm, *rstd = sm + F, *k1 = sm + 2 * F, *m1 = sm + 3 * F, *m2 = sm + 4 * F, *csum = sm + 5 * F, *part = sm + 7 * F;
    const int b = blockIdx.y, rank = (int)cluster.block_rank(), CL = (int)cluster.num_blocks();
    for (int f = threadIdx.x; f < F; f += THREADS) {
        in_scale_shift(stats, b, f, F, hw, eps, &mean[f], &rstd[f]);
        k1[f] = gamma[f] * rstd[f];
    }
    __syncthreads();
    const int f4n = F / 4;
    const int npix = max(0, min(pix_per_cta, hw - rank * pix_per_cta));
    const size_t base = ((size_t)b * hw + (size_t)rank * pix_per_cta) * F;
    const int c4 = threadIdx.x % f4n, prow = threadIdx.x / f4n, pstep = THREADS / f4n;
    float a1[4] = { 0.f, 0.f, 0.f, 0.f }, a2[4] = { 0.f, 0.f, 0.f, 0.f };
    float4 xc[CACHE ? IN_CACHE : 1], gc[CACHE ? IN_CACHE : 1];
    auto load = [&](int p, float4 &xv, float4 &g) {
        const size_t o = base + (size_t)p * F + 4 * c4;
        xv = __ldg(reinterpret_cast<const float4 *>(x + o));
        g = __ldg(reinterpret_cast<const float4 *>(dy + o));
        if (act) {
            const float4 av = __ldg(reinterpret_cast<const float4 *>(act + o));
            g.x *= av.x > 0.f ? 1.f : slope; g.y *= av.y > 0.f ? 1.f : slope;
            g.z *= av.z > 0.f ? 1.f : slope; g.w *= av.w > 0.f ? 1.f : slope;
        }
    };
    auto accumulate = [&](const float4 &xv, const float4 &g) {
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w }, gs[4] = { g.x, g.y, g.z, g.w };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - mean[4 * c4 + u]) * rstd[4 * c4 + u];
            a1[u] += gs[u];
            a2[u] = fmaf(gs[u], xh, a2[u]);
        }
    };
    if (CACHE) {
#pragma unroll
        for (int it = 0; it < IN_CACHE; ++it) {
            const int p = prow + it * pstep;
            xc[it] = make_float4(0.f, 0.f, 0.f, 0.f); gc[it] = xc[it];
            if (p < npix) load(p, xc[it], gc[it]);
        }
#pragma unroll
        for (int it = 0; it < IN_CACHE; ++it)
            if (prow + it * pstep < npix) accumulate(xc[it], gc[it]);
    } else {
        for (int p = prow; p < npix; p += pstep) {
            float4 xv, g;
            load(p, xv, g);
            accumulate(xv, g);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { part[threadIdx.x * 8 + u] = a1[u]; part[threadIdx.x * 8 + 4 + u] = a2[u]; }
    __syncthreads();
    if (threadIdx.x < 2 * F) {
        const int f = threadIdx.x >> 1, which = threadIdx.x & 1;
        const int cc4 = f >> 2, u = f & 3;
        float s = 0.f;
        for (int r = 0; r < pstep; ++r) s += part[(r * f4n + cc4) * 8 + which * 4 + u];
        csum[threadIdx.x] = s;                                   // [f][which]
    }
    cluster.sync();
    if (threadIdx.x < 2 * F) {
        double s = 0.0;
        for (int r = 0; r < CL; ++r) s += (double)cluster.map_shared_rank(csum, r)[threadIdx.x];
        const int f = threadIdx.x >> 1, which = threadIdx.x & 1;
        (which ? m2 : m1)[f] = (float)(s / hw);
        if (rank == 0) atomicAdd((which ? dgamma : dbeta) + f, (float)s);
    }
    cluster.sync();                                              // remote reads are done before any CTA of the cluster may exit
    auto apply = [&](int p, const float4 &xv, const float4 &g) {
        const int f = 4 * c4;
        const float xs[4] = { xv.x, xv.y, xv.z, xv.w }, gs[4] = { g.x, g.y, g.z, g.w };
        float o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float xh = (xs[u] - mean[f + u]) * rstd[f + u];
            o[u] = k1[f + u] * (gs[u] - m1[f + u] - xh * m2[f + u]);
        }
        const size_t off = base + (size_t)p * F + f;
        if (extra) {
            const float4 ev = __ldg(reinterpret_cast<const float4 *>(extra + off));
            o[0] += ev.x; o[1] += ev.y; o[2] += ev.z; o[3] += ev.w;
        }
        *reinterpret_cast<float4 *>(dx + off) = make_float4(o[0], o[1], o[2], o[3]);
    };
    if (CACHE) {
#pragma unroll
        for (int it = 0; it < IN_CACHE; ++it)
            if (prow + it * pstep < npix) apply(prow + it * pstep, xc[it], gc[it]);
    } else {
        for (int p = prow; p < npix; p += pstep) {
            float4 xv, g;
            load(p, xv, g);
            apply(p, xv, g);
        }
    }
}

// loss = mean((y - target)^2); dy = 2 (y - target) / n.  loss (fp64) must be zeroed by the caller.
__global__ void mse_loss_kernel(const float *__restrict__ y, const float *__restrict__ target, float *__restrict__ dy,
                                double *__restrict__ loss, long n4, float inv_n)
{
    float acc = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(y) + e), t = __ldg(reinterpret_cast<const float4 *>(target) + e);
        const float4 d = make_float4(a.x - t.x, a.y - t.y, a.z - t.z, a.w - t.w);
        acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        reinterpret_cast<float4 *>(dy)[e] = make_float4(2.f * d.x * inv_n, 2.f * d.y * inv_n, 2.f * d.z * inv_n, 2.f * d.w * inv_n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, (double)acc * inv_n);
}

// Keras RMSprop (train.py:201-202; rho 0.9, epsilon 1e-7): ms = rho*ms + (1-rho)*g^2 ; w -= lr * g / (sqrt(ms) + eps).
// grad_scale folds the 1/world_size of the data-parallel gradient average into the update.
__global__ void rmsprop_kernel(float *__restrict__ w, float *__restrict__ ms, const float *__restrict__ g, long n, float lr,
                               float rho, float eps, float grad_scale)
{
    // HBM-bound (20 bytes per variable): 128-bit accesses over the 16-byte aligned body, scalar tail
    const long n4 = (((uintptr_t)w | (uintptr_t)ms | (uintptr_t)g) & 15) == 0 ? n / 4 : 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
        const float4 gq = __ldg(reinterpret_cast<const float4 *>(g) + e);
        float4 mq = reinterpret_cast<float4 *>(ms)[e], wq = reinterpret_cast<float4 *>(w)[e];
        const float gv[4] = { gq.x * grad_scale, gq.y * grad_scale, gq.z * grad_scale, gq.w * grad_scale };
        float mv[4] = { mq.x, mq.y, mq.z, mq.w }, wv[4] = { wq.x, wq.y, wq.z, wq.w };
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            mv[u] = rho * mv[u] + (1.f - rho) * gv[u] * gv[u];
            wv[u] -= lr * gv[u] / (sqrtf(mv[u]) + eps);
        }
        reinterpret_cast<float4 *>(ms)[e] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        reinterpret_cast<float4 *>(w)[e] = make_float4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (long e = 4 * n4 + blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const float gv = g[e] * grad_scale;
        const float m = rho * ms[e] + (1.f - rho) * gv * gv;
        ms[e] = m;
        w[e] -= lr * gv / (sqrtf(m) + eps);
    }
}

static void in_grid(int B, int hw, int *chunks, int *pix)
{
    int c = (2 * 148 + B - 1) / B;
    int p = (hw + c - 1) / c;
    if (p < 16) p = 16;
    *pix = p;
    *chunks = (hw + p - 1) / p;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_instnorm_bwd(const float *x, const double *stats, const float *gamma, const float *dy, const float *act,
                                const float *extra, double *sums, float *dx, float *dgamma, float *dbeta, int B, int h, int w,
                                int F, float eps, float slope, void *stream)
{
    SKY_REQUIRE(x && stats && gamma && dy && sums && dx && dgamma && dbeta, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(F % 4 == 0 && TR_THREADS % (F / 4) == 0 && 2 * F <= TR_THREADS, SKY_ERR_UNSUPPORTED,
                "instance-norm backward needs filters in {4, 8, 16, 32, 64, 128} (got %d)", F);
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w;
    if (!getenv("SKY_INSTNORM_BWD_TWO_PASS")) {
        // one launch: a cluster of up to 8 CTAs per sample, partial sums through distributed shared memory (`sums` is not used)
        const int pstep = TR_THREADS / (F / 4);
        int CL = 8;
        while (CL > 1 && hw < CL * pstep) CL >>= 1;
        const int pix = (hw + CL - 1) / CL;
        const bool cache = pix <= IN_CACHE * pstep;
        const int threads = cache ? TR_THREADS : 1024;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL, B); cfg.blockDim = dim3(threads); cfg.stream = st;
        cfg.dynamicSmemBytes = (7 * F + threads * 8) * sizeof(float);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cache)
            SKY_CHECK_CUDA(cudaLaunchKernelEx(&cfg, instnorm_bwd_cluster_kernel<true, TR_THREADS>, x, stats, gamma, dy, act, extra, dx, dgamma, dbeta, hw, F,
                                              eps, slope, pix));
        else
            SKY_CHECK_CUDA(cudaLaunchKernelEx(&cfg, instnorm_bwd_cluster_kernel<false, 1024>, x, stats, gamma, dy, act, extra, dx, dgamma, dbeta, hw, F,
                                              eps, slope, pix));
        SKY_CHECK_LAUNCH();
        return SKY_OK;
    }
    int chunks, pix;
    in_grid(B, hw, &chunks, &pix);
    SKY_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * F * 2 * sizeof(double), st));
    dim3 grid(chunks, B);
    instnorm_bwd_reduce_kernel<<<grid, TR_THREADS, (2 * F + TR_THREADS * 8) * sizeof(float), st>>>(x, stats, dy, act, sums, hw, F, eps,
                                                                                                  slope, pix);
    SKY_CHECK_LAUNCH();
    instnorm_bwd_apply_kernel<<<grid, TR_THREADS, 5 * F * sizeof(float), st>>>(x, stats, gamma, dy, act, sums, extra, dx, dgamma, dbeta,
                                                                               hw, F, eps, slope, pix);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_mse_loss(const float *y, const float *target, float *dy, double *loss, long n, void *stream)
{
    SKY_REQUIRE(y && target && dy && loss && n > 0 && n % 4 == 0, SKY_ERR_INVALID, "bad arguments (n must be a positive multiple of 4)");
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
    long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    mse_loss_kernel<<<(int)blocks, 256, 0, st>>>(y, target, dy, loss, n / 4, 1.0f / (float)n);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_rmsprop_step(float *w, float *ms, const float *g, long n, float lr, float rho, float eps, float grad_scale,
                                void *stream)
{
    SKY_REQUIRE(w && ms && g && n > 0, SKY_ERR_INVALID, "bad arguments");
    long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rmsprop_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(w, ms, g, n, lr, rho, eps, grad_scale);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
