// Backward of the plain convolutions around the distortion-aware trunk (ops.conv2d / ops.deconv2d / Keras Conv2D of the encoder,
// decoders, sunRadNet, discriminator and VGG16: what TF autodiff derives from tf.nn.conv2d, ops.py:41, sunrad_net.py:12), and the
// pipelined weight-gradient kernel shared with the distortion-aware layers.
//
//   data gradient    dX = conv_transpose(dY, W).  Run as a FORWARD pass over dY with the flipped, transposed kernel
//                    (sky_conv2d_transpose_weights -> sky_da_pack_weights): stride-1 odd-k layers are then literally a SAME conv
//                    (band-staged / small-filter / direct tcgen05 kernels, epilogues included); strided and even-k layers use the
//                    direct kernel's transposed sampler (taps that do not land on a dY pixel contribute zero rows).  A gather, no atomics.
//                    The activation gradient of the layer below rides in the epilogue (SKY_EPI_MASK).
//   weight gradient  dW[(t,c), f] = sum_m Pix[m,(t,c)] * dY[m,f], Pix re-gathered on the fly (im2col never stored).  Warp-specialised:
//                    8 producer warps build MN-major TF32 operand tiles for 64 pixels per stage (A = 128 kernel rows of Pix, B = the dY
//                    tile) into a 2-3 stage mbarrier ring, one thread issues tcgen05.mma (M = 128 kernel rows, N = filters, K = 8
//                    pixels) into a TMEM accumulator that stays resident over the CTA's pixel partition, 4 warps drain it with vector
//                    reductions.  Sampler: plain (stride 1 / 2, TensorFlow SAME) or the distortion-aware geometry (da_sample).
#include <stdlib.h>
#include <string.h>

#include "strip_conv.cuh"

namespace sky {

// out[(t'*F + f)*C + c] = kernel[(t*C + c)*F + f], t = flip ? k*k-1-t' : t'
__global__ void transpose_weights_kernel(const float *__restrict__ kernel, float *__restrict__ out, int C, int F, int k2, int flip)
{
    const long total = (long)k2 * C * F;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C), f = (int)((e / C) % F), tp = (int)(e / ((long)C * F));
        const int t = flip ? k2 - 1 - tp : tp;
        out[e] = __ldg(kernel + ((size_t)t * C + c) * F + f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WG_PX = 64;                               // pixels (contraction depth) per pipeline stage = 8 MMAs of K = 8
constexpr int WG_PROD_WARPS = 8;
constexpr int WG_PROD_THREADS = WG_PROD_WARPS * 32;
constexpr int WG_THREADS = WG_PROD_THREADS + 32;        // + the MMA warp
constexpr int WG_A_BYTES = 4 * WG_PX * 128;             // four [64 px x 32 kernel rows] MN-major sub-tiles
constexpr int WG_TAB_BYTES = WG_PX * 4 * 32;            // per-stage geometry table (CornerRef per (pixel, tap of the row tile))

struct WgParams {
    const float *x, *dy, *offsets;
    float *dw;
    int B, h, w, C, F, k, k2, K;       // x [B,h,w,C]; dy [B,oh,ow,ldF] columns f0 .. f0+F-1; K = k2*C kernel rows
    int oh, ow, M;                     // dy map, M = B*oh*ow
    int plain, stride, ph0, pw0;       // plain: tap (a, b) of output pixel (i, j) reads x[i*s + a - ph0][j*s + b - pw0]
    int in_h, in_w;                    // distortion-aware: padded frame of da_sample (ph0 / pw0 = its front pads)
    int ldF, f0, Fp;                   // Fp = F rounded up to 32 (UMMA N)
    int C_store, ld_dw;                // kernel rows with c >= C_store are channel padding and are not stored; dw row = t*C_store + c
    int ntiles, tiles_per_part, stages, stage_bytes;
    uint32_t tmem_cols;
};

__device__ __forceinline__ void wg_red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t wg_idesc(uint32_t M, uint32_t N)
{
    return umma_idesc_tf32(M, N) | (1u << 15) | (1u << 16);      // both operands MN-major
}
// MN-major TF32 operand: SWIZZLE_128B with a 32-byte base — atoms of [4 k-rows x 128 B of MN]; lbo = bytes between MN atoms (32
// elements each), sbo = bytes between k atoms (4 rows each).  (The only legal MN-major TF32 layout; see da_conv_bwd.cu.)
__device__ __forceinline__ uint64_t wg_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ uint32_t wg_sw_offset(uint32_t row, uint32_t chunk16)
{
    return row * 128u + ((((chunk16 >> 1) ^ (row & 3u)) << 5) | ((chunk16 & 1u) << 4));
}
__device__ __forceinline__ uint4 wg_tf32x4(float4 v)
{
    uint4 u;
    u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
    return u;
}

struct WgPix { int base, iy0, ix0, ok; };   // plain sampler: x offset of the sample's image, first tap position, pixel inside M

// grid: (pixel partitions, row tiles of 128 kernel rows)
__global__ void __launch_bounds__(WG_THREADS, 1) conv2d_wgrad_kernel(const WgParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.stages * p.stage_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages, acc_full = empty0 + 8 * p.stages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int part = blockIdx.x, rt = blockIdx.y;
    const int tile_lo = part * p.tiles_per_part, tile_hi = min(p.ntiles, tile_lo + p.tiles_per_part);
    const int b_bytes = (p.Fp / 32) * WG_PX * 128;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full0 + 8 * s, WG_PROD_WARPS); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == WG_PROD_WARPS) { tmem_alloc(smem_u32(tmem_slot), p.tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < WG_PROD_WARPS) {
        // ============================== PRODUCERS ==============================
        // A item: (pixel, 16-byte chunk q of the 128 kernel rows): q is fixed per thread, so its (tap, channel) is too
        const int q = tid & 31, px0 = tid >> 5;
        const int kidx = rt * 128 + q * 4;
        const bool k_ok = kidx < p.K;
        const int t = k_ok ? kidx / p.C : 0, c = k_ok ? kidx % p.C : 0;
        const int ta = t / p.k, tb = t % p.k;
        const int t_first = (rt * 128) / p.C;                // distortion-aware tables: taps t_first .. t_first + 3 (C % 32 == 0)
        const int ntaps = p.C >= 128 ? 1 : 128 / p.C;
        const int f4n = p.Fp / 4;
        const bool dy_vec = (p.ldF % 4 == 0) && (p.f0 % 4 == 0) && (p.F % 4 == 0);
        uint32_t sg = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++sg) {
            const int s = sg % p.stages;
            mbar_wait(empty0 + 8 * s, ((sg / p.stages) & 1) ^ 1);
            uint8_t *a_tile = smem + s * p.stage_bytes;
            uint8_t *b_tile = a_tile + WG_A_BYTES;
            uint8_t *tab = b_tile + b_bytes;
            const int m0 = tile * WG_PX;
            // ---- per-stage geometry ----
            if (p.plain) {
                if (tid < WG_PX) {
                    const int m = m0 + tid;
                    WgPix e;
                    e.ok = m < p.M;
                    const int mm = e.ok ? m : 0;
                    const int j = mm % p.ow, i = (mm / p.ow) % p.oh, b = mm / (p.ow * p.oh);
                    e.base = b * p.h * p.w;
                    e.iy0 = i * p.stride - p.ph0;
                    e.ix0 = j * p.stride - p.pw0;
                    reinterpret_cast<WgPix *>(tab)[tid] = e;
                }
            } else {
                if (tid < WG_PX * ntaps) {
                    const int px = tid / ntaps, tl = tid % ntaps, tt = t_first + tl, m = m0 + px;
                    CornerRef cr;
#pragma unroll
                    for (int u = 0; u < 4; ++u) { cr.off[u] = -1; cr.w[u] = 0.f; }
                    if (tt < p.k2 && m < p.M) {
                        const int j = m % p.ow, i = (m / p.ow) % p.oh, b = m / (p.ow * p.oh);
                        const float2 yx = __ldg(reinterpret_cast<const float2 *>(p.offsets) + (size_t)i * p.k2 + tt);
                        const Sample sm = da_sample(i, j, tt / p.k, tt % p.k, yx.x, yx.y, p.in_h, p.in_w);
                        cr = da_corners(sm, b, p.h, p.w, p.C, p.ph0, p.pw0);
                    }
                    reinterpret_cast<CornerRef *>(tab)[px * 4 + tl] = cr;
                }
            }
            named_bar_sync(1, WG_PROD_THREADS);
            // ---- A: Pix[pixel][kernel row], MN-major.  All gathers of a batch are issued before the first store: the stage is bound by
            //      L2 latency, so loads in flight per thread are what counts (8 for the plain sampler, 4 pixels x 4 corners otherwise) ----
            uint8_t *a_dst = a_tile + (q >> 3) * (WG_PX * 128);
            if (p.plain) {
                float4 v[WG_PX / WG_PROD_WARPS];
#pragma unroll
                for (int r = 0; r < WG_PX / WG_PROD_WARPS; ++r) {
                    const WgPix e = reinterpret_cast<const WgPix *>(tab)[px0 + WG_PROD_WARPS * r];
                    const int iy = e.iy0 + ta, ix = e.ix0 + tb;
                    v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k_ok && e.ok && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w)
                        v[r] = __ldg(reinterpret_cast<const float4 *>(p.x + ((size_t)e.base + (size_t)iy * p.w + ix) * p.C + c));
                }
#pragma unroll
                for (int r = 0; r < WG_PX / WG_PROD_WARPS; ++r)
                    *reinterpret_cast<uint4 *>(a_dst + wg_sw_offset((uint32_t)(px0 + WG_PROD_WARPS * r), (uint32_t)(q & 7))) = wg_tf32x4(v[r]);
            } else {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 pv[4][4];
                    float wv[4][4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int px = px0 + WG_PROD_WARPS * (4 * half + r);
                        const CornerRef cr = reinterpret_cast<const CornerRef *>(tab)[px * 4 + (t - t_first)];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const bool ok = k_ok && cr.off[u] >= 0;
                            wv[r][u] = ok ? cr.w[u] : 0.f;
                            pv[r][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok) pv[r][u] = __ldg(reinterpret_cast<const float4 *>(p.x + cr.off[u] + c));
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            v.x = fmaf(wv[r][u], pv[r][u].x, v.x); v.y = fmaf(wv[r][u], pv[r][u].y, v.y);
                            v.z = fmaf(wv[r][u], pv[r][u].z, v.z); v.w = fmaf(wv[r][u], pv[r][u].w, v.w);
                        }
                        *reinterpret_cast<uint4 *>(a_dst + wg_sw_offset((uint32_t)(px0 + WG_PROD_WARPS * (4 * half + r)), (uint32_t)(q & 7))) = wg_tf32x4(v);
                    }
                }
            }
            // ---- B: dY[pixel][filter], MN-major ----
#pragma unroll 4
            for (int e = tid; e < WG_PX * f4n; e += WG_PROD_THREADS) {
                const int px = e / f4n, c4 = e % f4n, m = m0 + px, f = 4 * c4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M && f < p.F) {
                    const float *src = p.dy + (size_t)m * p.ldF + p.f0 + f;
                    if (dy_vec) {
                        v = __ldg(reinterpret_cast<const float4 *>(src));
                    } else {
                        v.x = __ldg(src);
                        if (f + 1 < p.F) v.y = __ldg(src + 1);
                        if (f + 2 < p.F) v.z = __ldg(src + 2);
                        if (f + 3 < p.F) v.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<uint4 *>(b_tile + (c4 >> 3) * (WG_PX * 128) + wg_sw_offset((uint32_t)px, (uint32_t)(c4 & 7))) = wg_tf32x4(v);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
        // ============================== DRAIN (warps 0-3: TMEM lane == kernel row of the tile) ==============================
        if (warp < 4 && tile_hi > tile_lo) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const int row = warp * 32 + lane, kr = rt * 128 + row;
            const int rt_t = kr / p.C, rt_c = kr % p.C;
            const bool row_ok = kr < p.K && rt_c < p.C_store;
            float *dst = p.dw + ((size_t)rt_t * p.C_store + rt_c) * p.ld_dw + p.f0;
            const bool vec = (p.ld_dw % 4 == 0) && (p.f0 % 4 == 0) && (p.F % 4 == 0);
            for (int c0 = 0; c0 < p.Fp; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16), r);
                tmem_ld_wait();
                if (row_ok) {
                    if (vec) {
#pragma unroll
                        for (int u = 0; u < 32; u += 4)
                            if (c0 + u < p.F)
                                wg_red_add_v4(dst + c0 + u, __uint_as_float(r[u]), __uint_as_float(r[u + 1]), __uint_as_float(r[u + 2]),
                                              __uint_as_float(r[u + 3]));
                    } else {
#pragma unroll
                        for (int u = 0; u < 32; ++u)
                            if (c0 + u < p.F) atomicAdd(dst + c0 + u, __uint_as_float(r[u]));
                    }
                }
            }
            tc_fence_before();
        }
    } else {
        // ============================== MMA ISSUER ==============================
        // the whole warp walks the loop and the barrier waits; one elected lane issues (descriptor words stay in uniform registers)
        if (tile_hi > tile_lo) {
            const uint32_t idesc = wg_idesc(128, (uint32_t)p.Fp);
            uint32_t s = 0, phase = 0, sg = 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile, ++sg) {
                mbar_wait(full0 + 8 * s, phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * p.stage_bytes), b0 = a0 + WG_A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int g8 = 0; g8 < WG_PX / 8; ++g8)
                        umma_tf32(tmem_base, wg_desc(a0 + g8 * 1024, WG_PX * 128, 512), wg_desc(b0 + g8 * 1024, WG_PX * 128, 512), idesc,
                                  (sg | (uint32_t)g8) != 0);
                    umma_commit(empty0 + 8 * s);
                }
                __syncwarp();
                if (++s == (uint32_t)p.stages) { s = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == WG_PROD_WARPS) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient of the 3-filter layers (conv1_f / conv1_u, generator.py:76,85: 7x7, 32 -> 3): dW[(t,c), f] = sum_m x[m @ t, c] dy[m, f]
// 1.2 GFLOP at B = 32, all input reuse: on the tensor-core kernel its cost is the 822 MB im2col gather (0.93 ms).  Here a persistent CTA
// walks 8 x 32 pixel tiles, stages the (8 + k - 1) x (32 + k - 1) input patch and the dy tile in shared memory, and each thread keeps the
// sums of 4 channels x up to 4 taps x F filters in registers across ALL its tiles (one atomicAdd per sum at the end).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SFW_TH = 8, SFW_TW = 32, SFW_THREADS = 128, SFW_MAXT = 4;

__global__ void __launch_bounds__(SFW_THREADS) conv2d_wgrad_smallf_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                                           float *__restrict__ dw, int B, int h, int w, int C, int F, int k)
{
    extern __shared__ __align__(16) float sfw[];
    const int k2 = k * k, r = k / 2, PW = SFW_TW + k - 1, PH = SFW_TH + k - 1, PS = C + 4;
    float *patch = sfw;                                          // [PH][PW][PS]
    float4 *dyt = reinterpret_cast<float4 *>(sfw + (size_t)PH * PW * PS);   // [SFW_TH * SFW_TW] (f padded to 4)
    const int tid = threadIdx.x, c4n = C / 4, tgn = SFW_THREADS / c4n;
    const int cg = tid % c4n, tg = tid / c4n;                    // 4-channel group, tap group: taps tg, tg + tgn, ...
    const int tiles_x = (w + SFW_TW - 1) / SFW_TW, tiles_y = (h + SFW_TH - 1) / SFW_TH, ntiles = tiles_x * tiles_y * B;
    float acc[SFW_MAXT][4][4];
#pragma unroll
    for (int a = 0; a < SFW_MAXT; ++a)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[a][u][f] = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y), trem = tile % (tiles_x * tiles_y);
        const int i0 = (trem / tiles_x) * SFW_TH, j0 = (trem % tiles_x) * SFW_TW;
        __syncthreads();
        for (int e = tid; e < PH * PW * c4n; e += SFW_THREADS) {
            const int c4 = e % c4n, px = (e / c4n) % PW, py = e / (c4n * PW);
            const int yy = i0 + py - r, xx = j0 + px - r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(reinterpret_cast<const float4 *>(x + (((size_t)b * h + yy) * w + xx) * C) + c4);
            *reinterpret_cast<float4 *>(patch + (size_t)(py * PW + px) * PS + 4 * c4) = v;
        }
        for (int e = tid; e < SFW_TH * SFW_TW; e += SFW_THREADS) {
            const int i = i0 + e / SFW_TW, j = j0 + e % SFW_TW;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < h && j < w) {
                const float *src = dy + (((size_t)b * h + i) * w + j) * F;
                v.x = __ldg(src);
                if (F > 1) v.y = __ldg(src + 1);
                if (F > 2) v.z = __ldg(src + 2);
                if (F > 3) v.w = __ldg(src + 3);
            }
            dyt[e] = v;
        }
        __syncthreads();
        for (int p0 = 0; p0 < SFW_TH * SFW_TW; p0 += 8) {        // 8 pixels of one tile row: their dy stays in registers over the taps
            float4 g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = dyt[p0 + q];
            const int py = p0 / SFW_TW, px = p0 % SFW_TW;
#pragma unroll
            for (int a = 0; a < SFW_MAXT; ++a) {
                const int t = tg + a * tgn;
                if (t >= k2) break;
                const float *src = patch + (size_t)((py + t / k) * PW + px + t % k) * PS + 4 * cg;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 xv = *reinterpret_cast<const float4 *>(src + (size_t)q * PS);
                    const float xs[4] = { xv.x, xv.y, xv.z, xv.w };
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[a][u][0] = fmaf(xs[u], g[q].x, acc[a][u][0]);
                        acc[a][u][1] = fmaf(xs[u], g[q].y, acc[a][u][1]);
                        acc[a][u][2] = fmaf(xs[u], g[q].z, acc[a][u][2]);
                        acc[a][u][3] = fmaf(xs[u], g[q].w, acc[a][u][3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < SFW_MAXT; ++a) {
        const int t = tg + a * tgn;
        if (t >= k2) break;
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if (f < F) atomicAdd(dw + ((size_t)t * C + 4 * cg + u) * F + f, acc[a][u][f]);
    }
}

// returns SKY_ERR_UNSUPPORTED (no error text) when the layer is outside what the kernel covers
static int launch_wgrad_smallf(const float *x, const float *dy, float *dw, int B, int h, int w, int C, int F, int k, cudaStream_t st)
{
    if (F > 4 || C % 4 != 0 || C > 128 || SFW_THREADS % (C / 4) != 0 || (k & 1) == 0) return SKY_ERR_UNSUPPORTED;
    const int tgn = SFW_THREADS / (C / 4);
    if (k * k > SFW_MAXT * tgn) return SKY_ERR_UNSUPPORTED;
    const size_t smem = ((size_t)(SFW_TH + k - 1) * (SFW_TW + k - 1) * (C + 4) + (size_t)SFW_TH * SFW_TW * 4) * sizeof(float);
    if (smem > 160 * 1024) return SKY_ERR_UNSUPPORTED;
    SKY_ENSURE_DYN_SMEM(conv2d_wgrad_smallf_kernel, 160 * 1024);
    const int ntiles = ((w + SFW_TW - 1) / SFW_TW) * ((h + SFW_TH - 1) / SFW_TH) * B;
    const int grid = ntiles < 2 * 148 ? ntiles : 2 * 148;
    conv2d_wgrad_smallf_kernel<<<grid, SFW_THREADS, smem, st>>>(x, dy, dw, B, h, w, C, F, k);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}


static int launch_wgrad(WgParams p, cudaStream_t st)
{
    p.Fp = round_up(p.F, 32);
    p.K = p.k2 * p.C;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < p.Fp) p.tmem_cols <<= 1;
    p.stage_bytes = WG_A_BYTES + (p.Fp / 32) * WG_PX * 128 + WG_TAB_BYTES;
    p.stages = (227 * 1024 - 2048) / p.stage_bytes;
    if (p.stages > 4) p.stages = 4;
    SKY_REQUIRE(p.stages >= 2, SKY_ERR_UNSUPPORTED, "weight gradient: %d filters do not fit two pipeline stages", p.F);
    p.ntiles = (p.M + WG_PX - 1) / WG_PX;
    const int row_tiles = (p.K + 127) / 128;
    // one CTA per SM fits (the stages take the shared memory): one wave of pixel partitions, fewer partial sums to reduce
    int parts = 148 / row_tiles;                    // floor: 17 x 9 = 153 CTAs would spill five CTAs into a second wave
    if (parts > p.ntiles) parts = p.ntiles;
    if (parts < 1) parts = 1;
    p.tiles_per_part = (p.ntiles + parts - 1) / parts;
    parts = (p.ntiles + p.tiles_per_part - 1) / p.tiles_per_part;
    const int smem = p.stages * p.stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
    SKY_ENSURE_DYN_SMEM(conv2d_wgrad_kernel, 227 * 1024);
    conv2d_wgrad_kernel<<<dim3(parts, row_tiles), WG_THREADS, smem, st>>>(p);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_conv2d_transpose_weights(const float *kernel, float *out, int C, int F, int k, int flip, void *stream)
{
    SKY_REQUIRE(kernel && out && C > 0 && F > 0 && k > 0, SKY_ERR_INVALID, "bad arguments");
    const long total = (long)k * k * C * F;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    transpose_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kernel, out, C, F, k * k, flip);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

static void same_pad(int n, int k, int s, int *out, int *front)
{
    *out = (n + s - 1) / s;
    const int total = (*out - 1) * s + k - n;
    *front = (total > 0 ? total : 0) / 2;
}

extern "C" int sky_conv2d_bwd_data(const float *dy, const void *packed_t, float *dx, const float *aux, int B, int h, int w, int C,
                                   int F, int k, int stride, int epilogue_flags, float slope, int math_mode, void *stream)
{
    const float *mask_src = aux;
    SKY_REQUIRE(dy && packed_t && dx, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(k >= 1 && k <= 15 && (stride == 1 || stride == 2), SKY_ERR_UNSUPPORTED, "kernel size %d / stride %d not supported", k, stride);
    SKY_REQUIRE(!(epilogue_flags & ~(SKY_EPI_MASK | SKY_EPI_RESIDUAL | SKY_EPI_FORCE_DIRECT)), SKY_ERR_INVALID, "the data gradient takes SKY_EPI_MASK or SKY_EPI_RESIDUAL only");
    SKY_REQUIRE((epilogue_flags & (SKY_EPI_MASK | SKY_EPI_RESIDUAL)) != (SKY_EPI_MASK | SKY_EPI_RESIDUAL), SKY_ERR_INVALID,
                "SKY_EPI_MASK and SKY_EPI_RESIDUAL share the aux tensor");
    SKY_REQUIRE(!(epilogue_flags & (SKY_EPI_MASK | SKY_EPI_RESIDUAL)) || aux, SKY_ERR_INVALID, "SKY_EPI_MASK / SKY_EPI_RESIDUAL without the aux tensor");
    int oh, ow, ph0, pw0;
    same_pad(h, k, stride, &oh, &ph0);
    same_pad(w, k, stride, &ow, &pw0);
    FwdArgs a;
    a.x = dy; a.offsets = nullptr; a.offsets_host = nullptr; a.packed = (const float *)packed_t; a.bias = nullptr;
    a.residual = mask_src; a.y = dx; a.stats = nullptr; a.B = B; a.h = oh; a.w = ow; a.C = F; a.F = C; a.k = k;
    a.flags = epilogue_flags; a.slope = slope; a.math_mode = math_mode; a.stream = (cudaStream_t)stream; a.plain_stride = stride;
    if (stride == 1 && (k & 1)) return conv2d_plain_entry(a, 1);      // a SAME conv of dy with the flipped, transposed kernel
    SKY_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)packed_t & 15) == 0, SKY_ERR_INVALID, "dy, dx and packed must be 16-byte aligned");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    a.transposed = 1; a.out_h = h; a.out_w = w; a.tp_ph0 = k - 1 - ph0; a.tp_pw0 = k - 1 - pw0;
    if (C > 256) {
        SKY_REQUIRE(C % 256 == 0, SKY_ERR_UNSUPPORTED, "data gradient: more than 256 input channels must come in multiples of 256 (got %d)", C);
        a.F = 256; a.ldF = C; a.nslices = C / 256;
    }
    if (!(epilogue_flags & SKY_EPI_FORCE_DIRECT)) {
        int rc = launch_fwd_strip_plain(a);        // row strips over dy (stride 2: one row class per output column parity)
        if (rc != SKY_ERR_UNSUPPORTED) return rc;
    }
    return launch_fwd_direct(a);
}

extern "C" int sky_conv2d_bwd_filter(const float *x, const float *dy, const float *offsets, float *dkernel, float *dbias, int B, int h, int w,
                                     int C, int C_store, int F, int k, int stride, int accumulate, void *stream)
{
    SKY_REQUIRE(x && dy && dkernel, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0 && C_store > 0 && C_store <= C, SKY_ERR_INVALID, "bad dimension");
    SKY_REQUIRE(k >= 1 && k <= 15 && (stride == 1 || stride == 2), SKY_ERR_UNSUPPORTED, "kernel size %d / stride %d not supported", k, stride);
    SKY_REQUIRE(C % 4 == 0, SKY_ERR_UNSUPPORTED, "weight gradient needs input channels %% 4 == 0 (got %d; image layers take the small-C kernel)", C);
    SKY_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0, SKY_ERR_INVALID, "x and dy must be 16-byte aligned");
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    cudaStream_t st = (cudaStream_t)stream;
    WgParams p;
    p.x = x; p.dy = dy; p.offsets = offsets; p.dw = dkernel;
    p.B = B; p.h = h; p.w = w; p.C = C; p.k = k; p.k2 = k * k; p.C_store = C_store; p.ld_dw = F; p.ldF = F;
    if (offsets) {
        SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
        SKY_REQUIRE(stride == 1 && C_store == C && (C == 32 || C == 64 || C % 128 == 0), SKY_ERR_UNSUPPORTED,
                    "distortion-aware weight gradient: stride 1 and 32, 64 or a multiple of 128 input channels (got %d)", C);
        int pht, pwt;
        pad_axis(h, k, &p.ph0, &pht);
        pad_axis(w, k, &p.pw0, &pwt);
        p.in_h = h + pht; p.in_w = w + pwt; p.plain = 0; p.stride = 1; p.oh = h; p.ow = w;
    } else {
        p.plain = 1; p.stride = stride; p.in_h = p.in_w = 0;
        same_pad(h, k, stride, &p.oh, &p.ph0);
        same_pad(w, k, stride, &p.ow, &p.pw0);
    }
    p.M = B * p.oh * p.ow;
    if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dkernel, 0, (size_t)p.k2 * C_store * F * sizeof(float), st));
    // plain layers with C % 32 == 0: the strip formulation (strip_wgrad.cu) — the input rows are staged once per pixel tile for all the
    // taps of a kernel row instead of being re-gathered per tap; SKY_WGRAD_KERNEL=gather keeps the kernels below (the cross-check)
    int rc_small = SKY_ERR_UNSUPPORTED;
    if (!offsets && C_store == C && C % 32 == 0 && !(F <= 4 && stride == 1) && !(getenv("SKY_WGRAD_KERNEL") && !strcmp(getenv("SKY_WGRAD_KERNEL"), "gather"))) {
        rc_small = launch_wgrad_strip(x, dy, nullptr, dkernel, B, h, w, C, F, k, stride, st);
        if (rc_small != SKY_OK && rc_small != SKY_ERR_UNSUPPORTED) return rc_small;
    }
    if (rc_small == SKY_ERR_UNSUPPORTED && !offsets && stride == 1 && C_store == C) rc_small = launch_wgrad_smallf(x, dy, dkernel, B, h, w, C, F, k, st);   // conv1_f / conv1_u
    if (rc_small != SKY_OK && rc_small != SKY_ERR_UNSUPPORTED) return rc_small;
    for (int f0 = 0; rc_small == SKY_ERR_UNSUPPORTED && f0 < F; f0 += 256) {          // more than 256 filters (d4: 512): column slices of the same dY
        p.f0 = f0; p.F = (F - f0) < 256 ? (F - f0) : 256;
        int rc = launch_wgrad(p, st);
        if (rc != SKY_OK) return rc;
    }
    if (dbias) {
        if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)F * sizeof(float), st));
        const int rc = launch_col_sum(dy, dbias, p.M, F, st);
        if (rc != SKY_OK) return rc;
    }
    return SKY_OK;
}
