// Band-staged, persistent, warp-specialised forward distortion-aware convolution for sm_100a (C % 32 == 0).
//
// One CTA per SM loops over tiles of TH x TW = 128 output pixels of one panorama.  Per tile and 32-channel chunk the
// input band the tile can touch (tile + halo derived from the offset table) is brought into shared memory by ONE TMA
// tensor load whose out-of-bounds zero fill *is* the zero halo of _pad_input (distortion_aware_ops.py:125-150).
// Warp roles (352 threads):
//   warps 0-3  producers   per k-block (chunk cc, tap t): exact sampling geometry (da_sample) for the thread's pixel ->
//                          table in smem; then each thread blends 8 (pixel, 16-byte chunk) items from the band
//                          (4 x LDS.128 + fma) and writes them TF32-rounded into the SWIZZLE_128B A stage.
//                          Corners the band does not hold (360-degree wrap at the seam, zenith row) are read from global.
//   warps 4-7  epilogue    tcgen05.ld the finished accumulator (double-buffered in TMEM), bias / LeakyReLU, stage
//                          through smem for coalesced stores, residual add, per-(sample, filter) sum / sum-of-squares
//                          for the instance norm that follows (generator.py:27-33).
//   warp 8     MMA         one lane issues tcgen05.mma.kind::tf32 (M=128, N=F_pad, K=8) x4 per k-block.
//   warp 9     weights     one lane streams the packed weight tile of each k-block with a bulk async copy.
//   warp 10    band        one lane issues the TMA tensor load of each (tile, chunk) band.
#include <math.h>

#include "da_conv.cuh"

namespace sky {

constexpr int BAND_THREADS = 352;
constexpr int WARP_MMA = 8, WARP_WLOAD = 9, WARP_BAND = 10;
constexpr int EPI_COLS = 32, EPI_STRIDE = 36;   // staging row stride (floats): odd multiple of 16 B -> conflict-free

struct BandParams {
    const float *x, *offsets, *packed, *bias, *residual;
    float *y;
    double *stats;
    int B, h, w, C, F, Fp, k, k2, CC, KB;
    int in_h, in_w, ph0, pw0;
    int TH, TW, tiles_x, tiles_y, ntiles;
    int hy_lo, hx_lo, BH, BW, band_bytes, band_stride, NB;
    int flags;
    float slope;
    uint32_t tmem_cols;
};

struct TabEntry {
    int off[4];   // >= 0: byte offset of the pixel inside the band buffer; -1: zero; <= -2: global element offset -(off+2)
    float w[4];
};

template <int STAGES, bool SPLIT3>
struct BandSmem {
    static constexpr int PLANES = SPLIT3 ? 2 : 1;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
    static __host__ __device__ int b_bytes(int Fp) { return Fp * BLOCK_K * 4; }
    static __host__ __device__ int stage_bytes(int Fp) { return PLANES * (A_BYTES + b_bytes(Fp)); }
    static constexpr int TABLE_BYTES = 2 * BLOCK_M * (int)sizeof(TabEntry);
    static constexpr int EPI_BYTES = BLOCK_M * EPI_STRIDE * 4;
    static constexpr int PART_BYTES = 4 * EPI_COLS * 2 * 4;
    static __host__ __device__ int num_bars(int NB) { return 2 * STAGES + 2 * NB + 4; }
    static __host__ __device__ int total_bytes(int Fp, int band_stride, int NB)
    {
        return STAGES * stage_bytes(Fp) + NB * band_stride + TABLE_BYTES + EPI_BYTES + PART_BYTES + num_bars(NB) * 8 + 16 + 1024;
    }
};

__device__ __forceinline__ void store_a(uint8_t *a_tile, int row, int chunk, float4 v, bool split3)
{
    uint4 hi;
    hi.x = f32_to_tf32_rna(v.x); hi.y = f32_to_tf32_rna(v.y); hi.z = f32_to_tf32_rna(v.z); hi.w = f32_to_tf32_rna(v.w);
    const uint32_t o = sw128_offset((uint32_t)row, (uint32_t)chunk);
    *reinterpret_cast<uint4 *>(a_tile + o) = hi;
    if (split3) {
        uint4 lo;
        lo.x = f32_to_tf32_rna(v.x - __uint_as_float(hi.x));
        lo.y = f32_to_tf32_rna(v.y - __uint_as_float(hi.y));
        lo.z = f32_to_tf32_rna(v.z - __uint_as_float(hi.z));
        lo.w = f32_to_tf32_rna(v.w - __uint_as_float(hi.w));
        *reinterpret_cast<uint4 *>(a_tile + BLOCK_M * BLOCK_K * 4 + o) = lo;
    }
}

__device__ __forceinline__ float4 fetch_corner(int off, const uint8_t *band, int chunk_bytes, const float *x, int ch_glob)
{
    if (off >= 0) return *reinterpret_cast<const float4 *>(band + off + chunk_bytes);
    if (off < -1) return __ldg(reinterpret_cast<const float4 *>(x + (size_t)(-(off + 2)) + ch_glob));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int STAGES, bool SPLIT3>
__global__ void __launch_bounds__(BAND_THREADS, 1)
da_conv2d_fwd_band_kernel(const BandParams p, const __grid_constant__ CUtensorMap tmap)
{
    using L = BandSmem<STAGES, SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = L::stage_bytes(p.Fp);
    const int b_bytes = L::b_bytes(p.Fp);
    uint8_t *bands = smem + STAGES * stage_bytes;
    TabEntry *table = reinterpret_cast<TabEntry *>(bands + p.NB * p.band_stride);
    float *epi = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(table) + L::TABLE_BYTES);
    float *part = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(epi) + L::EPI_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(part) + L::PART_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + L::num_bars(p.NB));
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES;
    const uint32_t band_full0 = empty0 + 8 * STAGES, band_empty0 = band_full0 + 8 * p.NB;
    const uint32_t tmem_full0 = band_empty0 + 8 * p.NB, tmem_empty0 = tmem_full0 + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 4 + 1);    // 4 producer warps + the weight loader's expect_tx arrive
            mbar_init(empty0 + 8 * s, 1);       // tcgen05.commit
        }
        for (int n = 0; n < p.NB; ++n) {
            mbar_init(band_full0 + 8 * n, 1);   // band loader's expect_tx arrive
            mbar_init(band_empty0 + 8 * n, 4);  // 4 producer warps
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);   // tcgen05.commit
            mbar_init(tmem_empty0 + 8 * a, 4);  // 4 epilogue warps
        }
        fence_mbar_init();
    }
    if (warp == WARP_MMA) {
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    }
    if (warp == WARP_BAND && lane == 0) prefetch_tmap(&tmap);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp < 4) {
        // ================================================ PRODUCERS ================================================
        const int chunk = tid & 7, row_base = tid >> 3;
        const int pty = tid / p.TW, ptx = tid % p.TW;       // this thread's pixel inside the tile (table duty)
        uint32_t kbg = 0, bandg = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int b_img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
            const int i = i0 + pty, j = j0 + ptx;
            const bool pix_ok = (i < p.h) && (j < p.w);
            const int by0 = i0 + p.hy_lo, bx0 = j0 + p.hx_lo;
            for (int cc = 0; cc < p.CC; ++cc, ++bandg) {
                const int nb = bandg % p.NB;
                mbar_wait(band_full0 + 8 * nb, (bandg / p.NB) & 1);
                const uint8_t *band = bands + nb * p.band_stride;
                const int ch_glob = cc * BLOCK_K + chunk * 4;
                for (int t = 0; t < p.k2; ++t, ++kbg) {
                    // ---- table duty: exact reference geometry for (pixel, tap) ----
                    TabEntry e;
#pragma unroll
                    for (int c = 0; c < 4; ++c) { e.off[c] = -1; e.w[c] = 0.f; }
                    if (pix_ok) {
                        const float yo = __ldg(p.offsets + ((size_t)i * p.k2 + t) * 2 + 0);
                        const float xo = __ldg(p.offsets + ((size_t)i * p.k2 + t) * 2 + 1);
                        const Sample s = da_sample(i, j, t / p.k, t % p.k, yo, xo, p.in_h, p.in_w);
                        const int ys[4] = { s.y0, s.y0, s.y1, s.y1 };
                        const int xs[4] = { s.x0, s.x1, s.x0, s.x1 };
                        e.w[0] = s.w0; e.w[1] = s.w1; e.w[2] = s.w2; e.w[3] = s.w3;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int yy = ys[c] - p.ph0, xx = xs[c] - p.pw0;
                            if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
                                const int by = yy - by0, bx = xx - bx0;
                                if (by >= 0 && by < p.BH && bx >= 0 && bx < p.BW) e.off[c] = (by * p.BW + bx) * (BLOCK_K * 4);
                                else e.off[c] = -2 - ((b_img * p.h + yy) * p.w + xx) * p.C;
                            }
                        }
                    }
                    TabEntry *tab = table + (kbg & 1) * BLOCK_M;
                    tab[tid] = e;
                    named_bar_sync(1, 128);
                    // ---- gather + blend into the A stage ----
                    const int s = kbg % STAGES;
                    mbar_wait(empty0 + 8 * s, ((kbg / STAGES) & 1) ^ 1);
                    uint8_t *a_tile = smem + s * stage_bytes;
#pragma unroll 4
                    for (int r = 0; r < BLOCK_M / 16; ++r) {
                        const int row = row_base + 16 * r;
                        const TabEntry q = tab[row];
                        const float4 p0 = fetch_corner(q.off[0], band, chunk * 16, p.x, ch_glob);
                        const float4 p1 = fetch_corner(q.off[1], band, chunk * 16, p.x, ch_glob);
                        const float4 p2 = fetch_corner(q.off[2], band, chunk * 16, p.x, ch_glob);
                        const float4 p3 = fetch_corner(q.off[3], band, chunk * 16, p.x, ch_glob);
                        float4 v;   // add_n order of distortion_aware_ops.py:112-113
                        v.x = fmaf(q.w[3], p3.x, fmaf(q.w[2], p2.x, fmaf(q.w[1], p1.x, q.w[0] * p0.x)));
                        v.y = fmaf(q.w[3], p3.y, fmaf(q.w[2], p2.y, fmaf(q.w[1], p1.y, q.w[0] * p0.y)));
                        v.z = fmaf(q.w[3], p3.z, fmaf(q.w[2], p2.z, fmaf(q.w[1], p1.z, q.w[0] * p0.z)));
                        v.w = fmaf(q.w[3], p3.w, fmaf(q.w[2], p2.w, fmaf(q.w[1], p1.w, q.w[0] * p0.w)));
                        store_a(a_tile, row, chunk, v, SPLIT3);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * s);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(band_empty0 + 8 * nb);   // this warp no longer reads the band buffer
            }
        }
    } else if (warp < 8) {
        // ================================================ EPILOGUE ================================================
        const int wq = warp - 4, etid = tid - 128;
        const bool vec_ok = (p.F % 4) == 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int b_img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
            const uint32_t acc = it & 1;
            mbar_wait(tmem_full0 + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * (uint32_t)p.Fp + ((uint32_t)(wq * 32) << 16);
            for (int c0 = 0; c0 < p.Fp; c0 += EPI_COLS) {
                const int ncols = min(EPI_COLS, p.Fp - c0);   // 32 or 16
                uint32_t r[32];
                if (ncols == 32) tmem_ld_32x32(taddr + (uint32_t)c0, r);
                else {
                    uint32_t r16[16];
                    tmem_ld_32x16(taddr + (uint32_t)c0, r16);
#pragma unroll
                    for (int q = 0; q < 16; ++q) { r[q] = r16[q]; r[q + 16] = 0u; }
                }
                tmem_ld_wait();
                {   // phase 1: row-owner applies bias / activation and parks the row in the staging tile
                    float *dst = epi + (wq * 32 + lane) * EPI_STRIDE;
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float o[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int f = c0 + q + u;
                            float val = __uint_as_float(r[q + u]);
                            if (f < p.F) {
                                val += __ldg(p.bias + f);
                                if (p.flags & SKY_EPI_LEAKY_RELU) val = val > 0.f ? val : val * p.slope;
                            } else val = 0.f;
                            o[u] = val;
                        }
                        *reinterpret_cast<float4 *>(dst + q) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                named_bar_sync(2, 128);
                {   // phase 2: coalesced stores (8 threads cover 128 contiguous bytes of one pixel), residual add
                    const int c4n = ncols / 4;
                    for (int idx = etid; idx < BLOCK_M * c4n; idx += 128) {
                        const int row = idx / c4n, c4 = idx % c4n;
                        const int ii = i0 + row / p.TW, jj = j0 + row % p.TW;
                        const int f = c0 + 4 * c4;
                        if (ii < p.h && jj < p.w && f < p.F) {
                            float *sp = epi + row * EPI_STRIDE + 4 * c4;
                            float4 v = *reinterpret_cast<float4 *>(sp);
                            const size_t go = ((size_t)(b_img * p.h + ii) * p.w + jj) * p.F + f;
                            if (vec_ok) {
                                if (p.flags & SKY_EPI_RESIDUAL) {
                                    const float4 rr = __ldg(reinterpret_cast<const float4 *>(p.residual + go));
                                    v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
                                    *reinterpret_cast<float4 *>(sp) = v;
                                }
                                *reinterpret_cast<float4 *>(p.y + go) = v;
                            } else {
                                float vv[4] = { v.x, v.y, v.z, v.w };
                                for (int u = 0; u < 4; ++u)
                                    if (f + u < p.F) {
                                        if (p.flags & SKY_EPI_RESIDUAL) vv[u] += __ldg(p.residual + go + u);
                                        p.y[go + u] = vv[u];
                                    }
                                *reinterpret_cast<float4 *>(sp) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                            }
                        }
                    }
                }
                if (p.stats) {   // phase 3: per-filter sum / sum of squares over the tile's valid pixels
                    named_bar_sync(2, 128);
                    const int c = etid & 31, g = etid >> 5;
                    float s1 = 0.f, s2 = 0.f;
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = g * 32 + rr;
                        const int ii = i0 + row / p.TW, jj = j0 + row % p.TW;
                        if (ii < p.h && jj < p.w) {
                            const float v = epi[row * EPI_STRIDE + c];
                            s1 += v;
                            s2 = fmaf(v, v, s2);
                        }
                    }
                    part[(g * EPI_COLS + c) * 2 + 0] = s1;
                    part[(g * EPI_COLS + c) * 2 + 1] = s2;
                    named_bar_sync(2, 128);
                    if (g == 0 && c < ncols && c0 + c < p.F) {
                        double d1 = 0.0, d2 = 0.0;
#pragma unroll
                        for (int gg = 0; gg < 4; ++gg) {
                            d1 += (double)part[(gg * EPI_COLS + c) * 2 + 0];
                            d2 += (double)part[(gg * EPI_COLS + c) * 2 + 1];
                        }
                        double *st = p.stats + ((size_t)b_img * p.F + c0 + c) * 2;
                        atomicAdd(st, d1);
                        atomicAdd(st + 1, d2);
                    }
                }
                named_bar_sync(2, 128);   // staging tile is rewritten by the next column chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * acc);
        }
    } else if (warp == WARP_MMA) {
        // ================================================ MMA ISSUER ================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BLOCK_M, (uint32_t)p.Fp);
            uint32_t kbg = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                mbar_wait(tmem_empty0 + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Fp;
                for (int kb = 0; kb < p.KB; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    mbar_wait(full0 + 8 * s, (kbg / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
                    const uint32_t b_hi = a_hi + L::PLANES * L::A_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                        const uint32_t koff = ks * UMMA_K * 4;
                        const uint64_t da = umma_desc_kmajor_sw128(a_hi + koff);
                        const uint64_t db = umma_desc_kmajor_sw128(b_hi + koff);
                        umma_tf32(d_tmem, da, db, idesc, (kb | ks) != 0);
                        if (SPLIT3) {
                            umma_tf32(d_tmem, umma_desc_kmajor_sw128(a_hi + L::A_BYTES + koff), db, idesc, 1);
                            umma_tf32(d_tmem, da, umma_desc_kmajor_sw128(b_hi + b_bytes + koff), idesc, 1);
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                }
                umma_commit(tmem_full0 + 8 * acc);
            }
        }
        __syncwarp();
    } else if (warp == WARP_WLOAD) {
        // ================================================ WEIGHT LOADER ================================================
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(L::PLANES * b_bytes);
            uint32_t kbg = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int t = 0; t < p.k2; ++t, ++kbg) {
                        const int s = kbg % STAGES;
                        mbar_wait(empty0 + 8 * s, ((kbg / STAGES) & 1) ^ 1);
                        const uint32_t dst = smem_u32(smem + s * stage_bytes + L::PLANES * L::A_BYTES);
                        mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                        // packed tiles are stored in the kernel variable's row order: k-block = tap * CC + chunk
                        const size_t kb_nat = (size_t)t * p.CC + cc;
                        bulk_g2s(dst, reinterpret_cast<const uint8_t *>(p.packed) + kb_nat * bytes, bytes, full0 + 8 * s);
                    }
        }
        __syncwarp();
    } else {
        // ================================================ BAND LOADER (TMA) ================================================
        if (lane == 0) {
            uint32_t bandg = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int b_img = tile / tiles_per_img, rem = tile % tiles_per_img;
                const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
                for (int cc = 0; cc < p.CC; ++cc, ++bandg) {
                    const int nb = bandg % p.NB;
                    mbar_wait(band_empty0 + 8 * nb, ((bandg / p.NB) & 1) ^ 1);
                    mbar_arrive_expect_tx(band_full0 + 8 * nb, (uint32_t)p.band_bytes);
                    tma_load_4d(smem_u32(bands + nb * p.band_stride), &tmap, cc * BLOCK_K, j0 + p.hx_lo, i0 + p.hy_lo, b_img,
                                band_full0 + 8 * nb);
                }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// Halo of input pixels (relative to an output pixel) that the regular taps can touch, from the host copy of the offset
// table.  Zenith-row taps whose x offset is about -2w (the reference's double 360-degree wrap) are left to the global
// fallback.  The estimate only affects speed: the kernel re-checks every corner against the band it actually holds.
static void compute_halo(const float *off, int h, int w, int k, int *hy_lo, int *hy_hi, int *hx_lo, int *hx_hi)
{
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    const int in_h = h + pht;
    int ylo = 0, yhi = 0, xlo = 0, xhi = 0;
    for (int i = 0; i < h; ++i)
        for (int t = 0; t < k * k; ++t) {
            const float yo = off[((size_t)i * k * k + t) * 2 + 0], xo = off[((size_t)i * k * k + t) * 2 + 1];
            if (!(fabsf(xo) <= 2.f * k + 2.f) || !(fabsf(yo) <= 2.f * k + 2.f)) continue;
            const int a = t / k, b = t % k;
            float y = (float)(i + a) + yo;
            y = fminf(fmaxf(y, 0.f), (float)(in_h - 1));
            int y0 = (int)floorf(y), y1 = y0 + 1;
            y0 = y0 < 0 ? 0 : (y0 > in_h - 1 ? in_h - 1 : y0);
            y1 = y1 < 0 ? 0 : (y1 > in_h - 1 ? in_h - 1 : y1);
            const int ry0 = y0 - ph0 - i, ry1 = y1 - ph0 - i;
            const float xr = (float)b + xo;
            const int rx0 = (int)floorf(xr - 2e-3f) - pw0, rx1 = (int)floorf(xr + 2e-3f) + 1 - pw0;
            ylo = ry0 < ylo ? ry0 : ylo; yhi = ry1 > yhi ? ry1 : yhi;
            xlo = rx0 < xlo ? rx0 : xlo; xhi = rx1 > xhi ? rx1 : xhi;
        }
    *hy_lo = ylo; *hy_hi = yhi; *hx_lo = xlo; *hx_hi = xhi;
}

template <int STAGES, bool SPLIT3>
static int launch_band(BandParams &p, const FwdArgs &a, int hy_span, int hx_span)
{
    using L = BandSmem<STAGES, SPLIT3>;
    static const int cand[8][2] = { { 8, 16 }, { 4, 32 }, { 16, 8 }, { 2, 64 }, { 1, 128 }, { 32, 4 }, { 64, 2 }, { 128, 1 } };
    long best_cost = -1;
    for (int nb_try = (p.CC > 1 ? 2 : 1); nb_try >= 1 && best_cost < 0; --nb_try)
        for (int c = 0; c < 8; ++c) {
            const int TH = cand[c][0], TW = cand[c][1];
            if ((TH > 2 * a.h && TH > 1) || (TW > 2 * a.w && TW > 1)) continue;
            const int BH = TH + hy_span, BW = TW + hx_span;
            if (BW > 256 || BH > 256) continue;
            const int band_bytes = BH * BW * BLOCK_K * 4;
            const int band_stride = round_up(band_bytes, 1024);
            if (L::total_bytes(p.Fp, band_stride, nb_try) > 227 * 1024) continue;
            const int tiles_y = (a.h + TH - 1) / TH, tiles_x = (a.w + TW - 1) / TW;
            const long cost = (long)tiles_y * tiles_x * BH * BW;
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                p.TH = TH; p.TW = TW; p.BH = BH; p.BW = BW; p.band_bytes = band_bytes; p.band_stride = band_stride;
                p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.NB = nb_try;
            }
        }
    if (best_cost < 0) return SKY_ERR_UNSUPPORTED;
    p.ntiles = p.tiles_x * p.tiles_y * a.B;

    CUtensorMap tmap;
    int rc = encode_nhwc_tensor_map(&tmap, a.x, a.B, a.h, a.w, a.C, BLOCK_K, p.BW, p.BH);
    if (rc != SKY_OK) return rc;

    const int smem = L::total_bytes(p.Fp, p.band_stride, p.NB);
    static bool configured = false;
    if (!configured) {
        SKY_CHECK_CUDA(cudaFuncSetAttribute(da_conv2d_fwd_band_kernel<STAGES, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SKY_CHECK_CUDA(cudaGetDevice(&dev));
        SKY_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
    da_conv2d_fwd_band_kernel<STAGES, SPLIT3><<<grid, BAND_THREADS, smem, a.stream>>>(p, tmap);
    SKY_CHECK_CUDA(cudaGetLastError());
    return SKY_OK;
}

int launch_fwd_band(const FwdArgs &a)
{
    if (a.C % BLOCK_K != 0 || a.offsets_host == nullptr) return SKY_ERR_UNSUPPORTED;
    BandParams p;
    p.x = a.x; p.offsets = a.offsets; p.packed = a.packed; p.bias = a.bias; p.residual = a.residual; p.y = a.y; p.stats = a.stats;
    p.B = a.B; p.h = a.h; p.w = a.w; p.C = a.C; p.F = a.F; p.Fp = f_pad_of(a.F); p.k = a.k; p.k2 = a.k * a.k;
    p.CC = a.C / BLOCK_K; p.KB = p.k2 * p.CC;
    int pht, pwt;
    pad_axis(a.h, a.k, &p.ph0, &pht);
    pad_axis(a.w, a.k, &p.pw0, &pwt);
    p.in_h = a.h + pht; p.in_w = a.w + pwt;
    p.flags = a.flags; p.slope = a.slope;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < 2 * p.Fp) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return SKY_ERR_UNSUPPORTED;
    int hy_lo, hy_hi, hx_lo, hx_hi;
    compute_halo(a.offsets_host, a.h, a.w, a.k, &hy_lo, &hy_hi, &hx_lo, &hx_hi);
    p.hy_lo = hy_lo; p.hx_lo = hx_lo;
    const int hy_span = hy_hi - hy_lo, hx_span = hx_hi - hx_lo;
    if (a.math_mode == SKY_MATH_TF32)
        return p.Fp <= 128 ? launch_band<3, false>(p, a, hy_span, hx_span) : launch_band<2, false>(p, a, hy_span, hx_span);
    return p.Fp <= 128 ? launch_band<2, true>(p, a, hy_span, hx_span) : SKY_ERR_UNSUPPORTED;
}

}  // namespace sky
