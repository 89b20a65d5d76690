// Row-strip convolution kernel for sm_100a (see strip_conv.cuh for the formulation): persistent, warp-specialised, tcgen05 / TMEM.
//
//   warps 0-15   producers   write strips: per (strip row, 16-byte channel chunk) two 128-bit global loads (the two input rows of the
//                            vertical blend), blend, TF32 rounding (+ the low part for 3xTF32), one 128-bit shared-memory store.
//                            k strips per (tile, 32-channel chunk) instead of k*k im2col tiles.
//   warps 16-19  epilogue    tcgen05.ld of the finished accumulator (double-buffered in TMEM) -> staging -> bias / LeakyReLU / residual /
//                            mask / ReLU / log decompression -> coalesced stores and the instance-norm moments per (sample, filter).
//   warp 20      MMA         one lane: per window 4 x tcgen05.mma.kind::tf32 (M = 128, N = filters of the slice, K = 8); the A descriptor is
//                            the strip in the no-swizzle K-major layout (8-row x 16-byte core matrices, planes of 16-byte channel chunks),
//                            started at the window's row, so a column shift costs nothing.
//   warp 21      weights     one lane: bulk async copy of each window's packed weight tile (128-byte-swizzled K-major, as packed by
//                            sky_da_pack_weights / sky_da_strip_pack_weights).
//
// Tile = one output row class x TW columns x NB panoramas (TW * NB = 128); tile row m = column * NB + panorama, so that a column shift of
// s is a row shift of s * NB for every panorama of the tile at once.
#include <math.h>

#include "strip_conv.cuh"

namespace sky {

constexpr int SC_PROD_WARPS = 16;
constexpr int SC_PROD_THREADS = SC_PROD_WARPS * 32;
constexpr int SC_EPI_WARP0 = SC_PROD_WARPS;           // 4 epilogue warps; % 4 == 0 keeps warp % 4 == TMEM lane quadrant
constexpr int SC_WARP_MMA = SC_PROD_WARPS + 4, SC_WARP_WLOAD = SC_PROD_WARPS + 5;
constexpr int SC_THREADS = (SC_PROD_WARPS + 6) * 32;
constexpr int SC_EPI_COLS = 16, SC_EPI_STRIDE = 20;   // staging row stride (floats): odd multiple of 16 B -> conflict-free
constexpr int SC_UNROLL = 3;                          // producer items in flight per thread (x 2 loads each)
static_assert(SC_EPI_WARP0 % 4 == 0, "epilogue warps must align with TMEM lane quadrants");

struct StripParams {
    const float *x;
    const uint8_t *packed;
    const float *bias, *residual;
    float *y;
    double *stats;
    const RowPlan *rows;
    const StripDesc *strips;
    const WinDesc *wins;
    int B, H, W, C, CC;                 // input tensor [B,H,W,C], CC = C / 32
    int OH, OW, ldF;                    // output tensor [B,OH,OW,ldF]
    int F, Fs;                          // valid filters of a slice, N of the MMA (multiple of 16)
    int slice_rows;                     // 1: slices are row ranges [slice * Fs, ...) of one packed image of tile_rows rows per plane
                                        // 0: every slice has its own packed image (slice_stride bytes apart) of tile_rows == Fs rows
    int tile_rows;
    size_t slice_stride;
    int wmul, wcc_stride;               // weight tile of (window, chunk cc) = wtile0 * wmul + cc * wcc_stride
    int ncols, ocs, TW, NB, log_nb, tiles_x, tiles_b, ntiles;
    int SR, PS, strip_bytes, NSB, NWB, G, group_threads;
    int da, in_h, in_w, ph0, pw0;       // distortion-aware column map (the reference's wrap in the padded frame) and exact taps
    int k;
    int flags;
    float slope;
    uint32_t tmem_cols;
};

template <bool SPLIT3>
struct StripSmem {
    static constexpr int PLANES = SPLIT3 ? 2 : 1;
    static __host__ __device__ int strip_bytes(int PS) { return round_up(PLANES * 8 * PS, 128); }
    static __host__ __device__ int b_stage(int Fs) { return PLANES * Fs * BLOCK_K * 4; }
    static constexpr int EPI_BYTES = BLOCK_M * SC_EPI_STRIDE * 4;
    static __host__ __device__ int num_bars(int NSB, int NWB) { return 2 * NSB + 2 * NWB + 4; }
    static __host__ __device__ int total_bytes(int PS, int Fs, int NSB, int NWB)
    {
        return round_up(NSB * strip_bytes(PS), 1024) + NWB * b_stage(Fs) + EPI_BYTES + num_bars(NSB, NWB) * 8 + 16 + 1024;
    }
};

// K-major operand without swizzle: core matrix = 8 rows x 16 bytes (contiguous 128 B); `sbo` = distance between 8-row groups,
// `lbo` = distance between the two 16-byte halves of a K = 8 step.  Any 16-byte aligned start address is legal, which is what lets a
// window start at an arbitrary strip row.
__device__ __forceinline__ uint64_t umma_desc_kmajor_interleaved(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    return d;                          // layout type 0: no swizzle
}

// The reference's treatment of a column index (distortion_aware_ops.py:76-77 on the float coordinate, :90-91 on the integer corners),
// restated on the integer part: `q` is a column in the PADDED frame before any wrap.  Returns the unpadded column, or -1 for a zero.
__device__ __forceinline__ int da_map_col(int q, int in_w, int pw0, int W)
{
    if (q < 0) q += in_w;                 // :76 (x < 0 -> x + in_w; the result is <= in_w - 1, so :77 does not fire after it)
    else if (q > in_w - 1) q -= in_w;     // :77
    if (q < 0) q += in_w;                 // :90
    if (q > in_w - 1) q -= in_w;          // :91
    const int c = q - pw0;
    return (c >= 0 && c < W) ? c : -1;
}

__device__ __forceinline__ void strip_store(uint8_t *buf, int PS, int rho, int c, float4 v, bool split3)
{
    uint4 hi;
    hi.x = f32_to_tf32_rna(v.x); hi.y = f32_to_tf32_rna(v.y); hi.z = f32_to_tf32_rna(v.z); hi.w = f32_to_tf32_rna(v.w);
    uint8_t *dst = buf + c * PS + rho * 16;
    *reinterpret_cast<uint4 *>(dst) = hi;
    if (split3) {
        uint4 lo;
        lo.x = f32_to_tf32_rna(v.x - __uint_as_float(hi.x));
        lo.y = f32_to_tf32_rna(v.y - __uint_as_float(hi.y));
        lo.z = f32_to_tf32_rna(v.z - __uint_as_float(hi.z));
        lo.w = f32_to_tf32_rna(v.w - __uint_as_float(hi.w));
        *reinterpret_cast<uint4 *>(dst + 8 * PS) = lo;
    }
}

template <bool SPLIT3>
__global__ void __launch_bounds__(SC_THREADS, 1) strip_conv_kernel(const StripParams p)
{
    using L = StripSmem<SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *strip_ring = smem;
    uint8_t *b_ring = smem + round_up(p.NSB * p.strip_bytes, 1024);
    const int b_stage = L::b_stage(p.Fs);
    const int b_plane = p.Fs * BLOCK_K * 4;
    float *epi = reinterpret_cast<float *>(b_ring + p.NWB * b_stage);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(epi) + L::EPI_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + L::num_bars(p.NSB, p.NWB));
    const uint32_t sfull0 = smem_u32(bars), sempty0 = sfull0 + 8 * p.NSB;
    const uint32_t bfull0 = sempty0 + 8 * p.NSB, bempty0 = bfull0 + 8 * p.NWB;
    const uint32_t tmem_full0 = bempty0 + 8 * p.NWB, tmem_empty0 = tmem_full0 + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < p.NSB; ++s) {
            mbar_init(sfull0 + 8 * s, p.group_threads / 32);   // the warps of the producer group that owns the buffer
            mbar_init(sempty0 + 8 * s, 1);                      // tcgen05.commit
        }
        for (int s = 0; s < p.NWB; ++s) {
            mbar_init(bfull0 + 8 * s, 1);                       // the weight loader's expect_tx arrive
            mbar_init(bempty0 + 8 * s, 1);                      // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);
            mbar_init(tmem_empty0 + 8 * a, 4);
        }
        fence_mbar_init();
    }
    if (warp == SC_WARP_MMA) {
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_row = p.tiles_x * p.tiles_b;

    if (warp < SC_PROD_WARPS) {
        // ================================================ PRODUCERS ================================================
        const int group = tid / p.group_threads, gtid = tid % p.group_threads, GT = p.group_threads;
        uint32_t sq = 0;                                   // strip sequence number, the same in every role
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int rp = tile / tiles_per_row, rem = tile % tiles_per_row;
            const int j0 = (rem / p.tiles_b) * p.TW, b0 = (rem % p.tiles_b) * p.NB;
            const RowPlan row = p.rows[rp];
            for (int cc = 0; cc < p.CC; ++cc) {
                for (int si = row.strip_begin; si < row.strip_end; ++si, ++sq) {
                    if ((int)(sq % p.G) != group) continue;
                    const StripDesc sd = p.strips[si];
                    const int sb = sq % p.NSB;
                    mbar_wait(sempty0 + 8 * sb, ((sq / p.NSB) & 1) ^ 1);
                    uint8_t *buf = strip_ring + sb * p.strip_bytes;
                    const int ch0 = cc * BLOCK_K;
                    if (sd.kind == 0) {
                        const int items = p.SR * 8;
                        const bool has0 = sd.r0 >= 0 && sd.wy0 != 0.f, has1 = sd.r1 >= 0 && sd.wy1 != 0.f;
                        for (int base = gtid; base < items; base += GT * SC_UNROLL) {
                            float4 v0[SC_UNROLL], v1[SC_UNROLL];
#pragma unroll
                            for (int u = 0; u < SC_UNROLL; ++u) {
                                const int idx = base + u * GT;
                                v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                                v1[u] = v0[u];
                                if (idx < items) {
                                    const int rho = idx >> 3, c = idx & 7;
                                    const int pi = rho >> p.log_nb, bimg = b0 + (rho & (p.NB - 1));
                                    int col = sd.cm * (j0 + sd.u0 + pi) + sd.c0;
                                    if (p.da) col = da_map_col(col + p.pw0, p.in_w, p.pw0, p.W);
                                    if (col >= 0 && col < p.W && bimg < p.B) {
                                        const float *src = p.x + ((size_t)bimg * p.H * p.W + col) * p.C + ch0 + c * 4;
                                        if (has0) v0[u] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)sd.r0 * p.W * p.C));
                                        if (has1) v1[u] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)sd.r1 * p.W * p.C));
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < SC_UNROLL; ++u) {
                                const int idx = base + u * GT;
                                if (idx < items) {
                                    float4 o;
                                    o.x = fmaf(sd.wy1, v1[u].x, sd.wy0 * v0[u].x);
                                    o.y = fmaf(sd.wy1, v1[u].y, sd.wy0 * v0[u].y);
                                    o.z = fmaf(sd.wy1, v1[u].z, sd.wy0 * v0[u].z);
                                    o.w = fmaf(sd.wy1, v1[u].w, sd.wy0 * v0[u].w);
                                    strip_store(buf, p.PS, idx >> 3, idx & 7, o, SPLIT3);
                                }
                            }
                        }
                    } else {
                        // exact tap: the reference's per-pixel geometry (da_sample) and its four-corner blend for every pixel of the tile
                        const int ta = sd.r0 / p.k, tb = sd.r0 % p.k;
                        for (int idx = gtid; idx < BLOCK_M * 8; idx += GT) {
                            const int rho = idx >> 3, c = idx & 7;
                            const int j = j0 + (rho >> p.log_nb), bimg = b0 + (rho & (p.NB - 1));
                            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (j < p.OW && bimg < p.B) {
                                const Sample sm = da_sample(row.out_row, j, ta, tb, sd.wy0, sd.wy1, p.in_h, p.in_w);
                                const CornerRef cr = da_corners(sm, bimg, p.H, p.W, p.C, p.ph0, p.pw0);
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    if (cr.off[q] < 0) continue;
                                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(p.x + cr.off[q] + ch0 + c * 4));
                                    o.x = fmaf(cr.w[q], pv.x, o.x); o.y = fmaf(cr.w[q], pv.y, o.y);
                                    o.z = fmaf(cr.w[q], pv.z, o.z); o.w = fmaf(cr.w[q], pv.w, o.w);
                                }
                            }
                            strip_store(buf, p.PS, rho, c, o, SPLIT3);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sfull0 + 8 * sb);
                }
            }
        }
    } else if (warp < SC_EPI_WARP0 + 4) {
        // ================================================ EPILOGUE ================================================
        const int wq = warp - SC_EPI_WARP0, etid = tid - SC_EPI_WARP0 * 32;
        const bool vec_ok = (p.F % 4) == 0 && (p.ldF % 4) == 0;
        const int f_base = slice * (p.slice_rows ? p.Fs : p.F);          // first filter of this slice in the layer
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int rp = tile / tiles_per_row, rem = tile % tiles_per_row;
            const int j0 = (rem / p.tiles_b) * p.TW, b0 = (rem % p.tiles_b) * p.NB;
            const RowPlan row = p.rows[rp];
            const uint32_t acc = it & 1;
            const bool no_terms = row.strip_begin == row.strip_end;      // a row class nothing contributes to: the accumulator is not written
            mbar_wait_sleep(tmem_full0 + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * (uint32_t)p.Fs + ((uint32_t)(wq * 32) << 16);
            for (int c0 = 0; c0 < p.Fs; c0 += SC_EPI_COLS) {
                {   // phase 1: the row owner (TMEM lane) parks 16 raw accumulator columns in the staging tile
                    uint32_t r[16];
                    tmem_ld_32x16(taddr + (uint32_t)c0, r);
                    tmem_ld_wait();
                    if (no_terms) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) r[q] = 0u;
                    }
                    uint4 *dst = reinterpret_cast<uint4 *>(epi + (wq * 32 + lane) * SC_EPI_STRIDE);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                }
                named_bar_sync(2, 128);
                {   // phase 2: each thread owns 4 fixed columns and the tile rows row0 + 32 n, which belong to ONE panorama (NB | 32)
                    const int c4 = etid & 3, row0 = etid >> 2;
                    const int f = c0 + 4 * c4;                            // filter inside the slice
                    const int bimg = b0 + (row0 & (p.NB - 1));
                    float bv[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (p.bias && f + u < p.F) bv[u] = __ldg(p.bias + f_base + f + u);
                    float s1[4] = { 0.f, 0.f, 0.f, 0.f }, s2[4] = { 0.f, 0.f, 0.f, 0.f };
                    if (f < p.F && bimg < p.B) {
#pragma unroll
                        for (int rr = 0; rr < BLOCK_M / 32; ++rr) {
                            const int m = row0 + 32 * rr;
                            const int jj = j0 + (m >> p.log_nb);
                            if (jj >= p.ncols) continue;
                            const int ocol = row.oc0 + p.ocs * jj;
                            const float4 raw = *reinterpret_cast<const float4 *>(epi + m * SC_EPI_STRIDE + 4 * c4);
                            float v[4] = { raw.x + bv[0], raw.y + bv[1], raw.z + bv[2], raw.w + bv[3] };
                            if (p.flags & SKY_EPI_LEAKY_RELU) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) v[u] = v[u] > 0.f ? v[u] : v[u] * p.slope;
                            }
                            const size_t go = ((size_t)(bimg * p.OH + row.out_row) * p.OW + ocol) * p.ldF + f_base + f;
                            if (p.flags & (SKY_EPI_RESIDUAL | SKY_EPI_MASK)) {
                                float rv[4] = { 0.f, 0.f, 0.f, 0.f };
                                if (vec_ok) {
                                    const float4 r4 = __ldg(reinterpret_cast<const float4 *>(p.residual + go));
                                    rv[0] = r4.x; rv[1] = r4.y; rv[2] = r4.z; rv[3] = r4.w;
                                } else {
                                    for (int u = 0; u < 4; ++u)
                                        if (f + u < p.F) rv[u] = __ldg(p.residual + go + u);
                                }
                                if (p.flags & SKY_EPI_RESIDUAL) {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) v[u] += rv[u];
                                } else {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) v[u] *= rv[u] > 0.f ? 1.f : p.slope;
                                }
                            }
                            if (p.flags & SKY_EPI_RELU) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) v[u] = fmaxf(v[u], 0.f);
                            }
                            if (p.flags & SKY_EPI_LOG_DECOMPRESS) {   // tf_utils.hdr_logDecompression, log(11) as fp32
#pragma unroll
                                for (int u = 0; u < 4; ++u) v[u] = (expf(v[u] * 2.3978953f) - 1.f) / 10.f;
                            }
                            if (vec_ok) {
                                *reinterpret_cast<float4 *>(p.y + go) = make_float4(v[0], v[1], v[2], v[3]);
                            } else {
                                for (int u = 0; u < 4; ++u)
                                    if (f + u < p.F) p.y[go + u] = v[u];
                                    else v[u] = 0.f;
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) { s1[u] += v[u]; s2[u] = fmaf(v[u], v[u], s2[u]); }
                        }
                    }
                    if (p.stats) {
                        // lanes l and l ^ o hold partial sums of the same 4 filters; they belong to the same panorama when
                        // ((l >> 2) ^ (o >> 2)) % NB == (l >> 2) % NB, i.e. o >= 4 * NB
                        for (int o = 4 * p.NB; o < 32; o <<= 1) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
                                s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
                            }
                        }
                        if ((lane >> 2) < p.NB && f < p.F && bimg < p.B) {
                            double *st = p.stats + ((size_t)bimg * p.ldF + f_base + f) * 2;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (f + u < p.F) {
                                    atomicAdd(st + 2 * u, (double)s1[u]);
                                    atomicAdd(st + 2 * u + 1, (double)s2[u]);
                                }
                        }
                    }
                }
                named_bar_sync(2, 128);   // the staging tile is rewritten by the next column pass
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * acc);
        }
    } else if (warp == SC_WARP_MMA) {
        // ================================================ MMA ISSUER ================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BLOCK_M, (uint32_t)p.Fs);
            uint32_t sq = 0, wq = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const RowPlan row = p.rows[tile / tiles_per_row];
                const uint32_t acc = it & 1;
                mbar_wait(tmem_empty0 + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Fs;
                uint32_t first = 0;
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int si = row.strip_begin; si < row.strip_end; ++si, ++sq) {
                        const int wb = __ldg(&p.strips[si].win_begin), we = __ldg(&p.strips[si].win_end);
                        const int sb = sq % p.NSB;
                        mbar_wait(sfull0 + 8 * sb, (sq / p.NSB) & 1);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(strip_ring + sb * p.strip_bytes);
                        for (int wi = wb; wi < we; ++wi, ++wq) {
                            const int start_row = __ldg(&p.wins[wi].start_row);
                            const int ws = wq % p.NWB;
                            mbar_wait(bfull0 + 8 * ws, (wq / p.NWB) & 1);
                            tc_fence_after();
                            const uint32_t b0 = smem_u32(b_ring + ws * b_stage);
                            const uint32_t a_w = a0 + (uint32_t)start_row * 16u;
#pragma unroll
                            for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                                const uint64_t da = umma_desc_kmajor_interleaved(a_w + 2 * ks * p.PS, (uint32_t)p.PS, 128u);
                                const uint64_t db = umma_desc_kmajor_sw128(b0 + ks * UMMA_K * 4);
                                umma_tf32(d_tmem, da, db, idesc, first);
                                first = 1;
                                if (SPLIT3) {
                                    const uint64_t da_lo = umma_desc_kmajor_interleaved(a_w + (8 + 2 * ks) * p.PS, (uint32_t)p.PS, 128u);
                                    const uint64_t db_lo = umma_desc_kmajor_sw128(b0 + b_plane + ks * UMMA_K * 4);
                                    umma_tf32(d_tmem, da_lo, db, idesc, 1);
                                    umma_tf32(d_tmem, da, db_lo, idesc, 1);
                                }
                            }
                            umma_commit(bempty0 + 8 * ws);
                        }
                        umma_commit(sempty0 + 8 * sb);
                    }
                umma_commit(tmem_full0 + 8 * acc);
            }
        }
        __syncwarp();
    } else {
        // ================================================ WEIGHT LOADER ================================================
        if (lane == 0) {
            const size_t tile_plane = (size_t)p.tile_rows * BLOCK_K * 4;              // one plane of one packed tile
            const size_t tile_bytes = (size_t)L::PLANES * tile_plane;
            const uint8_t *base = p.packed + (p.slice_rows ? (size_t)slice * p.Fs * (BLOCK_K * 4) : (size_t)slice * p.slice_stride);
            uint32_t wq = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const RowPlan row = p.rows[tile / tiles_per_row];
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int si = row.strip_begin; si < row.strip_end; ++si) {
                        const int wb = __ldg(&p.strips[si].win_begin), we = __ldg(&p.strips[si].win_end);
                        for (int wi = wb; wi < we; ++wi, ++wq) {
                            const int wt = __ldg(&p.wins[wi].wtile0) * p.wmul + cc * p.wcc_stride;
                            const int ws = wq % p.NWB;
                            mbar_wait(bempty0 + 8 * ws, ((wq / p.NWB) & 1) ^ 1);
                            const uint32_t dst = smem_u32(b_ring + ws * b_stage);
                            const uint8_t *src = base + (size_t)wt * tile_bytes;
                            mbar_arrive_expect_tx(bfull0 + 8 * ws, (uint32_t)b_stage);
                            bulk_g2s(dst, src, (uint32_t)b_plane, bfull0 + 8 * ws);
                            if (SPLIT3) bulk_g2s(dst + b_plane, src + tile_plane, (uint32_t)b_plane, bfull0 + 8 * ws);
                        }
                    }
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == SC_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// effective weights of a distortion-aware layer: tile (window, cc) = sum over the window's terms of coef * kernel[tap*C + cc*32 + kk, n],
// written as the K-major 128-byte-swizzled image the MMA reads (plane 0 = tf32(rna(v)), plane 1 (3xTF32) = tf32(rna(v - hi)))
// ---------------------------------------------------------------------------------------------------------------------
__global__ void strip_weff_pack_kernel(const float *__restrict__ kernel, float *__restrict__ packed, const int *__restrict__ term_begin,
                                       const WeffTerm *__restrict__ terms, int nwins, int C, int CC, int F, int Fp, int planes)
{
    const long total = (long)nwins * CC * Fp * BLOCK_K;
    const size_t tile_floats = (size_t)Fp * BLOCK_K;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        // n fastest: consecutive threads read consecutive filters of one kernel row (coalesced)
        const int n = (int)(e % Fp);
        const int kk = (int)((e / Fp) % BLOCK_K);
        const long wc = e / ((long)Fp * BLOCK_K);
        const int cc = (int)(wc % CC), w = (int)(wc / CC);
        float v = 0.f;
        if (n < F) {
            const int tb = term_begin[w], te = term_begin[w + 1];
            for (int q = tb; q < te; ++q) {
                const WeffTerm t = terms[q];
                v = fmaf(t.coef, kernel[((size_t)t.tap * C + cc * BLOCK_K + kk) * F + n], v);
            }
        }
        const uint32_t hi = f32_to_tf32_rna(v);
        const size_t o = (sw128_offset((uint32_t)n, (uint32_t)(kk >> 2)) >> 2) + (kk & 3);
        float *tile = packed + (size_t)wc * planes * tile_floats;
        tile[o] = __uint_as_float(hi);
        if (planes == 2) tile[tile_floats + o] = __uint_as_float(f32_to_tf32_rna(v - __uint_as_float(hi)));
    }
}

static int num_sms_cached()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

struct StripLaunch {
    const float *x; const void *packed; const float *bias, *residual; float *y; double *stats;
    int B, H, W, C;          // input
    int OH, OW;              // output map
    int F, ldF;              // filters of this launch (one packed image unless nimages > 1), row stride of y
    int nimages;             // > 1: `nimages` packed images of F filters each, slice_stride bytes apart (filters beyond 256)
    size_t image_stride;
    int k, flags, math_mode, da, in_h, in_w, ph0, pw0;
    float slope;
    cudaStream_t stream;
};

template <bool SPLIT3>
static int launch_strip(const StripLaunch &a, const StripPlan &pl)
{
    using L = StripSmem<SPLIT3>;
    StripParams p;
    p.x = a.x; p.packed = (const uint8_t *)a.packed; p.bias = a.bias; p.residual = a.residual; p.y = a.y; p.stats = a.stats;
    p.rows = pl.rows; p.strips = pl.strips; p.wins = pl.wins;
    p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.CC = a.C / BLOCK_K;
    p.OH = a.OH; p.OW = a.OW; p.ldF = a.ldF;
    p.ncols = pl.ncols; p.ocs = pl.ocs; p.TW = pl.TW; p.NB = pl.NB;
    p.log_nb = 0;
    while ((1 << p.log_nb) < p.NB) ++p.log_nb;
    p.tiles_x = (pl.ncols + pl.TW - 1) / pl.TW;
    p.tiles_b = (a.B + pl.NB - 1) / pl.NB;
    p.ntiles = pl.nrows * p.tiles_x * p.tiles_b;
    p.SR = pl.SR; p.PS = pl.SR * 16 + 16;               // odd multiple of 16 B: the 8 chunk planes of a row land in 8 different bank groups
    p.strip_bytes = L::strip_bytes(p.PS);
    p.da = a.da; p.in_h = a.in_h; p.in_w = a.in_w; p.ph0 = a.ph0; p.pw0 = a.pw0; p.k = a.k;
    p.flags = a.flags; p.slope = a.slope;
    const int Fp = f_pad_of(a.F);
    p.wmul = pl.weff ? p.CC : 1;
    p.wcc_stride = pl.weff ? 1 : a.k * a.k;
    p.tile_rows = Fp;
    int nslices = 1;
    p.slice_rows = 1; p.slice_stride = 0; p.F = a.F; p.Fs = Fp;
    if (a.nimages > 1) {
        nslices = a.nimages; p.slice_rows = 0; p.slice_stride = a.image_stride;
    } else if (Fp >= 128 && Fp % 32 == 0 && ((long)p.ntiles * 2 <= num_sms_cached() || SPLIT3)) {
        // few tiles (the trunk of 32x128 panoramas at B = 32 has 64): two CTAs per tile, each with half of the filters.  3xTF32 always
        // splits: its second operand plane would not leave room for the strip ring otherwise.
        if (a.F != Fp) return SKY_ERR_UNSUPPORTED;      // ragged filter counts are not needed by the path
        nslices = 2; p.Fs = Fp / 2; p.F = p.Fs;
    }
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < 2 * p.Fs) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return SKY_ERR_UNSUPPORTED;
    // ring depths: producer groups own whole strip buffers (NSB % G == 0, see the parity argument in the band kernel)
    const int budget = 227 * 1024;
    int best_nsb = 0, best_nwb = 0, best_g = 0;
    const int gs[3] = { 4, 2, 1 };
    for (int gi = 0; gi < 3 && !best_nsb; ++gi) {
        const int G = gs[gi];
        for (int nsb = 2 * G; nsb >= G && !best_nsb; nsb -= G)
            for (int nwb = 6; nwb >= 2; --nwb)
                if (L::total_bytes(p.PS, p.Fs, nsb, nwb) <= budget) { best_nsb = nsb; best_nwb = nwb; best_g = G; break; }
    }
    if (!best_nsb) return SKY_ERR_UNSUPPORTED;
    p.NSB = best_nsb; p.NWB = best_nwb; p.G = best_g; p.group_threads = SC_PROD_THREADS / best_g;
    const int smem = L::total_bytes(p.PS, p.Fs, p.NSB, p.NWB);
    SKY_ENSURE_DYN_SMEM((strip_conv_kernel<SPLIT3>), 227 * 1024);
    int gx = num_sms_cached() / nslices;
    if (gx < 1) gx = 1;
    if (gx > p.ntiles) gx = p.ntiles;
    strip_conv_kernel<SPLIT3><<<dim3(gx, nslices), SC_THREADS, smem, a.stream>>>(p);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

static int launch_strip_any(const StripLaunch &a, const StripPlan &pl)
{
    return a.math_mode == SKY_MATH_3XTF32 ? launch_strip<true>(a, pl) : launch_strip<false>(a, pl);
}

int launch_fwd_strip_plain(const FwdArgs &a)
{
    if (a.C % BLOCK_K != 0 || a.plain_stride <= 0) return SKY_ERR_UNSUPPORTED;
    if (a.flags & (SKY_EPI_SUN_BLEND | SKY_EPI_FORCE_DIRECT)) return SKY_ERR_UNSUPPORTED;
    if (a.F > 256) return SKY_ERR_UNSUPPORTED;
    const StripPlan *pl = nullptr;
    int OH, OW;
    if (a.transposed) {
        OH = a.out_h; OW = a.out_w;
    } else {
        OH = (a.h + a.plain_stride - 1) / a.plain_stride; OW = (a.w + a.plain_stride - 1) / a.plain_stride;
    }
    int rc = get_plan_plain(a.h, a.w, a.k, a.plain_stride, a.transposed, OH, OW, a.tp_ph0, a.tp_pw0, &pl);
    if (rc != SKY_OK) return rc;
    StripLaunch s;
    s.x = a.x; s.packed = a.packed; s.bias = a.bias; s.residual = a.residual; s.y = a.y; s.stats = a.stats;
    s.B = a.B; s.H = a.h; s.W = a.w; s.C = a.C; s.OH = OH; s.OW = OW;
    s.F = a.F; s.ldF = a.ldF > 0 ? a.ldF : a.F;
    s.nimages = a.nslices > 0 ? a.nslices : 1;
    const int K = a.k * a.k * a.C, KB = (K + BLOCK_K - 1) / BLOCK_K;
    s.image_stride = (size_t)KB * (a.math_mode == SKY_MATH_3XTF32 ? 2 : 1) * f_pad_of(a.F) * BLOCK_K * sizeof(float);
    s.k = a.k; s.flags = a.flags; s.math_mode = a.math_mode; s.da = 0; s.in_h = s.in_w = s.ph0 = s.pw0 = 0; s.slope = a.slope;
    s.stream = a.stream;
    return launch_strip_any(s, *pl);
}

}  // namespace sky

using namespace sky;

extern "C" size_t sky_da_strip_weight_bytes(const float *offsets_host, int h, int w, int C, int F, int k, int math_mode)
{
    if (!offsets_host || C <= 0 || F <= 0 || C % BLOCK_K != 0) return 0;
    const StripPlan *pl = nullptr;
    if (get_plan_da(offsets_host, h, w, k, &pl, false) != SKY_OK) return 0;
    const int planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    return (size_t)pl->nwins * (C / BLOCK_K) * planes * f_pad_of(F) * BLOCK_K * sizeof(float);
}

extern "C" int sky_da_strip_pack_weights(const float *kernel, const float *offsets_host, void *packed, int h, int w, int C, int F, int k,
                                         int math_mode, void *stream)
{
    SKY_REQUIRE(kernel && offsets_host && packed, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(C > 0 && F > 0 && C % BLOCK_K == 0 && F <= 256, SKY_ERR_UNSUPPORTED, "strip weights need C %% 32 == 0 and F <= 256 (C=%d F=%d)", C, F);
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl);
    if (rc != SKY_OK) return rc;
    const int Fp = f_pad_of(F), CC = C / BLOCK_K, planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    const long total = (long)pl->nwins * CC * Fp * BLOCK_K;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    strip_weff_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kernel, (float *)packed, pl->term_begin, pl->terms, pl->nwins, C, CC, F, Fp,
                                                                    planes);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_conv2d_fwd_strip(const float *x, const float *offsets_host, const void *packed, const float *bias, float *y,
                                       const float *residual, double *stats, int B, int h, int w, int C, int F, int k, int epilogue_flags,
                                       float slope, int math_mode, void *stream)
{
    SKY_REQUIRE(x && offsets_host && packed && bias && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension (B=%d h=%d w=%d C=%d F=%d)", B, h, w, C, F);
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    SKY_REQUIRE(C % BLOCK_K == 0 && F <= 256, SKY_ERR_UNSUPPORTED, "the strip kernel needs C %% 32 == 0 and F <= 256 (C=%d F=%d)", C, F);
    SKY_REQUIRE(!(epilogue_flags & (SKY_EPI_RESIDUAL | SKY_EPI_MASK)) || residual, SKY_ERR_INVALID, "SKY_EPI_RESIDUAL / SKY_EPI_MASK without its tensor");
    SKY_REQUIRE(!(epilogue_flags & SKY_EPI_SUN_BLEND), SKY_ERR_INVALID, "SKY_EPI_SUN_BLEND is taken by sky_conv2d_fwd_blend");
    SKY_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)packed & 15) == 0, SKY_ERR_INVALID, "x, y and packed must be 16-byte aligned");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl);
    if (rc != SKY_OK) return rc;
    StripLaunch s;
    s.x = x; s.packed = packed; s.bias = bias; s.residual = residual; s.y = y; s.stats = stats;
    s.B = B; s.H = h; s.W = w; s.C = C; s.OH = h; s.OW = w; s.F = F; s.ldF = F; s.nimages = 1; s.image_stride = 0;
    s.k = k; s.flags = epilogue_flags; s.math_mode = math_mode; s.da = 1; s.slope = slope; s.stream = (cudaStream_t)stream;
    int pht, pwt;
    pad_axis(h, k, &s.ph0, &pht);
    pad_axis(w, k, &s.pw0, &pwt);
    s.in_h = h + pht; s.in_w = w + pwt;
    rc = launch_strip_any(s, *pl);
    SKY_REQUIRE(rc != SKY_ERR_UNSUPPORTED, SKY_ERR_UNSUPPORTED, "the strip kernel does not cover this layer (F=%d, strip rows %d)", F, pl->SR);
    return rc;
}
