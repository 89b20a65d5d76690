// Band-staged, persistent, warp-specialised forward distortion-aware convolution for sm_100a (C % 32 == 0).
//
// One CTA per SM loops over tiles of TH x TW = 128 output pixels of one panorama.  Per tile and 32-channel chunk the
// input band the tile can touch (tile + halo derived from the offset table) is brought into shared memory by ONE TMA
// tensor load whose out-of-bounds zero fill *is* the zero halo of _pad_input (distortion_aware_ops.py:125-150).
//
// The sampling offsets depend on (row, tap) only (distortion_aware_ops.py:266-268), so RUN horizontally adjacent
// output pixels read RUN+1 adjacent input columns from the same two input rows with the same bilinear factors.  A
// producer thread therefore blends a whole run from 2*(RUN+1) 128-bit shared-memory loads instead of 4*RUN.
//
// Warp roles (23 warps):
//   warps 0-15  producers  split into groups that alternate k-blocks (chunk cc, tap t).  Per run: vertical then
//                          horizontal blend from the band, TF32 rounding, store into the SWIZZLE_128B A stage.  Runs
//                          flagged irregular (360-degree wrap at the seam, zenith row, band miss) take an exact
//                          per-pixel path.
//   warps 16-19 epilogue + geometry.  Geometry, one tile ahead: exact reference geometry (da_sample) for every
//                          (pixel, tap) of the next tile, reduced per run to {two band row offsets, 4 bilinear
//                          factors, regular?}.  Epilogue: tcgen05.ld the finished accumulator (double-buffered in
//                          TMEM), stage through smem, bias / LeakyReLU / residual, coalesced stores, and the
//                          per-(sample, filter) sum / sum-of-squares for the instance norm that follows
//                          (generator.py:27-33).
//   warp 20     MMA        one lane issues tcgen05.mma.kind::tf32 (M=128, N=F_pad, K=8) x4 per k-block.
//   warp 21     weights    one lane streams the packed weight tile of each k-block with a bulk async copy.
//   warp 22     band       one lane issues the TMA tensor load of each (tile, chunk) band.
#include <math.h>

#include "da_conv.cuh"

namespace sky {

constexpr int RUN = 4;                              // output pixels per producer item (TW % 8 == 0)
constexpr int NRUN = BLOCK_M / RUN;                 // run slots per tile (table stride)
constexpr int PROD_WARPS = 16;                      // producer warps, four per SM sub-partition
// The producer warps are split at run time into `ngroups` groups that alternate k-blocks (so that many A stages are
// being filled at once): 2 groups x 8 warps for 128-pixel tiles, 4 groups x 4 warps for 64-pixel tiles (one thread per
// (run, 16-byte chunk) item either way).  64-pixel tiles fill only half of the UMMA M = 128 rows; they are chosen when
// the problem has too few 128-pixel tiles to occupy the SMs (the trunk of 32x128 panoramas at B = 32 has 64).
constexpr int EPI_WARP0 = PROD_WARPS;               // 4 epilogue/geometry warps; EPI_WARP0 % 4 == 0 keeps warp%4 == TMEM lane quadrant
constexpr int WARP_MMA = PROD_WARPS + 4, WARP_WLOAD = PROD_WARPS + 5, WARP_BAND = PROD_WARPS + 6;
constexpr int BAND_THREADS = (PROD_WARPS + 7) * 32;
constexpr int EPI_COLS = 16, EPI_STRIDE = 20;       // staging row stride (floats): odd multiple of 16 B -> conflict-free
// A pipeline stage holds up to ATOMS k-blocks (two consecutive taps of one 32-channel chunk).  One full/empty barrier
// round trip costs the single MMA-issuing thread ~160 cycles (try_wait) + ~45 (commit), measured in
// tools/ubench_sync.cu; a 32-wide k-block only buys 4 x 67 cycles of tensor work, so stages are made two k-blocks deep.
constexpr int ATOMS = 1;
static_assert(EPI_WARP0 % 4 == 0, "epilogue warps must align with TMEM lane quadrants");

// Optional timeline probe (tools/trace_band.py): block 0 stamps %globaltimer at a few pipeline events when bit 21 of the
// flags word is set.  Never set in production.
__device__ unsigned long long g_band_trace[16];
__device__ __forceinline__ void trace_stamp(const int flags, int slot)
{
    if ((flags & (1 << 21)) && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_band_trace[slot] = t;
    }
}

struct BandParams {
    const float *x, *offsets, *packed, *bias, *residual;
    const float *aux;                      // SKY_EPI_SUN_BLEND: sky prediction (log domain), [pixels][3]
    float threshold;
    float *y;
    double *stats;
    int B, h, w, C, F, Fp, k, k2, CC, KB;
    int in_h, in_w, ph0, pw0;
    int TH, TW, tiles_x, tiles_y, ntiles;
    int tile_px, ngroups, group_threads;   // tile_px = TH*TW (64 or 128); producer grouping
    int plain, stride, oh, ow;             // plain != 0: identity sampler of a SAME conv (ops.py:41), taps at
                                           // (i*stride + a - ph0, j*stride + b - pw0); oh, ow: output map (== h, w otherwise)
    int hy_lo, hx_lo, BH, BW, band_bytes, band_stride, NB;
    int flags;
    float slope;
    uint32_t tmem_cols;
};

// Run table entry: one uint32 {bits 0-14: band offset of (row y0, first column) in 16-byte units, bit 15: regular?,
// bits 16-30: same offset for row y1, bit 31: integer taps (weights exactly (1,0,0,0): a plain conv run through the
// zero offset table, or an on-grid tap such as the centre row of the reference's table) -> straight copy} and floats {dy1, dy0, dx1, dx0} of the run's first pixel
// (the factors of distortion_aware_ops.py:103-106).
template <int STAGES, bool SPLIT3>
struct BandSmem {
    static constexpr int PLANES = SPLIT3 ? 2 : 1;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;
    static __host__ __device__ int b_bytes(int Fp) { return Fp * BLOCK_K * 4; }
    static constexpr int A_STAGE = ATOMS * PLANES * A_BYTES;
    static __host__ __device__ int b_stage(int Fp) { return ATOMS * PLANES * b_bytes(Fp); }
    static __host__ __device__ int stage_bytes(int Fp) { return STAGES * A_STAGE + STAGES * b_stage(Fp); }   // both rings
    // run table: double-buffered by tile, [k2][NRUN] x (packed uint32 offsets + float4 factors)
    static __host__ __device__ int table_bytes(int k2) { return 2 * k2 * NRUN * 20; }
    static constexpr int EPI_BYTES = BLOCK_M * EPI_STRIDE * 4;
    static constexpr int PART_BYTES = 128 * 8 * 4;
    static __host__ __device__ int num_bars(int NB) { return 2 * STAGES + 2 * NB + 8; }
    static __host__ __device__ int total_bytes(int Fp, int band_stride, int NB, int k2)
    {
        return stage_bytes(Fp) + NB * band_stride + table_bytes(k2) + EPI_BYTES + PART_BYTES + num_bars(NB) * 8 + 16 + 1024;
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

// Exact per-pixel path for irregular runs: one corner of one pixel, 4 channels.
__device__ __forceinline__ float4 fetch_exact(int yy_pad, int xx_pad, const BandParams &p, const uint8_t *band, int by0, int bx0,
                                              int b_img, int chunk, int ch_glob)
{
    const int yy = yy_pad - p.ph0, xx = xx_pad - p.pw0;
    if (yy < 0 || yy >= p.h || xx < 0 || xx >= p.w) return make_float4(0.f, 0.f, 0.f, 0.f);   // zero halo
    const int by = yy - by0, bx = xx - bx0;
    if (by >= 0 && by < p.BH && bx >= 0 && bx < p.BW)
        return *reinterpret_cast<const float4 *>(band + (by * p.BW + bx) * (BLOCK_K * 4) + chunk * 16);
    return __ldg(reinterpret_cast<const float4 *>(p.x + ((size_t)(b_img * p.h + yy) * p.w + xx) * p.C + ch_glob));
}

template <int STAGES, bool SPLIT3>
__global__ void __launch_bounds__(BAND_THREADS, 1)
da_conv2d_fwd_band_kernel(const BandParams p, const __grid_constant__ CUtensorMap tmap)
{
    using L = BandSmem<STAGES, SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    // round the dynamic smem base up to 1024 B by OFFSET arithmetic (a uintptr_t round trip would demote every access
    // below to generic LD/ST)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_bytes = L::b_bytes(p.Fp);
    const int b_stage = L::b_stage(p.Fp);
    uint8_t *b_ring = smem + STAGES * L::A_STAGE;
    uint8_t *bands = smem + L::stage_bytes(p.Fp);
    uint8_t *table = bands + p.NB * p.band_stride;
    const int tab_n = p.k2 * NRUN;                                                // entries per tile buffer
    float4 *rt_w = reinterpret_cast<float4 *>(table);                             // [2][k2][NRUN]
    uint32_t *rt_i = reinterpret_cast<uint32_t *>(table + 2 * tab_n * 16);        // [2][k2][NRUN]
    float *epi = reinterpret_cast<float *>(table + L::table_bytes(p.k2));
    float *part = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(epi) + L::EPI_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(part) + L::PART_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + L::num_bars(p.NB));
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES;
    const uint32_t band_full0 = empty0 + 8 * STAGES, band_empty0 = band_full0 + 8 * p.NB;
    const uint32_t tmem_full0 = band_empty0 + 8 * p.NB, tmem_empty0 = tmem_full0 + 16;
    const uint32_t rt_full0 = tmem_empty0 + 16, rt_empty0 = rt_full0 + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) trace_stamp(p.flags, 0);     // kernel start

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, p.group_threads / 32 + 1);   // the producer group that builds this k-block + the weight loader
            mbar_init(empty0 + 8 * s, 1);               // tcgen05.commit (producers and the weight loader both wait on it)
        }
        for (int n = 0; n < p.NB; ++n) {
            mbar_init(band_full0 + 8 * n, 1);   // band loader's expect_tx arrive
            mbar_init(band_empty0 + 8 * n, PROD_WARPS);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);   // tcgen05.commit
            mbar_init(tmem_empty0 + 8 * a, 4);  // 4 epilogue warps
            mbar_init(rt_full0 + 8 * a, 4);     // 4 geometry (= epilogue) warps
            mbar_init(rt_empty0 + 8 * a, PROD_WARPS);
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
    if (tid == 0) trace_stamp(p.flags, 1);     // prologue done (barriers, TMEM)

    if (warp < PROD_WARPS) {
        // ================================================ PRODUCERS ================================================
        const int group = tid / p.group_threads, gtid = tid % p.group_threads;
        const int chunk = gtid & 7, run = gtid >> 3;           // run < tile_px / RUN by construction of the grouping
        const int rty = (run * RUN) / p.TW, rtx = (run * RUN) % p.TW;   // first pixel of this thread's run inside the tile
        uint32_t sg = 0, bandg = 0, it = 0;     // sg: global stage counter (same sequence in every role)
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int b_img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
            const int by0 = i0 * p.stride + p.hy_lo, bx0 = j0 * p.stride + p.hx_lo;
            const int ri = i0 + rty, rj = j0 + rtx;
            const uint32_t tb = it & 1;
            mbar_wait(rt_full0 + 8 * tb, (it >> 1) & 1);              // geometry of this tile is in the run table
            if (tid == 0 && it == 0) trace_stamp(p.flags, 2);       // first run table ready
            const uint32_t *ti = rt_i + tb * tab_n + run;
            const float4 *tw = rt_w + tb * tab_n + run;
            for (int cc = 0; cc < p.CC; ++cc, ++bandg) {
                const int nb = bandg % p.NB;
                mbar_wait(band_full0 + 8 * nb, (bandg / p.NB) & 1);
                if (tid == 0 && bandg == 0) trace_stamp(p.flags, 3);   // first band landed
                const uint8_t *band = bands + nb * p.band_stride;
                const int ch_glob = cc * BLOCK_K + chunk * 4;
                for (int t0 = 0; t0 < p.k2; t0 += ATOMS, ++sg) {
                    if ((int)(sg % p.ngroups) != group) continue;   // another group builds this stage
                    const int natoms = min(ATOMS, p.k2 - t0);
                    const int s = sg % STAGES;
                    mbar_wait(empty0 + 8 * s, ((sg / STAGES) & 1) ^ 1);
                    if (tid == 0 && it == 0 && sg >= 16 && sg < 32) trace_stamp(p.flags, 8 + ((sg - 16) / 4) * 2);   // group 0: stage free
                    for (int at = 0; at < natoms; ++at) {
                        const int t = t0 + at;
                        const uint32_t ew = ti[t * NRUN];
                        const int4 e = make_int4((int)(ew & 0x7FFFu) << 4, (int)((ew >> 16) & 0x7FFFu) << 4, (int)((ew >> 15) & 1u),
                                                 (int)(ew >> 31));
                        const float4 f = tw[t * NRUN];
                        uint8_t *a_tile = smem + s * L::A_STAGE + at * (L::PLANES * L::A_BYTES);
                        if (p.flags & (1 << 16)) {
                            // timing experiment: no gather
                        } else if (e.z && e.w) {
                            // regular run with integer taps: the A rows are the band pixels themselves
                            const uint8_t *top = band + e.x + chunk * 16;
#pragma unroll
                            for (int q = 0; q < RUN; ++q)
                                store_a(a_tile, run * RUN + q, chunk,
                                        *reinterpret_cast<const float4 *>(top + q * p.stride * (BLOCK_K * 4)), SPLIT3);
                        } else if (e.z) {
                            // regular run: RUN+1 adjacent columns of two band rows
                            const uint8_t *top = band + e.x + chunk * 16, *bot = band + e.y + chunk * 16;
                            float4 v[RUN + 1];
#pragma unroll
                            for (int c = 0; c <= RUN; ++c) {
                                const float4 a = *reinterpret_cast<const float4 *>(top + c * (BLOCK_K * 4));
                                const float4 b = *reinterpret_cast<const float4 *>(bot + c * (BLOCK_K * 4));
                                v[c].x = fmaf(f.y, b.x, f.x * a.x);
                                v[c].y = fmaf(f.y, b.y, f.x * a.y);
                                v[c].z = fmaf(f.y, b.z, f.x * a.z);
                                v[c].w = fmaf(f.y, b.w, f.x * a.w);
                            }
#pragma unroll
                            for (int q = 0; q < RUN; ++q) {
                                float4 o;
                                o.x = fmaf(f.w, v[q + 1].x, f.z * v[q].x);
                                o.y = fmaf(f.w, v[q + 1].y, f.z * v[q].y);
                                o.z = fmaf(f.w, v[q + 1].z, f.z * v[q].z);
                                o.w = fmaf(f.w, v[q + 1].w, f.z * v[q].w);
                                store_a(a_tile, run * RUN + q, chunk, o, SPLIT3);
                            }
                        } else {
                            // irregular run: exact per-pixel geometry, corners from the band, from global, or the zero halo
                            const float2 yx = make_float2(f.x, f.y);     // the tap's offsets, left in the table by the geometry warps
                            const int ta = t / p.k, tbb = t % p.k;
                            for (int q = 0; q < RUN; ++q) {
                                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (ri < p.h && rj + q < p.w) {
                                    const Sample sm = da_sample(ri, rj + q, ta, tbb, yx.x, yx.y, p.in_h, p.in_w);
                                    const float4 p0 = fetch_exact(sm.y0, sm.x0, p, band, by0, bx0, b_img, chunk, ch_glob);
                                    const float4 p1 = fetch_exact(sm.y0, sm.x1, p, band, by0, bx0, b_img, chunk, ch_glob);
                                    const float4 p2 = fetch_exact(sm.y1, sm.x0, p, band, by0, bx0, b_img, chunk, ch_glob);
                                    const float4 p3 = fetch_exact(sm.y1, sm.x1, p, band, by0, bx0, b_img, chunk, ch_glob);
                                    o.x = fmaf(sm.w3, p3.x, fmaf(sm.w2, p2.x, fmaf(sm.w1, p1.x, sm.w0 * p0.x)));
                                    o.y = fmaf(sm.w3, p3.y, fmaf(sm.w2, p2.y, fmaf(sm.w1, p1.y, sm.w0 * p0.y)));
                                    o.z = fmaf(sm.w3, p3.z, fmaf(sm.w2, p2.z, fmaf(sm.w1, p1.z, sm.w0 * p0.z)));
                                    o.w = fmaf(sm.w3, p3.w, fmaf(sm.w2, p2.w, fmaf(sm.w1, p1.w, sm.w0 * p0.w)));
                                }
                                store_a(a_tile, run * RUN + q, chunk, o, SPLIT3);
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * s);
                    if (tid == 0 && it == 0 && sg >= 16 && sg < 32) trace_stamp(p.flags, 9 + ((sg - 16) / 4) * 2);   // group 0: stage filled
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(band_empty0 + 8 * nb);   // this warp no longer reads the band buffer
            }
            __syncwarp();
            if (tid == 0 && it == 0) trace_stamp(p.flags, 4);       // producers done with the first tile
            if (lane == 0) mbar_arrive(rt_empty0 + 8 * tb);         // ... nor this tile's run table
        }
    } else if (warp < EPI_WARP0 + 4) {
        // ================================================ EPILOGUE + GEOMETRY ================================================
        const int wq = warp - EPI_WARP0, etid = tid - EPI_WARP0 * 32;
        const bool vec_ok = (p.F % 4) == 0;
        int log_tw = 0;
        while ((1 << log_tw) < p.TW) ++log_tw;            // TW is a power of two
        // geometry duty: thread <-> tile pixel (row-major; TW % RUN == 0 => a run is RUN adjacent lanes)
        const int gt = etid;
        const int pty = gt / p.TW, ptx = gt % p.TW;
        const int grun = gt / RUN, gq = gt % RUN;
        const uint32_t run_mask = ((1u << RUN) - 1u) << (lane & ~(RUN - 1));
        uint32_t it = 0;
        // geometry runs one tile ahead of the epilogue: table(tile_0) first, then per tile: table(next), epilogue(this)
        for (int tile = blockIdx.x - (int)gridDim.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int gtile = tile + (int)gridDim.x;
            const uint32_t git = (tile < 0) ? 0u : it + 1;
            if (gtile < p.ntiles) {
                const int rem = gtile % tiles_per_img;
                const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
                const int by0 = i0 * p.stride + p.hy_lo, bx0 = j0 * p.stride + p.hx_lo;
                const int i = i0 + pty, j = j0 + ptx;
                const bool pix_ok = (gt < p.tile_px) && (i < p.oh) && (j < p.ow);
                const uint32_t tb = git & 1;
                mbar_wait_sleep(rt_empty0 + 8 * tb, ((git >> 1) & 1) ^ 1);
                uint32_t *ti = rt_i + tb * tab_n + grun;
                float4 *tw = rt_w + tb * tab_n + grun;
                // The row's offsets are fetched nine taps at a time (independent loads) instead of one dependent global round
                // trip per tap — this is on the critical path of a CTA's first tile, and for 7x7 layers (49 taps) the dependent
                // loads alone cost more than the tile's tensor work.
                const uint32_t inv_k = (65536u + (uint32_t)p.k - 1u) / (uint32_t)p.k;      // t / k == (t * inv_k) >> 16 for t < 225, k <= 15
                for (int t0 = 0; t0 < p.k2; t0 += 9) {
                    float2 pre[9];
                    if (!p.plain) {
#pragma unroll
                        for (int q = 0; q < 9; ++q)
                            pre[q] = (pix_ok && t0 + q < p.k2) ? __ldg(reinterpret_cast<const float2 *>(p.offsets) + (size_t)i * p.k2 + t0 + q)
                                                              : make_float2(0.f, 0.f);
                    }
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int t = t0 + q;
                        if (t >= p.k2) break;
                        const int a = (int)(((uint32_t)t * inv_k) >> 16), b = t - a * p.k;
                        Sample s;
                        s.y0 = s.y1 = s.x0 = s.x1 = 0; s.dy1 = s.dy0 = s.dx1 = s.dx0 = 0.f;
                        if (p.plain) {
                            // identity sampler: the tap itself, in the same "padded frame" convention (+ph0, +pw0) as da_sample
                            if (gq == 0 && pix_ok) {
                                const int r0 = pty * p.stride + a, c0 = ptx * p.stride + b;       // band-relative
                                ti[t * NRUN] = ((uint32_t)((r0 * p.BW + c0) * (BLOCK_K * 4)) >> 4) | (1u << 15) |
                                               (((uint32_t)((r0 * p.BW + c0) * (BLOCK_K * 4)) >> 4) << 16) | (1u << 31);
                                tw[t * NRUN] = make_float4(1.f, 0.f, 1.f, 0.f);
                            } else if (gq == 0) {
                                ti[t * NRUN] = (1u << 15);
                                tw[t * NRUN] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            continue;
                        }
                        if (pix_ok) s = da_sample(i, j, a, b, pre[q].x, pre[q].y, p.in_h, p.in_w);
                        // the run's first pixel defines the run; every in-image pixel must agree with it
                        const int src = lane & ~(RUN - 1);
                        const int fy0 = __shfl_sync(0xffffffffu, s.y0, src), fy1 = __shfl_sync(0xffffffffu, s.y1, src);
                        const int fx0 = __shfl_sync(0xffffffffu, s.x0, src);
                        const bool f_ok = __shfl_sync(0xffffffffu, (int)pix_ok, src) != 0;
                        bool agree = !pix_ok || (f_ok && s.y0 == fy0 && s.y1 == fy1 && s.x0 == fx0 + gq && s.x1 == s.x0 + 1);
                        const uint32_t votes = __ballot_sync(0xffffffffu, agree);
                        if (gq == 0) {
                            int4 e = make_int4(0, 0, 0, 0);
                            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (pix_ok) {
                                const int r0 = s.y0 - p.ph0 - by0, r1 = s.y1 - p.ph0 - by0, c0 = s.x0 - p.pw0 - bx0;
                                const bool inside = r0 >= 0 && r0 < p.BH && r1 >= 0 && r1 < p.BH && c0 >= 0 && c0 + RUN < p.BW;
                                // columns left of the image (x < 0 in the unpadded frame) are zero-filled by TMA only if they are
                                // not wrapped pixels: a wrapped index lands inside the image far away, so `inside` already fails
                                const bool regular = inside && ((votes & run_mask) == run_mask);
                                // irregular runs never use the offsets: keep them in range of the packed fields
                                const bool integer_taps = s.dy1 == 1.f && s.dy0 == 0.f && s.dx1 == 1.f && s.dx0 == 0.f;
                                e = regular ? make_int4((r0 * p.BW + c0) * (BLOCK_K * 4), (r1 * p.BW + c0) * (BLOCK_K * 4), 1,
                                                        integer_taps ? 1 : 0)
                                            : make_int4(0, 0, 0, 0);
                                f = make_float4(s.dy1, s.dy0, s.dx1, s.dx0);
                                if (!regular) f = make_float4(pre[q].x, pre[q].y, 0.f, 0.f);   // the exact per-pixel path needs the tap's offsets, not the factors
                            } else {
                                e.z = 1;   // run entirely outside the panorama (tile overhang): rows are never stored; read anything
                            }
                            ti[t * NRUN] = ((uint32_t)e.x >> 4) | ((uint32_t)e.z << 15) | (((uint32_t)e.y >> 4) << 16) | ((uint32_t)e.w << 31);
                            tw[t * NRUN] = f;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(rt_full0 + 8 * tb);

            }
            if (tile < 0) { it = (uint32_t)-1; continue; }     // prologue pass: only the first table
            const int b_img = tile / tiles_per_img, rem = tile % tiles_per_img;
            const int i0 = (rem / p.tiles_x) * p.TH, j0 = (rem % p.tiles_x) * p.TW;
            const uint32_t acc = it & 1;
            mbar_wait_sleep(tmem_full0 + 8 * acc, (it >> 1) & 1);   // a whole tile away: back off, leave issue slots to producers
            tc_fence_after();
            if (etid == 0 && it == 0) trace_stamp(p.flags, 5);      // first accumulator complete
            const uint32_t taddr = tmem_base + acc * (uint32_t)p.Fp + ((uint32_t)(wq * 32) << 16);
            for (int c0 = 0; c0 < ((p.flags & (1 << 19)) ? 0 : p.Fp); c0 += EPI_COLS) {   // bit 19: timing experiment, no epilogue
                const int ncols = EPI_COLS;                   // F_pad is a multiple of 16
                {   // phase 1: the row owner (TMEM lane) parks the raw accumulator row in the staging tile
                    static_assert(EPI_COLS == 16, "phase 1 moves 16 columns per pass");
                    uint32_t r[16];
                    tmem_ld_32x16(taddr + (uint32_t)c0, r);
                    tmem_ld_wait();
                    uint4 *dst = reinterpret_cast<uint4 *>(epi + (wq * 32 + lane) * EPI_STRIDE);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                }
                named_bar_sync(2, 128);
                {   // phase 2: each thread owns 4 fixed columns -> bias / activation / residual, coalesced 128-byte stores,
                    //          and the per-filter moments of what was stored
                    const int lg = (ncols == 32) ? 3 : 2;             // log2(float4 columns per row)
                    const int c4 = etid & ((1 << lg) - 1), row0 = etid >> lg, rstep = 128 >> lg;
                    const int f = c0 + 4 * c4;
                    float bv[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (p.bias && f + u < p.F) bv[u] = __ldg(p.bias + f + u);
                    float s1[4] = { 0.f, 0.f, 0.f, 0.f }, s2[4] = { 0.f, 0.f, 0.f, 0.f };
                    if (f < p.F) {
                        for (int row = row0; row < BLOCK_M; row += rstep) {
                            const int ii = i0 + (row >> log_tw), jj = j0 + (row & (p.TW - 1));
                            if (row >= p.tile_px || ii >= p.oh || jj >= p.ow) continue;
                            const float4 raw = *reinterpret_cast<const float4 *>(epi + row * EPI_STRIDE + 4 * c4);
                            float v[4] = { raw.x + bv[0], raw.y + bv[1], raw.z + bv[2], raw.w + bv[3] };
                            if (p.flags & SKY_EPI_LEAKY_RELU) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) v[u] = v[u] > 0.f ? v[u] : v[u] * p.slope;
                            }
                            const size_t go = ((size_t)(b_img * p.oh + ii) * p.ow + jj) * p.F + f;
                            if (p.flags & SKY_EPI_RESIDUAL) {
                                if (vec_ok) {
                                    const float4 rr = __ldg(reinterpret_cast<const float4 *>(p.residual + go));
                                    v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
                                } else {
                                    for (int u = 0; u < 4; ++u)
                                        if (f + u < p.F) v[u] += __ldg(p.residual + go + u);
                                }
                            }
                            if (p.flags & SKY_EPI_MASK) {   // activation gradient of the layer below (mask source in `residual`)
                                if (vec_ok) {
                                    const float4 mk = __ldg(reinterpret_cast<const float4 *>(p.residual + go));
                                    v[0] *= mk.x > 0.f ? 1.f : p.slope; v[1] *= mk.y > 0.f ? 1.f : p.slope;
                                    v[2] *= mk.z > 0.f ? 1.f : p.slope; v[3] *= mk.w > 0.f ? 1.f : p.slope;
                                } else {
                                    for (int u = 0; u < 4; ++u)
                                        if (f + u < p.F) v[u] *= __ldg(p.residual + go + u) > 0.f ? 1.f : p.slope;
                                }
                            }
                            if (p.flags & SKY_EPI_RELU) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) v[u] = fmaxf(v[u], 0.f);
                            }
                            if ((p.flags & SKY_EPI_SUN_BLEND) && f == 0) sun_blend3(v, p.aux + go, p.threshold);   // F == 3: go = 3 * pixel
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
                        // lanes with equal (lane & (c4n-1)) hold partial sums of the same 4 filters: butterfly over the
                        // remaining lane bits, then one fp64 atomic per (warp, filter)
                        for (int o = (1 << lg); o < 32; o <<= 1) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
                                s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
                            }
                        }
                        if (lane < (1 << lg) && f < p.F) {
                            double *st = p.stats + ((size_t)b_img * p.F + f) * 2;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (f + u < p.F) {
                                    atomicAdd(st + 2 * u, (double)s1[u]);
                                    atomicAdd(st + 2 * u + 1, (double)s2[u]);
                                }
                        }
                    }
                }
                named_bar_sync(2, 128);   // staging tile (and the partial sums) are rewritten by the next column chunk
            }
            tc_fence_before();
            __syncwarp();
            if (etid == 0 && it == 0) trace_stamp(p.flags, 6);      // first epilogue done
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * acc);
        }
    } else if (warp == WARP_MMA) {
        // ================================================ MMA ISSUER ================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(BLOCK_M, (uint32_t)p.Fp);
            uint32_t sg = 0, it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                mbar_wait(tmem_empty0 + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.Fp;
                uint32_t first = 0;
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int t0 = 0; t0 < p.k2; t0 += ATOMS, ++sg) {
                        const int natoms = min(ATOMS, p.k2 - t0);
                        const int s = sg % STAGES;
                        mbar_wait(full0 + 8 * s, (sg / STAGES) & 1);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(smem + s * L::A_STAGE);
                        const uint32_t b0 = smem_u32(b_ring + s * b_stage);
                        for (int at = 0; at < natoms; ++at) {
                            const uint32_t a_hi = a0 + at * (L::PLANES * L::A_BYTES);
                            const uint32_t b_hi = b0 + at * (L::PLANES * b_bytes);
#pragma unroll
                            for (int ks = 0; ks < ((p.flags & (1 << 20)) ? 0 : BLOCK_K / UMMA_K); ++ks) {   // bit 20: timing experiment, no MMA
                                const uint32_t koff = ks * UMMA_K * 4;
                                const uint64_t da = umma_desc_kmajor_sw128(a_hi + koff);
                                const uint64_t db = umma_desc_kmajor_sw128(b_hi + koff);
                                umma_tf32(d_tmem, da, db, idesc, first);
                                first = 1;
                                if (SPLIT3) {
                                    umma_tf32(d_tmem, umma_desc_kmajor_sw128(a_hi + L::A_BYTES + koff), db, idesc, 1);
                                    umma_tf32(d_tmem, da, umma_desc_kmajor_sw128(b_hi + b_bytes + koff), idesc, 1);
                                }
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
            const uint32_t atom_bytes = (uint32_t)(L::PLANES * b_bytes);
            // The ring only lets STAGES weight tiles be in flight, so every tile's first touch would pay the full
            // DRAM latency in turn: ask L2 for the whole packed kernel up front (all CTAs want the same bytes).
            for (int kb = 0; kb < p.KB; ++kb)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t *>(p.packed) +
                                                                              (size_t)kb * atom_bytes),
                             "r"(atom_bytes)
                             : "memory");
            uint32_t sg = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int t0 = 0; t0 < p.k2; t0 += ATOMS, ++sg) {
                        const int natoms = min(ATOMS, p.k2 - t0);
                        const int s = sg % STAGES;
                        mbar_wait(empty0 + 8 * s, ((sg / STAGES) & 1) ^ 1);
                        if (p.flags & (1 << 17)) { mbar_arrive(full0 + 8 * s); continue; }   // timing experiment: no weight load
                        const uint32_t bytes = atom_bytes * natoms;
                        mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                        // packed tiles are chunk-major (k-block = cc * k2 + tap): the taps of a stage are contiguous
                        const size_t kb0 = (size_t)cc * p.k2 + t0;
                        bulk_g2s(smem_u32(b_ring + s * b_stage), reinterpret_cast<const uint8_t *>(p.packed) + kb0 * atom_bytes, bytes,
                                 full0 + 8 * s);
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
                    mbar_wait_sleep(band_empty0 + 8 * nb, ((bandg / p.NB) & 1) ^ 1);
                    if (p.flags & (1 << 18)) { mbar_arrive(band_full0 + 8 * nb); continue; }   // timing experiment: no band load
                    mbar_arrive_expect_tx(band_full0 + 8 * nb, (uint32_t)p.band_bytes);
                    tma_load_4d(smem_u32(bands + nb * p.band_stride), &tmap, cc * BLOCK_K, j0 * p.stride + p.hx_lo, i0 * p.stride + p.hy_lo, b_img,
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
    if (tid == 0) trace_stamp(p.flags, 7);     // kernel end
}

}  // namespace sky
extern "C" int sky_debug_band_trace(unsigned long long *host_out16)
{
    SKY_CHECK_CUDA(cudaMemcpyFromSymbol(host_out16, sky::g_band_trace, sizeof(unsigned long long) * 16));
    return SKY_OK;
}
namespace sky {

// Halo of input pixels (relative to an output pixel) that the taps touch, from the host copy of the offset table.
// Every (row, tap) is evaluated with the reference's own arithmetic (same steps as da_sample) at the middle column, so
// the zenith-row taps — whose x offset of about -2w comes back as a small positive shift after the reference's two
// 360-degree wraps — are covered too.  Shifts beyond +-(2k+6) columns are left to the exact fallback path.  The estimate
// only affects speed: the kernel re-checks every corner against the band it actually holds.
void compute_halo(const float *off, int h, int w, int k, int *hy_lo, int *hy_hi, int *hx_lo, int *hx_hi)
{
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    const int in_h = h + pht, in_w = w + pwt;
    // The zenith row's shifted taps widen the band by ~5 columns for every tile; that only pays off when the zenith row
    // is a sizeable share of the map (h <= 16).  On tall maps those few runs take the exact fallback path instead.
    const int cap = h <= 16 ? 2 * k + 6 : k + 1;
    const int j = w / 2;
    int ylo = 0, yhi = 0, xlo = 0, xhi = 0;
    for (int i = 0; i < h; ++i)
        for (int t = 0; t < k * k; ++t) {
            const float yo = off[((size_t)i * k * k + t) * 2 + 0], xo = off[((size_t)i * k * k + t) * 2 + 1];
            if (!(yo == yo) || !(xo == xo)) continue;
            const int a = t / k, b = t % k;
            float y = (float)(i + a) + yo, x = (float)(j + b) + xo;
            y = fminf(fmaxf(y, 0.f), (float)(in_h - 1));
            if (x < 0.f) x = x + (float)in_w;
            if (x > (float)(in_w - 1)) x = x - (float)in_w;
            int y0 = (int)floorf(y), y1 = y0 + 1;
            y0 = y0 < 0 ? 0 : (y0 > in_h - 1 ? in_h - 1 : y0);
            y1 = y1 < 0 ? 0 : (y1 > in_h - 1 ? in_h - 1 : y1);
            // the second (integer) wrap of the reference shifts a still-negative x by in_w once more
            if (x < 0.f) x = x + (float)in_w;
            const int ry0 = y0 - ph0 - i, ry1 = y1 - ph0 - i;
            const float xr = x - (float)(j + pw0);                         // column shift relative to the output pixel
            // +-2e-3: the fractional part of x depends on j through fp32 rounding only
            const int rx0 = (int)floorf(xr - 2e-3f), rx1 = (int)floorf(xr + 2e-3f) + 1;
            if (rx0 < -cap || rx1 > cap) continue;                          // far away (or wrapped): exact fallback path
            ylo = ry0 < ylo ? ry0 : ylo; yhi = ry1 > yhi ? ry1 : yhi;
            xlo = rx0 < xlo ? rx0 : xlo; xhi = rx1 > xhi ? rx1 : xhi;
        }
    *hy_lo = ylo; *hy_hi = yhi; *hx_lo = xlo; *hx_hi = xhi;
}

template <int STAGES, bool SPLIT3>
static int launch_band(BandParams &p, const FwdArgs &a, int hy_span, int hx_span)
{
    using L = BandSmem<STAGES, SPLIT3>;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        SKY_CHECK_CUDA(cudaGetDevice(&dev));
        SKY_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // 128-pixel tiles, unless that leaves more than a quarter of the SMs without a tile: then 64-pixel tiles
    static const int cand128[5][2] = { { 8, 16 }, { 4, 32 }, { 16, 8 }, { 2, 64 }, { 1, 128 } };
    static const int cand64[4][2] = { { 8, 8 }, { 4, 16 }, { 2, 32 }, { 1, 64 } };
    BandParams best[2];
    long cost_of[2] = { -1, -1 };
    // A stage must belong to ONE producer group (STAGES % ngroups == 0): a group waits on its stage's empty barrier by
    // parity, which is only unambiguous when its consecutive uses of that stage are consecutive phases.
    for (int v = 0; v < ((STAGES % 4 == 0) ? 2 : 1); ++v) {
        const int px = v == 0 ? 128 : 64, ncand = v == 0 ? 5 : 4;
        for (int nb_try = (p.CC > 1 ? 2 : 1); nb_try >= 1 && cost_of[v] < 0; --nb_try)
            for (int c = 0; c < ncand; ++c) {
                const int TH = v == 0 ? cand128[c][0] : cand64[c][0], TW = v == 0 ? cand128[c][1] : cand64[c][1];
                if ((TH > 2 * p.oh && TH > 1) || (TW > 2 * p.ow && TW > 8)) continue;
                const int BH = (TH - 1) * p.stride + 1 + hy_span, BW = (TW - 1) * p.stride + 1 + hx_span;
                if (BW > 256 || BH > 256 || BW < RUN + 1) continue;
                const int band_bytes = BH * BW * BLOCK_K * 4;
                const int band_stride = round_up(band_bytes, 1024);
                if (L::total_bytes(p.Fp, band_stride, nb_try, p.k2) > 227 * 1024) continue;
                const int tiles_y = (p.oh + TH - 1) / TH, tiles_x = (p.ow + TW - 1) / TW;
                const long cost = (long)tiles_y * tiles_x * BH * BW;
                if (cost_of[v] < 0 || cost < cost_of[v]) {
                    cost_of[v] = cost;
                    best[v] = p;
                    best[v].TH = TH; best[v].TW = TW; best[v].BH = BH; best[v].BW = BW; best[v].band_bytes = band_bytes;
                    best[v].band_stride = band_stride; best[v].tiles_x = tiles_x; best[v].tiles_y = tiles_y; best[v].NB = nb_try;
                    best[v].tile_px = px;
                }
            }
    }
    long best_cost = -1;
    if (cost_of[0] >= 0) {
        const long tiles128 = (long)best[0].tiles_x * best[0].tiles_y * a.B;
        const bool few = tiles128 * 4 < (long)num_sms * 3;
        const int pick = (few && cost_of[1] >= 0) ? 1 : 0;
        p = best[pick];
        best_cost = cost_of[pick];
    } else if (cost_of[1] >= 0) {
        p = best[1];
        best_cost = cost_of[1];
    }
    if (best_cost < 0) return SKY_ERR_UNSUPPORTED;
    p.ntiles = p.tiles_x * p.tiles_y * a.B;
    p.ngroups = p.tile_px == 128 ? 2 : 4;
    p.group_threads = PROD_WARPS * 32 / p.ngroups;      // == (tile_px / RUN) * 8: one thread per (run, chunk)

    CUtensorMap tmap;
    int rc = encode_nhwc_tensor_map(&tmap, a.x, a.B, a.h, a.w, a.C, BLOCK_K, p.BW, p.BH);
    if (rc != SKY_OK) return rc;

    const int smem = L::total_bytes(p.Fp, p.band_stride, p.NB, p.k2);
    SKY_ENSURE_DYN_SMEM((da_conv2d_fwd_band_kernel<STAGES, SPLIT3>), 227 * 1024);
    const int grid = p.ntiles < num_sms ? p.ntiles : num_sms;
    da_conv2d_fwd_band_kernel<STAGES, SPLIT3><<<grid, BAND_THREADS, smem, a.stream>>>(p, tmap);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

int launch_fwd_band(const FwdArgs &a)
{
    if (a.C % BLOCK_K != 0 || (a.offsets_host == nullptr && a.plain_stride == 0)) return SKY_ERR_UNSUPPORTED;
    if (a.ldF > 0 && a.ldF != a.F) return SKY_ERR_UNSUPPORTED;      // filter slices take the direct kernel
    BandParams p;
    p.aux = a.aux; p.threshold = a.threshold;
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
    p.plain = 0; p.stride = 1; p.oh = a.h; p.ow = a.w;
    if (a.plain_stride > 0) {
        // tf.nn.conv2d SAME: out = ceil(n/s), total pad = max((out-1)*s + k - n, 0), the smaller half in front
        p.plain = 1; p.stride = a.plain_stride;
        p.oh = (a.h + p.stride - 1) / p.stride; p.ow = (a.w + p.stride - 1) / p.stride;
        const int th = (p.oh - 1) * p.stride + a.k - a.h, tw = (p.ow - 1) * p.stride + a.k - a.w;
        p.ph0 = (th > 0 ? th : 0) / 2; p.pw0 = (tw > 0 ? tw : 0) / 2;
        hy_lo = -p.ph0; hx_lo = -p.pw0; hy_hi = hy_lo + a.k - 1; hx_hi = hx_lo + a.k - 1;
    } else {
        compute_halo(a.offsets_host, a.h, a.w, a.k, &hy_lo, &hy_hi, &hx_lo, &hx_hi);
    }
    p.hy_lo = hy_lo; p.hx_lo = hx_lo;
    const int hy_span = hy_hi - hy_lo, hx_span = hx_hi - hx_lo;
    if (a.math_mode == SKY_MATH_TF32)
        return p.Fp <= 128 ? launch_band<4, false>(p, a, hy_span, hx_span) : SKY_ERR_UNSUPPORTED;
    return p.Fp <= 128 ? launch_band<2, true>(p, a, hy_span, hx_span) : SKY_ERR_UNSUPPORTED;
}

}  // namespace sky
