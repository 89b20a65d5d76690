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
#include <stdlib.h>

#include "strip_conv.cuh"

namespace sky {

// Warp w runs on scheduler (sub-partition) w % 4.  The single lane that issues the MMAs executes a dependent instruction chain per window
// (barrier wait, descriptor words, 4 x tcgen05.mma, commit); sharing its scheduler with busy producer warps made that chain ~1000
// cycles per window (timeline in tools/trace_strip.py) — so scheduler 0 hosts only the MMA warp, the weight loader and one (mostly
// sleeping) epilogue warp, and the 12 producer warps live on schedulers 1-3:
//   warps 0-3 epilogue (warp == TMEM lane quadrant), warp 4 MMA, warp 8 weights, warps 12 / 16 idle, every other warp of 5..19 a producer
constexpr int SC_PROD_WARPS = 12;
constexpr int SC_PROD_THREADS = SC_PROD_WARPS * 32;
constexpr int SC_EPI_WARP0 = 0;
constexpr int SC_WARP_MMA = 4, SC_WARP_WLOAD = 8;
constexpr int SC_THREADS = 20 * 32;
constexpr int SC_EPI_COLS = 16, SC_EPI_STRIDE = 20;   // staging row stride (floats): odd multiple of 16 B -> conflict-free
constexpr int SC_PAIR_MAX_N = 32;                      // widest layer (padded filters) that pairs tiles
constexpr int SC_MAX_STRIPS = 16, SC_MAX_WINS = 192; // strips / windows of one row class (staged in shared memory)
constexpr int SC_UNROLL2 = 5, SC_UNROLL1 = 8;          // producer items in flight per thread: strips blending two input rows / reading one
static_assert(SC_EPI_WARP0 % 4 == 0, "epilogue warps must align with TMEM lane quadrants");

// Optional timeline probe (tools/trace_strip.py): CTA (0, 0) stamps %globaltimer at pipeline events when bit 21 of the flags word is set.
// Never set in production.
__device__ unsigned long long g_strip_trace[64];
__device__ __forceinline__ void strip_stamp(const int flags, int slot)
{
    if ((flags & (1 << 21)) && blockIdx.x == 0 && blockIdx.y == 0 && slot < 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_strip_trace[slot] = t;
    }
}

static_assert(sizeof(StripDesc) == 40, "the producers copy the strip table word by word");

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
    int F, Fs;                          // filters of one packed image, N of the MMA (filters of a slice, multiple of 16)
    int nsub;                           // blockIdx.y = image * nsub + sub: sub-slice `sub` covers filters [sub * Fs, sub * Fs + Fs) of its image
    int tile_rows;                      // rows per plane of a packed tile (the image's padded filter count)
    size_t image_stride;                // bytes between the packed images of a layer with more than 256 filters
    int wmul, wcc_stride;               // weight tile of (window, chunk cc) = wtile0 * wmul + cc * wcc_stride
    int ncols, ocs, TW, NB, log_nb, tiles_x, tiles_b, ntiles;
    int pair, pairs_per_row, nwork;     // pair = 2: a CTA step covers two tiles of one row class (M = 256 through two accumulators) per weight tile
    int SR, PS, strip_bytes, NSB, NWB, G, group_threads;
    int da, in_h, in_w, ph0, pw0;       // da = 1: distortion-aware column map (the reference's wrap in the padded frame) and exact taps;
                                        // da = 2: its transpose (data gradient); 0: plain convolution (zero outside the map)
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
    static constexpr int PLAN_BYTES = 2 * (SC_MAX_STRIPS * 8 + SC_MAX_WINS * 4) + SC_MAX_STRIPS * 40;   // the current row class: one copy per
                                                                                     // single-lane role + the producers' strip table
    static __host__ __device__ int total_bytes(int PS, int Fs, int NSB, int NWB, int pair = 1)
    {
        return round_up(NSB * pair * strip_bytes(PS), 1024) + NWB * b_stage(Fs) + EPI_BYTES + PLAN_BYTES + num_bars(NSB, NWB) * 8 + 16 + 1024;
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

// The transposed map of the data gradient: `q` = input column c minus the forward shift s; returns the unique dy column j in [0, W)
// with da_map_col(j + s + pw0) == c, i.e. j = q + m * in_w for one |m| <= 2 (in_w > W, so at most one candidate is inside), or -1.
__device__ __forceinline__ int da_map_col_t(int q, int in_w, int W)
{
#pragma unroll
    for (int m = -2; m <= 2; ++m) {
        const int j = q + m * in_w;
        if (j >= 0 && j < W) return j;
    }
    return -1;
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

// One blended strip (kind 0): UN items (strip row, 16-byte channel chunk) per thread are in flight at a time — the stage is bound by the
// latency of the global loads, so loads in flight are what counts; a strip that reads ONE input row (plain convolutions, integer y)
// keeps half the registers per item and runs with a deeper batch.
template <int UN, bool TWO_ROWS, bool SPLIT3>
__device__ __forceinline__ void fill_strip(const StripParams &p, const StripDesc &sd, uint8_t *buf, int gtid, int GT, int j0, int b0, int ch0)
{
    // thread <-> (16-byte channel chunk c, strip rows r8, r8 + GT/8, ...): c, the channel pointer and the store column are fixed per
    // thread; a step of GT/8 strip rows is a fixed step of the store address.  The producers are bound by instruction issue (the address
    // arithmetic of an item), so everything that does not depend on the item is hoisted and all offsets are 32-bit (tensors < 2^31 elements).
    const int c = gtid & 7, r8 = gtid >> 3, rstep = GT >> 3;
    const float *x0 = p.x + (size_t)(sd.r0 >= 0 ? sd.r0 : 0) * p.W * p.C + ch0 + c * 4;
    const float *x1 = p.x + (size_t)(sd.r1 >= 0 ? sd.r1 : 0) * p.W * p.C + ch0 + c * 4;
    const bool has0 = sd.r0 >= 0 && sd.wy0 != 0.f;
    const int ubase = j0 + sd.u0, img_elems = p.H * p.W, nbm = p.NB - 1;
    uint8_t *dst0 = buf + c * p.PS;
    for (int rho0 = r8; rho0 < p.SR; rho0 += rstep * UN) {
        float4 v0[UN], v1[TWO_ROWS ? UN : 1];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rho = rho0 + u * rstep;
            v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (TWO_ROWS) v1[u] = v0[u];
            const int bimg = b0 + (rho & nbm);
            int col = sd.cm * (ubase + (rho >> p.log_nb)) + sd.c0;
            if (p.da == 1) col = da_map_col(col + p.pw0, p.in_w, p.pw0, p.W);
            else if (p.da == 2) col = da_map_col_t(col, p.in_w, p.W);
            if (rho < p.SR && (unsigned)col < (unsigned)p.W && bimg < p.B) {
                const int o = (bimg * img_elems + col) * p.C;
                if (has0) v0[u] = __ldg(reinterpret_cast<const float4 *>(x0 + o));
                if (TWO_ROWS) v1[u] = __ldg(reinterpret_cast<const float4 *>(x1 + o));
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int rho = rho0 + u * rstep;
            if (rho < p.SR) {
                float4 o;
                if (TWO_ROWS) {
                    o.x = fmaf(sd.wy1, v1[u].x, sd.wy0 * v0[u].x);
                    o.y = fmaf(sd.wy1, v1[u].y, sd.wy0 * v0[u].y);
                    o.z = fmaf(sd.wy1, v1[u].z, sd.wy0 * v0[u].z);
                    o.w = fmaf(sd.wy1, v1[u].w, sd.wy0 * v0[u].w);
                } else {
                    o.x = sd.wy0 * v0[u].x; o.y = sd.wy0 * v0[u].y; o.z = sd.wy0 * v0[u].z; o.w = sd.wy0 * v0[u].w;
                }
                uint8_t *dst = dst0 + rho * 16;
                uint4 hi;
                hi.x = f32_to_tf32_rna(o.x); hi.y = f32_to_tf32_rna(o.y); hi.z = f32_to_tf32_rna(o.z); hi.w = f32_to_tf32_rna(o.w);
                *reinterpret_cast<uint4 *>(dst) = hi;
                if (SPLIT3) {
                    uint4 lo;
                    lo.x = f32_to_tf32_rna(o.x - __uint_as_float(hi.x));
                    lo.y = f32_to_tf32_rna(o.y - __uint_as_float(hi.y));
                    lo.z = f32_to_tf32_rna(o.z - __uint_as_float(hi.z));
                    lo.w = f32_to_tf32_rna(o.w - __uint_as_float(hi.w));
                    *reinterpret_cast<uint4 *>(dst + 8 * p.PS) = lo;
                }
            }
        }
    }
}

template <bool SPLIT3, int PAIR>
__global__ void __launch_bounds__(SC_THREADS, 1) strip_conv_kernel(const StripParams p)
{
    using L = StripSmem<SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *strip_ring = smem;
    const int stage_bytes = PAIR * p.strip_bytes;            // one ring stage = the strips of the step's `pair` tiles
    uint8_t *b_ring = smem + round_up(p.NSB * stage_bytes, 1024);
    const int b_stage = L::b_stage(p.Fs);
    const int b_plane = p.Fs * BLOCK_K * 4;
    float *epi = reinterpret_cast<float *>(b_ring + p.NWB * b_stage);
    int *plan_scratch = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(epi) + L::EPI_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(plan_scratch) + L::PLAN_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + L::num_bars(p.NSB, p.NWB));
    const uint32_t sfull0 = smem_u32(bars), sempty0 = sfull0 + 8 * p.NSB;
    const uint32_t bfull0 = sempty0 + 8 * p.NSB, bempty0 = bfull0 + 8 * p.NWB;
    const uint32_t tmem_full0 = bempty0 + 8 * p.NWB, tmem_empty0 = tmem_full0 + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int image = blockIdx.y / p.nsub, sub = blockIdx.y % p.nsub;
    if (tid == 0) strip_stamp(p.flags, 0);                       // kernel start
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
    if (tid == 0) strip_stamp(p.flags, 1);                       // prologue done (barriers, TMEM)

    const bool is_producer = warp > 4 && (warp & 3) != 0;
    if (is_producer) {
        // ================================================ PRODUCERS ================================================
        const int ptid = (((warp >> 2) - 1) * 3 + (warp & 3) - 1) * 32 + lane;             // 0 .. SC_PROD_THREADS-1
        const int group = ptid / p.group_threads, gtid = ptid % p.group_threads, GT = p.group_threads;
        // strip sequence: buffer sb (ring of NSB), its consumer-phase parity, and whose turn it is (groups alternate strips; NSB % G == 0
        // makes a buffer belong to one group) — the same sequence in every role, kept with counters (no divisions in the loops)
        uint32_t sb = 0, sphase = 1, turn = 0;
        int nstamp = 0, cur_rp = -1;
        StripDesc *s_sd = reinterpret_cast<StripDesc *>(plan_scratch + 2 * (2 * SC_MAX_STRIPS + SC_MAX_WINS));
        for (int work = blockIdx.x; work < p.nwork; work += gridDim.x) {
            const int rp = work / p.pairs_per_row, t0 = (work % p.pairs_per_row) * PAIR;
            const RowPlan row = p.rows[rp];
            if (rp != cur_rp) {
                // the row class's strip table goes to shared memory once (one coalesced load instead of a dependent global load per strip)
                named_bar_sync(3, SC_PROD_THREADS);                       // everybody is done with the previous table
                const int nwords = (row.strip_end - row.strip_begin) * (int)(sizeof(StripDesc) / 4);
                const int *src = reinterpret_cast<const int *>(p.strips + row.strip_begin);
                for (int e = ptid; e < nwords; e += SC_PROD_THREADS) reinterpret_cast<int *>(s_sd)[e] = __ldg(src + e);
                named_bar_sync(3, SC_PROD_THREADS);
                cur_rp = rp;
            }
            for (int cc = 0; cc < p.CC; ++cc) {
                for (int si = row.strip_begin; si < row.strip_end; ++si) {
                    const uint32_t my_sb = sb, my_phase = sphase;
                    const bool mine = (int)turn == group;
                    if (++turn == (uint32_t)p.G) turn = 0;
                    if (++sb == (uint32_t)p.NSB) { sb = 0; sphase ^= 1; }
                    if (!mine) continue;
                    const StripDesc sd = s_sd[si - row.strip_begin];
                    mbar_wait_sleep(sempty0 + 8 * my_sb, my_phase);     // producers run ahead: back off, the issuing lane of the MMA warp shares the schedulers
                    const int ch0 = cc * BLOCK_K;
                    #pragma unroll
                    for (int tp = 0; tp < PAIR; ++tp) {
                    const int tq = t0 + tp;
                    if (tq >= tiles_per_row) break;                         // odd tile count: the pair's second accumulator is never stored
                    const int j0 = (tq / p.tiles_b) * p.TW, b0 = (tq % p.tiles_b) * p.NB;
                    uint8_t *buf = strip_ring + my_sb * stage_bytes + tp * p.strip_bytes;
                    if (sd.kind == 0) {
                        if (sd.r1 >= 0 && sd.wy1 != 0.f) fill_strip<SC_UNROLL2, true, SPLIT3>(p, sd, buf, gtid, GT, j0, b0, ch0);
                        else fill_strip<SC_UNROLL1, false, SPLIT3>(p, sd, buf, gtid, GT, j0, b0, ch0);
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
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sfull0 + 8 * my_sb);
                    if (gtid == 0 && work == (int)blockIdx.x && nstamp < 12) strip_stamp(p.flags, 40 + group * 3 + (nstamp++ % 3) );
                }
            }
        }
    } else if (warp < 4) {
        // ================================================ EPILOGUE ================================================
        const int wq = warp - SC_EPI_WARP0, etid = tid - SC_EPI_WARP0 * 32;
        const bool vec_ok = (p.F % 4) == 0 && (p.ldF % 4) == 0;
        const int f_base = image * p.F + sub * p.Fs;                     // first filter of this slice in the layer
        const int Fv = min(p.Fs, p.F - sub * p.Fs);                      // valid filters of this slice (the last one may be ragged)
        uint32_t it = 0;
        for (int work = blockIdx.x; work < p.nwork; work += gridDim.x, ++it) {
            const int rp = work / p.pairs_per_row, t0 = (work % p.pairs_per_row) * PAIR;
            const RowPlan row = p.rows[rp];
            const uint32_t acc = it & 1;
            const bool no_terms = row.strip_begin == row.strip_end;      // a row class nothing contributes to: the accumulator is not written
            mbar_wait_sleep(tmem_full0 + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            if (etid == 0 && it == 0) strip_stamp(p.flags, 4);   // first accumulator complete
            for (int tp = 0; tp < PAIR; ++tp) {
            const int tq = t0 + tp;
            if (tq >= tiles_per_row) break;
            const int j0 = (tq / p.tiles_b) * p.TW, b0 = (tq % p.tiles_b) * p.NB;
            const uint32_t taddr = tmem_base + (acc * (uint32_t)PAIR + (uint32_t)tp) * (uint32_t)p.Fs + ((uint32_t)(wq * 32) << 16);
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
                        if (p.bias && f + u < Fv) bv[u] = __ldg(p.bias + f_base + f + u);
                    float s1[4] = { 0.f, 0.f, 0.f, 0.f }, s2[4] = { 0.f, 0.f, 0.f, 0.f };
                    if (f < Fv && bimg < p.B) {
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
                                        if (f + u < Fv) rv[u] = __ldg(p.residual + go + u);
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
                                    if (f + u < Fv) p.y[go + u] = v[u];
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
                        if ((lane >> 2) < p.NB && f < Fv && bimg < p.B) {
                            double *st = p.stats + ((size_t)bimg * p.ldF + f_base + f) * 2;
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                if (f + u < Fv) {
                                    atomicAdd(st + 2 * u, (double)s1[u]);
                                    atomicAdd(st + 2 * u + 1, (double)s2[u]);
                                }
                        }
                    }
                }
                named_bar_sync(2, 128);   // the staging tile is rewritten by the next column pass
            }
            }
            tc_fence_before();
            __syncwarp();
            if (etid == 0 && it == 0) strip_stamp(p.flags, 5);   // first epilogue done
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * acc);
        }
    } else if (warp == SC_WARP_MMA) {
        // ================================================ MMA ISSUER ================================================
        // The row class's strip / window tables are staged in shared memory by the whole warp (one coalesced load per table) whenever the
        // row class changes: the issuing lane would otherwise pay a dependent global load per window.
        int2 *s_strip = reinterpret_cast<int2 *>(plan_scratch);
        int *s_start = plan_scratch + 2 * SC_MAX_STRIPS;
        const uint32_t idesc = umma_idesc_tf32(BLOCK_M, (uint32_t)p.Fs);
        // descriptor words that never change: A = no-swizzle K-major (LBO = plane stride, SBO = 128 B between 8-row groups), B = 128-byte
        // swizzled K-major (SBO = 1024 B); both version 1.  The issuing lane only adds row / k offsets to the low words.
        constexpr uint32_t A_HI = (128u >> 4) | (1u << 14), B_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t a_lbo = (((uint32_t)p.PS >> 4) & 0x3FFFu) << 16, a_kstep = (uint32_t)p.PS >> 3, a_lo_plane = (uint32_t)p.PS >> 1;
        const uint32_t b_lo_plane = (uint32_t)b_plane >> 4, strip_units = (uint32_t)p.strip_bytes >> 4;
        const uint32_t strip0 = smem_u32(strip_ring), bring0 = smem_u32(b_ring);
        uint32_t sb = 0, sphase = 0, ws = 0, wphase = 0, it = 0;
        int cur_rp = -1, nstrips = 0, w_first = 0;
        for (int work = blockIdx.x; work < p.nwork; work += gridDim.x, ++it) {
            const int rp = work / p.pairs_per_row;
            if (rp != cur_rp) {
                __syncwarp();
                const RowPlan row = p.rows[rp];
                nstrips = row.strip_end - row.strip_begin;
                int w_last = 0;
                w_first = 0;
                if (nstrips > 0) { w_first = __ldg(&p.strips[row.strip_begin].win_begin); w_last = __ldg(&p.strips[row.strip_end - 1].win_end); }
                for (int e = lane; e < nstrips; e += 32)
                    s_strip[e] = make_int2(__ldg(&p.strips[row.strip_begin + e].win_begin), __ldg(&p.strips[row.strip_begin + e].win_end));
                for (int e = lane; e < w_last - w_first; e += 32) s_start[e] = __ldg(&p.wins[w_first + e].start_row);
                cur_rp = rp;
                __syncwarp();
            }
            if (lane == 0 && it == 0) strip_stamp(p.flags, 2);   // MMA warp: row class staged
            {
                // the whole warp walks the loops and the barrier waits (warp-uniform control flow); one elected lane issues
                const uint32_t acc = it & 1;
                mbar_wait(tmem_empty0 + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)(PAIR * p.Fs);
                uint32_t first = 0;
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int si = 0; si < nstrips; ++si) {
                        const int2 wr = s_strip[si];
                        mbar_wait(sfull0 + 8 * sb, sphase);
                        tc_fence_after();
                        // low descriptor word of the strip: start address (16-byte units) | plane stride; a window adds its start row
                        const uint32_t a_lo0 = (((strip0 + sb * (uint32_t)stage_bytes) & 0x3FFFFu) >> 4) | a_lbo;
                        for (int wi = wr.x; wi < wr.y; ++wi) {
                            const uint32_t a_lo = a_lo0 + (uint32_t)s_start[wi - w_first];
                            mbar_wait(bfull0 + 8 * ws, wphase);
                            tc_fence_after();
                            const uint32_t b_lo = (((bring0 + ws * (uint32_t)b_stage) & 0x3FFFFu) >> 4) | (1u << 16);
                            if (elect_one()) {
                                for (int tp = 0; tp < PAIR; ++tp) {           // the same weight tile against the strip of every tile of the step
                                    const uint32_t a_t = a_lo + (uint32_t)tp * strip_units, d_t = d_tmem + (uint32_t)(tp * p.Fs);
#pragma unroll
                                    for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                                        const uint64_t da = ((uint64_t)A_HI << 32) | (a_t + ks * a_kstep);
                                        const uint64_t db = ((uint64_t)B_HI << 32) | (b_lo + ks * 2u);
                                        umma_tf32(d_t, da, db, idesc, first | (uint32_t)ks);
                                        if (SPLIT3) {
                                            const uint64_t da_lo = ((uint64_t)A_HI << 32) | (a_t + ks * a_kstep + a_lo_plane);
                                            const uint64_t db_lo = ((uint64_t)B_HI << 32) | (b_lo + ks * 2u + b_lo_plane);
                                            umma_tf32(d_t, da_lo, db, idesc, 1);
                                            umma_tf32(d_t, da, db_lo, idesc, 1);
                                        }
                                    }
                                }
                                umma_commit(bempty0 + 8 * ws);
                            }
                            __syncwarp();
                            first = 1;
                            if (++ws == (uint32_t)p.NWB) { ws = 0; wphase ^= 1; }
                        }
                        if (elect_one()) umma_commit(sempty0 + 8 * sb);
                        __syncwarp();
                        if (++sb == (uint32_t)p.NSB) { sb = 0; sphase ^= 1; }
                    }
                if (elect_one()) umma_commit(tmem_full0 + 8 * acc);
                __syncwarp();
                if (lane == 0 && it == 0) strip_stamp(p.flags, 3);            // last MMA of the first tile issued
            }
        }
    } else if (warp == SC_WARP_WLOAD) {
        // ================================================ WEIGHT LOADER ================================================
        int2 *s_strip = reinterpret_cast<int2 *>(plan_scratch + 2 * SC_MAX_STRIPS + SC_MAX_WINS);
        int *s_tile = plan_scratch + 2 * SC_MAX_STRIPS + SC_MAX_WINS + 2 * SC_MAX_STRIPS;
        const size_t tile_plane = (size_t)p.tile_rows * BLOCK_K * 4;              // one plane of one packed tile
        const size_t tile_bytes = (size_t)L::PLANES * tile_plane;
        const uint8_t *base = p.packed + (size_t)image * p.image_stride + (size_t)sub * p.Fs * (BLOCK_K * 4);
        uint32_t ws = 0, wphase = 1;                       // wphase: parity of the stage's previous consumer phase
        int cur_rp = -1, nstrips = 0, w_first = 0;
        for (int work = blockIdx.x; work < p.nwork; work += gridDim.x) {
            const int rp = work / p.pairs_per_row;
            if (rp != cur_rp) {
                __syncwarp();
                const RowPlan row = p.rows[rp];
                nstrips = row.strip_end - row.strip_begin;
                int w_last = 0;
                w_first = 0;
                if (nstrips > 0) { w_first = __ldg(&p.strips[row.strip_begin].win_begin); w_last = __ldg(&p.strips[row.strip_begin + nstrips - 1].win_end); }
                for (int e = lane; e < nstrips; e += 32)
                    s_strip[e] = make_int2(__ldg(&p.strips[row.strip_begin + e].win_begin), __ldg(&p.strips[row.strip_begin + e].win_end));
                for (int e = lane; e < w_last - w_first; e += 32) s_tile[e] = __ldg(&p.wins[w_first + e].wtile0) * p.wmul;
                cur_rp = rp;
                __syncwarp();
            }
            if (lane == 0) {
                for (int cc = 0; cc < p.CC; ++cc)
                    for (int si = 0; si < nstrips; ++si) {
                        const int2 wr = s_strip[si];
                        for (int wi = wr.x; wi < wr.y; ++wi) {
                            const int wt = s_tile[wi - w_first] + cc * p.wcc_stride;
                            mbar_wait_sleep(bempty0 + 8 * ws, wphase);
                            const uint32_t dst = smem_u32(b_ring + ws * b_stage);
                            const uint8_t *src = base + (size_t)wt * tile_bytes;
                            mbar_arrive_expect_tx(bfull0 + 8 * ws, (uint32_t)b_stage);
                            if (work == (int)blockIdx.x && cc == 0 && si == 0 && wi == wr.x) strip_stamp(p.flags, 7);   // first weight tile requested
                            bulk_g2s(dst, src, (uint32_t)b_plane, bfull0 + 8 * ws);
                            if (SPLIT3) bulk_g2s(dst + b_plane, src + tile_plane, (uint32_t)b_plane, bfull0 + 8 * ws);
                            if (++ws == (uint32_t)p.NWB) { ws = 0; wphase ^= 1; }
                        }
                    }
            }
            __syncwarp();
        }
    }

    __syncthreads();
    if (warp == SC_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
    if (tid == 0) strip_stamp(p.flags, 6);                       // kernel end
}

// ---------------------------------------------------------------------------------------------------------------------
// effective weights of a distortion-aware layer: tile (window, cc) = sum over the window's terms of coef * kernel[tap*C + cc*32 + kk, n],
// written as the K-major 128-byte-swizzled image the MMA reads (plane 0 = tf32(rna(v)), plane 1 (3xTF32) = tf32(rna(v - hi)))
// ---------------------------------------------------------------------------------------------------------------------
// transposed != 0 (data gradient): tile rows are the layer's INPUT channels (N = C, padded to Np), the contraction runs over its filters
// (KC = F / 32 chunks): tile (window, fc)[c][kk] = sum coef * kernel[tap*C + c, fc*32 + kk].
__global__ void strip_weff_pack_t_kernel(const float *__restrict__ kernel, float *__restrict__ packed, const int *__restrict__ term_begin,
                                         const WeffTerm *__restrict__ terms, int nwins, int C, int F, int Np, int planes)
{
    // one thread per (tile row n, 16-byte chunk): a 128-bit load per term (4 consecutive filters of kernel row tap*C + n), a 128-bit store;
    // the 8 threads of a row write its 128 bytes
    const int KC = F / BLOCK_K;
    const long total = (long)nwins * KC * Np * 8;
    const size_t tile_floats = (size_t)Np * BLOCK_K;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int chunk = (int)(e & 7);
        const int n = (int)((e >> 3) % Np);
        const long wc = e / ((long)Np * 8);
        const int fc = (int)(wc % KC), w = (int)(wc / KC);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < C) {
            const int tb = __ldg(term_begin + w), te = __ldg(term_begin + w + 1);
            for (int q = tb; q < te; ++q) {
                const WeffTerm t = terms[q];
                const float4 kv = __ldg(reinterpret_cast<const float4 *>(kernel + ((size_t)t.tap * C + n) * F + fc * BLOCK_K + 4 * chunk));
                v.x = fmaf(t.coef, kv.x, v.x); v.y = fmaf(t.coef, kv.y, v.y); v.z = fmaf(t.coef, kv.z, v.z); v.w = fmaf(t.coef, kv.w, v.w);
            }
        }
        uint4 hi;
        hi.x = f32_to_tf32_rna(v.x); hi.y = f32_to_tf32_rna(v.y); hi.z = f32_to_tf32_rna(v.z); hi.w = f32_to_tf32_rna(v.w);
        uint8_t *tile = reinterpret_cast<uint8_t *>(packed + (size_t)wc * planes * tile_floats);
        const uint32_t o = sw128_offset((uint32_t)n, (uint32_t)chunk);
        *reinterpret_cast<uint4 *>(tile + o) = hi;
        if (planes == 2) {
            uint4 lo;
            lo.x = f32_to_tf32_rna(v.x - __uint_as_float(hi.x)); lo.y = f32_to_tf32_rna(v.y - __uint_as_float(hi.y));
            lo.z = f32_to_tf32_rna(v.z - __uint_as_float(hi.z)); lo.w = f32_to_tf32_rna(v.w - __uint_as_float(hi.w));
            *reinterpret_cast<uint4 *>(tile + tile_floats * 4 + o) = lo;
        }
    }
}

// forward flavour: one block per (window, 32-channel chunk) tile.  The kernel variable is read along its filters (coalesced), the tile is
// written along its k values (128-byte rows): the transpose goes through shared memory.
__global__ void __launch_bounds__(256) strip_weff_pack_kernel(const float *__restrict__ kernel, float *__restrict__ packed,
                                                             const int *__restrict__ term_begin, const WeffTerm *__restrict__ terms, int C,
                                                             int CC, int F, int Fp, int planes)
{
    extern __shared__ float wsm[];                           // [32][Fp + 1]
    const int wc = blockIdx.x, cc = wc % CC, w = wc / CC, ld = Fp + 1;
    const int tb = __ldg(term_begin + w), te = __ldg(term_begin + w + 1);
    for (int e = threadIdx.x; e < BLOCK_K * Fp; e += blockDim.x) {
        const int kk = e / Fp, n = e - kk * Fp;
        float v = 0.f;
        if (n < F)
            for (int q = tb; q < te; ++q) {
                const WeffTerm t = terms[q];
                v = fmaf(t.coef, __ldg(kernel + ((size_t)t.tap * C + cc * BLOCK_K + kk) * F + n), v);
            }
        wsm[kk * ld + n] = v;
    }
    __syncthreads();
    const size_t tile_floats = (size_t)Fp * BLOCK_K;
    uint8_t *tile = reinterpret_cast<uint8_t *>(packed + (size_t)wc * planes * tile_floats);
    for (int e = threadIdx.x; e < Fp * 8; e += blockDim.x) {
        const int chunk = e & 7, n = e >> 3;
        const float v0 = wsm[(4 * chunk) * ld + n], v1 = wsm[(4 * chunk + 1) * ld + n], v2 = wsm[(4 * chunk + 2) * ld + n],
                    v3 = wsm[(4 * chunk + 3) * ld + n];
        uint4 hi;
        hi.x = f32_to_tf32_rna(v0); hi.y = f32_to_tf32_rna(v1); hi.z = f32_to_tf32_rna(v2); hi.w = f32_to_tf32_rna(v3);
        const uint32_t o = sw128_offset((uint32_t)n, (uint32_t)chunk);
        *reinterpret_cast<uint4 *>(tile + o) = hi;
        if (planes == 2) {
            uint4 lo;
            lo.x = f32_to_tf32_rna(v0 - __uint_as_float(hi.x)); lo.y = f32_to_tf32_rna(v1 - __uint_as_float(hi.y));
            lo.z = f32_to_tf32_rna(v2 - __uint_as_float(hi.z)); lo.w = f32_to_tf32_rna(v3 - __uint_as_float(hi.w));
            *reinterpret_cast<uint4 *>(tile + tile_floats * 4 + o) = lo;
        }
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
    p.tile_rows = Fp; p.F = a.F; p.image_stride = a.image_stride;
    const int nimages = a.nimages > 1 ? a.nimages : 1;
    // Filter sub-slices: every CTA builds its own strips, so slicing the filters costs producer work — worth it only while SMs would
    // otherwise idle (the trunk of 32x128 panoramas at B = 32 has 64 tiles; the 4x16 maps of sunRadNet / the discriminator have 16).
    // 3xTF32 keeps a second plane of both operands: slices of at most 64 filters leave room for the strip ring.
    int nsub = 1;
    while (Fp % (32 * nsub) == 0 && Fp / (2 * nsub) >= 32 && (long)p.ntiles * nimages * nsub * 2 <= num_sms_cached() + num_sms_cached() / 8)
        nsub *= 2;
    if (pl.max_strips_row > SC_MAX_STRIPS || pl.max_wins_row > SC_MAX_WINS) return SKY_ERR_UNSUPPORTED;
    // ring depths: producer groups own whole strip buffers (NSB % G == 0, see the parity argument in the band kernel).  When nothing
    // fits (3xTF32 keeps a second plane of both operands) the filters are sliced further.
    const int budget = 227 * 1024;
    int best_nsb = 0, best_nwb = 0, best_g = 0;
    // two tiles of one row class per CTA step (M = 256 through two accumulators) when there is more than a wave of tile pairs: every
    // weight tile is fetched once per pair and every barrier round trip of the issuing warp covers 8 MMAs.  Measured (B = 64, 64x256):
    // 7x7 32->32 1.18 -> 0.91 ms (N = 32 MMAs are issue-bound), 128->128 k3 unchanged (269 vs 264 TFLOP/s: producer-bound, and the
    // ring holds half as many stages) -> only for narrow layers
    const int tiles_per_row = p.tiles_x * p.tiles_b;
    p.pair = 1;
    if ((!SPLIT3 || getenv("SKY_STRIP_PAIR3")) && nsub == 1 && nimages == 1 && Fp <= SC_PAIR_MAX_N && tiles_per_row >= 2 && (long)pl.nrows * ((tiles_per_row + 1) / 2) >= 2L * num_sms_cached() &&
        !(a.flags & SKY_EPI_NO_PAIR) && !getenv("SKY_STRIP_NO_PAIR"))
        p.pair = 2;
    for (;;) {
        p.nsub = nsub; p.Fs = Fp / nsub;
        p.tmem_cols = 32;
        while ((int)p.tmem_cols < 2 * p.pair * p.Fs) p.tmem_cols <<= 1;
        const int gs[3] = { 4, 2, 1 };
        for (int gi = 0; gi < 3 && !best_nsb && p.tmem_cols <= 512; ++gi) {
            const int G = gs[gi];
            for (int nsb = 2 * G; nsb >= G && nsb >= 2 && !best_nsb; nsb -= G)
                for (int nwb = 6; nwb >= 3; --nwb)
                    if (L::total_bytes(p.PS, p.Fs, nsb, nwb, p.pair) <= budget) { best_nsb = nsb; best_nwb = nwb; best_g = G; break; }
        }
        if (best_nsb) break;
        if (p.pair == 2) { p.pair = 1; continue; }           // pairs do not fit: single tiles
        if (Fp % (32 * nsub) != 0) return SKY_ERR_UNSUPPORTED;
        nsub *= 2;                                            // nothing fits: slice the filters further
    }
    if (const char *ring = getenv("SKY_STRIP_RING")) {        // tuning: "G,NSB,NWB" replaces the choice above when it fits
        int g = 0, nsb = 0, nwb = 0;
        if (sscanf(ring, "%d,%d,%d", &g, &nsb, &nwb) == 3 && (g == 1 || g == 2 || g == 4) && nsb >= 2 && nsb % g == 0 && nwb >= 2 && nwb <= 16 &&
            L::total_bytes(p.PS, p.Fs, nsb, nwb, p.pair) <= budget) { best_g = g; best_nsb = nsb; best_nwb = nwb; }
    }
    if (getenv("SKY_STRIP_VERBOSE"))
        fprintf(stderr, "strip launch: C=%d F=%d Fs=%d nsub=%d pair=%d PS=%d strip_bytes=%d b_stage=%d G=%d NSB=%d NWB=%d smem=%d nrows=%d tiles/row=%d\n", a.C, a.F, p.Fs, nsub,
                p.pair, p.PS, L::strip_bytes(p.PS), L::b_stage(p.Fs), best_g, best_nsb, best_nwb, L::total_bytes(p.PS, p.Fs, best_nsb, best_nwb, p.pair), pl.nrows, tiles_per_row);
    const int nslices = nimages * nsub;
    p.pairs_per_row = (tiles_per_row + p.pair - 1) / p.pair;
    p.nwork = pl.nrows * p.pairs_per_row;
    p.NSB = best_nsb; p.NWB = best_nwb; p.G = best_g; p.group_threads = SC_PROD_THREADS / best_g;
    const int smem = L::total_bytes(p.PS, p.Fs, p.NSB, p.NWB, p.pair);
    int gx = num_sms_cached() / nslices;
    if (gx < 1) gx = 1;
    if (gx > p.nwork) gx = p.nwork;
    if (p.pair == 2) {
        SKY_ENSURE_DYN_SMEM((strip_conv_kernel<SPLIT3, 2>), 227 * 1024);
        strip_conv_kernel<SPLIT3, 2><<<dim3(gx, nslices), SC_THREADS, smem, a.stream>>>(p);
    } else {
        SKY_ENSURE_DYN_SMEM((strip_conv_kernel<SPLIT3, 1>), 227 * 1024);
        strip_conv_kernel<SPLIT3, 1><<<dim3(gx, nslices), SC_THREADS, smem, a.stream>>>(p);
    }
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
    SKY_REQUIRE(pl->terms != nullptr, SKY_ERR_CUDA, "strip plan tables are not on the device");
    const int Fp = f_pad_of(F), CC = C / BLOCK_K, planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    strip_weff_pack_kernel<<<pl->nwins * CC, 256, BLOCK_K * (Fp + 1) * sizeof(float), (cudaStream_t)stream>>>(kernel, (float *)packed, pl->term_begin,
                                                                                                             pl->terms, C, CC, F, Fp, planes);
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

// ---- data gradient through the transposed plan ------------------------------------------------------------------------------------
extern "C" size_t sky_da_strip_weight_bytes_t(const float *offsets_host, int h, int w, int C, int F, int k, int math_mode)
{
    if (!offsets_host || C <= 0 || F <= 0 || F % BLOCK_K != 0 || C > 256) return 0;
    const StripPlan *pl = nullptr;
    if (get_plan_da(offsets_host, h, w, k, &pl, false, true) != SKY_OK) return 0;
    if (pl->max_strips_row > SC_MAX_STRIPS || pl->max_wins_row > SC_MAX_WINS) return 0;
    const int planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    return (size_t)pl->nwins * (F / BLOCK_K) * planes * f_pad_of(C) * BLOCK_K * sizeof(float);
}

extern "C" int sky_da_strip_pack_weights_t(const float *kernel, const float *offsets_host, void *packed, int h, int w, int C, int F, int k,
                                           int math_mode, void *stream)
{
    SKY_REQUIRE(kernel && offsets_host && packed, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(C > 0 && F > 0 && F % BLOCK_K == 0 && C <= 256, SKY_ERR_UNSUPPORTED, "transposed strip weights need F %% 32 == 0 and C <= 256 (C=%d F=%d)", C, F);
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl, true, true);
    if (rc != SKY_OK) return rc;
    const int Np = f_pad_of(C), KC = F / BLOCK_K, planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    const long total = (long)pl->nwins * KC * Np * 8;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    strip_weff_pack_t_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kernel, (float *)packed, pl->term_begin, pl->terms, pl->nwins, C, F, Np, planes);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_conv2d_bwd_data_strip(const float *dy, const float *offsets_host, const void *packed_t, float *dx, const float *aux, int B,
                                            int h, int w, int C, int F, int k, int accumulate, int epilogue_flags, float slope, int math_mode,
                                            void *stream)
{
    SKY_REQUIRE(dy && offsets_host && packed_t && dx, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension (B=%d h=%d w=%d C=%d F=%d)", B, h, w, C, F);
    SKY_REQUIRE(k % 2 == 1 && k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    SKY_REQUIRE(F % BLOCK_K == 0 && C <= 256, SKY_ERR_UNSUPPORTED, "the strip data gradient needs F %% 32 == 0 and C <= 256 (C=%d F=%d)", C, F);
    SKY_REQUIRE(!(epilogue_flags & ~(SKY_EPI_MASK | SKY_EPI_NO_PAIR)), SKY_ERR_INVALID, "the data gradient takes SKY_EPI_MASK only");
    SKY_REQUIRE(!(epilogue_flags & SKY_EPI_MASK) || (aux && !accumulate), SKY_ERR_INVALID, "SKY_EPI_MASK needs its mask source and excludes accumulate");
    SKY_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)packed_t & 15) == 0, SKY_ERR_INVALID, "dy, dx and packed must be 16-byte aligned");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    const StripPlan *pl = nullptr;
    int rc = get_plan_da(offsets_host, h, w, k, &pl, true, true);
    if (rc != SKY_OK) return rc;
    StripLaunch s;
    s.x = dy; s.packed = packed_t; s.bias = nullptr; s.y = dx; s.stats = nullptr;
    s.residual = accumulate ? dx : aux;                      // accumulate: each output element is read and rewritten by the one thread that owns it
    s.B = B; s.H = h; s.W = w; s.C = F; s.OH = h; s.OW = w; s.F = C; s.ldF = C; s.nimages = 1; s.image_stride = 0;
    s.k = k; s.flags = (accumulate ? SKY_EPI_RESIDUAL : (epilogue_flags & SKY_EPI_MASK)) | (epilogue_flags & SKY_EPI_NO_PAIR); s.math_mode = math_mode; s.da = 2; s.slope = slope;
    s.stream = (cudaStream_t)stream;
    int pht, pwt;
    pad_axis(h, k, &s.ph0, &pht);
    pad_axis(w, k, &s.pw0, &pwt);
    s.in_h = h + pht; s.in_w = w + pwt;
    rc = launch_strip_any(s, *pl);
    SKY_REQUIRE(rc != SKY_ERR_UNSUPPORTED, SKY_ERR_UNSUPPORTED, "the strip kernel does not cover this data gradient (C=%d, strip rows %d)", C, pl->SR);
    return rc;
}

extern "C" int sky_debug_strip_trace(unsigned long long *host_out64)
{
    SKY_CHECK_CUDA(cudaMemcpyFromSymbol(host_out64, sky::g_strip_trace, sizeof(unsigned long long) * 64));
    return SKY_OK;
}
