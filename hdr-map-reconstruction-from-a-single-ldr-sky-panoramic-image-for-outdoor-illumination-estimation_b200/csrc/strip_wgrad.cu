// Weight gradient on the row-strip formulation (strip_conv.cuh).
//
// Forward:   y[i, j, :] = sum over strips (kernel rows) a, windows (integer column shifts) s of   V_{i,a}[j + s, :] . Weff_{i,a,s}
// so         dWeff_{i,a,s}[c, f] = sum over panoramas b and columns j of   V_{i,a}[b, j + s, c] * dy[b, i, j, f]
// and        dW[tap, c, f]      = sum over the (i, a, s) whose effective weight holds the tap of   coef * dWeff_{i,a,s}[c, f]
// (what TF autodiff derives from distortion_aware_ops.py:94-121: matmul^T, then the bilinear factors and gather_nd's scatter, here
// regrouped per window; a plain SAME convolution, ops.py:41, is the case coef = 1, one tap per window).
//
// The contraction index is the pixel, the ROW index of both natural layouts, so both operands are MN-major (SWIZZLE_128B with a 32-byte
// base, as in conv_bwd.cu): A = the blended strip [strip row][channel], written ONCE per pixel tile for all the windows of the strip
// (the im2col producer of conv2d_wgrad_kernel re-gathers it per tap); B = the dy tile [pixel][filter].  Strip row = column * 8 +
// panorama: a column shift is 8 rows = 1024 bytes, i.e. a window is the same A tile at another (swizzle-aligned) start address.
// Accumulators (one [C x F] block per window) stay in TMEM over all the pixel tiles a CTA owns of a unit and are added to dW with the
// tap coefficients by vector atomics.  32- / 64-channel layers interleave the operand rows as (column * chunks + 32-channel chunk) * 8 +
// panorama, so that the four MN atoms of one MMA are the chunks of four / two windows at consecutive column shifts (atom stride 1024 B).
// dy arrives by TMA through a tensor map with dimensions (filter, panorama, column, row) and CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B — the
// box lands as rows = column * 8 + panorama in exactly the MN-major operand layout.  Work: every accumulation unit (strip x windows that
// fit in TMEM) is split over P CTAs of TP pixel tiles, U units run at a time (launch_wgrad_strip's cost model), one drain per wave.
//
// Persistent, warp-specialised: two producer groups of 7 warps that take alternate pixel tiles (second input row by cp.async straight
// into the operand, first row by vector loads, blend in place; dy tile by TMA + in-place TF32 rounding), one MMA warp (elect-issued), 4 drain warps.
#include <stdlib.h>
#include <string.h>

#include "strip_conv.cuh"

namespace sky {

namespace {

constexpr int SW_THREADS = 20 * 32, SW_WARP_MMA = 4;
constexpr int SW_GROUPS = 2, SW_GROUP_WARPS = 7, SW_PROD_THREADS = SW_GROUP_WARPS * 32;   // two producer groups take alternate items (stages)
constexpr int SW_NB = 8;                                         // panoramas per pixel tile (tile = TW columns x 8 panoramas, KT = 8 TW pixels)
constexpr int SW_RP = SW_PROD_THREADS / 8;                       // strip rows (pixels) one pass of a producer group covers
constexpr int SW_MAX_GROUPS = 16;

struct SwParams {
    const float *x, *dy;
    float *dw;
    const RowPlan *rows;
    const StripDesc *strips;
    const WinDesc *wins;
    const int *term_begin;
    const WeffTerm *terms;
    const WgUnit *units;
    const WgGroup *groups;
    int B, H, W, C, OH, OW, F;
    int weff, da, in_h, in_w, ph0, pw0, k;
    int ncols, ocs;
    int SR, nreg, wpg;                 // strip rows (column * 8 + panorama); 32-channel chunks of the A tile that are written; windows per MMA group
    int nch;                           // A layout: 0 = one region of SR rows per 32-channel chunk (C >= 96); 1 / 2 = rows interleaved as
                                       // (column * nch + chunk) * 8 + panorama (C = 32 / 64), so that the MN atoms of an MMA are the chunks
                                       // of wpg = 4 / nch windows at consecutive column shifts, 1024 bytes apart
    int a_lbo, a_kstep;                // bytes between the MN atoms of A; 16-byte units between the K = 8 steps of a window
    int NF, ncc, nfc;                  // filters per MMA (N), channel chunks of 128, filter chunks of NF
    int TW, KT;                        // columns / pixels of a tile
    int tiles_x, tiles_b, ntiles;      // pixel tiles of one output row
    int nuidx, U, P, TP, nwaves;       // schedule: waves of U concurrent (unit, chunk) accumulations, each split over P CTAs of TP tiles
    int stages, a_bytes, stage_bytes;
    int trace;
    int dy_tma;                        // the dy tile arrives by TMA (F % 4 == 0) and is rounded to TF32 in place
    uint32_t tmem_cols;
};

__device__ __forceinline__ void sw_red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// MN-major TF32 operand, SWIZZLE_128B with a 32-byte base: atoms of [4 k-rows x 128 B of MN]; lbo = bytes between MN atoms (32 elements
// each), sbo = bytes between k atoms (4 rows each)
__device__ __forceinline__ uint32_t sw_offset(uint32_t row, uint32_t chunk16)
{
    return row * 128u + ((((chunk16 >> 1) ^ (row & 3u)) << 5) | ((chunk16 & 1u) << 4));
}
// byte offset of (strip row rho = column * 8 + panorama, 32-channel chunk r, 16-byte chunk c16) in the A tile
template <class P>
__device__ __forceinline__ uint32_t sw_a_off(const P &p, int rho, int r, int c16)
{
    if (p.nch) return sw_offset((uint32_t)((((rho >> 3) * p.nch + r) << 3) | (rho & 7)), (uint32_t)c16);
    return (uint32_t)(r * (p.SR * 128)) + sw_offset((uint32_t)rho, (uint32_t)c16);
}
__device__ __forceinline__ uint4 sw_tf32x4(float4 v)
{
    uint4 u;
    u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
    return u;
}

// Timeline probe (debug: SKY_WGRAD_TRACE=1): %globaltimer stamps of CTA 0 — MMA warp slots 4i .. 4i+2 (before / after the full wait,
// after the commits of its i-th item, i < 12), producer thread 0 slots 48 + 4i .. (stage free, strip written, dy tile written),
// 100 / 101 first drain begin / end, 102 / 103 kernel begin / end, 104 + i: the MMA warp has issued item 32 i.
__device__ unsigned long long g_wgrad_trace[128];
__device__ __forceinline__ void sw_stamp(int on, int slot)
{
    if (on && blockIdx.x == 0 && slot < 128) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_wgrad_trace[slot] = t;
    }
}

// Work of CTA b = (ub, pb) = (b / P, b % P): in wave w the tiles [pb * TP, pb * TP + TP) of accumulation w * U + ub.  The CTAs that run
// at the same time work on U neighbouring (output row, kernel row) strips, so the input / dy rows they share stay in L2.
struct ItemPos { int uidx, unit, cc, fc, j0, b0, tile; };
__device__ __forceinline__ bool item_pos(const SwParams &p, int ls, ItemPos &q)
{
    const int w = ls / p.TP, tt = ls - w * p.TP;
    const int ub = blockIdx.x / p.P, pb = blockIdx.x - ub * p.P;
    q.uidx = w * p.U + ub;
    q.tile = pb * p.TP + tt;
    if (q.uidx >= p.nuidx || q.tile >= p.ntiles) return false;
    const int per = p.ncc * p.nfc;
    q.unit = q.uidx / per;
    const int rem = q.uidx - q.unit * per;
    q.cc = rem / p.nfc; q.fc = rem - q.cc * p.nfc;
    q.j0 = (q.tile / p.tiles_b) * p.TW; q.b0 = (q.tile % p.tiles_b) * SW_NB;
    return true;
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, bool valid)
{
    const int sz = valid ? 16 : 0;                                     // 0: the 16 bytes are zero-filled, src is not read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One blended strip: thread <-> (16-byte chunk c16 of a 128-byte row, rows rr + 60 * pass, all NREG 32-channel regions).  The stage is
// bound by the bytes in flight per SM, not by arithmetic: the second input row travels by cp.async straight to the operand's final
// (swizzled) position — no registers — while the first row's vector loads are all in flight in registers; the blend then runs in place.
template <int NREG, int NPASS>
__device__ __forceinline__ void sw_fill_strip(const SwParams &p, const StripDesc &sd, uint8_t *a_tile, int c16, int rr, int j0, int b0, int ch0)
{
    // rr = the thread's first strip row of this call (rows rr, rr + SW_RP, ...)
    const bool has0 = sd.r0 >= 0 && sd.wy0 != 0.f, has1 = sd.r1 >= 0 && sd.wy1 != 0.f;
    const float *x0 = p.x + (size_t)(sd.r0 >= 0 ? sd.r0 : 0) * p.W * p.C + ch0;
    const float *x1 = p.x + (size_t)(sd.r1 >= 0 ? sd.r1 : 0) * p.W * p.C + ch0;
    const int ubase = j0 + sd.u0, img_elems = p.H * p.W;
    int off[NPASS];
#pragma unroll
    for (int ps = 0; ps < NPASS; ++ps) {
        const int rho = rr + ps * SW_RP;
        const int bimg = b0 + (rho & (SW_NB - 1));
        int col = sd.cm * (ubase + (rho >> 3)) + sd.c0;
        if (p.da) col = da_map_col(col + p.pw0, p.in_w, p.pw0, p.W);
        const bool ok = (unsigned)col < (unsigned)p.W && bimg < p.B;
        off[ps] = ok ? (bimg * img_elems + col) * p.C : -1;
    }
    if (has1) {
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps)
            if (rr + ps * SW_RP < p.SR) {
#pragma unroll
                for (int r = 0; r < NREG; ++r)
                    cp_async16(smem_u32(a_tile + sw_a_off(p, rr + ps * SW_RP, r, c16)), x1 + (off[ps] >= 0 ? off[ps] : 0) + r * 32, off[ps] >= 0);
            }
    }
    float4 v0[NPASS][NREG];
#pragma unroll
    for (int ps = 0; ps < NPASS; ++ps)
#pragma unroll
        for (int r = 0; r < NREG; ++r) {
            v0[ps][r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has0 && off[ps] >= 0 && rr + ps * SW_RP < p.SR) v0[ps][r] = __ldg(reinterpret_cast<const float4 *>(x0 + off[ps] + r * 32));
        }
    if (has1) cp_async_wait_all();
#pragma unroll
    for (int ps = 0; ps < NPASS; ++ps)
        if (rr + ps * SW_RP < p.SR) {
#pragma unroll
            for (int r = 0; r < NREG; ++r) {
                uint4 *ptr = reinterpret_cast<uint4 *>(a_tile + sw_a_off(p, rr + ps * SW_RP, r, c16));
                float4 v;
                v.x = sd.wy0 * v0[ps][r].x; v.y = sd.wy0 * v0[ps][r].y; v.z = sd.wy0 * v0[ps][r].z; v.w = sd.wy0 * v0[ps][r].w;
                if (has1) {
                    const uint4 u = *ptr;
                    v.x = fmaf(sd.wy1, __uint_as_float(u.x), v.x); v.y = fmaf(sd.wy1, __uint_as_float(u.y), v.y);
                    v.z = fmaf(sd.wy1, __uint_as_float(u.z), v.z); v.w = fmaf(sd.wy1, __uint_as_float(u.w), v.w);
                }
                *ptr = sw_tf32x4(v);
            }
        }
}

__global__ void __launch_bounds__(SW_THREADS, 1) strip_wgrad_kernel(const SwParams p, const __grid_constant__ CUtensorMap tmap_dy)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.stages * p.stage_bytes);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages, tma0 = empty0 + 8 * p.stages;
    const uint32_t acc_full = tma0 + 8 * p.stages, acc_empty = acc_full + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * p.stages + 2);
    int *s_grow = reinterpret_cast<int *>(tmem_slot + 4);                 // start rows of the current unit's groups (MMA warp)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nseq = p.nwaves * p.TP;                                     // this CTA's (wave, tile) slots; item_pos() says which exist

    if (tid == 0) sw_stamp(p.trace, 102);
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, SW_GROUP_WARPS);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(tma0 + 8 * s, 1);
        }
        if (p.dy_tma) prefetch_tmap(&tmap_dy);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_mbar_init();
    }
    if (warp == SW_WARP_MMA) { tmem_alloc(smem_u32(tmem_slot), p.tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp > 4 && warp < 5 + SW_GROUPS * SW_GROUP_WARPS) {
        // ================================================ PRODUCERS ================================================
        // two groups; group g fills the stages of the items ti = g, g + 2, ...: while one group waits for its loads the other blends
        const int group = (warp - 5) / SW_GROUP_WARPS;
        const int ptid = ((warp - 5) - group * SW_GROUP_WARPS) * 32 + lane;           // 0 .. SW_PROD_THREADS-1 within the group
        const int c16 = ptid & 7, rr = ptid >> 3;                                    // 16-byte chunk of a 128-byte row; first row
        const bool dy_vec = (p.F % 4) == 0;
        int cur_unit = -1, ti = -1;
        StripDesc sd{};
        RowPlan row{};
        for (int ls = 0; ls < nseq; ++ls) {
            ItemPos q;
            if (!item_pos(p, ls, q)) continue;
            ++ti;
            if ((ti & 1) != group) continue;
            if (q.unit != cur_unit) {
                const WgUnit u = p.units[q.unit];
                sd = p.strips[u.strip];
                row = p.rows[u.row];
                cur_unit = q.unit;
            }
            const uint32_t s = (uint32_t)ti % (uint32_t)p.stages, phase = (((uint32_t)ti / (uint32_t)p.stages) & 1u) ^ 1u;
            mbar_wait_sleep(empty0 + 8 * s, phase);
            if (ptid == 0 && ti < 12) sw_stamp(p.trace, 48 + 4 * ti);
            uint8_t *a_tile = smem + s * p.stage_bytes;
            uint8_t *b_tile = a_tile + p.a_bytes;
            if (p.dy_tma && ptid == 0) {
                // the dy tile travels by TMA while the strip is produced: one box of (32 filters, 8 panoramas, 8 columns) per 32 filters
                const int nfreg = p.NF >> 5;
                mbar_arrive_expect_tx(tma0 + 8 * s, (uint32_t)(nfreg * p.KT * 128));
                for (int r = 0; r < nfreg; ++r)
                    tma_load_4d(smem_u32(b_tile + r * (p.KT * 128)), &tmap_dy, q.fc * p.NF + r * 32, q.b0, row.oc0 + q.j0, row.out_row, tma0 + 8 * s);
            }
            const int ch0 = q.cc * 128 + c16 * 4;
            const int nreg = min(p.nreg, (p.C - q.cc * 128) >> 5);
            // ---- A: the strip, [strip row][channel] ----
            if (sd.kind == 0) {
                switch (nreg) {
                case 4: for (int r0 = rr; r0 < p.SR; r0 += 2 * SW_RP) sw_fill_strip<4, 2>(p, sd, a_tile, c16, r0, q.j0, q.b0, ch0); break;
                case 3: for (int r0 = rr; r0 < p.SR; r0 += 2 * SW_RP) sw_fill_strip<3, 2>(p, sd, a_tile, c16, r0, q.j0, q.b0, ch0); break;
                case 2: for (int r0 = rr; r0 < p.SR; r0 += 2 * SW_RP) sw_fill_strip<2, 2>(p, sd, a_tile, c16, r0, q.j0, q.b0, ch0); break;
                default: for (int r0 = rr; r0 < p.SR; r0 += 3 * SW_RP) sw_fill_strip<1, 3>(p, sd, a_tile, c16, r0, q.j0, q.b0, ch0); break;
                }
            } else {
                // exact tap: the reference's per-pixel geometry (da_sample) for every pixel of the tile; rows past the tile are zero
                const int ta = sd.r0 / p.k, tb = sd.r0 % p.k;
                for (int rho = rr; rho < p.SR; rho += SW_RP) {
                    const int j = q.j0 + (rho >> 3), bimg = q.b0 + (rho & (SW_NB - 1));
                    const bool ok = rho < p.KT && j < p.OW && bimg < p.B;
                    CornerRef cr;
                    if (ok) {
                        const Sample sm = da_sample(row.out_row, j, ta, tb, sd.wy0, sd.wy1, p.in_h, p.in_w);
                        cr = da_corners(sm, bimg, p.H, p.W, p.C, p.ph0, p.pw0);
                    }
                    for (int r = 0; r < nreg; ++r) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (cr.off[u] < 0) continue;
                                const float4 pv = __ldg(reinterpret_cast<const float4 *>(p.x + cr.off[u] + ch0 + r * 32));
                                v.x = fmaf(cr.w[u], pv.x, v.x); v.y = fmaf(cr.w[u], pv.y, v.y);
                                v.z = fmaf(cr.w[u], pv.z, v.z); v.w = fmaf(cr.w[u], pv.w, v.w);
                            }
                        }
                        *reinterpret_cast<uint4 *>(a_tile + sw_a_off(p, rho, r, c16)) = sw_tf32x4(v);
                    }
                }
            }
            if (ptid == 0 && ti < 12) sw_stamp(p.trace, 49 + 4 * ti);
            // ---- B: the dy tile, [pixel][filter] ----
            if (p.dy_tma) {
                mbar_wait(tma0 + 8 * s, phase ^ 1);
                const int n16 = (p.NF >> 5) * p.KT * 8;                             // 16-byte chunks: round to TF32 (rna) in place
                for (int e = ptid; e < n16; e += SW_PROD_THREADS) {
                    uint4 *ptr = reinterpret_cast<uint4 *>(b_tile) + e;
                    const uint4 v = *ptr;
                    *ptr = sw_tf32x4(make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)));
                }
            } else {
                const int nfreg = p.NF >> 5;
                const int f_first = q.fc * p.NF + c16 * 4;
                for (int px = rr; px < p.KT; px += SW_RP) {
                    const int bimg = q.b0 + (px & (SW_NB - 1)), colo = q.j0 + (px >> 3);
                    const bool ok = colo < p.ncols && bimg < p.B;
                    const float *src = p.dy + ((size_t)(bimg * p.OH + row.out_row) * p.OW + (ok ? row.oc0 + p.ocs * colo : 0)) * p.F + f_first;
                    const uint32_t so = sw_offset((uint32_t)px, (uint32_t)c16);
                    for (int r = 0; r < nfreg; ++r) {
                        const int f = f_first + r * 32;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok && f < p.F) {
                            if (dy_vec) v = __ldg(reinterpret_cast<const float4 *>(src + r * 32));
                            else {
                                v.x = __ldg(src + r * 32);
                                if (f + 1 < p.F) v.y = __ldg(src + r * 32 + 1);
                                if (f + 2 < p.F) v.z = __ldg(src + r * 32 + 2);
                                if (f + 3 < p.F) v.w = __ldg(src + r * 32 + 3);
                            }
                        }
                        *reinterpret_cast<uint4 *>(b_tile + r * (p.KT * 128) + so) = sw_tf32x4(v);
                    }
                }
            }
            if (ptid == 0 && ti < 12) sw_stamp(p.trace, 50 + 4 * ti);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
    } else if (warp < 4) {
        // ================================================ DRAIN ================================================
        // TMEM lane = channel (of window win[0]), or quarter `warp` = window win[warp] and lane = channel (32-channel layers)
        const bool vec = (p.F % 4) == 0;
        int n = 0;
        for (int w = 0; w < p.nwaves; ++w) {
            ItemPos q;
            if (!item_pos(p, w * p.TP, q)) continue;                      // no tile of this wave belongs to the CTA
            const WgUnit u = p.units[q.unit];
            mbar_wait_sleep(acc_full, n & 1);
            tc_fence_after();
            if (tid == 0 && n == 0) sw_stamp(p.trace, 100);
            const int cpw = 128 / p.wpg, wq = (warp * 32) / cpw;          // accumulator lanes per window; this warp's window of a group
            const int c = q.cc * 128 + (warp * 32) % cpw + lane;
            for (int g = u.group_begin; g < u.group_end; ++g) {
                const int win = __ldg(&p.groups[g].win[wq]);
                int t_lo = 0, t_hi = 1, tap1 = 0;
                if (win >= 0) {
                    if (p.weff) { t_lo = __ldg(p.term_begin + win); t_hi = __ldg(p.term_begin + win + 1); }
                    else tap1 = __ldg(&p.wins[win].wtile0);
                }
                for (int c0 = 0; c0 < p.NF; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + (uint32_t)((g - u.group_begin) * p.NF + c0) + ((uint32_t)(warp * 32) << 16), r);
                    tmem_ld_wait();
                    const int f0 = q.fc * p.NF + c0;
                    if (win < 0 || c >= p.C || f0 >= p.F) continue;
                    for (int ti = t_lo; ti < t_hi; ++ti) {
                        int tap = tap1;
                        float coef = 1.f;
                        if (p.weff) { const WeffTerm tm = p.terms[ti]; tap = tm.tap; coef = tm.coef; }
                        float *dst = p.dw + ((size_t)tap * p.C + c) * p.F + f0;
                        if (vec) {
#pragma unroll
                            for (int e = 0; e < 32; e += 4)
                                if (f0 + e < p.F)
                                    sw_red_add_v4(dst + e, coef * __uint_as_float(r[e]), coef * __uint_as_float(r[e + 1]),
                                                  coef * __uint_as_float(r[e + 2]), coef * __uint_as_float(r[e + 3]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (f0 + e < p.F) atomicAdd(dst + e, coef * __uint_as_float(r[e]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (tid == 0 && n == 0) sw_stamp(p.trace, 101);
            if (lane == 0) mbar_arrive(acc_empty);
            ++n;
        }
    } else if (warp == SW_WARP_MMA) {
        // ================================================ MMA ISSUER ================================================
        // the whole warp walks the loops and the barrier waits; one elected lane issues (descriptor words stay in uniform registers)
        const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)p.NF) | (1u << 15) | (1u << 16);       // both operands MN-major
        const uint32_t a_hi = (512u >> 4) | (1u << 14) | (1u << 29);                                   // sbo, descriptor version, SWIZZLE_128B_BASE32B
        const uint32_t a_lbo = (((uint32_t)p.a_lbo >> 4) & 0x3FFFu) << 16, b_lbo = (((uint32_t)(p.KT * 128) >> 4) & 0x3FFFu) << 16;
        const uint32_t a_kstep = (uint32_t)p.a_kstep, a_rowu = 8u * (uint32_t)(p.nch > 1 ? p.nch : 1);
        uint32_t s = 0, phase = 0;
        int cur_uidx = -1, ng = 0, n = 0, ti = -1;
        for (int ls = 0; ls < nseq; ++ls) {
            ItemPos q;
            if (!item_pos(p, ls, q)) continue;
            ++ti;
            const bool first_tile = q.uidx != cur_uidx;
            if (first_tile) {
                __syncwarp();
                const WgUnit u = p.units[q.unit];
                ng = u.group_end - u.group_begin;
                if (lane < ng) s_grow[lane] = __ldg(&p.groups[u.group_begin + lane].start_row);
                __syncwarp();
                if (n > 0) { mbar_wait(acc_empty, (n - 1) & 1); tc_fence_after(); }      // the previous unit's accumulators are drained
                cur_uidx = q.uidx;
                ++n;
            }
            const bool last_tile = (ls + 1) % p.TP == 0 || q.tile + 1 >= p.ntiles;
            if (lane == 0 && ti < 12) sw_stamp(p.trace, 4 * ti);
            mbar_wait(full0 + 8 * s, phase);
            tc_fence_after();
            if (lane == 0 && ti < 12) sw_stamp(p.trace, 4 * ti + 1);
            const uint32_t a0 = (smem_u32(smem + s * p.stage_bytes) & 0x3FFFFu) >> 4, b0 = a0 + ((uint32_t)p.a_bytes >> 4);
            if (elect_one()) {                   // one lane issues everything up to the commits (tcgen05.commit tracks the issuing thread's MMAs)
                for (int g = 0; g < ng; ++g) {
                    const uint32_t a_g = a0 + (uint32_t)s_grow[g] * a_rowu;                // start row (x chunks per column) * 128 B, in 16-byte units
                    const uint32_t d_g = tmem_base + (uint32_t)(g * p.NF);
#pragma unroll 8
                    for (int k8 = 0; k8 < p.TW; ++k8) {
                        const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)((a_g + k8 * a_kstep) | a_lbo);
                        const uint64_t db = ((uint64_t)a_hi << 32) | (uint64_t)((b0 + k8 * 64u) | b_lbo);
                        umma_tf32(d_g, da, db, idesc, (first_tile && k8 == 0) ? 0u : 1u);
                    }
                }
                umma_commit(empty0 + 8 * s);
                if (last_tile) umma_commit(acc_full);
            }
            __syncwarp();
            if (lane == 0 && ti < 12) sw_stamp(p.trace, 4 * ti + 2);
            if (lane == 0 && (ti & 31) == 0 && (ti >> 5) < 24) sw_stamp(p.trace, 104 + (ti >> 5));      // every 32nd item: the steady-state rate
            if (++s == (uint32_t)p.stages) { s = 0; phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) sw_stamp(p.trace, 103);
    if (warp == SW_WARP_MMA) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// dbias[f] = sum over pixels of dy[m, f] for a dense [M][F] matrix: the flat array is read with 256 consecutive threads per step, a
// thread's column is fixed when the step (a multiple of F) keeps e % F — so every thread is busy whatever F is (3 ... 512) and the loads are
// coalesced; the threads of a block that share a column meet in shared memory, one atomic per (block, column).
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS) sw_col_sum_kernel(const float *__restrict__ dy, float *__restrict__ db, long total, int F, int step)
{
    __shared__ float part[CS_THREADS];
    float s = 0.f;
    const int lane_e = threadIdx.x;                       // element index within a step; active while < step
    if (lane_e < step)
        for (long e = (long)blockIdx.x * step + lane_e; e < total; e += (long)gridDim.x * step) s += __ldg(dy + e);
    part[threadIdx.x] = s;
    __syncthreads();
    // threads t, t + F, t + 2F, ... (< step) hold the same column
    if (threadIdx.x < F && threadIdx.x < step) {
        float acc = 0.f;
        for (int t = threadIdx.x; t < step; t += F) acc += part[t];
        atomicAdd(db + threadIdx.x, acc);
    }
}
// F > 256: one thread per column, rows split over blockIdx.y
__global__ void sw_col_sum_wide_kernel(const float *__restrict__ dy, float *__restrict__ db, int M, int F)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * rows_per, hi = min(M, lo + rows_per);
    float s = 0.f;
    for (int m = lo; m < hi; ++m) s += dy[(size_t)m * F + f];
    if (hi > lo) atomicAdd(db + f, s);
}

int num_sms()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace

// Schedule of the strip weight gradient: every accumulation (unit x channel chunk x filter chunk; nuidx of them) is split over P CTAs of
// TP pixel tiles each (P * TP >= ntiles), U = SMs / P accumulations run at a time (a wave), and a CTA drains its accumulators once per
// wave.  P is chosen by a cost model in microseconds (about 1.5 per tile, 3.5 per drain); `big` (the tensors do not fit in L2): U is capped
// at ucap so that the rows a wave touches do.
void wgrad_schedule(int nuidx, int ntiles, int sms, bool big, int ucap, int *P_out, int *TP_out, int *U_out, int *nwaves_out)
{
    double best = 1e30;
    int bP = 1, bTP = ntiles, bU = sms < nuidx ? sms : nuidx;
    for (int P = 1; P <= ntiles && P <= sms; ++P) {
        const int TP = (ntiles + P - 1) / P, Pe = (ntiles + TP - 1) / TP;
        int U = sms / Pe;
        if (U > nuidx) U = nuidx;
        if (U < 1) continue;
        if (big && U > ucap && Pe < ntiles) continue;
        const int waves = (nuidx + U - 1) / U;
        const double cost = waves * (TP * 1.5 + 3.5);
        if (cost < best) { best = cost; bP = Pe; bTP = TP; bU = U; }
    }
    if (bU < 1) bU = 1;
    *P_out = bP; *TP_out = bTP; *U_out = bU; *nwaves_out = (nuidx + bU - 1) / bU;
}

// db[f] (+)= column sums of the dense matrix dy [M][F] (db zeroed by the caller)
int launch_col_sum(const float *dy, float *db, int M, int F, cudaStream_t st)
{
    if (F <= CS_THREADS) {
        const int step = (CS_THREADS / F) * F;             // elements per block step: whole rows
        const long total = (long)M * F;
        long blocks = (total + step - 1) / step;
        if (blocks > 4 * 148) blocks = 4 * 148;
        sw_col_sum_kernel<<<(int)blocks, CS_THREADS, 0, st>>>(dy, db, total, F, step);
    } else {
        int ysplit = (M + 31) / 32;
        if (ysplit > 4 * 148) ysplit = 4 * 148;
        sw_col_sum_wide_kernel<<<dim3((F + 127) / 128, ysplit), 128, 0, st>>>(dy, db, M, F);
    }
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

int launch_wgrad_strip(const float *x, const float *dy, const float *offsets_host, float *dw, int B, int h, int w, int C, int F, int k,
                       int stride, cudaStream_t stream)
{
    if (C % 32 != 0) return SKY_ERR_UNSUPPORTED;
    if ((long)B * h * w * (long)(C > F ? C : F) >= (1L << 31)) return SKY_ERR_UNSUPPORTED;
    SwParams p{};
    p.x = x; p.dy = dy; p.dw = dw;
    p.B = B; p.H = h; p.W = w; p.C = C; p.F = F; p.k = k;
    const int Fp = round_up(F, 32);
    p.NF = Fp < 128 ? Fp : 128;
    p.nfc = (Fp + p.NF - 1) / p.NF;
    p.wpg = C == 32 ? 4 : (C == 64 ? 2 : 1);
    p.nch = C == 32 ? 1 : (C == 64 ? 2 : 0);
    p.ncc = (C + 127) / 128;
    const int gmax = (512 / p.NF) < SW_MAX_GROUPS ? (512 / p.NF) : SW_MAX_GROUPS;
    const StripPlan *pl = nullptr;
    int rc;
    if (offsets_host) {
        if (stride != 1 || !(k & 1)) return SKY_ERR_UNSUPPORTED;
        rc = get_plan_da_wgrad(offsets_host, h, w, k, p.wpg, gmax, &pl);
        int pht, pwt;
        pad_axis(h, k, &p.ph0, &pht);
        pad_axis(w, k, &p.pw0, &pwt);
        p.in_h = h + pht; p.in_w = w + pwt; p.da = 1; p.OH = h; p.OW = w;
    } else {
        p.OH = (h + stride - 1) / stride; p.OW = (w + stride - 1) / stride;
        rc = get_plan_plain_wgrad(h, w, k, stride, p.OH, p.OW, p.wpg, gmax, &pl);
        p.da = 0;
    }
    if (rc != SKY_OK) return rc;
    if (pl->n_wg_units == 0) return SKY_OK;
    p.rows = pl->rows; p.strips = pl->strips; p.wins = pl->wins; p.term_begin = pl->term_begin; p.terms = pl->terms;
    p.units = pl->wg_units; p.groups = pl->wg_groups;
    p.weff = pl->weff; p.ncols = pl->ncols; p.ocs = pl->ocs;
    // tile width: 8 columns where the strip is 128 channels wide; 32-channel layers take 32 columns per tile (a stage costs a round trip
    // to L2 whatever it holds, and their tiles are small)
    p.nreg = C >= 128 ? 4 : C / 32;
    const int budget = 227 * 1024 - 2048;
    // (12 columns where the rows are long enough for the ragged last tile not to matter: a third less halo per strip, measured +5 %)
    for (p.TW = getenv("SKY_WGRAD_TW") ? atoi(getenv("SKY_WGRAD_TW")) : (p.wpg == 4 ? 32 : (p.wpg == 2 ? 16 : (pl->ncols >= 96 ? 12 : 8)));; p.TW -= (p.TW == 12 ? 4 : 8)) {
        if (p.TW > round_up(pl->ncols, 8)) p.TW = round_up(pl->ncols, 8);
        if (p.TW < 8) p.TW = 8;
        p.KT = p.TW * SW_NB;
        p.SR = (p.TW + pl->span_max + p.wpg - 1) * SW_NB;
        p.a_lbo = p.nch ? SW_NB * 128 : p.SR * 128;
        p.a_kstep = 64 * (p.nch > 1 ? p.nch : 1);
        // the A tile is sized for the four MN atoms an M = 128 MMA reads (64-channel layers leave two of them unwritten: their lanes are not stored)
        p.a_bytes = p.nch ? p.nch * p.SR * 128 : 4 * p.SR * 128;
        p.stage_bytes = p.a_bytes + (p.NF / 32) * p.KT * 128;
        p.stages = budget / p.stage_bytes;
        if (p.stages >= 2 || p.TW <= 8) break;
    }
    if (p.stages > 4) p.stages = 4;
    if (p.stages < 2) return SKY_ERR_UNSUPPORTED;
    p.tiles_x = (pl->ncols + p.TW - 1) / p.TW; p.tiles_b = (B + SW_NB - 1) / SW_NB; p.ntiles = p.tiles_x * p.tiles_b;
    p.nuidx = pl->n_wg_units * p.ncc * p.nfc;
    if ((long)p.nuidx * p.ntiles >= (1L << 30)) return SKY_ERR_UNSUPPORTED;
    {
        const bool big = ((double)B * h * w * C + (double)B * p.OH * p.OW * F) * 4.0 > 80e6;
        const int ucap = getenv("SKY_WGRAD_UCAP") ? atoi(getenv("SKY_WGRAD_UCAP")) : 16;
        wgrad_schedule(p.nuidx, p.ntiles, num_sms(), big, ucap, &p.P, &p.TP, &p.U, &p.nwaves);
    }
    const int grid = p.U * p.P;
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < pl->max_groups_unit * p.NF) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return SKY_ERR_UNSUPPORTED;
    // one CTA per SM (the TMEM allocation of a second one could block): always ask for more than half of the shared memory
    int smem = p.stages * p.stage_bytes + (3 * p.stages + 2) * 8 + 16 + SW_MAX_GROUPS * 4 + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;
    p.trace = getenv("SKY_WGRAD_TRACE") != nullptr;
    CUtensorMap tmap_dy;
    memset(&tmap_dy, 0, sizeof(tmap_dy));
    p.dy_tma = (F % 4 == 0) && pl->ocs == 1 && !getenv("SKY_WGRAD_NO_TMA");
    if (p.dy_tma) {
        rc = encode_dy_wgrad_tensor_map(&tmap_dy, dy, B, p.OH, p.OW, F, p.TW);
        if (rc != SKY_OK) return rc;
    }
    SKY_ENSURE_DYN_SMEM(strip_wgrad_kernel, 227 * 1024);
    strip_wgrad_kernel<<<grid, SW_THREADS, smem, stream>>>(p, tmap_dy);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

// Weight gradient of the distortion-aware layer on the strip formulation (same contract as sky_da_conv2d_bwd_filter, with the HOST copy
// of the offset table the strip plans are built from).
// accumulate != 0: dkernel / dbias are added to (the caller zeroed them, e.g. one memset of a flat gradient buffer per step).
extern "C" int sky_da_conv2d_bwd_filter_strip(const float *x, const float *dy, const float *offsets_host, float *dkernel, float *dbias, int B,
                                              int h, int w, int C, int F, int k, int accumulate, void *stream)
{
    SKY_REQUIRE(x && dy && offsets_host && dkernel, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    SKY_REQUIRE(C % 32 == 0, SKY_ERR_UNSUPPORTED, "the strip weight gradient needs C %% 32 == 0 (got %d)", C);
    SKY_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dkernel & 15) == 0, SKY_ERR_INVALID, "x, dy and dkernel must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dkernel, 0, (size_t)k * k * C * F * sizeof(float), st));
    const int rc = launch_wgrad_strip(x, dy, offsets_host, dkernel, B, h, w, C, F, k, 1, st);
    SKY_REQUIRE(rc == SKY_OK, rc, "no strip weight-gradient plan for B=%d h=%d w=%d C=%d F=%d k=%d", B, h, w, C, F, k);
    if (dbias) {
        const int M = B * h * w;
        if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)F * sizeof(float), st));
        return launch_col_sum(dy, dbias, M, F, st);
    }
    return SKY_OK;
}

/* debug: the 128 %globaltimer stamps CTA 0 of the last strip weight-gradient launch left when SKY_WGRAD_TRACE was set */
extern "C" int sky_debug_wgrad_trace(unsigned long long *host_out128)
{
    SKY_REQUIRE(host_out128, SKY_ERR_INVALID, "NULL pointer");
    SKY_CHECK_CUDA(cudaMemcpyFromSymbol(host_out128, g_wgrad_trace, sizeof(unsigned long long) * 128));
    return SKY_OK;
}

/* debug / tests (host only): the schedule launch_wgrad_strip picks — out4 = P, TP, U, waves */
extern "C" int sky_wgrad_schedule_info(int nuidx, int ntiles, int sms, int big, int ucap, int *out4)
{
    SKY_REQUIRE(out4 && nuidx > 0 && ntiles > 0 && sms > 0 && ucap > 0, SKY_ERR_INVALID, "bad arguments");
    wgrad_schedule(nuidx, ntiles, sms, big != 0, ucap, out4, out4 + 1, out4 + 2, out4 + 3);
    return SKY_OK;
}
