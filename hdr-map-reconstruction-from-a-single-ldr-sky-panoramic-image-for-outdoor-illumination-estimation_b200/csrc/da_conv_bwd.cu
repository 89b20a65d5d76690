// Backward of the distortion-aware convolution (what TF autodiff derives from distortion_aware_ops.py:50-123:
// matmul -> two matmul transposes, gather_nd -> scatter-add, pad -> slice; the offsets are constants).
//
//   dgrad  dX[corner(m,t), c] += w_corner(m,t) * dPix[m,(t,c)],   dPix[m,(t,c)] = sum_f dY[m,f] * W[(t,c),f]
//          per 128-pixel tile and k-block: tcgen05.mma (A = dY tile, K-major; B = 32 rows of the kernel variable, K-major in
//          its natural [k*k*C, F] layout) into TMEM, then the 128 row-owner threads scatter their 32 channels to the four
//          corners with red.global.add.v4.f32.  dPix (9x the size of dY) is never materialised.
//   wgrad  dW[(t,c), f] = sum_m Pix[m,(t,c)] * dY[m,f]
//          the contraction index is the pixel, which is the ROW index of both natural layouts, so both operands are fed
//          MN-major: A = four re-gathered Pix tiles [128 pixels x 32 channels] (one per tap of a tap group, M' = 128),
//          B = the dY tile [128 pixels x F].  Accumulator stays in TMEM across the CTA's pixel tiles; partial sums of the
//          pixel partitions meet in global memory with vector atomics.  Pix is re-gathered, never stored.
//   dbias  column sums of dY.
//
// Round-1 versions are phase-structured (load -> MMA -> drain per step, no warp specialisation) and gather straight from
// global/L2 with the exact per-pixel geometry; they are correct-first, the pipelined versions follow the forward kernel.
#include "da_conv.cuh"

namespace sky {

constexpr int BWD_THREADS = 256;

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// instruction descriptor with explicit operand majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t idesc_tf32_major(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn)
{
    return umma_idesc_tf32(M, N) | (a_mn << 15) | (b_mn << 16);
}
// MN-major TF32 operands have exactly one legal shared-memory layout: SWIZZLE_128B with a 32-byte base
// (layout type 1): atoms of [4 k-rows x 128 B of MN], 32-byte units XOR-ed with (k-row & 3).
// lbo = bytes between MN atoms (32 elements each), sbo = bytes between k atoms (4 rows each).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128_b32(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
    return d;
}
// byte offset of (k-row, 16-byte chunk) in a tile of 128-byte rows under that swizzle
__device__ __forceinline__ uint32_t sw128_b32_offset(uint32_t row, uint32_t chunk16)
{
    return row * 128u + ((((chunk16 >> 1) ^ (row & 3u)) << 5) | ((chunk16 & 1u) << 4));
}

struct BwdParams {
    const float *x, *offsets, *kernel, *dy;
    float *dx, *dw;
    int B, h, w, C, F, k, k2, CC, FC;   // CC = C/32, FC = F/32
    int in_h, in_w, ph0, pw0, M;
    int parts, tiles_per_part;           // wgrad: pixel partitions
    int ksplit;                          // dgrad: the k-blocks of a tile are shared out over gridDim.y CTAs (dx is accumulated
                                         // with atomics anyway), which fills the SMs when there are few pixel tiles
};

// ---------------------------------------------------------------------------------------------------------------------
// dgrad
// ---------------------------------------------------------------------------------------------------------------------
// smem: A = dY tile, FC atoms of [128 x 32] K-major (16 KB each) | B = 2 x (FC atoms of [32 x 32], 4 KB each) | barriers
__global__ void __launch_bounds__(BWD_THREADS) da_conv2d_dgrad_kernel(const BwdParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *a_tile = smem;
    uint8_t *b_tile = smem + p.FC * 16384;
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_tile + 2 * p.FC * 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const uint32_t bar0 = smem_u32(bars);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * BLOCK_M;

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 64); tmem_relinquish(); }

    // dY tile -> A (K-major in f), TF32-rounded; rows past M are zero
    for (int e = tid; e < BLOCK_M * p.F / 4; e += BWD_THREADS) {
        const int row = e / (p.F / 4), c4 = e % (p.F / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + row < p.M) v = __ldg(reinterpret_cast<const float4 *>(p.dy + (size_t)(m0 + row) * p.F) + c4);
        uint4 u;
        u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
        *reinterpret_cast<uint4 *>(a_tile + (c4 / 8) * 16384 + sw128_offset(row, c4 % 8)) = u;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(BLOCK_M, 32);

    // this thread's pixel (threads 0..127 own TMEM lane == tile row)
    const int m = m0 + (tid & 127);
    const bool m_ok = (tid < 128) && (m < p.M);
    const int j = m % p.w, i = (m / p.w) % p.h, b_img = m / (p.w * p.h);

    const int KB_all = p.k2 * p.CC;
    const int kb_per = (KB_all + p.ksplit - 1) / p.ksplit;
    const int kb_lo = blockIdx.y * kb_per;
    const int KB = min(kb_per, KB_all - kb_lo);          // this CTA's k-blocks: kb_lo .. kb_lo + KB - 1 (local index kb below)
    if (KB <= 0) {                                       // (uniform per CTA) nothing to do: release TMEM and leave
        __syncthreads();
        if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
        return;
    }
    // Scatter of one drained k-block.  The kernel is bound by the L2 reduction traffic (ncu: 73 % of the L2 peak on a 7x7 layer), and
    // half of it is redundant: along a row the right-hand corner of pixel j is the left-hand corner of pixel j + 1 (the offsets
    // depend on the row and the tap only), so lane l adds its left neighbour's right-corner contribution — passed down by warp
    // shuffles — to its own left-corner one and issues ONE vector reduction per 16 bytes and corner row instead of two.  Lanes whose
    // neighbour does not line up (360-degree wrap, zenith row, row ends, image edge) keep their own reduction.
    auto drain = [&](int pk) {
        const int pt = (kb_lo + pk) / p.CC, pcc = (kb_lo + pk) % p.CC;
        mbar_wait(bar0 + 8 * (pk & 1), (pk >> 1) & 1);
        tc_fence_after();
        if (warp < 4) {
            const int lane = tid & 31;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (pk & 1) * 32 + ((uint32_t)(warp * 32) << 16), r);
            tmem_ld_wait();
            CornerRef cr;
#pragma unroll
            for (int c = 0; c < 4; ++c) { cr.off[c] = -1; cr.w[c] = 0.f; }
            if (m_ok) {
                const float2 yx = __ldg(reinterpret_cast<const float2 *>(p.offsets) + (size_t)i * p.k2 + pt);
                const Sample s = da_sample(i, j, pt / p.k, pt % p.k, yx.x, yx.y, p.in_h, p.in_w);
                cr = da_corners(s, b_img, p.h, p.w, p.C, p.ph0, p.pw0);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (cr.w[c] == 0.f) cr.off[c] = -1;
            }
#pragma unroll
            for (int rowc = 0; rowc < 2; ++rowc) {          // corner rows: (y0: corners 0, 1), (y1: corners 2, 3)
                const int offL = cr.off[2 * rowc], offR = cr.off[2 * rowc + 1];
                const float wL = cr.w[2 * rowc], wR = cr.w[2 * rowc + 1];
                const int nbR = __shfl_up_sync(0xffffffffu, offR, 1);                   // left neighbour's right corner
                const bool take = lane > 0 && offL >= 0 && nbR == offL;                 // I carry my left neighbour's right corner
                const bool given = __shfl_down_sync(0xffffffffu, (int)take, 1) != 0 && lane < 31;   // my right corner is carried
                float *dstL = p.dx + (offL >= 0 ? offL : 0) + pcc * 32, *dstR = p.dx + (offR >= 0 ? offR : 0) + pcc * 32;
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    float vr[4], vl[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        vr[u] = wR * __uint_as_float(r[q + u]);
                        const float nb = __shfl_up_sync(0xffffffffu, vr[u], 1);
                        vl[u] = wL * __uint_as_float(r[q + u]) + (take ? nb : 0.f);
                    }
                    if (offL >= 0) red_add_v4(dstL + q, vl[0], vl[1], vl[2], vl[3]);
                    if (offR >= 0 && !given) red_add_v4(dstR + q, vr[0], vr[1], vr[2], vr[3]);
                }
            }
        }
        tc_fence_before();
    };
    for (int kb = 0; kb < KB; ++kb) {
        const int t = (kb_lo + kb) / p.CC, cc = (kb_lo + kb) % p.CC;
        uint8_t *bt = b_tile + (kb & 1) * p.FC * 4096;
        // B = kernel rows (t*C + cc*32 + n), n < 32: [32 x F] K-major in f
        for (int e = tid; e < 32 * p.F / 4; e += BWD_THREADS) {
            const int n = e / (p.F / 4), c4 = e % (p.F / 4);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p.kernel + (size_t)(t * p.C + cc * 32 + n) * p.F) + c4);
            uint4 u;
            u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
            *reinterpret_cast<uint4 *>(bt + (c4 / 8) * 4096 + sw128_offset(n, c4 % 8)) = u;
        }
        fence_proxy_async_smem();
        __syncthreads();
        const uint32_t d_tmem = tmem_base + (kb & 1) * 32;
        if (tid == 0) {
            tc_fence_after();
            for (int a = 0; a < p.FC; ++a)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_tf32(d_tmem, umma_desc_kmajor_sw128(smem_u32(a_tile + a * 16384) + ks * 32),
                              umma_desc_kmajor_sw128(smem_u32(bt + a * 4096) + ks * 32), idesc, (a | ks) != 0);
            umma_commit(bar0 + 8 * (kb & 1));
        }
        // drain the PREVIOUS k-block while this one's MMAs run
        if (kb > 0) drain(kb - 1);
        __syncthreads();
    }
    drain(KB - 1);   // the last k-block
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------------------------------
// grid: (parts, tap groups of 4, CC).  smem: A = 4 Pix tiles [128 px x 32 ch] (16 KB each, one per tap of the group) |
// B = dY tile as FC column groups of [128 px x 32 f] (16 KB each) | barrier.
__global__ void __launch_bounds__(BWD_THREADS) da_conv2d_wgrad_kernel(const BwdParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *a_tile = smem;                     // 4 x 16 KB
    uint8_t *b_tile = smem + 4 * 16384;         // FC x 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_tile + p.FC * 16384);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 1);
    CornerRef *ctab = reinterpret_cast<CornerRef *>(b_tile + p.FC * 16384 + 64);      // [4 taps][128 pixels]
    const uint32_t bar0 = smem_u32(bars);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int part = blockIdx.x, tg = blockIdx.y, cc = blockIdx.z;
    const int t_first = tg * 4;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.F) tmem_cols <<= 1;

    if (tid == 0) { mbar_init(bar0, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), tmem_cols); tmem_relinquish(); }
    // taps beyond k*k (last group) contribute zero rows
    for (int e = tid; e < 4 * 16384 / 16; e += BWD_THREADS) reinterpret_cast<uint4 *>(a_tile)[e] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = idesc_tf32_major(BLOCK_M, (uint32_t)p.F, 1, 1);
    const int ntiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int tile_lo = part * p.tiles_per_part, tile_hi = min(ntiles, tile_lo + p.tiles_per_part);
    const int chunk = tid & 7, row_base = tid >> 3;      // (row, 16-byte chunk) items: rows row_base + 32 r

    uint32_t iter = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++iter) {
        const int m0 = tile * BLOCK_M;
        // ---- B: dY tile, MN-major (row = pixel = contraction index) ----
        for (int e = tid; e < BLOCK_M * p.F / 4; e += BWD_THREADS) {
            const int row = e / (p.F / 4), c4 = e % (p.F / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < p.M) v = __ldg(reinterpret_cast<const float4 *>(p.dy + (size_t)(m0 + row) * p.F) + c4);
            uint4 u;
            u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
            *reinterpret_cast<uint4 *>(b_tile + (c4 / 8) * 16384 + sw128_b32_offset(row, c4 % 8)) = u;
        }
        // ---- geometry of the tile: one (tap, pixel) sample per table entry, computed once and shared by the eight threads that
        //      gather the entry's eight 16-byte channel chunks (the sampling arithmetic used to be redone by each of them: 58 % of
        //      the kernel's issue slots on a 7x7 layer) ----
        for (int e = tid; e < 4 * BLOCK_M; e += BWD_THREADS) {
            const int tl = e / BLOCK_M, row = e % BLOCK_M, t = t_first + tl, m = m0 + row;
            CornerRef cr;
#pragma unroll
            for (int c = 0; c < 4; ++c) { cr.off[c] = -1; cr.w[c] = 0.f; }
            if (t < p.k2 && m < p.M) {
                const int j = m % p.w, i = (m / p.w) % p.h, b_img = m / (p.w * p.h);
                const float2 yx = __ldg(reinterpret_cast<const float2 *>(p.offsets) + (size_t)i * p.k2 + t);
                const Sample s = da_sample(i, j, t / p.k, t % p.k, yx.x, yx.y, p.in_h, p.in_w);
                cr = da_corners(s, b_img, p.h, p.w, p.C, p.ph0, p.pw0);
            }
            ctab[e] = cr;
        }
        __syncthreads();
        // ---- A: re-gather Pix for the group's taps, 32 channels of chunk cc (corners from global/L2) ----
        for (int tl = 0; tl < 4; ++tl) {
            const int t = t_first + tl;
            if (t >= p.k2) break;
            for (int r = 0; r < BLOCK_M / 32; ++r) {
                const int row = row_base + 32 * r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                {
                    const CornerRef cr = ctab[tl * BLOCK_M + row];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (cr.off[c] < 0) continue;
                        const float4 px = __ldg(reinterpret_cast<const float4 *>(p.x + cr.off[c] + cc * 32 + chunk * 4));
                        v.x = fmaf(cr.w[c], px.x, v.x); v.y = fmaf(cr.w[c], px.y, v.y);
                        v.z = fmaf(cr.w[c], px.z, v.z); v.w = fmaf(cr.w[c], px.w, v.w);
                    }
                }
                uint4 u;
                u.x = f32_to_tf32_rna(v.x); u.y = f32_to_tf32_rna(v.y); u.z = f32_to_tf32_rna(v.z); u.w = f32_to_tf32_rna(v.w);
                *reinterpret_cast<uint4 *>(a_tile + tl * 16384 + sw128_b32_offset(row, chunk)) = u;
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int g8 = 0; g8 < BLOCK_M / 8; ++g8) {      // 8 pixels (K = 8) per MMA
                const uint64_t da = umma_desc_mnmajor_sw128_b32(smem_u32(a_tile) + g8 * 1024, 16384, 512);
                const uint64_t db = umma_desc_mnmajor_sw128_b32(smem_u32(b_tile) + g8 * 1024, 16384, 512);
                umma_tf32(tmem_base, da, db, idesc, (iter | (uint32_t)g8) != 0);
            }
            umma_commit(bar0);
        }
        mbar_wait(bar0, iter & 1);       // the tiles in smem may be overwritten only after the MMAs have read them
        tc_fence_after();
    }
    // ---- drain: row (tl*32 + c) of D is dW[(t_first+tl)*C + cc*32 + c, :] ----
    if (warp < 4 && tile_hi > tile_lo) {
        const int rowd = warp * 32 + (tid & 31), tl = rowd >> 5, c = rowd & 31;
        const int t = t_first + tl;
        for (int c0 = 0; c0 < p.F; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (uint32_t)c0 + ((uint32_t)(warp * 32) << 16), r);
            tmem_ld_wait();
            if (t < p.k2) {
                float *dst = p.dw + (size_t)(t * p.C + cc * 32 + c) * p.F + c0;
#pragma unroll
                for (int q = 0; q < 32; q += 4)
                    if (c0 + q < p.F)
                        red_add_v4(dst + q, __uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]),
                                   __uint_as_float(r[q + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dbias[f] = sum_m dY[m, f]
__global__ void col_sum_kernel(const float *__restrict__ dy, float *__restrict__ db, int M, int F)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * rows_per, hi = min(M, lo + rows_per);
    float s = 0.f;
    for (int m = lo; m < hi; ++m) s += dy[(size_t)m * F + f];
    atomicAdd(db + f, s);
}

static int fill_bwd_params(BwdParams &p, const float *x, const float *offsets, const float *kernel, const float *dy, int B, int h,
                           int w, int C, int F, int k)
{
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    SKY_REQUIRE(C % 32 == 0 && F % 32 == 0 && F <= 256, SKY_ERR_UNSUPPORTED,
                "backward kernels need C %% 32 == 0 and F %% 32 == 0, F <= 256 (got C=%d F=%d)", C, F);
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    p.x = x; p.offsets = offsets; p.kernel = kernel; p.dy = dy; p.dx = nullptr; p.dw = nullptr;
    p.B = B; p.h = h; p.w = w; p.C = C; p.F = F; p.k = k; p.k2 = k * k; p.CC = C / 32; p.FC = F / 32;
    int pht, pwt;
    pad_axis(h, k, &p.ph0, &pht);
    pad_axis(w, k, &p.pw0, &pwt);
    p.in_h = h + pht; p.in_w = w + pwt;
    p.M = B * h * w;
    p.parts = 1; p.tiles_per_part = 0; p.ksplit = 1;
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

extern "C" int sky_da_conv2d_bwd_data(const float *dy, const float *offsets, const float *kernel, float *dx, int B, int h, int w,
                                      int C, int F, int k, int accumulate, void *stream)
{
    BwdParams p;
    int rc = fill_bwd_params(p, nullptr, offsets, kernel, dy, B, h, w, C, F, k);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(dy && offsets && kernel && dx, SKY_ERR_INVALID, "NULL pointer");
    p.dx = dx;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) SKY_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)p.M * C * sizeof(float), st));
    const int smem = p.FC * 16384 + 2 * p.FC * 4096 + 64 + 1024;
    SKY_ENSURE_DYN_SMEM(da_conv2d_dgrad_kernel, 227 * 1024);
    const int tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    int ksplit = (2 * 148 + tiles - 1) / tiles;          // about two waves of CTAs
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.k2 * p.CC) ksplit = p.k2 * p.CC;
    p.ksplit = ksplit;
    da_conv2d_dgrad_kernel<<<dim3(tiles, ksplit), BWD_THREADS, smem, st>>>(p);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_da_conv2d_bwd_filter(const float *x, const float *dy, const float *offsets, float *dkernel, float *dbias, int B,
                                        int h, int w, int C, int F, int k, void *stream)
{
    BwdParams p;
    int rc = fill_bwd_params(p, x, offsets, nullptr, dy, B, h, w, C, F, k);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(x && dy && offsets && dkernel, SKY_ERR_INVALID, "NULL pointer");
    p.dw = dkernel;
    cudaStream_t st = (cudaStream_t)stream;
    SKY_CHECK_CUDA(cudaMemsetAsync(dkernel, 0, (size_t)p.k2 * C * F * sizeof(float), st));
    const int ntiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int groups = (p.k2 + 3) / 4;
    const int smem = 4 * 16384 + p.FC * 16384 + 64 + 4 * BLOCK_M * (int)sizeof(CornerRef) + 1024;
    const int per_sm = (227 * 1024) / (smem + 1024) > 1 ? (227 * 1024) / (smem + 1024) : 1;      // CTAs the shared memory lets an SM hold
    int parts = per_sm * 148 / (groups * p.CC);     // whole waves only (a second, part-filled wave costs as much as a full one)
    if (parts < 1) parts = 1;
    if (parts > ntiles) parts = ntiles;
    p.tiles_per_part = (ntiles + parts - 1) / parts;
    p.parts = (ntiles + p.tiles_per_part - 1) / p.tiles_per_part;
    SKY_ENSURE_DYN_SMEM(da_conv2d_wgrad_kernel, 227 * 1024);
    dim3 grid(p.parts, groups, p.CC);
    da_conv2d_wgrad_kernel<<<grid, BWD_THREADS, smem, st>>>(p);
    SKY_CHECK_LAUNCH();
    if (dbias) {
        SKY_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)F * sizeof(float), st));
        int ysplit = (p.M + 31) / 32;
        if (ysplit > 8 * 148) ysplit = 8 * 148;
        dim3 g2((F + 127) / 128, ysplit);
        col_sum_kernel<<<g2, 128, 0, st>>>(dy, dbias, p.M, F);
        SKY_CHECK_LAUNCH();
    }
    return SKY_OK;
}
