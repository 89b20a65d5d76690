// Forward distortion-aware convolution (distortion_aware_ops.py:50-123) for sm_100a.
//
//   sky_da_pack_weights     kernel [k*k*C, F] -> K-major SWIZZLE_128B tiles of 32 k-values x F_pad filters (TF32, rna)
//   sky_da_conv2d_fwd       warp-specialised implicit GEMM: producer warps evaluate the sampling geometry, gather the
//                           four corners with 128-bit loads, blend, and write the A tile in the UMMA swizzled layout;
//                           one thread streams the weight tiles with bulk async copies; one thread issues
//                           tcgen05.mma (kind::tf32) into TMEM; the producer warps then drain TMEM (tcgen05.ld), add
//                           the bias (+ LeakyReLU / residual) and store NHWC.
//   sky_da_conv2d_fwd_simt  fp32 CUDA-core restatement of the same contract (cross-check, odd shapes)
//   sky_resize_bilinear_fwd TF2 half-pixel bilinear resize (deconv2d.call :322)
#include "strip_conv.cuh"

namespace sky {

constexpr int NUM_PRODUCER_THREADS = 128;
constexpr int NUM_THREADS = NUM_PRODUCER_THREADS + 64;   // + MMA warp + weight-loader warp

// -----------------------------------------------------------------------------------------------------------------
// weight prepack
// -----------------------------------------------------------------------------------------------------------------
// packed[kb][plane][n_tile rows][32] : for k-block kb the image is exactly what the MMA expects in shared memory, so a
// stage is filled by one contiguous bulk copy.  plane 0 = tf32(rna(w)); plane 1 (3xTF32 only) = tf32(rna(w - hi)).
// k-block order: for C % 32 == 0 the tiles are stored chunk-major, kb = cc * k2 + tap (the order the conv kernels walk K:
// all taps of one 32-channel chunk, then the next chunk), so consecutive taps of a chunk are contiguous; otherwise
// k-block kb simply covers kernel rows [32 kb, 32 kb + 32).
// One block per (k-block, 32 filters) sub-tile: the variable is read along its filters (128-byte rows, coalesced), the tile is written
// along its k values (128-byte rows, 16 bytes per thread): the transpose goes through shared memory.
__global__ void __launch_bounds__(256) da_pack_weights_kernel(const float *__restrict__ kernel, float *__restrict__ packed, int K, int F, int Fp,
                                                             int KB, int planes, int C, int k2, int ldk)
{
    __shared__ float wsm[BLOCK_K][33];
    const int nsub = (Fp + 31) / 32;
    const size_t tile_floats = (size_t)Fp * BLOCK_K;
    for (int work = blockIdx.x; work < KB * nsub; work += gridDim.x) {
        const int kb = work / nsub, n0 = (work - kb * nsub) * 32;
        int kbase = kb * BLOCK_K;
        if (C % BLOCK_K == 0) {
            const int cc = kb / k2, t = kb % k2;
            kbase = t * C + cc * BLOCK_K;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < BLOCK_K * 32; e += 256) {
            const int kk = e >> 5, n = n0 + (e & 31);
            float v = 0.f;
            if (kbase + kk < K && n < F) v = __ldg(kernel + (size_t)(kbase + kk) * ldk + n);
            wsm[kk][e & 31] = v;
        }
        __syncthreads();
        uint8_t *tile = reinterpret_cast<uint8_t *>(packed + (size_t)kb * planes * tile_floats);
        {
            const int chunk = threadIdx.x & 7, nl = threadIdx.x >> 3, n = n0 + nl;
            if (n < Fp) {
                const float v0 = wsm[4 * chunk][nl], v1 = wsm[4 * chunk + 1][nl], v2 = wsm[4 * chunk + 2][nl], v3 = wsm[4 * chunk + 3][nl];
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
    }
}

// The same image for the FLIPPED, TRANSPOSED kernel of a plain conv (the data gradient run as a forward pass over dy, conv_bwd.cu),
// straight from the layer variable: the virtual kernel is kernelT[(t' * F + f) * C + c] = kernel[((k2 - 1 - t') * C + c) * F + f], its
// "input channels" are the layer's F filters and its "filters" the C channels (rows n of a tile; c0 = first channel of the slice).
// A tile row holds 32 consecutive f of one (t', c): the reads are contiguous along kk, no transposed copy in global memory.
__global__ void da_pack_weights_t_kernel(const float *__restrict__ kernel, float *__restrict__ packed, int Kt, int Cs, int Np, int KB,
                                         int planes, int C, int F, int k2, int c0)
{
    const long total = (long)KB * Np * BLOCK_K;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int kk = (int)(e % BLOCK_K);
        const int n = (int)((e / BLOCK_K) % Np);
        const int kb = (int)(e / ((long)BLOCK_K * Np));
        int kidx = kb * BLOCK_K + kk;
        if (F % BLOCK_K == 0) {
            const int cc = kb / k2, t = kb % k2;
            kidx = t * F + cc * BLOCK_K + kk;
        }
        float v = 0.f;
        if (kidx < Kt && n < Cs) {
            const int tp = kidx / F, f = kidx - tp * F;
            v = __ldg(kernel + ((size_t)(k2 - 1 - tp) * C + c0 + n) * F + f);
        }
        const uint32_t hi = f32_to_tf32_rna(v);
        const size_t tile_floats = (size_t)Np * BLOCK_K;
        const size_t o = (sw128_offset((uint32_t)n, (uint32_t)(kk >> 2)) >> 2) + (kk & 3);
        packed[((size_t)kb * planes + 0) * tile_floats + o] = __uint_as_float(hi);
        if (planes == 2) packed[((size_t)kb * planes + 1) * tile_floats + o] = __uint_as_float(f32_to_tf32_rna(v - __uint_as_float(hi)));
    }
}

// -----------------------------------------------------------------------------------------------------------------
// tensor-core forward
// -----------------------------------------------------------------------------------------------------------------
struct FwdParams {
    const float *x;
    const float *offsets;   // [h][k2][2]
    const float *packed;
    const float *bias;
    const float *residual;
    const float *aux;       // SKY_EPI_SUN_BLEND: sky prediction (log domain) [M][3]
    float threshold;
    float *y;
    double *stats;
    int B, h, w, C, F, Fp, k, k2, K, KB;
    int kb_per_split;       // split-K: k-blocks per blockIdx.y slice (KB when not split); split launches add raw partial sums into y
    int nslices;            // filter slices in blockIdx.z (each F filters, packed images / bias / y columns slice_* apart)
    size_t slice_packed_bytes;
    int ldF;                // row stride of y / residual / stats (filters of the whole layer; == F unless the launch is a filter slice)
    int in_h, in_w, ph0, pw0;
    int M;                  // B*oh*ow
    int oh, ow;             // output map (== h, w for the distortion-aware layers)
    int plain, stride;      // plain != 0: identity sampler (tf.nn.conv2d SAME), taps at (i*stride + a - ph0, j*stride + b - pw0)
    int flags;
    float slope;
};

template <int STAGES, bool SPLIT3>
struct FwdSmem {
    static constexpr int PLANES = SPLIT3 ? 2 : 1;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;                    // 16 KB per plane
    static __host__ __device__ int b_bytes(int Fp) { return Fp * BLOCK_K * 4; }
    static __host__ __device__ int stage_bytes(int Fp) { return PLANES * (A_BYTES + b_bytes(Fp)); }
    static __host__ __device__ int total_bytes(int Fp)
    {
        // stages | coord table (2 x 128 x 32 B) | barriers | tmem slot ; +1024 for manual alignment
        return STAGES * stage_bytes(Fp) + 2 * BLOCK_M * 32 + (2 * STAGES + 1) * 8 + 16 + 1024;
    }
};

// One 16-byte chunk (4 channels) of one A row: 4 corner loads, blend in the reference's add_n order (:112-113).
__device__ __forceinline__ float4 blend4(const float *__restrict__ x, const CornerRef &cr, int ch)
{
    float4 p[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        p[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cr.off[c] >= 0) p[c] = __ldg(reinterpret_cast<const float4 *>(x + cr.off[c] + ch));
    }
    float4 r;
    r.x = cr.w[0] * p[0].x; r.y = cr.w[0] * p[0].y; r.z = cr.w[0] * p[0].z; r.w = cr.w[0] * p[0].w;
#pragma unroll
    for (int c = 1; c < 4; ++c) {
        r.x = fmaf(cr.w[c], p[c].x, r.x);
        r.y = fmaf(cr.w[c], p[c].y, r.y);
        r.z = fmaf(cr.w[c], p[c].z, r.z);
        r.w = fmaf(cr.w[c], p[c].w, r.w);
    }
    return r;
}

__device__ __forceinline__ void store_a_chunk(uint8_t *a_tile, int row, int chunk, float4 v, bool split3)
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

// Sampling of one (pixel, tap): the reference's distortion-aware geometry, or the identity sampler of a plain conv.
__device__ __forceinline__ CornerRef sample_corners(const FwdParams &p, int b_img, int i, int j, int t)
{
    if (p.plain) {
        CornerRef r;
        int yy = i * p.stride + t / p.k - p.ph0, xx = j * p.stride + t % p.k - p.pw0;
        bool ok = true;
        if (p.plain == 2) {
            // transposed sampler (data gradient of a strided conv): x is dy, (i, j) a pixel of the gradient map; the tap lands on a
            // dy pixel only when (i + a - ph0, j + b - pw0) is a multiple of the stride
            yy = i + t / p.k - p.ph0; xx = j + t % p.k - p.pw0;
            ok = yy >= 0 && xx >= 0 && (yy % p.stride) == 0 && (xx % p.stride) == 0;
            yy /= p.stride; xx /= p.stride;
        }
        ok = ok && yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
        r.off[0] = ok ? ((b_img * p.h + yy) * p.w + xx) * p.C : -1;
        r.w[0] = 1.f;
        r.off[1] = r.off[2] = r.off[3] = -1;
        r.w[1] = r.w[2] = r.w[3] = 0.f;
        return r;
    }
    const float yo = __ldg(p.offsets + ((size_t)i * p.k2 + t) * 2 + 0);
    const float xo = __ldg(p.offsets + ((size_t)i * p.k2 + t) * 2 + 1);
    const Sample s = da_sample(i, j, t / p.k, t % p.k, yo, xo, p.in_h, p.in_w);
    return da_corners(s, b_img, p.h, p.w, p.C, p.ph0, p.pw0);
}

template <int STAGES, bool SPLIT3>
__global__ void __launch_bounds__(NUM_THREADS) da_conv2d_fwd_tc_kernel(const FwdParams p)
{
    using L = FwdSmem<STAGES, SPLIT3>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the .shared address space
    const int stage_bytes = L::stage_bytes(p.Fp);
    const int b_bytes = L::b_bytes(p.Fp);
    CornerRef *coord = reinterpret_cast<CornerRef *>(smem + STAGES * stage_bytes);           // [2][BLOCK_M]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * stage_bytes + 2 * BLOCK_M * 32);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 1);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tmem_full = smem_u32(bars + 2 * STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * BLOCK_M;
    const int kb_lo = blockIdx.y * p.kb_per_split, kb_hi = min(p.KB, kb_lo + p.kb_per_split);   // this CTA's k-blocks
    const bool split = p.kb_per_split < p.KB;
    const int f_slice = blockIdx.z * p.F;                                                            // first filter of this slice
    const uint8_t *packed = reinterpret_cast<const uint8_t *>(p.packed) + (size_t)blockIdx.z * p.slice_packed_bytes;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.Fp) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, NUM_PRODUCER_THREADS / 32 + 1);   // 4 producer warps + the weight loader
            mbar_init(empty0 + 8 * s, 1);                              // tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == 4) {   // TMEM allocation is warp-wide
        tmem_alloc(smem_u32(tmem_slot), tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== PRODUCERS: geometry -> gather -> blend -> swizzled A tile ===========================
        const bool fast = (p.C % BLOCK_K) == 0;
        const int chunk = tid & 7;            // 16-byte chunk inside the 128-byte row
        const int row_base = tid >> 3;        // rows row_base + 16*i
        int kb = kb_lo;
        if (fast) {
            const int cpt = p.C / BLOCK_K;    // k-blocks per tap
            // this thread's own pixel (row = tid) for the per-tap coordinate table
            const int m = m0 + tid;
            const bool m_ok = m < p.M;
            const int j = m % p.ow, i = (m / p.ow) % p.oh, b_img = m / (p.ow * p.oh);
            (void)cpt;
            for (int kk = kb_lo; kk < kb_hi; ++kk) {      // chunk-major k-block order (matches the packed weights)
                const int cc = kk / p.k2, t = kk % p.k2;
                CornerRef cr;
                if (m_ok) {
                    cr = sample_corners(p, b_img, i, j, t);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) { cr.off[c] = -1; cr.w[c] = 0.f; }
                }
                CornerRef *tab = coord + (kk & 1) * BLOCK_M;
                tab[tid] = cr;
                named_bar_sync(1, NUM_PRODUCER_THREADS);
                {
                    const int s = (kb - kb_lo) % STAGES;
                    mbar_wait(empty0 + 8 * s, (((kb - kb_lo) / STAGES) & 1) ^ 1);
                    uint8_t *a_tile = smem + s * stage_bytes;
                    const int ch = cc * BLOCK_K + chunk * 4;
#pragma unroll 4
                    for (int r = 0; r < BLOCK_M / 16; ++r) {
                        const int row = row_base + 16 * r;
                        const CornerRef c = tab[row];
                        store_a_chunk(a_tile, row, chunk, blend4(p.x, c, ch), SPLIT3);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8 * s);
                    ++kb;
                }
            }
        } else if ((p.C & 3) == 0) {
            // C % 4 == 0 (the 6 -> 8 channel-padded sunRadNet input, sunrad_net.py:37): a 16-byte chunk of the A row is four
            // channels of ONE tap, so it is one sampling + one 128-bit gather per corner, like the fast path
            int pj[BLOCK_M / 16], pi[BLOCK_M / 16], pb[BLOCK_M / 16];
#pragma unroll
            for (int r = 0; r < BLOCK_M / 16; ++r) {
                const int m = m0 + row_base + 16 * r;
                pj[r] = m % p.ow; pi[r] = (m / p.ow) % p.oh; pb[r] = m < p.M ? m / (p.ow * p.oh) : -1;
            }
            for (; kb < kb_hi; ++kb) {
                const int s = (kb - kb_lo) % STAGES;
                const int kidx = kb * BLOCK_K + chunk * 4;
                const bool k_ok = kidx < p.K;
                const int t = k_ok ? kidx / p.C : 0, c = kidx % p.C;
                mbar_wait(empty0 + 8 * s, (((kb - kb_lo) / STAGES) & 1) ^ 1);
                uint8_t *a_tile = smem + s * stage_bytes;
#pragma unroll
                for (int r = 0; r < BLOCK_M / 16; ++r) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k_ok && pb[r] >= 0) v = blend4(p.x, sample_corners(p, pb[r], pi[r], pj[r], t), c);
                    store_a_chunk(a_tile, row_base + 16 * r, chunk, v, SPLIT3);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * s);
            }
        } else {
            // generic channel counts (3-channel image layers, C not a multiple of 4): geometry per element
            for (; kb < kb_hi; ++kb) {
                const int s = (kb - kb_lo) % STAGES;
                mbar_wait(empty0 + 8 * s, (((kb - kb_lo) / STAGES) & 1) ^ 1);
                uint8_t *a_tile = smem + s * stage_bytes;
                for (int r = 0; r < BLOCK_M / 16; ++r) {
                    const int row = row_base + 16 * r;
                    const int m = m0 + row;
                    float v[4] = { 0.f, 0.f, 0.f, 0.f };
                    if (m < p.M) {
                        const int j = m % p.ow, i = (m / p.ow) % p.oh, b_img = m / (p.ow * p.oh);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int kidx = kb * BLOCK_K + chunk * 4 + e;
                            if (kidx < p.K) {
                                const int t = kidx / p.C, c = kidx % p.C;
                                const CornerRef cr = sample_corners(p, b_img, i, j, t);
                                float acc = 0.f;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float px = cr.off[q] >= 0 ? __ldg(p.x + cr.off[q] + c) : 0.f;
                                    acc = (q == 0) ? cr.w[0] * px : fmaf(cr.w[q], px, acc);
                                }
                                v[e] = acc;
                            }
                        }
                    }
                    store_a_chunk(a_tile, row, chunk, make_float4(v[0], v[1], v[2], v[3]), SPLIT3);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full0 + 8 * s);
            }
        }

        // =========================== EPILOGUE: TMEM -> registers -> bias/activation/residual -> NHWC ===========================
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int m = m0 + warp * 32 + lane;         // TMEM lane == tile row
        const uint32_t taddr_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool vec_ok = (p.F % 4) == 0 && (p.ldF % 4) == 0;
        for (int c0 = 0; c0 < p.Fp; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x16(taddr_row + (uint32_t)c0, r);
            tmem_ld_wait();
            if (m < p.M && split) {
                // split-K: add this CTA's partial sums to y (zeroed by the launcher); bias / activation are applied by
                // conv_finalize_kernel once every split has landed
                float *dst = p.y + (size_t)m * p.ldF + f_slice + c0;
                if (vec_ok && c0 + 16 <= p.F) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + q), "f"(__uint_as_float(r[q])),
                                     "f"(__uint_as_float(r[q + 1])), "f"(__uint_as_float(r[q + 2])), "f"(__uint_as_float(r[q + 3]))
                                     : "memory");
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c0 + q < p.F) atomicAdd(dst + q, __uint_as_float(r[q]));
                }
            } else if (m < p.M) {
                float o[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int f = c0 + q;
                    float val = __uint_as_float(r[q]);
                    if (f < p.F) {
                        if (p.bias) val += __ldg(p.bias + f_slice + f);
                        if (p.flags & SKY_EPI_LEAKY_RELU) val = val > 0.f ? val : val * p.slope;
                        if (p.flags & SKY_EPI_RESIDUAL) val += __ldg(p.residual + (size_t)m * p.ldF + f_slice + f);
                        if (p.flags & SKY_EPI_MASK) val *= __ldg(p.residual + (size_t)m * p.ldF + f_slice + f) > 0.f ? 1.f : p.slope;
                        if (p.flags & SKY_EPI_RELU) val = fmaxf(val, 0.f);
                    }
                    o[q] = val;
                }
                if ((p.flags & SKY_EPI_SUN_BLEND) && c0 == 0) sun_blend3(o, p.aux + (size_t)m * 3, p.threshold);   // F == 3
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int f = c0 + q;
                    if (f < p.F) {
                        if (p.flags & SKY_EPI_LOG_DECOMPRESS) o[q] = (expf(o[q] * 2.3978953f) - 1.f) / 10.f;   // log(11) as fp32
                        if (p.stats) {   // generic path: plain atomics (the band-staged kernel reduces per tile first)
                            double *st = p.stats + ((size_t)(m / (p.oh * p.ow)) * p.ldF + f_slice + f) * 2;
                            atomicAdd(st, (double)o[q]);
                            atomicAdd(st + 1, (double)o[q] * (double)o[q]);
                        }
                    }
                }
                float *dst = p.y + (size_t)m * p.ldF + f_slice + c0;
                if (vec_ok && c0 + 16 <= p.F) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4) *reinterpret_cast<float4 *>(dst + q) = make_float4(o[q], o[q + 1], o[q + 2], o[q + 3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c0 + q < p.F) dst[q] = o[q];
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // =========================== MMA ISSUER (the warp walks the loop, one elected lane issues) ===========================
        {
            const uint32_t idesc = umma_idesc_tf32(BLOCK_M, (uint32_t)p.Fp);
            uint32_t s = 0, phase = 0;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                mbar_wait(full0 + 8 * s, phase);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
                const uint32_t b_hi = a_hi + L::PLANES * L::A_BYTES;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                        const uint32_t koff = ks * UMMA_K * 4;
                        const uint64_t da = umma_desc_kmajor_sw128(a_hi + koff);
                        const uint64_t db = umma_desc_kmajor_sw128(b_hi + koff);
                        umma_tf32(tmem_base, da, db, idesc, ((kb - kb_lo) | ks) != 0);
                        if (SPLIT3) {
                            const uint64_t da_lo = umma_desc_kmajor_sw128(a_hi + L::A_BYTES + koff);
                            const uint64_t db_lo = umma_desc_kmajor_sw128(b_hi + b_bytes + koff);
                            umma_tf32(tmem_base, da_lo, db, idesc, 1);
                            umma_tf32(tmem_base, da, db_lo, idesc, 1);
                        }
                    }
                    umma_commit(empty0 + 8 * s);            // smem stage free once these MMAs have read it
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(tmem_full);         // accumulator complete
            __syncwarp();
        }
    } else {
        // =========================== WEIGHT LOADER (one elected lane, bulk async copies) ===========================
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(L::PLANES * b_bytes);
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                const int s = (kb - kb_lo) % STAGES;
                mbar_wait(empty0 + 8 * s, (((kb - kb_lo) / STAGES) & 1) ^ 1);
                const uint32_t dst = smem_u32(smem + s * stage_bytes + L::PLANES * L::A_BYTES);
                mbar_arrive_expect_tx(full0 + 8 * s, bytes);
                bulk_g2s(dst, packed + (size_t)kb * bytes, bytes, full0 + 8 * s);
            }
        }
        __syncwarp();
    }

    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// -----------------------------------------------------------------------------------------------------------------
// CUDA-core restatement
// -----------------------------------------------------------------------------------------------------------------
__global__ void da_conv2d_fwd_simt_kernel(const float *__restrict__ x, const float *__restrict__ offsets,
                                          const float *__restrict__ kernel, const float *__restrict__ bias,
                                          float *__restrict__ y, int B, int h, int w, int C, int F, int k)
{
    const int k2 = k * k;
    int ph0, pht, pw0, pwt;
    pad_axis(h, k, &ph0, &pht);
    pad_axis(w, k, &pw0, &pwt);
    const long total = (long)B * h * w * F;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int f = (int)(o % F);
        const long m = o / F;
        const int j = (int)(m % w), i = (int)((m / w) % h), b_img = (int)(m / ((long)w * h));
        float acc = 0.f;
        for (int t = 0; t < k2; ++t) {
            const float yo = offsets[((size_t)i * k2 + t) * 2 + 0], xo = offsets[((size_t)i * k2 + t) * 2 + 1];
            const Sample s = da_sample(i, j, t / k, t % k, yo, xo, h + pht, w + pwt);
            const CornerRef cr = da_corners(s, b_img, h, w, C, ph0, pw0);
            for (int c = 0; c < C; ++c) {
                float pix = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float px = cr.off[q] >= 0 ? x[cr.off[q] + c] : 0.f;
                    pix = (q == 0) ? __fmul_rn(cr.w[0], px) : __fadd_rn(pix, __fmul_rn(cr.w[q], px));   // add_n order
                }
                acc = fmaf(pix, kernel[((size_t)t * C + c) * F + f], acc);
            }
        }
        y[o] = acc + bias[f];
    }
}

// TF2 bilinear resize, half-pixel centres: src = (dst + 0.5) * (in/out) - 0.5 ; lower = max(floor, 0) ;
// upper = min(ceil, n-1) ; lerp = src - floor(src) ; top/bottom lerp in x, then lerp in y.  VEC channels per thread.
template <int VEC>
__global__ void resize_bilinear_kernel(const float *__restrict__ x, float *__restrict__ y, int B, int h, int w, int C,
                                       int oh, int ow)
{
    const float sy = __fdiv_rn((float)h, (float)oh), sx = __fdiv_rn((float)w, (float)ow);
    const int cv = C / VEC;
    const long total = (long)B * oh * ow * cv;
    for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
        const int c = (int)(o % cv) * VEC;
        const int ox = (int)((o / cv) % ow), oy = (int)((o / ((long)cv * ow)) % oh), b = (int)(o / ((long)cv * ow * oh));
        const float fy = __fsub_rn(__fmul_rn(__fadd_rn((float)oy, 0.5f), sy), 0.5f);
        const float fx = __fsub_rn(__fmul_rn(__fadd_rn((float)ox, 0.5f), sx), 0.5f);
        const float fly = floorf(fy), flx = floorf(fx);
        const int ylo = max((int)fly, 0), yhi = min((int)ceilf(fy), h - 1);
        const int xlo = max((int)flx, 0), xhi = min((int)ceilf(fx), w - 1);
        const float ly = __fsub_rn(fy, fly), lx = __fsub_rn(fx, flx);
        const float *img = x + (size_t)b * h * w * C + c;
        const float *ptl = img + ((size_t)ylo * w + xlo) * C, *ptr = img + ((size_t)ylo * w + xhi) * C;
        const float *pbl = img + ((size_t)yhi * w + xlo) * C, *pbr = img + ((size_t)yhi * w + xhi) * C;
        float tl[VEC], tr[VEC], bl[VEC], br[VEC], out[VEC];
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(tl) = __ldg(reinterpret_cast<const float4 *>(ptl));
            *reinterpret_cast<float4 *>(tr) = __ldg(reinterpret_cast<const float4 *>(ptr));
            *reinterpret_cast<float4 *>(bl) = __ldg(reinterpret_cast<const float4 *>(pbl));
            *reinterpret_cast<float4 *>(br) = __ldg(reinterpret_cast<const float4 *>(pbr));
        } else {
            tl[0] = ptl[0]; tr[0] = ptr[0]; bl[0] = pbl[0]; br[0] = pbr[0];
        }
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
            const float top = __fadd_rn(tl[u], __fmul_rn(__fsub_rn(tr[u], tl[u]), lx));
            const float bot = __fadd_rn(bl[u], __fmul_rn(__fsub_rn(br[u], bl[u]), lx));
            out[u] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
        }
        float *dst = y + (((size_t)b * oh + oy) * ow + ox) * C + c;
        if (VEC == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<float4 *>(out);
        else dst[0] = out[0];
    }
}

// y = act(y + bias) after a split-K launch (flags: SKY_EPI_LEAKY_RELU, or SKY_EPI_MASK with its mask source)
__global__ void conv_finalize_kernel(float *__restrict__ y, const float *__restrict__ bias, const float *__restrict__ mask, long total4,
                                     int F4, int flags, float slope)
{
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total4; e += (long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4 *>(y)[e];
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + (int)(e % F4));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (flags & SKY_EPI_MASK) {
            const float4 mk = __ldg(reinterpret_cast<const float4 *>(mask) + e);
            v.x *= mk.x > 0.f ? 1.f : slope; v.y *= mk.y > 0.f ? 1.f : slope;
            v.z *= mk.z > 0.f ? 1.f : slope; v.w *= mk.w > 0.f ? 1.f : slope;
        }
        if (flags & SKY_EPI_LEAKY_RELU) {
            v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
            v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
        }
        reinterpret_cast<float4 *>(y)[e] = v;
    }
}

static int check_conv_args(int B, int h, int w, int C, int F, int k)
{
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && F > 0, SKY_ERR_INVALID, "non-positive dimension (B=%d h=%d w=%d C=%d F=%d)", B, h, w, C, F);
    SKY_REQUIRE(k % 2 == 1, SKY_ERR_EVEN_KERNEL, "kernel_size must be odd number, current kernel size : %d", k);
    SKY_REQUIRE(k >= 3 && k <= 15, SKY_ERR_UNSUPPORTED, "kernel_size %d outside the supported odd range 3..15", k);
    SKY_REQUIRE((long)B * h * w * (long)(C > F ? C : F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    return SKY_OK;
}

template <int STAGES, bool SPLIT3>
static int launch_fwd(const FwdParams &p, cudaStream_t st)
{
    using L = FwdSmem<STAGES, SPLIT3>;
    const int smem = L::total_bytes(p.Fp);
    SKY_ENSURE_DYN_SMEM((da_conv2d_fwd_tc_kernel<STAGES, SPLIT3>), 227 * 1024);
    SKY_REQUIRE(smem <= 227 * 1024, SKY_ERR_UNSUPPORTED, "shared memory %d B exceeds 227 KB (F=%d)", smem, p.F);
    const int tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const int ksplit = (p.KB + p.kb_per_split - 1) / p.kb_per_split;
    if (ksplit > 1) SKY_CHECK_CUDA(cudaMemsetAsync(p.y, 0, (size_t)p.M * p.ldF * sizeof(float), st));
    da_conv2d_fwd_tc_kernel<STAGES, SPLIT3><<<dim3(tiles, ksplit, p.nslices), NUM_THREADS, smem, st>>>(p);
    SKY_CHECK_LAUNCH();
    if (ksplit > 1) {
        const long total4 = (long)p.M * p.ldF / 4;
        int blocks = (int)((total4 + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        conv_finalize_kernel<<<blocks, 256, 0, st>>>(p.y, p.bias, p.residual, total4, p.ldF / 4, p.flags, p.slope);
        SKY_CHECK_LAUNCH();
    }
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

// Layers with more than 256 filters (sunRadNet d4: 512, sunrad_net.py:40) are run as slices of <= 256 filters, each with
// its own packed image; the images are stored back to back.
static inline int slice_count(int F) { return (F + 255) / 256; }
static inline int slice_filters(int F, int s) { return (F - 256 * s) < 256 ? (F - 256 * s) : 256; }
static inline size_t slice_bytes(int C, int Fs, int k, int math_mode)
{
    const int K = k * k * C, KB = (K + BLOCK_K - 1) / BLOCK_K, Fp = f_pad_of(Fs);
    const int planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    return (size_t)KB * planes * Fp * BLOCK_K * sizeof(float);
}

extern "C" size_t sky_da_packed_weight_bytes(int C, int F, int k, int math_mode)
{
    if (C <= 0 || F <= 0 || k <= 0) return 0;
    size_t total = 0;
    for (int s = 0; s < slice_count(F); ++s) total += slice_bytes(C, slice_filters(F, s), k, math_mode);
    return total;
}

extern "C" int sky_da_pack_weights(const float *kernel, void *packed, int C, int F, int k, int math_mode, void *stream)
{
    SKY_REQUIRE(kernel && packed, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(C > 0 && F > 0 && k > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    const int K = k * k * C, KB = (K + BLOCK_K - 1) / BLOCK_K;
    const int planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    uint8_t *dst = (uint8_t *)packed;
    for (int s = 0; s < slice_count(F); ++s) {
        const int Fs = slice_filters(F, s), Fp = f_pad_of(Fs);
        const int work = KB * ((Fp + 31) / 32), blocks = work < 148 * 8 ? work : 148 * 8;
        da_pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kernel + 256 * s, (float *)dst, K, Fs, Fp, KB, planes, C, k * k, F);
        SKY_CHECK_LAUNCH();
        dst += slice_bytes(C, Fs, k, math_mode);
    }
    return SKY_OK;
}

// Packed image of the flipped, transposed kernel of a plain conv layer [k*k*C, F]: byte for byte what sky_conv2d_transpose_weights
// (flip = 1) followed by sky_da_pack_weights(F, C, k) produces (sky_da_packed_weight_bytes(F, C, k, math_mode) bytes), in one pass.
extern "C" int sky_conv2d_pack_weights_t(const float *kernel, void *packed, int C, int F, int k, int math_mode, void *stream)
{
    SKY_REQUIRE(kernel && packed, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(C > 0 && F > 0 && k > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    const int Kt = k * k * F, KB = (Kt + BLOCK_K - 1) / BLOCK_K;
    const int planes = math_mode == SKY_MATH_3XTF32 ? 2 : 1;
    uint8_t *dst = (uint8_t *)packed;
    for (int s = 0; s < slice_count(C); ++s) {
        const int Cs = slice_filters(C, s), Np = f_pad_of(Cs);
        const long total = (long)KB * Np * BLOCK_K;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        da_pack_weights_t_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kernel, (float *)dst, Kt, Cs, Np, KB, planes, C, F, k * k, 256 * s);
        SKY_CHECK_LAUNCH();
        dst += slice_bytes(F, Cs, k, math_mode);
    }
    return SKY_OK;
}

int sky::launch_fwd_direct(const FwdArgs &a)
{
    FwdParams p;
    p.x = a.x; p.offsets = a.offsets; p.packed = a.packed; p.bias = a.bias; p.residual = a.residual; p.y = a.y; p.stats = a.stats;
    p.aux = a.aux; p.threshold = a.threshold; p.ldF = a.ldF > 0 ? a.ldF : a.F;
    p.B = a.B; p.h = a.h; p.w = a.w; p.C = a.C; p.F = a.F; p.Fp = f_pad_of(a.F); p.k = a.k; p.k2 = a.k * a.k;
    p.K = a.k * a.k * a.C; p.KB = (p.K + BLOCK_K - 1) / BLOCK_K;
    int pht, pwt;
    pad_axis(a.h, a.k, &p.ph0, &pht);
    pad_axis(a.w, a.k, &p.pw0, &pwt);
    p.in_h = a.h + pht; p.in_w = a.w + pwt;
    p.oh = a.h; p.ow = a.w; p.plain = 0; p.stride = 1;
    if (a.plain_stride > 0) {
        // tf.nn.conv2d SAME: out = ceil(n/s), total pad = max((out-1)*s + k - n, 0), the smaller half in front
        p.plain = 1; p.stride = a.plain_stride;
        p.oh = (a.h + p.stride - 1) / p.stride; p.ow = (a.w + p.stride - 1) / p.stride;
        int th = (p.oh - 1) * p.stride + a.k - a.h, tw = (p.ow - 1) * p.stride + a.k - a.w;
        p.ph0 = (th > 0 ? th : 0) / 2; p.pw0 = (tw > 0 ? tw : 0) / 2;
    }
    if (a.transposed) {
        p.plain = 2; p.stride = a.plain_stride; p.oh = a.out_h; p.ow = a.out_w; p.ph0 = a.tp_ph0; p.pw0 = a.tp_pw0;
    }
    p.M = a.B * p.oh * p.ow;
    p.flags = a.flags; p.slope = a.slope;
    p.nslices = a.nslices > 0 ? a.nslices : 1;
    p.slice_packed_bytes = slice_bytes(a.C, a.F, a.k, a.math_mode);
    p.kb_per_split = p.KB;
    {
        // Few output tiles and a long K (sunRadNet d3 / d4: 16 tiles, 64 / 128 k-blocks): split K over blockIdx.y so the SMs are
        // used; the partial sums meet in y through vector reductions and conv_finalize_kernel applies bias + LeakyReLU.
        const int tiles = (p.M + BLOCK_M - 1) / BLOCK_M * p.nslices;
        const bool simple_epilogue = !(a.flags & ~(SKY_EPI_LEAKY_RELU | SKY_EPI_MASK | SKY_EPI_FORCE_DIRECT)) && a.stats == nullptr &&
                                     (a.flags & (SKY_EPI_LEAKY_RELU | SKY_EPI_MASK)) != (SKY_EPI_LEAKY_RELU | SKY_EPI_MASK);
        if (simple_epilogue && tiles * 2 <= 148 && p.KB >= 16 && (p.ldF % 4) == 0 && (a.F % 4) == 0 && (a.ldF == 0 || a.nslices > 0)) {
            int ksplit = 148 / tiles;
            if (ksplit > p.KB / 4) ksplit = p.KB / 4;
            if (ksplit > 1) p.kb_per_split = (p.KB + ksplit - 1) / ksplit;
        }
    }
    if (a.math_mode == SKY_MATH_TF32) {
        // 3 stages of 16 KB + Fp*128 B keep two CTAs resident per SM for F <= 128
        return p.Fp <= 128 ? launch_fwd<3, false>(p, a.stream) : launch_fwd<4, false>(p, a.stream);
    }
    return p.Fp <= 128 ? launch_fwd<3, true>(p, a.stream) : launch_fwd<2, true>(p, a.stream);
}

extern "C" int sky_da_conv2d_fwd(const float *x, const float *offsets, const float *offsets_host, const void *packed,
                                 const float *bias, float *y, const float *residual, double *stats, int B, int h, int w,
                                 int C, int F, int k, int epilogue_flags, float slope, int math_mode, void *stream)
{
    int rc = check_conv_args(B, h, w, C, F, k);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(x && offsets && packed && bias && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(F <= 256, SKY_ERR_UNSUPPORTED, "filters=%d > 256 not supported by the tensor-core path", F);
    SKY_REQUIRE(!(epilogue_flags & SKY_EPI_RESIDUAL) || residual, SKY_ERR_INVALID, "SKY_EPI_RESIDUAL without a residual pointer");
    SKY_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)packed & 15) == 0, SKY_ERR_INVALID, "x, y and packed must be 16-byte aligned");
    SKY_REQUIRE(math_mode == SKY_MATH_TF32 || math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", math_mode);
    FwdArgs a;
    a.x = x; a.offsets = offsets; a.offsets_host = offsets_host; a.packed = (const float *)packed; a.bias = bias;
    a.residual = residual; a.y = y; a.stats = stats; a.B = B; a.h = h; a.w = w; a.C = C; a.F = F; a.k = k;
    a.flags = epilogue_flags; a.slope = slope; a.math_mode = math_mode; a.stream = (cudaStream_t)stream;
    a.plain_stride = 0;
    if (offsets_host != nullptr && (C % BLOCK_K) == 0 && !(epilogue_flags & SKY_EPI_FORCE_DIRECT)) {
        rc = launch_fwd_band(a);
        if (rc != SKY_ERR_UNSUPPORTED) return rc;
    }
    return launch_fwd_direct(a);
}

// Plain SAME conv, any filter count: slices of <= 256 filters, each a launch over its own packed image.
static int conv2d_plain(FwdArgs a, int stride)
{
    SKY_REQUIRE(a.B > 0 && a.h > 0 && a.w > 0 && a.C > 0 && a.F > 0, SKY_ERR_INVALID, "non-positive dimension");
    SKY_REQUIRE(a.k >= 1 && a.k <= 15, SKY_ERR_UNSUPPORTED, "kernel size %d outside 1..15", a.k);
    SKY_REQUIRE(stride == 1 || stride == 2, SKY_ERR_UNSUPPORTED, "stride %d not supported (the path uses 1 and 2)", stride);
    SKY_REQUIRE(a.x && a.packed && a.y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(!(a.flags & (SKY_EPI_RESIDUAL | SKY_EPI_MASK)) || a.residual, SKY_ERR_INVALID, "SKY_EPI_RESIDUAL / SKY_EPI_MASK without its tensor");
    SKY_REQUIRE((a.flags & (SKY_EPI_RESIDUAL | SKY_EPI_MASK)) != (SKY_EPI_RESIDUAL | SKY_EPI_MASK), SKY_ERR_INVALID,
                "SKY_EPI_RESIDUAL and SKY_EPI_MASK share one tensor argument");
    SKY_REQUIRE(!(a.flags & SKY_EPI_SUN_BLEND) || (a.aux && a.F == 3 && a.threshold > 0.f), SKY_ERR_INVALID,
                "SKY_EPI_SUN_BLEND needs the sky prediction, 3 filters and a positive threshold");
    SKY_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.y & 15) == 0 && ((uintptr_t)a.packed & 15) == 0, SKY_ERR_INVALID, "x, y and packed must be 16-byte aligned");
    SKY_REQUIRE(a.math_mode == SKY_MATH_TF32 || a.math_mode == SKY_MATH_3XTF32, SKY_ERR_INVALID, "unknown math_mode %d", a.math_mode);
    SKY_REQUIRE((long)a.B * a.h * a.w * (long)(a.C > a.F ? a.C : a.F) < (1L << 31), SKY_ERR_UNSUPPORTED, "tensor exceeds 2^31 elements");
    a.offsets = nullptr; a.offsets_host = nullptr; a.plain_stride = stride;
    const int F = a.F;
    if (F <= 4) {
        // conv1_f / conv1_u (7x7, 32 -> 3): on the row-strip kernel the 49 taps are 7 strips and N = 16 MMAs; the fp32 small-filter kernel
        // keeps the layers the strip kernel does not take (the fused inference tail, C % 32 != 0)
        // (only the raw conv + bias of the train step: the inference tails, fused or not, stay on ONE fp32 kernel so that
        // sky_conv2d_fwd_blend and conv + sky_blend_split remain the same arithmetic bit for bit)
        if ((a.C % BLOCK_K) == 0 && a.flags == 0) {
            int rc = launch_fwd_strip_plain(a);
            if (rc != SKY_ERR_UNSUPPORTED) return rc;
        }
        int rc = launch_fwd_smallf(a);
        if (rc != SKY_ERR_UNSUPPORTED) return rc;
    }
    if (F <= 256) {
        if ((a.C % BLOCK_K) == 0 && !(a.flags & SKY_EPI_FORCE_DIRECT)) {
            int rc = (a.flags & SKY_EPI_FORCE_BAND) ? SKY_ERR_UNSUPPORTED : launch_fwd_strip_plain(a);   // row strips: k strips instead of k*k im2col tiles
            if (rc != SKY_ERR_UNSUPPORTED) return rc;
            rc = launch_fwd_band(a);            // identity sampler in the band-staged kernel (any stride the band fits)
            if (rc != SKY_ERR_UNSUPPORTED) return rc;
        }
        return launch_fwd_direct(a);
    }
    if (F % 256 == 0) {          // equal slices: one launch, slices in blockIdx.z / blockIdx.y
        FwdArgs b = a;
        b.F = 256; b.ldF = F; b.nslices = F / 256;
        if ((a.C % BLOCK_K) == 0 && !(a.flags & (SKY_EPI_FORCE_DIRECT | SKY_EPI_FORCE_BAND))) {
            int rc = launch_fwd_strip_plain(b);
            if (rc != SKY_ERR_UNSUPPORTED) return rc;
        }
        return launch_fwd_direct(b);
    }
    const uint8_t *packed = (const uint8_t *)a.packed;
    for (int s = 0; s < slice_count(F); ++s) {
        FwdArgs b = a;
        b.F = slice_filters(F, s); b.ldF = F;
        b.packed = (const float *)packed; b.bias = a.bias + 256 * s; b.y = a.y + 256 * s;
        b.residual = a.residual ? a.residual + 256 * s : nullptr;
        b.stats = a.stats ? a.stats + 2 * 256 * s : nullptr;
        int rc = launch_fwd_direct(b);
        if (rc != SKY_OK) return rc;
        packed += slice_bytes(a.C, b.F, a.k, a.math_mode);
    }
    return SKY_OK;
}

int sky::conv2d_plain_entry(FwdArgs a, int stride) { return conv2d_plain(a, stride); }

extern "C" int sky_conv2d_fwd(const float *x, const void *packed, const float *bias, float *y, const float *residual, double *stats,
                              int B, int h, int w, int C, int F, int k, int stride, int epilogue_flags, float slope,
                              int math_mode, void *stream)
{
    SKY_REQUIRE(!(epilogue_flags & SKY_EPI_SUN_BLEND), SKY_ERR_INVALID, "SKY_EPI_SUN_BLEND is taken by sky_conv2d_fwd_blend");
    FwdArgs a;
    a.x = x; a.packed = (const float *)packed; a.bias = bias;
    a.residual = residual; a.y = y; a.stats = stats; a.B = B; a.h = h; a.w = w; a.C = C; a.F = F; a.k = k;
    a.flags = epilogue_flags; a.slope = slope; a.math_mode = math_mode; a.stream = (cudaStream_t)stream;
    return conv2d_plain(a, stride);
}

extern "C" int sky_conv2d_fwd_blend(const float *x, const void *packed, const float *bias, float *y, const float *residual,
                                    const float *sky_gamma, float threshold, int B, int h, int w, int C, int k,
                                    int epilogue_flags, float slope, int math_mode, void *stream)
{
    FwdArgs a;
    a.x = x; a.packed = (const float *)packed; a.bias = bias;
    a.residual = residual; a.y = y; a.stats = nullptr; a.B = B; a.h = h; a.w = w; a.C = C; a.F = 3; a.k = k;
    a.aux = sky_gamma; a.threshold = threshold;
    a.flags = epilogue_flags | SKY_EPI_SUN_BLEND; a.slope = slope; a.math_mode = math_mode; a.stream = (cudaStream_t)stream;
    return conv2d_plain(a, 1);
}

extern "C" int sky_da_conv2d_fwd_simt(const float *x, const float *offsets, const float *kernel, const float *bias, float *y,
                                      int B, int h, int w, int C, int F, int k, void *stream)
{
    int rc = check_conv_args(B, h, w, C, F, k);
    if (rc != SKY_OK) return rc;
    SKY_REQUIRE(x && offsets && kernel && bias && y, SKY_ERR_INVALID, "NULL pointer");
    const long total = (long)B * h * w * F;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    da_conv2d_fwd_simt_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, offsets, kernel, bias, y, B, h, w, C, F, k);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}

extern "C" int sky_resize_bilinear_fwd(const float *x, float *y, int B, int h, int w, int C, int oh, int ow, void *stream)
{
    SKY_REQUIRE(x && y, SKY_ERR_INVALID, "NULL pointer");
    SKY_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && oh > 0 && ow > 0, SKY_ERR_INVALID, "non-positive dimension");
    const bool vec = (C % 4 == 0) && (((uintptr_t)x & 15) == 0) && (((uintptr_t)y & 15) == 0);
    const long total = (long)B * oh * ow * (vec ? C / 4 : C);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (vec) resize_bilinear_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, B, h, w, C, oh, ow);
    else resize_bilinear_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, B, h, w, C, oh, ow);
    SKY_CHECK_LAUNCH();
    return SKY_OK;
}
