// Shared host/device helpers for the skydome_b200 library: error reporting across the C ABI and the thin inline-PTX
// layer (mbarrier, bulk async copy, tcgen05/TMEM) the sm_100a kernels are written against.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/skydome_b200.h"

namespace sky {

void set_error(const char *fmt, ...);

#define SKY_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            sky::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SKY_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

// after every kernel launch: counts it (sky_launch_count) and surfaces launch-configuration errors
void count_launch();
#define SKY_CHECK_LAUNCH()                   \
    do {                                     \
        sky::count_launch();                 \
        SKY_CHECK_CUDA(cudaGetLastError());  \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: opt in once per device, thread-safe.
// (A per-process bool would leave a second GPU of the same process — or a racing second thread — without the opt-in.)
#define SKY_ENSURE_DYN_SMEM(kernel, bytes)                                                                       \
    do {                                                                                                         \
        static std::atomic<unsigned long long> _done{ 0 };                                                       \
        int _dev = 0;                                                                                            \
        SKY_CHECK_CUDA(cudaGetDevice(&_dev));                                                                    \
        const unsigned long long _bit = 1ull << (_dev & 63);                                                     \
        if (!(_done.load(std::memory_order_acquire) & _bit)) {                                                   \
            SKY_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));    \
            _done.fetch_or(_bit, std::memory_order_release);                                                     \
        }                                                                                                        \
    } while (0)

#define SKY_REQUIRE(cond, code, ...)     \
    do {                                 \
        if (!(cond)) {                   \
            sky::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

// Padding chosen by _pad_input (distortion_aware_ops.py:125-150) for one axis at stride 1.
__host__ __device__ inline void pad_axis(int n, int k, int *before, int *total)
{
    int same_out = n, valid_out = n - k + 1;
    if (same_out == valid_out) { *before = 0; *total = 0; }
    else { *total = k - 1; *before = (k - 1) / 2; }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while when the phase is still pending; test_wait never does)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// named barrier that also OR-reduces a predicate over the participating threads
__device__ __forceinline__ bool named_bar_red_or(uint32_t id, uint32_t nthreads, uint32_t pred)
{
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.or.pred p, %2, %3, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(pred), "r"(id), "r"(nthreads)
        : "memory");
    return r != 0;
}
// wait with back-off, for roles whose barrier is far away (keeps issue slots free for the producer warps)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}

// One lane of a converged warp: the warp-uniform way to single out the thread that issues tcgen05.mma / tcgen05.commit.  With the
// surrounding loop executed by the whole warp, descriptor words stay in uniform registers; under `if (lane == 0)` the compiler moves
// them there lane by lane (R2UR + a BRA.U.ANY loop) before every MMA.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 operands, fp32 accumulate.  Issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when all previously issued tcgen05.mma of this thread have completed (implies
// tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp reads TMEM lane base+t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte rows (32 tf32), SWIZZLE_128B: 8-row groups of 1024 B (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell).  `saddr` = shared-space byte address of the tile (1024-aligned) plus the in-row
// k offset (multiples of 32 B).
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;            // leading-dim byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // stride-dim byte offset: next 8-row group
    d |= (uint64_t)1 << 46;            // descriptor version
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
// instruction descriptor: kind::tf32, D=f32, A,B K-major, M x N
__device__ __host__ inline uint32_t umma_idesc_tf32(uint32_t M, uint32_t N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint32_t f32_to_tf32_rna(float v)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// byte offset of (row, 16-byte chunk) inside a K-major SWIZZLE_128B tile
__device__ __host__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk)
{
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}
#endif  // __CUDACC__

}  // namespace sky

#ifdef __CUDACC__
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
namespace sky {
// 4-D tiled TMA load (fastest dim first: c, x, y, b) into shared memory; out-of-bounds elements are zero-filled,
// which is exactly the zero halo of _pad_input (distortion_aware_ops.py:125-150).
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap *tmap, int c, int x, int y, int b, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst_smem), "l"(tmap), "r"(c), "r"(x), "r"(y), "r"(b), "r"(bar)
        : "memory");
}
// 2-D tiled TMA load (column, row) of a row-major fp32 matrix; out-of-bounds elements are zero-filled.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *tmap, int col, int row, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(tmap), "r"(col), "r"(row), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// Host: encode a 4-D fp32 NHWC tensor map with box (box_c, box_w, box_h, 1), no swizzle.
int encode_nhwc_tensor_map(CUtensorMap *out, const float *base, int B, int h, int w, int C, int box_c, int box_w, int box_h);
// Host: dy [B, OH, OW, F] (F % 4 == 0) for the strip weight gradient, box = (32 f, 8 panoramas, box_cols columns, 1 row), 32-byte-atom 128-byte swizzle
int encode_dy_wgrad_tensor_map(CUtensorMap *out, const float *base, int B, int OH, int OW, int F, int box_cols);
// Host: encode a row-major fp32 matrix [rows][cols] with box (box_cols, box_rows), no swizzle (cols % 4 == 0, 16-byte aligned base).
int encode_2d_tensor_map(CUtensorMap *out, const float *base, long rows, long cols, int box_cols, int box_rows);
int encode_2d_tensor_map_sw(CUtensorMap *out, const float *base, long rows, long cols, int box_cols, int box_rows, int swizzle128);
}  // namespace sky
#endif
