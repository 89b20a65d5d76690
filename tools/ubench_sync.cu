// Micro-benchmark of the per-k-block control costs on sm_100a: mbarrier try_wait on a completed phase,
// tcgen05.commit, tcgen05.mma issue, fence.proxy.async.  One CTA, timings by clock64 in one thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200/csrc/sky_common.cuh"
using namespace sky;
namespace sky { void set_error(const char*, ...) {} }

__global__ void __launch_bounds__(128) k(long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 65536);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 64; ++i) mbar_init(smem_u32(bars + i), 1); fence_mbar_init(); }
    for (int i = tid; i < 16384; i += 128) reinterpret_cast<float *>(smem)[i] = 1.0f;
    if (warp == 0) { tmem_alloc(smem_u32(slot), 256); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const int N = 256;
        // complete phase 0 of barriers 0..31
        for (int i = 0; i < 32; ++i) mbar_arrive(smem_u32(bars + i));
        long long t0 = clock64();
        for (int i = 0; i < N; ++i) mbar_wait(smem_u32(bars + (i & 31)), 0);
        long long t1 = clock64();
        out[0] = (t1 - t0) / N;                       // try_wait, already complete
        t0 = clock64();
        for (int i = 0; i < N; ++i) fence_proxy_async_smem();
        t1 = clock64();
        out[1] = (t1 - t0) / N;
        // commits with nothing outstanding; each to its own barrier (32..63), then wait all
        t0 = clock64();
        for (int i = 0; i < 32; ++i) umma_commit(smem_u32(bars + 32 + i));
        t1 = clock64();
        out[2] = (t1 - t0) / 32;                      // commit issue cost
        for (int i = 0; i < 32; ++i) mbar_wait(smem_u32(bars + 32 + i), 0);
        long long t2 = clock64();
        out[3] = (t2 - t0);                           // until all 32 commits landed
        // MMA issue cost: 64 MMAs M128 N128 K8 tf32 back to back, then commit + wait
        const uint32_t idesc = umma_idesc_tf32(128, 128);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem + 16384));
        t0 = clock64();
        for (int i = 0; i < 64; ++i) umma_tf32(tmem, da, db, idesc, i != 0);
        t1 = clock64();
        out[4] = (t1 - t0) / 64;                      // issue cost per MMA
        umma_commit(smem_u32(bars + 0));              // phase 1 of barrier 0
        mbar_wait(smem_u32(bars + 0), 1);
        t2 = clock64();
        out[5] = (t2 - t0) / 64;                      // execution time per MMA (throughput)
        // single MMA + commit + wait latency
        t0 = clock64();
        umma_tf32(tmem, da, db, idesc, 1);
        umma_commit(smem_u32(bars + 1));
        mbar_wait(smem_u32(bars + 1), 1);
        t1 = clock64();
        out[6] = t1 - t0;
        // ping-pong: commit -> wait, 64 times (commit-to-visible latency, nothing outstanding)
        t0 = clock64();
        for (int i = 0; i < 64; ++i) { umma_commit(smem_u32(bars + 2)); mbar_wait(smem_u32(bars + 2), (i + 1) & 1); }
        t1 = clock64();
        out[7] = (t1 - t0) / 64;
        // arrive -> wait by the same thread
        t0 = clock64();
        for (int i = 0; i < 64; ++i) { mbar_arrive(smem_u32(bars + 3)); mbar_wait(smem_u32(bars + 3), (i + 1) & 1); }
        t1 = clock64();
        out[8] = (t1 - t0) / 64;
        // 64 MMAs of N=256
        const uint32_t idesc2 = umma_idesc_tf32(128, 256);
        t0 = clock64();
        for (int i = 0; i < 64; ++i) umma_tf32(tmem, da, db, idesc2, 1);
        umma_commit(smem_u32(bars + 4));
        mbar_wait(smem_u32(bars + 4), 0);
        t1 = clock64();
        out[9] = (t1 - t0) / 64;
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int main()
{
    long long *d, h[16] = {0};
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 128, 100 * 1024>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("try_wait(complete) %lld cyc | fence.proxy.async %lld | commit issue %lld | 32 commits landed %lld | mma issue %lld | mma exec(N128,K8) %lld | mma+commit+wait latency %lld | commit->wait %lld | arrive->wait %lld | mma exec(N256,K8) %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
    return 0;
}
