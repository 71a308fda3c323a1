// Self-test harness for the tcgen05 building blocks: runs ONE 128 x N x (16*steps) UMMA problem whose operand
// images (exact shared-memory byte layouts) and descriptor fields come from the caller, so tests/ can pin the
// descriptor conventions (no-swizzle core matrices, 16-byte shifted starts, SW128 K-advance) against a matmul.
#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100.cuh"

namespace bcad {

struct SelftestParams {
    int N, steps;
    int a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout;
    int a_koff[64], b_koff[64];
    int mode;      // 0: bf16 operands, fp32 accumulators -> D fp32 [128][N];  1: fp16 operands, FP16 accumulators -> the raw 32-bit TMEM
                   //    cells [128][N/2] (pins how tcgen05 packs half-precision accumulators: two per column)
};

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint4* __restrict__ a_img, int a_bytes, const uint4* __restrict__ b_img, int b_bytes,
                     SelftestParams p, float* __restrict__ D) {
    using namespace sm100;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
    for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base, 256);
        tmem_relinquish();
    }
    fence_proxy_async();           // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = p.mode == 1 ? (make_idesc_f16(128, p.N) & ~(1u << 4)) : make_idesc_bf16(128, p.N);     // D format 0 = F16
        for (int k = 0; k < p.steps; ++k) {
            const uint64_t ad = make_smem_desc(smem_u32(sa) + p.a_koff[k], p.a_lbo, p.a_sbo, p.a_layout);
            const uint64_t bd = make_smem_desc(smem_u32(sb) + p.b_koff[k], p.b_lbo, p.b_sbo, p.b_layout);
            umma_bf16(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (p.mode == 1) {
        for (int c0 = 0; c0 < p.N / 2; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            const int row = warp * 32 + (tid & 31);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (c0 + j < p.N / 2) D[(size_t)row * (p.N / 2) + c0 + j] = v[j];
        }
    } else
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        const int row = warp * 32 + (tid & 31);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < p.N) D[(size_t)row * p.N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- micro-benchmarks (design evidence, not product): cycles for `reps` back-to-back UMMAs on fixed smem operands,
// and for `reps` TMEM->register loads by `nwarps` warps.  out[0] = MMA cycles (CTA 0), out[1] = TMEM-load cycles, out[3] = slowest
// CTA's MMA cycles.  `grid` CTAs run the same loop (one per SM: is the per-MMA time a property of the SM or of the loaded
// chip?); with `concurrent` the TMEM loads (warps 1..ld_warps) and `st_warps` warps of 16-byte shared-memory stores run WHILE
// thread 0 issues the MMAs (the fused conv kernel's situation) instead of after them.
__global__ void __launch_bounds__(256, 1)
umma_bench_kernel(int N, int a_layout, int b_layout, int a_lbo, int a_sbo, int b_lbo, int b_sbo, int reps, int ld_warps,
                  int ld_x16, int a_off, int alternate, int concurrent, int st_warps, uint4* __restrict__ gscratch, long long* __restrict__ out) {
    using namespace sm100;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (64 * 1024) / 16; i += 256) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0, 0);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        done = 0;
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base, 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t ad = make_smem_desc(smem_u32(smem) + a_off, a_lbo, a_sbo, a_layout);
        const uint64_t ad2 = make_smem_desc(smem_u32(smem) + a_off + (alternate ? 4096 : 0), a_lbo, a_sbo, a_layout);
        const uint64_t bd = make_smem_desc(smem_u32(smem) + 32768, b_lbo, b_sbo, b_layout);
        const uint64_t bd2 = make_smem_desc(smem_u32(smem) + 32768 + (alternate ? 2048 : 0), b_lbo, b_sbo, b_layout);
        const long long t0 = clock64();
        for (int k = 0; k < reps; k += 2) {
            umma_bf16(tmem, ad, bd, idesc, 1u);
            umma_bf16(tmem, ad2, bd2, idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long dt = clock64() - t0;
        if (blockIdx.x == 0) out[0] = dt;
        atomicMax(reinterpret_cast<unsigned long long*>(out + 3), (unsigned long long)dt);
        done = 1;
    }
    if (!concurrent) {
        __syncthreads();
        tc_fence_after();
    }
    if (warp >= 1 && warp <= ld_warps) {
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        float acc = 0.f;
        const long long t0 = clock64();
        long long n = 0;
        for (int k = 0; concurrent ? !done : k < reps; ++k, ++n) {
            if (ld_x16) {
                float v[16];
                tmem_ld16(tmem + lane_off + 256 + ((k * 16) & 255), v);
                tmem_ld_wait();
                acc += v[0] + v[15];
            } else {
                float v[32];
                tmem_ld32(tmem + lane_off + 256 + ((k * 32) & 255), v);
                tmem_ld_wait();
                acc += v[0] + v[31];
            }
        }
        const long long dt = clock64() - t0;
        if (tid == 32 && blockIdx.x == 0) { out[1] = dt; out[4] = n; }
        if (acc == 123.456f) out[2] = 1;      // keep the loads alive
    } else if (concurrent && warp >= 5 && warp < 5 + st_warps) {
        // 16-byte stores into a scratch area of shared memory (an im2col build's traffic) until the MMAs are done
        long long n = 0;
        if (gscratch != nullptr) {                        // ... or coalesced 16-byte stores to HBM (the conv epilogue's traffic: 1 MB per CTA, revisited)
            uint4* dst = gscratch + (size_t)blockIdx.x * 65536 + (tid & 127);
            for (int k = 0; !done; ++k, ++n) dst[(size_t)(k & 511) * 128] = make_uint4(k, k, k, k);
        } else {
            uint4* dst = reinterpret_cast<uint4*>(smem + 49152) + (tid & 127);
            for (int k = 0; !done; ++k, ++n) dst[(k & 7) * 128] = make_uint4(k, k, k, k);
        }
        if ((tid & 31) == 0 && warp == 5 && blockIdx.x == 0) out[5] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace bcad

extern "C" int bcad_selftest_umma_bench(const int32_t* p /*N,a_layout,b_layout,a_lbo,a_sbo,b_lbo,b_sbo,reps,ld_warps,ld_x16,a_off,alternate,grid,concurrent,st_warps,stores_to_hbm*/,
                                        long long* out_dev, void* stream) {
    using namespace bcad;
    BCAD_REQUIRE(p && out_dev, "selftest bench: null argument");
    BCAD_REQUIRE(p[12] >= 1 && p[12] <= 1024 && p[8] >= 0 && p[8] <= 4 && p[14] >= 0 && p[14] <= 3, "selftest bench: bad grid / warp counts");
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    BCAD_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, 6 * sizeof(long long), (cudaStream_t)stream));
    static uint4* gscratch = nullptr;                     // 1 MB per CTA for the stores-to-HBM variant (diagnostic tool: kept for the process)
    if (p[15] && gscratch == nullptr) BCAD_CUDA_CHECK(cudaMalloc((void**)&gscratch, (size_t)1024 * 65536 * sizeof(uint4)));
    umma_bench_kernel<<<p[12], 256, 64 * 1024, (cudaStream_t)stream>>>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[13], p[14],
                                                                        p[15] ? gscratch : nullptr, out_dev);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

extern "C" int bcad_selftest_umma(const void* a_img_dev, int a_bytes, const void* b_img_dev, int b_bytes,
                                  const int32_t* params_host, float* d_dev, void* stream) {
    using namespace bcad;
    BCAD_REQUIRE(a_img_dev && b_img_dev && params_host && d_dev, "selftest: null argument");
    BCAD_REQUIRE(a_bytes % 16 == 0 && b_bytes % 16 == 0, "selftest: images must be multiples of 16 bytes");
    SelftestParams p;
    memcpy(&p, params_host, sizeof(p));
    BCAD_REQUIRE(p.N >= 16 && p.N <= 256 && p.N % 16 == 0 && p.steps >= 1 && p.steps <= 64, "selftest: bad N/steps");
    BCAD_REQUIRE(p.mode == 0 || (p.mode == 1 && p.N % 32 == 0), "selftest: mode must be 0, or 1 with N a multiple of 32");
    const size_t smem = ((size_t)(a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
    BCAD_REQUIRE(smem <= 200 * 1024, "selftest: operand images too large");
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(a_img_dev), a_bytes,
                                                                 reinterpret_cast<const uint4*>(b_img_dev), b_bytes, p, d_dev);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}
