// Tensor-core kernels of the training step: the second conv block's forward, input gradient and weight gradient as
// split-operand ("x3") tcgen05 implicit GEMMs that read and write the fp32 NHWC tensors of the fp32 path directly.
//
// Every fp32 value v is split on the fly into hi = round16(v) and lo = round16(v - hi); a product is a_hi b_hi + a_lo b_hi + a_hi b_lo
// with fp32 accumulation in TMEM.  The forward splits into fp16 pairs (2^-22 relative); the two gradient kernels split BOTH operands into
// bf16 pairs (2^-16, but fp32's exponent range: no loss scaling; one MMA takes one 16-bit type -- mixing fp16 and bf16 operands raises
// "illegal instruction").  Reference: Classes/CNNModel.py:227-240 (conv forward), :320-355 (dX, dF of a conv block),
// ADCNNM.py:72-78 / autograd.
//
//   conv3x3_x3_kernel<CIN, COUT, ABF>   y[b, oy, ox, :] = act(sum_taps x[b, oy+dy-pad, ox+dx-pad, :] . w[tap] + bias)      (fwd: 32 -> 64, dgrad: 64 -> 32)
//   wgrad3x3_x3_kernel                   dW[tap][ci][co] = sum_{b, y, x} X[b, y+dy-pad, x+dx-pad, ci] dY[b, y, x, co]           (32 -> 64; pixels = K, MN-major operands)
//   dense_fwd_x3_kernel                  z[b][u] = sum_k p[b][k] W[u][k]                                                       (split-K over the CTAs)
//   dense_bwd_x3_kernel<0 / 1>           dW[u][k] = sum_b dz[b][u] p[b][k]   /   g[b][k] = sum_u dz[b][u] W[u][k]               (streamed fp32 tiles, MN-major)
//   CUDA-core helpers of the same step: conv0_fwd_fused / conv0_bwd_fused (first block, one input channel, lane = filter), unpool_mask,
//   maxpool2x2_nhwc, colsum64, pack_w_x3.
//
// Common structure (as the second block of conv_fused_kernel): a persistent CTA walks (image, row band) items; producer warps convert
// input rows into a shared-memory ring of "C8-planar" rows [plane hi|lo][channel octet][pixel slot][8 x 16 bit] -- a tap's column shift is a
// +16-byte descriptor offset, a row is loaded once per band --; one thread issues the MMAs; four epilogue warps drain TMEM.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdio.h>

#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100.cuh"
#include "sm100_train.h"

namespace bcad {

using namespace sm100;

namespace {

constexpr int TC_XP = 130;                    // pixel slots per ring row (128 + 2 halo): LBO = 2080 B keeps the producers' 16-byte stores conflict-free
constexpr int TC_PW = 8;                      // producer warps
constexpr int TC_THREADS = 32 * (TC_PW + 1 + 4);
constexpr int TC_NB = 4;                      // accumulator buffers (output rows in flight between the issuer and the epilogue)

__device__ __forceinline__ uint32_t split_hi_lo(float a, float b, uint32_t& lo, bool bf) {
    if (bf) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
        lo = *reinterpret_cast<const uint32_t*>(&l);
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    lo = *reinterpret_cast<const uint32_t*>(&l);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// instruction descriptor, kind::f16, fp32 accumulate; formats 0 = fp16, 1 = bf16; major bits 15 (A) / 16 (B): 1 = MN-major
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N, int a_fmt, int b_fmt, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CIN, int COUT>
struct TcSmem {
    static constexpr int CHUNKS = CIN / 8;
    static constexpr int LBO = TC_XP * 16;
    static constexpr int PLANEB = CHUNKS * LBO;
    static constexpr int ROWB = 2 * PLANEB;
    static constexpr int R = (CIN <= 32) ? 6 : 4;
    static constexpr int WPL = 9 * CHUNKS * COUT * 16;                 // one weight plane [tap][octet][COUT][16 B]
    static constexpr int OFF_W = 0;
    static constexpr int OFF_RING = 2 * WPL;
    static constexpr int OFF_BIAS = OFF_RING + R * ROWB;
    static constexpr int OFF_BAR = OFF_BIAS + COUT * 4;
    static constexpr int TOTAL = OFF_BAR + 256;
};

// row of `n8` items of 8 fp32 values -> hi / lo 16-byte items of a ring row (256 producer threads)
// 8 fp32 values -> one hi and one lo 16-byte item
__device__ __forceinline__ void split8(const float4& v0, const float4& v1, uint4& hi, uint4& lo, bool bf) {
    hi.x = split_hi_lo(v0.x, v0.y, lo.x, bf);
    hi.y = split_hi_lo(v0.z, v0.w, lo.y, bf);
    hi.z = split_hi_lo(v1.x, v1.y, lo.z, bf);
    hi.w = split_hi_lo(v1.z, v1.w, lo.w, bf);
}

constexpr int TC_U = 4;                       // producer items in flight per thread: all loads of a group are issued before the first conversion

// row of W * CIN fp32 values -> hi / lo 16-byte items of a ring row (256 producer threads)
template <int CIN>
__device__ __forceinline__ void produce_row(const float* __restrict__ src, int W, uint8_t* row, int planeb, int pad, int ptid, bool bf) {
    constexpr int CHUNKS = CIN / 8;
    const int n = W * CHUNKS;
    for (int it0 = ptid; it0 < n; it0 += 32 * TC_PW * TC_U) {
        float4 v0[TC_U], v1[TC_U];
#pragma unroll
        for (int u = 0; u < TC_U; ++u) {
            const int it = it0 + u * 32 * TC_PW;
            if (it < n) {
                v0[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)it * 8));           // item (px, ch) = 8 consecutive floats of the row
                v1[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)it * 8 + 4));
            }
        }
#pragma unroll
        for (int u = 0; u < TC_U; ++u) {
            const int it = it0 + u * 32 * TC_PW;
            if (it < n) {
                const int px = it / CHUNKS, ch = it % CHUNKS;
                uint4 hi, lo;
                split8(v0[u], v1[u], hi, lo, bf);
                uint8_t* dst = row + ch * (TC_XP * 16) + (px + pad) * 16;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + planeb) = lo;
            }
        }
    }
}

template <int CIN, int COUT, bool ABF>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_x3_kernel(TcConvArgs a) {
    using L = TcSmem<CIN, COUT>;
    constexpr int R = L::R, NB = TC_NB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem + L::OFF_W;
    uint8_t* s_ring = smem + L::OFF_RING;
    float* s_bias = reinterpret_cast<float*>(smem + L::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                 // [R] producers -> MMA (count TC_PW)
    uint64_t* empty = bars + R;            // [R] MMA -> producers
    uint64_t* tfull = bars + 2 * R;        // [NB] MMA -> epilogue
    uint64_t* tempty = bars + 2 * R + NB;  // [NB] epilogue -> MMA (count 4)
    uint64_t* wbar = bars + 2 * R + 2 * NB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 2 * NB + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (R * L::ROWB) / 16; i += TC_THREADS) reinterpret_cast<uint4*>(s_ring)[i] = make_uint4(0, 0, 0, 0);   // halo slots stay zero
    if (tid < COUT) s_bias[tid] = a.bias ? a.bias[tid] : 0.f;
    if (tid == 0) {
        for (int i = 0; i < R; ++i) { mbar_init(&full[i], TC_PW); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    constexpr uint32_t TCOLS = (NB * COUT <= 128) ? 128 : 256;
    if (warp == 0) {
        tmem_alloc(tmem_slot, TCOLS);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_items = a.B * a.bands;

    if (warp < TC_PW) {
        // ================================ producers: fp32 NHWC rows -> hi / lo ring rows ================================
        uint32_t g = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / a.bands, y0 = (item % a.bands) * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            for (int i = 0; i < nrows + 2; ++i, ++g) {
                const uint32_t slot = g % R;
                if (g >= (uint32_t)R) mbar_wait(&empty[slot], ((g / R) - 1) & 1);
                const int in_row = y0 - a.pad + i;
                if (in_row >= 0 && in_row < a.H) {
                    produce_row<CIN>(a.x + ((size_t)b * a.H + in_row) * a.W * CIN, a.W, s_ring + slot * L::ROWB, L::PLANEB, a.pad, tid, ABF);
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[slot]);
            }
        }
    } else if (warp == TC_PW) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr int WB = 2 * L::WPL;
            mbar_arrive_expect_tx(wbar, WB);
            for (int off = 0; off < WB; off += 16384) bulk_g2s(s_w + off, a.w_img + off, min(16384, WB - off), wbar);
        }
        const bool leader = elect_one();
        constexpr uint32_t idesc = tc_idesc(128, COUT, ABF ? 1 : 0, ABF ? 1 : 0, 0, 0);     // kind::f16 wants A and B of ONE 16-bit type
        constexpr uint32_t d_hi = (uint32_t)(128 >> 4) | (1u << 14);                 // SBO 128, descriptor version 1
        constexpr uint32_t a_lo_t = (uint32_t)(L::LBO >> 4) << 16;
        constexpr uint32_t b_lo_t = (uint32_t)((COUT * 16) >> 4) << 16;
        const uint32_t w_base = smem_u32(s_w), ring_base = smem_u32(s_ring);
        mbar_wait(wbar, 0);
        uint32_t gbase = 0, acc_it = 0;
        auto wait_full = [&](uint32_t gs) { mbar_wait(&full[gs % R], (gs / R) & 1u); };
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int y0 = (item % a.bands) * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            for (int r = 0; r < nrows; ++r, ++acc_it) {
                if (r == 0) { wait_full(gbase); wait_full(gbase + 1); }
                wait_full(gbase + r + 2);
                const uint32_t j = acc_it % NB;
                if (acc_it >= (uint32_t)NB) mbar_wait(&tempty[j], ((acc_it / NB) - 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + j * COUT;
                uint32_t acc = 0;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int in_row = y0 - a.pad + r + dy;
                    if (in_row < 0 || in_row >= a.H) continue;                      // padding row: contributes nothing
                    const uint32_t row_base = ring_base + ((gbase + (uint32_t)(r + dy)) % R) * L::ROWB;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                        for (int ks = 0; ks < CIN / 16; ++ks)
#pragma unroll
                            for (int combo = 0; combo < 3; ++combo) {           // a_hi w_hi, a_lo w_hi, a_hi w_lo
                                const uint32_t a_addr = row_base + (combo == 1 ? L::PLANEB : 0) + ks * 2 * L::LBO + dx * 16;
                                const uint32_t b_addr = w_base + (combo == 2 ? L::WPL : 0) + ((dy * 3 + dx) * L::CHUNKS + 2 * ks) * (COUT * 16);
                                umma_f16_if(leader, d_tmem, desc64(a_lo_t | ((a_addr & 0x3FFFFu) >> 4), d_hi), desc64(b_lo_t | ((b_addr & 0x3FFFFu) >> 4), d_hi),
                                            idesc, acc);
                                acc = 1u;
                            }
                }
                umma_commit_if(leader, &empty[(gbase + (uint32_t)r) % R]);
                umma_commit_if(leader, &tfull[j]);
            }
            umma_commit_if(leader, &empty[(gbase + (uint32_t)nrows) % R]);
            umma_commit_if(leader, &empty[(gbase + (uint32_t)nrows + 1) % R]);
            gbase += (uint32_t)nrows + 2;
        }
    } else {
        // ================================ epilogue: TMEM -> (+ bias, LeakyReLU) -> fp32 NHWC ================================
        const int quad = warp & 3;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int x = quad * 32 + lane;
        uint32_t acc_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / a.bands, y0 = (item % a.bands) * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            for (int r = 0; r < nrows; ++r, ++acc_it) {
                const uint32_t j = acc_it % NB;
                mbar_wait(&tfull[j], (acc_it / NB) & 1);
                tc_fence_after();
                float* dst = a.y + (((size_t)b * a.Ho + (y0 + r)) * a.Wo + x) * COUT;
#pragma unroll
                for (int half = 0; half < COUT / 32; ++half) {
                    float v[32];
                    tmem_ld32(tmem + lane_off + j * COUT + half * 32, v);
                    tmem_ld_wait();
                    if (half == COUT / 32 - 1) {                    // the row is in registers: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[j]);
                    }
                    if (x < a.Wo) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 o;
                            o.x = v[4 * q] + s_bias[half * 32 + 4 * q];
                            o.y = v[4 * q + 1] + s_bias[half * 32 + 4 * q + 1];
                            o.z = v[4 * q + 2] + s_bias[half * 32 + 4 * q + 2];
                            o.w = v[4 * q + 3] + s_bias[half * 32 + 4 * q + 3];
                            o.x = o.x > 0.f ? o.x : a.alpha * o.x;
                            o.y = o.y > 0.f ? o.y : a.alpha * o.y;
                            o.z = o.z > 0.f ? o.z : a.alpha * o.z;
                            o.w = o.w > 0.f ? o.w : a.alpha * o.w;
                            *reinterpret_cast<float4*>(dst + half * 32 + 4 * q) = o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// fp32 [9][CIN][CoutPad] (the fp32 path's packed filters) -> fp16 hi / lo planes [plane][tap][CIN/8][COUT][8]
__global__ void pack_w_x3_kernel(const float* __restrict__ w, uint8_t* __restrict__ img, int CIN, int COUT, int CoutPad, int bf) {
    const int n = 9 * CIN * COUT;
    const size_t wpl = (size_t)9 * (CIN / 8) * COUT * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int co = i % COUT, ci = (i / COUT) % CIN, tap = i / (COUT * CIN);
        const float v = w[((size_t)tap * CIN + ci) * CoutPad + co];
        const size_t off = (((size_t)tap * (CIN / 8) + ci / 8) * COUT + co) * 16 + (ci % 8) * 2;
        if (bf) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img + wpl + off) = __float2bfloat16_rn(v - __bfloat162float(hi));
        } else {
            const __half hi = __float2half_rn(v);
            *reinterpret_cast<__half*>(img + off) = hi;
            *reinterpret_cast<__half*>(img + wpl + off) = __float2half_rn(v - __half2float(hi));
        }
    }
}

// 2x2 / 2 max pool of an NHWC fp32 map (floor: Classes/CNNModel.py:245-261, nn.MaxPool2d(2)); C % 4 == 0
__global__ void maxpool2x2_nhwc_kernel(const float* __restrict__ y, float* __restrict__ p, int Ho, int Wo, int Hp, int Wp, int C4, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        size_t t = i / C4;
        const int px = (int)(t % Wp);
        t /= Wp;
        const int py = (int)(t % Hp);
        const size_t b = t / Hp;
        const float4* r0 = reinterpret_cast<const float4*>(y) + ((b * Ho + 2 * py) * Wo + 2 * px) * C4 + c;
        const float4* r1 = r0 + (size_t)Wo * C4;
        const float4 a0 = r0[0], a1 = r0[C4], a2 = r1[0], a3 = r1[C4];
        float4 o;
        o.x = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x));
        o.y = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
        o.z = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z));
        o.w = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
        reinterpret_cast<float4*>(p)[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 conv block, 32 -> 64 channels:  dW[dy][dx][ci][co] = sum_{b, y, x} X[b, y+dy-pad, x+dx-pad, ci] dY[b, y, x, co].
// The pixel index is the GEMM's K: both operands are read MN-major (8 channels = one 16-byte item, consecutive pixels 16 bytes apart --
// exactly the C8-planar rows above), so a tap's column shift is again a descriptor offset.  A work unit is (image, output row pair, 64-pixel
// half): the A operand stacks the two dY rows (M = 2 x 64 filters), the B operand is one of the four X rows y'-pad .. y'-pad+3 (N = 32).
// X row k feeds tap dy = k of the first dY row (accumulator lanes 0-63) and tap dy = k - 1 of the second (lanes 64-127), so twelve 32-column
// accumulators T[dx][k] collect everything; they stay in TMEM for the CTA's whole life and are dumped once (partials[cta][dx][k][lane][ci]).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int WG_STAGES = 3;
constexpr int WG_XS = 66;                                       // X pixel slots per half row (64 + 2 halo)
constexpr int WG_XROWB = 2 * 4 * WG_XS * 16;                    // [plane][4 octets][66][16 B]
constexpr int WG_XPLANE = 4 * WG_XS * 16;
constexpr int WG_YPLANE = 2 * 8 * 64 * 16;                      // [row r][8 octets][64 px][16 B]
constexpr int WG_YB = 2 * WG_YPLANE;
constexpr int WG_STAGEB = 4 * WG_XROWB + WG_YB;
constexpr int WG_OFF_BAR = WG_STAGES * WG_STAGEB;
constexpr int WG_TOTAL = WG_OFF_BAR + 128;
constexpr int WG_THREADS = 32 * (TC_PW + 1 + 4);

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad3x3_x3_kernel(TcWgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_OFF_BAR);
    uint64_t* full = bars;                       // [S] producers -> MMA (count TC_PW)
    uint64_t* empty = bars + WG_STAGES;          // [S] MMA -> producers
    uint64_t* done = bars + 2 * WG_STAGES;       // all MMAs retired -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], TC_PW); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int npairs = (a.Ho + 1) / 2, nhalf = (a.Wo + 63) / 64;
    const int n_units = a.B * npairs * nhalf;

    if (warp < TC_PW) {
        // ================================ producers ================================
        uint32_t g = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++g) {
            const int h = u % nhalf, p = (u / nhalf) % npairs, b = u / (nhalf * npairs);
            const int x0 = h * 64;
            const uint32_t st = g % WG_STAGES;
            if (g >= (uint32_t)WG_STAGES) mbar_wait(&empty[st], ((g / WG_STAGES) - 1) & 1);
            uint8_t* sx = smem + st * WG_STAGEB;
            uint8_t* sy = sx + 4 * WG_XROWB;
            // X rows (bf16 pairs, like dY: one MMA takes one 16-bit type): slot s = pixel x0 + s - pad, zero outside the map
            for (int it0 = tid; it0 < 4 * WG_XS * 4; it0 += 32 * TC_PW * TC_U) {
                float4 v0[TC_U], v1[TC_U];
                bool ok[TC_U];
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    const int ch = it & 3, sl = (it >> 2) % WG_XS, k = it / (4 * WG_XS);
                    const int in_row = 2 * p - a.pad + k, px = x0 + sl - a.pad;
                    ok[u] = it < 4 * WG_XS * 4 && in_row >= 0 && in_row < a.H && px >= 0 && px < a.W;
                    if (ok[u]) {
                        const float* src = a.x + (((size_t)b * a.H + in_row) * a.W + px) * 32 + ch * 8;
                        v0[u] = __ldg(reinterpret_cast<const float4*>(src));
                        v1[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    if (it >= 4 * WG_XS * 4) continue;
                    const int ch = it & 3, sl = (it >> 2) % WG_XS, k = it / (4 * WG_XS);
                    const int in_row = 2 * p - a.pad + k;
                    if (in_row < 0 || in_row >= a.H) continue;                   // the issuer skips this row
                    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                    if (ok[u]) split8(v0[u], v1[u], hi, lo, true);
                    uint8_t* dst = sx + k * WG_XROWB + ch * (WG_XS * 16) + sl * 16;
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + WG_XPLANE) = lo;
                }
            }
            // dY rows (bf16 pairs): slot s = pixel x0 + s, zero outside the map / below the last row
            for (int it0 = tid; it0 < 2 * 64 * 8; it0 += 32 * TC_PW * TC_U) {
                float4 v0[TC_U], v1[TC_U];
                bool ok[TC_U];
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    const int ch = it & 7, sl = (it >> 3) & 63, r = it >> 9;
                    const int y = 2 * p + r, px = x0 + sl;
                    ok[u] = y < a.Ho && px < a.Wo;
                    if (ok[u]) {
                        const float* src = a.dy + (((size_t)b * a.Ho + y) * a.Wo + px) * 64 + ch * 8;
                        v0[u] = __ldg(reinterpret_cast<const float4*>(src));
                        v1[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    const int ch = it & 7, sl = (it >> 3) & 63, r = it >> 9;
                    uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                    if (ok[u]) split8(v0[u], v1[u], hi, lo, true);
                    uint8_t* dst = sy + (r * 8 + ch) * (64 * 16) + sl * 16;
                    *reinterpret_cast<uint4*>(dst) = hi;
                    *reinterpret_cast<uint4*>(dst + WG_YPLANE) = lo;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
        }
    } else if (warp == TC_PW) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        constexpr uint32_t idesc = tc_idesc(128, 32, 1, 1, 1, 1);              // A = dY, B = X: bf16, both MN-major
        // MN-major, no swizzle: LBO = byte distance between 8-pixel (K) groups = 128, SBO = byte distance between channel octets
        constexpr uint32_t a_hi = (uint32_t)((64 * 16) >> 4) | (1u << 14);
        constexpr uint32_t b_hi = (uint32_t)((WG_XS * 16) >> 4) | (1u << 14);
        constexpr uint32_t lbo_t = (uint32_t)(128 >> 4) << 16;
        uint32_t started = 0;                                                    // bit dx * 4 + k: accumulator T[dx][k] has been written
        uint32_t g = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++g) {
            const int p = (u / nhalf) % npairs;
            const uint32_t st = g % WG_STAGES;
            mbar_wait(&full[st], (g / WG_STAGES) & 1);
            tc_fence_after();
            const uint32_t sx = smem_u32(smem + st * WG_STAGEB), sy = sx + 4 * WG_XROWB;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int in_row = 2 * p - a.pad + k;
                if (in_row < 0 || in_row >= a.H) continue;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t bit = 1u << (dx * 4 + k);
                    uint32_t acc = (started & bit) ? 1u : 0u;
                    started |= bit;
                    const uint32_t d_tmem = tmem + (dx * 4 + k) * 32;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int combo = 0; combo < 3; ++combo) {               // dY_hi X_hi, dY_lo X_hi, dY_hi X_lo
                            const uint32_t a_addr = sy + (combo == 1 ? WG_YPLANE : 0) + ks * 256;
                            const uint32_t b_addr = sx + k * WG_XROWB + (combo == 2 ? WG_XPLANE : 0) + (ks * 16 + dx) * 16;
                            umma_f16_if(leader, d_tmem, desc64(lbo_t | ((a_addr & 0x3FFFFu) >> 4), a_hi), desc64(lbo_t | ((b_addr & 0x3FFFFu) >> 4), b_hi), idesc, acc);
                            acc = 1u;
                        }
                }
            }
            umma_commit_if(leader, &empty[st]);
        }
        // accumulators that never received an MMA hold garbage: tell the epilogue which ones are live (before the commit it waits on)
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1) = started;
        __threadfence_block();
        __syncwarp();
        umma_commit_if(leader, done);
    } else {
        // ================================ epilogue: dump the accumulators once ================================
        const int quad = warp & 3;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        __syncwarp();
        const uint32_t started = *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1);
        float* dst = a.partials + (size_t)blockIdx.x * (12 * 128 * 32);
#pragma unroll 1
        for (int t = 0; t < 12; ++t) {
            float v[32];
            tmem_ld32(tmem + lane_off + t * 32, v);
            tmem_ld_wait();
            const bool live = (started >> t) & 1u;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(dst + ((size_t)t * 128 + row) * 32 + 4 * q) =
                    live ? make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// dW[(dy, dx)][ci][co] = sum over CTAs of T[dx][dy][lane co][ci] + T[dx][dy + 1][lane 64 + co][ci], into the fp32 path's [9][32][CoutPad] layout
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw, int CoutPad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // over 9 * 32 * 64
    if (i >= 9 * 32 * 64) return;
    const int co = i & 63, ci = (i >> 6) & 31, tap = i >> 11;
    const int dy = tap / 3, dx = tap % 3;
    const size_t o0 = ((size_t)(dx * 4 + dy) * 128 + co) * 32 + ci, o1 = ((size_t)(dx * 4 + dy + 1) * 128 + 64 + co) * 32 + ci;
    float s0 = 0.f, s1 = 0.f;
    for (int c = 0; c < nparts; ++c) {
        s0 += part[(size_t)c * (12 * 128 * 32) + o0];
        s1 += part[(size_t)c * (12 * 128 * 32) + o1];
    }
    dw[((size_t)tap * 32 + ci) * CoutPad + co] = s0 + s1;
}

// out[n] = sum_k A[k][64] over a tall matrix (bias gradient of the 64-filter block): fixed slabs of rows per CTA, fixed-order second pass
__global__ void __launch_bounds__(256) colsum64_part_kernel(const float* __restrict__ A, float* __restrict__ part, size_t K, int slab) {
    __shared__ float sm[4][64];
    const int n = threadIdx.x & 63, q = threadIdx.x >> 6;
    const size_t k0 = (size_t)blockIdx.x * slab, k1 = k0 + slab < K ? k0 + slab : K;
    float acc = 0.f;
    for (size_t k = k0 + q; k < k1; k += 4) acc += A[k * 64 + n];
    sm[q][n] = acc;
    __syncthreads();
    if (q == 0) part[(size_t)blockIdx.x * 64 + n] = (sm[0][n] + sm[1][n]) + (sm[2][n] + sm[3][n]);
}
__global__ void colsum64_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ out) {
    const int n = threadIdx.x;
    float acc = 0.f;
    for (int c = 0; c < nparts; ++c) acc += part[(size_t)c * 64 + n];
    out[n] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// CUDA-core pieces of the fast backward.
// unpool_mask: dz = (pool switch ? g : 0) * LeakyReLU'(y) in ONE pass over y (Classes/CNNModel.py:263-280 + :300-304; the two-kernel
// form reads and writes the full-resolution gradient twice).  Same arithmetic, same results.
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpool_mask_kernel(const float4* __restrict__ g, const float4* __restrict__ y, float4* __restrict__ dz,
                                                          int B, int Ho, int Wo, int C4, int first_only, float alpha) {
    const int Hp = Ho / 2, Wp = Wo / 2, Hc = (Ho + 1) / 2, Wc = (Wo + 1) / 2;
    const size_t total = (size_t)B * Hc * Wc * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        size_t r = i / C4;
        const int wx = (int)(r % Wc); r /= Wc;
        const int wy = (int)(r % Hc);
        const size_t b = r / Hc;
        const size_t ybase = (b * Ho) * Wo * C4 + c;
        if (wy < Hp && wx < Wp) {
            const float4 gv = g[((b * Hp + wy) * Wp + wx) * C4 + c];
            float4 v[4];
            size_t idx[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                idx[q] = ybase + ((size_t)(2 * wy + (q >> 1)) * Wo + (2 * wx + (q & 1))) * C4;
                v[q] = y[idx[q]];
            }
            float4 o[4];
            auto lane = [&](float g1, float a0, float a1, float a2, float a3, float& o0, float& o1, float& o2, float& o3) {
                const float m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
                bool s0 = (a0 == m), s1 = (a1 == m), s2 = (a2 == m), s3 = (a3 == m);
                if (first_only) { s1 = s1 && !s0; s2 = s2 && !(s0 || s1); s3 = s3 && !(s0 || s1 || s2); }
                o0 = (s0 ? g1 : 0.f) * (a0 > 0.f ? 1.f : alpha);
                o1 = (s1 ? g1 : 0.f) * (a1 > 0.f ? 1.f : alpha);
                o2 = (s2 ? g1 : 0.f) * (a2 > 0.f ? 1.f : alpha);
                o3 = (s3 ? g1 : 0.f) * (a3 > 0.f ? 1.f : alpha);
            };
            lane(gv.x, v[0].x, v[1].x, v[2].x, v[3].x, o[0].x, o[1].x, o[2].x, o[3].x);
            lane(gv.y, v[0].y, v[1].y, v[2].y, v[3].y, o[0].y, o[1].y, o[2].y, o[3].y);
            lane(gv.z, v[0].z, v[1].z, v[2].z, v[3].z, o[0].z, o[1].z, o[2].z, o[3].z);
            lane(gv.w, v[0].w, v[1].w, v[2].w, v[3].w, o[0].w, o[1].w, o[2].w, o[3].w);
#pragma unroll
            for (int q = 0; q < 4; ++q) dz[idx[q]] = o[q];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int yy = 2 * wy + (q >> 1), xx = 2 * wx + (q & 1);
                if (yy < Ho && xx < Wo) dz[ybase + ((size_t)yy * Wo + xx) * C4] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// First conv block (one input channel, 32 filters, 3x3): max-pool backward, LeakyReLU' and the weight / bias gradients in ONE pass --
// the 32-channel full-resolution gradient (0.5 GB per 64 images at 256x256) is never written: a warp (lane = filter) walks the pool
// windows of a row pair, routes g to the window's maximum, and accumulates dF[tap][lane] += dz * x[tap] from a 4x4 patch of the input row
// strip staged in shared memory.  Per-CTA partials, fixed-order final sum (deterministic).
// dF[ky][kx][f] = sum dz[b,y,x,f] x[b, y+ky-pad, x+kx-pad]   (explainability.py:58-59 / Classes/CNNModel.py:340-352 for C = 1), db[f] = sum dz.
// ---------------------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------------------
// First conv block forward for training (one input channel, 32 filters, 3x3): y = LeakyReLU(conv + b) and its 2x2 max pool in one pass,
// lane = filter (a pixel's 32 outputs are one coalesced 128-byte store), the 4x4 input patch of a pool window broadcast from a shared-memory
// row strip, the 9 weights + bias of the lane's filter in registers.  Same structure as conv0_bwd_fused_kernel below.
// Reference: Classes/CNNModel.py:227-261 for C = 1 / nn.Conv2d(1, 32, 3) + LeakyReLU + MaxPool2d(2) (ADCNNM.py:48-53).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int C0F_WARPS = 8;
__global__ void __launch_bounds__(32 * C0F_WARPS) conv0_fwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                                        float* __restrict__ y, float* __restrict__ p, int B, int H, int W, int Ho, int Wo,
                                                                        int pad, float alpha) {
    extern __shared__ float s_x[];                              // [4 rows][pitch]: column j = input column j - pad, zero outside the image
    const int Hp = Ho / 2, Wp = Wo / 2, Hc = (Ho + 1) / 2, Wc = (Wo + 1) / 2;      // windows incl. the odd last row / column (computed, not pooled)
    const int pitch = (W + 2 * pad + 4 + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wt[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg(w + t * 32 + lane);              // packed [tap][Cin = 1][CoutPad = 32]
    const float bl = __ldg(bias + lane);
    const int units = B * Hc;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = u / Hc, wy = u % Hc;
        __syncthreads();
        for (int i = threadIdx.x; i < 4 * pitch; i += blockDim.x) {
            const int r = i / pitch, j = i % pitch;
            const int row = 2 * wy - pad + r, col = j - pad;
            s_x[i] = (row >= 0 && row < H && col >= 0 && col < W) ? __ldg(x + ((size_t)b * H + row) * W + col) : 0.f;
        }
        __syncthreads();
        const bool row1 = 2 * wy + 1 < Ho;
        float* y0 = y + (((size_t)b * Ho + 2 * wy) * Wo) * 32 + lane;
        float* y1 = y0 + (size_t)Wo * 32;
        for (int wx = warp; wx < Wc; wx += C0F_WARPS) {
            float pt[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 lo = *reinterpret_cast<const float2*>(s_x + r * pitch + 2 * wx);
                const float2 hi = *reinterpret_cast<const float2*>(s_x + r * pitch + 2 * wx + 2);
                pt[r][0] = lo.x; pt[r][1] = lo.y; pt[r][2] = hi.x; pt[r][3] = hi.y;
            }
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int qy = q >> 1, qx = q & 1;
                float acc = 0.f;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(pt[qy + ky][qx + kx], wt[ky * 3 + kx], acc);
                acc += bl;
                o[q] = acc > 0.f ? acc : alpha * acc;
            }
            const bool col1 = 2 * wx + 1 < Wo;
            y0[(size_t)(2 * wx) * 32] = o[0];
            if (col1) y0[(size_t)(2 * wx + 1) * 32] = o[1];
            if (row1) {
                y1[(size_t)(2 * wx) * 32] = o[2];
                if (col1) y1[(size_t)(2 * wx + 1) * 32] = o[3];
            }
            if (wy < Hp && wx < Wp) p[(((size_t)b * Hp + wy) * Wp + wx) * 32 + lane] = fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3]));
        }
    }
}

constexpr int C0B_WARPS = 8;
__global__ void __launch_bounds__(32 * C0B_WARPS) conv0_bwd_fused_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ x,
                                                                        float* __restrict__ part, int B, int H, int W, int Ho, int Wo, int pad,
                                                                        int first_only, float alpha) {
    extern __shared__ float s_x[];                              // [4 rows][pitch]: column j = input column j - pad, zero outside the image
    __shared__ float s_red[C0B_WARPS][10][32];
    const int Hp = Ho / 2, Wp = Wo / 2;
    const int pitch = (W + 2 * pad + 4 + 1) & ~1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[9], accb = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    const int units = B * Hp;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = u / Hp, wy = u % Hp;
        __syncthreads();                                        // the previous unit's patch reads are done
        for (int i = threadIdx.x; i < 4 * pitch; i += blockDim.x) {
            const int r = i / pitch, j = i % pitch;
            const int row = 2 * wy - pad + r, col = j - pad;
            s_x[i] = (row >= 0 && row < H && col >= 0 && col < W) ? __ldg(x + ((size_t)b * H + row) * W + col) : 0.f;
        }
        __syncthreads();
        const float* y0 = y + (((size_t)b * Ho + 2 * wy) * Wo) * 32 + lane;
        const float* y1 = y0 + (size_t)Wo * 32;
        const float* gr = g + (((size_t)b * Hp + wy) * Wp) * 32 + lane;
        for (int wx = warp; wx < Wp; wx += C0B_WARPS) {
            const float gv = __ldg(gr + (size_t)wx * 32);
            const float a0 = __ldg(y0 + (size_t)(2 * wx) * 32), a1 = __ldg(y0 + (size_t)(2 * wx + 1) * 32);
            const float a2 = __ldg(y1 + (size_t)(2 * wx) * 32), a3 = __ldg(y1 + (size_t)(2 * wx + 1) * 32);
            const float m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
            bool s0 = (a0 == m), s1 = (a1 == m), s2 = (a2 == m), s3 = (a3 == m);
            if (first_only) { s1 = s1 && !s0; s2 = s2 && !(s0 || s1); s3 = s3 && !(s0 || s1 || s2); }
            float d[4];
            d[0] = (s0 ? gv : 0.f) * (a0 > 0.f ? 1.f : alpha);
            d[1] = (s1 ? gv : 0.f) * (a1 > 0.f ? 1.f : alpha);
            d[2] = (s2 ? gv : 0.f) * (a2 > 0.f ? 1.f : alpha);
            d[3] = (s3 ? gv : 0.f) * (a3 > 0.f ? 1.f : alpha);
            accb += (d[0] + d[1]) + (d[2] + d[3]);
            float p[4][4];                                      // input patch: rows 2 wy - pad + r, columns 2 wx - pad + c (broadcast reads)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 lo = *reinterpret_cast<const float2*>(s_x + r * pitch + 2 * wx);
                const float2 hi = *reinterpret_cast<const float2*>(s_x + r * pitch + 2 * wx + 2);
                p[r][0] = lo.x; p[r][1] = lo.y; p[r][2] = hi.x; p[r][3] = hi.y;
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    float t = acc[ky * 3 + kx];
                    t = fmaf(d[0], p[ky][kx], t);
                    t = fmaf(d[1], p[ky][kx + 1], t);
                    t = fmaf(d[2], p[ky + 1][kx], t);
                    t = fmaf(d[3], p[ky + 1][kx + 1], t);
                    acc[ky * 3 + kx] = t;
                }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) s_red[warp][t][lane] = acc[t];
    s_red[warp][9][lane] = accb;
    __syncthreads();
    for (int i = threadIdx.x; i < 320; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < C0B_WARPS; ++w) sum += s_red[w][i / 32][i % 32];
        part[(size_t)blockIdx.x * 320 + i] = sum;
    }
}
__global__ void conv0_bwd_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ dw, float* __restrict__ db) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= 320) return;
    float sum = 0.f;
    for (int c = 0; c < nparts; ++c) sum += part[(size_t)c * 320 + i];
    if (i < 288) dw[i] = sum;
    else db[i - 288] = sum;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// The first dense layer's two backward GEMMs against the big (units x flat, 268 MB at the canonical shape) weight matrix:
//   MODE 0 (weight gradient)  dW[u][k] = sum_b dz[b][u] p[b][k]      M = units (halves of 128), N = 128 columns per tile, K = batch (<= 64)
//   MODE 1 (input gradient)   g[b][k]  = sum_u dz[b][u] W[u][k]      M = 128 columns per tile, N = batch (64), K = units in four stages of 64
// Both stream fp32 tiles of 64 rows x 128 contiguous columns (rows of p, resp. of W) that enter the MMA MN-major: the producers split them
// into bf16 hi / lo 16-byte items (8 columns of one row) at [column octet][row / 8][row % 8] -- the no-swizzle MN-major canonical layout
// (LBO = 128 B between 8-row groups, SBO = the octet pitch, padded to 1040 B so the producers' stores do not collide).  dz is resident.
// Reference: Classes/CNNModel.py:307-318 (dW = d_out^T . input, d_input = d_out . W), torch autograd of nn.Linear (ADCNNM.py:57-66).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int DG_NST = 4;                               // streamed stages
constexpr int DG_OCT = 1040;                            // pitch of a column octet: 8 row groups x 128 B, + 16 B
constexpr int DG_SPLANE = 16 * DG_OCT;                  // one plane of a streamed tile (128 columns)
constexpr int DG_STAGEB = 2 * DG_SPLANE;
constexpr int DG_RPLANE = 32 * DG_OCT;                  // resident dz plane: 32 octets (256 units MN-major, or 256 units as 32 K-chunks K-major)
constexpr int DG_OFF_R = DG_NST * DG_STAGEB;
constexpr int DG_OFF_STG = DG_OFF_R + 2 * DG_RPLANE;     // 4 epilogue warps x [32 rows][33] floats: accumulator rows -> row-contiguous stores
constexpr int DG_OFF_BAR = DG_OFF_STG + 4 * 32 * 33 * 4;
constexpr int DG_TOTAL = DG_OFF_BAR + 256;
constexpr int DG_THREADS = 32 * (TC_PW + 1 + 4);

template <int MODE>
__global__ void __launch_bounds__(DG_THREADS, 1) dense_bwd_x3_kernel(DenseBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_res = smem + DG_OFF_R;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DG_OFF_BAR);
    uint64_t* full = bars;                       // [NST] producers -> MMA (count TC_PW)
    uint64_t* empty = bars + DG_NST;             // [NST] MMA -> producers
    uint64_t* tfull = bars + 2 * DG_NST;         // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * DG_NST + 2;    // [2] epilogue -> MMA (count 4)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DG_NST + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t TCOLS = (MODE == 0) ? 512 : 128;
    if (tid == 0) {
        for (int i = 0; i < DG_NST; ++i) { mbar_init(&full[i], TC_PW); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, TCOLS);
        tmem_relinquish();
    }
    // resident dz (bf16 hi / lo): 16-byte item = 8 consecutive units of one image, MODE 0: MN-major [unit octet][b / 8][b % 8],
    // MODE 1: K-major [unit octet = K chunk][b]; rows b >= B and units >= a.units are zero
    for (int it = tid; it < 32 * 64; it += DG_THREADS) {
        const int oct = it & 31, b = it >> 5;
        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
        if (b < a.B && oct * 8 < a.units) {
            const float* src = a.dz + (size_t)b * a.units + oct * 8;
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            hi.x = split_hi_lo(v0.x, v0.y, lo.x, true);
            hi.y = split_hi_lo(v0.z, v0.w, lo.y, true);
            hi.z = split_hi_lo(v1.x, v1.y, lo.z, true);
            hi.w = split_hi_lo(v1.z, v1.w, lo.w, true);
        }
        uint8_t* dst = s_res + oct * DG_OCT + ((MODE == 0) ? ((b >> 3) * 128 + (b & 7) * 16) : (b * 16));
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + DG_RPLANE) = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_tiles = (int)(a.flat / 128);
    constexpr int SPT = (MODE == 0) ? 1 : 4;                     // streamed stages per output tile
    const int kq = (MODE == 0) ? 1 : a.units / 64;              // ... of which carry data (MODE 1: units / 64 row groups of W)

    if (warp < TC_PW) {
        // ================================ producers ================================
        uint32_t g = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int q = 0; q < SPT; ++q) {
                if (MODE == 1 && q >= kq) break;
                const uint32_t st = g % DG_NST;
                if (g >= (uint32_t)DG_NST) mbar_wait(&empty[st], ((g / DG_NST) - 1) & 1);
                uint8_t* dstb = smem + st * DG_STAGEB;
                const int nrows = (MODE == 0) ? a.B : 64;
                const float* base = a.src + (size_t)(MODE == 0 ? 0 : q * 64) * a.flat + (size_t)t * 128;
                for (int it0 = tid; it0 < 16 * 64; it0 += 32 * TC_PW * TC_U) {          // all loads of a group first: DRAM latency paid once
                    float4 v0[TC_U], v1[TC_U];
#pragma unroll
                    for (int u = 0; u < TC_U; ++u) {
                        const int it = it0 + u * 32 * TC_PW;
                        const int oct = it & 15, k = it >> 4;
                        if (k < nrows) {
                            const float* src = base + (size_t)k * a.flat + oct * 8;
                            v0[u] = __ldg(reinterpret_cast<const float4*>(src));
                            v1[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TC_U; ++u) {
                        const int it = it0 + u * 32 * TC_PW;
                        const int oct = it & 15, k = it >> 4;
                        uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                        if (k < nrows) split8(v0[u], v1[u], hi, lo, true);
                        uint8_t* dst = dstb + oct * DG_OCT + (k >> 3) * 128 + (k & 7) * 16;
                        *reinterpret_cast<uint4*>(dst) = hi;
                        *reinterpret_cast<uint4*>(dst + DG_SPLANE) = lo;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[st]);
                ++g;
            }
    } else if (warp == TC_PW) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        const uint32_t res = smem_u32(s_res);
        constexpr uint32_t mn_hi = (uint32_t)(DG_OCT >> 4) | (1u << 14);             // MN-major: SBO = octet pitch
        constexpr uint32_t mn_lo = (uint32_t)(128 >> 4) << 16;                      //           LBO = 128 B between 8-row (K) groups
        constexpr uint32_t k_hi = (uint32_t)(128 >> 4) | (1u << 14);                // K-major resident dz: SBO = 128 B (8 images)
        constexpr uint32_t k_lo = (uint32_t)(DG_OCT >> 4) << 16;                    //           LBO = chunk pitch
        uint32_t g = 0, acc_it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++acc_it) {
            const uint32_t j = acc_it & 1;
            if (acc_it >= 2) mbar_wait(&tempty[j], ((acc_it >> 1) - 1) & 1);
            if (MODE == 0) {
                constexpr uint32_t idesc = tc_idesc(128, 128, 1, 1, 1, 1);
                const uint32_t st = g % DG_NST;
                mbar_wait(&full[st], (g / DG_NST) & 1);
                tc_fence_after();
                const uint32_t sb = smem_u32(smem + st * DG_STAGEB);
                const int halves = a.units / 128;
                for (int h = 0; h < halves; ++h) {
                    uint32_t acc = 0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int combo = 0; combo < 3; ++combo) {           // dz_hi p_hi, dz_lo p_hi, dz_hi p_lo
                            const uint32_t a_addr = res + (combo == 1 ? DG_RPLANE : 0) + h * 16 * DG_OCT + ks * 256;
                            const uint32_t b_addr = sb + (combo == 2 ? DG_SPLANE : 0) + ks * 256;
                            umma_f16_if(leader, tmem + j * 256 + h * 128, desc64(mn_lo | ((a_addr & 0x3FFFFu) >> 4), mn_hi),
                                        desc64(mn_lo | ((b_addr & 0x3FFFFu) >> 4), mn_hi), idesc, acc);
                            acc = 1u;
                        }
                }
                umma_commit_if(leader, &empty[st]);
                ++g;
            } else {
                constexpr uint32_t idesc = tc_idesc(128, 64, 1, 1, 1, 0);
                uint32_t acc = 0;
                for (int q = 0; q < kq; ++q, ++g) {
                    const uint32_t st = g % DG_NST;
                    mbar_wait(&full[st], (g / DG_NST) & 1);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(smem + st * DG_STAGEB);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                        for (int combo = 0; combo < 3; ++combo) {           // W_hi dz_hi, W_lo dz_hi, W_hi dz_lo
                            const uint32_t a_addr = sb + (combo == 1 ? DG_SPLANE : 0) + ks * 256;
                            const uint32_t b_addr = res + (combo == 2 ? DG_RPLANE : 0) + (q * 8 + ks * 2) * DG_OCT;
                            umma_f16_if(leader, tmem + j * 64, desc64(mn_lo | ((a_addr & 0x3FFFFu) >> 4), mn_hi),
                                        desc64(k_lo | ((b_addr & 0x3FFFFu) >> 4), k_hi), idesc, acc);
                            acc = 1u;
                        }
                    umma_commit_if(leader, &empty[st]);
                }
            }
            umma_commit_if(leader, &tfull[j]);
        }
    } else {
        // ================================ epilogue ================================
        const int quad = warp & 3;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;
        uint32_t acc_it = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++acc_it) {
            const uint32_t j = acc_it & 1;
            mbar_wait(&tfull[j], (acc_it >> 1) & 1);
            tc_fence_after();
            if (MODE == 0) {
                const int halves = a.units / 128;
                // a lane holds 32 consecutive columns of ITS row (rows are `flat` floats apart): through a padded per-warp tile so that a
                // warp's store instruction writes four 128-byte row segments instead of thirty-two 16-byte pieces 1 MB apart
                float* stg = reinterpret_cast<float*>(smem + DG_OFF_STG) + (warp - (TC_PW + 1)) * (32 * 33);
                for (int h = 0; h < halves; ++h) {
                    float* dst = a.out + (size_t)(h * 128 + quad * 32) * a.flat + (size_t)t * 128;
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        float v[32];
                        tmem_ld32(tmem + lane_off + j * 256 + h * 128 + c * 32, v);
                        tmem_ld_wait();
                        if (h == halves - 1 && c == 3) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty[j]);
                        }
#pragma unroll
                        for (int q = 0; q < 32; ++q) stg[lane * 33 + q] = v[q];
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + (lane >> 3), c4 = lane & 7;
                            const float* sp = stg + r * 33 + c4 * 4;
                            *reinterpret_cast<float4*>(dst + (size_t)r * a.flat + c * 32 + c4 * 4) = make_float4(sp[0], sp[1], sp[2], sp[3]);
                        }
                        __syncwarp();
                    }
                }
            } else {
                float v[64];
                tmem_ld32(tmem + lane_off + j * 64, *reinterpret_cast<float(*)[32]>(v));
                tmem_ld32(tmem + lane_off + j * 64 + 32, *reinterpret_cast<float(*)[32]>(v + 32));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[j]);
                float* dst = a.out + (size_t)t * 128 + row;
#pragma unroll
                for (int b = 0; b < 64; ++b)
                    if (b < a.B) dst[(size_t)b * a.flat] = v[b];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// The first dense layer's forward against the same big matrix:  z[b][u] = sum_k p[b][k] W[u][k], split-K over the CTAs.
// Both operands are K-major (k contiguous in memory): D[u][b] with A = W rows (M = 128-unit halves), B = p rows (N = 64 images), fp16 hi / lo
// pairs; 64 columns of k per stage.  The two (or one) 64-column accumulators live in TMEM for the CTA's whole k range and are written once
// as split-K partials [cta][b][u] -- the layout the fp32 path's reduction / fused head already consume.
// Reference: Classes/CNNModel.py:177-189 (z = input . W^T + b), nn.Linear (ADCNNM.py:77).
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int DF_NST = 2;
constexpr int DF_ALBO = 256 * 16 + 16;                  // pitch of a K chunk (8 columns) of the W tile: 256 unit rows x 16 B, + 16 B
constexpr int DF_APLANE = 8 * DF_ALBO;
constexpr int DF_BLBO = 64 * 16 + 16;
constexpr int DF_BPLANE = 8 * DF_BLBO;
constexpr int DF_STAGEB = 2 * DF_APLANE + 2 * DF_BPLANE;
constexpr int DF_OFF_BAR = DF_NST * DF_STAGEB;
constexpr int DF_TOTAL = DF_OFF_BAR + 128;

__global__ void __launch_bounds__(DG_THREADS, 1) dense_fwd_x3_kernel(DenseFwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DF_OFF_BAR);
    uint64_t* full = bars;                       // [NST] producers -> MMA (count TC_PW)
    uint64_t* empty = bars + DF_NST;             // [NST] MMA -> producers
    uint64_t* done = bars + 2 * DF_NST;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DF_NST + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < DF_NST; ++i) { mbar_init(&full[i], TC_PW); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_chunks = (int)(a.flat / 64);
    const int per = (n_chunks + (int)gridDim.x - 1) / (int)gridDim.x;
    const int c0 = blockIdx.x * per, c1 = min(n_chunks, c0 + per);
    const int halves = a.units / 128;

    if (warp < TC_PW) {
        uint32_t g = 0;
        for (int c = c0; c < c1; ++c, ++g) {
            const uint32_t st = g % DF_NST;
            if (g >= (uint32_t)DF_NST) mbar_wait(&empty[st], ((g / DF_NST) - 1) & 1);
            uint8_t* sa = smem + st * DF_STAGEB;
            uint8_t* sb = sa + 2 * DF_APLANE;
            // W tile: item (unit row m, chunk k8) = 8 consecutive k of row m; all loads of a group of TC_U items first
            const int na = a.units * 8;
            for (int it0 = tid; it0 < na; it0 += 32 * TC_PW * TC_U) {
                float4 v0[TC_U], v1[TC_U];
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    if (it < na) {
                        const float* src = a.w + (size_t)(it >> 3) * a.flat + (size_t)c * 64 + (it & 7) * 8;
                        v0[u] = __ldg(reinterpret_cast<const float4*>(src));
                        v1[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
                    }
                }
#pragma unroll
                for (int u = 0; u < TC_U; ++u) {
                    const int it = it0 + u * 32 * TC_PW;
                    if (it < na) {
                        uint4 hi, lo;
                        split8(v0[u], v1[u], hi, lo, false);
                        uint8_t* dst = sa + (it & 7) * DF_ALBO + (it >> 3) * 16;
                        *reinterpret_cast<uint4*>(dst) = hi;
                        *reinterpret_cast<uint4*>(dst + DF_APLANE) = lo;
                    }
                }
            }
            // p tile: item (image b, chunk k8); images >= B are zero
            for (int it = tid; it < 64 * 8; it += 32 * TC_PW) {
                const int k8 = it & 7, b = it >> 3;
                uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
                if (b < a.B) {
                    const float* src = a.p + (size_t)b * a.flat + (size_t)c * 64 + k8 * 8;
                    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                    split8(v0, v1, hi, lo, false);
                }
                uint8_t* dst = sb + k8 * DF_BLBO + b * 16;
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + DF_BPLANE) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[st]);
        }
    } else if (warp == TC_PW) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = tc_idesc(128, 64, 0, 0, 0, 0);
        constexpr uint32_t sbo_hi = (uint32_t)(128 >> 4) | (1u << 14);
        constexpr uint32_t a_lo_t = (uint32_t)(DF_ALBO >> 4) << 16, b_lo_t = (uint32_t)(DF_BLBO >> 4) << 16;
        uint32_t g = 0, acc = 0;
        for (int c = c0; c < c1; ++c, ++g) {
            const uint32_t st = g % DF_NST;
            mbar_wait(&full[st], (g / DF_NST) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + st * DF_STAGEB), sb = sa + 2 * DF_APLANE;
            for (int h = 0; h < halves; ++h)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int combo = 0; combo < 3; ++combo) {               // W_hi p_hi, W_lo p_hi, W_hi p_lo
                        const uint32_t a_addr = sa + (combo == 1 ? DF_APLANE : 0) + ks * 2 * DF_ALBO + h * 128 * 16;
                        const uint32_t b_addr = sb + (combo == 2 ? DF_BPLANE : 0) + ks * 2 * DF_BLBO;
                        umma_f16_if(leader, tmem + h * 64, desc64(a_lo_t | ((a_addr & 0x3FFFFu) >> 4), sbo_hi), desc64(b_lo_t | ((b_addr & 0x3FFFFu) >> 4), sbo_hi),
                                    idesc, (acc || ks || combo) ? 1u : 0u);
                    }
            acc = 1u;
            umma_commit_if(leader, &empty[st]);
        }
        umma_commit_if(leader, done);
    } else {
        const int quad = warp & 3;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int row = quad * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        float* dst = a.partials + (size_t)blockIdx.x * a.B * a.units;
        for (int h = 0; h < halves; ++h) {
            float v[64];
            tmem_ld32(tmem + lane_off + h * 64, *reinterpret_cast<float(*)[32]>(v));
            tmem_ld32(tmem + lane_off + h * 64 + 32, *reinterpret_cast<float(*)[32]>(v + 32));
            tmem_ld_wait();
            const bool live = c1 > c0;                                      // a CTA without chunks contributes zeros
#pragma unroll
            for (int b = 0; b < 64; ++b)
                if (b < a.B) dst[(size_t)b * a.units + h * 128 + row] = live ? v[b] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int CIN, int COUT, bool ABF>
int launch_conv_t(const TcConvArgs& a, int grid, cudaStream_t s) {
    using L = TcSmem<CIN, COUT>;
    static_assert(L::TOTAL <= 227 * 1024, "conv3x3_x3: shared memory budget");
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_x3_kernel<CIN, COUT, ABF>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    conv3x3_x3_kernel<CIN, COUT, ABF><<<grid, TC_THREADS, L::TOTAL, s>>>(a);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

}  // namespace

bool conv3x3_x3_supported(int Cin, int Cout, int k, int W, int Wo, int pad) {
    return k == 3 && ((Cin == 32 && Cout == 64) || (Cin == 64 && Cout == 32)) && Wo <= 128 && W + 2 * pad <= TC_XP && pad >= 0 && pad <= 2;
}

size_t conv3x3_x3_weight_bytes(int Cin, int Cout) { return (size_t)2 * 9 * (Cin / 8) * Cout * 16; }

int launch_pack_w_x3(const float* w, uint8_t* img, int Cin, int Cout, int CoutPad, bool bf16, cudaStream_t s) {
    pack_w_x3_kernel<<<cdiv(9 * Cin * Cout, 256), 256, 0, s>>>(w, img, Cin, Cout, CoutPad, bf16 ? 1 : 0);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_maxpool2x2_nhwc(const float* y, float* p, int B, int Ho, int Wo, int C, cudaStream_t s) {
    BCAD_REQUIRE(C % 4 == 0, "maxpool2x2_nhwc: channels must be a multiple of 4");
    const int Hp = Ho / 2, Wp = Wo / 2;
    const size_t n = (size_t)B * Hp * Wp * (C / 4);
    if (n == 0) return BCAD_OK;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    maxpool2x2_nhwc_kernel<<<grid, 256, 0, s>>>(y, p, Ho, Wo, Hp, Wp, C / 4, n);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

bool wgrad3x3_x3_supported(int Cin, int Cout, int k, int W, int Wo, int pad) {
    return k == 3 && Cin == 32 && Cout == 64 && Wo <= 128 && W <= 128 && pad >= 0 && pad <= 1;
}

size_t wgrad3x3_x3_partial_floats(int sms) { return (size_t)sms * 12 * 128 * 32; }

int launch_wgrad3x3_x3(const TcWgradArgs& a, float* dw, int CoutPad, int sms, cudaStream_t s) {
    static_assert(WG_TOTAL <= 227 * 1024, "wgrad3x3_x3: shared memory budget");
    BCAD_REQUIRE(wgrad3x3_x3_supported(32, 64, 3, a.W, a.Wo, a.pad), "wgrad3x3_x3: unsupported shape (W %d, Wo %d, pad %d)", a.W, a.Wo, a.pad);
    const int units = a.B * ((a.Ho + 1) / 2) * ((a.Wo + 63) / 64);
    const int grid = units < sms ? units : sms;
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(wgrad3x3_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_TOTAL));
    wgrad3x3_x3_kernel<<<grid, WG_THREADS, WG_TOTAL, s>>>(a);
    BCAD_CUDA_CHECK(cudaGetLastError());
    wgrad_reduce_kernel<<<cdiv(9 * 32 * 64, 256), 256, 0, s>>>(a.partials, grid, dw, CoutPad);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_unpool_mask(const float* g, const float* y, float* dz, int B, int Ho, int Wo, int C, int first_only, float alpha, cudaStream_t s) {
    BCAD_REQUIRE(C % 4 == 0, "unpool_mask: channels must be a multiple of 4");
    const size_t total = (size_t)B * ((Ho + 1) / 2) * ((Wo + 1) / 2) * (C / 4);
    const int blocks = (int)std::min<size_t>((size_t)148 * 16, (total + 255) / 256);
    unpool_mask_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), reinterpret_cast<float4*>(dz), B, Ho, Wo, C / 4,
                                              first_only, alpha);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_conv0_fwd_fused(const float* x, const float* w, const float* bias, float* y, float* p, int B, int H, int W, int Ho, int Wo, int pad, float alpha,
                           int sms, cudaStream_t s) {
    const int units = B * ((Ho + 1) / 2);
    if (units == 0) return BCAD_OK;
    const int grid = std::min(units, sms * 4);
    const int pitch = (W + 2 * pad + 4 + 1) & ~1;
    conv0_fwd_fused_kernel<<<grid, 32 * C0F_WARPS, (size_t)4 * pitch * sizeof(float), s>>>(x, w, bias, y, p, B, H, W, Ho, Wo, pad, alpha);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// scratch: conv0_bwd_fused_parts() * 320 floats; dw: [9][32] (tap-major, filters innermost), db: [32]
int conv0_bwd_fused_parts(int sms) { return sms * 4; }
int launch_conv0_bwd_fused(const float* g, const float* y, const float* x, float* scratch, float* dw, float* db, int B, int H, int W, int Ho, int Wo,
                           int pad, int first_only, float alpha, int sms, cudaStream_t s) {
    const int units = B * (Ho / 2);
    if (units == 0) return BCAD_OK;
    const int grid = std::min(units, conv0_bwd_fused_parts(sms));
    const int pitch = (W + 2 * pad + 4 + 1) & ~1;
    conv0_bwd_fused_kernel<<<grid, 32 * C0B_WARPS, (size_t)4 * pitch * sizeof(float), s>>>(g, y, x, scratch, B, H, W, Ho, Wo, pad, first_only, alpha);
    conv0_bwd_final_kernel<<<2, 160, 0, s>>>(scratch, grid, dw, db);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

bool dense_bwd_x3_supported(int B, int units, long long flat) { return B >= 1 && B <= 64 && (units == 128 || units == 256) && flat % 128 == 0; }

// mode 0: out = dW [units][flat] from dz [B][units] and src = p [B][flat]; mode 1: out = g [B][flat] from dz and src = W [units][flat]
int launch_dense_bwd_x3(const DenseBwdArgs& a, int mode, int sms, cudaStream_t s) {
    static_assert(DG_TOTAL <= 227 * 1024, "dense_bwd_x3: shared memory budget");
    BCAD_REQUIRE(dense_bwd_x3_supported(a.B, a.units, a.flat), "dense_bwd_x3: unsupported shape (B %d, units %d, flat %lld)", a.B, a.units, a.flat);
    const int tiles = (int)(a.flat / 128);
    const int grid = tiles < sms ? tiles : sms;
    if (mode == 0) {
        BCAD_CUDA_CHECK(cudaFuncSetAttribute(dense_bwd_x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_TOTAL));
        dense_bwd_x3_kernel<0><<<grid, DG_THREADS, DG_TOTAL, s>>>(a);
    } else {
        BCAD_CUDA_CHECK(cudaFuncSetAttribute(dense_bwd_x3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_TOTAL));
        dense_bwd_x3_kernel<1><<<grid, DG_THREADS, DG_TOTAL, s>>>(a);
    }
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

bool dense_fwd_x3_supported(int B, int units, long long flat) { return B >= 1 && B <= 64 && (units == 128 || units == 256) && flat % 64 == 0; }
int dense_fwd_x3_splits(long long flat, int sms) {
    const int chunks = (int)(flat / 64);
    const int grid = chunks < sms ? chunks : sms;
    const int per = (chunks + grid - 1) / grid;
    return (chunks + per - 1) / per;                          // every launched CTA owns at least one chunk
}
// partials: [dense_fwd_x3_splits][B][units] fp32
int launch_dense_fwd_x3(const DenseFwdArgs& a, int sms, cudaStream_t s) {
    static_assert(DF_TOTAL <= 227 * 1024, "dense_fwd_x3: shared memory budget");
    BCAD_REQUIRE(dense_fwd_x3_supported(a.B, a.units, a.flat), "dense_fwd_x3: unsupported shape (B %d, units %d, flat %lld)", a.B, a.units, a.flat);
    const int grid = dense_fwd_x3_splits(a.flat, sms);
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(dense_fwd_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DF_TOTAL));
    dense_fwd_x3_kernel<<<grid, DG_THREADS, DF_TOTAL, s>>>(a);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_colsum64(const float* A, float* scratch, float* out, size_t K, cudaStream_t s) {
    const int slab = 2048;
    const int nparts = (int)((K + slab - 1) / slab);
    BCAD_REQUIRE(nparts <= 4096, "colsum64: %zu rows exceed the scratch (4096 slabs)", K);
    colsum64_part_kernel<<<nparts, 256, 0, s>>>(A, scratch, K, slab);
    colsum64_final_kernel<<<1, 64, 0, s>>>(scratch, nparts, out);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_conv3x3_x3(const TcConvArgs& a0, int Cin, int Cout, int sms, cudaStream_t s) {
    BCAD_REQUIRE(conv3x3_x3_supported(Cin, Cout, 3, a0.W, a0.Wo, a0.pad), "conv3x3_x3: unsupported shape %d -> %d, W %d", Cin, Cout, a0.W);
    TcConvArgs a = a0;
    // rows per work item: the split that minimises the busiest CTA's rows (+ 2 halo rows of producer work per item)
    int best = 8, best_cost = 1 << 30;
    for (int br = 4; br <= 64; br *= 2) {
        const int items = a.B * cdiv(a.Ho, br);
        const int cost = cdiv(items, sms) * (br + 1);
        if (cost < best_cost) { best_cost = cost; best = br; }
    }
    a.band_rows = best;
    a.bands = cdiv(a.Ho, a.band_rows);
    const int items = a.B * a.bands;
    const int grid = items < sms ? items : sms;
    if (Cin == 32 && Cout == 64) return a.a_bf16 ? launch_conv_t<32, 64, true>(a, grid, s) : launch_conv_t<32, 64, false>(a, grid, s);
    return a.a_bf16 ? launch_conv_t<64, 32, true>(a, grid, s) : launch_conv_t<64, 32, false>(a, grid, s);
}

}  // namespace bcad
