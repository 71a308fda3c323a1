// First conv block with MANY input channels on tcgen05 (sm_100a): the deployed "advanced" classifier feeds 64- and
// 256-channel bottleneck features (app.py:568-593, SURVEY 8a a16 / 8d secondary shapes), and GRADCAM.py feeds 3-channel
// images (padded to 16 here).
//
//   nhwc_to_c8      fp32 [B][H][W][C]  ->  fp16 C8-planar [B][H][CinPad/8][W][8]   (one pass, HBM-bound, channels zero-padded)
//   conv_wide       3x3 conv + bias + LeakyReLU + 2x2 max-pool, implicit GEMM:
//                   M = 128 pixels of an output row segment, N = Cout, K = 9 taps x KC channels per step, the channel
//                   GROUPS (Cin / KC of them) accumulate into the same TMEM columns, so a CTA keeps a whole band of
//                   RH = 256/Cout output rows in tensor memory (2 halves x 256 columns: the epilogue of one band overlaps
//                   the MMAs of the next) while the producer streams (group, input row) planes through a shared-memory
//                   ring and the group's weight image through a double buffer.  Every input plane is read once per band
//                   (+ 2 halo rows), the weights come from L2.
//   Output: pooled map, C8-planar fp16 -- the input layout of conv_igemm (sm100_kernels.cu).
// Reference semantics: Conv2d(padding=1) / valid conv + bias + LeakyReLU + MaxPool2d(2) (ADCNNM.py:48,76; Classes/CNNModel.py:227-261).
#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100.cuh"
#include "sm100_kernels.h"

namespace bcad {

using namespace sm100;

__device__ __forceinline__ uint32_t wh2u(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }

// ---------------------------------------------------------------------------------------------------------------
// thread = (pixel, channel octet): reads 32 contiguous bytes (one sector), the lanes of a warp write 512 contiguous bytes
__global__ void __launch_bounds__(256) nhwc_to_c8_kernel(const float* __restrict__ x, __half* __restrict__ out, int W, int C,
                                                         int octets, size_t rows) {
    const size_t per_row = (size_t)octets * W;
    const size_t total = rows * per_row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / per_row;
        const int rem = (int)(i - row * per_row);
        const int o = rem / W, px = rem - o * W;
        const float* src = x + (row * W + px) * C + o * 8;
        float v[8];
        if (o * 8 + 8 <= C && (C & 3) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (o * 8 + e < C) ? __ldg(src + e) : 0.f;
        }
        const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
        const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
        reinterpret_cast<uint4*>(out)[i] = make_uint4(wh2u(h0), wh2u(h1), wh2u(h2), wh2u(h3));
    }
}

// fp16x3 form: thread = (pixel, REAL channel octet); writes the hi halves to octets o and 2*octets + o, the lo halves to octets + o
__global__ void __launch_bounds__(256) nhwc_to_c8_x3_kernel(const float* __restrict__ x, __half* __restrict__ out, int W, int C,
                                                            int octets, size_t rows) {
    const size_t per_row = (size_t)octets * W;
    const size_t total = rows * per_row;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / per_row;
        const int rem = (int)(i - row * per_row);
        const int o = rem / W, px = rem - o * W;
        const float* src = x + (row * W + px) * C + o * 8;
        float v[8];
        if (o * 8 + 8 <= C && (C & 3) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (o * 8 + e < C) ? __ldg(src + e) : 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const __half2 h = __floats2half2_rn(v[e], v[e + 1]);
            const float2 hf = __half22float2(h);
            hi[e >> 1] = wh2u(h);
            lo[e >> 1] = pack_f16(v[e] - hf.x, v[e + 1] - hf.y);
        }
        uint4* dst = reinterpret_cast<uint4*>(out) + (row * 3 * octets + o) * W + px;
        const uint4 vh = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[0] = vh;
        dst[(size_t)octets * W] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dst[(size_t)2 * octets * W] = vh;
    }
}

int launch_nhwc_to_c8_x3(const float* x, __half* out, int B, int H, int W, int C, int CinPad, cudaStream_t s) {
    BCAD_REQUIRE(CinPad % 8 == 0 && CinPad >= C, "nhwc_to_c8_x3: bad channel padding %d for %d", CinPad, C);
    const size_t total = (size_t)B * H * (CinPad / 8) * W;
    const int blocks = (int)std::min<size_t>((size_t)148 * 16, (total + 255) / 256);
    nhwc_to_c8_x3_kernel<<<blocks, 256, 0, s>>>(x, out, W, C, CinPad / 8, (size_t)B * H);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_nhwc_to_c8(const float* x, __half* out, int B, int H, int W, int C, int CinPad, cudaStream_t s) {
    BCAD_REQUIRE(CinPad % 8 == 0 && CinPad >= C, "nhwc_to_c8: bad channel padding %d for %d", CinPad, C);
    const size_t total = (size_t)B * H * (CinPad / 8) * W;
    const int blocks = (int)std::min<size_t>((size_t)148 * 16, (total + 255) / 256);
    nhwc_to_c8_kernel<<<blocks, 256, 0, s>>>(x, out, W, C, CinPad / 8, (size_t)B * H);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int WD_XP = 136;              // pixel slots per ring row (128 + 2 halo, multiple of 8)
constexpr int WD_THREADS = 320;         // producer warp, MMA warp, 8 epilogue warps
// ring stages of 2 input rows.  Deep prefetch: a stage is ~1 us of MMA work (groups of 64 channels with a 3-stage ring measured
// 9 % slower).  Measured and dropped: reading the caller's fp32 NHWC directly (converter warps, cp.async staging) -- a channel
// group is then a 128-byte piece of every pixel record, and that strided HBM pattern cost as much as the separate streaming
// conversion pass saves (1.03 ms vs 0.54 + 0.475 ms at 128 x 256x256x64).
constexpr int wd_stages(int) { return 8; }

template <int KC, int COUT>
struct WideSmem {
    static constexpr int CHUNKS = KC / 8;
    static constexpr int LBO = WD_XP * 16;
    static constexpr int ROWB = CHUNKS * LBO;
    static constexpr int WG_BYTES = 9 * CHUNKS * COUT * 16;     // weight image of one channel group
    static constexpr int BIAS_TILE = 2 * COUT * 16;
    static constexpr int ONES_TILE = 2 * 128 * 16;
    static constexpr int OFF_W = 0;                              // 2 weight buffers
    static constexpr int OFF_BIAS = OFF_W + 2 * WG_BYTES;
    static constexpr int OFF_ONES = OFF_BIAS + BIAS_TILE;
    static constexpr int OFF_ZERO = OFF_ONES + ONES_TILE;
    static constexpr int OFF_RING = OFF_ZERO + ROWB;
    static constexpr int STAGES = wd_stages(KC);
    static constexpr int OFF_BAR = OFF_RING + STAGES * 2 * ROWB;
    static constexpr int TOTAL = OFF_BAR + 256;
    static constexpr int RH = 256 / COUT;                        // output rows per TMEM half
};

template <int KC, int COUT>
__global__ void __launch_bounds__(WD_THREADS, 1) conv_wide_kernel(WideArgs a) {
    using L = WideSmem<KC, COUT>;
    constexpr int S = L::STAGES, RH = L::RH;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem + L::OFF_W;
    uint8_t* s_bias = smem + L::OFF_BIAS;
    uint8_t* s_ones = smem + L::OFF_ONES;
    uint8_t* s_zero = smem + L::OFF_ZERO;
    uint8_t* s_ring = smem + L::OFF_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                  // [S]  producer -> MMA   (input rows)
    uint64_t* empty = bars + S;             // [S]  MMA -> producer
    uint64_t* wfull = bars + 2 * S;         // [2]  producer -> MMA   (weights of a group)
    uint64_t* wempty = bars + 2 * S + 2;    // [2]  MMA -> producer
    uint64_t* tfull = bars + 2 * S + 4;     // [2]  MMA -> epilogue   (a band's accumulators)
    uint64_t* tempty = bars + 2 * S + 6;    // [2]  epilogue -> MMA
    uint64_t* bbar = bars + 2 * S + 8;      // bias tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (L::ROWB * (1 + S * 2)) / 16; i += WD_THREADS)         // zero row + ring (halo slots)
        reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < L::ONES_TILE / 16; i += WD_THREADS)                     // octet 0: halves {1,1,0,...}; octet 1: zeros
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        mbar_init(bbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int per_img = a.ybands * a.xsegs;
    const int n_items = a.B * per_img;
    const int octets_all = a.G * L::CHUNKS;                            // channel octets of the whole input
    const size_t in_row_elems = (size_t)octets_all * a.W * 8;          // fp16 elements per input row
    const bool stream_w = (a.G > 1);                                   // one group: its weights stay resident

    if (warp == 0) {
        // ================================ producer ================================
        if (lane == 0) {
            mbar_arrive_expect_tx(bbar, L::BIAS_TILE);
            bulk_g2s(s_bias, a.w_img + (size_t)a.G * L::WG_BYTES, L::BIAS_TILE, bbar);
        }
        uint32_t g = 0, wg = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / per_img, rem = item % per_img;
            const int yb = rem / a.xsegs, x0 = (rem % a.xsegs) * 128;
            const int y0 = yb * RH;
            const int nrows = min(RH, a.Ho - y0);
            const int nstages = (nrows + 1) / 2 + 1;
            const int s_lo = max(0, a.pad - x0), s_hi = min(130, a.W - x0 + a.pad);
            const bool zero_hi = (a.xsegs > 1) && (s_hi < 130);
            const bool zero_lo = (a.xsegs > 1) && (s_lo > 0);
            for (int grp = 0; grp < a.G; ++grp) {
                if (stream_w || wg == 0) {
                    const uint32_t ws = wg & 1;
                    if (lane == 0) {
                        if (wg >= 2) mbar_wait(&wempty[ws], ((wg >> 1) - 1) & 1);
                        mbar_arrive_expect_tx(&wfull[ws], L::WG_BYTES);
                        const uint8_t* wsrc = a.w_img + (size_t)grp * L::WG_BYTES;
                        for (int off = 0; off < L::WG_BYTES; off += 16384)
                            bulk_g2s(s_w + ws * L::WG_BYTES + off, wsrc + off, min(16384, L::WG_BYTES - off), &wfull[ws]);
                    }
                    __syncwarp();
                }
                ++wg;
                const __half* inb = a.in + (size_t)b * a.H * in_row_elems + ((size_t)grp * L::CHUNKS * a.W + (x0 - a.pad + s_lo)) * 8;
                for (int q = 0; q < nstages; ++q, ++g) {
                    const uint32_t slot = g % S;
                    if (g >= S) mbar_wait(&empty[slot], ((g / S) - 1) & 1);
                    if (zero_hi || zero_lo) {
                        for (int i = lane; i < 2 * L::CHUNKS; i += 32) {
                            uint8_t* rowp = s_ring + (slot * 2 + (i / L::CHUNKS)) * L::ROWB + (i % L::CHUNKS) * L::LBO;
                            if (zero_lo) *reinterpret_cast<uint4*>(rowp) = make_uint4(0, 0, 0, 0);
                            if (zero_hi) *reinterpret_cast<uint4*>(rowp + s_hi * 16) = make_uint4(0, 0, 0, 0);
                        }
                        fence_proxy_async();
                        __syncwarp();
                    }
                    if (lane == 0) {
                        uint32_t bytes = 0;
                        for (int r = 0; r < 2; ++r) {
                            const int in_row = y0 - a.pad + 2 * q + r;
                            if (in_row >= 0 && in_row < a.H) bytes += L::CHUNKS * (s_hi - s_lo) * 16;
                        }
                        if (bytes) mbar_arrive_expect_tx(&full[slot], bytes);
                        else mbar_arrive(&full[slot]);
                        for (int r = 0; r < 2; ++r) {
                            const int in_row = y0 - a.pad + 2 * q + r;
                            if (in_row < 0 || in_row >= a.H) continue;
                            uint8_t* dst = s_ring + (slot * 2 + r) * L::ROWB + s_lo * 16;
                            const __half* src = inb + (size_t)in_row * in_row_elems;
                            for (int c = 0; c < L::CHUNKS; ++c)
                                bulk_g2s(dst + c * L::LBO, src + (size_t)c * a.W * 8, (s_hi - s_lo) * 16, &full[slot]);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(128, COUT);
        mbar_wait(bbar, 0);
        uint32_t g = 0, wg = 0, it = 0;
        const uint32_t w_base = smem_u32(s_w), zero_base = smem_u32(s_zero), ring_base = smem_u32(s_ring);
        constexpr uint32_t d_hi = (uint32_t)(128 >> 4) | (1u << 14);            // SBO = 128 B, descriptor version 1
        constexpr uint32_t a_lo_t = (uint32_t)(L::LBO >> 4) << 16;
        constexpr uint32_t b_lo_t = (uint32_t)((COUT * 16) >> 4) << 16;
        const uint32_t ones_lo = ((uint32_t)((128 * 16) >> 4) << 16) | ((smem_u32(s_ones) & 0x3FFFFu) >> 4);
        const uint32_t bias_lo = b_lo_t | ((smem_u32(s_bias) & 0x3FFFFu) >> 4);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int yb = (item % per_img) / a.xsegs;
            const int y0 = yb * RH;
            const int nrows = min(RH, a.Ho - y0);
            const int npairs = (nrows + 1) / 2;
            const uint32_t half = it & 1;
            if (it >= 2) mbar_wait(&tempty[half], ((it >> 1) - 1) & 1);
            tc_fence_after();
            for (int grp = 0; grp < a.G; ++grp, ++wg) {
                const uint32_t ws = stream_w ? (wg & 1) : 0;
                if (stream_w) mbar_wait(&wfull[ws], (wg >> 1) & 1);
                else if (wg == 0) mbar_wait(&wfull[0], 0);
                const uint32_t b_lo0 = b_lo_t | (((w_base + ws * L::WG_BYTES) & 0x3FFFFu) >> 4);
                for (int p = 0; p < npairs; ++p, ++g) {
                    if (p == 0) mbar_wait(&full[g % S], (g / S) & 1);
                    mbar_wait(&full[(g + 1) % S], ((g + 1) / S) & 1);
                    tc_fence_after();
                    for (int r = 0; r < 2; ++r) {
                        if (2 * p + r >= nrows) break;
                        const uint32_t d_tmem = tmem + half * 256 + (2 * p + r) * COUT;
                        if (grp == 0)       // bias K-step: D = ones x {b_hi, b_lo}
                            umma_f16_if(leader, d_tmem, desc64(ones_lo, d_hi), desc64(bias_lo, d_hi), idesc, 0u);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const int i = 2 * p + r + dy;                    // band-local input row
                            const int in_row = y0 - a.pad + i;
                            uint32_t row_base;
                            if (in_row < 0 || in_row >= a.H) row_base = zero_base;
                            else row_base = ring_base + ((((g + (i >> 1) - p) % S) << 1) + (i & 1)) * L::ROWB;
                            const uint32_t a_lo0 = a_lo_t | ((row_base & 0x3FFFFu) >> 4);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int ks = 0; ks < KC / 16; ++ks)
                                    umma_f16_if(leader, d_tmem, desc64(a_lo0 + (uint32_t)((ks * 2 * L::LBO + dx * 16) >> 4), d_hi),
                                                desc64(b_lo0 + (uint32_t)((((dy * 3 + dx) * L::CHUNKS + 2 * ks) * (COUT * 16)) >> 4), d_hi),
                                                idesc, 1u);
                        }
                    }
                    umma_commit_if(leader, &empty[g % S]);
                }
                umma_commit_if(leader, &empty[g % S]);                  // the (band, group)'s last stage
                ++g;
                if (stream_w) umma_commit_if(leader, &wempty[ws]);
            }
            umma_commit_if(leader, &tfull[half]);
        }
    } else {
        // ================================ epilogue (8 warps) ================================
        const int quad = warp & 3;
        const int set = (warp - 2) >> 2;                 // the two warps of a TMEM quadrant split the (row pair, 32-channel) tasks
        const __half2 alpha2 = __float2half2_rn(a.alpha);
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        constexpr int CH = COUT / 32;
        uint32_t it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int b = item / per_img, rem = item % per_img;
            const int yb = rem / a.xsegs;
            const int x = (rem % a.xsegs) * 128 + quad * 32 + lane;
            const int y0 = yb * RH;
            const int nrows = min(RH, a.Ho - y0);
            const int npairs = (nrows + 1) / 2;
            const uint32_t half = it & 1;
            mbar_wait(&tfull[half], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int t = set; t < npairs * CH; t += 2) {
                const int pair = t / CH, ch = t % CH;
                const bool has1 = (2 * pair + 1 < nrows);
                const int py = (y0 + 2 * pair) >> 1, px = x >> 1;
                const bool pool_ok = has1 && py < a.Hp && px < a.Wp && !(x & 1);
                float v0[32], v1[32];
                tmem_ld32(tmem + lane_off + half * 256 + (2 * pair) * COUT + ch * 32, v0);
                tmem_ld32(tmem + lane_off + half * 256 + (2 * pair + 1) * COUT + ch * 32, v1);
                tmem_ld_wait();
                if (a.split_out) {
                    // fp16x3: LeakyReLU and the pool in fp32, then the pooled value as hi + lo halves (2*COUT virtual channels)
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float a0 = fmaxf(v0[q], a.alpha * v0[q]), a1 = fmaxf(v1[q], a.alpha * v1[q]);
                        const float mv = fmaxf(a0, a1);
                        v0[q] = fmaxf(mv, __shfl_xor_sync(0xffffffffu, mv, 1));
                    }
                    if (pool_ok) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const float m0 = v0[cc * 8 + e], m1 = v0[cc * 8 + e + 1];
                                const __half2 h = __floats2half2_rn(m0, m1);
                                const float2 hf = __half22float2(h);
                                ph[e >> 1] = wh2u(h);
                                pl[e >> 1] = pack_f16(m0 - hf.x, m1 - hf.y);
                            }
                            uint4* dh = reinterpret_cast<uint4*>(a.pool_c8) + (((size_t)b * a.Hp + py) * (2 * COUT / 8) + ch * 4 + cc) * a.Wp + px;
                            *dh = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *(dh + (size_t)(COUT / 8) * a.Wp) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    }
                    continue;
                }
                __half2 m[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const __half2 h0 = __floats2half2_rn(v0[2 * q], v0[2 * q + 1]);
                    const __half2 h1 = __floats2half2_rn(v1[2 * q], v1[2 * q + 1]);
                    const __half2 a0 = __hmax2(h0, __hmul2(h0, alpha2)), a1 = __hmax2(h1, __hmul2(h1, alpha2));
                    const __half2 mv = __hmax2(a0, a1);
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, wh2u(mv), 1);
                    m[q] = __hmax2(mv, *reinterpret_cast<const __half2*>(&o));
                }
                if (pool_ok) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int chunk = ch * 4 + cc;
                        uint4* d = reinterpret_cast<uint4*>(a.pool_c8) + (((size_t)b * a.Hp + py) * (COUT / 8) + chunk) * a.Wp + px;
                        *d = make_uint4(wh2u(m[cc * 4]), wh2u(m[cc * 4 + 1]), wh2u(m[cc * 4 + 2]), wh2u(m[cc * 4 + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[half]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KC, int COUT>
static int launch_wide_t(const WideArgs& a, int sms, cudaStream_t s) {
    using L = WideSmem<KC, COUT>;
    static_assert(L::TOTAL <= 227 * 1024, "conv_wide: shared memory budget");
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(conv_wide_kernel<KC, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    const int items = a.B * a.ybands * a.xsegs;
    const int grid = items < sms ? items : sms;
    conv_wide_kernel<KC, COUT><<<grid, WD_THREADS, L::TOTAL, s>>>(a);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int conv_wide_group_channels(int CinPad, int Cout) {
    if (CinPad % 32 == 0) return 32;
    return 16;
}

int conv_wide_rows_per_band(int Cout) { return 256 / Cout; }

int launch_conv_wide(const WideArgs& a, int CinPad, int Cout, int sms, cudaStream_t s) {
    const int KC = conv_wide_group_channels(CinPad, Cout);
    BCAD_REQUIRE(CinPad % 16 == 0 && a.G * KC == CinPad, "conv_wide: %d channels in %d groups of %d", CinPad, a.G, KC);
    BCAD_REQUIRE(a.xsegs == cdiv(a.Wo, 128) && a.ybands == cdiv(a.Ho, conv_wide_rows_per_band(Cout)), "conv_wide: bad tiling");
    if (Cout == 32) {
        if (KC == 32) return launch_wide_t<32, 32>(a, sms, s);
        return launch_wide_t<16, 32>(a, sms, s);
    }
    if (Cout == 64) {
        if (KC == 32) return launch_wide_t<32, 64>(a, sms, s);
        return launch_wide_t<16, 64>(a, sms, s);
    }
    set_error("conv_wide: Cout=%d not supported (32 or 64)", Cout);
    return BCAD_ERR_INVALID;
}

}  // namespace bcad
