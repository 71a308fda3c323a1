// Both conv blocks of the canonical network in ONE persistent kernel, second generation (sm_100a, fp16 mode, 1 -> 32 -> 64 filters).
//
// Same contract as conv_fused_kernel (sm100_fused.cu): the pooled first-block map never goes to HBM, it is produced row by row
// into the shared-memory ring the second block's implicit GEMM reads its A operand from.  What changed is the first block:
//
//  * "patch-union" operands: a pooled pixel's 4x4 input patch IS the K = 16 row of its A operand (16 halves, built once per
//    tile), and the B operand holds the 3x3 filter of every pool class at that class's offset inside the patch
//    (N = 4 classes x 32 filters = 128).  ONE tcgen05.mma (128 x 128 x 16) per tile of 128 pooled pixels instead of four,
//    one tensor-core round trip per tile instead of two, one image build instead of two.
//  * the input patch rows arrive by TENSOR-MAP TMA (cp.async.bulk.tensor.3d, out-of-bounds zero fill = the conv padding): two
//    boxes of 4 rows x 136 floats per tile, issued two tiles ahead by the otherwise idle loader warp.  The innermost box coordinate
//    has to be a multiple of 16 bytes (measured: -1 raises "illegal instruction"), so for pad = 1 the boxes start 3 columns early
//    and a thread reads a patch row as 4 + 8 + 4 bytes.  No global loads, no boundary branches, no patch registers alive across
//    the tensor-core round trip.
//  * software pipeline inside a team: the image of tile n+1 is built while the MMA of tile n is in flight (two 4 KB images per
//    team), and the MMA of tile n+1 is requested as soon as tile n's accumulators are in registers, so the round trip overlaps
//    the ring store of tile n and the build of tile n+2.
//  * two teams with 128 TMEM columns each (all four pool classes at once); the second block keeps conv_fused_kernel's structure
//    (row pairs, 2 x 2 x 64 accumulator columns).  [measured, r02j: three teams + per-row double buffering (2 x 64 columns) starves the
//    tensor pipe -- a two-row queue is shallower than the issuer's wake-up + issue latency: 0.326 vs 0.278 ms with idle teams]
//  * the bias of the first block is added in fp32 after the class maximum (max commutes with adding a per-filter constant).
//
// Warp roles (default F2_TEAM_WARPS = 4, F2_EPI_WARPS = 4: 14 warps, 128 registers): 0-7 two first-block teams (team t produces ring row
// t of every 2-row stage), 8 second-block MMA issuer, 9 weight loader + the teams' TMA boxes and MMAs, 10-13 second-block epilogue.
// TMEM (512 columns): [0,256) second block (2 buffers x 2 rows x 64), [256 + 128 t, +128) team t.
// Compile-time variants (tools/build_variant.sh; profiles/r02_fused2_variants.md): 8-warp teams (two threads per pooled pixel) and / or 8
// epilogue warps (32 filters each) -- fewer registers per thread, measured slower or equal.
// Measured (512 x 256x256x1, one B200): 577 k SM cycles per launch against 800 k of conv_fused_kernel; under sustained load the chip is at
// its power cap and the SM clock inside this kernel settles near 1.3 GHz (clock64 against %globaltimer), so the time gain is smaller.
// Reference semantics: Conv2d + bias + LeakyReLU + MaxPool2d(2), twice (ADCNNM.py:48,72-76; Classes/CNNModel.py:227-261).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100.cuh"
#include "sm100_kernels.h"

namespace bcad {

using namespace sm100;

namespace {

__device__ __forceinline__ uint32_t f2_h2u(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }

constexpr int F2_TEAMS = 2;
#ifndef F2_TEAM_WARPS
#define F2_TEAM_WARPS 4
#endif
#ifndef F2_EPI_WARPS
#define F2_EPI_WARPS 4
#endif
constexpr int F2_TW = F2_TEAM_WARPS;     // warps per team: 8 = two threads per pooled pixel (each owns 16 of the 32 filters and 2 of the 4 patch rows), 4 = one
constexpr int F2_NH = 8 / F2_TW;         // 16-filter halves per team thread
constexpr int F2_EW = F2_EPI_WARPS;      // second-block epilogue warps: 4 = one per TMEM lane quadrant, 8 = two (32 of the 64 filters each)
constexpr int F2_THREADS = 32 * F2_TW * F2_TEAMS + 64 + 32 * F2_EW;     // teams, MMA + loader warps, epilogue warps
static_assert((F2_TW == 4 || F2_TW == 8) && (F2_EW == 4 || F2_EW == 8), "conv_fused2: warp layout");
constexpr int F2_STAGES = 4;             // ring stages of 2 first-block output rows
constexpr int F2_XP = 136;               // pixel slots per ring row
constexpr int F2_C0 = 32, F2_C1 = 64;    // filters of the two blocks
constexpr int F2_BOXW = 136;             // floats per input box row: 2 * 64 pooled pixels + 2 + the alignment offset, rounded to 16 bytes
constexpr int F2_BOXB = 4 * F2_BOXW * 4; // bytes per box (4 input rows)
constexpr int F2_BOXS = 2176;            // box slot in shared memory (128-byte aligned)

struct F2Smem {
    static constexpr int CHUNKS = F2_C0 / 8;
    static constexpr int LBO = F2_XP * 16;
    static constexpr int ROWB = CHUNKS * LBO;
    static constexpr int WBYTES = 9 * CHUNKS * F2_C1 * 16;
    static constexpr int BIAS_TILE = 2 * 2 * F2_C1 * 16;               // [2 chunks][2 x 64 rows][16 B]: the bias rows twice (N = 128 bias MMA)
    static constexpr int ONES_TILE = 2 * 128 * 16;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_ONES = OFF_W1 + WBYTES + BIAS_TILE;
    static constexpr int OFF_ZERO = OFF_ONES + ONES_TILE;
    static constexpr int OFF_RING = OFF_ZERO + ROWB;
    static constexpr int OFF_IM = OFF_RING + F2_STAGES * 2 * ROWB;        // 2 teams x 2 images x [2 chunks][128 rows][16 B]
    static constexpr int OFF_W0 = OFF_IM + F2_TEAMS * 2 * 4096;           // [2 chunks][128 n][16 B]
    static constexpr int OFF_IN = OFF_W0 + 4096;                          // 2 teams x 2 buffers x 2 boxes
    static constexpr int OFF_B0 = OFF_IN + F2_TEAMS * 4 * F2_BOXS;        // first-block bias, fp32 [32]
    static constexpr int OFF_BAR = OFF_B0 + F2_C0 * 4;
    static constexpr int TOTAL = OFF_BAR + 512;
};
static_assert(F2Smem::OFF_IN % 128 == 0, "TMA box destination alignment");
static_assert(F2Smem::OFF_BAR % 8 == 0, "mbarrier alignment");

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

#ifdef F2_TRACE
__device__ long long g_f2_trace[16];
__device__ long long g_f2_cta[1024];        // per CTA: cycles of its second-block issuer | %smid << 48
#define F2_T(x) x
#else
#define F2_T(x)
#endif

__global__ void __launch_bounds__(F2_THREADS, 1) conv_fused2_kernel(FusedArgs a, const __grid_constant__ CUtensorMap xmap) {
    using L = F2Smem;
    constexpr int S = F2_STAGES, COUT = F2_C1, CIN = F2_C0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem + L::OFF_W1;
    uint8_t* s_ones = smem + L::OFF_ONES;
    uint8_t* s_zero = smem + L::OFF_ZERO;
    uint8_t* s_ring = smem + L::OFF_RING;
    uint8_t* s_w0 = smem + L::OFF_W0;
    float* s_b0 = reinterpret_cast<float*>(smem + L::OFF_B0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                           // [S] teams -> MMA        (count 16: 8 warps x 2 teams)
    uint64_t* empty = bars + S;                      // [S] MMA -> teams
    uint64_t* tfull = bars + 2 * S;                  // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * S + 2;             // [2] epilogue -> MMA      (count 4)
    uint64_t* wbar = bars + 2 * S + 4;               // second-block weights landed
    uint64_t* tbar = bars + 2 * S + 5;               // [2] a team's first-block MMA has retired
    uint64_t* ready = tbar + F2_TEAMS;               // [2] a team's next image is built and its accumulators are drained (count 8)
    uint64_t* inbar = ready + F2_TEAMS;              // [2][2] a team's input boxes have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(inbar + 2 * F2_TEAMS);
    uint32_t* f2_stop = tmem_slot + 1;               // set by the second block's issuer when its last item is issued

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (L::ROWB * (1 + S * 2)) / 16; i += F2_THREADS)          // zero row + ring (halo slots stay zero)
        reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < L::ONES_TILE / 16; i += F2_THREADS)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 4096 / 16; i += F2_THREADS)
        reinterpret_cast<uint4*>(s_w0)[i] = reinterpret_cast<const uint4*>(a.w0_img)[i];
    if (tid < F2_C0) s_b0[tid] = a.b0[tid];
    if (tid == 0) {
        *f2_stop = 0u;
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], F2_TW * F2_TEAMS); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], F2_EW); }
        for (int i = 0; i < F2_TEAMS; ++i) { mbar_init(&tbar[i], 1); mbar_init(&ready[i], F2_TW); }
        for (int i = 0; i < 2 * F2_TEAMS; ++i) mbar_init(&inbar[i], 1);
        mbar_init(wbar, 1);
        fence_barrier_init();
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
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
    const int n_items = a.B * a.bands;

    // The CTA's (item, stage) steps are numbered g = 0, 1, ...; a stage is two ring rows and team t produces row t of each.  Every step
    // is a tile of team t, including the (at most one per image) tile whose row lies above / below the map: the second block reads
    // zeros instead of that ring row, so such a tile runs through the pipeline on out-of-bounds (zero) input and is simply not stored.
    struct Cursor { int item, q, nstages, y0; };
    auto enter = [&](Cursor& c) {
        if (c.item >= n_items) return;
        c.y0 = (c.item % a.bands) * a.band_rows;
        c.nstages = (min(a.band_rows, a.Ho - c.y0) + 1) / 2 + 1;
    };
    auto step = [&](Cursor& c) {
        if (++c.q < c.nstages) return;
        c.q = 0;
        c.item += gridDim.x;
        enter(c);
    };
    const bool teams_on = !(a.debug & 16);                        // debug bit 4: first-block teams idle (timing experiment)

    if (warp < F2_TW * F2_TEAMS) {
        // ================================ first-block teams ================================
        const int team = warp / F2_TW, tw = warp & 3;
        const int hf0 = (F2_TW == 8) ? ((warp >> 2) & 1) : 0;     // this thread's halves [hf0, hf0 + F2_NH): half h = filters [16 h, +16) and patch rows 2 h, 2 h + 1
        uint8_t* s_img = smem + L::OFF_IM + team * 8192;
        uint8_t* s_in = smem + L::OFF_IN + team * 4 * F2_BOXS;
        uint64_t* t_inbar = inbar + 2 * team;
        const uint32_t t_tmem = tmem + 256 + team * 128 + ((uint32_t)(tw * 32) << 16);
        const int px = tw * 32 + lane;                            // pooled pixel of the first block = ring pixel slot - pad
        const __half2 alpha2 = __float2half2_rn(a.alpha);
        // A-operand image of a tile: row = pooled pixel, K slot r * 4 + c = patch[r][c] (fp16), chunk 0 = patch rows 0-1, chunk 1 = rows 2-3
        auto build = [&](uint32_t vi) {
            if (!(a.debug & 256)) mbar_wait(&t_inbar[vi & 1u], (vi >> 1) & 1u);
#pragma unroll
            for (int hh = 0; hh < F2_NH; ++hh) {
                const int hf = hf0 + hh;
                const uint8_t* src = s_in + (vi & 1u) * (2 * F2_BOXS) + (px >> 6) * F2_BOXS + ((px & 63) * 2 + a.xoff) * 4 + 2 * hf * (F2_BOXW * 4);
                uint32_t wd[4];
                if (a.xoff & 1) {                                 // odd float offset: 4 + 8 + 4 bytes per patch row
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float p0 = *reinterpret_cast<const float*>(src + r * (F2_BOXW * 4));
                        const float2 p12 = *reinterpret_cast<const float2*>(src + r * (F2_BOXW * 4) + 4);
                        const float p3 = *reinterpret_cast<const float*>(src + r * (F2_BOXW * 4) + 12);
                        wd[2 * r] = pack_f16(p0, p12.x);
                        wd[2 * r + 1] = pack_f16(p12.y, p3);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float2 v0 = *reinterpret_cast<const float2*>(src + r * (F2_BOXW * 4));
                        const float2 v1 = *reinterpret_cast<const float2*>(src + r * (F2_BOXW * 4) + 8);
                        wd[2 * r] = pack_f16(v0.x, v0.y);
                        wd[2 * r + 1] = pack_f16(v1.x, v1.y);
                    }
                }
                *reinterpret_cast<uint4*>(s_img + (vi & 1u) * 4096 + hf * 2048 + px * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
            fence_proxy_async();                                  // generic-proxy stores -> tensor core reads
        };
        auto request = [&]() {                                    // this warp's part of "image built + accumulators drained"
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[team]);
        };

        Cursor c{(int)blockIdx.x, 0, 0, 0};
        enter(c);
        if (teams_on) {
            build(0u);
            request();
        }
        F2_T(long long tr_build = 0; long long tr_tbar = 0; long long tr_fold = 0; long long tr_empty = 0; long long tr_store = 0; long long tr_tiles = 0;)
        F2_T(const long long tr_begin = clock64();)
        for (uint32_t g = 0; c.item < n_items; ++g) {
            uint32_t pk[8 * F2_NH];
            const int py = c.y0 - a.pad + 2 * c.q + team;         // first-block pooled row (= second-block input row) of this tile
            if (teams_on) {
                F2_T(const long long q0 = clock64();)
                const bool have_next = (c.q + 1 < c.nstages) || (c.item + (int)gridDim.x < n_items);
                if (have_next) build(g + 1u);                     // overlaps this tile's MMA
                F2_T(const long long q1 = clock64();)
                mbar_wait(&tbar[team], g & 1u);
                tc_fence_after();
                F2_T(const long long q2 = clock64();)
                // max over the four pool classes, + bias, round to fp16, LeakyReLU(v) = max(v, alpha v) for 0 <= alpha <= 1
#pragma unroll
                for (int hh = 0; hh < F2_NH; ++hh) {
                    const int c0 = 16 * (hf0 + hh);
                    float v0[16], v1[16], v2[16], v3[16];
                    tmem_ld16(t_tmem + c0, v0);
                    tmem_ld16(t_tmem + F2_C0 + c0, v1);
                    tmem_ld16(t_tmem + 2 * F2_C0 + c0, v2);
                    tmem_ld16(t_tmem + 3 * F2_C0 + c0, v3);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(s_b0 + c0 + e);
                        const float m0 = fmaxf(fmaxf(v0[e], v1[e]), fmaxf(v2[e], v3[e])) + bb.x;
                        const float m1 = fmaxf(fmaxf(v0[e + 1], v1[e + 1]), fmaxf(v2[e + 1], v3[e + 1])) + bb.y;
                        const float m2 = fmaxf(fmaxf(v0[e + 2], v1[e + 2]), fmaxf(v2[e + 2], v3[e + 2])) + bb.z;
                        const float m3 = fmaxf(fmaxf(v0[e + 3], v1[e + 3]), fmaxf(v2[e + 3], v3[e + 3])) + bb.w;
                        const __half2 h01 = __floats2half2_rn(m0, m1), h23 = __floats2half2_rn(m2, m3);
                        pk[hh * 8 + (e >> 1)] = f2_h2u(__hmax2(h01, __hmul2(h01, alpha2)));
                        pk[hh * 8 + (e >> 1) + 1] = f2_h2u(__hmax2(h23, __hmul2(h23, alpha2)));
                    }
                }
                if (have_next) request();                         // the MMA of the next tile may start: it overlaps the ring store below
                F2_T(const long long q3 = clock64(); tr_build += q1 - q0; tr_tbar += q2 - q1; tr_fold += q3 - q2; ++tr_tiles;)
            }
            const uint32_t slot = g % S;
            F2_T(const long long q4 = clock64();)
            if (g >= (uint32_t)S) mbar_wait(&empty[slot], ((g / S) - 1) & 1);
            F2_T(const long long q5 = clock64(); tr_empty += q5 - q4;)
            if (teams_on && py >= 0 && py < a.H1) {
                if (px < a.W1) {
                    uint8_t* rowp = s_ring + (slot * 2 + team) * L::ROWB + (px + a.pad) * 16;
#pragma unroll
                    for (int k = 0; k < 2 * F2_NH; ++k) {
                        const int oc = 2 * hf0 + k;
                        const uint4 val = make_uint4(pk[k * 4], pk[k * 4 + 1], pk[k * 4 + 2], pk[k * 4 + 3]);
                        *reinterpret_cast<uint4*>(rowp + oc * L::LBO) = val;
                        if (a.p1_out != nullptr)
                            reinterpret_cast<uint4*>(a.p1_out)[(((size_t)(c.item / a.bands) * a.H1 + py) * (F2_C0 / 8) + oc) * a.W1 + px] = val;
                    }
                }
                fence_proxy_async();                              // ring row (generic-proxy stores) -> tensor core reads
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            F2_T(tr_store += clock64() - q5;)
            step(c);
        }
        F2_T(if (blockIdx.x == 0 && tid == 0) { g_f2_trace[0] = clock64() - tr_begin; g_f2_trace[1] = tr_build; g_f2_trace[2] = tr_tbar;
                                                g_f2_trace[3] = tr_fold; g_f2_trace[4] = tr_empty; g_f2_trace[5] = tr_store; g_f2_trace[6] = tr_tiles; })
    } else if (warp == F2_TW * F2_TEAMS + 1) {
        // ============ second-block weights; then, on behalf of the teams: their input boxes (TMA) and their MMAs ============
        if (lane == 0) {
            constexpr int WB = L::WBYTES + L::BIAS_TILE;
            mbar_arrive_expect_tx(wbar, WB);
            for (int off = 0; off < WB; off += 16384) bulk_g2s(s_w + off, a.w1_img + off, min(16384, WB - off), wbar);
        }
        // (one thread of the CTA issues the first block's MMAs on behalf of the teams; polling `ready` from inside the second
        //  block's issue loop starved that loop -- sm100_fused.cu, tools/fz_decompose.sh)
        const bool leader = elect_one();
        const bool lead0 = leader && !(a.debug & 512);        // debug bit 9: the teams' MMA instructions run predicated off
        constexpr uint32_t idesc0 = make_idesc_f16(128, 4 * F2_C0);
        constexpr uint64_t op_tmpl = ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);   // LBO 2048, SBO 128
        const uint64_t im_desc0 = op_tmpl | (uint64_t)((smem_u32(smem + L::OFF_IM) & 0x3FFFFu) >> 4);
        const uint64_t w0_desc = op_tmpl | (uint64_t)((smem_u32(s_w0) & 0x3FFFFu) >> 4);
        const int nbox = a.W1 > 64 ? 2 : 1;
        // input boxes of tile `vi` of team t: image rows 2 py - pad .. + 3, columns from -pad - xoff and from 128 - pad - xoff (136 wide; xoff makes
        // the box start a multiple of 4 columns: the innermost TMA coordinate has to be 16-byte aligned).  Out-of-bounds elements arrive as zeros.
        auto tma_in = [&](int t, const Cursor& c, uint32_t vi) {
            if (!leader || (a.debug & 256) || !teams_on) return;
            uint8_t* dst = smem + L::OFF_IN + t * 4 * F2_BOXS + (vi & 1u) * (2 * F2_BOXS);
            uint64_t* bar = &inbar[2 * t + (vi & 1u)];
            const int iy0 = 2 * (c.y0 - a.pad + 2 * c.q + t) - a.pad;
            const int b = c.item / a.bands;
            mbar_arrive_expect_tx(bar, (uint32_t)(nbox * F2_BOXB));
            tma_load_3d(dst, &xmap, -a.pad - a.xoff, iy0, b, bar);
            if (nbox == 2) tma_load_3d(dst + F2_BOXS, &xmap, 128 - a.pad - a.xoff, iy0, b, bar);
        };
        // both teams walk the same steps; pf = the step whose boxes are requested next (two tiles ahead of the one being served)
        Cursor pf[F2_TEAMS];
        uint32_t served[F2_TEAMS];
#pragma unroll
        for (int t = 0; t < F2_TEAMS; ++t) {
            pf[t] = Cursor{(int)blockIdx.x, 0, 0, 0};
            enter(pf[t]);
            served[t] = 0;
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k)
                if (pf[t].item < n_items) { tma_in(t, pf[t], k); step(pf[t]); }
        }
        const volatile uint32_t* stop = f2_stop;
        while (!*stop) {
            if (a.debug & 2048) __nanosleep(64);                  // back off between polling rounds: the chip is power-capped, a hot spin costs clock
#pragma unroll
            for (int t = 0; t < F2_TEAMS; ++t) {
                const uint32_t par = served[t] & 1u;
                if (mbar_test_wait(&ready[t], par)) {
                    tc_fence_after();
                    umma_f16_if(lead0, tmem + 256 + t * 128, im_desc0 + (uint64_t)((t * 8192 + par * 4096) >> 4), w0_desc, idesc0, 0u);
                    umma_commit_if(leader, &tbar[t]);
                    // the image of tile served[t] is built, so its box buffer is free: fetch the boxes of tile served[t] + 2 into it
                    if (pf[t].item < n_items) { tma_in(t, pf[t], served[t] + 2u); step(pf[t]); }
                    ++served[t];
                }
            }
        }
    } else if (warp == F2_TW * F2_TEAMS) {
        // ================================ second-block MMA issuer (as conv_fused_kernel) ================================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(128, COUT), idesc2 = make_idesc_f16(128, 2 * COUT);
        F2_T(long long tr_full = 0; long long tr_tempty = 0; long long tr_issue = 0;)
        F2_T(const long long tr_begin = clock64(); unsigned long long tr_ns0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_ns0));)
        mbar_wait(wbar, 0);
        uint32_t g = 0, acc_it = 0;
        const uint32_t w_base = smem_u32(s_w), zero_base = smem_u32(s_zero), ring_base = smem_u32(s_ring);
        constexpr uint32_t d_hi = (uint32_t)(128 >> 4) | (1u << 14);
        constexpr uint32_t a_lo_t = (uint32_t)(L::LBO >> 4) << 16;
        constexpr uint32_t b_lo_t = (uint32_t)((3 * COUT * 16) >> 4) << 16;      // K chunks of one (dx, k-step) are 3 taps apart
        const uint32_t b_lo0 = b_lo_t | ((w_base & 0x3FFFFu) >> 4);
        const uint32_t bias_lo = ((uint32_t)((2 * COUT * 16) >> 4) << 16) | (((w_base + L::WBYTES) & 0x3FFFFu) >> 4);   // [2 chunks][128 rows][16 B]
        const uint32_t ones_lo = ((uint32_t)((128 * 16) >> 4) << 16) | ((smem_u32(s_ones) & 0x3FFFFu) >> 4);
        const bool lead2 = leader && !(a.debug & 128);           // debug bit 7: the second block's MMA instructions run predicated off
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int band = item % a.bands;
            const int y0 = band * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            const int npairs = (nrows + 1) / 2;
            for (int p = 0; p < npairs; ++p, ++g, ++acc_it) {
                F2_T(const long long w0 = clock64();)
                if (p == 0) mbar_wait(&full[g % S], (g / S) & 1);
                mbar_wait(&full[(g + 1) % S], ((g + 1) / S) & 1);
                F2_T(const long long w1 = clock64(); tr_full += w1 - w0;)
                const uint32_t j = acc_it & 1;
                if (acc_it >= 2) mbar_wait(&tempty[j], ((acc_it >> 1) - 1) & 1);
                F2_T(const long long w2 = clock64(); tr_tempty += w2 - w1;)
                tc_fence_after();
                // Weight image (pair layout): per (dx, 8-channel chunk) the three dy taps are contiguous, [w(dy=2) | w(1) | w(0)], 64 rows x 16 B each,
                // so [w(2) | w(1)] and [w(1) | w(0)] are valid N = 128 operands.  Input row i of the pair (i = 0..3 below the pair's first output
                // row) feeds output row 0 with tap dy = i and output row 1 with dy = i - 1: rows 1 and 2 do both with ONE N = 128 MMA into the two
                // adjacent accumulators (the A operand is read once instead of twice); rows 0 and 3 are N = 64.  25 MMAs per pair instead of 38,
                // same products in the same order per accumulator.
                auto row_desc = [&](int i_abs) -> uint32_t {
                    const int in_row = y0 - a.pad + i_abs;
                    uint32_t row_base;
                    if (in_row < 0 || in_row >= a.H1) row_base = zero_base;
                    else row_base = ring_base + ((((g + (i_abs >> 1) - p) % S) << 1) + (i_abs & 1)) * L::ROWB;
                    return a_lo_t | ((row_base & 0x3FFFFu) >> 4);
                };
                auto w_off = [&](int dy, int dx, int ks) -> uint32_t { return (uint32_t)((((dx * L::CHUNKS + 2 * ks) * 3 + (2 - dy)) * (COUT * 16)) >> 4); };
                const uint32_t d_tmem = tmem + j * (2 * COUT);
                const bool pair = (2 * p + 1 < nrows) && !(a.debug & 1024);      // debug bit 10: no paired taps (A/B timing)
                if (a.debug & 8) {                                               // debug bit 3: no second-block MMAs (timing experiment)
                } else if (pair) {
                    umma_f16_if(lead2, d_tmem, desc64(ones_lo, d_hi), desc64(bias_lo, d_hi), idesc2, 0u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t a_lo0 = row_desc(2 * p + i);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int ks = 0; ks < CIN / 16; ++ks) {
                                const uint64_t ad = desc64(a_lo0 + (uint32_t)((ks * 2 * L::LBO + dx * 16) >> 4), d_hi);
                                if (i == 0) umma_f16_if(lead2, d_tmem, ad, desc64(b_lo0 + w_off(0, dx, ks), d_hi), idesc, 1u);
                                else if (i == 3) umma_f16_if(lead2, d_tmem + COUT, ad, desc64(b_lo0 + w_off(2, dx, ks), d_hi), idesc, 1u);
                                else umma_f16_if(lead2, d_tmem, ad, desc64(b_lo0 + w_off(i, dx, ks), d_hi), idesc2, 1u);     // [w(i) | w(i-1)]
                            }
                    }
                } else {
                    for (int r = 0; r < 2; ++r) {
                        if (2 * p + r >= nrows) break;
                        umma_f16_if(lead2, d_tmem + r * COUT, desc64(ones_lo, d_hi), desc64(bias_lo, d_hi), idesc, 0u);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint32_t a_lo0 = row_desc(2 * p + r + dy);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int ks = 0; ks < CIN / 16; ++ks)
                                    umma_f16_if(lead2, d_tmem + r * COUT, desc64(a_lo0 + (uint32_t)((ks * 2 * L::LBO + dx * 16) >> 4), d_hi),
                                                desc64(b_lo0 + w_off(dy, dx, ks), d_hi), idesc, 1u);
                        }
                    }
                }
                umma_commit_if(leader, &empty[g % S]);
                umma_commit_if(leader, &tfull[j]);
                F2_T(tr_issue += clock64() - w2;)
            }
            umma_commit_if(leader, &empty[g % S]);
            ++g;
        }
        // every team MMA has retired by now (the second block consumed their rows): release the loader warp's service loop
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(f2_stop) = 1u;
        F2_T(if (lane == 0 && blockIdx.x < 1024) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                                                 g_f2_cta[blockIdx.x] = (clock64() - tr_begin) | ((long long)smid << 48); })
        F2_T(if (blockIdx.x == 0 && lane == 0) { g_f2_trace[8] = clock64() - tr_begin; g_f2_trace[9] = tr_full; g_f2_trace[10] = tr_tempty;
                                                 g_f2_trace[11] = tr_issue; g_f2_trace[12] = acc_it;
                                                 unsigned long long ns1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1)); g_f2_trace[13] = (long long)(ns1 - tr_ns0); })
    } else {
        // ================================ second-block epilogue (one or two warps per TMEM lane quadrant; as conv_fused_kernel) ================================
        const int quad = warp & 3;
        const int ew = warp - (F2_TW * F2_TEAMS + 2);             // 8 epilogue warps: warps ew and ew + 4 share a quadrant, one half of the filters each
        const int half_lo = (F2_EW == 8) ? (ew >> 2) : 0, half_hi = (F2_EW == 8) ? half_lo + 1 : COUT / 32;
        const __half2 alpha2 = __float2half2_rn(a.alpha);
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const int x = quad * 32 + lane;
        uint32_t acc_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = item / a.bands, band = item % a.bands;
            const int y0 = band * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            const int npairs = (nrows + 1) / 2;
            for (int p = 0; p < npairs; ++p, ++acc_it) {
                const uint32_t j = acc_it & 1;
                mbar_wait(&tfull[j], (acc_it >> 1) & 1);
                tc_fence_after();
                if (a.debug & 4) {                         // timing experiment: epilogue does nothing
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[j]);
                    continue;
                }
                const int t0 = y0 + 2 * p;
                const bool has1 = (2 * p + 1 < nrows);
                const int py = t0 >> 1, pxx = x >> 1;
                const bool pool_ok = has1 && py < a.Hp && pxx < a.Wp && !(x & 1);
#pragma unroll 1
                for (int half = half_lo; half < half_hi; ++half) {
                    float v0[32], v1[32];
                    tmem_ld32(tmem + lane_off + j * (2 * COUT) + half * 32, v0);
                    tmem_ld32(tmem + lane_off + j * (2 * COUT) + COUT + half * 32, v1);
                    tmem_ld_wait();
                    __half2 a0[16], a1[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const __half2 h0 = __floats2half2_rn(v0[2 * q], v0[2 * q + 1]);
                        const __half2 h1 = __floats2half2_rn(v1[2 * q], v1[2 * q + 1]);
                        a0[q] = __hmax2(h0, __hmul2(h0, alpha2));
                        a1[q] = __hmax2(h1, __hmul2(h1, alpha2));
                    }
                    if (a.act != nullptr && x < a.Wo) {
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int chunk = half * 4 + cc;
                            uint4* d0 = reinterpret_cast<uint4*>(a.act) + (((size_t)b * a.Ho + t0) * (COUT / 8) + chunk) * a.Wo + x;
                            *d0 = make_uint4(f2_h2u(a0[cc * 4]), f2_h2u(a0[cc * 4 + 1]), f2_h2u(a0[cc * 4 + 2]), f2_h2u(a0[cc * 4 + 3]));
                            if (has1) {
                                uint4* d1 = d0 + (size_t)(COUT / 8) * a.Wo;
                                *d1 = make_uint4(f2_h2u(a1[cc * 4]), f2_h2u(a1[cc * 4 + 1]), f2_h2u(a1[cc * 4 + 2]), f2_h2u(a1[cc * 4 + 3]));
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const __half2 mv = __hmax2(a0[q], a1[q]);
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, f2_h2u(mv), 1);
                        a0[q] = __hmax2(mv, *reinterpret_cast<const __half2*>(&o));
                    }
                    if (pool_ok && a.pool_fc != nullptr) {
                        const int row = b & 127;
                        uint8_t* base = a.pool_fc + ((((size_t)(b >> 7) * a.Hp * a.Wp) + (size_t)py * a.Wp + pxx) * 128 + row) * 128;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int chunk = (half * 4 + cc) ^ (row & 7);
                            *reinterpret_cast<uint4*>(base + chunk * 16) =
                                make_uint4(f2_h2u(a0[cc * 4]), f2_h2u(a0[cc * 4 + 1]), f2_h2u(a0[cc * 4 + 2]), f2_h2u(a0[cc * 4 + 3]));
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace

// the tensor-map TMA needs 16-byte aligned image rows; everything else is conv_fused_supported()'s contract
bool conv_fused2_supported(const float* x, int H, int W) {
    (void)H;
    return (W % 4) == 0 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 && encode_tiled_fn() != nullptr;
}

int launch_conv_fused2(const FusedArgs& a, int sms, cudaStream_t s) {
    static_assert(F2Smem::TOTAL <= 227 * 1024, "conv_fused2: shared memory budget");
    BCAD_REQUIRE(a.band_rows % 2 == 0 && a.bands == cdiv(a.Ho, a.band_rows), "conv_fused2: bad banding");
    BCAD_REQUIRE(a.W1 <= 128 && a.Wo <= 128, "conv_fused2: second-block map wider than 128");
    BCAD_REQUIRE(a.plain0 && a.b0 != nullptr, "conv_fused2: fp16 mode only (union weight image + fp32 bias)");
    BCAD_REQUIRE(a.xoff >= 0 && a.xoff <= 6, "conv_fused2: bad box offset %d", a.xoff);
    EncodeTiledFn enc = encode_tiled_fn();
    BCAD_REQUIRE(enc != nullptr, "conv_fused2: cuTensorMapEncodeTiled is not available from this driver");
    // the input as a 3-D tensor (W, H, B) of fp32; a box = 4 rows x 136 columns of one image, out-of-bounds elements read as 0
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
    const cuuint64_t gstr[2] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.W * a.H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)F2_BOXW, 4, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("conv_fused2: cuTensorMapEncodeTiled failed (%d) for a %dx%dx%d input", (int)cr, a.B, a.H, a.W);
        return BCAD_ERR_CUDA;
    }
    const int items = a.B * a.bands;
    const int grid = items < sms ? items : sms;
    BCAD_CUDA_CHECK(cudaFuncSetAttribute(conv_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2Smem::TOTAL));
    conv_fused2_kernel<<<grid, F2_THREADS, F2Smem::TOTAL, s>>>(a, map);
    BCAD_CUDA_CHECK(cudaGetLastError());
#ifdef F2_TRACE
    static int n_launch = 0;
    ++n_launch;
    const char* every = getenv("BCAD_F2_TRACE_AT");       // print the trace of the N-th launch (it ran back to back with the ones before it)
    if ((a.debug & 64) || (every != nullptr && n_launch == atoi(every))) {
        long long h[16];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_f2_trace, sizeof(h));
        fprintf(stderr, "f2_trace team0: total %lld | build %lld | wait mma %lld | fold %lld | wait ring slot %lld | store %lld | tiles %lld\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
        {
            static long long hc[1024];
            cudaMemcpyFromSymbol(hc, g_f2_cta, sizeof(long long) * (grid < 1024 ? grid : 1024));
            long long mn = 1LL << 60, mx = 0, sum = 0; int imn = 0, imx = 0;
            for (int i = 0; i < grid && i < 1024; ++i) {
                const long long v = hc[i] & ((1LL << 48) - 1);
                sum += v;
                if (v < mn) { mn = v; imn = i; }
                if (v > mx) { mx = v; imx = i; }
            }
            fprintf(stderr, "f2_trace per-CTA issuer cycles: min %lld (cta %d, sm %lld) mean %lld max %lld (cta %d, sm %lld)\n", mn, imn, hc[imn] >> 48,
                    sum / grid, mx, imx, hc[imx] >> 48);
            fprintf(stderr, "f2_trace per-CTA cycles/1000:");
            for (int i = 0; i < grid && i < 1024; ++i) fprintf(stderr, " %lld", (hc[i] & ((1LL << 48) - 1)) / 1000);
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "f2_trace issuer: total %lld | wait full %lld | wait tempty %lld | issue %lld | pairs %lld | %lld ns = %.0f MHz (launch %d)\n", h[8], h[9], h[10],
                h[11], h[12], h[13], h[13] > 0 ? 1e3 * (double)h[8] / (double)h[13] : 0.0, n_launch);
    }
#endif
    return BCAD_OK;
}

}  // namespace bcad
