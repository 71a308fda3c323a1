// Both conv blocks of the canonical network in ONE persistent kernel (sm_100a, fp16 mode, Cin = 1 -> 32 -> 64 filters).
//
// The pooled output of the first block never goes to HBM: four 128-thread "teams" compute it row by row on the tensor core
// (the conv_first_tc_kernel scheme: im2col rows with hi/lo-split operands, pool-class accumulators in one TMEM lane, pooled
// epilogue) and store it straight into the shared-memory ring the second block's implicit GEMM reads its A operand from
// (the conv_igemm_kernel scheme: ring row = C8-planar pixels 16 B apart, tap shift = descriptor start offset, bias K-step,
// TMEM double-buffered accumulators, half2 epilogue with fused 2x2 pool, fp16 A store + fc1 A-tile store).
// Per 512 x 256x256 images this removes 0.54 GB of HBM writes and 0.55 GB of reads, and -- more important -- the first
// block's CUDA-core work (instruction-issue bound) now overlaps the second block's HBM-bound stores on the same SM.
//
// Warp roles (22 warps): 0-15 four first-block teams (team t produces ring row t&1 of the stages with parity t>>1),
// 16 second-block MMA issuer, 17 weight loader + issuer of the teams' (first-block) MMAs, 18-21 second-block epilogue.  TMEM (512 columns): [0,256) second block
// (2 buffers x 2 rows x 64), [256 + 64 t, +64) team t.  A team's tile (128 pooled pixels) runs in TWO passes of two pool
// classes each (2 x 32 accumulator columns; the running maximum stays in registers as packed halves): four teams are what
// it takes to hide a tile's dependent build -> MMA -> TMEM-load chain, and four 128-column teams would not fit in TMEM.
// Reference semantics: Conv2d/valid conv + bias + LeakyReLU + MaxPool2d(2), twice (ADCNNM.py:48,76; Classes/CNNModel.py:227-261).
#include <stdio.h>

#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100.cuh"
#include "sm100_kernels.h"

namespace bcad {

using namespace sm100;

__device__ __forceinline__ uint32_t fh2u(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }

constexpr int FZ_TEAMS = 4;
constexpr int FZ_THREADS = 128 * FZ_TEAMS + 64 + 128;     // teams, MMA + loader warps, 4 epilogue warps
constexpr int FZ_STAGES = 4;             // ring stages of 2 first-block output rows
constexpr int FZ_XP = 136;               // pixel slots per ring row
constexpr int FZ_C0 = 32, FZ_C1 = 64;    // filters of the two blocks

struct FusedSmem {
    static constexpr int CHUNKS = FZ_C0 / 8;
    static constexpr int LBO = FZ_XP * 16;
    static constexpr int ROWB = CHUNKS * LBO;
    static constexpr int WBYTES = 9 * CHUNKS * FZ_C1 * 16;
    static constexpr int BIAS_TILE = 2 * FZ_C1 * 16;
    static constexpr int ONES_TILE = 2 * 128 * 16;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_ONES = OFF_W1 + WBYTES + BIAS_TILE;
    static constexpr int OFF_ZERO = OFF_ONES + ONES_TILE;
    static constexpr int OFF_RING = OFF_ZERO + ROWB;
    static constexpr int OFF_IM = OFF_RING + FZ_STAGES * 2 * ROWB;       // 4 teams x 16 KB im2col images (two pool classes)
    static constexpr int OFF_W0 = OFF_IM + FZ_TEAMS * 16384;                    // [4 chunks][32][16 B]
    static constexpr int OFF_BAR = OFF_W0 + 4 * FZ_C0 * 16;
    static constexpr int TOTAL = OFF_BAR + 256;
};

// PLAIN: first-block operands as plain fp16 (the fp16 mode): K slots [x(9) 1 1 0..] against [w(9) b_hi b_lo 0..], one K-step
// per pool class and two 16-byte im2col stores per row; !PLAIN: the hi/lo split image (three stores, two K-steps)
#ifdef FZ_TRACE
__device__ long long g_fz_trace[8];
#define FZ_T(x) x
#else
#define FZ_T(x)
#endif

template <bool PLAIN>
__global__ void __launch_bounds__(FZ_THREADS, 1) conv_fused_kernel(FusedArgs a) {
    using L = FusedSmem;
    constexpr int S = FZ_STAGES, COUT = FZ_C1, CIN = FZ_C0;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem + L::OFF_W1;
    uint8_t* s_ones = smem + L::OFF_ONES;
    uint8_t* s_zero = smem + L::OFF_ZERO;
    uint8_t* s_ring = smem + L::OFF_RING;
    uint8_t* s_w0 = smem + L::OFF_W0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* full = bars;                       // [S] teams -> MMA        (count 8: 4 warps x 2 teams)
    uint64_t* empty = bars + S;                  // [S] MMA -> teams
    uint64_t* tfull = bars + 2 * S;              // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * S + 2;         // [2] epilogue -> MMA     (count 4)
    uint64_t* wbar = bars + 2 * S + 4;           // second-block weights landed
    uint64_t* tbar = bars + 2 * S + 5;           // [4] a team's first-block MMAs have retired
    uint64_t* ready = bars + 2 * S + 5 + FZ_TEAMS;   // [4] a team's im2col image is built and its accumulator columns are drained (count 4)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5 + 2 * FZ_TEAMS);
    uint32_t* fz_stop = tmem_slot + 1;           // set by the second block's issuer when its last item is done

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (L::ROWB * (1 + S * 2)) / 16; i += FZ_THREADS)          // zero row + ring (halo slots stay zero)
        reinterpret_cast<uint4*>(s_zero)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < L::ONES_TILE / 16; i += FZ_THREADS)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    for (int i = tid; i < ((PLAIN ? 2 : 4) * FZ_C0 * 16) / 16; i += FZ_THREADS)
        reinterpret_cast<uint4*>(s_w0)[i] = reinterpret_cast<const uint4*>(a.w0_img)[i];
    if (tid == 0) {
        *fz_stop = 0u;
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 8); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        for (int i = 0; i < FZ_TEAMS; ++i) { mbar_init(&tbar[i], 1); mbar_init(&ready[i], 4); }
        mbar_init(wbar, 1);
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
    const int n_items = a.B * a.bands;

    if (warp < 4 * FZ_TEAMS) {
        // ================================ first-block teams ================================
        const int team = warp >> 2, tw = warp & 3, ttid = tid & 127;
        const int trow = team & 1, tpar = team >> 1;             // ring row inside a stage; parity of the stages this team serves
        uint8_t* s_a = smem + L::OFF_IM + team * 16384;           // 2 classes x [4 chunks][128 rows][16 B]
        const uint32_t t_tmem = tmem + 256 + team * 64;
        const uint32_t lane_off = (uint32_t)(tw * 32) << 16;
        const int px = ttid;                                      // pooled pixel of the first block = ring pixel slot - pad
        const __half2 alpha2 = __float2half2_rn(a.alpha);
        // the global sequence of (item, stage) steps is numbered g = 0, 1, ...; this team serves the steps with g % 2 == tpar
        struct Cursor { int item, q, nstages, b, y0; };
        auto enter = [&](Cursor& c) {
            if (c.item >= n_items) return;
            c.b = c.item / a.bands;
            c.y0 = (c.item % a.bands) * a.band_rows;
            c.nstages = (min(a.band_rows, a.Ho - c.y0) + 1) / 2 + 1;
        };
        auto step = [&](Cursor& c) {
            if (++c.q < c.nstages) return;
            c.q = 0;
            c.item += gridDim.x;
            enter(c);
        };
        auto row_of = [&](const Cursor& c) { return c.y0 - a.pad + 2 * c.q + trow; };     // first-block pooled row (= second-block input row)
        auto valid_row = [&](const Cursor& cc) { const int r = row_of(cc); return cc.item < n_items && r >= 0 && r < a.H1 && !(a.debug & 16); };
        float patch[16], lo[16];
        auto load_patch = [&](const Cursor& c) {
            const int py = row_of(c);
            const float* xb = a.x + (size_t)c.b * a.H * a.W;
            const int iy0 = 2 * py - a.pad, ix0 = 2 * px - a.pad;
            if (iy0 >= 0 && iy0 + 3 < a.H && ix0 >= 0 && ix0 + 3 < a.W) {
                const float* p0 = xb + (size_t)iy0 * a.W + ix0;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) patch[r * 4 + cc] = __ldg(p0 + (size_t)r * a.W + cc);
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int iy = iy0 + r, ix = ix0 + cc;
                        patch[r * 4 + cc] = (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) ? __ldg(xb + (size_t)iy * a.W + ix) : 0.f;
                    }
            }
        };
        // im2col rows of the two classes of pool-window row `qr`; K slots [x(9) 1 | x_lo(9) 1 | x(9) 0 0 0] (conv_first_tc_kernel)
        auto build = [&](int qr) {
#pragma unroll
            for (int qc = 0; qc < 2; ++qc) {
                float xt[9], lt[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int pi = (qr + t / 3) * 4 + (qc + t % 3);
                    xt[t] = patch[pi];
                    lt[t] = lo[pi];
                }
                uint32_t wd[16];
                wd[0] = pack_f16(xt[0], xt[1]); wd[1] = pack_f16(xt[2], xt[3]); wd[2] = pack_f16(xt[4], xt[5]); wd[3] = pack_f16(xt[6], xt[7]);
                wd[4] = pack_f16(xt[8], 1.f);
                if constexpr (PLAIN) {
                    wd[5] = 0x00003C00u;                      // (1, 0): the b_lo slot
                    wd[6] = wd[7] = 0u;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
                        *reinterpret_cast<uint4*>(s_a + qc * 8192 + ch * 2048 + ttid * 16) =
                            make_uint4(wd[ch * 4], wd[ch * 4 + 1], wd[ch * 4 + 2], wd[ch * 4 + 3]);
                    continue;
                }
                wd[5] = pack_f16(lt[0], lt[1]); wd[6] = pack_f16(lt[2], lt[3]); wd[7] = pack_f16(lt[4], lt[5]); wd[8] = pack_f16(lt[6], lt[7]);
                wd[9] = pack_f16(lt[8], 1.f);
                wd[10] = wd[0]; wd[11] = wd[1]; wd[12] = wd[2]; wd[13] = wd[3];
                wd[14] = wd[4] & 0x0000ffffu;
                wd[15] = 0u;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    *reinterpret_cast<uint4*>(s_a + qc * 8192 + ch * 2048 + ttid * 16) =
                        make_uint4(wd[ch * 4], wd[ch * 4 + 1], wd[ch * 4 + 2], wd[ch * 4 + 3]);
            }
        };
        // hand the image to the MMA warp: ONE warp issues every tcgen05.mma of the CTA, so it can slot a team's four MMAs between
        // two rows of the second block instead of letting them queue behind a whole row pair (38 MMAs, ~2500 cycles)
        // (r02 experiment, dropped: the team issuing its own MMAs after a 128-thread named barrier instead of this hand-over measured
        //  0.531 vs 0.510 ms for the kernel and 0.412 vs 0.394 ms for the teams alone: the hop through the loader warp is not the cost)
        auto issue = [&]() {
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[team]);
        };
        // running maximum over the window's classes, 32 filters as 16 packed halves (rounding to fp16 commutes with max)
        __half2 best[16];
        auto fold = [&](bool first) {
#pragma unroll
            for (int c0 = 0; c0 < FZ_C0; c0 += 16) {
                float v0[16], v1[16];
                tmem_ld16(t_tmem + lane_off + c0, v0);
                tmem_ld16(t_tmem + lane_off + FZ_C0 + c0, v1);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const __half2 m = __floats2half2_rn(fmaxf(v0[e], v1[e]), fmaxf(v0[e + 1], v1[e + 1]));
                    best[(c0 + e) >> 1] = first ? m : __hmax2(best[(c0 + e) >> 1], m);
                }
            }
        };
        Cursor c{(int)blockIdx.x, 0, 0, 0, 0}, n{};
        enter(c);
        if (tpar == 1 && c.item < n_items) step(c);              // teams 2,3 start at step 1
        if (valid_row(c)) load_patch(c);
        uint32_t phase = 0;
        for (uint32_t g = tpar; c.item < n_items; g += 2) {
            const bool valid = valid_row(c);
            if (valid) {
#pragma unroll
                for (int e = 0; e < 16; ++e) lo[e] = patch[e] - __half2float(__float2half_rn(patch[e]));
                build(0);                                         // the image is free: this team waited for its last MMAs below
                issue();
                mbar_wait(&tbar[team], phase);
                phase ^= 1;
                tc_fence_after();
                build(1);                                         // (the tensor core is done reading the image)
                fold(true);
                tc_fence_before();
                issue();
            }
            n = c;
            step(n);
            if (n.item < n_items) step(n);                       // two steps ahead: the other team pair serves the one in between
            if (valid_row(n)) load_patch(n);                      // next row's inputs: the loads fly during the MMAs / epilogue
            const uint32_t slot = g % S;
            if (g >= (uint32_t)S) mbar_wait(&empty[slot], ((g / S) - 1) & 1);
            if (valid) {
                mbar_wait(&tbar[team], phase);
                phase ^= 1;
                tc_fence_after();
                fold(false);
                tc_fence_before();
                if (px < a.W1) {
                    uint8_t* rowp = s_ring + (slot * 2 + trow) * L::ROWB + (px + a.pad) * 16;
                    const int py = row_of(c);
#pragma unroll
                    for (int oc = 0; oc < FZ_C0 / 8; ++oc) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 m = best[oc * 4 + e];
                            pk[e] = fh2u(__hmax2(m, __hmul2(m, alpha2)));     // LeakyReLU(v) = max(v, alpha v), 0 <= alpha <= 1
                        }
                        const uint4 val = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(rowp + oc * L::LBO) = val;
                        if (a.p1_out != nullptr)
                            reinterpret_cast<uint4*>(a.p1_out)[(((size_t)c.b * a.H1 + py) * (FZ_C0 / 8) + oc) * a.W1 + px] = val;
                    }
                }
                fence_proxy_async();                              // ring row (generic-proxy stores) -> tensor core reads
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[slot]);
            c = n;
        }
    } else if (warp == 4 * FZ_TEAMS + 1) {
        // ================================ second-block weights, then the teams' MMA issuer ================================
        if (lane == 0) {
            constexpr int WB = L::WBYTES + L::BIAS_TILE;
            mbar_arrive_expect_tx(wbar, WB);
            for (int off = 0; off < WB; off += 16384) bulk_g2s(s_w + off, a.w1_img + off, min(16384, WB - off), wbar);
        }
        // The first block's MMAs are issued HERE, on behalf of the teams, not by the second block's issuer: polling four `ready`
        // barriers costs ~600 cycles a round (mbarrier.test_wait), and inside the second block's issue loop those rounds (two per
        // row pair + every wait) kept its single thread from feeding the tensor pipe: 79 cycles per 128x64x16 MMA with the teams
        // idle against the pipe's own 48 (tools/fz_decompose.sh, tools/umma_microbench.py).  Any thread of the CTA may issue
        // tcgen05.mma; the two issuers' instructions interleave in the pipe's queue exactly as they did before.
        const bool leader = elect_one();
        constexpr uint32_t idesc0 = make_idesc_f16(128, FZ_C0);
        constexpr uint64_t a_tmpl = ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        constexpr uint64_t b_tmpl = ((uint64_t)((FZ_C0 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t im_desc0 = a_tmpl | (uint64_t)((smem_u32(smem + L::OFF_IM) & 0x3FFFFu) >> 4);
        const uint64_t w0_desc0 = b_tmpl | (uint64_t)((smem_u32(s_w0) & 0x3FFFFu) >> 4);
        uint32_t rphase = 0;                                  // bit t = parity team t's `ready` barrier is expected to complete next
        const volatile uint32_t* stop = fz_stop;
        while (!*stop) {
#pragma unroll
            for (int t = 0; t < FZ_TEAMS; ++t) {
                if (mbar_test_wait(&ready[t], (rphase >> t) & 1)) {
                    rphase ^= 1u << t;
                    tc_fence_after();
#pragma unroll
                    for (int qc = 0; qc < 2; ++qc)
#pragma unroll
                        for (int ks = 0; ks < (PLAIN ? 1 : 2); ++ks)
                            umma_f16_if(leader, tmem + 256 + t * 64 + qc * FZ_C0, im_desc0 + (uint64_t)((t * 16384 + qc * 8192 + ks * 4096) >> 4),
                                        w0_desc0 + (uint64_t)((ks * 2 * FZ_C0 * 16) >> 4), idesc0, ks);
                    umma_commit_if(leader, &tbar[t]);
                }
            }
        }
    } else if (warp == 4 * FZ_TEAMS) {
        // ================================ second-block MMA issuer ================================
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(128, COUT);
        FZ_T(long long tr_full = 0; long long tr_tempty = 0; long long tr_issue = 0; long long tr_serv = 0; long long tr_nserv = 0;)
        FZ_T(const long long tr_begin = clock64();)
        auto wait_serving = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };    // (the loader warp serves the teams now)
        mbar_wait(wbar, 0);
        uint32_t g = 0, acc_it = 0;
        const uint32_t w_base = smem_u32(s_w), zero_base = smem_u32(s_zero), ring_base = smem_u32(s_ring);
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int band = item % a.bands;
            const int y0 = band * a.band_rows;
            const int nrows = min(a.band_rows, a.Ho - y0);
            const int npairs = (nrows + 1) / 2;
            for (int p = 0; p < npairs; ++p, ++g, ++acc_it) {
                FZ_T(const long long w0 = clock64(); const long long sv0 = tr_serv;)
                if (p == 0) wait_serving(&full[g % S], (g / S) & 1);
                wait_serving(&full[(g + 1) % S], ((g + 1) / S) & 1);
                FZ_T(const long long w1 = clock64(); const long long sv1 = tr_serv; tr_full += (w1 - w0) - (sv1 - sv0);)
                const uint32_t j = acc_it & 1;
                if (acc_it >= 2) wait_serving(&tempty[j], ((acc_it >> 1) - 1) & 1);
                FZ_T(const long long w2 = clock64(); const long long sv2 = tr_serv; tr_tempty += (w2 - w1) - (sv2 - sv1);)
                tc_fence_after();
                constexpr uint32_t d_hi = (uint32_t)(128 >> 4) | (1u << 14);
                constexpr uint32_t a_lo_t = (uint32_t)(L::LBO >> 4) << 16;
                constexpr uint32_t b_lo_t = (uint32_t)((COUT * 16) >> 4) << 16;
                const uint32_t b_lo0 = b_lo_t | ((w_base & 0x3FFFFu) >> 4);
                const uint32_t ones_lo = ((uint32_t)((128 * 16) >> 4) << 16) | ((smem_u32(s_ones) & 0x3FFFFu) >> 4);
                const bool lead2 = leader && !(a.debug & 128);       // debug bit 7: the second block's MMA instructions run predicated off
                for (int r = 0; r < 2; ++r) {
                    if (2 * p + r >= nrows || (a.debug & 8)) break;     // debug bit 3: no second-block MMAs (timing experiment)
                    const uint32_t d_tmem = tmem + j * (2 * COUT) + r * COUT;
                    umma_f16_if(lead2, d_tmem, desc64(ones_lo, d_hi), desc64(b_lo0 + (uint32_t)(L::WBYTES >> 4), d_hi), idesc, 0u);
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const int i = 2 * p + r + dy;
                        const int in_row = y0 - a.pad + i;
                        uint32_t row_base;
                        if (in_row < 0 || in_row >= a.H1) row_base = zero_base;
                        else row_base = ring_base + ((((g + (i >> 1) - p) % S) << 1) + (i & 1)) * L::ROWB;
                        const uint32_t a_lo0 = a_lo_t | ((row_base & 0x3FFFFu) >> 4);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int ks = 0; ks < CIN / 16; ++ks)
                                umma_f16_if(lead2, d_tmem, desc64(a_lo0 + (uint32_t)((ks * 2 * L::LBO + dx * 16) >> 4), d_hi),
                                            desc64(b_lo0 + (uint32_t)((((dy * 3 + dx) * L::CHUNKS + 2 * ks) * (COUT * 16)) >> 4), d_hi),
                                            idesc, 1u);
                    }
                }
                umma_commit_if(leader, &empty[g % S]);
                umma_commit_if(leader, &tfull[j]);
                FZ_T(tr_issue += (clock64() - w2) - (tr_serv - sv2);)
            }
            umma_commit_if(leader, &empty[g % S]);
            ++g;
        }
        // every team MMA has retired by now (the second block consumed their rows): release the loader warp's service loop
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(fz_stop) = 1u;
        FZ_T(if (blockIdx.x == 0 && lane == 0) { g_fz_trace[0] = clock64() - tr_begin; g_fz_trace[1] = tr_full; g_fz_trace[2] = tr_tempty;
                                                 g_fz_trace[3] = tr_issue; g_fz_trace[4] = tr_serv; g_fz_trace[5] = tr_nserv; g_fz_trace[6] = acc_it; })
    } else {
        // ================================ second-block epilogue (4 warps, one per TMEM lane quadrant) ================================
        const int quad = warp & 3;
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
                for (int half = 0; half < COUT / 32; ++half) {
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
                            *d0 = make_uint4(fh2u(a0[cc * 4]), fh2u(a0[cc * 4 + 1]), fh2u(a0[cc * 4 + 2]), fh2u(a0[cc * 4 + 3]));
                            if (has1) {
                                uint4* d1 = d0 + (size_t)(COUT / 8) * a.Wo;
                                *d1 = make_uint4(fh2u(a1[cc * 4]), fh2u(a1[cc * 4 + 1]), fh2u(a1[cc * 4 + 2]), fh2u(a1[cc * 4 + 3]));
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const __half2 mv = __hmax2(a0[q], a1[q]);
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, fh2u(mv), 1);
                        a0[q] = __hmax2(mv, *reinterpret_cast<const __half2*>(&o));
                    }
                    if (pool_ok && a.pool_fc != nullptr) {
                        const int row = b & 127;
                        uint8_t* base = a.pool_fc + ((((size_t)(b >> 7) * a.Hp * a.Wp) + (size_t)py * a.Wp + pxx) * 128 + row) * 128;
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int chunk = (half * 4 + cc) ^ (row & 7);
                            *reinterpret_cast<uint4*>(base + chunk * 16) =
                                make_uint4(fh2u(a0[cc * 4]), fh2u(a0[cc * 4 + 1]), fh2u(a0[cc * 4 + 2]), fh2u(a0[cc * 4 + 3]));
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

bool conv_fused_supported(int Cin, int C0, int C1, int W1, int Wo1, bool x3) {
    return Cin == 1 && C0 == FZ_C0 && C1 == FZ_C1 && W1 <= 128 && Wo1 <= 128 && !x3;   // (+ conv LeakyReLU slope in [0, 1]: checked by the tensor path)
}

int launch_conv_fused(const FusedArgs& a, int sms, cudaStream_t s) {
    static_assert(FusedSmem::TOTAL <= 227 * 1024, "conv_fused: shared memory budget");
    BCAD_REQUIRE(a.band_rows % 2 == 0 && a.bands == cdiv(a.Ho, a.band_rows), "conv_fused: bad banding");
    BCAD_REQUIRE(a.W1 <= 128 && a.Wo <= 128, "conv_fused: second-block map wider than 128");
    const int items = a.B * a.bands;
    const int grid = items < sms ? items : sms;
    if (a.plain0) {
        BCAD_CUDA_CHECK(cudaFuncSetAttribute(conv_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::TOTAL));
        conv_fused_kernel<true><<<grid, FZ_THREADS, FusedSmem::TOTAL, s>>>(a);
    } else {
        BCAD_CUDA_CHECK(cudaFuncSetAttribute(conv_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem::TOTAL));
        conv_fused_kernel<false><<<grid, FZ_THREADS, FusedSmem::TOTAL, s>>>(a);
    }
    BCAD_CUDA_CHECK(cudaGetLastError());
#ifdef FZ_TRACE
    if (a.debug & 64) {
        long long h[8];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_fz_trace, sizeof(h));
        fprintf(stderr, "fz_trace: total %lld | wait full %lld | wait tempty %lld | issue second block %lld | serve teams %lld (%lld requests) | pairs %lld\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    }
#endif
    return BCAD_OK;
}

}  // namespace bcad
