// Tiny U-Net encoder front (BASELINE config 3) on the tensor cores:
//   tiny_unet_numpy  (Classes/unet.py:61-73):  conv(1->16)+ReLU+pool -> conv(16->32)+ReLU+pool -> conv(32->64)+ReLU
//   average_pool     (Classes/ImageSegmentation.py:145-163)  -> the CNN's input (22,22,64) for a 256x256 image
// with the reference's conv2d(..., 'same') QUIRK (unet.py:19-27): every conv output is allocated at the PADDED size, its
// trailing two rows / columns stay zero.  For an input with H, W multiples of 4 the quirk decomposes cleanly:
//   * conv1: the pooled map is the true (H/2 x W/2) map plus ONE zero row / column (the pool window of the two zero rows)
//     -> conv_first_pool_kernel (CUDA cores, K = 9) writes the true part; the zero border is never materialised:
//   * conv2 sees that zero row / column exactly like its own zero padding, EXCEPT that it also produces one more output
//     row / column (index H/2: taps dy = -1 only).  The true (H/2 x W/2) outputs + ReLU + pool run on tcgen05
//     (conv_igemm_kernel<16,32>); the extra row / column -- (W/2 + H/2 + 1) pixels -- is a small CUDA-core kernel that writes the
//     border of the pooled map directly (its pool partner is a quirk-zero row, and ReLU >= 0);
//   * conv3 (conv_igemm_kernel<32,64>, tcgen05) runs on the (H/4+1 x W/4+1) pooled map; its two quirk-zero rows / columns only
//     matter to what reads it: the output kernel below treats rows / columns >= H/4+1 as zero.
// Operands are fp16, accumulation fp32 (the 16-bit mode's tolerance: 1e-2 relative to the map's scale); the fp32 CUDA-core
// route (bcad_conv_block) stays for every other shape and for fp32-grade results.  All buffers live in a handle: no
// allocation on the forward path.
#include <string.h>

#include <mutex>
#include <new>
#include <vector>

#include "../../include/bcad.h"
#include "common.cuh"
#include "sm100_kernels.h"

namespace bcad {

struct UnetFront {
    int H = 0, W = 0, max_batch = 0, device = 0, sms = 148;
    int H1 = 0, W1 = 0;            // pooled conv1 map (true part)
    int H2 = 0, W2 = 0;            // pooled conv2 map INCLUDING the border row / column: H1/2 + 1
    bool have_kernels = false;
    float* d_w1 = nullptr;         // [9][16] fp32
    float* d_b1 = nullptr;         // zeros [16]
    uint8_t* d_w2_img = nullptr;   // igemm weight image 16 -> 32 (+ zero bias tile)
    float* d_w2_f32 = nullptr;     // [9][16][32] fp32 values of the SAME fp16-rounded weights (border kernel)
    uint8_t* d_w3_img = nullptr;   // igemm weight image 32 -> 64 (+ zero bias tile)
    __half* p1 = nullptr;          // [B][H1][2][W1][8]
    __half* p2 = nullptr;          // [B][H2][4][W2][8]
    __half* a3 = nullptr;          // [B][H2][8][W2][8]
    std::mutex mu;
    cudaEvent_t done = nullptr;
    std::vector<void*> allocs;
    int64_t launches = 0;
    bool profiling = false;        // one CUDA event before every launch of the last chunk (+ one at the end)
    cudaEvent_t prof[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool prof_valid = false;
};
static const char* const kUnetStage[5] = {"unet_conv1_first_pool", "unet_conv2_igemm_tcgen05", "unet_conv2_border", "unet_conv3_igemm_tcgen05",
                                          "unet_out_avgpool"};

// conv2's extra output row (index H1: input row H1-1 under the TOP kernel row) and column (index W1: input column W1-1 under the
// LEFT kernel column), ReLU, then the pooled border of P2: row H2-1 (W2 entries) and column W2-1 (H2-1 entries).  One CTA per image.
__global__ void __launch_bounds__(256)
unet_conv2_border_kernel(const __half* __restrict__ p1, const float* __restrict__ w /*[9][16][32]*/, __half* __restrict__ p2, int H1, int W1,
                         int H2, int W2) {
    __shared__ float s_w[9 * 16 * 32];
    for (int i = threadIdx.x; i < 9 * 16 * 32; i += 256) s_w[i] = w[i];
    __syncthreads();
    const int b = blockIdx.x;
    const __half* pb = p1 + (size_t)b * H1 * 2 * W1 * 8;
    auto in = [&](int y, int x, int c) -> float {          // C8-planar [y][c/8][x][8]; zero outside the true map
        if (y < 0 || y >= H1 || x < 0 || x >= W1) return 0.f;
        return __half2float(pb[(((size_t)y * 2 + (c >> 3)) * W1 + x) * 8 + (c & 7)]);
    };
    auto row_out = [&](int x, int f) -> float {             // true conv2 output at (H1, x): only kernel row 0 meets data
        float acc = 0.f;
        for (int dx = 0; dx < 3; ++dx)
            for (int c = 0; c < 16; ++c) acc = fmaf(in(H1 - 1, x + dx - 1, c), s_w[((0 * 3 + dx) * 16 + c) * 32 + f], acc);
        return fmaxf(acc, 0.f);
    };
    auto col_out = [&](int y, int f) -> float {             // true conv2 output at (y, W1): only kernel column 0 meets data
        float acc = 0.f;
        for (int dy = 0; dy < 3; ++dy)
            for (int c = 0; c < 16; ++c) acc = fmaf(in(y + dy - 1, W1 - 1, c), s_w[((dy * 3 + 0) * 16 + c) * 32 + f], acc);
        return fmaxf(acc, 0.f);
    };
    __half* ob = p2 + (size_t)b * H2 * 4 * W2 * 8;
    const int n_row = W2, n_col = H2 - 1;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < (n_row + n_col) * 32; i += 256 * gridDim.y) {
        const int f = i & 31, pos = i >> 5;
        float v;
        int py, px;
        if (pos < n_row) {                                  // pooled row H2-1: windows {(H1, 2j), (H1, 2j+1)} over a quirk-zero row
            py = H2 - 1; px = pos;
            v = (2 * px + 1 <= W1) ? fmaxf(row_out(2 * px, f), row_out(2 * px + 1, f)) : row_out(2 * px, f);
        } else {                                            // pooled column W2-1: windows {(2i, W1), (2i+1, W1)}
            py = pos - n_row; px = W2 - 1;
            v = fmaxf(col_out(2 * py, f), col_out(2 * py + 1, f));
        }
        ob[(((size_t)py * 4 + (f >> 3)) * W2 + px) * 8 + (f & 7)] = __float2half_rn(v);
    }
}

// conv3 map (C8-planar fp16, true size h x w) -> fp32 NHWC output: pool = 0: the reference's `bn` tensor (h+2, w+2, C) with its two
// quirk-zero rows / columns; pool > 0: average_pool(pool) of that tensor (floor dims; the zero rows count in the mean).
// One thread per (output pixel, channel octet): 16-byte loads of 8 channels per source pixel, two 16-byte stores.
__global__ void __launch_bounds__(256)
unet_out_kernel(const __half* __restrict__ a3, float* __restrict__ out, int h, int w, int C, int pool, int Hn, int Wn, size_t total8) {
    const int oct = C / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
        const int o = (int)(i % oct);
        size_t r = i / oct;
        const int ox = (int)(r % Wn);
        r /= Wn;
        const int oy = (int)(r % Hn);
        const size_t b = r / Hn;
        const uint4* ab = reinterpret_cast<const uint4*>(a3) + b * (size_t)h * oct * w;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int p = pool > 0 ? pool : 1;
        for (int u = 0; u < p; ++u)
            for (int v = 0; v < p; ++v) {
                const int y = oy * p + u, x = ox * p + v;
                if (y >= h || x >= w) continue;               // the quirk's zero rows / columns
                const uint4 q = __ldg(ab + ((size_t)y * oct + o) * w + x);
                const float2 f0 = unpack_f16(q.x), f1 = unpack_f16(q.y), f2 = unpack_f16(q.z), f3 = unpack_f16(q.w);
                acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y; acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
            }
        const float inv = 1.f / (float)(p * p);
        float4* dst = reinterpret_cast<float4*>(out + ((b * Hn + oy) * (size_t)Wn + ox) * C + o * 8);
        dst[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
        dst[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    }
}

static int ualloc(UnetFront* u, void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return BCAD_ERR_NOMEM;
    }
    u->allocs.push_back(*p);
    return BCAD_OK;
}

// rows per work item of a persistent conv kernel: the even band height that minimises the busiest SM's row count
// (waves x band rows), e.g. 256 images of a 65-row map on 148 SMs: 22-row bands (6 waves x 22) instead of 64 + 1 (2 waves x 64)
static int pick_band_rows(int n, int Ho, int xsegs, int sms) {
    int best = 2;
    long long best_cost = -1;
    for (int br = 2; br <= 64; br += 2) {
        const int bands = cdiv(Ho, br);
        const long long items = (long long)n * bands * xsegs;
        const long long cost = ((items + sms - 1) / sms) * (br + 2);      // + 2: the halo rows every band re-reads
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && br > best)) { best_cost = cost; best = br; }
    }
    if (best > Ho) best = cdiv(Ho, 2) * 2;
    return best;
}

static uint16_t f2h_bits(float f) {
    const __half h = __float2half_rn(f);
    uint16_t q;
    memcpy(&q, &h, 2);
    return q;
}

// igemm weight image: [tap*(Cin/8)+chunk][cout][8] fp16, then the bias tile [2][cout][8] (zeros: the U-Net convs have no bias)
static std::vector<uint16_t> igemm_image(const float* k /*(3,3,Cin,F)*/, int Cin, int F) {
    const int chunks = Cin / 8;
    std::vector<uint16_t> img((size_t)9 * chunks * F * 8 + (size_t)2 * F * 8, 0);
    for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < Cin; ++c)
            for (int f = 0; f < F; ++f)
                img[(((size_t)tap * chunks + (c >> 3)) * F + f) * 8 + (c & 7)] = f2h_bits(k[((size_t)tap * Cin + c) * F + f]);
    return img;
}

}  // namespace bcad

using namespace bcad;

extern "C" {

int bcad_unet_create(int H, int W, int max_batch, int device, bcad_unet** out) {
    BCAD_REQUIRE(out != nullptr, "unet_create: null argument");
    *out = nullptr;
    BCAD_REQUIRE(H >= 8 && W >= 8 && H % 4 == 0 && W % 4 == 0,
                 "unet_create: the tensor-core front takes single-channel images with H and W multiples of 4 (got %dx%d); use "
                 "bcad_conv_block for other shapes", H, W);
    BCAD_REQUIRE(max_batch >= 1 && max_batch <= 65535, "unet_create: max_batch %d out of range", max_batch);
    int ndev = 0;
    BCAD_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    BCAD_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d visible)", device, ndev);
    UnetFront* u = new (std::nothrow) UnetFront();
    if (!u) { set_error("out of host memory"); return BCAD_ERR_NOMEM; }
    u->H = H; u->W = W; u->max_batch = max_batch; u->device = device;
    u->H1 = H / 2; u->W1 = W / 2; u->H2 = u->H1 / 2 + 1; u->W2 = u->W1 / 2 + 1;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaDeviceGetAttribute(&u->sms, cudaDevAttrMultiProcessorCount, device);
    const size_t mb = (size_t)max_batch;
    int rc = ualloc(u, (void**)&u->d_w1, 9 * 16 * 4);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->d_b1, 16 * 4);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->d_w2_img, ((size_t)9 * 2 * 32 * 8 + 2 * 32 * 8) * 2);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->d_w2_f32, (size_t)9 * 16 * 32 * 4);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->d_w3_img, ((size_t)9 * 4 * 64 * 8 + 2 * 64 * 8) * 2);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->p1, mb * u->H1 * u->W1 * 16 * 2);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->p2, mb * u->H2 * u->W2 * 32 * 2);
    if (rc == BCAD_OK) rc = ualloc(u, (void**)&u->a3, mb * u->H2 * u->W2 * 64 * 2);
    if (rc == BCAD_OK && cudaMemset(u->d_b1, 0, 16 * 4) != cudaSuccess) rc = BCAD_ERR_CUDA;
    if (rc == BCAD_OK && cudaEventCreateWithFlags(&u->done, cudaEventDisableTiming) != cudaSuccess) rc = BCAD_ERR_CUDA;
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != BCAD_OK) {
        bcad_unet_destroy(reinterpret_cast<bcad_unet*>(u));
        return rc;
    }
    *out = reinterpret_cast<bcad_unet*>(u);
    return BCAD_OK;
}

void bcad_unet_destroy(bcad_unet* uu) {
    if (!uu) return;
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(u->device);
    cudaDeviceSynchronize();
    for (void* p : u->allocs) cudaFree(p);
    if (u->done) cudaEventDestroy(u->done);
    for (cudaEvent_t e : u->prof) if (e) cudaEventDestroy(e);
    if (prev >= 0) cudaSetDevice(prev);
    delete u;
}

int bcad_unet_set_kernels(bcad_unet* uu, const float* k1, const float* k2, const float* k3) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    BCAD_REQUIRE(u && k1 && k2 && k3, "unet_set_kernels: null argument");
    std::lock_guard<std::mutex> lock(u->mu);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(u->device);
    cudaDeviceSynchronize();
    const std::vector<uint16_t> img2 = igemm_image(k2, 16, 32), img3 = igemm_image(k3, 32, 64);
    std::vector<float> w2((size_t)9 * 16 * 32);
    for (size_t i = 0; i < w2.size(); ++i) w2[i] = __half2float(__float2half_rn(k2[i]));     // the values the tensor core sees
    cudaError_t e = cudaMemcpy(u->d_w1, k1, 9 * 16 * 4, cudaMemcpyHostToDevice);              // (3,3,1,16) == [9][16]
    if (e == cudaSuccess) e = cudaMemcpy(u->d_w2_img, img2.data(), img2.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(u->d_w2_f32, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(u->d_w3_img, img3.data(), img3.size() * 2, cudaMemcpyHostToDevice);
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) {
        set_error("unet_set_kernels: %s", cudaGetErrorString(e));
        return BCAD_ERR_CUDA;
    }
    u->have_kernels = true;
    return BCAD_OK;
}

int bcad_unet_out_shape(bcad_unet* uu, int avg_pool, int* out_h, int* out_w, int* out_c) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    BCAD_REQUIRE(u && out_h && out_w && out_c && avg_pool >= 0, "unet_out_shape: bad argument");
    const int bh = u->H2 + 2, bw = u->W2 + 2;               // the reference's bn tensor: padded-size output of conv3
    *out_h = avg_pool > 0 ? bh / avg_pool : bh;
    *out_w = avg_pool > 0 ? bw / avg_pool : bw;
    *out_c = 64;
    return BCAD_OK;
}

int bcad_unet_forward(bcad_unet* uu, const float* x_dev, int B, int avg_pool, float* out_dev, void* stream) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    BCAD_REQUIRE(u && x_dev && out_dev, "unet_forward: null argument");
    BCAD_REQUIRE(B >= 1 && avg_pool >= 0, "unet_forward: bad batch / pool size");
    if (!u->have_kernels) { set_error("unet_forward: call bcad_unet_set_kernels first"); return BCAD_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != u->device) cudaSetDevice(u->device);
    std::lock_guard<std::mutex> lock(u->mu);
    int oh = 0, ow = 0, oc = 0;
    bcad_unet_out_shape(uu, avg_pool, &oh, &ow, &oc);
    int rc = BCAD_OK;
    if (cudaStreamWaitEvent(s, u->done, 0) != cudaSuccess) rc = BCAD_ERR_CUDA;
    for (int b0 = 0; b0 < B && rc == BCAD_OK; b0 += u->max_batch) {
        const int n = std::min(u->max_batch, B - b0);
        const float* x = x_dev + (size_t)b0 * u->H * u->W;
        auto mark = [&](int i) { if (u->profiling && u->prof[i]) cudaEventRecord(u->prof[i], s); };
        mark(0);
        // conv1 (1 -> 16) + ReLU + pool: CUDA cores (K = 9), true pooled map H1 x W1
        rc = launch_conv_first_pool(x, u->d_w1, u->d_b1, u->p1, n, u->H, u->W, 1, 16, 0.f, s);
        mark(1);
        // conv2 (16 -> 32) + ReLU + pool on tcgen05, pooled output written into the (H2 x W2)-pitched map
        if (rc == BCAD_OK) {
            IgemmArgs a;
            a.in = u->p1; a.w_img = u->d_w2_img; a.act = nullptr; a.pool_fc = nullptr; a.pool_c8 = u->p2;
            a.B = n; a.H = u->H1; a.W = u->W1; a.Ho = u->H1; a.Wo = u->W1; a.Hp = u->H2; a.Wp = u->W2; a.pad = 1;
            a.xsegs = cdiv(a.Wo, 128);
            a.band_rows = pick_band_rows(n, a.Ho, a.xsegs, u->sms);
            a.bands = cdiv(a.Ho, a.band_rows);
            a.alpha = 0.f; a.debug = 0;
            rc = launch_conv_igemm(a, 16, 32, false, u->sms, s);
        }
        mark(2);
        if (rc == BCAD_OK) {
            unet_conv2_border_kernel<<<dim3(n, 8), 256, 0, s>>>(u->p1, u->d_w2_f32, u->p2, u->H1, u->W1, u->H2, u->W2);
            if (cudaGetLastError() != cudaSuccess) { set_error("unet border kernel launch failed"); rc = BCAD_ERR_CUDA; }
        }
        mark(3);
        // conv3 (32 -> 64) + ReLU on tcgen05 over the (H2 x W2) map
        if (rc == BCAD_OK) {
            IgemmArgs a;
            a.in = u->p2; a.w_img = u->d_w3_img; a.act = u->a3; a.pool_fc = nullptr; a.pool_c8 = nullptr;
            a.B = n; a.H = u->H2; a.W = u->W2; a.Ho = u->H2; a.Wo = u->W2; a.Hp = u->H2 / 2; a.Wp = u->W2 / 2; a.pad = 1;
            a.xsegs = cdiv(a.Wo, 128);
            a.band_rows = pick_band_rows(n, a.Ho, a.xsegs, u->sms);
            a.bands = cdiv(a.Ho, a.band_rows);
            a.alpha = 0.f; a.debug = 0;
            rc = launch_conv_igemm(a, 32, 64, false, u->sms, s);
        }
        mark(4);
        if (rc == BCAD_OK) {
            const size_t total8 = (size_t)n * oh * ow * 8;
            const int grid = (int)std::min<size_t>((total8 + 255) / 256, (size_t)u->sms * 16);
            unet_out_kernel<<<grid, 256, 0, s>>>(u->a3, out_dev + (size_t)b0 * oh * ow * 64, u->H2, u->W2, 64, avg_pool, oh, ow, total8);
            if (cudaGetLastError() != cudaSuccess) { set_error("unet output kernel launch failed"); rc = BCAD_ERR_CUDA; }
        }
        mark(5);
        u->prof_valid = u->profiling;
        u->launches += 5;
    }
    if (rc == BCAD_OK && cudaEventRecord(u->done, s) != cudaSuccess) rc = BCAD_ERR_CUDA;
    if (prev >= 0 && prev != u->device) cudaSetDevice(prev);
    return rc;
}

int bcad_unet_set_profiling(bcad_unet* uu, int on) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    BCAD_REQUIRE(u, "null handle");
    std::lock_guard<std::mutex> lock(u->mu);
    if (on)
        for (cudaEvent_t& e : u->prof)
            if (e == nullptr) BCAD_CUDA_CHECK(cudaEventCreate(&e));
    u->profiling = on != 0;
    u->prof_valid = false;
    return BCAD_OK;
}

int bcad_unet_profile_get(bcad_unet* uu, int i, char* name_buf, int name_cap, float* ms) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    BCAD_REQUIRE(u && ms && i >= 0 && i < 5, "unet_profile_get: bad argument");
    if (!u->prof_valid) { set_error("unet_profile_get: no profiled forward (bcad_unet_set_profiling(u, 1) first)"); return BCAD_ERR_STATE; }
    BCAD_CUDA_CHECK(cudaEventSynchronize(u->prof[i + 1]));
    BCAD_CUDA_CHECK(cudaEventElapsedTime(ms, u->prof[i], u->prof[i + 1]));
    if (name_buf && name_cap > 0) {
        strncpy(name_buf, kUnetStage[i], name_cap - 1);
        name_buf[name_cap - 1] = 0;
    }
    return BCAD_OK;
}

int64_t bcad_unet_launch_count(bcad_unet* uu) {
    UnetFront* u = reinterpret_cast<UnetFront*>(uu);
    return u ? u->launches : -1;
}

}  // extern "C"
