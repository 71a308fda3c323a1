// Training-step kernels, fp32 CUDA cores: loss + top gradient, weight / bias gradients of the
// dense and conv layers, SGD with per-tensor norm clipping (Classes/CNNModel.py:217-222, 372-394) and Adam
// (torch.optim.Adam as used at ADCNNM.py:88).  Correctness-first: these reuse the fp32 path's layouts (NHWC
// activations, conv weights packed [tap][Cin][CoutPad], dense (out,in)) so the optimiser updates the very buffers the
// forward kernels read.
#include "../../include/bcad.h"
#include "common.cuh"
#include "kernels.h"

namespace bcad {

// loss[b] = -log(max(p_label, 1e-12)) (Classes/CNNModel.py:360-367); dz_last = (probs - onehot) / B (:299, mean over batch)
__global__ void ce_loss_topgrad_kernel(const float* __restrict__ probs, const int32_t* __restrict__ labels,
                                       float* __restrict__ loss, float* __restrict__ dz, int B, int nc, float inv_b) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int y = labels[b];
    for (int c = 0; c < nc; ++c) {
        const float p = probs[(size_t)b * nc + c];
        dz[(size_t)b * nc + c] = (p - (c == y ? 1.f : 0.f)) * inv_b;
        if (c == y) loss[b] = -logf(fmaxf(p, 1e-12f));
    }
}

int launch_ce_loss_topgrad(const float* probs, const int32_t* labels, float* loss, float* dz, int B, int nc, cudaStream_t s) {
    ce_loss_topgrad_kernel<<<cdiv(B, 128), 128, 0, s>>>(probs, labels, loss, dz, B, nc, 1.0f / (float)B);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

__global__ void leaky_from_z_kernel(const float* __restrict__ z, float* __restrict__ h, float alpha, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) h[i] = leaky(z[i], alpha);
}

int launch_leaky_from_z(const float* z, float* h, float alpha, int64_t n, cudaStream_t s) {
    const int blocks = (int)min((int64_t)148 * 8, (n + 255) / 256);
    leaky_from_z_kernel<<<blocks, 256, 0, s>>>(z, h, alpha, (size_t)n);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// h[b][u] (*)= mask[b][off + u]: dropout multipliers (0 or 1/(1-p)) of one hidden layer, rows of length ld in `mask`
__global__ void mul_mask_kernel(float* __restrict__ h, const float* __restrict__ mask, int units, int ld, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / units;
        h[i] *= mask[b * ld + (i - b * units)];
    }
}

int launch_mul_mask(float* h, const float* mask, int B, int units, int ld, cudaStream_t s) {
    const int64_t n = (int64_t)B * units;
    const int blocks = (int)min((int64_t)148 * 8, (n + 255) / 256);
    mul_mask_kernel<<<blocks, 256, 0, s>>>(h, mask, units, ld, (size_t)n);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// C[M][N] = sum_k A[k][M] * B[k][N]  (dense weight gradient dW = dz^T . input; K = batch)
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       float* __restrict__ C, int M, int N, int K) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64 + ty * 4, n0 = blockIdx.x * 64 + tx * 4;
    float acc[4][4] = {};
    const bool mv = (M % 4 == 0) && (m0 + 4 <= M), nv = (N % 4 == 0) && (n0 + 4 <= N);
    for (int k = 0; k < K; ++k) {
        float a[4], b[4];
        if (mv) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(A + (size_t)k * M + m0));
            a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = (m0 + i < M) ? __ldg(A + (size_t)k * M + m0 + i) : 0.f;
        }
        if (nv) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)k * N + n0));
            b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = (n0 + j < N) ? __ldg(Bm + (size_t)k * N + n0 + j) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (m0 + i >= M) continue;
        float* dst = C + (size_t)(m0 + i) * N + n0;
        if (nv) *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        else
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + j < N) dst[j] = acc[i][j];
    }
}

int launch_sgemm_tn(const float* A, const float* Bm, float* C, int M, int N, int K, cudaStream_t s) {
    dim3 grid(cdiv(N, 64), cdiv(M, 64));
    BCAD_REQUIRE(grid.y <= 65535, "sgemm_tn: M too large");
    sgemm_tn_kernel<<<grid, 256, 0, s>>>(A, Bm, C, M, N, K);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// out[n] = sum_k A[k][n]   (bias gradients)
__global__ void colsum_kernel(const float* __restrict__ A, float* __restrict__ out, int K, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += A[(size_t)k * N + n];
    out[n] = acc;
}

int launch_colsum(const float* A, float* out, int K, int N, cudaStream_t s) {
    colsum_kernel<<<cdiv(N, 128), 128, 0, s>>>(A, out, K, N);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// =====================================================================================================
// conv weight gradient: dF[tap][c][f] = sum_{b,y,x} dz[b,y,x,f] * in[b, y+ky-pad, x+kx-pad, c]   (explainability.py:58-59),
// db[f] = sum dz.  One CTA per (image, band of rows); thread = (8 filters, one input channel), 9 x 8 accumulators;
// per-CTA partials reduced in a fixed order afterwards (deterministic).
// =====================================================================================================
template <int K>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ in, float* __restrict__ part_w, float* __restrict__ part_b,
                  int H, int W, int Cin, int Cout, int CoutPad, int pad, int Ho, int Wo, int band_rows, int bands) {
    extern __shared__ float sm[];
    const int NG = CoutPad / 8, CL = 256 / NG;                 // filter groups, input-channel lanes
    float* s_dz = sm;                                           // [32 px][CoutPad]
    float* s_in = sm + 32 * CoutPad;                            // [K rows][32+K-1 px][CL]
    const int tid = threadIdx.x, cg = tid % NG, cl = tid / NG;
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int y0 = band * band_rows, y1 = min(Ho, y0 + band_rows);
    const size_t wsz = (size_t)K * K * Cin * CoutPad;
    float* pw = part_w + (size_t)blockIdx.x * wsz;
    float accb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) accb[j] = 0.f;
    for (int c0 = 0; c0 < Cin; c0 += CL) {
        const int c = c0 + cl;
        float acc[K * K][8];
#pragma unroll
        for (int t = 0; t < K * K; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
        for (int y = y0; y < y1; ++y)
            for (int x0 = 0; x0 < Wo; x0 += 32) {
                __syncthreads();
                for (int i = tid; i < 32 * CoutPad; i += 256) {
                    const int f = i % CoutPad, px = i / CoutPad;
                    s_dz[i] = (x0 + px < Wo && f < Cout) ? dz[(((size_t)b * Ho + y) * Wo + x0 + px) * Cout + f] : 0.f;
                }
                for (int i = tid; i < K * (32 + K - 1) * CL; i += 256) {
                    const int ci = i % CL, rest = i / CL;
                    const int px = rest % (32 + K - 1), r = rest / (32 + K - 1);
                    const int iy = y + r - pad, ix = x0 + px - pad;
                    s_in[i] = (c0 + ci < Cin && iy >= 0 && iy < H && ix >= 0 && ix < W)
                                  ? in[(((size_t)b * H + iy) * W + ix) * Cin + c0 + ci] : 0.f;
                }
                __syncthreads();
                for (int px = 0; px < 32; ++px) {
                    const float4 d0 = *reinterpret_cast<const float4*>(&s_dz[px * CoutPad + cg * 8]);
                    const float4 d1 = *reinterpret_cast<const float4*>(&s_dz[px * CoutPad + cg * 8 + 4]);
                    const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                    if (c0 == 0 && cl == 0)
#pragma unroll
                        for (int j = 0; j < 8; ++j) accb[j] += dv[j];
#pragma unroll
                    for (int ky = 0; ky < K; ++ky)
#pragma unroll
                        for (int kx = 0; kx < K; ++kx) {
                            const float xv = s_in[((ky * (32 + K - 1)) + px + kx) * CL + cl];
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[ky * K + kx][j] = fmaf(xv, dv[j], acc[ky * K + kx][j]);
                        }
                }
            }
        if (c < Cin)
#pragma unroll
            for (int t = 0; t < K * K; ++t) {
                float* dst = pw + ((size_t)t * Cin + c) * CoutPad + cg * 8;
                *reinterpret_cast<float4*>(dst) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
            }
    }
    if (cl == 0) {
        float* dst = part_b + (size_t)blockIdx.x * CoutPad + cg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = accb[j];
    }
}


// First-layer variant (Cin <= 4): with so few input channels the (filters x channels) thread map above leaves most of the
// CTA idle, so here thread = (8 filters, one PIXEL lane): dz is read straight from global (the NG threads of a pixel cover
// its CoutPad floats contiguously), the K*K input taps come through L1, and the pixel lanes are summed at the end --
// shuffles inside a warp, then a fixed-order pass over the 8 warps in shared memory (deterministic).
template <int K>
__global__ void __launch_bounds__(256)
conv_wgrad_smallcin_kernel(const float* __restrict__ dz, const float* __restrict__ in, float* __restrict__ part_w,
                           float* __restrict__ part_b, int H, int W, int Cin, int Cout, int CoutPad, int pad, int Ho, int Wo,
                           int band_rows, int bands) {
    __shared__ float red[8][8][33];                             // [warp][filter in group][cg (<= 32)], +1 pad; one tap at a time
    const int NG = CoutPad / 8, PL = 256 / NG;                  // filter groups (<= 32), pixel lanes
    const int tid = threadIdx.x, cg = tid % NG, pl = tid / NG, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / bands, band = blockIdx.x % bands;
    const int y0 = band * band_rows, y1 = min(Ho, y0 + band_rows);
    const size_t wsz = (size_t)K * K * Cin * CoutPad;
    float* pw = part_w + (size_t)blockIdx.x * wsz;
    const bool vec = (Cout % 8 == 0);
    for (int c = 0; c < Cin; ++c) {
        float acc[K * K + 1][8];                                // last row: the bias gradient (c == 0 only)
#pragma unroll
        for (int t = 0; t <= K * K; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
        for (int y = y0; y < y1; ++y)
            for (int x = pl; x < Wo; x += PL) {
                float dv[8];
                const float* dp = dz + (((size_t)b * Ho + y) * Wo + x) * Cout + cg * 8;
                if (vec && cg * 8 < Cout) {
                    const float4 d0 = __ldg(reinterpret_cast<const float4*>(dp)), d1 = __ldg(reinterpret_cast<const float4*>(dp) + 1);
                    dv[0] = d0.x; dv[1] = d0.y; dv[2] = d0.z; dv[3] = d0.w; dv[4] = d1.x; dv[5] = d1.y; dv[6] = d1.z; dv[7] = d1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) dv[j] = (cg * 8 + j < Cout) ? __ldg(dp + j) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[K * K][j] += dv[j];
#pragma unroll
                for (int ky = 0; ky < K; ++ky)
#pragma unroll
                    for (int kx = 0; kx < K; ++kx) {
                        const int iy = y + ky - pad, ix = x + kx - pad;
                        const float xv = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(in + (((size_t)b * H + iy) * W + ix) * Cin + c) : 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[ky * K + kx][j] = fmaf(xv, dv[j], acc[ky * K + kx][j]);
                    }
            }
        // sum the pixel lanes: lanes of a warp that share cg differ by multiples of NG
        const int nval = (c == 0) ? K * K + 1 : K * K;
#pragma unroll
        for (int t = 0; t <= K * K; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = acc[t][j];
                for (int o = NG; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[t][j] = v;
            }
#pragma unroll
        for (int t = 0; t <= K * K; ++t) {
            if (t >= nval) break;
            __syncthreads();
            if (lane < NG)
#pragma unroll
                for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[t][j];
            __syncthreads();
            if (tid < 8 * NG) {
                const int g = tid % NG, j = tid / NG;
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[w][j][g];
                if (t < K * K) pw[((size_t)t * Cin + c) * CoutPad + g * 8 + j] = v;
                else part_b[(size_t)blockIdx.x * CoutPad + g * 8 + j] = v;
            }
        }
    }
}

// out[i] = sum_p part[p][i]  (fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int nparts, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int p = 0; p < nparts; ++p) v += part[(size_t)p * n + i];
        out[i] = v;
    }
}

// rows per CTA band such that B * bands <= max_ctas (bounds the partial buffer)
int conv_wgrad_band_rows(int B, int Ho, int max_ctas) {
    int rows = 8;
    while ((int64_t)B * cdiv(Ho, rows) > max_ctas && rows < Ho) rows *= 2;
    return rows;
}

int launch_conv_wgrad(const float* dz, const float* in, float* part_w, float* part_b, float* dw, float* db, int B, int H, int W,
                      int Cin, int Cout, int CoutPad, int k, int pad, int Ho, int Wo, int band_rows, cudaStream_t s) {
    BCAD_REQUIRE(k >= 1 && k <= 3, "training supports conv ksize 1..3 (got %d)", k);
    BCAD_REQUIRE(CoutPad == 32 || CoutPad == 64 || CoutPad == 128 || CoutPad == 256,
                 "training supports conv layers with up to 32/64/128/256 (padded) filters, got %d", CoutPad);
    const int bands = cdiv(Ho, band_rows);
    const int NG = CoutPad / 8, CL = 256 / NG;
    const int ncta = B * bands;
    const size_t smem = (size_t)(32 * CoutPad + k * (32 + k - 1) * CL) * sizeof(float);
    if (Cin <= 4) {                                           // first layer: pixel-parallel variant
        if (k == 1) conv_wgrad_smallcin_kernel<1><<<ncta, 256, 0, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
        else if (k == 2) conv_wgrad_smallcin_kernel<2><<<ncta, 256, 0, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
        else conv_wgrad_smallcin_kernel<3><<<ncta, 256, 0, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
    } else if (k == 1) conv_wgrad_kernel<1><<<ncta, 256, smem, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
    else if (k == 2) conv_wgrad_kernel<2><<<ncta, 256, smem, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
    else conv_wgrad_kernel<3><<<ncta, 256, smem, s>>>(dz, in, part_w, part_b, H, W, Cin, Cout, CoutPad, pad, Ho, Wo, band_rows, bands);
    BCAD_CUDA_CHECK(cudaGetLastError());
    const size_t wsz = (size_t)k * k * Cin * CoutPad;
    reduce_partials_kernel<<<(int)min((size_t)1024, (wsz + 255) / 256), 256, 0, s>>>(part_w, dw, ncta, wsz);
    reduce_partials_kernel<<<1, 256, 0, s>>>(part_b, db, ncta, (size_t)CoutPad);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// dgrad weights from the packed forward weights: dk[(t2*Cout + f)*CinPad + c] = w[(t*Cin + c)*CoutPad + f], t2 = flipped tap
__global__ void repack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ dk, int k, int Cin, int Cout, int CoutPad, int CinPad) {
    const int total = k * k * Cout * CinPad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % CinPad, rest = i / CinPad;
        const int f = rest % Cout, t2 = rest / Cout;
        const int ky = k - 1 - t2 / k, kx = k - 1 - t2 % k;
        dk[i] = (c < Cin) ? w[((size_t)(ky * k + kx) * Cin + c) * CoutPad + f] : 0.f;
    }
}

int launch_repack_dgrad(const float* w, float* dk, int k, int Cin, int Cout, int CoutPad, int CinPad, cudaStream_t s) {
    const int total = k * k * Cout * CinPad;
    repack_dgrad_kernel<<<cdiv(total, 256), 256, 0, s>>>(w, dk, k, Cin, Cout, CoutPad, CinPad);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// =====================================================================================================
// optimisers.  One CTA per tensor for the norm (tensors: every W and every b, as the reference clips them separately).
// =====================================================================================================
__global__ void __launch_bounds__(1024) l2norm_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)g[i] * (double)g[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
        *out = (float)sqrt(t);
    }
}

// w -= lr * g * scale, scale = max_norm / (norm + 1e-6) if norm > max_norm else 1   (Classes/CNNModel.py:217-222, 385-394)
__global__ void sgd_clip_update_kernel(float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ norm,
                                       float lr, float max_norm, size_t n) {
    const float nr = *norm;
    const float scale = (max_norm > 0.f && nr > max_norm) ? max_norm / (nr + 1e-6f) : 1.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        w[i] -= lr * (g[i] * scale);
}

// torch.optim.Adam (no weight decay, no amsgrad): m,v EMA, bias correction, w -= lr * mhat / (sqrt(vhat) + eps)
__global__ void adam_update_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m1,
                                   float* __restrict__ m2, float lr, float b1, float b2, float eps, float bc1, float bc2, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float a = b1 * m1[i] + (1.f - b1) * gi;
        const float v = b2 * m2[i] + (1.f - b2) * gi * gi;
        m1[i] = a;
        m2[i] = v;
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        w[i] -= (lr / bc1) * (a / denom);
    }
}

int launch_l2norm(const float* g, size_t n, float* out, cudaStream_t s) {
    l2norm_kernel<<<1, 1024, 0, s>>>(g, n, out);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

int launch_sgd_clip_update(float* w, const float* g, const float* norm, float lr, float max_norm, size_t n, cudaStream_t s) {
    const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
    sgd_clip_update_kernel<<<blocks, 256, 0, s>>>(w, g, norm, lr, max_norm, n);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

// the same update on 4 elements per thread (16-byte loads / stores; identical arithmetic per element): the big tensors are pure HBM traffic
__global__ void __launch_bounds__(256) adam_update_vec4_kernel(float4* __restrict__ w, const float4* __restrict__ g, float4* __restrict__ m1,
                                                               float4* __restrict__ m2, float lr, float b1, float b2, float eps, float bc1, float bc2, size_t n4) {
    auto one = [&](float& wi, float gi, float& mi, float& vi) {
        const float a = b1 * mi + (1.f - b1) * gi;
        const float v = b2 * vi + (1.f - b2) * gi * gi;
        mi = a;
        vi = v;
        const float denom = sqrtf(v) / sqrtf(bc2) + eps;
        wi -= (lr / bc1) * (a / denom);
    };
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 gi = g[i];
        float4 wi = w[i], mi = m1[i], vi = m2[i];
        one(wi.x, gi.x, mi.x, vi.x);
        one(wi.y, gi.y, mi.y, vi.y);
        one(wi.z, gi.z, mi.z, vi.z);
        one(wi.w, gi.w, mi.w, vi.w);
        m1[i] = mi;
        m2[i] = vi;
        w[i] = wi;
    }
}

int launch_adam_update(float* w, const float* g, float* m1, float* m2, float lr, float b1, float b2, float eps, int step, size_t n,
                       cudaStream_t s) {
    const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
    const bool vec = n % 4 == 0 && n >= 4096 && (((uintptr_t)w | (uintptr_t)g | (uintptr_t)m1 | (uintptr_t)m2) & 15) == 0;
    if (vec) {
        const int blocks4 = (int)min((size_t)148 * 16, (n / 4 + 255) / 256);
        adam_update_vec4_kernel<<<blocks4, 256, 0, s>>>(reinterpret_cast<float4*>(w), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m1),
                                                        reinterpret_cast<float4*>(m2), lr, b1, b2, eps, bc1, bc2, n / 4);
        BCAD_CUDA_CHECK(cudaGetLastError());
        return BCAD_OK;
    }
    const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
    adam_update_kernel<<<blocks, 256, 0, s>>>(w, g, m1, m2, lr, b1, b2, eps, bc1, bc2, n);
    BCAD_CUDA_CHECK(cudaGetLastError());
    return BCAD_OK;
}

}  // namespace bcad
