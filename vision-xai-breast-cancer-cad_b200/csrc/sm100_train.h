// Launchers of the training step's tensor-core kernels (sm100_train.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bcad {

struct TcConvArgs {
    const float* x;        // fp32 NHWC [B][H][W][Cin]
    const uint8_t* w_img;  // launch_pack_w_x3 image: hi / lo planes [plane][tap][Cin/8][Cout][8], fp16 -- or bf16 when a_bf16 (same type as the input)
    const float* bias;     // fp32 [Cout] or nullptr
    float* y;              // fp32 NHWC [B][Ho][Wo][Cout]
    int B, H, W, Ho, Wo, pad;
    float alpha;           // LeakyReLU slope (1 = identity)
    int a_bf16;            // the input splits into bf16 pairs (gradients: fp32's range) instead of fp16 pairs
    int bands = 0, band_rows = 0;   // set by the launcher
};
bool conv3x3_x3_supported(int Cin, int Cout, int k, int W, int Wo, int pad);
size_t conv3x3_x3_weight_bytes(int Cin, int Cout);
// w: the fp32 path's packed filters [9][Cin][CoutPad]
int launch_pack_w_x3(const float* w, uint8_t* img, int Cin, int Cout, int CoutPad, bool bf16, cudaStream_t s);
int launch_conv3x3_x3(const TcConvArgs& a, int Cin, int Cout, int sms, cudaStream_t s);
struct TcWgradArgs {
    const float* x;        // the block's input, fp32 NHWC [B][H][W][32]
    const float* dy;       // gradient w.r.t. the block's pre-activation, fp32 NHWC [B][Ho][Wo][64]
    float* partials;       // [grid][12][128][32] fp32 scratch (wgrad3x3_x3_partial_floats)
    int B, H, W, Ho, Wo, pad;
};
bool wgrad3x3_x3_supported(int Cin, int Cout, int k, int W, int Wo, int pad);
size_t wgrad3x3_x3_partial_floats(int sms);
// dw: [9][32][CoutPad] fp32 (the fp32 path's gradient layout), overwritten
int launch_wgrad3x3_x3(const TcWgradArgs& a, float* dw, int CoutPad, int sms, cudaStream_t s);
struct DenseBwdArgs {
    const float* dz;       // [B][units] fp32: gradient w.r.t. the layer's pre-activation
    const float* src;      // mode 0: the layer's input p [B][flat]; mode 1: its weights W [units][flat]
    float* out;            // mode 0: dW [units][flat]; mode 1: g [B][flat]
    int B, units;
    long long flat;
};
bool dense_bwd_x3_supported(int B, int units, long long flat);
int launch_dense_bwd_x3(const DenseBwdArgs& a, int mode, int sms, cudaStream_t s);
struct DenseFwdArgs {
    const float* p;        // the layer's input [B][flat] fp32
    const float* w;        // its weights [units][flat] fp32
    float* partials;       // split-K partials [dense_fwd_x3_splits(flat, sms)][B][units]
    int B, units;
    long long flat;
};
bool dense_fwd_x3_supported(int B, int units, long long flat);
int dense_fwd_x3_splits(long long flat, int sms);
int launch_dense_fwd_x3(const DenseFwdArgs& a, int sms, cudaStream_t s);
// dz = (pool switch ? g : 0) * (y > 0 ? 1 : alpha): max-pool backward + LeakyReLU' in one pass (NHWC fp32, C % 4 == 0)
int launch_unpool_mask(const float* g, const float* y, float* dz, int B, int Ho, int Wo, int C, int first_only, float alpha, cudaStream_t s);
// first conv block with ONE input channel and 32 filters, forward: y (post-activation) and its 2x2 max pool in one pass; w = packed [9][1][32]
int launch_conv0_fwd_fused(const float* x, const float* w, const float* bias, float* y, float* p, int B, int H, int W, int Ho, int Wo, int pad, float alpha,
                           int sms, cudaStream_t s);
// first conv block with ONE input channel and 32 filters: pool backward + LeakyReLU' + weight / bias gradients without writing dz
int conv0_bwd_fused_parts(int sms);
int launch_conv0_bwd_fused(const float* g, const float* y, const float* x, float* scratch, float* dw, float* db, int B, int H, int W, int Ho, int Wo,
                           int pad, int first_only, float alpha, int sms, cudaStream_t s);
// out[64] = column sums of A[K][64]; scratch: 4096 * 64 floats
int launch_colsum64(const float* A, float* scratch, float* out, size_t K, cudaStream_t s);
int launch_maxpool2x2_nhwc(const float* y, float* p, int B, int Ho, int Wo, int C, cudaStream_t s);

}  // namespace bcad
