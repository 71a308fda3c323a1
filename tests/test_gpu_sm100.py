"""Pins the tcgen05 conventions the tensor path relies on (descriptor fields, operand images), through the
C-ABI diagnostic bcad_selftest_umma: one 128 x N x K UMMA against a float matmul of the same bf16 operands."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LAYOUT_NONE, LAYOUT_SW128 = 0, 2


def _bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def image_noswizzle(X: torch.Tensor) -> np.ndarray:
    """[R,K] -> bytes laid out [K/8][R][8]: core matrices of 8 rows x 16 B are contiguous (128 B)."""
    R, K = X.shape
    return np.ascontiguousarray(_bf16_bits(X).reshape(R, K // 8, 8).transpose(1, 0, 2))


def image_sw128(X: torch.Tensor) -> np.ndarray:
    """[R,K] (K % 64 == 0, R % 8 == 0) -> [K/64][R][128 B] with 16-byte chunk c stored at c ^ (r & 7)."""
    R, K = X.shape
    bits = _bf16_bits(X).reshape(R, K // 64, 8, 8).transpose(1, 0, 2, 3)       # [kb][r][chunk][8]
    out = np.empty_like(bits)
    r = np.arange(R)
    for c in range(8):
        out[:, r, c ^ (r & 7), :] = bits[:, r, c, :]
    return np.ascontiguousarray(out)


def run_umma(a_img, b_img, N, steps, a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout, a_koff, b_koff, mode=0):
    import bcad_b200
    lib = bcad_b200._lib.load()
    params = np.zeros(8 + 128 + 1, np.int32)
    params[:8] = [N, steps, a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout]
    params[8:8 + steps] = a_koff
    params[72:72 + steps] = b_koff
    params[136] = mode
    a_dev = torch.from_numpy(a_img.view(np.int16).reshape(-1).copy()).cuda()
    b_dev = torch.from_numpy(b_img.view(np.int16).reshape(-1).copy()).cuda()
    d = torch.full((128, N if mode == 0 else N // 2), float("nan"), device="cuda")
    rc = lib.bcad_selftest_umma(C.c_void_p(a_dev.data_ptr()), a_dev.numel() * 2, C.c_void_p(b_dev.data_ptr()),
                                b_dev.numel() * 2, C.c_void_p(params.ctypes.data), C.c_void_p(d.data_ptr()), None)
    bcad_b200._lib.check(rc)
    torch.cuda.synchronize()
    return d.cpu()


@pytest.mark.parametrize("N", [64, 32, 256])
@pytest.mark.parametrize("K", [16, 32, 288])
@pytest.mark.parametrize("shift", [0, 1, 2, 5])
def test_noswizzle_kmajor_with_shifted_start(N, K, shift):
    """conv implicit GEMM: A rows are consecutive pixels 16 B apart; a tap shift is a start-address offset."""
    if K // 16 > 64 or (128 + 8 + N) * K * 2 > 190 * 1024:
        pytest.skip("operand images exceed the harness's shared memory")
    g = torch.Generator().manual_seed(N * 1000 + K + shift)
    RA = 128 + 8
    A = torch.randn(RA, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    steps = K // 16
    a_koff = [shift * 16 + 2 * k * RA * 16 for k in range(steps)]
    b_koff = [2 * k * N * 16 for k in range(steps)]
    got = run_umma(image_noswizzle(A), image_noswizzle(B), N, steps, RA * 16, 128, LAYOUT_NONE, N * 16, 128, LAYOUT_NONE,
                   a_koff, b_koff)
    want = A[shift:shift + 128].float() @ B.float().T
    err = (got - want).abs().max().item()
    assert err <= 1e-3 * max(1.0, want.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("N", [256, 64])
@pytest.mark.parametrize("K", [64, 256])
def test_sw128_kmajor_k_advance(N, K):
    """fc1 GEMM: 128-byte swizzled rows, K advanced by +32 B inside the 64-element block, +tile per block."""
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    steps = K // 16
    a_koff = [(k // 4) * 128 * 128 + (k % 4) * 32 for k in range(steps)]
    b_koff = [(k // 4) * N * 128 + (k % 4) * 32 for k in range(steps)]
    got = run_umma(image_sw128(A), image_sw128(B), N, steps, 16, 1024, LAYOUT_SW128, 16, 1024, LAYOUT_SW128, a_koff, b_koff)
    want = A.float() @ B.float().T
    err = (got - want).abs().max().item()
    assert err <= 1e-3 * max(1.0, want.abs().max().item()), f"max err {err}"


@pytest.mark.parametrize("N", [32, 64, 128])
def test_fp16_accumulators_still_take_one_tmem_column_each(N):
    """kind::f16 with D format F16: measured here because a packed layout (two accumulators per 32-bit TMEM cell) would have let the
    fused conv kernel's first block run all four pool classes in one pass.  It is NOT packed: accumulator n sits alone in the low
    half of column n, so half-precision accumulators save no tensor memory.  Values = the fp32 product sum
    rounded to fp16."""
    g = torch.Generator().manual_seed(N)
    A = torch.randn(128, 16, generator=g).to(torch.float16)
    B = torch.randn(N, 16, generator=g).to(torch.float16)

    def img(X):
        R, K = X.shape
        return np.ascontiguousarray(X.view(torch.int16).numpy().astype(np.uint16).reshape(R, K // 8, 8).transpose(1, 0, 2))
    got = run_umma(img(A), img(B), N, 1, 128 * 16, 128, LAYOUT_NONE, N * 16, 128, LAYOUT_NONE, [0], [0], mode=1)
    cells = got.numpy().view(np.uint32)                                   # the first N/2 columns, raw
    lo = (cells & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    want = (A.float() @ B.float().T).numpy().astype(np.float16).astype(np.float32)[:, :N // 2]
    ulp = np.maximum(np.abs(want), 2.0 ** -14) * 2.0 ** -10
    assert np.all(np.abs(lo - want) <= ulp)
