"""GPU: the library's three GEMM kernels against torch (bf16 inputs, fp32 accumulate) on the
shapes the path uses.  bf16 x bf16 products are exact in fp32, so the only difference is the
order of the fp32 sum: tolerance 1e-3 relative to the row scale."""
import pytest
import torch

from helpers import package

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import importlib
    package()
    lib = importlib.import_module("mr-mt3_b200._lib")
    return lib.Engine()


def _check(eng, which, M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed + M + 7 * N + 13 * K)
    a = torch.randn((M, K), generator=g).bfloat16()
    w = (torch.randn((N, K), generator=g) * K ** -0.5).bfloat16()
    got = eng.test_gemm(a, w, which).cpu()
    want = a.double() @ w.double().T
    err = (got.double() - want).abs().max().item()
    assert err < 2e-3, (which, M, N, K, err)


# encoder / projection shapes: (M, N, K)
BIG = [(256, 512, 512), (512, 1152, 512), (300, 384, 512), (1000, 512, 384), (1024, 2048, 512),
       (640, 512, 1024), (128, 6144, 512), (77, 1536, 512)]


@pytest.mark.parametrize("M,N,K", BIG)
def test_gemm_mma_sync(eng, M, N, K):
    _check(eng, 0, M, N, K)


@pytest.mark.parametrize("M,N,K", BIG)
def test_gemm_tcgen05(eng, M, N, K):
    _check(eng, 1, M, N, K)


@pytest.mark.parametrize("M,N,K", [(256, 1152, 512), (64, 512, 384), (33, 2048, 512), (256, 512, 1024), (5, 384, 512)])
def test_gemm_decode_single_shot(eng, M, N, K):
    _check(eng, 2, M, N, K)


def test_tcgen05_large(eng):
    _check(eng, 1, 65536, 1152, 512)


# CTA-pair tiles (tcgen05.mma.cta_group::2, 256 x BN per two-CTA cluster) against the single-CTA kernel:
# same products summed in the same order, so the outputs must be BIT-identical; against fp64 as above.
# Shapes: every BN the pair kernel has (256 / 192 / 128), M not a multiple of 256 (the second CTA of the
# last pair works on rows past M), M < 256 (launcher falls back to single), K = 64 (one k-block), both
# K-major (which 1, bf16-out 3) and the data-gradient form (which 5, B operand MN-major).
PAIR = [(1, 256, 256, 64), (1, 512, 192, 512), (1, 384, 128, 384), (1, 1000, 1152, 512), (1, 4096, 2048, 512),
        (1, 8192, 512, 1024), (1, 130, 512, 512), (3, 2048, 1152, 512), (3, 777, 512, 384),
        (5, 1000, 512, 1152), (5, 4096, 384, 512), (5, 8192, 1024, 512), (5, 256, 256, 128)]


@pytest.mark.parametrize("which,M,N,K", PAIR)
def test_gemm_tcgen05_cta_pair_equals_single(eng, which, M, N, K):
    g = torch.Generator().manual_seed(which + M + 7 * N + 13 * K)
    a = torch.randn((M, K), generator=g).bfloat16()
    w = (torch.randn((K, N) if which == 5 else (N, K), generator=g) * K ** -0.5).bfloat16()
    try:
        eng.set_option("gemm_2cta", 1)
        pair = eng.test_gemm(a, w, which)
        eng.set_option("gemm_2cta", 0)
        single = eng.test_gemm(a, w, which)
    finally:
        eng.set_option("gemm_2cta", -1)
    if which == 3:                                   # the hook's buffer holds M * N bf16 in its first half
        pair = pair.view(torch.bfloat16).reshape(-1)[:M * N].reshape(M, N).float()
        single = single.view(torch.bfloat16).reshape(-1)[:M * N].reshape(M, N).float()
    assert torch.equal(pair, single)
    want = a.double() @ (w.double() if which == 5 else w.double().T)
    err = (pair.cpu().double() - want).abs().max().item()
    assert err < (2e-2 if which == 3 else 2e-3), (which, M, N, K, err)   # which 3 adds the bf16 output rounding


# the fine-tune backward's two operand forms, at the reduction depth of configs[4] (32 x 1024 rows)
@pytest.mark.parametrize("R,M,N", [(32768, 512, 384), (32768, 1152, 512), (32768, 2048, 512), (10240, 6144, 512),
                                   (1000, 512, 512), (32768, 512, 1024), (32768, 1536, 512)])
def test_gemm_tcgen05_wgrad_form_split_k(eng, R, M, N):
    """which = 4: C (M, N) = A^T W over R rows, both operands row-major (MN-major tcgen05 operands),
    split-K chosen as the weight-gradient path chooses it + the fixed-order reduce kernel."""
    g = torch.Generator().manual_seed(R + 3 * M + 5 * N)
    a = (torch.randn((R, M), generator=g) * R ** -0.5).bfloat16()
    w = torch.randn((R, N), generator=g).bfloat16()
    got = eng.test_gemm(a, w, 4).cpu()
    want = (a.double().T @ w.double())
    err = (got.double() - want).abs().max().item()
    assert got.shape == (M, N) and err < 2e-3, (R, M, N, err)


@pytest.mark.parametrize("M,K,N", [(32768, 1536, 512), (32768, 512, 384), (8192, 2048, 512), (32768, 1152, 512),
                                   (777, 384, 512), (32768, 512, 1024)])
def test_gemm_tcgen05_dgrad_form(eng, M, K, N):
    """which = 5: C (M, N) = A W with W (K, N) row-major as stored (B operand MN-major)."""
    g = torch.Generator().manual_seed(M + 3 * K + 5 * N)
    a = torch.randn((M, K), generator=g).bfloat16()
    w = (torch.randn((K, N), generator=g) * K ** -0.5).bfloat16()
    got = eng.test_gemm(a, w, 5).cpu()
    want = a.double() @ w.double()
    err = (got.double() - want).abs().max().item()
    assert got.shape == (M, N) and err < 2e-3, (M, K, N, err)
