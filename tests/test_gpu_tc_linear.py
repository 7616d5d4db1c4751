"""The hand-written tcgen05 linear (tgm_tc_linear, csrc/tc_linear.cu: fp32 operands split in flight
into two TF32 terms, three tensor-core products accumulated in TMEM) against a float64 reference:
the DyGFormer shapes (tgm/nn/encoder/dygformer.py:80-143 at config-5 size: 25 600 tokens, E = 200)
and ragged shapes that exercise row / column / K tails.  The tolerance is the 1e-5 parity bar of
the aggregation path with margin: max abs error <= 4e-6 at unit-scale outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tgm_b200 import _cabi  # noqa: E402

DEV = 'cuda:0'


def _run(S, N, K, residual, gelu, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    A = torch.randn(S, K, generator=g, device=DEV) * scale
    W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    b = torch.randn(N, generator=g, device=DEV)
    R = torch.randn(S, N, generator=g, device=DEV) if residual else None
    out = torch.full((S, N), float('nan'), device=DEV)
    _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(),
                                        _cabi.ptr(R), int(gelu), out.data_ptr(),
                                        torch.cuda.current_stream(DEV).cuda_stream))
    want = A.double() @ W.double().T + b.double()
    if residual:
        want = want + R.double()
    if gelu == 2:
        want = want.relu()
    elif gelu:
        want = 0.5 * want * (1 + torch.erf(want / 2 ** 0.5))
    fp32 = torch.nn.functional.linear(A, W, b)  # cuBLAS fp32 for scale
    if residual:
        fp32 = fp32 + R
    if gelu == 2:
        fp32 = fp32.relu()
    elif gelu:
        fp32 = torch.nn.functional.gelu(fp32)
    return out, want, fp32


@pytest.mark.parametrize('S,N,K,residual,gelu', [
    (25600, 600, 200, False, False),   # in_proj
    (25600, 200, 200, True, False),    # out_proj + residual
    (25600, 800, 200, False, True),    # FFN1 + GELU
    (25600, 200, 800, True, False),    # FFN2 + residual
    (130, 100, 36, False, False),      # row tail, narrow tile, K tail inside one chunk
    (257, 404, 44, True, False),       # three column tiles of 136, K tail in a second chunk
    (4096, 172, 400, False, True),
    (1, 4, 4, False, False),
    (12600, 104, 548, False, False),   # TGAT layer 1: folded output product (attn_fold.cu)
    (12600, 172, 104, False, 2),       # merge layer fc1 + ReLU
    (600, 172, 172, False, False),     # short matrix: narrow column tiles over more SMs
    (600, 272, 888, False, False),     # two-wide, many chunks
    (3000, 64, 33 * 4, False, 2),      # chunk count not a multiple of the stage counts
], ids=['in_proj', 'out_proj', 'ffn1_gelu', 'ffn2', 'tails_a', 'tails_b', 'wiki_out', 'tiny',
        'tgat_out', 'merge_relu', 'short', 'short_wide', 'chunks_5'])
def test_tc_linear_matches_float64(S, N, K, residual, gelu):
    out, want, fp32 = _run(S, N, K, residual, gelu)
    assert bool(torch.isfinite(out).all()), 'unwritten or non-finite outputs'
    err = float((out.double() - want).abs().max())
    err32 = float((fp32.double() - want).abs().max())
    scale = max(1.0, float(want.abs().max()) / 8)
    assert err <= 4e-6 * scale, f'max abs err {err:.3e} (cuBLAS fp32: {err32:.3e})'


def test_tc_linear_residual_may_alias_out_and_large_inputs():
    S, N, K = 2048, 200, 200
    g = torch.Generator(device=DEV).manual_seed(3)
    A = torch.randn(S, K, generator=g, device=DEV) * 30
    W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    b = torch.zeros(N, device=DEV)
    X = torch.randn(S, N, generator=g, device=DEV)
    want = A.double() @ W.double().T + X.double()
    _cabi.check(_cabi.lib.tgm_tc_linear(S, N, K, A.data_ptr(), W.data_ptr(), b.data_ptr(),
                                        X.data_ptr(), 0, X.data_ptr(),
                                        torch.cuda.current_stream(DEV).cuda_stream))
    err = float((X.double() - want).abs().max())  # dot products of scale 30: errors scale with them
    assert err <= 4e-6 * 30


def test_tc_linear_argument_errors():
    A = torch.zeros(8, 6, device=DEV)
    with pytest.raises(_cabi.TGMNativeError, match='N % 4 == 0'):
        _cabi.check(_cabi.lib.tgm_tc_linear(8, 8, 6, A.data_ptr(), A.data_ptr(), A.data_ptr(), None,
                                            0, A.data_ptr(), None))


@pytest.mark.parametrize('M,N,K,act,with_bias', [
    (600, 888, 172, 0, False),   # TGAT layer 2: x-side qk product
    (600, 272, 888, 0, False),   # folded output product
    (600, 172, 444, 2, True),    # merge layer fc1 + ReLU
    (600, 172, 172, 0, True),
    (33, 31, 50, 2, True),       # tails in every dimension, K % 4 != 0 (scalar loads)
    (1, 1, 1, 0, True),
    (4096, 104, 548, 0, False),
], ids=['qk', 'out', 'fc1_relu', 'fc2', 'tails', 'one', 'limit'])
def test_small_gemm_matches_float64(M, N, K, act, with_bias):
    """tgm_small_gemm (csrc/small_gemm.cu): 32x32 SIMT tiles, fused bias / ReLU."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, device=DEV)
    W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    b = torch.randn(N, generator=g, device=DEV) if with_bias else None
    out = torch.full((M + 1, N), float('nan'), device=DEV)
    _cabi.check(_cabi.lib.tgm_small_gemm(M, N, K, A.data_ptr(), W.data_ptr(), _cabi.ptr(b), act,
                                         out.data_ptr(), torch.cuda.current_stream(DEV).cuda_stream))
    want = A.double() @ W.double().T
    if with_bias:
        want = want + b.double()
    if act == 2:
        want = want.relu()
    assert bool(torch.isnan(out[M]).all()), 'wrote past the last row'
    err = float((out[:M].double() - want).abs().max())
    # one fp32 FMA chain over K (what an SGEMM does): the bar is the path's 1e-5, not the 4e-6 of
    # the chunk-promoted tensor-core kernel
    assert err <= 1e-5 * max(1.0, float(want.abs().max()) / 8), err
