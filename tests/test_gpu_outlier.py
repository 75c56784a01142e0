"""GPU tests of the LLM.int8()-style outlier-split matmul (row G3).  PARITY
UNPINNED against the reference (it only stores ``threshold``); the oracle is
oracle/oracle_np.py:int8_outlier_matmul, defined by this repository.
Tolerance: 1e-2 relative (max-normalised), like the other GEMM rows; the
integer part (outlier set J, int8 codes, int32 accumulators) is exact, so the
error is only the 16-bit rounding of the output and fp32 accumulation order."""
import numpy as np
import pytest
import torch

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu
TOL = 1e-2
OUTLIER_COLS = [7, 513, 1024, 2049, 3071, 4000]


def make_x(M, K, dtype, seed, outliers=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    if outliers:
        for c in OUTLIER_COLS:
            if c < K:
                x[:, c] *= 20.0
    return x.to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(128, 128, 1), (256, 512, 16), (200, 320, 7), (512, 4096, 33), (4096, 4096, 256),
                                   (1024, 2048, 300), (256, 8192, 5), (384, 16384, 16)])      # the last two take the split-K path
def test_outlier_matmul_matches_oracle(dtype, shape):
    from quanta_b200.nn import int8_outlier_matmul, rowwise_quantize_sym
    N, K, M = shape
    g = torch.Generator().manual_seed(N + K + M)
    w = torch.randn(N, K, generator=g) * 0.02
    b = (torch.randn(N, generator=g) * 0.1).to(dtype)
    x = make_x(M, K, dtype, seed=M)
    qw, cw = rowwise_quantize_sym(w.cuda())
    # weight codes are the reference's B1-symmetric arithmetic: bit-exact with the oracle
    qo, co = O.rowwise_quantize_sym(w.numpy())
    assert np.array_equal(qw.cpu().numpy(), qo)
    assert np.array_equal(cw.cpu().numpy().view(np.uint32), co.view(np.uint32))
    y = int8_outlier_matmul(x.cuda(), qw, cw, 6.0, b.cuda())
    assert y.shape == (M, N) and y.dtype == dtype
    ref, J = O.int8_outlier_matmul(x.float().numpy(), qo, co, 6.0, b.float().numpy(),
                                   "bf16" if dtype == torch.bfloat16 else "fp16")
    assert len(J) == sum(1 for c in OUTLIER_COLS if c < K)
    err = float(np.abs(y.float().cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < TOL, f"rel err {err:.3e} N={N} K={K} M={M} {dtype}"


def test_outlier_matmul_edge_cases():
    from quanta_b200.nn import int8_outlier_matmul, rowwise_quantize_sym
    g = torch.Generator().manual_seed(5)
    w = torch.randn(256, 256, generator=g) * 0.02
    qw, cw = rowwise_quantize_sym(w.cuda())
    qo, co = O.rowwise_quantize_sym(w.numpy())
    # no outliers at all, an all-zero row, and every column an outlier
    for name, x in (("none", make_x(5, 256, torch.bfloat16, 1, outliers=False) * 0.5),
                    ("zero-row", torch.cat([torch.zeros(1, 256), torch.randn(3, 256, generator=g)]).to(torch.bfloat16)),
                    ("all", (torch.randn(4, 256, generator=g) * 100).to(torch.bfloat16))):
        y = int8_outlier_matmul(x.cuda(), qw, cw, 6.0, None).float().cpu().numpy()
        ref, J = O.int8_outlier_matmul(x.float().numpy(), qo, co, 6.0, None, "bf16")
        err = float(np.abs(y - ref).max() / (np.abs(ref).max() + 1e-30))
        assert err < TOL, f"{name}: rel err {err:.3e} (|J| = {len(J)})"


def test_linear8bitlt_outlier_split_module():
    from quanta_b200.nn import Linear8bitLt
    torch.manual_seed(0)
    lin = Linear8bitLt(512, 256, bias=True, has_fp16_weights=False, threshold=6.0, outlier_split=True).cuda()
    w = lin.weight.detach().clone()
    x = make_x(9, 512, torch.float16, 3).cuda()
    y = lin(x)
    assert y.shape == (9, 256) and y.dtype == torch.float16
    ref_fp = torch.nn.functional.linear(x.float(), w.float(), lin.bias.float())
    assert float((y.float() - ref_fp).abs().max() / ref_fp.abs().max()) < 0.03      # int8 quantization error only
