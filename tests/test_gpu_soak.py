"""Soak / robustness tests of the TMA-ring kernels: every kernel family that recycles
shared-memory stages or hands data between CTAs is repeated >= 1000 times at several sizes and must reproduce
its first result bit for bit (a stage released too early, a stale counter or a missed fence shows up as a
differing repetition); plus the grid-barrier time-out path, which must raise instead of killing the context."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import quanta_b200 as q
    return q


def _soak(fn, reps, check_every=50):
    ref = fn()
    ref = [t.clone() for t in (ref if isinstance(ref, (tuple, list)) else (ref,))]
    bad = 0
    for i in range(reps):
        out = fn()
        out = out if isinstance(out, (tuple, list)) else (out,)
        if i % check_every == check_every - 1 or i == reps - 1:
            bad += sum(0 if torch.equal(a, b) else 1 for a, b in zip(out, ref))
        else:
            # cheap on-device comparison every iteration, one host sync per `check_every`
            for a, b in zip(out, ref):
                if a.numel():
                    bad_flag = (a != b).any()
                    _soak.acc = bad_flag if getattr(_soak, "acc", None) is None else (_soak.acc | bad_flag)
    if getattr(_soak, "acc", None) is not None:
        bad += int(_soak.acc.item())
        _soak.acc = None
    return bad


@pytest.mark.parametrize("shape", [(1024, 1024), (4096, 4096), (1536, 11008), (257, 4160)])
@pytest.mark.parametrize("mode", ["block4pack", "block8", "tensor8", "dim0", "backend_sym", "nf4"])
def test_quantize_rings_soak(Q, mode, shape):
    from quanta_b200 import backends as QB
    g = torch.Generator(device="cuda").manual_seed(shape[0] + shape[1])
    x = torch.randn(*shape, device="cuda", generator=g)
    fn = {"block4pack": lambda: Q.quantize_4bit(x, blocksize=64, packed=True),
          "block8": lambda: Q.quantize_8bit(x, blocksize=64),
          "tensor8": lambda: Q.quantize_8bit(x),
          "dim0": lambda: Q.quantize_8bit(x, per_channel=True),
          "backend_sym": lambda: QB.quantize_8bit(x, False, True),
          "nf4": lambda: (lambda r: (r[0], r[2]))(Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True))}[mode]
    assert _soak(fn, 1000) == 0


@pytest.mark.parametrize("case", [(4096, 4096, 1, 4), (4096, 4096, 16, 4), (1024, 14336, 8, 4), (2048, 4096, 16, 8),
                                  (4096, 4096, 64, 4), (2048, 4096, 256, 4), (1024, 2048, 128, 8)])
def test_gemm_soak(Q, case):
    """Both dequant-GEMM kernels (weight-stream kernel for M <= 16, tcgen05 kernel above): TMA rings, stream-K
    partials and counters; identical output on every repetition and the workspace left clean."""
    from quanta_b200.nn import linear_wna16
    N, K, M, bits = case
    g = torch.Generator(device="cuda").manual_seed(N + K + M)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.02
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16)
    qf = Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64)
    assert _soak(lambda: linear_wna16(x, *qf, b, bits=bits, blocksize=64, out_features=N), 1000) == 0


@pytest.mark.parametrize("tokens", [1, 16, 256])
def test_outlier_matmul_soak(Q, tokens):
    from quanta_b200.nn.functional import int8_outlier_matmul, rowwise_quantize_sym
    g = torch.Generator(device="cuda").manual_seed(tokens)
    w = torch.randn(2048, 4096, device="cuda", generator=g) * 0.02
    x = torch.randn(tokens, 4096, device="cuda", generator=g)
    x[:, [7, 513, 1024]] *= 20
    x = x.to(torch.bfloat16)
    qw, cw = rowwise_quantize_sym(w)
    assert _soak(lambda: int8_outlier_matmul(x, qw, cw, 6.0, None), 300) == 0


def test_grid_barrier_timeout_raises_and_recovers(Q):
    """A grid barrier that cannot complete (here: the arrival counter is poisoned so that it wraps and never
    reaches the grid size) used to __trap() and kill the context.  Now the kernel gives up after ~1 s, reports
    through the host-mapped flag, leaves the header clean; the next call raises QuantaError and the one after
    that is correct again."""
    import numpy as np
    from quanta_b200 import _host, _lib
    from oracle import oracle_np as O
    g = torch.Generator().manual_seed(5)
    x = torch.randn(512, 1024, generator=g)
    xd = x.cuda()
    q0 = Q.quantize_8bit(xd)[0].clone()                       # creates the per-stream workspace
    ws = _host.quantize_workspace(xd.device, 256)
    ws[32:36].view(torch.int32).fill_(-1)                      # arrive counter (int index 8) = 0xFFFFFFFF
    Q.quantize_8bit(xd)                                         # times out inside the kernel, context survives
    torch.cuda.synchronize()
    with pytest.raises(_lib.QuantaError):
        Q.quantize_8bit(xd)
    q2, s2, z2 = Q.quantize_8bit(xd)                            # header was reset: back to normal
    qo, so, zo = O.quantize_affine(x.numpy(), 8, O.MODE_TENSOR)
    assert np.array_equal(q2.cpu().numpy(), qo) and torch.equal(q2, q0)
    assert float(s2) == float(so) and float(z2) == float(zo)


# ---- guard bands: compute-sanitizer is closed on this pool (profiles/r02_sanitizer.txt), so out-of-bounds WRITES are
# ---- looked for with our own canaries: every output lives inside a larger sentinel-filled buffer, odd sizes included
def _guarded(nbytes, dtype=torch.uint8, pad=4096):
    raw = torch.full((nbytes + 2 * pad,), 0xA5, dtype=torch.uint8, device="cuda")
    return raw, raw[pad:pad + nbytes], pad


def _intact(raw, nbytes, pad):
    return bool((raw[:pad] == 0xA5).all()) and bool((raw[pad + nbytes:] == 0xA5).all())


@pytest.mark.parametrize("n_rows,n_cols", [(520, 1024), (257, 4160), (33, 192), (1, 64)])
def test_quantize_entries_stay_inside_their_outputs(Q, n_rows, n_cols):
    from quanta_b200 import _lib, _host
    L = _lib.lib()
    dev = torch.device("cuda")
    x = torch.randn(n_rows, n_cols, device=dev)
    n = n_rows * n_cols
    st = _host.stream_ptr(dev)
    for mode, block, bits, pack in ((_lib.MODE_BLOCK, 64, 4, 1), (_lib.MODE_BLOCK, 64, 4, 0), (_lib.MODE_BLOCK, 64, 8, 0),
                                    (_lib.MODE_TENSOR, 0, 8, 0), (_lib.MODE_TENSOR, 0, 4, 0), (_lib.MODE_DIM0, 0, 8, 0)):
        rows, cols = (n_rows, n_cols) if mode == _lib.MODE_DIM0 else (1, n)
        nparam = n // 64 if mode == _lib.MODE_BLOCK else (n_cols if mode == _lib.MODE_DIM0 else 1)
        qbytes = (n + 1) // 2 if pack else n
        rq, q, pad = _guarded(qbytes)
        rs, s, _ = _guarded(nparam * 4)
        rz, z, _ = _guarded(nparam * 4)
        ws = _host.quantize_workspace(dev, L.quanta_workspace_bytes(_lib.OP_QUANTIZE_AFFINE, rows if mode == _lib.MODE_DIM0 else 1,
                                                                    cols if mode == _lib.MODE_DIM0 else 1))
        rc = L.quanta_quantize_affine(x.data_ptr(), _lib.F32, rows, cols, mode, block, bits, pack, q.data_ptr(), s.data_ptr(),
                                      z.data_ptr(), ws.data_ptr(), ws.numel(), st)
        assert rc == 0, (mode, bits, pack, rc)
        torch.cuda.synchronize()
        assert _intact(rq, qbytes, pad) and _intact(rs, nparam * 4, pad) and _intact(rz, nparam * 4, pad), (mode, bits, pack)
        # and back: dequantize into a guarded fp32 buffer
        ro, o, _ = _guarded(n * 4)
        rc = L.quanta_dequantize_affine(q.data_ptr(), pack, rows, cols, mode, block, s.data_ptr(), z.data_ptr(), o.data_ptr(), _lib.F32, st)
        assert rc == 0
        torch.cuda.synchronize()
        assert _intact(ro, n * 4, pad)
        d = o.view(torch.float32).reshape(n_rows, n_cols)
        assert float((d - x).abs().max()) <= float(s.view(torch.float32).max()) * 0.51 + 1e-6


@pytest.mark.parametrize("case", [(200, 512, 7, 4), (136, 1536, 16, 4), (200, 320, 33, 8), (384, 1024, 130, 4), (128, 256, 1, 8)])
def test_gemm_stays_inside_y(Q, case):
    from quanta_b200 import _lib, _host
    L = _lib.lib()
    dev = torch.device("cuda")
    N, K, M, bits = case
    w = torch.randn(N, K, device=dev) * 0.02
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    qf = Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64)
    ry, y, pad = _guarded(M * N * 2)
    ws = _host.gemm_workspace(dev, L.quanta_workspace_bytes(_lib.OP_GEMM, M, N))
    rc = L.quanta_gemm_wna16(x.data_ptr(), _lib.BF16, qf[0].data_ptr(), bits, qf[1].data_ptr(), qf[2].data_ptr(), 64, None,
                             y.data_ptr(), M, N, K, ws.data_ptr(), ws.numel(), _host.stream_ptr(dev))
    assert rc == 0
    torch.cuda.synchronize()
    assert _intact(ry, M * N * 2, pad)
    from quanta_b200.nn import linear_wna16
    assert torch.equal(y.view(torch.bfloat16).reshape(M, N), linear_wna16(x, *qf, None, bits=bits, blocksize=64, out_features=N))
    assert int(ws[:65536].view(torch.int32).abs().sum()) == 0            # counters left zero
