"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py [family ...]

Each case is checked against torch / the composition oracle so that a sanitizer-clean run is also a correct one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
from quanta_b200 import backends as QB
from quanta_b200.nn import linear_wna16, linear_nf4a16
from quanta_b200.nn.functional import int8_outlier_matmul, rowwise_quantize_sym

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)


def rnd(*shape, scale=0.02):
    return torch.randn(*shape, device=dev, generator=g) * scale


def quantize_ring():
    x = rnd(520, 1024)
    for bits, fn in ((4, Q.quantize_4bit), (8, Q.quantize_8bit)):
        q, s, z = fn(x, blocksize=64)
        d = (Q.dequantize_4bit if bits == 4 else Q.dequantize_8bit)(q, s, z, blocksize=64)
        assert float((d - x).abs().max()) <= float(s.max()) * 0.51
    p, s, z = Q.quantize_4bit(x, blocksize=64, packed=True)
    assert torch.equal(Q.unpack_4bit_tensor(p)[: x.numel()], Q.quantize_4bit(x, blocksize=64)[0].reshape(-1))
    Q.quantize_4bit_many([x[:128], x[128:384], x[384:]], blocksize=64, packed=True)


def fused_tensor():
    x = rnd(520, 1024, scale=1.0)
    for fn in (Q.quantize_8bit, Q.quantize_4bit):
        q, s, z = fn(x)
        assert float(z) == float(x.min())
    QB.quantize_8bit(x, False, True); QB.quantize_4bit(x, False, False)


def dim0():
    x = rnd(520, 1024, scale=1.0)
    q, s, z = Q.quantize_8bit(x, per_channel=True)
    assert torch.equal(z.reshape(-1), x.min(dim=0).values)
    QB.quantize_8bit(x, True, True); QB.quantize_8bit(x, True, False)


def dequantize():
    x = rnd(300, 1000, scale=1.0)
    q, s, z = Q.quantize_8bit(x)
    assert torch.allclose(Q.dequantize_8bit(q, s, z), x, atol=float(s))
    q, s, z = QB.quantize_8bit(x, False, True)
    QB.dequantize_8bit(q, s, z)
    c = torch.randint(0, 16, (1001,), device=dev, dtype=torch.uint8)
    pk, _ = Q.pack_4bit_tensor(c)
    assert torch.equal(Q.unpack_4bit_tensor(pk)[:1001], c)


def gemm():
    for (N, K, M, bits) in ((256, 512, 5, 4), (384, 2048, 16, 4), (256, 1024, 9, 8), (384, 1024, 40, 4), (256, 512, 200, 8)):
        w, x = rnd(N, K), torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
        b = rnd(N, scale=0.1).to(torch.bfloat16)
        qf = Q.quantize_4bit(w, blocksize=64, packed=True) if bits == 4 else Q.quantize_8bit(w, blocksize=64)
        y = linear_wna16(x, *qf, b, bits=bits, blocksize=64, out_features=N)
        wd = (Q.dequantize_4bit(*qf, blocksize=64, packed=True, shape=(N, K)) if bits == 4 else Q.dequantize_8bit(*qf, blocksize=64))
        ref = x.float() @ wd.t() + b.float()
        assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1e-2, (N, K, M, bits)
    # scatter epilogue into three buffers
    from quanta_b200.nn.functional import linear_wna16_scatter
    N, K, M, ldy, col0 = 384, 512, 48, 1024, 256
    w, x = rnd(N, K), torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    qf = Q.quantize_4bit(w, blocksize=64, packed=True)
    outs = [torch.zeros(M, ldy, dtype=torch.bfloat16, device=dev) for _ in range(3)]
    linear_wna16_scatter(x, *qf, None, ([o.data_ptr() for o in outs], ldy), col0, bits=4, blocksize=64, out_features=N)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[2])


def nf4():
    x = rnd(520, 1024)
    q, lv, am = Q.quantize_4bit(x, quant_type="nf4", blocksize=64, packed=True)
    d = Q.dequantize_4bit(q, lv, am, quant_type="nf4", blocksize=64, packed=True, shape=x.shape)
    assert float((d - x).abs().max()) < 0.02
    Q.quantize_4bit(x, quant_type="nf4")
    Q.quantize_8bit(x, quant_type="nf8", blocksize=64); Q.quantize_8bit(x, quant_type="fp8"); Q.quantize_4bit(x, quant_type="fp4")
    xa = torch.randn(7, 1024, device=dev, generator=g).to(torch.float16)
    linear_nf4a16(xa, q, am, None, blocksize=64, out_features=520)


def outlier():
    w = rnd(256, 1024)
    qw, cw = rowwise_quantize_sym(w)
    for tokens in (1, 40):
        x = torch.randn(tokens, 1024, device=dev, generator=g)
        x[:, [3, 500]] *= 20
        int8_outlier_matmul(x.to(torch.bfloat16), qw, cw, 6.0, None)


FAMILIES = {"quantize_ring": quantize_ring, "fused_tensor": fused_tensor, "dim0": dim0, "dequantize": dequantize,
            "gemm": gemm, "nf4": nf4, "outlier": outlier}
for name in (sys.argv[1:] or list(FAMILIES)):
    FAMILIES[name]()
    torch.cuda.synchronize()
    print("ok", name, flush=True)
