"""The drop-in exercised through the REFERENCE's own code: ``integration/apply_dropin.py``
adds ``Quanta/backends/cuda/`` and the INTEGRATION.md §3 device checks to a copy of the unmodified reference
(``oracle/_ref/Quanta``, vendored by ``make -C oracle``), and the patched package is imported in a fresh
interpreter.

CPU part (runs everywhere): the patch applies, ``Quanta.backends.CUDA_AVAILABLE`` flips to True, CPU tensors
still run the reference's own code (bit-identical to the unpatched package), and the reference's own test
file still passes.
GPU part (-m gpu): the reference's dispatcher forwards CUDA tensors to the sm_100a kernels with results
bit-identical to its CPU backend, and the reference's CUDA twin tests
(Quanta/tests/test_quantization.py:34-124) pass on the new kernels."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "oracle", "_ref", "Quanta")

needs_ref = pytest.mark.skipif(not os.path.isdir(REF_PKG), reason="oracle/_ref not built (make -C oracle)")


def patched_tree(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    try:
        import apply_dropin
    finally:
        sys.path.pop(0)
    apply_dropin.apply(REF_PKG, str(tmp_path))
    return str(tmp_path)


def run_py(code, tree, extra_path=()):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=os.pathsep.join([tree, ROOT, *extra_path]))
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


@needs_ref
def test_manifest_matches_vendored_reference():
    """oracle/_ref is the unmodified reference: every file matches the sha256 manifest written at vendoring time,
    and (in the build container) the reference tree itself."""
    import hashlib
    man = os.path.join(ROOT, "oracle", "_ref", "MANIFEST.sha256")
    lines = [l.split() for l in open(man) if l.strip()]
    assert len(lines) >= 15
    for digest, rel in lines:
        data = open(os.path.join(ROOT, "oracle", "_ref", rel), "rb").read()
        assert hashlib.sha256(data).hexdigest() == digest, rel
        live = os.path.join("/root/reference", rel)
        if os.path.exists(live):
            assert open(live, "rb").read() == data, rel


@needs_ref
def test_dropin_enables_the_dispatcher_and_keeps_cpu_behaviour(tmp_path):
    tree = patched_tree(tmp_path)
    out = run_py("""
        import json, torch
        import Quanta.backends as B
        from Quanta.functional.quantization import quantize_8bit, dequantize_8bit, quantize_4bit
        from Quanta.utils.utils import pack_4bit_tensor, unpack_4bit_tensor
        t = torch.tensor([-1.0, 0.0, 1.0, 2.0])
        q, s, z = quantize_8bit(t)
        q4, s4, z4 = quantize_4bit(t)
        pk, shp = pack_4bit_tensor(torch.arange(16, dtype=torch.uint8))
        qb, sb, zb = B.quantize_8bit(t, False, True)
        print(json.dumps({"cuda_available": B.CUDA_AVAILABLE, "file": B.__file__, "q": q.tolist(), "s": float(s), "z": float(z),
                          "q4": q4.tolist(), "pk": pk.tolist(), "qb": qb.tolist(), "sb": float(sb),
                          "deq": dequantize_8bit(q, s, z).tolist(), "unpk": unpack_4bit_tensor(pk).tolist()}))
        """, tree)
    r = json.loads(out.strip().splitlines()[-1])
    assert r["cuda_available"] is True and r["file"].startswith(tree)
    # SURVEY Appendix B known answers: the CPU path is still the reference's own
    assert r["q"] == [0, 85, 170, 255] and abs(r["s"] - 0.011764706112) < 1e-12 and r["z"] == -1.0
    assert r["q4"] == [0, 5, 10, 15] and r["deq"] == [-1.0, 0.0, 1.0, 2.0]
    assert r["pk"] == [0x10, 0x32, 0x54, 0x76, 0x98, 0xBA, 0xDC, 0xFE] and r["unpk"] == list(range(16))
    assert r["qb"] == [64, 128, 192, 255] and r["sb"] == 63.5


@needs_ref
def test_reference_test_file_passes_on_the_patched_tree_cpu(tmp_path):
    tree = patched_tree(tmp_path)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join([tree, ROOT]))
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-m", "",
                          os.path.join(tree, "Quanta", "tests", "test_quantization.py")],
                         capture_output=True, text=True, env=env, cwd=tree, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout


@needs_ref
@pytest.mark.gpu
def test_reference_cuda_twins_pass_on_the_new_kernels(tmp_path):
    """Quanta/tests/test_quantization.py: 4 CPU tests + their 4 *_cuda twins (:34-49, :61-71, :87-101, :114-124) —
    8 passed, 0 skipped (the twins skip on a box without the drop-in's kernels or without a GPU)."""
    tree = patched_tree(tmp_path)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=os.pathsep.join([tree, ROOT]))
    out = subprocess.run([sys.executable, "-m", "pytest", "-v", "-p", "no:cacheprovider", "-m", "", "-rs",
                          os.path.join(tree, "Quanta", "tests", "test_quantization.py")],
                         capture_output=True, text=True, env=env, cwd=tree, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "8 passed" in out.stdout and "skipped" not in out.stdout, out.stdout[-1500:]
    for name in ("test_8bit_quantization_cuda", "test_8bit_quantization_per_channel_cuda", "test_4bit_quantization_cuda",
                 "test_quantization_edge_cases_cuda"):
        assert name + " PASSED" in out.stdout, out.stdout[-1500:]


@needs_ref
@pytest.mark.gpu
def test_reference_dispatcher_runs_cuda_tensors_on_the_new_kernels(tmp_path):
    """Quanta.backends.quantize_*bit / dequantize_*bit (backends/__init__.py:42-127) with CUDA tensors: the
    dispatcher picks 'cuda', libquanta_b200.so is mapped into the process, and codes / scale / zero-point /
    dequantized values are bit-identical to the reference's CPU backend on the same inputs."""
    tree = patched_tree(tmp_path)
    out = run_py("""
        import json, torch
        import Quanta.backends as B
        assert B.CUDA_AVAILABLE
        g = torch.Generator().manual_seed(7)
        bad = []
        for shape in [(64, 96), (257, 33), (5, 7, 9)]:
            x = torch.randn(*shape, generator=g)
            for bits, qf, df in ((8, B.quantize_8bit, B.dequantize_8bit), (4, B.quantize_4bit, B.dequantize_4bit)):
                for per_channel in (False, True):
                    for symmetric in (True, False):
                        assert B._get_backend(x.cuda()) == "cuda" and B._get_backend(x) == "cpu"
                        qc, sc, zc = qf(x, per_channel, symmetric)
                        qg, sg, zg = qf(x.cuda(), per_channel, symmetric)
                        dc, dg = df(qc, sc, zc), df(qg, sg, zg)
                        ok = (qg.is_cuda and torch.equal(qg.cpu(), qc) and sg.shape == sc.shape
                              and torch.equal(sg.cpu().view(torch.int32), sc.view(torch.int32))
                              and torch.equal(zg.cpu().view(torch.int32), zc.view(torch.int32))
                              and torch.equal(dg.cpu().view(torch.int32), dc.view(torch.int32)))
                        if not ok:
                            bad.append((shape, bits, per_channel, symmetric))
        mapped = any("libquanta_b200.so" in l for l in open("/proc/self/maps"))
        print(json.dumps({"bad": bad, "mapped": mapped}))
        """, tree)
    r = json.loads(out.strip().splitlines()[-1])
    assert r["mapped"], "the CUDA path did not load libquanta_b200.so"
    assert r["bad"] == [], r["bad"]
