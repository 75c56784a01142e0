"""Per-tensor NF4 / nf8 quantize (abs-max pass + code pass), CUDA-graph replay over rotating inputs."""
import sys, json; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, quanta_b200 as Q
for shape in ((4096, 4096), (11008, 4096)):
    xs = [torch.randn(shape, device="cuda") * 0.02 for _ in range(4)]
    for name, fn in (("nf4", lambda x: Q.quantize_4bit(x, quant_type="nf4")), ("nf8", lambda x: Q.quantize_8bit(x, quant_type="nf8"))):
        for i in range(3): fn(xs[i % 4])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side):
            outs = [fn(xs[i % 4]) for i in range(20)]
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(shape, name, "per-tensor us", round(e0.elapsed_time(e1) * 1e3 / 20, 2), flush=True)
