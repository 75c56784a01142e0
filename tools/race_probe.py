"""Repeat the per-tensor / per-channel quantize on one input and compare every run with the first:
any difference is a race.  Prints where (row, tile, owning CTA) the runs differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q

torch.manual_seed(0)
for shape in ((4096, 4096), (2000, 4096), (11008, 4096)):
    for mode in ("tensor", "dim0"):
        x = torch.randn(shape, device="cuda")
        fn = (lambda t: Q.quantize_8bit(t)) if mode == "tensor" else (lambda t: Q.quantize_8bit(t, per_channel=True))
        ref = fn(x)[0].clone()
        bad_runs = 0
        for it in range(200):
            q = fn(x)[0]
            if not torch.equal(q, ref):
                bad_runs += 1
                d = (q.reshape(-1) != ref.reshape(-1)).nonzero().reshape(-1)
                rows = torch.unique(d // 32)
                n_tiles = (x.numel() // 32 + 127) // 128
                tiles = torch.unique(rows // 128)
                print(f"{shape} {mode} run {it}: {d.numel()} codes differ in {rows.numel()} rows, tiles {tiles[:8].tolist()}"
                      f" (of {n_tiles}), rows%128 {(rows % 128)[:8].tolist()}", flush=True)
                if bad_runs > 5:
                    break
        print(f"{shape} {mode}: {bad_runs} bad runs of 200", flush=True)
