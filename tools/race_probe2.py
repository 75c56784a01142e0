"""For corrupted rows of the per-tensor quantize, find which input row the observed codes came from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quanta_b200 as Q
torch.manual_seed(0)
shape = (11008, 4096)
x = torch.randn(shape, device="cuda")
ref = Q.quantize_8bit(x)[0].clone().reshape(-1, 32)
n_tiles = ref.shape[0] // 128
G = 148
found = 0
for it in range(400):
    q = Q.quantize_8bit(x)[0].reshape(-1, 32)
    bad = (q != ref).any(dim=1).nonzero().reshape(-1)
    if bad.numel() == 0:
        continue
    for r in bad[:6].tolist():
        src = (ref == q[r]).all(dim=1).nonzero().reshape(-1).tolist()
        t, ts = r // 128, [s // 128 for s in src]
        c = (t * G) // n_tiles
        while (n_tiles * (c + 1)) // G <= t: c += 1
        while (n_tiles * c) // G > t: c -= 1
        t0, t1 = (n_tiles * c) // G, (n_tiles * (c + 1)) // G
        print(f"run {it}: row {r} (tile {t} = t0+{t - t0} of cta {c} [{t0},{t1}), row%128 {r % 128}) holds the codes of rows {src[:4]} -> tiles {ts[:4]}"
              f" offsets {[s - t0 for s in ts[:4]]} row%128 {[s % 128 for s in src[:4]]}", flush=True)
    found += 1
    if found >= 6:
        break
print("done", found)
