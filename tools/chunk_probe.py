import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from oracle import fixtures as fx
from _util import make_generator
m = make_generator(fx.V1, precision="bf16").cuda()
res = {}
with torch.no_grad():
    for B in (1, 2, 4, 16):
        mel = fx.synthetic_mel(B, 800, seed=7).cuda()
        for _ in range(3): m(mel)
        acc = {}
        for _ in range(5):
            for r in m.profile_layers(mel):
                acc.setdefault(r["name"], []).append(r["ms"])
        res[B] = {k: sum(v) / len(v) for k, v in acc.items()}
names = list(res[16].keys())
def stage(n):
    if n.startswith("resblocks."):
        return int(n.split(".")[1]) // 3
    return {"ups.0": 0, "ups.1": 1, "ups.2": 2, "ups.3": 3}.get(n, -1)
print("per-item ms by stage (B=1,2,4,16):")
for s in (-1, 0, 1, 2, 3):
    print(s, [round(sum(v for k, v in res[B].items() if stage(k) == s) / B, 4) for B in (1, 2, 4, 16)])
print("total per item", [round(sum(res[B].values()) / B, 4) for B in (1, 2, 4, 16)])
for k in names:
    if stage(k) in (2, 3):
        print(k, [round(res[B][k] / B, 4) for B in (1, 2, 4, 16)])
