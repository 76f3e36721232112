"""Short inputs: the ResBlocks of a stage on separate streams (api.cu::concurrent_eligible) against the serial
schedule, interleaved in one process — eager back-to-back, one synchronised request at a time, and CUDA-graph replay.

    python tools/concurrent_blocks_ab.py [B] [T]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import fixtures as fx
from _util import make_generator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mel = fx.synthetic_mel(B, T, seed=7).cuda()
models = {}
for name, v in (("serial", "0"), ("concurrent", os.environ.get("KELEMS", "2560"))):
    os.environ["HG_CONCURRENT_KELEMS"] = v
    m = make_generator(fx.V1, precision="bf16").cuda()
    with torch.no_grad():
        y = m(mel)
    models[name] = (m, y.clone())
assert torch.equal(models["serial"][1], models["concurrent"][1])
graphs = {n: m.make_graphed(B, T) for n, (m, _) in models.items()}
for n in graphs:
    assert torch.equal(graphs[n](mel), models[n][1])
res = {n: {"eager": [], "graph": [], "single": []} for n in models}
with torch.no_grad():
    for rnd in range(6):
        for n, (m, _) in models.items():
            for _ in range(10): m(mel)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): m(mel)
            e1.record(); torch.cuda.synchronize()
            res[n]["eager"].append(e0.elapsed_time(e1) / 100)
            # one forward at a time, host-synchronised: what a single request sees
            t0 = time.perf_counter()
            for _ in range(50):
                m(mel); torch.cuda.synchronize()
            res[n]["single"].append((time.perf_counter() - t0) / 50 * 1e3)
            g = graphs[n]
            for _ in range(10): g(mel)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(100): g(mel)
            e1.record(); torch.cuda.synchronize()
            res[n]["graph"].append(e0.elapsed_time(e1) / 100)
print(f"B={B} T={T}")
for n in res:
    print(n, {k: round(min(v), 4) for k, v in res[n].items()}, {k: round(sum(v) / len(v), 4) for k, v in res[n].items()})
