import json, os, sys
ROOT="/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import torch
from oracle import fixtures as fx
from _util import make_generator
m = make_generator(fx.V2_NARROW, precision="bf16").cuda()
mel = fx.synthetic_mel(16, 800, seed=7).cuda()
with torch.no_grad():
    m(mel); m(mel)
    rows = m.profile_layers(mel)
tot=sum(r["ms"] for r in rows)
print("total", tot)
agg={}
for r in rows:
    key=(r.get("c_in"), r.get("c_out"), r.get("k"), r.get("kind"), r.get("tensor_core"))
    a=agg.setdefault(key,[0,0]); a[0]+=r["ms"]; a[1]+=1
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]): print(k, round(v[0],3), v[1])
