import json,sys,math
rows = json.load(open(sys.argv[1]))
B,T=16,800
rates=[8,8,2,2]
tot=sum(r["ms"] for r in rows)
print("total",round(tot,2))
agg={}
names={r['name'] for r in rows}
for r in rows:
    if r["kind"]<0: continue
    n=r["name"]
    if n=="conv_pre": fi=T; st="pre"
    elif n.startswith("ups."): i=int(n.split(".")[1]); fi=T*math.prod(rates[:i]); st=f"ups{i}"
    elif n=="conv_post": fi=T*256; st="post"
    else:
        i=int(n.split(".")[1])//3; fi=T*math.prod(rates[:i+1]); st=f"stage{i} k={r['k']:2d} {'c1' if 'convs1' in n else 'c2'}"
    fl=2*B*fi*r["c_in"]*r["c_out"]*r["k"]
    if ".convs2." in n and n.replace(".convs2.",".convs1.") not in names: fl*=2; st=st.replace("c2","pair")
    a=agg.setdefault(st,[0,0,0,r]); a[0]+=r["ms"]; a[1]+=fl; a[2]+=1
stage={}
for k,(ms,fl,n,r) in agg.items():
    print(f"{k:20s} n={n} {ms:7.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s  {r.get('kernel','')} ms={r.get('m_subtiles')} st={r.get('stages')} res={r.get('weights_resident')} nb={r.get('slab_buffers')} smem={r.get('smem_bytes')}")
    s=k.split()[0]; stage[s]=stage.get(s,0)+ms
print({k:round(v,2) for k,v in stage.items()})
