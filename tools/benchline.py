import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["vocab"], "tok/s %.3fM"%(d["value"]/1e6), "step %.3f ms"%d["ms_per_step"], "lookup %.3f ms (%.1f%%)"%(r["ms_per_launch"],100*r["frac"]), "decode %.3f ms (%.1f%%)"%(r["decode"]["ms_per_call"],100*r["decode"]["frac"]), "coder %.3f"%r["coder_kernel_ms"])
