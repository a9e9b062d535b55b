import csv, io, subprocess, sys, re
rep=sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
lines=out.splitlines()
rows=list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
recs=[]
for r in rows[1:]:
    if len(r)<len(hdr): continue
    try: n=int(r[ix["# Samples"]])
    except: continue
    recs.append((int(r[ix["Address"]],16), r[ix["Source"]], n, int(r[ix["Instructions Executed"]] or 0), {h:int(r[ix[h]] or 0) for h in hdr if h.startswith("stall_") and "Not" not in h}))
base=recs[0][0]
tot=sum(r[2] for r in recs)
# markers
marks=[i for i,r in enumerate(recs) if re.search(r"BAR\.SYNC|UTMASTG|SYNCS\.PHASECHK", r[1])]
prev=0
import collections
for m in marks+[len(recs)]:
    seg=recs[prev:m+1]
    n=sum(r[2] for r in seg)
    if n>tot*0.005:
        ex=max((r[3] for r in seg), default=0)
        st=collections.Counter()
        for r in seg:
            st.update(r[4])
        fma=sum(r[3] for r in seg if re.match(r"^(@!?U?P\d+\s+)?(FFMA2|FMUL2|FADD2)", r[1].strip()))
        fs=sum(r[3] for r in seg if re.match(r"^(@!?U?P\d+\s+)?(FFMA|FMUL|FADD)\b", r[1].strip()))
        allx=sum(r[3] for r in seg)
        top=", ".join(f"{k[6:]}={100*v/n:.0f}%" for k,v in st.most_common(4))
        print(f"{seg[0][0]-base:6x}-{seg[-1][0]-base:6x} samples {100*n/tot:5.1f}%  exec {allx/1e6:7.1f}M (packed {fma/1e6:6.1f}M scalarF {fs/1e6:5.1f}M)  ends: {seg[-1][1][:28]:28s} {top}")
    prev=m+1
