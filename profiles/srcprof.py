#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python srcprof.py src.csv [top]"""
import csv, sys
def f(x):
    try: return float(x.replace(',',''))
    except Exception: return 0.0
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 25
sections=[];cur=None
for r in rows:
    if len(r)>=2 and r[0]=="File Path":
        cur={'file':r[1],'rows':[],'hdr':None}; sections.append(cur)
    elif len(r)>5 and r[0]=="Line No":
        cur['hdr']=r
    elif cur is not None and len(r)>5:
        cur['rows'].append(r)
for s in sections:
    h=s['hdr']
    iI=h.index("Instructions Executed"); iT=h.index("Thread Instructions Executed"); iS=h.index("# Samples")
    tot=sum(f(r[iI]) for r in s['rows']); totS=sum(f(r[iS]) for r in s['rows'])
    print("==",s['file'],"rows",len(s['rows']),"total inst %.3f G"%(tot/1e9),"samples",totS)
    agg={}
    for r in s['rows']:
        key=(r[0],r[1][:100])
        a=agg.setdefault(key,[0,0,0])
        a[0]+=f(r[iI]); a[1]+=f(r[iT]); a[2]+=f(r[iS])
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
        print(f"{a[0]/1e6:9.1f}M inst {a[0]/max(tot,1)*100:5.1f}%  thr/inst={a[1]/max(a[0],1):5.1f} samples={a[2]/max(totS,1)*100:5.1f}%  L{k[0]}: {k[1]}")
