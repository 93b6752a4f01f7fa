import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1],errors='ignore')))
hdr=None
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r; ki=r.index('Kernel Name'); vi=r.index('Metric Value'); ui=r.index('Metric Unit')
        continue
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='us': v*=1e3
    elif r[ui]=='ms': v*=1e6
    n=r[ki][:80]
    agg[n][0]+=1; agg[n][1]+=v
tot=sum(v[1] for v in agg.values())
print("total %.1f us"%(tot/1e3))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1]): print("%-80s %6d %12.1f us %5.1f%%"%(k,v[0],v[1]/1e3,100*v[1]/tot))
