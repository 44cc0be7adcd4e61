import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    n=r[ki].split("(")[0][:70]; v=float(r[vi].replace(",",""))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{n:70s} n={a[0]:3d} total {a[1]/1e3:8.1f} us  mean {a[1]/a[0]/1e3:7.1f}")
print("total us", tot/1e3, "launches", sum(a[0] for a in agg.values()))
