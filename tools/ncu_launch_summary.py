import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ki][:60]; v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1e6 if u in ("ns","nsecond") else (v/1e3 if u in ("us","usecond") else v)
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print(f"{k:62s} {n:4d} {t:10.3f} ms  avg {t/n:8.3f}")
