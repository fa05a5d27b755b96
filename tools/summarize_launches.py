import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; data=rows[i+1:]; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in data:
    k=r[ki][:70]; v=float(r[vi].replace(',',''))
    agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print("kernel,launches,total_ns,avg_ns,share")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('"%s",%d,%.0f,%.0f,%.3f'%(k,v[0],v[1],v[1]/v[0],v[1]/tot))
