import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]
iK=hdr.index('Kernel Name'); iM=hdr.index('Metric Name'); iV=hdr.index('Metric Value'); iID=hdr.index('ID'); iG=hdr.index('Grid Size')
d={}
for r in rows[1:]:
    d.setdefault(r[iID],{'k':r[iK].split('(')[0].replace('void ','').replace('msau::','')[:28],'g':r[iG]})[r[iM]]=float(r[iV].replace(',',''))
tot=sum(v['gpu__time_duration.sum'] for v in d.values())
print('total ms', round(tot/1e6,2), 'launches', len(d))
agg={}
for v in d.values():
    a=agg.setdefault(v['k'],[0,0.0]); a[0]+=1; a[1]+=v['gpu__time_duration.sum']
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"  {k:30s} n={n:4d} {t/1e6:8.2f} ms  {100*t/tot:5.1f}%")
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
for v in sorted(d.values(), key=lambda v:-v['gpu__time_duration.sum'])[:n]:
    print(f"{v['k']:28s} {v['g']:18s} {v['gpu__time_duration.sum']/1e3:8.1f}us rd {v['dram__bytes_read.sum']/1e6:7.1f}MB wr {v['dram__bytes_write.sum']/1e6:7.1f}MB warps {v['sm__warps_active.avg.pct_of_peak_sustained_active']:5.1f}% tensor {v['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:5.1f}%")
