import csv, collections, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
with open(src) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict(); tot=0
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1e3
    elif unit=='ms': v*=1e3
    name=re.sub(r'\(.*','',name)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
out=[f"# {title}", "# source: ncu --metrics gpu__time_duration.sum --clock-control none ; cold-cache, serialised: compare SHARES",
     "kernel,launches,total_us,avg_us,share"]
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    out.append(f"{k},{n},{t:.1f},{t/n:.1f},{t/tot:.3f}")
open(dst,'w').write("\n".join(out)+"\n")
print("\n".join(out[:22]))
