import csv, collections, sys
path = sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0,[]])
for row in r:
    name=row['Kernel Name']
    try: v=float(row['Metric Value'].replace(',',''))
    except: continue
    unit=row['Metric Unit']
    if unit=='us': v*=1e3
    elif unit=='ms': v*=1e6
    agg[name][0]+=1; agg[name][1]+=v; agg[name][2].append(v)
tot=sum(v[1] for v in agg.values())
print(f"total {tot/1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    big=sorted(v[2])[-3:]
    print(f"{v[1]/1e6:9.3f} ms {v[0]:6d} x  {100*v[1]/tot:5.1f}%  max {big[-1]/1e3:8.1f} us  {k[:120]}")
