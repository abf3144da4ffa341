import csv, collections, re, sys
path = sys.argv[1]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for row in rows:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',',''))/1e3
    name=re.sub(r'<.*','',name)[:70]
    agg[name][0]+=1; agg[name][1]+=v; tot+=v
print("total us %.1f kernels %d" % (tot, sum(a[0] for a in agg.values())))
mine=sum(t for k,(n,t) in agg.items() if 'mpc::' in k or 'tc::' in k)
print("ours us %.1f share %.3f" % (mine, mine/tot))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print("%-72s %5d %10.1f us %5.1f%%"%(k,n,t,100*t/tot))
