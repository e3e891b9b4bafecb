import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
h=None
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[start+1:]:
    if len(r)>vi:
        try: d[(r[ki].split('(')[0], r[mi])].append(float(r[vi].replace(',','')))
        except: pass
tot=0
for k,v in sorted(d.items()):
    if 'time' in k[1]:
        print("%-28s %3d launches  %8.1f us" % (k[0], len(v), sum(v)/len(v)/1000)); tot+=sum(v)/len(v)/1000
print("sum %.1f us"%tot)
