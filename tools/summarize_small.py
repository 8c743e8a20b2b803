import csv, collections, sys
for w in sys.argv[1:] or ("chol", "lu"):
    rows = list(csv.reader(open(f'gpurun_out/small_{w}.csv')))
    hi = next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    seq = [(r[ki].split('(')[0], float(r[vi].replace(',',''))/1e3) for r in rows[hi+1:] if len(r) > vi]
    seq = seq[len(seq)//2:]
    a = collections.defaultdict(lambda:[0,0.0])
    for n_, us in seq: a[n_][0]+=1; a[n_][1]+=us
    t = sum(v[1] for v in a.values())
    print(w, "launches", len(seq), "sum us", round(t))
    for k_, v in sorted(a.items(), key=lambda kv:-kv[1][1]):
        print(f"   {k_[:70]:70s} n={v[0]:4d} total={v[1]:8.1f} us avg={v[1]/v[0]:7.1f}")
