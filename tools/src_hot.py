import csv, sys, subprocess
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
cols = {n: hdr.index(n) for n in stalls}
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) <= isamp: continue
    try: s = int(r[isamp])
    except: continue
    data.append((s, idx, r))
tot = sum(s for s,_,_ in data)
agg = {k:0 for k in cols}
for s, _, r in data:
    for k, i in cols.items():
        try: agg[k] += int(r[i])
        except: pass
print("total samples", tot, "instructions", len(data))
print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v})
for s, idx, r in sorted(data, key=lambda x:-x[0])[:top]:
    print(s, idx, r[isrc][:100], 'exec=', r[iex], {k[6:]:r[i] for k,i in cols.items() if r[i] not in ('0','')})
