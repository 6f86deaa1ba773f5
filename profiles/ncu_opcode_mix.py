import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi+1:] if len(r)==len(hdr)]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
num = lambda v: int(float(v)) if v not in ("", None) else 0
agg = collections.Counter(); ex = collections.Counter()
for r in data:
    s = r[isrc].strip()
    s = re.sub(r'^@!?U?P\d+\s+', '', s)
    op = s.split()[0] if s else '?'
    op = op.split('.')[0] if not op.startswith(('LDL','STL','LDS','STS','LDTM','STTM','UTC','SYNCS','BAR','MUFU','LDG','STG')) else '.'.join(op.split('.')[:2])
    agg[op] += num(r[isamp]); ex[op] += num(r[iex])
tot = sum(agg.values()); tex = sum(ex.values())
for op, v in agg.most_common(40):
    print("%-18s samples %8d (%5.1f%%)  executed %12d (%5.1f%%)" % (op, v, 100*v/tot, ex[op], 100*ex[op]/tex))
