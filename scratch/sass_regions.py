"""Aggregate an `ncu --page source --csv` export by contiguous SASS regions (first kernel instance)."""
import csv, sys
from collections import Counter
path=sys.argv[1]; B=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
end=starts[1] if len(starts)>1 else len(rows)
hdr=rows[starts[0]+1]; data=[r for r in rows[starts[0]+2:end] if len(r)==len(hdr)]
iex=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); isrc=hdr.index('Source')
tot=sum(int(r[iex]) for r in data); ts=sum(int(r[isamp]) for r in data)
print('sass', len(data), 'warp-inst', tot, 'samples', ts)
def op(s):
    t=s.split()
    return t[1] if t[0].startswith('@') else t[0]
for s in range(0,len(data),B):
    blk=data[s:s+B]
    e=sum(int(r[iex]) for r in blk); sm=sum(int(r[isamp]) for r in blk)
    c=Counter(op(r[isrc]) for r in blk).most_common(6)
    print(str(s).rjust(5), f"{100*e/tot:5.1f}% inst {100*sm/max(ts,1):5.1f}% samp", c)
