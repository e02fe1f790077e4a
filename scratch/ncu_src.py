"""Summarise an ncu --page source --csv export: hottest SASS lines by stall samples."""
import csv, sys
path = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index('# Samples')].isdigit()]
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
tot_s = sum(int(r[isamp]) for r in data); tot_i = sum(int(r[iex]) for r in data)
print('sass lines', len(data), 'samples', tot_s, 'warp-inst', tot_i)
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:ntop]
for i in sorted(top):
    r = data[i]
    print(str(i).rjust(5), r[ia].strip()[:90].ljust(90), r[isamp].rjust(6), f"{100*int(r[isamp])/max(tot_s,1):5.1f}%", r[iex].rjust(9))
