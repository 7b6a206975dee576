"""List the hottest SASS instructions of an `ncu --page source --print-source sass --csv` export.
usage: python profiles/sass_hot.py sass.csv [top_n]"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
top=int(sys.argv[2]) if len(sys.argv)>2 else 60
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    try: ex=int(r[idx['Instructions Executed']]); sm=int(r[idx['# Samples']])
    except ValueError: continue
    data.append((ex,sm,r[idx['Address']][-5:],r[idx['Source']]))
tot=sum(d[0] for d in data); tots=sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", tots, "static instrs", len(data))
# print in program order those with >=0.4% of executed instrs
for ex,sm,ad,src in data:
    if ex/tot>0.004 or sm/max(1,tots)>0.006:
        print(f"{ad} {ex/tot*100:5.2f}% ex {sm/max(1,tots)*100:5.2f}% smp  {src[:110]}")
