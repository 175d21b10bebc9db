"""Top stall-sample instructions (and per-source-line totals) from an .ncu-rep source page."""
import csv, subprocess, sys, collections
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(['ncu','-i',path,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
data=[]
for r in rows[2:]:
    if len(r)<=isamp: continue
    try: data.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
    except ValueError: pass
tot=sum(d[1] for d in data)
print('instructions',len(data),'samples',tot)
# cumulative profile in 20 buckets of instruction index
B=25; n=len(data)
for b in range(B):
    seg=data[b*n//B:(b+1)*n//B]
    sm=sum(d[1] for d in seg)
    ops=collections.Counter(d[0].split()[0] if not d[0].startswith('@') else d[0].split()[1] for d in seg)
    print(f'[{b*n//B:5d}-{(b+1)*n//B:5d}] {100*sm/tot:5.1f}%  ', ', '.join(f'{k}:{v}' for k,v in ops.most_common(5)))
print('--- top instructions')
for i,d in sorted(enumerate(data), key=lambda x:-x[1][1])[:topn]:
    print(f'{i:5d} {d[1]:6d} {d[2]:9d}  {d[0][:100]}')
