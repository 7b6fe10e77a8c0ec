import csv,sys
f=sys.argv[1]
rows=list(csv.reader(open(f)))
hdr=rows[1]; data=rows[2:]
iA=hdr.index("Address"); iS=hdr.index("Source"); iN=hdr.index("# Samples"); iE=hdr.index("Instructions Executed"); iT=hdr.index("Thread Instructions Executed")
segs=[]; cur=dict(start=None,inst=0,samp=0,thr=0,n=0,mufu=0,first=None)
tot=0;totS=0
for r in data:
    if len(r)<=iT or not r[iE].replace(",","").isdigit(): continue
    e=int(r[iE].replace(",","") or 0); s=int(r[iN].replace(",","") or 0); t=int(r[iT].replace(",","") or 0)
    if cur['start'] is None: cur['start']=r[iA]
    cur['inst']+=e; cur['samp']+=s; cur['thr']+=t; cur['n']+=1
    if 'MUFU' in r[iS]: cur['mufu']+=e
    tot+=e; totS+=s
    if 'BAR.SYNC' in r[iS] or 'EXIT' in r[iS]:
        cur['end']=r[iA]; segs.append(cur); cur=dict(start=None,inst=0,samp=0,thr=0,n=0,mufu=0)
if cur['n']: cur['end']='end'; segs.append(cur)
print("total warp-inst",tot,"samples",totS)
for s in segs:
    if s['inst']==0 and s['samp']==0: continue
    print(f"{s['start']:>8s}-{s['end']:>8s} n={s['n']:5d} inst={s['inst']:12d} ({s['inst']/tot*100:5.1f}%) samples={s['samp']:7d} ({s['samp']/max(totS,1)*100:5.1f}%) thr/inst={s['thr']/max(s['inst'],1):5.1f} mufu={s['mufu']}")
