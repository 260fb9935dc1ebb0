import csv, sys, subprocess
rep = sys.argv[1]; fname=sys.argv[2]; ranges=[(a.split(':')[0], int(a.split(':')[1].split('-')[0]), int(a.split('-')[1])) for a in sys.argv[3:]]
txt = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur=None; hdr=None; agg={}; other=[0,0]
for r in rows:
    if not r: continue
    if r[0] in ('File Path','File Name'): cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0].isdigit():
        try:
            s=int(r[hdr.index('# Samples')]); i=int(r[hdr.index('Instructions Executed')])
        except Exception: continue
        ln=int(r[0]); hit=False
        if cur==fname:
            for n,a,b in ranges:
                if a<=ln<=b:
                    agg.setdefault(n,[0,0]); agg[n][0]+=s; agg[n][1]+=i; hit=True; break
        if not hit:
            agg.setdefault('other:'+cur,[0,0]); agg['other:'+cur][0]+=s; agg['other:'+cur][1]+=i
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]): print("%-28s samples %5.1f%%  instr %5.1f%%" % (k, 100*v[0]/ts, 100*v[1]/ti))
