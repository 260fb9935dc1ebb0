import csv, sys, subprocess
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur=None; hdr=None; out=[]
for r in rows:
    if not r: continue
    if r[0] in ('File Path','File Name'): cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0].isdigit():
        try:
            si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
            out.append((int(r[si]), int(r[ii]), cur, int(r[0]), r[1][:120]))
        except Exception: pass
tot=sum(o[0] for o in out) or 1; toti=sum(o[1] for o in out) or 1
print('total samples',tot,'warp instr',toti)
for o in sorted(out,reverse=True)[:topn]: print("%5.1f%% smp %5.1f%% ins %s:%d  %s" % (100*o[0]/tot, 100*o[1]/toti, o[2], o[3], o[4]))
