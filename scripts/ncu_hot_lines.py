import csv, sys, re
# usage: hot.py <ncu-rep> <kernel-regex>
import subprocess
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
h=None
for i,r in enumerate(rows):
    if 'Instructions Executed' in r: h=i; break
hdr=rows[h]
ii=hdr.index('Instructions Executed'); wi=hdr.index('Warp Stall Sampling (All Samples)')
si=hdr.index('Source')
# rows: cuda lines have "Line No"? figure structure: print first few
agg={}
cur=None
for r in rows[h+1:]:
    if len(r)<=ii: continue
    src=r[si]
    first=r[0]
    # cuda source rows have numeric line in col0 and no address
    if not first.startswith('0x'):
        cur=(first, src.strip()[:100]); agg.setdefault(cur,[0,0]); 
        try:
            agg[cur][0]+=int(r[ii] or 0); agg[cur][1]+=int(r[wi] or 0)
        except: pass
tot=sum(v[0] for v in agg.values()); tots=sum(v[1] for v in agg.values())
print('total inst',tot,'samples',tots, 'lines', len(agg))
print('--- top by instructions')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:32]: print('%6.2f%% %6.2f%% L%s %s'%(100*v[0]/max(tot,1),100*v[1]/max(tots,1),k[0],k[1]))
print('--- top by samples')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print('%6.2f%% %6.2f%% L%s %s'%(100*v[0]/max(tot,1),100*v[1]/max(tots,1),k[0],k[1]))
