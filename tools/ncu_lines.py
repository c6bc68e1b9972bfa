"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report captured with
`--set full --import-source on` (library built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep tc_nn_kernel 30
"""
import csv,collections,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]; top=int(sys.argv[3]) if len(sys.argv)>3 else 25
out=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv','-k','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg=collections.defaultdict(lambda:[0.0,0.0,0.0]); hdr=None; src={}
def num(x):
    try: return float(x.replace(',',''))
    except: return 0.0
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': hdr=r; ix={n:i for i,n in enumerate(hdr)}; continue
    if hdr and len(r)==len(hdr):
        try: ln=int(r[0])
        except: continue
        key=(cur,ln); src[key]=r[1]
        if r[2]!='':
            a=agg[key]; a[0]+=num(r[ix['# Samples']]); a[1]+=num(r[ix['Instructions Executed']]); a[2]+=num(r[ix['Thread Instructions Executed']])
tot=sum(a[1] for a in agg.values()); tots=sum(a[0] for a in agg.values())
print("total warp instr %.3e samples %d avg lanes %.1f"%(tot,tots,sum(a[2] for a in agg.values())/tot))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:top]:
    print(f"{a[1]/tot*100:5.1f}% instr {a[0]/tots*100:5.1f}% samp lanes {a[2]/max(a[1],1):5.1f}  {k[0]}:{k[1]}  {src[k].strip()[:88]}")
