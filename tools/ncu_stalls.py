"""Per-CUDA-source-line stall REASONS of one kernel in an .ncu-rep (needs -lineinfo and --import-source on):
   python tools/ncu_stalls.py rep.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname=""; hdr=None; agg=[]
for r in rows:
    if r and r[0]=="File Name": fname=r[1].split("/")[-1]
    if "Instructions Executed" in r: hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0].strip().isdigit():
        d={h:r[i] for i,h in enumerate(hdr)}
        agg.append((fname,int(r[0]),r[1].strip(),d))
keys=['stall_no_inst','stall_wait','stall_long_sb','stall_short_sb','stall_not_selected','stall_selected','stall_dispatch','stall_math','stall_branch_resolving']
def I(d,k):
    try: return int(d[k])
    except: return 0
tots=sum(I(d,'Warp Stall Sampling (All Samples)') for _,_,_,d in agg); tot=sum(I(d,'Instructions Executed') for _,_,_,d in agg)
print('samples',tots,'inst',tot)
print('%6s %6s | '%('smp%','ins%')+' '.join(k[6:12].rjust(6) for k in keys))
for f,ln,src,d in sorted(agg,key=lambda a:-I(a[3],'Warp Stall Sampling (All Samples)'))[:int(sys.argv[3]) if len(sys.argv)>3 else 60]:
    print('%6.2f %6.2f | '%(100*I(d,'Warp Stall Sampling (All Samples)')/tots,100*I(d,'Instructions Executed')/tot)+' '.join('%6d'%I(d,k) for k in keys)+f'  {f}:{ln} {src[:80]}')
