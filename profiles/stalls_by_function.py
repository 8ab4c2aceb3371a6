#!/usr/bin/env python3
"""Aggregate the source page of an `ncu --set full --import-source on` report by routine of kernel_warp.cu:
share of warp-state samples, share of executed instructions and the stall mix of every routine.

    python profiles/stalls_by_function.py gpurun_out/prof.ncu-rep parseoggvorbis_b200/csrc/kernel_warp.cu > profiles/rNN_..._stalls_by_function.txt
(the source file must be the one the profiled library was built from: routines are found by line number)"""
import csv,sys,collections,subprocess,re
rep=sys.argv[1]; src=sys.argv[2]
lines=open(src).read().split('\n')
ranges=[]
for i,l in enumerate(lines,1):
    m=re.match(r'^(?:static )?__(?:device|global)__.*?\b(\w+)\(',l)
    if m and not l.startswith('\t'): ranges.append((i,m.group(1)))
def fn_of(line):
    name='?'
    for st,n in ranges:
        if st<=line: name=n
    return name
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass,cuda"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur_file=None; hdr=None
stalls=['stall_barrier','stall_branch_resolving','stall_dispatch','stall_lg','stall_long_sb','stall_math','stall_mio','stall_no_inst','stall_not_selected','stall_selected','stall_short_sb','stall_wait']
agg=collections.defaultdict(lambda: collections.Counter())
cur_key=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur_file=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No":
        hdr=r; idx={s:hdr.index(s) for s in stalls}; iI=hdr.index("Instructions Executed"); continue
    if hdr is None or len(r)<10: continue
    if r[0]!="": cur_key=(cur_file,int(r[0])); continue
    f = fn_of(cur_key[1]) if cur_key[0]==src.split('/')[-1] else cur_key[0]
    if f in('fft_core.cuh','sm_100_rt.hpp'): f='fft(dft8/cmul)'
    for s in stalls:
        try: agg[f][s]+=int(r[idx[s]])
        except: pass
    try: agg[f]['inst']+=int(r[iI])
    except: pass
tot=sum(sum(v[s] for s in stalls) for v in agg.values())
print("%-22s %7s %6s | "%("function","samples","inst%")+" ".join(s.replace('stall_','')[:7].rjust(7) for s in stalls))
ti=sum(v['inst'] for v in agg.values())
for f,v in sorted(agg.items(), key=lambda kv:-sum(kv[1][s] for s in stalls))[:16]:
    n=sum(v[s] for s in stalls)
    print("%-22s %6.1f%% %5.1f%% | "%(f[:22],100*n/tot,100*v['inst']/ti)+" ".join(("%d"%(100*v[s]/max(1,n))).rjust(7) for s in stalls))
