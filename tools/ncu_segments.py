#!/usr/bin/env python3
"""Warp-stall samples of a kernel per barrier-delimited segment of its SASS.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; ncu_segments.py src.csv"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,'# Samples') for r in data)
print('total samples',tot, 'instrs', len(data))
stalls=['stall_barrier','stall_dispatch','stall_lg','stall_long_sb','stall_math','stall_mio','stall_no_inst','stall_not_selected','stall_selected','stall_short_sb','stall_wait','stall_branch_resolving']
seg=[];new=lambda i:{'start':i,'n':0,'samples':0,'exec':0,**{s:0 for s in stalls}}
cur=new(0)
for i,r in enumerate(data):
    src=r[ix['Source']]
    cur['n']+=1; cur['samples']+=f(r,'# Samples'); cur['exec']+=f(r,'Instructions Executed')
    for s in stalls: cur[s]+=f(r,s)
    if 'BAR.SYNC' in src or src.strip().startswith('BAR'):
        cur['end']=i; seg.append(cur); cur=new(i+1)
cur['end']=len(data)-1; seg.append(cur)
for s in seg:
    print(f"instr {s['start']:5d}-{s['end']:5d} n={s['n']:5d} exec={s['exec']/1e6:8.2f}M samples={s['samples']:8.0f} ({100*s['samples']/tot:5.1f}%) ", ' '.join(f"{k[6:]}={100*s[k]/max(1,s['samples']):.0f}" for k in stalls if s[k]/max(1,s['samples'])>0.04))
