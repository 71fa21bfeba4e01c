#!/usr/bin/env python3
"""Per-basic-block view of an ncu source capture: consecutive SASS instructions with equal execution counts are one block;
prints instruction count, executions, lanes, issue share, stall-sample share, long-scoreboard share and source lines.
    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/sass_blocks.py <cubin> <kernel-substring> sass.csv [min issue share %]"""
import csv,re,subprocess,sys
cubin,pat,ncu=sys.argv[1:4]
txt=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout
cur=None;line=None;file_=None;insts=[]
for ln in txt.splitlines():
    m=re.match(r"\s*\.text\.(\S+):",ln)
    if m: cur=m.group(1);continue
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',ln)
    if m: file_,line=m.group(1).split("/")[-1],int(m.group(2));continue
    if cur and pat in cur:
        m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",ln)
        if m: insts.append((int(m.group(1),16),file_,line,m.group(2)))
rows=list(csv.reader(open(ncu)));h=rows[1];d=rows[2:]
ia,ie,it,ns,lsb=[h.index(k) for k in ["Address","Instructions Executed","Thread Instructions Executed","# Samples","stall_long_sb"]]
base=int(d[0][ia],16)
W={int(r[ia],16)-base:(int(r[ie]),int(r[it]),int(r[ns]),int(r[lsb])) for r in d}
tot_e=sum(w[0] for w in W.values());tot_s=sum(w[2] for w in W.values())
blocks=[];cb=None
o0=insts[0][0]
for off,f,l,t in insts:
    w=W.get(off-o0,(0,0,0,0))
    if cb and cb['e']==w[0] and cb['t']==w[1]:
        cb['n']+=1;cb['s']+=w[2];cb['l']+=w[3];cb['lines'].add((f[:14].replace('volpath_',''),l))
    else:
        cb={'e':w[0],'t':w[1],'n':1,'s':w[2],'l':w[3],'lines':{(f[:14].replace('volpath_',''),l)},'first':t};blocks.append(cb)
for b in blocks:
    share=100*b['e']*b['n']/tot_e
    if share<float(sys.argv[4] if len(sys.argv)>4 else 0.3): continue
    ls=sorted(b['lines'])
    print("n=%3d ex/1e6=%6.0f thr=%5.1f issue=%5.2f%% smp=%5.2f%% lsb=%5.2f%% %s"%(b['n'],b['e']/1e6,b['t']/max(b['e'],1),share,100*b['s']/tot_s,100*b['l']/tot_s," ".join("%s:%d"%x for x in ls)[:150]))
