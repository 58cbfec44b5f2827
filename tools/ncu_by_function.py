"""Stall samples of an ncu capture by source function.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_by_function.py src.csv [git revision of the captured tree]
file:line of every sampled source line -> the enclosing function of that file (inlined code is attributed to the function its
source line lives in); columns: share of the samples outside the barriers, share of the executed instructions, stall reasons."""
import csv, sys, collections, re, subprocess
path = sys.argv[1]
root = '/root/repo/automatedvaletparking_b200/csrc/'
rows = list(csv.reader(open(path, newline='')))
fname=None; hdr=None; lines=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if r[0]=='' or hdr is None: continue
    d=dict(zip(hdr[4:], r[4:]))
    def g(k):
        try: return float(d.get(k,'0') or 0)
        except ValueError: return 0.0
    lines.append((fname,int(r[0]),g('# Samples')-g('stall_barrier'),g('Instructions Executed'), g('stall_wait'), g('stall_no_inst'), g('stall_long_sb'), g('stall_short_sb'), g('stall_branch_resolving')))
# function boundaries from the CURRENT sources is wrong for old captures; use git show of a given rev if provided
rev = sys.argv[2] if len(sys.argv)>2 else None
funcs={}
import os
for f in set(l[0] for l in lines):
    p=root+f
    try:
        if rev: src=subprocess.run(['git','-C','/root/repo','show',rev+':automatedvaletparking_b200/csrc/'+f],capture_output=True,text=True).stdout.split('\n')
        else: src=open(p).read().split('\n')
    except Exception: continue
    b=[]
    for i,s in enumerate(src,1):
        m=re.match(r'^\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|AVP_HD|static).*?([A-Za-z_0-9]+)\s*\(',s)
        if m and not s.strip().startswith('//') and ('{' in s or s.rstrip().endswith(',') or s.rstrip().endswith(')')): b.append((i,m.group(1)))
    funcs[f]=b
def fn(f,l):
    b=funcs.get(f,[]); name='?'
    for i,n in b:
        if i<=l: name=n
        else: break
    return name
agg=collections.Counter(); inst=collections.Counter(); st=collections.defaultdict(lambda:[0,0,0,0,0])
for f,l,nb,ins,w,ni,lsb,ssb,br in lines:
    k=(f,fn(f,l)); agg[k]+=nb; inst[k]+=ins
    for j,v in enumerate((w,ni,lsb,ssb,br)): st[k][j]+=v
T=sum(agg.values()); TI=sum(inst.values())
print('non-barrier samples',T,'instructions %.1fG'%(TI/1e9))
for k,v in agg.most_common(45):
    s=st[k]
    print('%5.1f%%  inst %5.1f%%  wait %3.0f%% noinst %3.0f%% lsb %3.0f%% ssb %3.0f%% br %3.0f%%  %s:%s'%(100*v/T,100*inst[k]/TI,*(100*x/max(v,1) for x in s),k[0],k[1]))
