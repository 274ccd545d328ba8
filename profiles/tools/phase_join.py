import csv,re,collections,sys
sass,srccsv=sys.argv[1],sys.argv[2]
amap={}
cur=None
fn=None
for l in open(sass):
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\.text\.(\S+):',l)
    if m: fn=m.group(1); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m and fn and 'solve_kernel' in fn:
        amap[int(m.group(1),16)]=(cur,m.group(2).strip())
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; data=rows[2:]
iA=hdr.index('Address'); iS=hdr.index('Source'); iSm=hdr.index('# Samples'); iN=hdr.index('Instructions Executed')
names=('stall_barrier','stall_short_sb','stall_wait','stall_long_sb','stall_no_inst','stall_selected','stall_branch_resolving','stall_math','stall_mio','stall_not_selected','stall_lg','stall_dispatch')
cols={h:hdr.index(h) for h in names}
base=None
per=collections.defaultdict(collections.Counter)
tot=0; matched=0; n=0
for r in data:
    try: a=int(r[iA],16)
    except: continue
    if base is None: base=a
    off=a-base
    sm=int(r[iSm]); tot+=sm; n+=1
    if off in amap:
        (cur,txt)=amap[off]
        t=r[iS].split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        if op in txt: matched+=1
        key=cur if cur else ('?',0)
        per[key]['samples']+=sm; per[key]['inst']+=int(r[iN])
        for h,i in cols.items():
            try: per[key][h]+=int(r[i])
            except: pass
print('total samples',tot,'rows',n,'matched',matched)
PER_IT=float(sys.argv[3]) if len(sys.argv)>3 else 1024*271.118
rng=[(0,292,'setup'),(293,392,'scaling'),(393,781,'refactor'),(782,845,'helpers'),(846,899,'a'),(900,923,'b1'),(924,937,"b1'"),(938,976,'b2'),(977,1001,"b2'"),(1002,1033,'b3'),(1034,1068,'c'),(1069,1310,'check'),(1311,1400,'loop/epilogue')]
def phase(f,ln):
    if f!='pqp_kp_core3.cuh': return f
    for a,b,nm in rng:
        if a<=ln<=b: return nm
    return 'other'
ph=collections.defaultdict(collections.Counter)
for (f,ln),c in per.items():
    for k,v in c.items(): ph[phase(f,ln)][k]+=v
for p,c in sorted(ph.items(), key=lambda kv:-kv[1]['samples']):
    s=max(1,c['samples'])
    print(f"{p:22s} samples {100*c['samples']/tot:5.1f}%  inst/it {c['inst']/PER_IT:7.0f}  barrier {100*c['stall_barrier']/s:3.0f}% short_sb {100*c['stall_short_sb']/s:3.0f}% wait {100*c['stall_wait']/s:3.0f}% long_sb {100*c['stall_long_sb']/s:3.0f}% sel {100*c['stall_selected']/s:3.0f}% noinst {100*c['stall_no_inst']/s:3.0f}% branch {100*c['stall_branch_resolving']/s:3.0f}% mio {100*c['stall_mio']/s:3.0f}% math {100*c['stall_math']/s:3.0f}%")
print()
for (f,ln),c in sorted(per.items(), key=lambda kv:-kv[1]['samples'])[:22]:
    print(f,ln,f"{100*c['samples']/tot:.1f}%", 'inst/it %.0f'%(c['inst']/PER_IT),'bar',c['stall_barrier'],'ssb',c['stall_short_sb'],'wait',c['stall_wait'],'sel',c['stall_selected'])
