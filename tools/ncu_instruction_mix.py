import csv, gzip, sys, collections, re
csv.field_size_limit(1<<30)
path=sys.argv[1]
rows=list(csv.reader(gzip.open(path,'rt')))
# may contain multiple kernels; take the one with the most instructions
tables=[];cur=None
for r in rows:
    if r and r[0]=="Kernel Name":
        cur={'name':r[1],'hdr':None,'rows':[]}; tables.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    cur['rows'].append(r)
for t in tables:
    h=t['hdr']; ia=h.index("Source"); ie=h.index("Instructions Executed"); isamp=h.index("# Samples")
    tot=sum(float(r[ie]) for r in t['rows'] if len(r)>ie); ts=sum(float(r[isamp]) for r in t['rows'] if len(r)>isamp)
    if tot==0: continue
    print("==",t['name'][:90],"warp-instr executed",int(tot),"samples",int(ts))
    byop=collections.Counter(); bys=collections.Counter()
    for r in t['rows']:
        if len(r)<=ie: continue
        op=r[ia].strip().split()
        if not op: continue
        o=op[0]
        if o.startswith('@'): o=op[1]
        o=o.split('.')[0]
        byop[o]+=float(r[ie]); bys[o]+=float(r[isamp])
    for o,v in byop.most_common(22):
        print(f"   {o:10s} {100*v/tot:5.1f}% instr   {100*bys[o]/ts:5.1f}% samples")
    stall=[c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    sc=collections.Counter()
    for r in t['rows']:
        for c in stall:
            sc[c]+=float(r[h.index(c)] or 0)
    print("   stalls:", ", ".join(f"{k[6:]} {100*v/ts:.0f}%" for k,v in sc.most_common(8)))
