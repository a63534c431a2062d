import re,sys,subprocess,collections
cubin=sys.argv[1]; pat=sys.argv[2] if len(sys.argv)>2 else ''
nb=float(sys.argv[3]) if len(sys.argv)>3 else 32
lo=int(sys.argv[4],16) if len(sys.argv)>4 else 0
hi=int(sys.argv[5],16) if len(sys.argv)>5 else 1<<30
out=subprocess.run(['cuobjdump','-sass',cubin],capture_output=True,text=True).stdout
cur=None; funcs=collections.OrderedDict()
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); funcs[cur]=collections.Counter(); continue
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)',line)
    if m and cur and lo<=int(m.group(1),16)<hi:
        op=m.group(3)
        base=op.split('.')[0]
        if base=='IMAD':
            if '.WIDE' in op: k='IMAD.WIDE'
            elif '.HI' in op: k='IMAD.HI'
            elif '.MOV' in op: k='IMAD.MOV'
            elif '.X' in op: k='IMAD.X'
            elif '.IADD' in op: k='IMAD.IADD'
            elif '.SHL' in op: k='IMAD.SHL'
            else: k='IMAD'
        else: k=base
        funcs[cur][k]+=1
for f,c in funcs.items():
    if pat not in f: continue
    heavy4=c['IMAD.WIDE']+c['IMAD.HI']; heavy2=c['IMAD']+c['IMAD.MOV']+c['IMAD.X']+c['IMAD.IADD']+c['IMAD.SHL']+c['HFMA2']+c['FMUL']+c['FFMA']
    alu=sum(c[k] for k in ['IADD3','LOP3','SHF','ISETP','SEL','LEA','MOV','VIADD','PRMT','PLOP3','IABS','CS2R','VIMNMX'])
    tot=sum(c.values())
    one=c['IADD3']+c['VIADD']; two=alu-one
    mem=sum(v for k,v in c.items() if k in ('LDS','STS','LDG','STG','LDL','STL','LDC','LDCU','ATOMS','RED','SYNCS','UBLKCP'))
    rest=tot-(heavy4+heavy2+alu+mem)
    cost=4*heavy4+2*heavy2+one+2*two+mem+rest
    print(f"{f[:70]}\n  ISSUE-COST model (WIDE/HI 4, IMAD* 2, IADD3 1, other ALU 2, rest 1): {cost} ({cost/nb:.1f}/bfly)  [fma {4*heavy4+2*heavy2}, iadd3 {one}, alu2 {2*two}, mem {mem}, rest {rest}]")
    print(f"  total={tot} ({tot/nb:.1f}/bfly) heavy_cycles={4*heavy4+2*heavy2} ({(4*heavy4+2*heavy2)/nb:.1f}/bfly) alu={alu} ({alu/nb:.1f}/bfly)")
    print('  '+' '.join(f"{k}:{v}" for k,v in c.most_common(24)))
