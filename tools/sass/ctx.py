# print the in-range instructions of a kernel matching a regex, with N lines of context: python ctx.py cubin kernelpat lo hi regex [ctx]
import re,sys,subprocess
cubin,pat,lo,hi,rx=sys.argv[1],sys.argv[2],int(sys.argv[3],16),int(sys.argv[4],16),sys.argv[5]
ctx=int(sys.argv[6]) if len(sys.argv)>6 else 3
out=subprocess.run(['cuobjdump','-sass',cubin],capture_output=True,text=True).stdout
cur=None; L=[]
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); continue
    if cur and pat in cur:
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',line)
        if m and lo<=int(m.group(1),16)<hi: L.append((m.group(1),m.group(2).strip()))
hits=[i for i,(a,t) in enumerate(L) if re.search(rx,t)]
shown=set()
for i in hits[:int(sys.argv[7]) if len(sys.argv)>7 else 12]:
    for j in range(max(0,i-ctx),min(len(L),i+ctx+1)):
        if j not in shown: print(('>> ' if j==i else '   ')+L[j][0],L[j][1]); shown.add(j)
    print('   --')
