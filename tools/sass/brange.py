import re,subprocess,sys
cubin=sys.argv[1]; pat=sys.argv[2]
out=subprocess.run(['cuobjdump','-sass',cubin],capture_output=True,text=True).stdout
cur=None; lines=[]
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); continue
    if cur and pat in cur:
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',line)
        if m: lines.append((int(m.group(1),16),m.group(2)))
# main loop: backward branch immediately followed by EXIT
for i,(a,t) in enumerate(lines[:-1]):
    m=re.search(r'BRA 0x([0-9a-f]+)',t)
    if m and int(m.group(1),16)<a and lines[i+1][1].strip().startswith('EXIT'):
        print(f"{int(m.group(1),16):x} {a+0x10:x}"); break
