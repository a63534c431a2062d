import re,subprocess,sys
# pass A with one item per CTA executes the peeled first copy of the item loop: from the first "@P0 EXIT" to the second
out=subprocess.run(['cuobjdump','-sass',sys.argv[1]],capture_output=True,text=True).stdout
cur=None; ex=[]
for line in out.splitlines():
    m=re.search(r'Function : (\S+)',line)
    if m: cur=m.group(1); continue
    if cur and sys.argv[2] in cur:
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+@P0 EXIT',line)
        if m: ex.append(int(m.group(1),16))
print(f"{ex[0]:x} {ex[1]+0x10:x}" if len(ex)>=2 else "")
