#!/usr/bin/env python3
"""tools/sass_loop_count.py <lib.so> <mangled kernel> -- static SASS view of a kernel (run here, no GPU): total instructions, backward
branches, and the opcode histogram of its largest loop (the factor-block loop of the row kernels)."""
import sys,re,collections,subprocess
lib,fn=sys.argv[1],sys.argv[2]
out=subprocess.run(['cuobjdump','-sass','-fun',fn,lib],capture_output=True,text=True).stdout
rows=[]
for l in out.splitlines():
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;',l)
    if m: rows.append((int(m.group(1),16),m.group(2)))
print('total',len(rows))
# find backward branches
back=[]
for a,ins in rows:
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)',ins)
    if m:
        t=int(m.group(1),16)
        if t<a: back.append((t,a))
print('backward branches',[(hex(t),hex(a),(a-t)//16) for t,a in back])
if back:
    t,a=max(back,key=lambda x:x[1]-x[0])
    c=collections.Counter()
    for x,ins in rows:
        if t<=x<=a:
            tk=ins.split()
            op=tk[1] if tk[0].startswith('@') else tk[0]
            c[op.split('.')[0]]+=1
    print('main loop',(a-t)//16+1,c.most_common(30))
open('/tmp/last.sass','w').write('\n'.join('%04x %s'%r for r in rows))
