#!/bin/bash
# tools/sass_stats.sh <lib.so> <kernel-substring> — SASS opcode histogram of one kernel (whole body) and of its hottest loop.
lib=$1; pat=$2
cuobjdump -sass "$lib" | awk -v pat="$pat" '/Function :/{f=index($0,pat)>0} f' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -e 's#/\*[0-9a-f]*\*/##g' | awk '{$1=$1};1' > /tmp/sass_kernel.txt
echo "total instructions: $(wc -l < /tmp/sass_kernel.txt)"
python3 - <<'PY'
import re,collections
lines=[l.strip() for l in open("/tmp/sass_kernel.txt")]
# find backward branches -> loops; take the innermost loop with the largest body containing MUFU
addr=0
ins=[]
for l in lines:
    ins.append(l)
# addresses are implicit: 16 bytes per instruction.  The hot loop = the smallest loop holding two full Philox
# blocks of two photons per lane (16 events = 32 MUFU): the full-cohort instantiation of the walk loop
loops=[]
for i,l in enumerate(ins):
    m=re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)",l)
    if m:
        tgt=int(m.group(1),16)//16
        if tgt<=i: loops.append((tgt,i))
best=None
for (s,e) in loops:
    body=ins[s:e+1]
    if sum("MUFU" in b for b in body) >= 32:
        if best is None or (e-s)<(best[1]-best[0]): best=(s,e)
if best:
    s,e=best
    body=ins[s:e+1]
    print(f"hot loop: instructions {s}..{e} = {e-s+1}")
    c=collections.Counter()
    for b in body:
        b=re.sub(r"^@!?U?P\d+\s+","",b)
        op=b.split()[0].rstrip(";")
        c[op.split(".")[0] + ("."+op.split(".")[1] if op.startswith(("IMAD","MUFU","FFMA2","ISETP")) and "." in op else "")]+=1
    for k,v in c.most_common(): print(f"  {v:4d} {k}")
PY
