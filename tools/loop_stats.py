#!/usr/bin/env python3
"""Static opcode mix of the inner tick loop of each vdt_rollout_fast_kernel instantiation."""
import re,collections,sys,subprocess
so = sys.argv[1] if len(sys.argv)>1 else "roboken-fmskf-robot-controller_b200/librobotick_b200.so"
txt=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout
parts=re.split(r'\n\s*Function : ', txt)
for p in parts[1:]:
    name=p.split('\n')[0]
    if 'vdt_rollout_fast_kernelILb0' in name or 'full_rollout' in name:
        lines=[l for l in p.split('\n') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/',l)]
        addr=lambda l:int(re.match(r'\s+/\*([0-9a-f]+)\*/',l).group(1),16)
        for l in lines:
            if 'BRA' in l:
                m2=re.search(r'0x([0-9a-f]+)\s*;',l)
                if m2:
                    tgt=int(m2.group(1),16)
                    if tgt<addr(l) and 100<(addr(l)-tgt)//16<900:
                        body=[x for x in lines if tgt<=addr(x)<=addr(l)]
                        ops=collections.Counter()
                        for x in body:
                            t=x.split('*/',1)[1].strip().rstrip(';').split()
                            op=t[1] if t[0].startswith('@') else t[0]
                            ops[op.split('.')[0]]+=1
                        m=re.search(r'ELi(\d)E',name)
                        print(name[:60], "occ", m.group(1) if m else '?', "loop instrs", len(body), dict(ops.most_common(16)))
