#!/usr/bin/env python3
"""Static view of a kernel's loops from `cuobjdump -sass`: every backward branch with the opcode mix of the
instructions it spans.   python tools/sass_loops.py file.cubin <kernel-substring> [min_instr]"""
import collections
import re
import subprocess
import sys

cubin, key = sys.argv[1], sys.argv[2]
min_instr = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
fn, body = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        body[fn] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn:
        body[fn].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in body.items():
    if key not in fn:
        continue
    print(fn, len(ins), "instructions")
    addr = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            span = ins[addr[tgt]:k + 1]
            if len(span) < min_instr:
                continue
            ops = collections.Counter()
            for _, s in span:
                w = s.split()
                op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
                ops[op] += 1
            print(f"  loop {tgt:#x}..{a:#x}: {len(span)} instr:", ", ".join(f"{o} {c}" for o, c in ops.most_common()))
