#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in an object file with their instruction mix.
    python tools/sass_loops.py opticalflowhs_b200/build/hs_stream_t4.o 'k_jacobi_streamILi4ELi0ELb0' [min_len] [max_len]
Only innermost candidates are interesting: pass max_len to hide the enclosing loops."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
maxlen = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ins = None, []
for l in names.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(ins)}
print(f"{len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr and minlen <= i - addr[tgt] + 1 <= maxlen:
            c = collections.Counter()
            for _, x in ins[addr[tgt]:i + 1]:
                x = re.sub(r"^@!?U?P\d+\s+", "", x)
                c[x.split()[0].split(".")[0]] += 1
            print(hex(tgt), hex(a), i - addr[tgt] + 1, dict(c.most_common()))
