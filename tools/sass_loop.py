#!/usr/bin/env python
"""Opcode histogram of the largest loop of one kernel in libgokalman_b200.so (cuobjdump -sass).
usage: tools/sass_loop.py <mangled-function-name> [max_loop_instrs]"""
import collections, re, subprocess, sys
fun = sys.argv[1]
maxlen = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, "gokalman_b200/libgokalman_b200.so"], capture_output=True, text=True).stdout
ins = []
for l in txt.splitlines():
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print("total instrs", len(ins))
backs = []
for addr, t in ins:
    m = re.search(r'BRA\S*\s+.*?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < addr:
        backs.append((addr, int(m.group(1), 16)))
        print(hex(addr), t, "-> back", (addr - int(m.group(1), 16)) // 16)
cands = [b for b in backs if (b[0] - b[1]) // 16 <= maxlen]
a, b = max(cands, key=lambda x: x[0] - x[1])
body = [t for ad, t in ins if b <= ad <= a]
ops = collections.Counter(re.sub(r'^@!?U?P\d\s+', '', t).split()[0] for t in body)
print("loop", hex(b), hex(a), len(body))
fp64 = sum(v for k, v in ops.items() if k.startswith(("DFMA", "DADD", "DMUL", "DSETP", "MUFU.R")))
print("fp64-pipe-ish", fp64)
for k, v in ops.most_common(50):
    print("%-28s %d" % (k, v))
