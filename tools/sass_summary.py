#!/usr/bin/env python3
"""Per-kernel SASS opcode summary of the built library (cuobjdump -sass analiticcl_b200/libanaliticcl_b200.so).

usage: sass_summary.py [out.txt]

For every kernel: total SASS instructions, code size, registers, and the counts of the opcode families that matter
here (global/shared loads and stores, atomics, TMA bulk copies UBLKCP + mbarrier SYNCS, warp votes/shuffles/match,
f64 arithmetic, byte/SIMD min-max, integer multiply-add).  Committed under profiles/ per round.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "analiticcl_b200", "libanaliticcl_b200.so")
FAMILIES = [
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDC", r"^LDC|^ULDC"), ("LDL/STL", r"^LDL|^STL"),
    ("ATOM/RED", r"^ATOM|^RED|^ATOMS|^ATOMG"), ("UBLKCP", r"^UBLKCP"), ("UTMALDG", r"^UTMA"), ("SYNCS", r"^SYNCS"),
    ("VOTE/MATCH/SHFL", r"^VOTE|^MATCH|^SHFL|^REDUX"), ("BAR", r"^BAR|^WARPSYNC"),
    ("F64", r"^DADD|^DMUL|^DFMA|^DSETP|^MUFU\.RCP64H|^F2F\.F64|^I2F\.F64"), ("VIMNMX/VIADD", r"^VIMNMX|^VIADD|^VABSDIFF"),
    ("IMAD/IADD3/LOP3", r"^IMAD|^IADD3|^LOP3|^LEA|^SHF|^PRMT"), ("BRA/CALL", r"^BRA|^CALL|^RET|^BSSY|^BSYNC|^EXIT"),
]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+).*?SHARED:(\d+)", res):
    regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
fn = None
ops = collections.defaultdict(collections.Counter)
arch = None
for line in sass.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if fn and m:
        ops[fn][m.group(1)] += 1


def short(name):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = d.replace("(anonymous namespace)::", "").replace("anl::", "").replace("void ", "")
    d = re.sub(r"\(.*", "", d)
    return d


out = [f"SASS summary of analiticcl_b200/libanaliticcl_b200.so ({arch}); cuobjdump -sass / -res-usage", ""]
hdr = f"{'kernel':44s} {'instr':>7s} {'KB':>6s} {'regs':>4s} {'smem':>6s}  " + " ".join(f"{n:>8s}" for n, _ in FAMILIES)
out.append(hdr)
for fn_ in sorted(ops, key=lambda f: -sum(ops[f].values())):
    c = ops[fn_]
    tot = sum(c.values())
    fam = []
    for _, rx in FAMILIES:
        fam.append(sum(v for k, v in c.items() if re.match(rx, k)))
    r = regs.get(fn_, (0, 0))
    nm = short(fn_)
    if len(nm) > 44:
        nm = nm[:41] + "..."
    out.append(f"{nm:44s} {tot:7d} {tot * 16 / 1024:6.1f} {r[0]:4d} {r[1]:6d}  " + " ".join(f"{v:8d}" for v in fam))
out.append("")
out.append("TMA / mbarrier evidence (kernels with UBLKCP or SYNCS):")
for fn_ in ops:
    tma = {k: v for k, v in ops[fn_].items() if k.startswith(("UBLKCP", "UTMA", "SYNCS"))}
    if tma:
        out.append(f"  {short(fn_)}: " + ", ".join(f"{k} x{v}" for k, v in sorted(tma.items())))
text = "\n".join(out) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
