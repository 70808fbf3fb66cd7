#!/usr/bin/env python3
"""Static SASS instruction count per CUDA source line for one kernel of the built object.

usage: sass_lines.py <kernel-name-substring> [top_n]     (needs -lineinfo; reads analiticcl_b200/build/kernels.cu.o)
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "analiticcl_b200", "build", "kernels.cu.o")
want = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=td, capture_output=True, text=True).stdout
fn, line, fname = None, 0, "?"
cnt, tot = {}, 0
for l in dis.splitlines():
    m = re.match(r"^\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fname, line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if fn and want in fn and re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[(fname, line)] = cnt.get((fname, line), 0) + 1
        tot += 1
print(f"{want}: {tot} SASS instructions = {tot * 16 / 1024:.1f} KB")
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{c:6d}  {f}:{ln}")
