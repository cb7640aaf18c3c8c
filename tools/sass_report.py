"""Static per-kernel summary of the built library (no GPU): registers / shared memory from cuobjdump --dump-resource-usage and the
counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (UTCHMMA, UTMALDG, UTMAPF, LDTM, UTCBAR) or its absence.
    python tools/sass_report.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dclip_b200  # noqa: E402

lib = dclip_b200.lib_path()
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True, check=True).stdout
usage = {}
fn = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        fn = m.group(1)
    elif fn and "REG:" in line:
        usage[fn] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line)}
        fn = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTCHMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "HMMA", "FCHK")
counts = collections.defaultdict(collections.Counter)
total = collections.Counter()
op = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
for line in sass.splitlines():
    if "Function :" in line:
        fn = line.split("Function :", 1)[1].strip()
        continue
    m = op.match(line)
    if m:
        total[fn] += 1
        if m.group(1) in KEYS:
            counts[fn][m.group(1)] += 1
demangled = dict(zip(total, subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()))
print(f"# {os.path.relpath(lib)}: {len(total)} kernels (sm_100a SASS); static, no GPU needed: python tools/sass_report.py")
print("# regs  smem_static  stack  instrs  " + "  ".join(f"{k:>7s}" for k in KEYS) + "  kernel")
for fn in sorted(total, key=lambda f: (-counts[f]["UTCHMMA"], demangled[f])):
    u = usage.get(fn, {})
    name = re.sub(r"\(.*", "", demangled[fn]).replace("void ", "").replace("dclip::", "")
    print(f"{u.get('REG', 0):6d}  {u.get('SHARED', 0):11d}  {u.get('STACK', 0):5d}  {total[fn]:6d}  " +
          "  ".join(f"{counts[fn][k]:7d}" for k in KEYS) + f"  {name}")
