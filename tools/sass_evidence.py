"""Write profiles/<name>: per kernel of the in-tree .so, the SASS lines that carry the Blackwell-native instructions
(tcgen05 / tensor memory / TMA).  python tools/sass_evidence.py [profiles/r02_sass_blackwell_mnemonics.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rlsolver_b200", "_C", "librlsolver_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_blackwell_mnemonics.txt")
WANT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCCP|UTMAPF)\b")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = {}
for m in re.finditer(r"Function : (\S+)", sass):
    names[m.group(1)] = None
demangled = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(names, demangled))
out = ["SASS evidence of the Blackwell-native instructions in rlsolver_b200/_C/librlsolver_b200.so",
       "(cuobjdump -sass of the in-tree build for sm_100a; per kernel: instruction count, then every line that carries",
       " UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), UTCBAR (tcgen05.commit))",
       ""]
totals = collections.Counter()
cur, lines, hits = None, 0, []


def flush():
    if cur is not None and hits:
        out.append(f"== {names.get(cur, cur)}")
        out.append(f"   {lines} SASS lines, {len(hits)} tensor-memory / TMA instructions")
        out.extend("   " + h.strip() for h in hits)
        out.append("")


for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, lines, hits = m.group(1), 0, []
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        lines += 1
        w = WANT.search(line)
        if w:
            hits.append(line.split("/*", 2)[0] + "/*" + line.split("/*", 2)[1] if False else line[:line.rfind("/*")] if line.count("/*") > 1 else line)
            totals[w.group(1)] += 1
flush()
out.append("totals: " + ", ".join(f"{k} x {v}" for k, v in sorted(totals.items())))
open(out_path, "w").write("\n".join(out) + "\n")
print(out[-1])
