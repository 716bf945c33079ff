"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (cuobjdump -sass of the shipped library).
Run in the build container:  python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "minimax-speech_b200", "libls_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "MUFU.EX2", "MUFU.TANH", "HMMA", "FFMA2"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur)  # drop the parameter list
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
print("SASS summary of minimax-speech_b200/libls_b200.so (sm_100a), per kernel; wgmma / mma.sync (HMMA) must be 0 in the tensor kernels")
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>9s}" for k in MN))
tot = collections.Counter()
for name, c in counts.items():
    if not any(c[k] for k in MN[:6]):
        continue
    short = name if len(name) <= 70 else name[:67] + "..."
    print(f"{short:70s} {c['instructions']:7d} " + " ".join(f"{c[k]:9d}" for k in MN))
    tot.update(c)
print(f"{'TOTAL (kernels with tcgen05 / TMA instructions)':70s} {tot['instructions']:7d} " + " ".join(f"{tot[k]:9d}" for k in MN))
