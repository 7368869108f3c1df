"""SASS evidence of the Blackwell-native kernels: per-kernel counts of the tensor-core / TMEM / TMA mnemonics in libamb200.so
(`cuobjdump -sass`; the PTX names never appear in SASS — tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG,
cp.async.bulk = UBLKCP; HMMA would be the legacy mma.sync path).  Runs on the CPU build box:
    python tools/sass_census.py > profiles/r2_sass_tcgen05.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "afford-motion_b200", "csrc", "libamb200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|LDTM|STTM|UTMALDG(?:\.[0-9A-Z.]+)?|UTMASTG(?:\.[0-9A-Z.]+)?|UBLKCP|UTCBAR(?:\.[0-9A-Z.]+)?|HMMA|SYNCS|FFMA2|MUFU\.EX2)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            name = re.sub(r"^void ", "", name)
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        for tok in PAT.findall(ln):
            cur[re.sub(r"\.(2D|3D|1D)", "", tok)] += 1
        cur["_instructions"] += 1 if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln) else 0
    tot = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; {len(per)} kernels)")
    print("# kernel | SASS instructions | tensor-core / TMEM / TMA mnemonics")
    for k, c in per.items():
        keys = {a: b for a, b in c.items() if a != "_instructions" and a not in ("MUFU.EX2", "SYNCS")}
        tot.update(keys)
        if any(a.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP")) for a in keys):
            print(f"{k} | {c['_instructions']} | " + ", ".join(f"{a} x{b}" for a, b in sorted(keys.items())))
    print("# totals: " + ", ".join(f"{a} x{b}" for a, b in sorted(tot.items())))
    print(f"# HMMA (legacy mma.sync) x{tot.get('HMMA', 0)}")


if __name__ == "__main__":
    sys.exit(main())
