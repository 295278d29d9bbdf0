#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05.mma = UTCHMMA, tcgen05.ld / st =
LDTM / STTM, TMA loads / stores = UTMALDG / UTMASTG, legacy mma.sync = HMMA (must be 0).  Usage: python tools/sass_summary.py [lib.so]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "complex_prompt_diffusion_b200/libcpd_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU.EX2", "FMNMX3", "LDS", "STS", "LD.E", "ST.E"]
counts, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\((anonymous namespace)?.*$", "", name)[:70]
        counts[name] = dict.fromkeys(MN, 0)
        continue
    if name is None:
        continue
    for k in MN:
        if re.search(r"(?<![A-Z.])" + re.escape(k) + r"(?![A-Z0-9])", line):
            counts[name][k] += 1
print(f"# cuobjdump -sass {lib}: SASS mnemonic counts per kernel (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA)")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in MN))
tot = dict.fromkeys(MN, 0)
for n in sorted(counts):
    c = counts[n]
    for k in MN:
        tot[k] += c[k]
    print(f"{n:70s} " + " ".join(f"{c[k]:8d}" for k in MN))
print(f"{'TOTAL':70s} " + " ".join(f"{tot[k]:8d}" for k in MN))
