"""profiles/sass_opcodes.txt: per-kernel counts of the Blackwell-specific SASS opcodes in the shipped libmilb200.so
(cuobjdump -sass; no GPU needed): tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM), TMA loads/stores (UTMALDG / UTMASTG), cluster
barriers / multicast commits (UTCBAR), multimem (NVSwitch) loads/stores, plus total instruction count.
Usage: python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "llm-guided-multimodal-mil_b200", "libmilb200.so")
PAT = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
KEYS = [("UTC*MMA (tcgen05.mma)", re.compile(r"^UTC\w*MMA")), ("  of which .2CTA", re.compile(r"^UTC\w*MMA.*2CTA")),
        ("LDTM (tcgen05.ld)", re.compile(r"^LDTM")), ("UTMALDG (TMA load)", re.compile(r"^UTMALDG")),
        ("UTMASTG (TMA store)", re.compile(r"^UTMASTG")), ("UTCBAR (tcgen05.commit)", re.compile(r"^UTCBAR")),
        ("multimem ld_reduce/st", re.compile(r"MULTIMEM|LDGMC|STGMC|REDGMC|\.MC")), ("HMMA/QMMA (legacy mma.sync)", re.compile(r"^[HQI]MMA")),
        ("FFMA", re.compile(r"^FFMA"))]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, totals = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[kern] = collections.Counter()
            continue
        m = PAT.match(line)
        if m and kern:
            op = m.group(1)
            counts[kern]["instructions"] += 1
            for name, rx in KEYS:
                if rx.search(op):
                    counts[kern][name] += 1
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# columns: " + " | ".join(k for k, _ in KEYS) + " | instructions")
    for k, c in counts.items():
        if not any(c[name] for name, _ in KEYS[:8]):
            continue
        print(f"{k[-110:]:110s} " + " ".join(f"{c[name]:5d}" for name, _ in KEYS) + f" {c['instructions']:6d}")
        for name, _ in KEYS:
            totals[name] += c[name]
    print("TOTAL".ljust(110) + " " + " ".join(f"{totals[name]:5d}" for name, _ in KEYS))
    print(f"# kernels in the library: {len(counts)}")


if __name__ == "__main__":
    main()
