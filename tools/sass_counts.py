#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, bulk copies (TMA engine) -> UBLKCP/UTMALDG, tcgen05.commit -> UTCBAR; legacy paths HMMA / HGMMA
must be absent.      python tools/sass_counts.py > profiles/r02_sass_counts.txt     (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nerfmeshes_b200", "lib", "libnerfmeshes_b200.so")
MN = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "USETMAXREG", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "HGMMA", "LDGSTS", "DFMA", "ATOMS", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts, sizes = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = fn.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            fn = re.sub(r"^void ", "", fn)
            fn = re.sub(r"\(.*", "", fn)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            sizes[fn] += 1
            for k in MN:
                if op == k:
                    counts[fn][k] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel; columns with no hits anywhere are omitted")
    used = [k for k in MN if any(c[k] for c in counts.values())]
    print(f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k in used))
    for fn, c in counts.items():
        print(f"{fn[:58]:58s} {sizes[fn]:7d} " + " ".join(f"{c[k]:8d}" for k in used))
    legacy = sum(c["HMMA"] + c["HGMMA"] for c in counts.values())
    print(f"# legacy tensor-core instructions (HMMA / HGMMA): {legacy}")


if __name__ == "__main__":
    main()
