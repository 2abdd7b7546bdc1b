"""SASS evidence: instruction counts per kernel of libpassport_sm100.so (cuobjdump -sass).  Usage:
    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepipr_b200", "lib", "libpassport_sm100.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "DFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, variants = None, collections.OrderedDict(), collections.defaultdict(set)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            base = op.split(".")[0]
            if base in OPS:
                counts[cur][base] += 1
                if base == "UTCHMMA":
                    variants[cur].add(op)
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence (cuobjdump -sass deepipr_b200/lib/libpassport_sm100.so, sm_100a): instruction counts per kernel")
    print("# UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32 share the opcode; the operand format is in the instruction")
    print("# descriptor), UTMALDG = cp.async.bulk.tensor (TMA loads), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
    print("# UTCATOMSWS = tcgen05.alloc / dealloc / relinquish, SYNCS = mbarrier ops.")
    total_hmma = 0
    for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
        c = counts[n]
        total_hmma += c["HMMA"]
        if c["UTCHMMA"] or c["UTMALDG"] or "stem" in d or "maxpool" in d:
            short = d.split("(")[0]
            print(f"{short:70s} " + " ".join(f"{k}={c[k]}" for k in OPS if c[k]) +
                  (f"   [{', '.join(sorted(variants[n]))}]" if variants[n] else ""))
    print(f"HMMA (mma.sync / wmma) instructions in the whole library: {total_hmma}")


if __name__ == "__main__":
    main()
