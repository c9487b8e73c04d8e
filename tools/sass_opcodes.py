"""Counts of the Blackwell-specific SASS opcodes per kernel of libvyom_b200.so (cuobjdump -sass): the evidence that the
tensor-core kernels issue tcgen05 MMAs (UTCHMMA / UTCQMMA ...), read TMEM (LDTM), and move tiles with TMA (UTMALDG / UTMASTG /
UBLKCP), per /opt/skills/guides/B200_PROFILING.md. Writes a markdown table to stdout.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vyomai_b200", "csrc", "libvyom_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "MUFU.EX2"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name)
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    print("# SASS opcode counts per kernel (cuobjdump -sass vyomai_b200/csrc/libvyom_b200.so, sm_100a)\n")
    print("Instantiations of one template are summed; kernels with none of the listed opcodes are collected in the last row.\n")
    agg = collections.OrderedDict()
    for k, c in counts.items():
        base = re.sub(r"<.*", "", k).replace("void ", "").replace("vy::", "")
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(c)
    print("| kernel | instantiations | instructions | " + " | ".join(OPS) + " |")
    print("|---|---|---|" + "---|" * len(OPS))
    rest = []
    for k, (n, c) in agg.items():
        if not any(c[o] for o in OPS):
            rest.append(k)
            continue
        print(f"| {k} | {n} | {c['_all']} | " + " | ".join(str(c[o]) for o in OPS) + " |")
    print(f"\nKernels with none of these opcodes (plain LDG/STG streaming kernels): {', '.join(sorted(rest))}")


if __name__ == "__main__":
    sys.exit(main())
