"""SASS evidence: per-kernel histogram of the opcodes that matter on sm_100a, from the in-tree libkge_b200.so.

    python profiles/sass_hist.py > profiles/<tag>_sass.md

UTC*MMA = tcgen05.mma (gdesc = operand from shared memory, tmem = operand from tensor memory), LDTM/STTM = tcgen05.ld/st,
UTMALDG = TMA tile load (cp.async.bulk.tensor), UBLKCP = cp.async.bulk, UBLKRED = cp.reduce.async.bulk (row scatter of the
update, into local or peer HBM), FFMA2/FMUL2/FADD2 = packed fp32x2 (RotatE), SYNCS = mbarrier ops, REDG/ATOMG = global
reductions / atomics (".SYS" = system scope: NVLink peers), MUFU = special-function unit, HMMA would be the legacy path."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dgl-ke_b200", "lib", "libkge_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA(tmem A)", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UBLKRED", "SYNCS", "UTCBAR", "REDG", "REDG.SYS", "ATOMG", "LDG",
        "STG", "LDS", "STS", "MUFU", "FFMA", "FFMA2", "FMUL2", "FADD2", "HMMA", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::|kge::|^void ", "", fn)
            fn = fn.split("(")[0] if "<" not in fn else fn[:fn.index(">(") + 1] if ">(" in fn else fn[:90]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);", line)
        if not (m and fn):
            continue
        op, args = m.group(1), m.group(2)
        base = op.split(".")[0]
        hist[fn][base] += 1
        if base == "UTCHMMA" and args.strip().startswith("tmem"):
            hist[fn]["UTCHMMA(tmem A)"] += 1
        if base == "REDG" and ".SYS" in op:
            hist[fn]["REDG.SYS"] += 1
        hist[fn]["_total"] += 1
    print("# SASS opcode histogram of dgl-ke_b200/lib/libkge_b200.so (cuobjdump -sass), one row per kernel\n")
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for fn, c in hist.items():
        if c["_total"] < 20:
            continue
        print("| `%s` | %d | " % (fn[:70], c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")


if __name__ == "__main__":
    main()
