#!/usr/bin/env python3
"""Per-kernel SASS evidence of the built objects (no GPU needed): for every kernel of an object file, the
instances of the Blackwell-specific mnemonics (tcgen05 = UTC*, LDTM / STTM, TMA = UTMA* / UBLKCP, mbarrier = SYNCS,
ELECT) and the resource usage `cuobjdump -res-usage` reports (registers, shared memory, stack = spills).

    python tools/sass_summary.py fanlin-rs_b200/csrc/build/kernels_fused_tc3.o > profiles/r02/sass_tc3_mnemonics.txt

The names follow /opt/skills/guides/B200_PROFILING.md: UTCIMMA = tcgen05.mma kind::i8, UTCHMMA = kind::f16, UTCBAR =
tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UTMAPF = its L2 prefetch, UBLKCP = cp.async.bulk,
UTCATOMSWS = TMEM alloc / dealloc.
"""
import collections
import re
import subprocess
import sys

KEEP = re.compile(r"^(UTC|LDTM|STTM|UTMA|UBLKCP|SYNCS|ELECT|FENCE|F2FP|I2FP|HMMA|IMMA|STS\.128|LDS\.128|STG\.E\.(64|128)|LDG\.E\.(64|128)|"
                  r"CCTL|NANOSLEEP|BAR|WARPSYNC|FFMA2|HFMA2|PRMT|REDUX|SHFL)")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
    return out.splitlines()


def main():
    for obj in sys.argv[1:]:
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True, check=True).stdout
        usage = {}
        fn = None
        for line in res.splitlines():
            m = re.match(r"\s*Function (\S+):", line)
            if m:
                fn = m.group(1)
            elif fn and "REG:" in line:
                usage[fn] = line.strip()
                fn = None
        kernels = collections.OrderedDict()
        cur = None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            if cur is None:
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                cur["__total__"] += 1
                op = m.group(1)
                if KEEP.match(op):
                    cur[op] += 1
        names = list(kernels)
        pretty = dict(zip(names, demangle(names)))
        print(f"# cuobjdump -sass / -res-usage {obj} (sm_100a), tools/sass_summary.py")
        for k, c in kernels.items():
            short = re.sub(r"\(fanlin::.*$", "", pretty[k]).replace("void fanlin::(anonymous namespace)::", "").replace("fanlin::(anonymous namespace)::", "")
            print(f"\n## {short}   [{c['__total__']} SASS instructions]")
            if k in usage:
                print(f"#  {usage[k]}")
            for op, n in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
                if op != "__total__":
                    print(f"{n:7d} {op}")


if __name__ == "__main__":
    main()
