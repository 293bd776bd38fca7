#!/usr/bin/env python3
"""Where a kernel's warps wait: the warp-stall samples of an `ncu --set full --import-source on` report, summed per stall
reason over the whole kernel and listed for the SASS instructions that collected the most samples.  Reads the report
here (no GPU needed):

    python tools/ncu_stalls.py gpurun_out/ncu_tc3_r02b.ncu-rep > profiles/r02/ncu_tc3_v2_stalls.txt
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernel = rows[0][1] if rows and rows[0][0] == "Kernel Name" else "?"
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    reasons = [n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    total = {n: 0 for n in reasons}
    per = []
    executed = 0
    for r in body:
        n_s = int(r[col["# Samples"]] or 0)
        executed += int(r[col["Instructions Executed"]] or 0)
        st = {n: int(r[col[n]] or 0) for n in reasons}
        for n in reasons:
            total[n] += st[n]
        per.append((n_s, r[col["Source"]].strip(), st, int(r[col["Instructions Executed"]] or 0)))
    all_s = sum(p[0] for p in per)
    print(f"# {rep}: warp-stall samples (ncu --page source), tools/ncu_stalls.py")
    print(f"# kernel: {kernel}")
    print(f"# {len(body)} SASS instructions, {executed} warp instructions executed, {all_s} samples\n")
    print("## samples per stall reason (whole kernel)")
    for n, v in sorted(total.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"{v:9d}  {100.0 * v / max(all_s, 1):5.1f} %  {n}")
    print(f"\n## the {top_n} instructions with the most samples (share of all samples; executed count; the instruction's main stall reasons)")
    for n_s, src, st, ex in sorted(per, key=lambda p: -p[0])[:top_n]:
        main_r = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"{n_s:8d}  {100.0 * n_s / max(all_s, 1):5.1f} %  x{ex:<10d} {src[:70]:70s} {main_r}")
    # instruction mix by opcode (executed warp instructions)
    mix = {}
    for n_s, src, st, ex in per:
        op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else (src.split()[0] if src else "?")
        mix[op] = mix.get(op, 0) + ex
    print("\n## executed warp instructions by opcode (top 20)")
    for op, v in sorted(mix.items(), key=lambda kv: -kv[1])[:20]:
        print(f"{v:12d}  {100.0 * v / max(executed, 1):5.1f} %  {op}")


if __name__ == "__main__":
    main()
