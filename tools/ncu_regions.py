#!/usr/bin/env python
"""Group the SASS page of an .ncu-rep into code regions (runs of instructions with equal execution
counts): executed warp instructions, share, PC samples, average active lanes, first instruction.
usage: python tools/ncu_regions.py rep.ncu-rep [min_share_pct] [warps]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isrc, iex, ith, ismp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed",
                                                         "Thread Instructions Executed", "# Samples"))
    ins = [(r[isrc].strip(), int(r[iex]), int(r[ith]), int(r[ismp])) for r in rows[2:] if len(r) > ismp and r[iex].isdigit()]
    total = sum(i[1] for i in ins)
    samples = sum(i[3] for i in ins) or 1
    print(f"# {rep}: {total} warp instructions, {samples} PC samples")
    print("sass lines      n   exec/instr    share  samples  lanes  first instruction")
    k = 0
    while k < len(ins):
        j = k
        while j + 1 < len(ins) and abs(ins[j + 1][1] - ins[k][1]) <= 0.02 * max(ins[k][1], 1):
            j += 1
        ex = sum(i[1] for i in ins[k:j + 1])
        th = sum(i[2] for i in ins[k:j + 1])
        sm = sum(i[3] for i in ins[k:j + 1])
        if 100.0 * ex / total >= min_share:
            print(f"{k:4d}-{j:<6d} {j - k + 1:5d} {ins[k][1]:12d} {100.0 * ex / total:7.1f}% {100.0 * sm / samples:7.1f}% "
                  f"{th / max(ex, 1):6.1f}  {ins[k][0][:60]}")
        k = j + 1


if __name__ == "__main__":
    main()
