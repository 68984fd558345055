"""Aggregate the SASS source page of an .ncu-rep: samples / executed instructions per opcode and per contiguous hot region.
python tools/ncu_source_hot.py report.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def load(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], stderr=subprocess.DEVNULL).decode()
    lines = raw.splitlines()
    # first line: kernel name; second: header
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[1:] if len(r) == len(hdr)]


def main(rep):
    rows = load(rep)
    tot_s = sum(int(r["# Samples"]) for r in rows)
    tot_i = sum(int(r["Instructions Executed"]) for r in rows)
    print("instructions: %d SASS lines, %d warp-instructions executed, %d samples" % (len(rows), tot_i, tot_s))
    by_op = defaultdict(lambda: [0, 0])
    for r in rows:
        src = r["Source"].strip()
        src = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = src.split()[0] if src else "?"
        op = ".".join(op.split(".")[:2])
        by_op[op][0] += int(r["Instructions Executed"])
        by_op[op][1] += int(r["# Samples"])
    print("\nper opcode (executed %, samples %):")
    for op, (i, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
        print("  %-18s %6.2f %%  %6.2f %%" % (op, 100.0 * i / tot_i, 100.0 * s / max(tot_s, 1)))
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    print("\nstall reasons over all samples:")
    st = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
    for c, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-24s %6.2f %%" % (c, 100.0 * v / max(tot_s, 1)))
    # regions: split the address space into chunks of 256 instructions
    print("\nregions of 256 SASS lines (executed %, samples %, first BAR/label-ish instruction):")
    for s0 in range(0, len(rows), 256):
        ch = rows[s0:s0 + 256]
        i = sum(int(r["Instructions Executed"]) for r in ch)
        s = sum(int(r["# Samples"]) for r in ch)
        if 100.0 * i / tot_i >= 0.5 or 100.0 * s / max(tot_s, 1) >= 0.5:
            print("  [%5d, %5d)  %6.2f %%  %6.2f %%" % (s0, s0 + len(ch), 100.0 * i / tot_i, 100.0 * s / max(tot_s, 1)))
    print("\ntop 40 instructions by samples:")
    top = sorted(enumerate(rows), key=lambda kv: -int(kv[1]["# Samples"]))[:40]
    for idx, r in top:
        reasons = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
        print("  #%5d %6.2f %%  exec %9s  %-58s %s" % (idx, 100.0 * int(r["# Samples"]) / max(tot_s, 1), r["Instructions Executed"], r["Source"].strip()[:58],
                                                  ", ".join("%s %d" % (c[6:], v) for v, c in reasons)))


if __name__ == "__main__":
    main(sys.argv[1])
