"""Dynamic code working set of a kernel from an ncu report (needs --import-source on):
python tools/ncu_workset.py report.ncu-rep
Prints how many 128-byte instruction lines cover 50/80/90/95/99/99.9 % of the executed warp instructions,
and the address layout of the hot lines (so that one can see what has to fit the instruction cache)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    ins = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0].startswith("0x"):
            d = dict(zip(hdr, r))
            ins.append((int(r[0], 16), int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), d["Source"].strip()))
    base = ins[0][0]
    total = sum(i[1] for i in ins)
    print("static: %d instructions = %.1f KB; executed warp instructions %d; never executed: %d instructions" %
          (len(ins), len(ins) * 16 / 1024, total, sum(1 for i in ins if i[1] == 0)))
    lines = {}
    for a, c, s, _ in ins:
        k = (a - base) // 128
        lines[k] = lines.get(k, 0) + c
    order = sorted(lines.items(), key=lambda x: -x[1])
    acc, j = 0, 0
    marks = [0.5, 0.8, 0.9, 0.95, 0.99, 0.999]
    for n, (k, c) in enumerate(order, 1):
        acc += c
        while j < len(marks) and acc >= marks[j] * total:
            print("  %5.1f %% of executed instructions: %4d lines = %5.1f KB" % (100 * marks[j], n, n * 128 / 1024))
            j += 1
    if len(sys.argv) > 2:  # hot lines in address order: KB offset, share
        hot = sorted(k for k, c in order if c >= float(sys.argv[2]) * total / 100)
        print("lines with >= %s %% each: %d" % (sys.argv[2], len(hot)))
        for k in hot:
            print("   +%5.1f KB  %.2f %%" % (k * 128 / 1024, 100.0 * lines[k] / total))


if __name__ == "__main__":
    main()
