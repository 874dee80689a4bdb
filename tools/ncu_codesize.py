"""Static SASS bytes and executed share per source line of one file, from an ncu report (--import-source on):
python tools/ncu_codesize.py report.ncu-rep file.cu [min_bytes]
For every source line: SASS instructions attributed to it (all inlined copies), how many were ever executed, and
its share of the executed warp instructions.  Tells what has to shrink for the hot path to fit the instruction cache."""
import csv
import subprocess
import sys


def main():
    rep, path = sys.argv[1], sys.argv[2]
    min_b = int(sys.argv[3]) if len(sys.argv) > 3 else 96
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    base = path.split("/")[-1]
    hdr, fname, key = None, "", None
    per = {}
    tot = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            key = (fname, int(r[0]))
        elif hdr and len(r) == len(hdr) and r[0] == "" and r[2].startswith("0x"):
            i = int(r[hdr.index("Instructions Executed")] or 0)
            a = per.setdefault(key, [0, 0, 0])
            a[0] += 1
            a[1] += 1 if i else 0
            a[2] += i
            tot += i
    src = open(path).read().split("\n")
    print("total executed %d; static instructions %d" % (tot, sum(a[0] for a in per.values())))
    for (f, ln), (n, live, ex) in sorted(per.items()):
        if n * 16 >= min_b or 100.0 * ex / tot >= 0.5:
            text = src[ln - 1].strip()[:90] if f == base and ln <= len(src) else f
            print("%5d %5d B (%4d live) %5.2f%% | %s" % (ln, n * 16, live * 16, 100.0 * ex / tot, text))


if __name__ == "__main__":
    main()
