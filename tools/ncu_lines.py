"""Per-source-line samples/instructions of one file from an ncu report, with the source text:
python tools/ncu_lines.py report.ncu-rep path/to/file.cu [min_pct]"""
import csv
import subprocess
import sys


def main():
    rep, path = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    base = path.split("/")[-1]
    hdr, fname, data, other = None, "", {}, {}
    tot_s = tot_i = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            s, i = int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0)
            tot_s += s
            tot_i += i
            tgt = data if fname == base else other
            key = int(r[0]) if fname == base else fname
            a = tgt.setdefault(key, [0, 0, {}])
            a[0] += s
            a[1] += i
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    a[2][h[6:]] = a[2].get(h[6:], 0) + int(d[h] or 0)
    src = open(path).read().split("\n")
    print("total samples %d instructions %d" % (tot_s, tot_i))
    for k, (s, i, st) in sorted(other.items(), key=lambda x: -x[1][0]):
        print("other %-28s %5.1f%% smp %5.1f%% ins" % (k, 100.0 * s / tot_s, 100.0 * i / tot_i))
    for ln in sorted(data):
        s, i, st = data[ln]
        if 100.0 * s / tot_s >= min_pct or 100.0 * i / tot_i >= min_pct:
            top = max(st, key=st.get) if st else ""
            print("%5d %5.2f%% smp %5.2f%% ins %-9s| %s" % (ln, 100.0 * s / tot_s, 100.0 * i / tot_i, top, src[ln - 1].strip()[:100]))


if __name__ == "__main__":
    main()
