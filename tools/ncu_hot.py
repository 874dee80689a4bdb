"""Top source lines of a kernel from an ncu report: python tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, fname, data = None, "", []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-" and r[0].isdigit():
            d = dict(zip(hdr, r))
            d["file"] = fname
            data.append(d)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = sum(int(d["# Samples"] or 0) for d in data)
    tot_i = sum(int(d["Instructions Executed"] or 0) for d in data)
    print("total instructions %d, samples %d" % (tot_i, tot_s))
    agg = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
    print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot_s, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for d in sorted(data, key=lambda x: -int(x["# Samples"] or 0))[:top]:
        s = int(d["# Samples"] or 0)
        i = int(d["Instructions Executed"] or 0)
        ms = max(stalls, key=lambda k: int(d[k] or 0))
        print("%5.1f%% smp %5.1f%% ins  %-10s %s:%s  %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, ms[6:], d["file"][:18], d["Line No"],
                                                       d["Source"].strip()[:100]))


if __name__ == "__main__":
    main()
