"""Static SASS size of one kernel by source line: python tools/sass_lines.py file.cubin kernel_substr src.cu [min_instr]
(nvdisasm -g prints a //## File "...", line N marker before each group of instructions; needs -lineinfo)."""
import re
import subprocess
import sys


def main():
    cubin, kern, src = sys.argv[1], sys.argv[2], sys.argv[3]
    min_i = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    base = src.split("/")[-1]
    lines = open(src).read().split("\n")
    cur_file, cur_line, in_k = "", 0, False
    per, tot, other = {}, 0, {}
    for ln in out:
        if ln.startswith(".text."):
            in_k = kern in ln
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
            continue
        if in_k and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            tot += 1
            if cur_file == base:
                per[cur_line] = per.get(cur_line, 0) + 1
            else:
                other[cur_file] = other.get(cur_file, 0) + 1
    print("kernel %s: %d instructions = %.1f KB" % (kern, tot, tot * 16 / 1024))
    for f, c in sorted(other.items(), key=lambda x: -x[1]):
        print("  other %-30s %5d" % (f, c))
    for l in sorted(per):
        if per[l] >= min_i:
            print("%5d %5d | %s" % (l, per[l], lines[l - 1].strip()[:110]))


if __name__ == "__main__":
    main()
