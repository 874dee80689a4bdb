"""SASS of the instructions a range of source lines compiled to, in one kernel:
python tools/sass_excerpt.py file.cubin kernel_substr src_basename first_line last_line [max_instr]
(nvdisasm -g prints a '//## File "...", line N' marker before each group of instructions; needs -lineinfo)."""
import re
import subprocess
import sys


def main():
    cubin, kern, base, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    cap = int(sys.argv[6]) if len(sys.argv) > 6 else 80
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    in_k, on, n, last = False, False, 0, None
    for ln in out:
        if ln.startswith(".text."):
            in_k = kern in ln
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            on = m.group(1).endswith(base) and lo <= int(m.group(2)) <= hi
            if in_k and on and last != m.group(2):
                print("    // %s:%s" % (base, m.group(2)))
                last = m.group(2)
            continue
        if in_k and on and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln):
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.rstrip()))
            n += 1
            if n >= cap:
                print("    ... (cut at %d instructions)" % cap)
                return


if __name__ == "__main__":
    main()
