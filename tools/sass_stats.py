"""Developer probe: registers / spills (ptxas -v log) and SASS opcode mix of the Jacobi kernels of a cubin.
usage: sass_stats.py <ptxas_log> <cubin> [T,MODE,VAR ...]"""
import collections
import re
import subprocess
import sys


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    log, cubin = sys.argv[1], sys.argv[2]
    want = [tuple(a.split(",")) for a in sys.argv[3:]] or [("7", "0", "0"), ("7", "1", "0"), ("6", "0", "0"), ("6", "1", "0"), ("7", "0", "3"), ("7", "0", "2")]
    txt = open(log).read()
    for name, stack, ss, sl, regs in re.findall(
            r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
        m = re.search(r"jacobi_stream_kernel<(\d+), (\d+), (\d+)>", demangle(name))
        if m and m.groups() in want:
            print(m.groups(), "regs", regs, "stack", stack, "spill", ss, sl)
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        m = re.search(r"jacobi_stream_kernel<(\d+), (\d+), (\d+)>", demangle(f.split("\n")[0].strip()))
        if not m or m.groups() not in want:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in ins)
        print(m.groups(), "instructions", len(ins), c.most_common(22))


if __name__ == "__main__":
    main()
