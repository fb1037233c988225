"""Static SASS histogram of one kernel in an object file / shared library (no GPU needed):
    python tools/sass_hist.py <file.o|.so> <substring of the kernel name> [--loop]
Prints the instruction count by opcode for the whole function and, with --loop, for the largest backward-branch loop
(the z march of the fused 3-D kernels) -- the number DESIGN.md quotes as "instructions per voxel column and slice"."""
import collections
import re
import subprocess
import sys


def main():
    path, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs = out.split("Function : ")
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print(name, "total", len(ins))
        # largest loop: backward branch with the largest span
        best = None
        for addr, txt in ins:
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", txt)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
        sel = ins
        if "--loop" in sys.argv and best:
            sel = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
            print("loop", hex(best[0]), hex(best[1]), "instructions", len(sel))
        h = collections.Counter()
        for _, t in sel:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            h[t.split()[0].split(".")[0]] += 1
        print(" ".join(f"{k}:{v}" for k, v in h.most_common(30)))
        reuse = sum(1 for _, t in sel if ".reuse" in t and t.lstrip("@!UP0123456789 ").startswith("FFMA2"))
        print("FFMA2 with .reuse:", reuse)


main()
