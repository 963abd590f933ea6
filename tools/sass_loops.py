#!/usr/bin/env python
"""Static view of a kernel's loops: dump the SASS of one function from an object file and list every
backward branch with the number of instructions (and the opcode histogram) of the loop it closes.

    python tools/sass_loops.py mantaray_b200/csrc/build/mr_kernels_fast.o 'trace_kernelILi2ELi1ELi0ELb1ELi1E' [--dump out.sass]
"""
import collections, re, subprocess, sys


def function_sass(obj, pattern):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = pattern in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    t = ins.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


if __name__ == "__main__":
    obj, pat = sys.argv[1], sys.argv[2]
    ins = function_sass(obj, pat)
    if "--dump" in sys.argv:
        with open(sys.argv[sys.argv.index("--dump") + 1], "w") as f:
            for a, s in ins:
                f.write(f"/*{a:04x}*/ {s}\n")
    print("static instructions:", len(ins))
    for a, s in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:U?P\d,\s*|!U?P\d,\s*)?(?:U?P\d,\s*)?0x([0-9a-f]+)", s)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            body = [x for x in ins if t <= x[0] <= a]
            h = collections.Counter(opcode(x[1]) for x in body)
            print(f"loop {t:#06x}..{a:#06x}: {len(body)} instrs; " + " ".join(f"{k}:{v}" for k, v in h.most_common(14)))
