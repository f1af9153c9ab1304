"""Opcode mix of the innermost hot loop of a kernel in a cubin / .so (no GPU needed).

    python tools/sass_loop_mix.py <lib.so> <mangled-kernel-substring> [min_loop_len]

Finds backward branches in the kernel's SASS, takes the loop with the most instructions whose body
contains no other loop's back edge start outside of it, and prints the opcode histogram of its body
(the per-iteration instruction count the issue-bound analysis in DESIGN.md uses)."""
import re
import subprocess
import sys
from collections import Counter


def kernel_sass(lib, key):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if key in name:
            return name, b
    raise SystemExit(f"kernel {key} not found")


def main():
    lib, key = sys.argv[1], sys.argv[2]
    name, body = kernel_sass(lib, key)
    ins = []
    for ln in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    print(f"{name}: {len(ins)} instructions, loops (start, end, n): "
          + ", ".join(f"({a:#x},{b:#x},{sum(1 for x, _ in ins if a <= x <= b)})" for a, b in loops))
    want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    cands = [(sum(1 for x, _ in ins if a <= x <= b), a, b) for a, b in loops]
    cands = [c for c in cands if c[0] >= want]
    n, a, b = min(cands) if want else max(cands)
    mix = Counter()
    for x, text in ins:
        if a <= x <= b:
            t = re.sub(r"^@!?U?P\d+\s+", "", text)
            mix[t.split()[0].split(".")[0]] += 1
    print(f"loop {a:#x}..{b:#x}: {n} instructions")
    for op, c in mix.most_common():
        print(f"  {op:10s} {c}")


if __name__ == "__main__":
    main()
