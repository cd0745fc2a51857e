#!/usr/bin/env python
"""Lists the loops (backward branches) of one kernel in an object file with their instruction counts and the
opcode mix -- a quick check of instructions per frame before spending GPU time.

    python tools/sass_loops.py hubertfa_b200/csrc/build/hfa_dp.o warp_any_kernelILb0 [min_instrs]
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, ins = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and pat in fn:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    print(f"{len(ins)} instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\S+)?\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx and i - addr_idx[tgt] + 1 >= min_n:
                body = ins[addr_idx[tgt]:i + 1]
                c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
                print(f"loop {tgt:#x}..{a:#x}: {len(body)} instr  " + " ".join(f"{k}:{v}" for k, v in c.most_common(14)))


if __name__ == "__main__":
    main()
