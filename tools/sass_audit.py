#!/usr/bin/env python
"""Count instruction classes in the SASS of one kernel (cuobjdump -sass output), split into the main loop
(between the first backward-branch target and the branch) and the whole function.
usage: sass_audit.py <sass file> <mangled-name substring>"""
import re, sys, collections

def main():
    path, pat = sys.argv[1], sys.argv[2]
    lines = open(path).read().split("\n")
    start = None
    for i, l in enumerate(lines):
        if "Function :" in l and pat in l:
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found")
    end = len(lines)
    for i in range(start + 1, len(lines)):
        if "Function :" in lines[i]:
            end = i
            break
    ins = []  # (addr, opcode, text)
    rx = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")
    for l in lines[start:end]:
        m = rx.search(l)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    # the hot loop
    best = (-1, 0, 0)  # the backward branch whose span holds the most wide multiplies
    for a, op, txt in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", txt)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a:
                    w = sum(1 for x in ins if tgt <= x[0] <= a and x[1].startswith("IMAD.WIDE"))
                    if w > best[0]:
                        best = (w, tgt, a)
    def classify(op):
        if op.startswith("IMAD.WIDE"):
            return "IMAD.WIDE" + (".X" if ".X" in op else "")
        if op.startswith("IMAD"):
            for k in ("MOV", "IADD", "SHL", "HI", "X"):
                if "." + k in op:
                    return "IMAD." + k
            return "IMAD"
        return op.split(".")[0]
    for name, sel in (("function", ins), ("loop [%#x..%#x]" % (best[1], best[2]), [x for x in ins if best[1] <= x[0] <= best[2]])):
        c = collections.Counter(classify(op) for _, op, _ in sel)
        tot = sum(c.values())
        fma = sum(v for k, v in c.items() if k.startswith("IMAD") or k in ("FFMA", "FMUL", "FADD", "HFMA2"))
        print(f"== {name}: {tot} instructions, fma-pipe {fma}")
        for k, v in c.most_common(30):
            print(f"   {k:16s} {v}")

main()
