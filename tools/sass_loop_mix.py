#!/usr/bin/env python
"""Static instruction mix of a kernel's loops, from the SASS of the built library (cuobjdump; no GPU needed).
For every backward branch the body [target, branch] is listed with its instruction count and its FP64 arithmetic count; the
requested loop (default: the one with the most FP64 instructions per instruction, i.e. the hot loop) gets an opcode histogram.

  python tools/sass_loop_mix.py [--kernel 'k_rollout_lin_grid<false, false>'] [--so path]

DESIGN section 3 quotes the headline kernel's warp-uniform loop from this: 124 static instructions of which 69 FP64
(34 DMUL + 30 DADD + 5 DFMA); ncu counts 120 executed per warp-step.
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DMUL", "DADD", "DFMA")


def kernels(so):
    out = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = {"reg": int(m.group(2)), "stack": int(m.group(3)), "shared": int(m.group(4)), "local": int(m.group(5))}
    names = list(res)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return {re.sub(r"\(.*", "", d).replace("void cck::", ""): (n, res[n]) for n, d in zip(names, dem)}


def sass(so, mangled):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, so], capture_output=True, text=True, check=True).stdout
    ins = []
    for ln in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]


def loops(ins):
    index = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in index:
            body = ins[index[int(m.group(1), 16)]:i + 1]
            ops = [opcode(x[1]) for x in body]
            out.append({"begin": int(m.group(1), 16), "end": a, "n": len(body), "fp64": sum(o in FP64 for o in ops), "hist": collections.Counter(ops)})
    return out


def hot_loop(ls):
    """The loop with the highest FP64 density among those with at least 30 FP64 instructions."""
    c = [l for l in ls if l["fp64"] >= 30]
    return max(c, key=lambda l: l["fp64"] / l["n"]) if c else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--so", default=os.path.join(ROOT, "chance-constrained-k-cbs_b200", "libcck_b200.so"))
    ap.add_argument("--kernel", default="k_rollout_lin_grid<false, false>")
    a = ap.parse_args()
    ks = kernels(a.so)
    mangled, res = ks[a.kernel]
    ins = sass(a.so, mangled)
    ls = loops(ins)
    print(f"{a.kernel}: {len(ins)} SASS instructions, {res['reg']} registers, {res['local']} B local, {res['stack']} B stack")
    print("loops (backward branches): begin end instructions fp64")
    for l in sorted(ls, key=lambda l: l["n"]):
        print(f"  {l['begin']:#07x} {l['end']:#07x} {l['n']:5d} {l['fp64']:4d}")
    h = hot_loop(ls)
    if h:
        print(f"hot loop {h['begin']:#x}-{h['end']:#x}: {h['n']} instructions, {h['fp64']} FP64 arithmetic "
              f"({', '.join(f'{h['hist'][o]} {o}' for o in FP64)}); issue-cycle model 2*fp64/(2*fp64+other) = "
              f"{2 * h['fp64'] / (2 * h['fp64'] + h['n'] - h['fp64']):.3f}")
        for op, n in h["hist"].most_common():
            print(f"    {op:10s} {n}")
    whole = collections.Counter(opcode(t) for _, t in ins)
    print("whole kernel:", ", ".join(f"{whole[o]} {o}" for o in ("DMUL", "DADD", "DFMA", "DSETP", "UBLKCP", "LDGSTS", "SYNCS", "ATOMS", "STS", "LDS", "LDG", "STG") if whole[o]))


if __name__ == "__main__":
    main()
