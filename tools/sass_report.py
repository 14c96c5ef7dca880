#!/usr/bin/env python
"""Blackwell evidence from the built library, no GPU needed: per kernel, the SASS instruction counts that show what it is made
of (TMA tensor / bulk copies, mbarrier operations, packed f32x2 arithmetic, SFU calls), and registers / spills / shared memory
from the ptxas logs of the same build.

    python tools/sass_report.py [regex ...] > profiles/sass_rNN.md          (after `make -C ideal-gan_b200/csrc`; the regexes select the
                                                                             kernels of the main table, default all)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ideal-gan_b200", "idealgan", "libidealgan.so")
BUILD = os.path.join(ROOT, "ideal-gan_b200", "csrc", "build")

PATTERNS = collections.OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"), ("FMUL2", r"\bFMUL2"),
    ("FADD2", r"\bFADD2"), ("FFMA", r"\bFFMA\b"), ("MUFU", r"\bMUFU"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("LDG", r"\bLDG"),
    ("STG", r"\bSTG"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED)\b"), ("SHFL", r"\bSHFL"), ("LDL/STL", r"\b(LDL|STL)\b")])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def sass_counts():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur and "/*" in line and ";" in line:
            kernels[cur]["instructions"] += 1
            for k, pat in PATTERNS.items():
                if re.search(pat, line):
                    kernels[cur][k] += 1
    return arch, kernels


def ptxas_info():
    info = {}
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".ptxas.log"):
            continue
        cur = None
        for line in open(os.path.join(BUILD, f)):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1)
                info[cur] = {"file": f.replace(".ptxas.log", ".cu")}
                continue
            if cur:
                m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
                if m:
                    info[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
                m = re.search(r"Used (\d+) registers", line)
                if m:
                    info[cur]["regs"] = int(m.group(1))
                    sm = re.search(r"(\d+) bytes smem", line)
                    info[cur]["smem"] = int(sm.group(1)) if sm else 0
    return info


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ig::", "").replace("(anonymous namespace)::", "")


def main():
    arch, kernels = sass_counts()
    info = ptxas_info()
    dm = demangle(list(kernels))
    print("# SASS / ptxas evidence of the built library\n")
    print(f"`cuobjdump -sass ideal-gan_b200/idealgan/libidealgan.so`: arch = {', '.join(arch)}; {len(kernels)} kernels, "
          f"{sum(k['instructions'] for k in kernels.values())} instructions.  Produced by `tools/sass_report.py` from the binaries of this commit.\n")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("Library totals: " + ", ".join(f"{k} {tot[k]}" for k in PATTERNS) + "\n")
    print("No `UTC*MMA` / TMEM instruction is present and none is expected: the largest contraction on this path is 2 x ne per voxel, in fp32.\n")
    focus = [a for a in sys.argv[1:] if not a.startswith("-")]
    cols = ["regs", "spill st/ld B", "smem B"] + list(PATTERNS)
    if focus:
        print("Main table: kernels matching " + ", ".join(f"`{f}`" for f in focus) + " (the instantiations the BASELINE configurations launch at ne = 6, "
              "the ne = 12 ones, and the kernels without an echo-count template).\n")
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    rows = []
    for mangled, c in kernels.items():
        name = short(dm.get(mangled, mangled))
        if focus and not any(re.search(f, name) for f in focus):
            continue
        i = info.get(mangled, {})
        rows.append((name, i, c))
    rows.sort(key=lambda r: r[0])
    for name, i, c in rows:
        cells = [str(i.get("regs", "?")), f"{i.get('spill_st', '?')}/{i.get('spill_ld', '?')}", str(i.get("smem", "?"))] + [str(c[k]) if c[k] else "" for k in PATTERNS]
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    spilling = [(short(dm.get(m, m)), i) for m, i in info.items() if i.get("spill_st", 0) or i.get("spill_ld", 0)]
    print(f"\n## Instantiations that spill ({len(spilling)} of {len(info)})\n")
    for name, i in sorted(spilling, key=lambda x: -x[1].get("spill_st", 0)):
        print(f"* `{name}` ({i['file']}): {i.get('regs')} registers, {i.get('spill_st')} B stores / {i.get('spill_ld')} B loads")


if __name__ == "__main__":
    main()
