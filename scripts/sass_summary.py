"""SASS / resource summary of the built library (runs here, no GPU):  python scripts/sass_summary.py > profiles/<name>.md
Per kernel family: instantiations, registers, spills, shared memory (cuobjdump --dump-resource-usage) and the counts of
the instructions that prove which hardware path a kernel uses (B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, HMMA = mma.sync, MUFU = special-function unit)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vln-magic_b200", "lib", "libmagic_b200.so")
OPS = ("UTCHMMA", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU", "SYNCS", "ACQBULK",
       "LDGSTS", "REDUX", "ATOMG", "RED.", "STL", "LDL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def family(name):
    m = re.match(r"(?:void )?(?:\(anonymous namespace\)::)?(\w+)", name)
    return m.group(1) if m else name


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, n_instr, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None or "/*" not in ln:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        n_instr[cur] += 1
        op = m.group(1)
        for key in OPS:
            if op.startswith(key):
                counts[cur][key] += 1
    names = demangle(sorted(usage))
    fam = collections.defaultdict(list)
    for mangled in usage:
        fam[family(names[mangled])].append(mangled)
    digest = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    stamp = open(os.path.join(os.path.dirname(LIB), "build.stamp")).read().strip()[:16]
    print(f"# SASS / resource summary of `vln-magic_b200/lib/libmagic_b200.so` (sha256 {digest}…, build.stamp {stamp}…)\n")
    print("`python scripts/sass_summary.py` (cuobjdump --dump-resource-usage / -sass, sm_100a). Per kernel family: number of "
          "instantiations, register range, stack bytes (spills), static shared memory, SASS instruction count, and how many "
          "of the marker instructions the family's kernels contain in total.\n")
    print("| kernel family | inst. | regs | max stack (B) | max static smem (B) | SASS instr. | marker instructions |")
    print("|---|---:|---:|---:|---:|---:|---|")
    for f in sorted(fam, key=lambda k: -sum(n_instr[m] for m in fam[k])):
        ms = fam[f]
        regs = [usage[m][0] for m in ms]
        tot = collections.Counter()
        for m in ms:
            tot.update(counts[m])
        marks = ", ".join(f"{k} {v}" for k, v in tot.most_common() if v) or "—"
        r = f"{min(regs)}" if min(regs) == max(regs) else f"{min(regs)}–{max(regs)}"
        print(f"| `{f}` | {len(ms)} | {r} | {max(usage[m][1] for m in ms)} | {max(usage[m][2] for m in ms)} | "
              f"{sum(n_instr[m] for m in ms)} | {marks} |")
    print(f"\n{len(usage)} kernels in {len(fam)} families.")


if __name__ == "__main__":
    sys.exit(main())
