"""Static census of the shipped library (no GPU needed): per kernel the registers, shared memory
and spill figures `cuobjdump -res-usage` reports, and counts of the SASS mnemonics that show which
hardware paths a kernel uses (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM =
tcgen05.ld / st, SYNCS = mbarrier, MUFU = special-function unit, DADD/DFMA/DMUL = fp64).

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import collections
import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parents[1]
LIB = ROOT / "mininf_b200" / "_lib" / "libmininf_b200.so"
WATCH = ["UTCHMMA", "UTMALDG", "UTMAPF", "UTMASTG", "LDTM", "STTM", "SYNCS", "MUFU", "DADD", "DFMA", "DMUL",
         "LDG", "STG", "ATOM", "RED", "FFMA", "HFMA2", "F2FP"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*$", "", name)               # drop the argument list
    return name.replace("void ", "")


def main():
    usage = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True).stdout
    rows = {}
    current = None
    for line in usage.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            current = m.group(1)
            continue
        if current and "REG:" in line:
            fields = dict(re.findall(r"(\w+):(\d+)", line))
            rows[current] = fields
            current = None
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    lengths = collections.Counter()
    current = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            current = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and current:
            lengths[current] += 1
            op = m.group(1)
            for key in WATCH:
                if op == key or op.startswith(key + "."):
                    counts[current][key] += 1
    names = demangle(sorted(rows))
    print("# Static census of `mininf_b200/_lib/libmininf_b200.so` (sm_100a)\n")
    print("Produced by `python tools/sass_census.py` from `cuobjdump -res-usage` and `cuobjdump -sass` of the library "
          "the tests and the bench load; no GPU involved. SHARED is static shared memory (dynamic shared memory of "
          "the tcgen05 kernels is requested at launch), STACK > 0 would mean spills to local memory.\n")
    print("| kernel | REG | SHARED | STACK | SASS instr | " + " | ".join(WATCH) + " |")
    print("|---|---|---|---|---|" + "---|" * len(WATCH))
    for mangled in sorted(rows, key=lambda k: names[k]):
        f = rows[mangled]
        c = counts[mangled]
        print(f"| `{short(names[mangled])}` | {f.get('REG', '?')} | {f.get('SHARED', '?')} | {f.get('STACK', '?')} | "
              f"{lengths[mangled]} | " + " | ".join(str(c[k]) if c[k] else "" for k in WATCH) + " |")
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("\nLibrary totals: " + ", ".join(f"{k} {total[k]}" for k in WATCH if total[k]) + ".")


if __name__ == "__main__":
    sys.exit(main())
