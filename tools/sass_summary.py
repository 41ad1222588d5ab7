"""Per-kernel counts of the SASS mnemonics that prove the Blackwell machinery (tcgen05 MMA, tensor
memory loads / stores, TMA loads / stores) in the built library:

    python tools/sass_summary.py [> profiles/rN_sass_summary.txt]

Runs `cuobjdump -sass` on libovdet.so (build it first: python -m ovdet.build) and prints one row per
kernel that contains at least one of the mnemonics, plus totals.  Mnemonics
(/opt/skills/guides/B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit,
LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
UTMAPF = TMA prefetch, SYNCS = mbarrier operations, UTCATOMSWS = tensor-memory allocation."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "real-time-zero-shot-open-vocabulary-object-detection-using-a-lightweight_b200", "libovdet.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UTCATOMSWS"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    current = None
    arch = set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            counts[current] = collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        if current is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[current]["instructions"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[current][mn] += 1
    names = demangle(list(counts))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (arch: {', '.join(sorted(arch))}; {len(counts)} kernels)")
    print("# " + " ".join(f"{m:>10}" for m in MNEMONICS) + "  instructions  kernel")
    total = collections.Counter()
    for fn, c in counts.items():
        total.update(c)
        if not any(c[m] for m in MNEMONICS):
            continue
        short = names[fn].replace("(anonymous namespace)::", "").replace("ovdet::", "")
        short = re.sub(r"^void ", "", short)
        cut = short.find(">(")
        short = short[:cut + 1] if cut >= 0 else re.sub(r"\(.*", "", short)
        print("  " + " ".join(f"{c[m]:>10}" for m in MNEMONICS) + f"  {c['instructions']:>12}  {short}")
    print("# " + " ".join(f"{total[m]:>10}" for m in MNEMONICS) + f"  {total['instructions']:>12}  TOTAL (all kernels)")


if __name__ == "__main__":
    sys.exit(main())
