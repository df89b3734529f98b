"""Per-kernel count of the Blackwell-only SASS instructions in the shipped library:
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UTMASTG / UTMAREDG (TMA tensor store / reduce-add), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier).
    python tools/sass_summary.py [libvnet_b200.so] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

MNEMONICS = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS")


def summarise(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in MNEMONICS:
                per[cur][op] += 1
    return per


def demangle(names):
    try:
        r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, r))
    except Exception:
        return {n: n for n in names}


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "vnet_tensorflow_b200", "libvnet_b200.so")
    per = summarise(lib)
    names = demangle(list(per))
    total = collections.Counter()
    print("# cuobjdump -sass %s: Blackwell-specific instruction counts per kernel" % os.path.basename(lib))
    print("# %-70s %s" % ("kernel", " ".join("%8s" % m for m in MNEMONICS)))
    for k, c in per.items():
        total.update(c)
        if sum(c.values()):
            short = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("vnb::", "")
            print("%-72s %s" % (short[:72], " ".join("%8d" % c[m] for m in MNEMONICS)))
    print("%-72s %s" % ("TOTAL", " ".join("%8d" % total[m] for m in MNEMONICS)))
