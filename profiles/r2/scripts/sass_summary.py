"""Opcode census of the kernels in a built library (cuobjdump -sass): which memory / logic / tensor-memory / bulk-copy
instructions each kernel really contains.  usage: sass_summary.py lib.so [lib2.so ...] > profiles/r2/sass_opcodes.md"""
import collections, re, subprocess, sys

WATCH = ["LOP3", "LDS", "STS", "LDSM", "LDG", "STG", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS",
         "BRX", "BRA", "BAR", "POPC", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "MATCH", "LDL", "STL", "LDC", "ISETP", "IADD3", "IMAD", "PRMT", "MOV"]


def census(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
            cur["_total"] += 1
    return kernels


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return name


for lib in sys.argv[1:]:
    ks = census(lib)
    print("## %s\n" % lib)
    cols = [w for w in WATCH if any(k[w] for k in ks.values())]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for name, c in ks.items():
        print("| `%s` | %d | " % (demangle(name), c["_total"]) + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    print()
