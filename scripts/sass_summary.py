"""cuobjdump -sass of the built library -> per-kernel counts of the Blackwell-specific instructions (tcgen05 MMA / TMEM
loads / TMA loads and stores / cluster barriers) plus one sample line each: profiles/sass_r02.md.  Runs without a GPU.
    python scripts/sass_summary.py > profiles/sass_r02.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yolact_b200", "libyolact_b200.so")
PAT = collections.OrderedDict([
    ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"),
    ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF/CCTL", r"\bUTMAPF|\bUTMACCTL"),
    ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster)", r"\bUCGABAR"), ("HFMA2", r"\bHFMA2"), ("SHFL", r"\bSHFL"),
    ("LDG.E.128", r"\bLDG\.E\.128|\bLDG\.E\.(?:CONSTANT\.)?128|LDG\.E\.128\.CONSTANT"), ("POPC", r"\bPOPC"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (.*)$", line)
        if m:
            cur = m.group(1).strip()
            kernels[cur] = []
            continue
        if cur and "/*" in line:
            kernels[cur].append(line)
    dem = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, dem)) if len(dem) == len(kernels) else {k: k for k in kernels}
    print("# SASS evidence (cuobjdump -sass yolact_b200/libyolact_b200.so, sm_100a)\n")
    print("Counts of the Blackwell-specific instructions per kernel (instantiations of one template are summed).\n")
    agg = collections.OrderedDict()
    sample = {}
    for k, lines in kernels.items():
        short = names[k].replace("void ", "").replace("yb::<unnamed>::", "").replace("<unnamed>::", "").replace("yb::", "")
        base = re.sub(r"[<(].*", "", short)
        a = agg.setdefault(base, collections.Counter())
        a["instantiations"] += 1
        a["instructions"] += len(lines)
        for key, pat in PAT.items():
            hits = [l for l in lines if re.search(pat, l)]
            a[key] += len(hits)
            if hits and (base, key) not in sample:
                sample[(base, key)] = re.sub(r"\s+", " ", re.sub(r"^\s*/\*[0-9a-f]+\*/", "", hits[0]).split("/*")[0]).strip()
    cols = ["instantiations", "instructions"] + list(PAT)
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for base, a in agg.items():
        print("| `%s` | " % base + " | ".join(str(a[c]) for c in cols) + " |")
    print("\n## one sample line per (kernel, instruction)\n")
    for (base, key), l in sample.items():
        if key in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"):
            print("* `%s` / %s: `%s`" % (base, key, l))


if __name__ == "__main__":
    main()
