"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA, bulk copies), legacy HMMA,
plus FFMA2 (packed fp32x2) -- from `cuobjdump -sass` of the built library.

    python tools/sass_summary.py [path/to/libdmf_b200.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "FFMA2", "MUFU", "REDUX", "SYNCS")


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            cur["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    cur[k] += 1
    print(f"# {os.path.basename(path)}: SASS mnemonic counts per kernel (cuobjdump -sass)")
    print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
    tot = collections.Counter()
    for name, c in per.items():
        tot.update(c)
        if any(c[k] for k in KEYS[:9]) or c["FFMA2"]:
            print(f"{name[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    print(f"{'TOTAL (all ' + str(len(per)) + ' kernels)':58s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "disentagled_multimodal_fusion_b200", "libdmf_b200.so"))
