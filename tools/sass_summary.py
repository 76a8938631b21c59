"""Write profiles/r02c_sass_summary.txt (or the path given as the first argument): per-kernel counts of the SASS mnemonics that prove what the built library runs on
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor loads / stores, LDTM = tcgen05.ld from TMEM, FFMA2 = packed fma.rn.f32x2,
SYNCS / UTCBAR = mbarrier / tcgen05.commit).  Run where the library was built:   python tools/sass_summary.py
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "usot_b200", "libusot_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "HMMA", "LDGSTS", "ATOMG", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
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
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    cur[mn] += 1
                    break
            cur["_total"] += 1
    arch = re.findall(r"arch = (sm_\w+)", out)
    lines = [f"SASS summary of usot_b200/libusot_b200.so (cuobjdump -sass; architectures: {sorted(set(arch))})",
             "kernel".ljust(58) + "".join(m.rjust(9) for m in MNEMONICS) + "   total"]
    tot = collections.Counter()
    for k, c in per.items():
        lines.append(k[:57].ljust(58) + "".join(str(c.get(m, 0)).rjust(9) for m in MNEMONICS) + str(c["_total"]).rjust(8))
        tot.update(c)
    lines.append("ALL KERNELS".ljust(58) + "".join(str(tot.get(m, 0)).rjust(9) for m in MNEMONICS) + str(tot["_total"]).rjust(8))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02c_sass_summary.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:3] + lines[-1:]))
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
