"""Counts the Blackwell-specific SASS mnemonics per kernel of libft3d.so (`cuobjdump -sass`): UTC*MMA = tcgen05.mma,
LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA unit), LDGSTS = cp.async gathers, REDG = red.global.add."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "fusiontransformer_b200" / "libft3d.so"
MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "REDG", "SYNCS", "HMMA")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + mn + r"(\.|\b)", line):
                cnt[cur][mn] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print("%-46s " % "kernel" + " ".join("%8s" % m for m in MNEMONICS))
    for mangled, name in sorted(zip(cnt, names), key=lambda kv: kv[1]):
        c = cnt[mangled]
        if not any(c[m] for m in ("UTCHMMA", "UTCQMMA", "LDTM", "UBLKCP", "LDGSTS", "REDG")):
            continue
        short = re.sub(r"\(.*", "", name).replace("ft3d::", "").replace("void ", "")
        print("%-46s " % short[:46] + " ".join("%8d" % c[m] for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
