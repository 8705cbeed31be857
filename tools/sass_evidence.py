"""profiles/sass_r2.txt: per kernel of libapg_b200.so, how many tcgen05 / TMEM / bulk-copy instructions its SASS holds
(cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld /
st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier, HMMA = legacy mma.sync)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "apg_trajectory_tracking_b200", "libapg_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "HMMA", "MUFU", "ACQBULK", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, fn = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in MNEMONICS:
                counts[fn][op] += 1
    lines = ["SASS instruction counts per kernel of libapg_b200.so (sm_100a), cuobjdump -sass; tools/sass_evidence.py",
             f"{'kernel':58s} " + " ".join(f"{m:>8s}" for m in MNEMONICS)]
    for fn, c in counts.items():
        if sum(c.values()):
            lines.append(f"{fn[:58]:58s} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
    text = "\n".join(lines) + "\n"
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r2.txt")
    open(path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
