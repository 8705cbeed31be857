"""Map ncu SASS-level stall samples of a kernel to CUDA source lines.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled function name substring>"""
import csv, io, re, subprocess, sys, collections

rep, kre, cubin, fn = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the export holds one section per matching launch: "Kernel Name",<name> / header / sass rows; take the first
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and re.search(kre, r[1]))
hi = next(i for i in range(start, len(rows)) if "Source" in rows[i] and "Address" in rows[i])
hdr = rows[hi]
si, ss = hdr.index("# Samples"), hdr.index("Source")
sass = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr) and r[si].isdigit():
        sass.append((r[ss].strip(), int(r[si])))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line, file_, instrs = None, None, []
insec = False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        insec = fn in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        file_, line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        instrs.append((file_, line, m.group(2).strip()))
print(f"ncu sass rows {len(sass)}, nvdisasm instrs {len(instrs)}")
n = min(len(sass), len(instrs))
agg = collections.Counter()
tot = sum(s for _, s in sass)
for (txt, smp), (f, l, itxt) in zip(sass[:n], instrs[:n]):
    agg[(f, l)] += smp
src_cache = {}
def src(f, l):
    import glob
    if f not in src_cache:
        c = glob.glob(f"apg_trajectory_tracking_b200/csrc/{f}")
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    s = src_cache[f]
    return s[l - 1].strip()[:100] if 0 < l <= len(s) else ""
print("total samples", tot)
for (f, l), s in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    print(f"{s:6d} {100 * s / tot:5.1f}%  {f}:{l}  {src(f, l)}")
