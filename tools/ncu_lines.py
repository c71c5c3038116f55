"""Warp-stall samples per SOURCE line of one kernel: joins `ncu --page source --csv` (SASS + samples) with the line table of
`nvdisasm -g` on the cubin of the same build.  usage: ncu_lines.py report.ncu-rep libkdsl.so KERNEL_MANGLED_PREFIX [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, so, kprefix = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + kprefix))
off2line, cur = {}, None
for ln in sass[start + 1:]:
    if ln.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][0], 16)
agg, why, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    key = off2line.get(int(r[0], 16) - base, ("?", 0))
    s = int(r[idx["# Samples"]] or 0)
    agg[key] += s
    tot += s
    for h in stall:
        v = int(r[idx[h]] or 0)
        if v:
            why[key][h[6:]] += v
print("# %s: %d warp-state samples" % (rows[0][1], tot))
for (f, l), v in agg.most_common(top):
    try:
        text = open(f).read().split("\n")[l - 1].strip()[:100]
    except Exception:
        text = ""
    print("%6d %5.1f%%  %s:%d  %-100s %s" % (v, 100.0 * v / max(tot, 1), os.path.basename(f), l, text, dict(why[(f, l)].most_common(2))))
