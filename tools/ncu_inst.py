"""Executed warp instructions per SOURCE line of one kernel (same join as ncu_lines.py, counting 'Instructions Executed').
usage: ncu_inst.py report.ncu-rep libkdsl.so KERNEL_MANGLED_PREFIX [top]"""
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
base = int(rows[2][0], 16)
agg, ops, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    key = off2line.get(int(r[0], 16) - base, ("?", 0))
    n = int(r[idx["Instructions Executed"]] or 0)
    agg[key] += n
    tot += n
    ops[key][r[1].split()[0] if r[1].split()[0][0] != "@" else r[1].split()[1]] += n
print("# %s: %d warp instructions" % (rows[0][1], tot))
for (f, l), v in agg.most_common(top):
    try:
        text = open(f).read().split("\n")[l - 1].strip()[:90]
    except Exception:
        text = ""
    print("%11d %5.1f%%  %s:%d  %-90s %s" % (v, 100.0 * v / max(tot, 1), os.path.basename(f), l, text, dict(ops[(f, l)].most_common(4))))
