"""BASELINE.json config 5: lattice-size x walker-count sweep on one GPU through bench.py (device-timed value only).
Writes one JSON line per case to gpurun_out/config_sweep.jsonl and prints a table."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [(6, 1024), (6, 4096), (6, 16384), (6, 65536), (8, 4096), (8, 16384), (12, 1024), (12, 4096), (12, 16384), (18, 1024), (18, 4096)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in c.split("x")) for c in sys.argv[1:]]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "config_sweep.jsonl"), "w")
print("%6s %8s %14s %10s  %s" % ("sites", "walkers", "walker-sweeps/s", "ms/step", "kernel ms per step"))
for n, nw in cases:
    steps = 12 if n <= 8 else (6 if n == 12 else 3)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--lattice", str(n), "--walkers-per-gpu", str(nw), "--steps", str(steps),
           "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--thermalization", str(2 * 3 * n * n)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(n, nw, "FAILED", r.stderr[-300:])
        continue
    d = json.loads(line[-1])
    d["config"]["lattice_n"] = n
    out.write(json.dumps(d) + "\n"); out.flush()
    km = {k: round(v / d["steps"], 2) for k, v in d["kernel_ms"].items() if v > 0}
    print("%6d %8d %14.3e %10.2f  %s  E/site=%.5f" % (3 * n * n, nw, d["value"], d["ms_per_step"], km, d["observables"]["E_per_site"]), flush=True)
