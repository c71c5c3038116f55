"""Developer probe: k_reeval_fused (inverse_variant 6 / default) against the separate inverse + GEMM path (variant 5)
and the simple kernels (variant 1): max |dW|, refresh time of `nw` walkers, phase cycles of CTA 0."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tunings = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
eng = kd.Engine(ham, nw, 0)
eng.set_config(ku, kdn); eng.set_rng(kd.walker_states(1234, nw))
eng.set_option('inverse_variant', 5); eng.refresh()
eng.sweep(2 * ns, -1)                       # decorrelate the walkers
eng.set_option('inverse_variant', 1); eng.refresh()
ws = [0, 3, nw - 1]
W_ref = [(eng.get_W(w, 0).copy(), eng.get_W(w, 1).copy()) for w in ws]
out = (C.c_longlong * 16)()
ctas = int(os.environ.get("FUSED_CTAS", "0"))
eng.set_option("fused_ctas", ctas)
for variant, tuning in [(5, 0)] + [(6, t) for t in tunings]:
    eng.set_option("inverse_variant", variant)
    eng.set_option("inverse_tuning", tuning)
    eng.refresh()
    eng._L.kdsl_debug_inverse_phases(eng._h, out)
    eng.set_profiling(True); eng.reset_timers()
    for _ in range(3):
        eng.refresh()
    eng._L.kdsl_debug_inverse_phases(eng._h, out)
    err = max(max(np.abs(eng.get_W(w, 0) - a).max(), np.abs(eng.get_W(w, 1) - b).max()) for w, (a, b) in zip(ws, W_ref))
    tm = {k: round(x["ms"] / 3, 3) for k, x in eng.timers().items() if x["ms"] > 0}
    v = np.array(out[:8], dtype=float)
    print("variant %d tuning %d: |dW| vs variant 1 = %.2e   ms per refresh of %d walkers: %s  total %.3f" % (variant, tuning, err, nw, tm, sum(tm.values())))
    print("    CTA 0 phase cycles (sum over its items, 3 refreshes):", [int(x) for x in v])
    eng.set_profiling(False)
eng.close()
