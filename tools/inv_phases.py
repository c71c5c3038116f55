import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
eng = kd.Engine(ham, nw, 0); eng.set_option('inverse_variant', 1); eng.set_config(ku, kdn); eng.refresh()
W_ref = eng.get_W(3, 1).copy()
out = (C.c_longlong * 16)()
names = ["1 panel load", "2 panel LU", "3 publish+moves", "4 columns (U_K, pivot rows)", "5 panel cols", "6 GEMM update"]
# (variant 0 = k_inverse_v4 has four phases: load, pivot loop, publish + gather, GEMM update)
eng.set_option('gemm_variant', 1)
for variant, tuning in ((5, 0), (4, 0)):
    eng.set_option("inverse_variant", variant)
    eng.set_option("inverse_tuning", tuning)
    eng.refresh()
    eng._L.kdsl_debug_inverse_phases(eng._h, out)
    eng.set_profiling(True); eng.reset_timers()
    eng.refresh()
    eng._L.kdsl_debug_inverse_phases(eng._h, out)
    v = np.array(out[:6], dtype=float) / 2     # two launches (up, down)
    err = np.abs(eng.get_W(3, 1) - W_ref).max()
    print("variant %d tuning %d: cycles per matrix (CTA 0) total %.0f  |dW| vs first %.2e" % (variant, tuning, v.sum(), err))
    print("   " + "  ".join("%s=%.0f" % (nme.split()[0] + nme.split()[1][:5], c) for nme, c in zip(names, v)))
    if variant in (5, 6):
        print("    v5 (per matrix): phase A %.0f  team P factor %.0f  team G update %.0f (measured by warp 8)  gather %.0f  loop %.0f  P waits for G %.0f" % (v[0], v[1], v[2], v[3], v[4], v[5]))
    print("   ", {k: round(x["ms"], 3) for k, x in eng.timers().items() if x["ms"] > 0})
    g = np.array(out[5:8], dtype=float) / (2 * nw)
    print("    gemm CTA (blockIdx.x = 0) cycles: prologue %.0f  main loop %.0f  epilogue %.0f" % tuple(g))
