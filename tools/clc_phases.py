"""Developer probe: per-phase SM cycles of k_inverse_cl_c (cluster 0: its pivot CTA and its first update CTA).
Needs the instrumented library:  make -C kagomedsl.jl_b200/csrc libkdsl_ticks.so;  KDSL_LIB=.../libkdsl_ticks.so python tools/clc_phases.py 12 2048"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kagomedsl.jl_b200 as kd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
opts = [o.split("=") for o in sys.argv[3:]]
lat = kd.DoubleKagome(1.0, n, n, (True, True), (True, False)); ns = kd.ns(lat)
ham = kd.Hamiltonian(ns // 2, ns // 2, lat, B=0.02); ku, kdn = kd.init_conf_qr(ham, ns, ns // 2)
eng = kd.Engine(ham, nw, 0)
for k, v in opts:
    eng.set_option(k, int(v))
eng.set_config(ku, kdn); eng.refresh()
out = (C.c_longlong * 16)()
eng._L.kdsl_debug_inverse_phases(eng._h, out)
eng.set_profiling(True); eng.reset_timers()
eng.refresh()
eng._L.kdsl_debug_inverse_phases(eng._h, out)
v = np.array(out[8:15], dtype=float) / max(out[15], 1)
print("items of cluster 0: %d; cycles per matrix" % out[15])
print("  P CTA: next-panel update %.0f  factor %.0f  barrier wait %.0f   total %.0f" % (v[0], v[1], v[2], v[:3].sum()))
print("  G CTA: operand load %.0f  gather %.0f  update(warp 0) %.0f  barrier wait %.0f   total %.0f" % (v[3], v[4], v[5], v[6], v[3:7].sum()))
print("  ", {k: round(x["ms"], 3) for k, x in eng.timers().items() if x["ms"] > 0})
