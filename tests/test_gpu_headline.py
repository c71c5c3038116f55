"""GPU parity at the BASELINE.json headline sizes: oracle replays at 432 and 972 sites, zero-flux and B != 0 chains,
the 432-site energy check, and the call-order / two-handle regressions of the shared-memory opt-in.

Bars as in test_gpu_parity.py: configurations / Z_mu / counters bit-exact under a replayed proposal sequence, W and
O_L within 1e-10 relative of the FP64 (ComplexF64 for B != 0) oracle; <E>/site within 4 sigma of the oracle chain.
"""
import json
import os
import threading

import numpy as np
import pytest

import _util as U

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kd():
    import kagomedsl.jl_b200 as kd_
    if kd_._lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return kd_


def chain_states(kd, ham, ns, N_up, nw, n_therm, seed):
    """`nw` decorrelated, chain-typical Mott configurations: the QR start state (src/MonteCarlo.jl:326-357) advanced by
    `n_therm` sweeps on the GPU.  Only a generator of valid starting points: every comparison below starts from them
    on BOTH sides."""
    ku0, kd0 = kd.init_conf_qr(ham, ns, N_up)
    eng = kd.Engine(ham, nw)
    eng.set_config(ku0, kd0)
    eng.set_rng(kd.walker_states(seed, nw))
    eng.refresh()
    eng.sweep(n_therm, -1)
    ku, kdn = eng.get_config()
    eng.close()
    return ku.astype(np.int64), kdn.astype(np.int64)


def replay_against_oracle(kd, ham, ku, kdn, r, bond, chunks, options=None, dtype="f64", check_W=True):
    """advance GPU and oracle by the same replayed proposals, compare after every chunk"""
    nw = ku.shape[0]
    eng = kd.Engine(ham, nw)
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    eng.set_config(ku, kdn)
    eng.refresh()
    orc = U.oracle_walkers(ham, ku, kdn, dtype=dtype)
    done = 0
    for chunk in chunks:
        eng.replay(r[done:done + chunk], bond[done:done + chunk])
        for w, mc in enumerate(orc):
            for s in range(done, done + chunk):
                mc.sweep(replay=(r[s, w], int(bond[s, w]), 1))
                mc.sweeps = mc.sweeps + 1
        done += chunk
        gku, gkd = eng.get_config()
        z, zr = eng.Z()
        ol = eng.measure()                                                   # Woodbury form: pending updates included
        for w, mc in enumerate(orc):
            oku, okd = mc.kappa()
            assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd), f"kappa differs, walker {w} after {done} sweeps"
            assert z[w] == zr[w] == U.O.Z(ham.nn, oku, okd)
            ref = mc.getOL()
            assert abs(ol[w] - ref) <= TOL * max(1.0, abs(ref)), f"O_L differs, walker {w} after {done} sweeps"
            if check_W:
                Wu, Wd = mc.W()
                assert U.relerr(eng.get_W(w, 0), Wu) < TOL
                assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    acc, acc_w, _ = eng.accumulators(per_walker=True)
    for w, mc in enumerate(orc):
        assert acc_w[w] == mc.counters()[0]
    assert acc[kd._lib.ACC_N_REACH] == sum(mc.counters()[1] for mc in orc)
    assert acc[kd._lib.ACC_N_REFRESH] == sum(mc.counters()[2] for mc in orc)
    tm = eng.timers()
    eng.close()
    return acc, tm


@pytest.mark.parametrize("options", [{}, {"update_variant": 0}, {"inverse_variant": 5}])
def test_replay_432_sites(kd, options):
    """BASELINE config 3 lattice (12x12, 432 sites): 4 walkers x 500 replayed sweeps across the re-evaluations at sweeps
    216 and 432, default path (Woodbury updates + k_reeval_fused; k_flush_wb runs with two row blocks per species here),
    the reference-style rank-1 path and the three-kernel re-evaluation."""
    lat, ham = U.problem(12, 12)
    ns, nw, n = kd.ns(lat), 4, 500
    ku, kdn = chain_states(kd, ham, ns, ns // 2, nw, 3000, 41)
    rng = np.random.default_rng(432)
    # walkers 2, 3 draw r from [0, 0.45): several times the acceptance rate, so that flushes (>= 16 pending updates)
    # happen many times between two re-evaluations
    r = rng.random((n, nw)) * np.array([1.0, 1.0, 0.45, 0.45])
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    acc, tm = replay_against_oracle(kd, ham, ku, kdn, r, bond, (1, 215, 1, 120, 163), options)
    assert acc[kd._lib.ACC_N_REFRESH] >= 3
    if not options:
        assert tm["update"]["flushes"] >= 4                                  # the delayed-update pass really ran


def test_replay_972_sites(kd):
    """BASELINE config 4 lattice (18x18, 972 sites, N = 486: three-kernel re-evaluation with the RPT = 2 inverse):
    2 walkers x 1000 replayed sweeps across the re-evaluations at sweeps 486 and 972"""
    lat, ham = U.problem(18, 18)
    ns, nw, n = kd.ns(lat), 2, 1000
    ku, kdn = chain_states(kd, ham, ns, ns // 2, nw, 4000, 97)
    rng = np.random.default_rng(972)
    r = rng.random((n, nw)) * np.array([1.0, 0.45])
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    acc, tm = replay_against_oracle(kd, ham, ku, kdn, r, bond, (486, 1, 486, 27))
    assert acc[kd._lib.ACC_N_REFRESH] >= 2 and tm["update"]["flushes"] >= 2


def test_replay_zero_flux_108_sites(kd):
    """BASELINE config 2, zero-flux tables (scripts/zero_flux.jl:15-42): 6 walkers x 1200 replayed sweeps"""
    lat, ham = U.problem(6, 6, flux="zero")
    ns, nw, n = kd.ns(lat), 6, 1200
    ku, kdn = chain_states(kd, ham, ns, ns // 2, nw, 2000, 13)
    rng = np.random.default_rng(108)
    r = rng.random((n, nw))
    r[:, 3:] *= 0.45
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    replay_against_oracle(kd, ham, ku, kdn, r, bond, (1, 53, 54, 400, 692))


def test_replay_c128_8x8_landau_level(kd):
    """scripts/LL.jl:11-25: 8x8 (192 sites), antiPBC = (false, true), Peierls flux B = imbalance * pi / (n1 n2 2 sqrt 3)
    with imbalance = 2, N_up = 97, N_down = 95: ComplexF64 engine against the c128 oracle, 4 walkers x 400 sweeps"""
    n1 = n2 = 8
    B = 2 * np.pi / (n1 * n2 * 2 * np.sqrt(3.0))
    lat, ham = U.problem(n1, n2, (True, True), (False, True), "pi", 97, B)
    assert np.iscomplexobj(ham.U_up) and np.abs(np.asarray(ham.U_up).imag).max() > 1e-3
    ns, nw, n = kd.ns(lat), 4, 400
    ku, kdn = chain_states(kd, ham, ns, 97, nw, 1500, 5)
    rng = np.random.default_rng(192)
    r = rng.random((n, nw)) * np.array([1.0, 1.0, 0.45, 0.45])
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    acc, _ = replay_against_oracle(kd, ham, ku, kdn, r, bond, (1, 94, 1, 200, 104), dtype="c128")
    assert acc[kd._lib.ACC_N_REFRESH] >= 3


@pytest.mark.parametrize("options", [{}, {"inverse_variant": 7, "flush_variant": 4, "gemm_variant": 4}])
def test_replay_c128_432_sites(kd, options):
    """ComplexF64 engine at the headline lattice (12x12, 432 sites, Peierls flux): N = 216 runs the production instantiations
    of the complex tensor-pipe kernels -- k_inverse_cl_c with panels of 16 columns and two CTAs per SM, k_gemm_W_dmma_c with
    3 x 9 tiles, k_flush_dmma_c with two 216-row blocks -- against the c128 oracle: 3 walkers x 450 replayed sweeps across the
    re-evaluations at sweeps 216 and 432; two walkers driven at ~3x the acceptance rate so that the flush runs often.
    Second case: the embedding inverse, the FMA flush and the FMA product on the same stream."""
    lat, ham = U.problem(12, 12, (True, True), (True, False), "pi", None, 0.02)
    assert np.iscomplexobj(ham.U_up) and np.abs(np.asarray(ham.U_up).imag).max() > 1e-3
    ns, nw, n = kd.ns(lat), 3, 450
    ku, kdn = chain_states(kd, ham, ns, ns // 2, nw, 2000, 11)
    rng = np.random.default_rng(4320)
    r = rng.random((n, nw)) * np.array([1.0, 0.45, 0.45])
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    acc, tm = replay_against_oracle(kd, ham, ku, kdn, r, bond, (1, 215, 1, 120, 113), options, dtype="c128")
    assert acc[kd._lib.ACC_N_REFRESH] >= 2 and tm["update"]["flushes"] >= 3


def test_c128_972_sites_refresh_and_rank1_path(kd):
    """ComplexF64 engine at 18x18 (972 sites, N = 486 > 256): k_inverse_cl_c with 512-thread CTAs and panels of 8 columns,
    k_gemm_W_dmma_c with 7 x 21 tiles; the Woodbury kernels' operands exceed the shared memory at this size, so the engine
    falls back to the reference's immediate rank-1 update (documented).  W against numpy, then 60 device-RNG sweeps:
    maintained W == re-evaluated W, incremental Z_mu == recount."""
    lat, ham = U.problem(18, 18, (True, True), (True, False), "pi", None, 0.01)
    ns, nw = kd.ns(lat), 3
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
    eng = kd.Engine(ham, nw)
    assert eng.is_complex
    eng.set_config(ku0, kd0)
    eng.set_rng(kd.walker_states(972, nw))
    eng.refresh()
    Uu, Ud = np.asarray(ham.U_up), np.asarray(ham.U_down)
    assert U.relerr(eng.get_W(0, 0), Uu @ np.linalg.inv(kd.tilde_U(Uu, ku0))) < 1e-9
    assert U.relerr(eng.get_W(2, 1), Ud @ np.linalg.inv(kd.tilde_U(Ud, kd0))) < 1e-9
    eng.sweep(60, -1)
    ku, kdn = eng.get_config()
    z, zr = eng.Z()
    assert np.array_equal(z, zr)
    Wm = [eng.get_W(w, 0) for w in range(nw)]
    assert any(not np.array_equal(ku[w], ku0) for w in range(nw))               # the chains moved
    eng.refresh()
    for w in range(nw):
        Wr = Uu @ np.linalg.inv(kd.tilde_U(Uu, ku[w]))
        assert U.relerr(eng.get_W(w, 0), Wr) < 1e-9 and U.relerr(Wm[w], Wr) < 1e-8
    eng.close()


def test_chains_are_reproducible_and_Z_is_right_for_every_handle_size(kd):
    """Regression (round 2): host-to-device copies used to run on the legacy default stream while the engine's stream is
    non-blocking; the kernel that counts Z_mu after set_config could start before the tail of the kappa upload had landed.
    A 512-walker handle at 432 sites showed it: Z_mu of the last walkers wrong, chains different from run to run.  All
    copies now go through the engine's stream.  Two runs must agree bit for bit, Z_mu must equal the recount right after
    set_config, and a handle of 512 walkers must reproduce walkers 0..511 of a 1024-walker handle."""
    lat, ham = U.problem(12, 12)
    ns = kd.ns(lat)
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
    states = kd.walker_states(1234, 1024)

    def run(nw):
        eng = kd.Engine(ham, nw)
        eng.set_config(ku0, kd0)
        z, zr = eng.Z()
        assert np.array_equal(z, zr) and np.all(z == z[0])
        eng.set_rng(states[:nw])
        eng.refresh()
        eng.sweep(120, -1)
        ku, kdn = eng.get_config()
        z, zr = eng.Z()
        assert np.array_equal(z, zr)
        rng = eng.get_rng()
        eng.close()
        return ku, kdn, rng

    a = run(512)
    for _ in range(3):
        b = run(512)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    c = run(1024)
    assert all(np.array_equal(x, y[:512]) for x, y in zip(a, c))


@pytest.mark.parametrize("n,flux,B,nw,n_sweeps,picks", [
    (12, "pi", 0.0, 4096, 450, (0, 7, 8, 1023, 2048, 3333, 4094, 4095)),      # headline batch
    (6, "pi", 0.0, 4096, 500, (0, 443, 444, 2047, 4095)),                     # BASELINE config 2 on k_resident (444 resident CTAs)
    (6, "zero", 0.0, 4096, 500, (1, 445, 4000)),
    (12, "pi", 0.02, 4096, 450, (0, 9, 2222, 4095)),                          # ComplexF64 engine, full batch
    (18, "pi", 0.0, 1024, 500, (0, 511, 1023)),                               # BASELINE config 4: cluster inverse + product
])
def test_full_batch_matches_oracle_chains(kd, n, flux, B, nw, n_sweeps, picks):
    """BASELINE's full single-GPU batches (device Xoshiro streams, sweeps across the periodic re-evaluations): walkers
    picked across the batch (first, last, CTA / resident-slot / list boundaries) must be bit-identical to the oracle's
    Carlo loop started from the same state -- kappa, acceptance count, O_L sum, W -- and every walker's incremental
    Z_mu must equal the recount.  (The small replay tests cannot see effects that need a full machine: thousands of
    (walker, species) items per persistent CTA, hundreds of flushes per launch, every resident slot taken.)"""
    lat, ham = U.problem(n, n, (True, True), (True, False), flux, None, B)
    dtype = "c128" if B != 0.0 else "f64"
    ns = kd.ns(lat)
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
    states = kd.walker_states(2026 + n, nw)
    eng = kd.Engine(ham, nw)
    assert eng.is_complex == (B != 0.0)
    eng.set_config(ku0, kd0)
    eng.set_rng(states)
    eng.refresh()
    eng.sweep(n_sweeps, thermalization=100)
    gku, gkd = eng.get_config()
    z, zr = eng.Z()
    assert np.array_equal(z, zr)
    acc, acc_w, ol_w = eng.accumulators(per_walker=True)
    assert acc[kd._lib.ACC_N_SINGULAR] == 0
    for w in picks:
        mc = U.O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, dtype)
        mc.set_kappa(ku0, kd0)
        mc.reevaluateW()
        st, _ = mc.run(U.O.Xoshiro(states[w]), n_sweeps, 100)
        oku, okd = mc.kappa()
        assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd), f"walker {w}: configuration differs from the oracle chain"
        assert acc_w[w] == st[0]
        assert abs(ol_w[w] - st[1]) <= 1e-9 * max(1.0, abs(st[1]))
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL and U.relerr(eng.get_W(w, 1), Wd) < TOL
    eng.close()


def test_measure_right_after_refresh_432_and_two_handles(kd):
    """Regression: set_config -> refresh -> measure as the FIRST calls on a handle at 432 sites (k_measure_wb needs 92 KB
    of dynamic shared memory; the opt-in used to be set inside the first flush launch, process-wide), and a second
    handle in the same process."""
    lat, ham = U.problem(12, 12)
    ns, nw = kd.ns(lat), 3
    ku, kdn = chain_states(kd, ham, ns, ns // 2, nw, 1500, 3)
    refs = [mc.getOL() for mc in U.oracle_walkers(ham, ku, kdn)]
    eng_a = kd.Engine(ham, nw)
    eng_b = kd.Engine(ham, nw)
    for eng in (eng_a, eng_b):
        eng.set_config(ku, kdn)
        eng.refresh()
        ol = eng.measure()
        for w in range(nw):
            assert abs(ol[w] - refs[w]) <= TOL * max(1.0, abs(refs[w]))
    # resume just below a measurement sweep: the cadence measurement is the first Woodbury kernel of the handle
    eng_c = kd.Engine(ham, nw)
    eng_c.set_config(ku, kdn)
    eng_c.refresh()
    eng_c.sweeps = ns // 2 - 1
    eng_c.sweep(1, thermalization=0)
    eng_c.synchronize()
    assert eng_c.accumulators()[kd._lib.ACC_N_OL] == nw
    lat18, ham18 = U.problem(18, 18)
    e18 = kd.Engine(ham18, 1)
    with pytest.raises(kd.KdslError, match="shared memory"):                 # kmax = 32 at 972 sites does not fit: clear
        e18.set_option("flush_every", 16)                                    # error at the option, not a launch failure
    e18.close()
    for eng in (eng_a, eng_b, eng_c):
        eng.close()


def test_singular_walker_is_frozen_and_reported(kd):
    """A walker whose re-evaluation met a singular tilde_U stops proposing, keeps its flag, gives no O_L samples, and
    kdsl_synchronize reports KDSL_ERR_SINGULAR (the SingularException of src/MonteCarlo.jl:596-603); the other walkers
    of the batch carry on."""
    lat, ham = U.problem(2, 2, (False, False), (False, False))
    Ud = np.zeros((12, 6)); Ud[:6] = np.eye(6)                               # down orbitals live on sites 1..6 only
    ham2 = kd.Hamiltonian(6, 6, ham.U_up, Ud, ham.H_mat, ham.nn)
    ku = np.zeros((2, 12), dtype=np.int64); kdn = np.zeros((2, 12), dtype=np.int64)
    for w, dn_sites in enumerate((np.arange(6, 12), np.arange(0, 6))):       # walker 0 singular, walker 1 fine
        kdn[w, dn_sites] = np.arange(1, 7)
        ku[w, np.setdiff1d(np.arange(12), dn_sites)] = np.arange(1, 7)
    eng = kd.Engine(ham2, 2)
    eng.set_config(ku, kdn)
    with pytest.raises(kd.SingularException):
        eng.refresh()
    eng.set_rng(kd.walker_states(4, 2))
    eng.sweep(300, thermalization=0)
    with pytest.raises(kd.SingularException):
        eng.synchronize()
    gku, gkd = eng.get_config()
    assert np.array_equal(gku[0], ku[0]) and np.array_equal(gkd[0], kdn[0])  # frozen
    fl = eng.flags()
    assert fl[0] & 1
    _, n_ol = eng.last_OL()
    assert n_ol[0] == 0
    eng.close()


@pytest.mark.slow
def test_energy_432_sites_within_error_bars(kd):
    """north_star correctness (3): <E>/site at the 432-site pi-flux DSL within statistical error bars of the reference
    chain.  GPU: 4096 walkers x 40 bins; oracle (f64 instantiation; the law of the chain does not depend on the storage
    type): one walker per host core x 500 bins, run here, plus the long oracle run committed under tests/golden/
    (16 walkers x 8000 bins, made by tools/energy_check.py)."""
    lat, ham = U.problem(12, 12)
    ns = kd.ns(lat)
    n_occ = ns // 2
    ku0, kd0 = kd.init_conf_qr(ham, ns, n_occ)
    nw, bins_gpu, therm = 4096, 40, 20 * ns // n_occ * n_occ
    eng = kd.Engine(ham, nw)
    eng.set_config(ku0, kd0)
    eng.set_rng(kd.walker_states(1234, nw))
    eng.refresh()
    eng.sweep(therm, -1)
    eng.reset_accumulators()
    eng.sweep(bins_gpu * n_occ, 0)
    acc, acc_w, ol_w = eng.accumulators(per_walker=True)
    eng.close()
    e_w = ol_w / bins_gpu / ns
    e_gpu, s_gpu = e_w.mean(), e_w.std(ddof=1) / np.sqrt(nw)
    assert acc[kd._lib.ACC_N_SINGULAR] == 0 and s_gpu < 5e-5
    # the committed long oracle chain
    gold = json.load(open(os.path.join(GOLDEN, "energy_432.json")))
    e_ref, s_ref = gold["oracle_chain"]["E_per_site"], gold["oracle_chain"]["stderr"]
    assert abs(e_gpu - e_ref) < 4.0 * np.hypot(s_gpu, s_ref), (e_gpu, s_gpu, e_ref, s_ref)
    # a live oracle chain on the host cores (bounded: ~15 s)
    cores = min(os.cpu_count() or 1, 32)
    bins_cpu = 500
    bonds = np.asarray(ham.nn, dtype=np.int32)
    res = [None] * cores

    def work(t):
        mc = U.O.MC(bonds, ham.U_up, ham.U_down, "f64")
        mc.set_kappa(ku0, kd0)
        mc.reevaluateW()
        g = U.O.Xoshiro.from_seed(99 + 7919 * t)
        mc.run(g, therm, 10 ** 12)
        st = np.zeros(4)
        mc.run(g, bins_cpu * n_occ, 0, stats=st)
        res[t] = st
    ths = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    [th.start() for th in ths]
    [th.join() for th in ths]
    e_c = np.array([r_[1] / r_[3] / ns for r_ in res])
    e_cpu, s_cpu = e_c.mean(), e_c.std(ddof=1) / np.sqrt(cores)
    assert abs(e_gpu - e_cpu) < 4.0 * np.hypot(s_gpu, s_cpu), (e_gpu, s_gpu, e_cpu, s_cpu)


def test_energy_c128_landau_level_192_sites_within_error_bars(kd):
    """The same statistical check for the ComplexF64 engine: scripts/LL.jl's 8x8 lattice (192 sites, antiPBC = (false, true),
    Peierls flux with imbalance 2, N_up = 97, N_down = 95).  GPU: 2048 walkers x 40 bins on the complex tensor-pipe kernels;
    oracle (c128 instantiation): one walker per host core x 600 bins, run here.  |dE| < 4 sigma."""
    n1 = n2 = 8
    B = 2 * np.pi / (n1 * n2 * 2 * np.sqrt(3.0))
    lat, ham = U.problem(n1, n2, (True, True), (False, True), "pi", 97, B)
    ns = kd.ns(lat)
    n_occ = min(ham.N_up, ham.N_down)
    ku0, kd0 = kd.init_conf_qr(ham, ns, ham.N_up)
    nw, bins_gpu, therm = 2048, 40, 40 * n_occ
    eng = kd.Engine(ham, nw)
    assert eng.is_complex
    eng.set_config(ku0, kd0)
    eng.set_rng(kd.walker_states(4321, nw))
    eng.refresh()
    eng.sweep(therm, -1)
    eng.reset_accumulators()
    eng.sweep(bins_gpu * n_occ, 0)
    acc, acc_w, ol_w = eng.accumulators(per_walker=True)
    eng.close()
    assert acc[kd._lib.ACC_N_SINGULAR] == 0 and acc[kd._lib.ACC_N_OL] == nw * bins_gpu
    e_w = ol_w / bins_gpu / ns
    e_gpu, s_gpu = e_w.mean(), e_w.std(ddof=1) / np.sqrt(nw)
    cores = min(os.cpu_count() or 1, 32)
    bins_cpu = 600
    bonds = np.asarray(ham.nn, dtype=np.int32)
    res = [None] * cores

    def work(t):
        mc = U.O.MC(bonds, ham.U_up, ham.U_down, "c128")
        mc.set_kappa(ku0, kd0)
        mc.reevaluateW()
        g = U.O.Xoshiro.from_seed(77 + 7919 * t)
        mc.run(g, therm, 10 ** 12)
        st = np.zeros(4)
        mc.run(g, bins_cpu * n_occ, 0, stats=st)
        res[t] = st
    ths = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    [th.start() for th in ths]
    [th.join() for th in ths]
    e_c = np.array([r_[1] / r_[3] / ns for r_ in res])
    e_cpu, s_cpu = e_c.mean(), e_c.std(ddof=1) / np.sqrt(cores)
    print("E/site c128 192 sites: GPU %.6f +- %.6f, oracle chain %.6f +- %.6f" % (e_gpu, s_gpu, e_cpu, s_cpu))
    assert s_gpu < 2e-4 and abs(e_gpu - e_cpu) < 4.0 * np.hypot(s_gpu, s_cpu), (e_gpu, s_gpu, e_cpu, s_cpu)


def test_two_host_threads_drive_two_handles_concurrently(kd):
    """One handle per host thread (SURVEY 8(b) threading): two threads sweep their own handles on the same GPU at the same
    time (ctypes releases the GIL, the kernels of the two streams interleave); each must end exactly where it ends alone."""
    lat, ham = U.problem(8, 8)
    ns, nw, n = kd.ns(lat), 256, 400
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)

    def run(first, out, k):
        eng = kd.Engine(ham, nw)
        eng.set_config(ku0, kd0)
        eng.set_rng(kd.walker_states(99, nw, first_walker=first))
        eng.refresh()
        for _ in range(4):
            eng.sweep(n // 4, thermalization=50)
        out[k] = (eng.get_config(), eng.accumulators(per_walker=True), eng.get_rng())
        eng.close()

    alone = {}
    run(0, alone, 0)
    run(nw, alone, 1)
    both = {}
    ths = [threading.Thread(target=run, args=(k * nw, both, k)) for k in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for k in range(2):
        (ku_a, kd_a), (acc_a, accw_a, ol_a), rng_a = alone[k]
        (ku_b, kd_b), (acc_b, accw_b, ol_b), rng_b = both[k]
        assert np.array_equal(ku_a, ku_b) and np.array_equal(kd_a, kd_b) and np.array_equal(rng_a, rng_b)
        assert np.array_equal(accw_a, accw_b) and np.array_equal(ol_a, ol_b) and np.array_equal(acc_a, acc_b)


def test_two_gpus_reduce_accumulators_through_the_c_abi(kd):
    """kdsl_comm_init_all + kdsl_group_accumulators_allreduce: one process, one handle per GPU (the Julia host's layout);
    the NCCL sum equals the sum of the per-handle vectors.  Needs two visible GPUs (gpurun --gpus 2)."""
    if kd._lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    lat, ham = U.problem(4, 3)
    ns, nw = kd.ns(lat), 32
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
    engines = []
    for dev in (0, 1):
        eng = kd.Engine(ham, nw, device=dev)
        eng.set_config(ku0, kd0)
        eng.set_rng(kd.walker_states(77, nw, first_walker=dev * nw))
        eng.refresh()
        engines.append(eng)
    grp = kd.dist.Group(engines)
    assert engines[1].comm_info() == (1, 2)
    grp.sweep(500, thermalization=36)
    total = grp.accumulators()
    parts = [e.accumulators() for e in engines]
    assert np.allclose(total, parts[0] + parts[1], rtol=1e-13, atol=0)
    assert total[kd._lib.ACC_WALKER_SWEEPS] == 2 * nw * 500
    assert not np.array_equal(parts[0], parts[1])                            # disjoint streams
    for e in engines:
        e.close()
