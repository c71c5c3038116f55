"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): configurations / Z / counters bit-exact under a replayed proposal
sequence; W, determinant ratios and O_L within 1e-10 relative of the FP64 reference path.
"""
import numpy as np
import pytest

import _util as U

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def kd():
    import kagomedsl.jl_b200 as kd_
    if kd_._lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    return kd_


@pytest.mark.parametrize("n1,n2,PBC,anti,flux", [
    (2, 2, (False, False), (False, False), "pi"),       # the reference's own test lattice (test-MonteCarlo.jl:480)
    (4, 3, (True, True), (True, False), "pi"),          # SURVEY config 1
    (6, 6, (True, True), (True, False), "pi"),          # config 2
    (6, 6, (True, True), (True, False), "zero"),
])
def test_refresh_matches_oracle(kd, n1, n2, PBC, anti, flux):
    lat, ham = U.problem(n1, n2, PBC, anti, flux)
    ns, nw = kd.ns(lat), 6
    rng = np.random.default_rng(11)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw)
    eng = kd.Engine(ham, nw)
    eng.set_config(ku, kdn)
    eng.refresh()
    z, zr = eng.Z()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd) < TOL
        assert z[w] == zr[w] == U.O.Z(ham.nn, ku[w], kdn[w])
    # rows of W at occupied sites are unit vectors (SURVEY 8(a) invariant)
    W0 = eng.get_W(0, 0)
    occ = np.nonzero(ku[0])[0]
    assert np.allclose(W0[occ, :][np.arange(len(occ)), ku[0][occ] - 1], 1.0, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("variant", [0, 1, 4, 5, 6, 7])
def test_refresh_imbalanced_filling_and_variants(kd, variant):
    """N_up != N_down (scripts/FP.jl), N not a multiple of 8; fused re-evaluation (0 / 6), simple kernels (1),
    gather + blocked DMMA inverse + product (4 / 5), cluster inverse (7)"""
    lat, ham = U.problem(4, 3, N_up=20)
    ns, nw = kd.ns(lat), 5
    rng = np.random.default_rng(21)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    eng = kd.Engine(ham, nw)
    eng.set_option("inverse_variant", variant)
    eng.set_config(ku, kdn)
    eng.refresh()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    eng.close()


@pytest.mark.parametrize("n,N_up,cluster,row_slices,nw", [
    (6, None, 4, 8, 90), (6, 50, 2, 1, 20), (8, None, 4, 8, 50), (8, 90, 5, 3, 50), (8, None, 8, 2, 24), (8, None, 3, 64, 24)])
def test_cluster_inverse_matches_oracle(kd, n, N_up, cluster, row_slices, nw):
    """k_inverse_cl (one matrix per thread-block cluster: 1 pivot CTA + cluster-1 update CTAs; the default for
    256 < Np <= 512, selected here on small lattices by inverse_variant 7): several panels of 24 columns incl. a
    partial last one, more (walker, species) items than resident clusters, all cluster sizes' work splits"""
    lat, ham = U.problem(n, n, N_up=N_up)
    ns = kd.ns(lat)
    rng = np.random.default_rng(100 + n + cluster)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    eng = kd.Engine(ham, nw)
    eng.set_option("inverse_variant", 7)
    eng.set_option("inverse_cluster", cluster)
    eng.set_option("inverse_row_slices", row_slices)
    eng.set_config(ku, kdn)
    for rep in range(2):                                         # the second pass reuses the per-cluster scratch buffers
        eng.set_W(0, 0, np.zeros((ns, ham.N_up)))
        eng.refresh()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    assert eng.accumulators()[kd._lib.ACC_N_SINGULAR] == 0
    eng.close()


@pytest.mark.parametrize("n1,n2,N_up,cluster,row_slices,nw", [
    (6, 6, None, 4, 4, 90), (6, 6, 50, 2, 1, 20), (8, 8, None, 4, 4, 50), (8, 8, 90, 5, 3, 50), (8, 8, None, 8, 2, 24),
    (8, 8, None, 3, 16, 24), (4, 3, 20, 4, 4, 40), (4, 3, None, 2, 4, 9)])
def test_cluster_reeval_matches_oracle(kd, n1, n2, N_up, cluster, row_slices, nw):
    """k_reeval_cl (the whole of reevaluateW! in one kernel, one matrix per thread-block cluster; the default for
    256 < Np <= 512, forced here on small lattices by inverse_variant 8): single-panel matrices (first step = last step),
    partial last panels, N_up != N_down, more items than resident clusters, every cluster size's work split"""
    lat, ham = U.problem(n1, n2, N_up=N_up)
    ns = kd.ns(lat)
    rng = np.random.default_rng(300 + n1 + cluster)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    eng = kd.Engine(ham, nw)
    try:
        eng.set_option("inverse_variant", 8)
    except kd.KdslError:
        eng.close()
        pytest.skip("k_reeval_cl is a developer variant (make DEV=1): measured slower than the product paths, DESIGN.md 4.6")
    eng.set_option("reeval_cluster", cluster)
    eng.set_option("reeval_rs", row_slices)
    eng.set_config(ku, kdn)
    for rep in range(2):                                         # the second pass reuses the per-cluster workspaces
        eng.set_W(0, 0, np.zeros((ns, ham.N_up)))
        eng.refresh()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    assert eng.accumulators()[kd._lib.ACC_N_SINGULAR] == 0
    eng.close()


@pytest.mark.parametrize("variant", [7, 8])
def test_cluster_inverse_singular_matrix_is_flagged(kd, variant):
    """a singular tilde_U inside a batch of the cluster inverse: that walker is flagged, the clusters go on with the
    other items (the singular flag travels through the per-cluster scratch: every CTA leaves the item at the same step)"""
    lat, ham = U.problem(6, 6)
    ns, nw = kd.ns(lat), 45
    rng = np.random.default_rng(77)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    Ud = np.array(ham.U_down, dtype=np.float64).copy()
    bad_sites = np.nonzero(kdn[7])[0]
    Ud[bad_sites[0], :] = 0.0                                     # walker 7 (and whoever shares the site): a zero row in tilde_U_down
    ham2 = kd.Hamiltonian(ham.N_up, ham.N_down, ham.U_up, Ud, ham.H_mat, ham.nn)
    expect_bad = np.array([kdn[w, bad_sites[0]] != 0 for w in range(nw)])
    assert expect_bad[7] and not expect_bad.all()
    eng = kd.Engine(ham2, nw)
    try:
        eng.set_option("inverse_variant", variant)
    except kd.KdslError:
        eng.close()
        pytest.skip("developer variant (make DEV=1)")
    eng.set_config(ku, kdn)
    with pytest.raises(kd.SingularException):
        eng.refresh()
    fl = eng.flags()
    assert np.array_equal((fl & 1) != 0, expect_bad)
    for w in np.nonzero(~expect_bad)[0][:6]:
        Wd_ref = Ud @ np.linalg.inv(kd.tilde_U(Ud, kdn[w]))
        assert U.relerr(eng.get_W(int(w), 1), Wd_ref) < 1e-8
    eng.close()


@pytest.mark.parametrize("cols_per_item,N_up", [(8, None), (5, None), (27, None), (7, 20)])
def test_update_W_matches_oracle_and_formula(kd, cols_per_item, N_up):
    """rank-1 kernel (flat walk over a slab of `cols_per_item` columns): slabs that do not divide N, a slab larger
    than N, and imbalanced filling"""
    lat, ham = U.problem(4, 3, N_up=N_up)
    ns, nw = kd.ns(lat), 5
    rng = np.random.default_rng(5)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_cols_per_item", cols_per_item)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw)
    eng.set_config(ku, kdn)
    eng.refresh()
    Wu0 = [rng.standard_normal((ns, ham.N_up)) for _ in range(nw)]
    Wd0 = [rng.standard_normal((ns, ham.N_down)) for _ in range(nw)]
    for w in range(nw):
        eng.set_W(w, 0, Wu0[w])
        eng.set_W(w, 1, Wd0[w])
    walkers = np.array([3, 0, 4], dtype=np.int32)
    l_up = np.array([1, ham.N_up, 7]); K_up = np.array([ns, 1, 20])
    l_dn = np.array([2, 5, ham.N_down]); K_dn = np.array([1, ns, 9])
    eng.set_option("update_variant", 0)                                         # the immediate rank-1 kernel
    eng.update_W(walkers, l_up, K_up, l_dn, K_dn)
    for m, w in enumerate(walkers):
        for spin, (W0, l, K) in enumerate(((Wu0[w], l_up[m], K_up[m]), (Wd0[w], l_dn[m], K_dn[m]))):
            ref = np.asfortranarray(W0.copy())
            U.O.update_W(ref, int(l), int(K), "f64")
            got = eng.get_W(int(w), spin)
            assert np.array_equal(got, ref) or U.relerr(got, ref) < 1e-14      # same fma order as the oracle
            # element formula of the reference test (test-MonteCarlo.jl:235-252)
            I, j = 3, 4
            d = 1.0 if j == l else 0.0
            assert np.isclose(got[I - 1, j - 1], W0[I - 1, j - 1] - W0[I - 1, l - 1] / W0[K - 1, l - 1] * (W0[K - 1, j - 1] - d), rtol=1e-12)
    for w in (1, 2):                                                           # untouched walkers
        assert np.array_equal(eng.get_W(w, 0), Wu0[w])
    eng.close()


@pytest.mark.parametrize("variant", [3, 2, 0])
@pytest.mark.parametrize("n1,n2,nw,n_sweeps", [(2, 2, 8, 600), (4, 3, 8, 1500), (6, 6, 6, 2500)])
def test_replay_trajectory_bit_exact(kd, n1, n2, nw, n_sweeps, variant):
    """replayed (r, bond) sequence: kappa, Z, acceptance counters bit-exact; W within 1e-10.
    variant 3 = walker resident in shared memory (k_resident, the default on these small lattices), 2 = delayed updates
    in Woodbury form (the default on large ones), 0 = immediate rank-1 update like the reference"""
    PBC, anti = ((False, False), (False, False)) if n1 == 2 else ((True, True), (True, False))
    lat, ham = U.problem(n1, n2, PBC, anti)
    ns = kd.ns(lat)
    rng = np.random.default_rng(2024 + n1)
    # (the reference's QR start state is numerically singular for spin-down on the 12-site lattice,
    #  so parity runs start from well-conditioned random Mott states, one per walker)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=200.0)
    nb = len(ham.nn)
    r = rng.random((n_sweeps, nw))
    bond = rng.integers(1, nb + 1, size=(n_sweeps, nw)).astype(np.int32)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_config(ku, kdn)
    eng.refresh()
    orc = U.oracle_walkers(ham, ku, kdn)
    done = 0
    for chunk in (1, 7, n_sweeps // 3, n_sweeps - 8 - n_sweeps // 3):
        eng.replay(r[done:done + chunk], bond[done:done + chunk])
        acc_o = 0
        for w, mc in enumerate(orc):
            for s in range(done, done + chunk):
                mc.sweep(replay=(r[s, w], int(bond[s, w]), 1))
                mc.sweeps = mc.sweeps + 1
        done += chunk
        gku, gkd = eng.get_config()
        z, zr = eng.Z()
        for w, mc in enumerate(orc):
            oku, okd = mc.kappa()
            assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd), f"kappa differs, walker {w} after {done} sweeps"
            assert z[w] == zr[w] == U.O.Z(ham.nn, oku, okd)
            Wu, Wd = mc.W()
            assert U.relerr(eng.get_W(w, 0), Wu) < TOL
            assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    acc, acc_w, _ = eng.accumulators(per_walker=True)
    for w, mc in enumerate(orc):
        c = mc.counters()
        assert acc_w[w] == c[0]
    assert acc[kd._lib.ACC_N_REACH] == sum(mc.counters()[1] for mc in orc)
    assert acc[kd._lib.ACC_N_REFRESH] == sum(mc.counters()[2] for mc in orc)
    assert eng.sweeps == n_sweeps
    eng.close()


@pytest.mark.parametrize("variant", [3, 2, 0])
def test_device_rng_matches_xoshiro_stream(kd, variant):
    """device-drawn random numbers follow Julia's Xoshiro256++ conventions: same
    trajectory and same final generator state as the oracle fed with the same initial states"""
    lat, ham = U.problem(4, 3)
    ns, nw, n = kd.ns(lat), 8, 3000
    ku0, kd0 = U.well_conditioned_mott(np.random.default_rng(8), ham, ns, ham.N_up, 1, cond_max=200.0)
    ku0, kd0 = ku0[0], kd0[0]
    states = kd.walker_states(1234, nw)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_config(ku0, kd0)
    eng.set_rng(states)
    eng.refresh()
    eng.sweep(n, thermalization=100)
    gku, gkd = eng.get_config()
    gst = eng.get_rng()
    acc, acc_w, ol_w = eng.accumulators(per_walker=True)
    tot = np.zeros(4)
    for w in range(nw):
        mc = U.oracle_walkers(ham, ku0[None], kd0[None])[0]
        g = U.O.Xoshiro(states[w])
        st, _ = mc.run(g, n, 100)
        oku, okd = mc.kappa()
        assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd)
        assert np.array_equal(gst[w], g.s)
        assert acc_w[w] == st[0]
        assert abs(ol_w[w] - st[1]) <= 1e-9 * max(1.0, abs(st[1]))
        tot += st
    assert acc[kd._lib.ACC_N_OL] == tot[3]
    assert abs(acc[kd._lib.ACC_SUM_OL] - tot[1]) <= 1e-9 * abs(tot[1])
    assert acc[kd._lib.ACC_WALKER_SWEEPS] == n * nw
    eng.close()


def test_measure_matches_oracle_getOL(kd):
    lat, ham = U.problem(6, 6)
    ns, nw = kd.ns(lat), 6
    rng = np.random.default_rng(3)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw)
    eng = kd.Engine(ham, nw)
    eng.set_config(ku, kdn)
    eng.refresh()
    ol = eng.measure()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        ref = mc.getOL()
        assert abs(ol[w] - ref) <= TOL * max(1.0, abs(ref))
    eng.close()


@pytest.mark.parametrize("variant", [2])
def test_measure_with_pending_factors(kd, variant):
    """O_L evaluated from W0 + pending delayed updates equals the oracle's O_L on the same trajectory"""
    lat, ham = U.problem(4, 3)
    ns, nw, n = kd.ns(lat), 8, 17                      # 17 sweeps: no refresh since sweep 0, factors pending
    rng = np.random.default_rng(77)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=200.0)
    r = rng.random((n, nw)) * 0.3                       # small r: many acceptances
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_config(ku, kdn)
    eng.refresh()
    eng.sweeps = 1
    eng.replay(r, bond)
    ol = eng.measure()
    orc = U.oracle_walkers(ham, ku, kdn)
    n_acc = 0
    for w, mc in enumerate(orc):
        mc.sweeps = 1
        for s in range(n):
            n_acc += mc.sweep(replay=(r[s, w], int(bond[s, w]), 1)) & 1
            mc.sweeps = mc.sweeps + 1
        ref = mc.getOL()
        assert abs(ol[w] - ref) <= TOL * max(1.0, abs(ref))
    assert n_acc > nw                                   # the test really exercised pending factors
    eng.close()


def test_error_paths(kd):
    lat, ham = U.problem(2, 2, (False, False), (False, False))
    eng = kd.Engine(ham, 2)
    with pytest.raises(kd.KdslError):
        eng.sweep(1)                                    # no configuration yet
    bad = np.zeros((2, 12), dtype=np.int64)
    with pytest.raises(kd.KdslError):
        eng.set_config(bad, bad)                        # not a Mott state
    ku0, kd0 = kd.init_conf_qr(ham, 12, 6)
    eng.set_config(ku0, kd0)
    with pytest.raises(kd.KdslError):
        eng.sweep(1)                                    # W stale
    eng.refresh()
    eng.sweep(10)
    for name, value in (("inverse_variant", 9), ("flush_variant", 4), ("flush_variant", 7), ("inverse_variant", 12), ("no_such_option", 1)):
        with pytest.raises(kd.KdslError):
            eng.set_option(name, value)                 # ComplexF64-only variants / unknown names are rejected on the real engine
    eng.close()
    # rank-deficient orbitals -> SingularException (reference: test-MonteCarlo.jl:429-477)
    Ud = np.zeros((12, 6)); Ud[:6] = np.eye(6)
    ham2 = kd.Hamiltonian(6, 6, Ud, Ud, ham.H_mat, ham.nn)
    eng2 = kd.Engine(ham2, 1)
    ku = np.zeros(12, dtype=np.int64); ku[6:] = np.arange(1, 7)
    kdn = np.zeros(12, dtype=np.int64); kdn[:6] = np.arange(1, 7)
    eng2.set_config(ku, kdn)
    with pytest.raises(kd.SingularException):
        eng2.refresh()
    assert eng2.flags()[0] & 1
    eng2.close()


def test_fused_refresh_singular_species_is_isolated(kd):
    """k_reeval_fused refreshes the two species independently: a singular tilde_U of one species flags the walker
    (SingularException, src/MonteCarlo.jl:596-603) and leaves that species' W alone, the other species is re-evaluated"""
    lat, ham = U.problem(2, 2, (False, False), (False, False))
    Ud = np.zeros((12, 6)); Ud[:6] = np.eye(6)                  # down orbitals live on sites 1..6 only
    ham2 = kd.Hamiltonian(6, 6, ham.U_up, Ud, ham.H_mat, ham.nn)
    rng = np.random.default_rng(5)
    ku, kdn = U.well_conditioned_mott(rng, ham, 12, 6, 2)
    # walker 0: down particles on sites 7..12 -> tilde_U_down = 0 (singular); walker 1: on sites 1..6 -> identity
    for w, dn_sites in enumerate((np.arange(6, 12), np.arange(0, 6))):
        ku[w] = 0; kdn[w] = 0
        kdn[w, dn_sites] = np.arange(1, 7)
        ku[w, np.setdiff1d(np.arange(12), dn_sites)] = np.arange(1, 7)
    eng = kd.Engine(ham2, 2)
    eng.set_config(ku, kdn)
    with pytest.raises(kd.SingularException):
        eng.refresh()
    fl = eng.flags()
    assert fl[0] & 1 and not fl[1] & 1
    for w in range(2):                                           # the up species is fine for both walkers
        cond = np.linalg.cond(kd.tilde_U(ham.U_up, ku[w]))
        if cond < 1e8:
            mc = U.oracle_walkers(ham2, ku[w:w + 1], kdn[w:w + 1], refresh=False)[0]
            Wu_ref = ham.U_up @ np.linalg.inv(kd.tilde_U(ham.U_up, ku[w]))
            assert U.relerr(eng.get_W(w, 0), Wu_ref) < 1e-8 * max(1.0, cond)
    Wd1 = eng.get_W(1, 1)                                        # walker 1, down: W = U_dn inv(I) = U_dn
    assert U.relerr(Wd1, Ud) < TOL
    assert np.all(eng.get_W(0, 1) == 0.0)                        # walker 0, down: never written
    eng.close()


def test_exact_energy_12_sites(kd):
    """<E>/site of the chain's stationary law |psi|^2 / Z_mu on the reference's 2x2 OBC test lattice:
    exact enumeration gives -0.3714938624 (tests/golden/exact_energies.json)"""
    lat, ham = U.problem(2, 2, (False, False), (False, False))
    nw, therm, n = 2048, 2000, 12000
    mc = kd.MC({"n1": 2, "n2": 2, "PBC": (False, False), "N_up": 6, "N_down": 6, "n_walkers": nw})
    ctx = kd.MCContext({"thermalization": therm, "seed": 99, "binsize": 1})
    kd.init_(mc, ctx, {"n1": 2, "n2": 2, "N_up": 6})
    kd.run_(mc, ctx, therm + n)
    out, acc_w, ol_w = mc.engine.accumulators(per_walker=True)
    _, n_w = mc.engine.last_OL()
    per_walker = ol_w / n_w / 12.0
    mean = per_walker.mean()
    err = per_walker.std(ddof=1) / np.sqrt(nw)
    assert abs(mean - (-0.3714938624)) < 5 * err + 1e-5, (mean, err)
    assert err < 5e-4


def test_reweighted_energy_is_the_psi2_average_12_sites(kd):
    """The chain samples |psi|^2 / Z_mu; the Z_mu-reweighted estimator <O_L Z_mu> / <Z_mu> (kdsl_set_observables) must give
    the |psi|^2 average instead: exact enumeration -0.3720882491 vs -0.3714938624 for the chain's own law (derived.json)."""
    nw, therm, n = 2048, 2000, 12000
    mc = kd.MC({"n1": 2, "n2": 2, "PBC": (False, False), "N_up": 6, "N_down": 6, "n_walkers": nw})
    ctx = kd.MCContext({"thermalization": therm, "seed": 7, "binsize": 1})
    kd.init_(mc, ctx, {"n1": 2, "n2": 2, "N_up": 6})
    mc.engine.set_observables()
    kd.run_(mc, ctx, therm + n)
    obs = mc.engine.observables()
    out, _, ol_w = mc.engine.accumulators(per_walker=True)
    _, n_w = mc.engine.last_OL()
    err = (ol_w / n_w / 12.0).std(ddof=1) / np.sqrt(nw)
    assert obs["n"] == out[kd._lib.ACC_N_OL]
    assert abs(obs["energy_psi2"] - (-0.3720882491)) < 6 * err + 2e-5, (obs["energy_psi2"], err)
    assert abs(obs["energy_psi2"] - (-0.3714938624)) > 3 * err               # ... and it is NOT the chain-law value


@pytest.mark.parametrize("variant", [3, 2, 0])
def test_extra_observables_match_oracle(kd, variant):
    """S(q) and the Z_mu-weighted sums taken at the :OL cadence equal the oracle-side evaluation (numpy structure factor,
    oracle Z and getOL) on the same replayed chain"""
    lat, ham = U.problem(4, 3)
    ns, nw, n = kd.ns(lat), 6, 400
    rng = np.random.default_rng(17)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=200.0)
    r = rng.random((n, nw)) * 0.6
    bond = rng.integers(1, len(ham.nn) + 1, size=(n, nw)).astype(np.int32)
    coords = np.array([kd.get_site_coord(lat, s) for s in range(1, ns + 1)], dtype=np.float64)
    qs = np.array([[0.0, 0.0], [np.pi, 0.0], [2 * np.pi / 3, 2 * np.pi / np.sqrt(3.0)], [0.3, -1.1]])
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_observables(qs, coords)
    eng.set_config(ku, kdn)
    eng.refresh()
    eng.replay(r, bond, thermalization=0)
    got = eng.observables()
    cos_t, sin_t = np.cos(qs @ coords.T), np.sin(qs @ coords.T)
    ref = np.zeros(4 + 2 * len(qs))
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn)):
        for s in range(n):
            mc.sweep(replay=(r[s, w], int(bond[s, w]), 1))
            mc.sweeps = mc.sweeps + 1
            if mc.sweeps % (ns // 2) == 0:                                    # Carlo.measure! cadence (:630)
                oku, okd = mc.kappa()
                z, ol, sq = U.O.Z(ham.nn, oku, okd), mc.getOL(), U.O.structure_factor(oku, cos_t, sin_t)
                ref[0] += 1; ref[1] += z; ref[2] += ol * z
                ref[4:4 + len(qs)] += sq; ref[4 + len(qs):] += sq * z
    assert got["n"] == ref[0] == nw * (n // (ns // 2))
    assert got["sum_Z"] == ref[1]
    assert abs(got["sum_OL_Z"] - ref[2]) <= 1e-10 * abs(ref[2])
    assert np.allclose(got["sum_Sq"], ref[4:4 + len(qs)], rtol=1e-12, atol=1e-12)
    assert np.allclose(got["sum_Sq_Z"], ref[4 + len(qs):], rtol=1e-12, atol=1e-10)
    assert abs(got["Sq_chain"][0]) < 1e-12                                    # S(q = 0) = (sum Sz)^2 / ns = 0 at half filling
    eng.close()


def _run_chain(kd, ham, ku, kdn, states, n_sweeps, options):
    nw = ku.shape[0]
    eng = kd.Engine(ham, nw)
    for k, v in options.items():
        eng.set_option(k, v)
    eng.set_config(ku, kdn)
    eng.set_rng(states)
    eng.refresh()
    eng.sweep(n_sweeps, thermalization=0)
    out = (eng.get_config(), eng.accumulators(per_walker=True), eng.get_rng().copy(),
           [eng.get_W(w, s) for w in range(nw) for s in (0, 1)], eng.Z())
    eng.close()
    return out


@pytest.mark.parametrize("options", [
    {"update_variant": 2},
    {"update_variant": 2, "fuse_sweeps": 0},
    {"update_variant": 2, "flush_every": 4},
    {"update_variant": 2, "flush_every": 3, "flush_threshold": 5},
    {"update_variant": 2, "flush_every": 16, "flush_threshold": 16},
    {"update_variant": 2, "flush_every": 8, "flush_threshold": 2},   # threshold below the cadence: a walker is still listed once
    {"update_variant": 0},
])
def test_launch_grouping_and_flush_cadence_do_not_change_the_chain(kd, options):
    """The update algorithm (walker resident in shared memory = the default at 108 sites, Woodbury delayed updates,
    immediate rank-1), fusing proposals into one launch and the flush cadence are scheduling choices: configurations,
    counters, RNG states and O_L sums must be identical to the default path; W agrees to rounding."""
    lat, ham = U.problem(6, 6)
    ns, nw, n_sweeps = kd.ns(lat), 24, 700
    rng = np.random.default_rng(77)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=500.0)
    states = kd.walker_states(99, nw)
    ref = _run_chain(kd, ham, ku, kdn, states, n_sweeps, {})
    got = _run_chain(kd, ham, ku, kdn, states, n_sweeps, options)
    assert np.array_equal(ref[0][0], got[0][0]) and np.array_equal(ref[0][1], got[0][1])
    assert np.array_equal(ref[2], got[2])                                   # Xoshiro states
    assert np.array_equal(ref[1][1], got[1][1])                             # accepted moves per walker
    assert np.array_equal(ref[4][0], got[4][0]) and np.array_equal(got[4][0], got[4][1])   # Z_mu, incremental == recount
    assert ref[1][0][kd._lib.ACC_N_REFRESH] == got[1][0][kd._lib.ACC_N_REFRESH]
    assert np.allclose(ref[1][2], got[1][2], rtol=1e-11, atol=1e-11)         # O_L sums per walker
    for a, b in zip(ref[3], got[3]):
        assert U.relerr(a, b) < 1e-11
    if options == {"update_variant": 2, "fuse_sweeps": 0}:                   # same arithmetic, different launches: bit-identical
        ref2 = _run_chain(kd, ham, ku, kdn, states, n_sweeps, {"update_variant": 2})
        for a, b in zip(ref2[3], got[3]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("n,nw,n_sweeps", [(8, 24, 400), (10, 16, 400), (12, 48, 500), (18, 6, 1000)])
def test_full_size_invariants(kd, n, nw, n_sweeps):
    """BASELINE configs 3 and 4 (432 / 972 sites): size-independent properties instead of an oracle replay.
    W tilde_U = U (the definition of W), unit rows at occupied sites, incremental Z_mu == recount, and after a chain
    with delayed updates the maintained W equals a from-scratch re-evaluation of the final configuration."""
    lat, ham = U.problem(n, n)
    ns = kd.ns(lat)
    ku0, kd0 = kd.init_conf_qr(ham, ns, ns // 2)
    ku = np.tile(ku0, (nw, 1)); kdn = np.tile(kd0, (nw, 1))
    eng = kd.Engine(ham, nw)
    eng.set_config(ku, kdn)
    eng.set_rng(kd.walker_states(5, nw))
    eng.refresh()
    eng.sweep(n_sweeps, thermalization=0)
    z, zr = eng.Z()
    assert np.array_equal(z, zr)
    gku, gkd = eng.get_config()
    assert not np.array_equal(gku[0], gku[1])                               # the walkers decorrelated
    W_run = [(eng.get_W(w, 0), eng.get_W(w, 1)) for w in (0, nw - 1)]
    ol_run = eng.measure()
    eng.refresh()
    ol_ref = eng.measure()
    assert np.allclose(ol_run, ol_ref, rtol=1e-9, atol=1e-9)
    for (Wu, Wd), w in zip(W_run, (0, nw - 1)):
        for spin, (W, U_, kap) in enumerate(((Wu, ham.U_up, gku[w]), (Wd, ham.U_down, gkd[w]))):
            Wf = eng.get_W(w, spin)
            assert U.relerr(W, Wf) < 1e-9                                   # rank-1 history vs from scratch
            Ut = kd.tilde_U(U_, kap)
            assert np.max(np.abs(Wf @ Ut - U_)) < 1e-10 * max(1.0, np.max(np.abs(Wf)))
            occ = np.nonzero(kap)[0]
            E = np.zeros_like(Wf[occ]); E[np.arange(len(occ)), kap[occ] - 1] = 1.0
            assert np.array_equal(Wf[occ], E)                               # exact unit rows
    assert eng.accumulators()[kd._lib.ACC_N_SINGULAR] == 0
    eng.close()


# ---- ComplexF64 mode: Peierls flux B != 0, complex Hermitian hopping matrix ----------------

def _complex_problem(kd, n1, n2, B, N_up=None):
    lat, ham = U.problem(n1, n2, (True, True), (True, False), "pi", N_up, B)
    assert np.iscomplexobj(ham.U_up) and np.abs(np.asarray(ham.U_up).imag).max() > 1e-3
    return lat, ham


@pytest.mark.parametrize("n1,n2,B", [(4, 3, 0.37), (6, 6, 0.11)])
def test_complex_refresh_update_measure_match_oracle(kd, n1, n2, B):
    lat, ham = _complex_problem(kd, n1, n2, B)
    ns, nw = kd.ns(lat), 5
    rng = np.random.default_rng(31)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e4)
    eng = kd.Engine(ham, nw)
    assert eng.is_complex
    eng.set_config(ku, kdn)
    orc = U.oracle_walkers(ham, ku, kdn, dtype="c128")
    # inverse_variant 0 (default) / 9: complex blocked Gauss-Jordan by thread-block clusters on split (re, im) planes;
    # 5 / 7: inverse through the real 2N x 2N embedding on the blocked DMMA kernels (one CTA / one cluster per matrix);
    # 1: the reference-style unblocked complex Gauss-Jordan
    for variant in (1, 5, 7, 9, 0):
        eng.set_option("inverse_variant", variant)
        eng.set_W(0, 0, np.zeros_like(np.asarray(orc[0].W()[0])))           # make sure the refresh really rewrites W
        eng.refresh()
        ol = eng.measure()
        for w, mc in enumerate(orc):
            Wu, Wd = mc.W()
            assert np.abs(np.asarray(Wu).imag).max() > 1e-6                 # genuinely complex
            assert U.relerr(eng.get_W(w, 0), Wu) < TOL
            assert U.relerr(eng.get_W(w, 1), Wd) < TOL
            assert abs(ol[w] - mc.getOL()) < TOL * max(1.0, abs(mc.getOL()))
    # explicit rank-1 moves (update_W!, src/MonteCarlo.jl:279-292) against the oracle's complex update
    mv_w, lu, Ku, ld, Kd = [], [], [], [], []
    for w in range(nw):
        up_sites = np.nonzero(ku[w])[0]; dn_sites = np.nonzero(kdn[w])[0]
        i, s = int(up_sites[w % len(up_sites)]), int(dn_sites[(3 * w + 1) % len(dn_sites)])
        mv_w.append(w); lu.append(int(ku[w][i])); Ku.append(s + 1); ld.append(int(kdn[w][s])); Kd.append(i + 1)
    eng.update_W(mv_w, lu, Ku, ld, Kd)
    for w, mc in enumerate(orc):
        Wu, Wd = mc.W()
        Wu2 = U.O.update_W(np.array(Wu), lu[w], Ku[w], "c128")
        Wd2 = U.O.update_W(np.array(Wd), ld[w], Kd[w], "c128")
        assert U.relerr(eng.get_W(w, 0), Wu2) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd2) < TOL
    with pytest.raises(kd.KdslError):
        eng.set_option("flush_variant", 3)                                  # the bulk-async flush is real-only
    eng.close()


@pytest.mark.parametrize("n,N_up,B,cluster,row_slices,nw", [
    (6, None, 0.11, 4, 8, 80), (6, 50, 0.11, 2, 1, 20), (8, 97, 0.05, 4, 8, 40), (8, None, 0.05, 5, 3, 40), (8, None, 0.05, 8, 2, 20)])
def test_complex_cluster_inverse_matches_oracle(kd, n, N_up, B, cluster, row_slices, nw):
    """k_inverse_cl_c (ComplexF64 engine: complex blocked Gauss-Jordan, one matrix per thread-block cluster, split re / im
    planes, four real DMMAs per complex block product): N not a multiple of 8, N_up != N_down (scripts/LL.jl filling),
    more (walker, species) items than resident clusters, all cluster sizes' work splits, scratch reuse"""
    lat, ham = _complex_problem(kd, n, n, B, N_up=N_up)
    ns = kd.ns(lat)
    rng = np.random.default_rng(500 + n + cluster)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    eng = kd.Engine(ham, nw)
    assert eng.is_complex
    eng.set_option("inverse_variant", 9)
    eng.set_option("inverse_cluster", cluster)
    eng.set_option("inverse_row_slices", row_slices)
    eng.set_config(ku, kdn)
    for rep in range(2):
        eng.set_W(0, 0, np.zeros((ns, ham.N_up), dtype=np.complex128))
        eng.refresh()
    for w, mc in enumerate(U.oracle_walkers(ham, ku, kdn, dtype="c128")):
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL
        assert U.relerr(eng.get_W(w, 1), Wd) < TOL
    assert eng.accumulators()[kd._lib.ACC_N_SINGULAR] == 0
    eng.close()


def test_complex_cluster_inverse_singular_matrix_is_flagged(kd):
    lat, ham = _complex_problem(kd, 6, 6, 0.11)
    ns, nw = kd.ns(lat), 45
    rng = np.random.default_rng(78)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=1e5)
    Ud = np.array(ham.U_down, dtype=np.complex128).copy()
    bad_sites = np.nonzero(kdn[7])[0]
    Ud[bad_sites[0], :] = 0.0
    ham2 = kd.Hamiltonian(ham.N_up, ham.N_down, ham.U_up, Ud, ham.H_mat, ham.nn)
    expect_bad = np.array([kdn[w, bad_sites[0]] != 0 for w in range(nw)])
    assert expect_bad[7] and not expect_bad.all()
    eng = kd.Engine(ham2, nw)
    assert eng.is_complex
    eng.set_config(ku, kdn)
    with pytest.raises(kd.SingularException):
        eng.refresh()
    fl = eng.flags()
    assert np.array_equal((fl & 1) != 0, expect_bad)
    for w in np.nonzero(~expect_bad)[0][:6]:
        Wd_ref = Ud @ np.linalg.inv(kd.tilde_U(Ud, kdn[w]))
        assert U.relerr(eng.get_W(int(w), 1), Wd_ref) < 1e-8
    eng.close()


@pytest.mark.parametrize("variant,flush", [(2, 0), (2, 7), (2, 4), (0, 0)])
def test_complex_replay_and_device_rng_match_oracle(kd, variant, flush):
    """ComplexF64 chain: replayed proposals -> kappa / Z_mu / counters bit-exact, W within 1e-10; device Xoshiro ->
    same trajectory, counters and O_L sums as the oracle's Carlo loop.  variant 2 = Woodbury delayed updates (default;
    flush 0 = tensor-pipe flush k_flush_dmma2_c with two CTAs per SM, 7 = k_flush_dmma_c with one, 4 = FMA flush k_flush_c), 0 = the reference's immediate rank-1 update."""
    lat, ham = _complex_problem(kd, 4, 3, 0.37)
    ns, nw, n_sweeps = kd.ns(lat), 6, 900
    rng = np.random.default_rng(8)
    ku, kdn = U.well_conditioned_mott(rng, ham, ns, ham.N_up, nw, cond_max=300.0)
    r = rng.random((n_sweeps, nw))
    r[:, 3:] *= 0.45                                                        # more acceptances: pending updates, flushes
    bond = rng.integers(1, len(ham.nn) + 1, size=(n_sweeps, nw)).astype(np.int32)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_option("flush_variant", flush)
    eng.set_config(ku, kdn)
    eng.refresh()
    eng.replay(r, bond)
    ol = eng.measure()                                                      # (Woodbury form: pending updates included)
    orc = U.oracle_walkers(ham, ku, kdn, dtype="c128")
    gku, gkd = eng.get_config()
    z, zr = eng.Z()
    for w, mc in enumerate(orc):
        for s in range(n_sweeps):
            mc.sweep(replay=(r[s, w], int(bond[s, w]), 1))
            mc.sweeps = mc.sweeps + 1
        oku, okd = mc.kappa()
        assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd)
        assert abs(ol[w] - mc.getOL()) <= TOL * max(1.0, abs(mc.getOL()))
        assert z[w] == zr[w] == U.O.Z(ham.nn, oku, okd)
        Wu, Wd = mc.W()
        assert U.relerr(eng.get_W(w, 0), Wu) < TOL and U.relerr(eng.get_W(w, 1), Wd) < TOL
    acc, acc_w, _ = eng.accumulators(per_walker=True)
    assert [int(a) for a in acc_w] == [mc.counters()[0] for mc in orc]
    assert acc[kd._lib.ACC_N_REFRESH] == sum(mc.counters()[2] for mc in orc)
    eng.close()
    # device RNG
    states = kd.walker_states(3, nw)
    eng = kd.Engine(ham, nw)
    eng.set_option("update_variant", variant)
    eng.set_option("flush_variant", flush)
    eng.set_config(ku, kdn)
    eng.set_rng(states)
    eng.refresh()
    eng.sweep(n_sweeps, thermalization=ns // 2)
    gku, gkd = eng.get_config()
    acc, acc_w, ol_w = eng.accumulators(per_walker=True)
    for w in range(nw):
        mc = U.O.MC(np.asarray(ham.nn, dtype=np.int32), ham.U_up, ham.U_down, "c128")
        mc.set_kappa(ku[w], kdn[w]); mc.reevaluateW()
        st, _ = mc.run(U.O.Xoshiro(states[w]), n_sweeps, ns // 2)
        oku, okd = mc.kappa()
        assert np.array_equal(gku[w], oku) and np.array_equal(gkd[w], okd)
        assert acc_w[w] == st[0]
        assert abs(ol_w[w] - st[1]) <= 1e-9 * max(1.0, abs(st[1]))
    eng.close()
