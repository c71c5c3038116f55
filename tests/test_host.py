"""Host-side logic of the product (lattice / Hamiltonian / bond tables / MC plumbing) against the
reference's known answers and against the oracle.  CPU only: nothing here touches the GPU."""
import numpy as np
import pytest

import kagomedsl.jl_b200 as kd
from oracle import oracle as O


def test_double_kagome_constructor():                    # reference test-Lattice.jl:1-8, src/Lattice.jl:68-97
    with pytest.raises(AssertionError):
        kd.DoubleKagome(1.0, 3, 3, (False, False))
    DK = kd.DoubleKagome(1.0, 4, 3, (False, False))
    assert kd.ns(DK) == 36
    assert np.allclose(DK.a1, [4.0, 0.0]) and np.allclose(DK.a2, [1.0, np.sqrt(3.0)])
    assert np.allclose(DK.r[5], [2.5, 0.5 * np.sqrt(3.0)])
    with pytest.raises(ValueError, match="Cannot have antiperiodic boundary conditions without periodic"):
        kd.DoubleKagome(1.0, 4, 3, (False, False), (True, False))


@pytest.mark.parametrize("n1,n2,PBC,anti,B,flux", [
    (4, 3, (False, False), (False, False), 0.0, "pi"), (4, 3, (True, False), (True, False), 0.0, "pi"),
    (4, 3, (True, True), (True, True), 0.1, "pi"), (2, 2, (True, True), (False, True), 0.0, "zero"),
    (6, 6, (True, True), (True, False), 0.0, "pi"), (12, 12, (True, True), (True, False), 0.0, "pi"),
])
def test_hmat_and_bonds_equal_oracle(n1, n2, PBC, anti, B, flux):
    li, lx = (None, None) if flux == "pi" else (kd.zero_link_in, kd.zero_link_inter)
    lat = kd.DoubleKagome(1.0, n1, n2, PBC, anti)
    Hp = kd.Hmat(lat, link_in=li, link_inter=lx, B=B)
    Ho = O.Lattice(1.0, n1, n2, PBC, anti).hmat(li, lx, B)
    assert np.array_equal(Hp, Ho) or np.allclose(Hp, Ho, rtol=0, atol=1e-15)
    assert kd.get_nn(Hp) == [tuple(b) for b in O.get_nn(Ho).tolist()]          # bit-exact bond table


def test_reference_hamiltonian_kats():                   # test-Hamiltonian.jl:4-61, 84-95
    H = kd.Hmat(kd.DoubleKagome(1.0, 4, 3, (False, False)))
    assert H[0, 1] == -1 and H[0, 2] == -1 and np.isclose(H[2, 12], 1) and np.allclose(H, H.conj().T)
    HB = kd.Hmat(kd.DoubleKagome(1.0, 4, 3, (True, True), (True, True)))
    assert HB[0, 10] == 1 and HB[0, 26] == -1 and HB[10, 26] == 1
    ham = kd.Hamiltonian(18, 18, kd.DoubleKagome(1.0, 4, 3, (True, False)))
    assert ham.U_up.shape == (36, 18) and ham.U_down.shape == (36, 18)
    assert np.allclose(ham.U_up.T @ ham.U_up, np.eye(18), atol=1e-12)


def test_geometry_helpers():                             # test-Hamiltonian.jl:467-685
    lat = kd.DoubleKagome(1.0, 4, 3, (True, True))
    assert kd.unitcell_diff(lat, lat.a1 + lat.a2, [0.0, 0.0]) == (1, 1)
    assert np.array_equal(kd.get_site_coord(lat, 13), [1.0, np.sqrt(3.0)])
    assert (1, 0, 1.0) in kd.get_boundary_shifts(lat, 3, 9)
    sh = kd.get_boundary_shifts(lat, 3, 9)
    assert (-1, 0, 1.0) in sh and (3, 0, 1.0) in sh
    assert sorted(sh) == sorted(O.Lattice(1.0, 4, 3, (True, True)).get_boundary_shifts(3, 9))
    with pytest.raises(AssertionError):
        kd.get_boundary_shifts(lat, 3, 3)
    T = np.zeros((36, 36), dtype=np.complex128)
    kd.apply_boundary_conditions_(T, kd.DoubleKagome(1.0, 4, 3, (True, True), (True, False)), 1, 11, {(1, 5, -1, 0): 1.0}, 0.0)
    assert np.isclose(T[0, 10], -1.0)
    with pytest.raises(AssertionError):
        kd.apply_boundary_conditions_(T, lat, 1, 2, {(1, 5, -1, 0): 1.0}, 0.0)


def test_Sz_spinInteraction_getxprime():                 # test-Hamiltonian.jl:97-135, 148-331
    assert kd.Sz(1, [1, 0, 2], [0, 2, 0]) == 0.5 and kd.Sz(2, [1, 0, 2], [0, 2, 0]) == -0.5
    with pytest.raises(ValueError, match="doubly occupied"):
        kd.Sz(2, [1, 2, 0], [0, 2, 1])
    with pytest.raises(IndexError):
        kd.Sz(4, [1, 2, 0], [0, 2, 1])
    with pytest.raises(ValueError, match="DimensionMismatch"):
        kd.Sz(1, [1, 2, 0, 1], [0, 2, 1])
    with pytest.raises(ValueError, match="unoccupied"):
        kd.Sz(1, [0], [0])
    ku, kdn = [1, 0, 2], [0, 1, 0]
    xp = {}
    kd.spinInteraction_(xp, ku, kdn, 2, 1)
    assert xp == {(2, 1, 1, 1): -0.5}
    xp = {}
    kd.spinInteraction_(xp, ku, kdn, 1, 2)
    assert xp == {(2, 1, 1, 1): -0.5}
    xp = {}
    kd.spinInteraction_(xp, ku, kdn, 1, 3)
    assert xp == {}
    kd.spinInteraction_(xp, ku, kdn, 2, 3)
    assert xp[(2, 2, 3, 1)] == -0.5
    xp = {(2, 1, 1, 1): 0.25}
    kd.spinInteraction_(xp, ku, kdn, 1, 2)
    assert np.isclose(xp[(2, 1, 1, 1)], -0.25)
    xp = {}
    kd.spinInteraction_(xp, [1], [0], 1, 1)
    assert xp == {}
    with pytest.raises(IndexError):
        kd.spinInteraction_({}, [], [], 1, 1)
    ham = kd.Hamiltonian(1, 0, kd.DoubleKagome(1.0, 4, 3, (False, False)))
    xp = kd.getxprime(ham, [1] + [0] * 35, [0] + list(range(1, 36)))
    assert len(xp) == 3 and xp[(-1, -1, -1, -1)] == (len(ham.nn) - 2) * 0.25 - 0.5
    assert xp[(2, 1, 1, 1)] == -0.5 and xp[(3, 1, 1, 2)] == -0.5
    assert xp == O.getxprime(np.asarray(ham.nn), [1] + [0] * 35, [0] + list(range(1, 36)))


def test_tilde_U_Z_is_occupied():                        # test-MonteCarlo.jl:54-183, 296-303
    Um = np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]])
    r = kd.tilde_U(Um, [2, 3, 1])
    assert np.array_equal(r[0], Um[2]) and np.array_equal(r[1], Um[0]) and np.array_equal(r[2], Um[1])
    with pytest.raises(ValueError, match="not valid"):
        kd.tilde_U(np.array([[1.0, 2], [3, 4]]), [0, 0])
    with pytest.raises(IndexError):
        kd.tilde_U(np.array([[1.0, 2], [3, 4]]), [3, 1])
    with pytest.raises(ValueError, match="DimensionMismatch"):
        kd.tilde_U(np.array([[1.0, 2], [3, 4]]), [1, 2, 3])
    assert kd.tilde_U(np.array([[1, 2], [3, 4]]), [1, 2]).dtype.kind == "i"
    assert kd.tilde_U(np.zeros((0, 0)), []).shape == (0, 0)
    assert kd.Z([(1, 2), (2, 3), (1, 3)], [0, 1, 0], [1, 0, 2]) == 2
    k = [1, 0, 2, 0]
    assert kd.is_occupied(k, 1) and not kd.is_occupied(k, 2) and kd.is_occupied(k, 3) and not kd.is_occupied(k, 4)
    with pytest.raises(IndexError):
        kd.is_occupied(k, 5)


def test_MC_constructs_without_gpu_and_init_conf_qr():   # test-MonteCarlo.jl:8-52
    mc = kd.MC({"n1": 4, "n2": 3, "PBC": (True, False), "N_up": 18, "N_down": 18})
    assert mc.W_up.shape == (36, 18) and not mc.W_up.any() and not mc.kappa_up.any()
    rng = np.random.default_rng(0)
    ham = kd.Hamiltonian(6, 6, rng.random((12, 6)), rng.random((12, 6)), np.zeros((12, 12)), [])
    ku, kdn = kd.init_conf_qr(ham, 12, 6)
    assert np.count_nonzero(ku) == 6 and np.count_nonzero(kdn) == 6
    assert sorted(ku[ku != 0]) == list(range(1, 7)) and sorted(kdn[kdn != 0]) == list(range(1, 7))
    assert np.all((ku != 0) ^ (kdn != 0))
    assert abs(np.linalg.det(kd.tilde_U(ham.U_up, ku))) > np.finfo(float).eps
    assert abs(np.linalg.det(kd.tilde_U(ham.U_down, kdn))) > np.finfo(float).eps
    # the benchmark lattices start from a well-conditioned state for both species
    lat = kd.DoubleKagome(1.0, 6, 6, (True, True), (True, False))
    h2 = kd.Hamiltonian(54, 54, lat)
    ku, kdn = kd.init_conf_qr(h2, 108, 54)
    assert np.linalg.cond(kd.tilde_U(h2.U_up, ku)) < 50 and np.linalg.cond(kd.tilde_U(h2.U_down, kdn)) < 50
    assert h2.gap() > 0.8


def test_checkpoint_roundtrip_and_evaluables():          # test-MonteCarlo.jl:502-524, src/MonteCarlo.jl:675-687
    params = {"n1": 2, "n2": 2, "PBC": (False, False), "N_up": 6, "N_down": 6}
    mc = kd.MC(params)
    mc._kappa_up = np.array([1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 0])
    mc._kappa_down = np.array([0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6])
    group = {}
    kd.write_checkpoint(mc, group)
    assert group["kappa_up"].dtype == np.int64 and set(group) == {"kappa_up", "kappa_down"}
    mc2 = kd.MC(params)
    kd.read_checkpoint_(mc2, group, defer=True)
    assert np.array_equal(mc2.kappa_up, mc._kappa_up) and np.array_equal(mc2.kappa_down, mc._kappa_down)
    # the same group as a file (.npz with the reference's dataset names / dtypes, plus ctx.sweeps)
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "run0001.npz")
        c1 = kd.MCContext({"seed": 1}); c1.sweeps = 4321
        kd.save_checkpoint_npz(mc, path, c1)
        with np.load(path) as f:
            assert f["kappa_up"].dtype == np.int64 and f["kappa_down"].dtype == np.int64 and int(f["sweeps"]) == 4321
        mc3, c3 = kd.MC(params), kd.MCContext({"seed": 1})
        kd.load_checkpoint_npz(mc3, path, c3, defer=True)
        assert c3.sweeps == 4321 and np.array_equal(mc3.kappa_down, mc._kappa_down)
    ev = kd.Evaluator()
    kd.register_evaluables(kd.MC, ev, params)
    ctx = kd.MCContext({"binsize": 3, "seed": 123, "thermalization": 10})
    kd.measure_(ctx, "OL", -4.8)
    kd.measure_(ctx, "OL", -5.2)
    assert np.isclose(ev.results(ctx)["energy"], -5.0 / 12)
    assert ctx.thermalization_sweeps == 10 and ctx.sweeps == 0 and not ctx.is_thermalized()


def test_host_xoshiro_equals_oracle_stream():
    a = kd.Xoshiro(seed=42)
    b = O.Xoshiro.from_seed(42)
    assert [a.next_u64() for _ in range(20)] == [b.next_u64() for _ in range(20)]
    assert [a.rand() for _ in range(20)] == [b.rand() for _ in range(20)]
    assert [a.rand_index(864) for _ in range(50)] == [b.rand_index(864) for _ in range(50)]
    assert a.rand_index(1) == 1 == b.rand_index(1)
    assert np.array_equal(kd.walker_states(1234, 5), O.seed_states(1234, 5))
    assert np.array_equal(kd.walker_states(1234, 3, first_walker=2), O.seed_states(1234, 5)[2:])


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point fails loudly (KDSL_ERR_CUDA)"""
    from kagomedsl.jl_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    ham = kd.Hamiltonian(6, 6, kd.DoubleKagome(1.0, 2, 2, (False, False)))
    with pytest.raises(kd.KdslError) as e:
        kd.Engine(ham, 4)
    assert e.value.code == _lib.KDSL_ERR_CUDA
    mc = kd.MC({"n1": 2, "n2": 2, "PBC": (False, False), "N_up": 6, "N_down": 6})
    with pytest.raises(kd.KdslError):
        kd.init_(mc, kd.MCContext({"seed": 1}), {"n1": 2, "n2": 2, "N_up": 6})


def test_bench_clock_sampler_window():
    """bench.py reports only the nvidia-smi samples that fall inside the timed region (and says so when it had to widen)"""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    mk = lambda t, mhz, pcap="Not Active": (t, f"{mhz}, 1965, 600.0, Not Active, Not Active, Not Active, {pcap}\n")
    lines = [mk(0.0, 210), mk(1.0, 1900), mk(2.0, 1965), mk(2.05, 1950, "Active"), mk(2.1, 1965), mk(2.15, 1965), mk(3.0, 300)]
    c = bench.ClockSampler.summarise(lines, 2.0, 2.15, 1.0)
    assert c["samples"] == 4 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["window"] == "timed region"
    c = bench.ClockSampler.summarise(lines, 2.12, 2.15, 1.0)          # too short: widened over the warm-up
    assert c["samples"] == 5 and c["window"].startswith("warm-up")
    assert bench.ClockSampler.summarise([], 0.0, 1.0, 0.0)["sm_mhz"] is None
