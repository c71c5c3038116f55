"""GPU parity fuzz: a few seconds of random problems per run (tests/_fuzz.py; tools/fuzz_parity.py runs it for longer)."""
import pytest

pytestmark = pytest.mark.gpu


def test_random_problems_match_the_oracle_chains():
    import kagomedsl.jl_b200 as kd
    if kd._lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box")
    import _fuzz
    n_ok, n_skip = _fuzz.run_fuzz(20261018, 8.0)
    assert n_ok >= 20, (n_ok, n_skip)
