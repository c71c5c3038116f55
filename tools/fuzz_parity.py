"""Randomised parity fuzz over random problems (tests/_fuzz.py): usage  fuzz_parity.py SEED SECONDS [big]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _fuzz
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
big = len(sys.argv) > 3 and sys.argv[3] == "big"          # lattices up to 12 x 12 (432 sites)
ok, skip = _fuzz.run_fuzz(seed, budget, big)
print(f"fuzz seed {seed}: {ok} cases bit-identical to the oracle chains, {skip} skipped (ill-conditioned start / unsupported), {budget:.0f} s")
