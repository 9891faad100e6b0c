"""romis_b200/csrc/cod_fixed.hpp (the R-OMIS solve with a compile-time system size: everything in registers on the GPU) against
include/romis_cod.h (the definition the oracle runs), on the CPU: same rank and the same bits of all three solutions for hundreds
of thousands of systems -- sums of outer products like the technique matrices (full rank, every deficient rank down to 0,
ill-conditioned, repeated techniques, tiny entries), diagonal, unsymmetric, NaN / Inf entries; N = 6 (the reference's default
k = 5), 2, 4 and 11.  The GPU side of the claim is tests/test_gpu_romis.py: image bit-exact against the oracle."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fixed_size_cod_equals_the_generic_routine_bit_for_bit(tmp_path):
    exe = str(tmp_path / "cod_fixed_check")
    r = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-w", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "romis_b200", "csrc"), os.path.join(ROOT, "tests", "native", "cod_fixed_check.cpp"), "-o", exe],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout[-3000:]
    ranks = dict(kv.split(":") for kv in r.stdout.split("by rank:")[1].split("\n")[0].split())
    assert all(int(ranks[str(k)]) > 1000 for k in range(7)), f"every rank of the 6x6 system must be exercised: {ranks}"
