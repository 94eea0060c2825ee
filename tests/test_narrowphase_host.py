"""Pre-GPU logic check: the product's device arithmetic (csrc/narrowphase.cuh, also compilable for the host)
against the oracle, bit for bit, on 1M seeded random + lattice pairs.  The product library never calls the host
instantiation; this only catches logic slips before GPU time is spent."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_instantiation_matches_oracle(tmp_path):
    exe = tmp_path / "host_narrowphase_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_narrowphase_check.cpp")], check=True)
    res = subprocess.run([str(exe), "1000000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "mismatches 0" in res.stdout
