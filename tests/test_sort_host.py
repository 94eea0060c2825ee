"""CPU check of the event-queue sort's pure functions (collider-rs_b200/csrc/sort_cut.cuh compiled for the host by
tests/host_sort_check.cpp): order-preserving time image, monotone bucket cuts, key / record decoding."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sort_cuts_are_monotone_in_key_order(tmp_path):
    exe = tmp_path / "host_sort_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", str(exe), os.path.join(ROOT, "tests", "host_sort_check.cpp")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "HOST_SORT_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]
