"""CPU check of the column formulation of the point-triangle distance (baby_shark_b200/csrc/bs_ptdist.cuh, what k_eval
runs on the device) against the oracle's closest_point (triangle3.rs:317-382 restated in oracle/bso_convert.h): the header
is compiled for the host with -ffp-contract=off and must give the same bits on random, degenerate, sliver and
lattice-aligned triangles."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_column_distance_bits_equal_closest_point(tmp_path):
    exe = str(tmp_path / "ptdist_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-o", exe, os.path.join(ROOT, "tests", "host", "ptdist_check.cpp")])
    out = subprocess.run([exe, "80000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert " 0 mismatches" in out.stdout
