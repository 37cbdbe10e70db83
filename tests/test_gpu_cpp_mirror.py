"""GPU test (-m gpu): the C++ host mirror (include/baby_shark.hpp) compiled against libbshark_cuda.so runs the
reference's own tests for this path (7944 known answer, voxel remesher, dual contouring example)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_mirror(tmp_path):
    exe = str(tmp_path / "mirror_test")
    lib_dir = os.path.join(ROOT, "baby_shark_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"),
                           "-L", lib_dir, "-lbshark_cuda", "-Wl,-rpath," + lib_dir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "mirror ok" in out.stdout, out.stdout + out.stderr
