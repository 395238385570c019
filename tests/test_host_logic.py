"""Host-side logic of the C++ mirror that needs no GPU (DataPoints descriptors and concatenate's common-descriptor rule, the RAM /
hard-drive CellManagers behind the reference's interface with their VTK files, the YAML reader's tree): a small C++ program
(tests/cpp/test_host_logic.cpp) compiled against the headers of norlab_icp_mapper_b200/host/ and run here."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_logic_cpp(tmp_path):
    exe = tmp_path / "test_host_logic"
    src = os.path.join(ROOT, "tests", "cpp", "test_host_logic.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-o", str(exe), src], check=True)
    cells = tmp_path / "cells"
    cells.mkdir()
    r = subprocess.run([str(exe), str(cells)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for tag in ("ok descriptors", "ok concatenate", "ok cell managers", "ok yaml tree"):
        assert tag in r.stdout
