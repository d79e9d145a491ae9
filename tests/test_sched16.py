"""CPU test of the conflict-free half-warp scheduler (csrc/sched16.cuh is host/device code): compiled
with g++ and run on random, ragged and adversarial row patterns."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("impl", ["bvn", "colouring"])   # the device kernel runs the Birkhoff-von Neumann variant
@pytest.mark.parametrize("nl", [16, 8])            # half-warp / 8-byte gathers and quarter-warp / 16-byte gathers
def test_sched16_edge_colouring_on_cpu(tmp_path, nl, impl):
    exe = tmp_path / "test_sched16"
    subprocess.check_call(["g++", "-O2", "-std=c++17", f"-DTNL={nl}", *(["-DUSE_BVN"] if impl == "bvn" else []),
                           "-I", os.path.join(ROOT, "sparsifiedkmeans_b200", "csrc"),
                           "-o", str(exe), os.path.join(ROOT, "tests", "native", "test_sched16.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad 0" in out.stdout


def test_sched16_under_address_and_ub_sanitizers(tmp_path):
    """The same scheduler header under -fsanitize=address,undefined (both forms, half-warp problems): no
    out-of-bounds scratch access, no undefined shifts in the bit masks."""
    for extra in (["-DUSE_BVN"], []):
        exe = tmp_path / ("san" + str(len(extra)))
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                               "-DTNL=16", *extra, "-I", os.path.join(ROOT, "sparsifiedkmeans_b200", "csrc"),
                               "-o", str(exe), os.path.join(ROOT, "tests", "native", "test_sched16.cpp")])
        out = subprocess.run([str(exe)], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-3000:]
