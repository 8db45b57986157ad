"""GPU: the firmware-side C++ shim of INTEGRATION.md is real code -- examples/vdt_shim.hpp (VEHICLE_CTRL / MOTOR_IF_M2006
with the reference's member names over the C-ABI) compiled with g++ against librobotick_b200.so replays BASELINE
configs[0] (1 vehicle, 10 s at 1 kHz) and reproduces the golden trace of the compiled reference bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from roboken_fmskf_robot_controller_b200 import _cabi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_replays_c1_golden(tmp_path):
    exe = str(tmp_path / "c1_replay")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "examples"),
                    os.path.join(ROOT, "examples", "c1_replay.cpp"), "-o", exe, _cabi.LIB_PATH, "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout
    rows = {}
    for line in out.splitlines():
        f = line.split()
        rows[int(f[0])] = [int(x, 16) for x in f[1:10]] + [int(x) & 0xFFFFFFFF for x in f[10:14]]
    g = np.load(os.path.join(ROOT, "tests", "golden", "vdt_golden.npz"))
    assert sorted(rows) == list(g["c1_rows"])
    got = np.array([rows[t] for t in g["c1_rows"]], dtype=np.uint32)
    np.testing.assert_array_equal(got, g["c1_trace"][:, :13, 0])
