"""GPU: the IMU, arm and manager shims of INTEGRATION.md are compiled code -- examples/{imt,adt,rmt}_shim.hpp (the
reference's class and member names over the C-ABI) built with g++ against librobotick_b200.so and driven as the
firmware's task loops drive the originals -- and reproduce, bit for bit, the traces of the reference compiled for x86
(tests/golden/*.npz) and the literal values of SURVEY.md Appendix D (the reference's own POS_CMD_SEQ_DEBUG_2 included)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("shims") / "shim_replay")
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "examples"),
                    os.path.join(ROOT, "examples", "shim_replay.cpp"), "-o", path, _cabi.LIB_PATH, "-Wl,-rpath," + libdir], check=True)
    return path


def run(exe, what, blob, tmp_path):
    f = tmp_path / (what + ".bin")
    f.write_bytes(blob)
    return subprocess.run([exe, what, str(f)], check=True, capture_output=True, text=True, timeout=600).stdout.splitlines()


def test_imu_shim_replays_golden_and_appendix_d(exe, tmp_path):
    g = np.load(os.path.join(GOLD, "imu_golden.npz"))
    n, K = 64, 32
    regs, have = streams.imu_samples(n, K, seed=0x5EED, drop_every=8)
    for i in (0, 5, 63):
        blob = b"".join(regs[u, :, i].astype("<i2").tobytes() + struct.pack("<i", int(have[u, i])) for u in range(K))
        rows = run(exe, "imu", blob, tmp_path)
        assert len(rows) == K
        for u, line in enumerate(rows):
            f = line.split()
            data = np.array([int(x, 16) for x in f[1:17]], dtype=np.uint32)
            np.testing.assert_array_equal(data, g["out"][u, :, i, :].reshape(16), err_msg=f"IMU {i} update {u}")
            assert int(f[17], 16) == int(data[11])  # getYawDate() == Data.angle[2]
            if u > 0:
                assert int(f[18]) == (0 if have[u, i] else 1)  # isError(): no quaternion frame since the last update
    # boot quaternion (32767, 0, 0, 0), then one sample; then an update without a quaternion frame
    boot = np.zeros(16, dtype="<i2")
    boot[12] = 32767
    smp = np.array([2048, -1024, 512, 164, -328, 16384, 11, -22, 33, -16384, 8192, -24576, 23170, 100, -200, 23170], dtype="<i2")
    rows = run(exe, "imu", boot.tobytes() + struct.pack("<i", 1) + smp.tobytes() + struct.pack("<i", 1) + (smp // 2).astype("<i2").tobytes() + struct.pack("<i", 0), tmp_path)
    d = np.array([int(x, 16) for x in rows[1].split()[1:17]], dtype=np.uint32).view(np.float32)
    want = [1, 0.5, -0.25, 10.0097656, 20.0195312, -1000, 11, 22, -33, 90, 45, -135, 0.00305166468, -0.00610332936, 0.707070708, 0.707070708]
    np.testing.assert_array_equal(d, np.array(want, dtype=np.float32))
    last = rows[2].split()
    assert last[18] == "1" and np.array([int(last[17], 16)], dtype=np.uint32).view(np.float32)[0] == np.float32(-135.0)  # held


def test_arm_shim_replays_debug_sequences(exe, tmp_path):
    g = np.load(os.path.join(GOLD, "arm_golden.npz"))
    K = int(g["K"])

    def seq_blob(img):
        b = struct.pack("<IB3x", int(img[0]), int(img[1]) & 0xFF)
        for k in range(32):
            b += struct.pack("<I", int(img[4 + 8 * k])) + img[5 + 8 * k:10 + 8 * k].astype("<u4").tobytes()
        assert len(b) == 776
        return b

    for i in (0, 1, 2, 7):  # 0..2: the reference's own POS_CMD_SEQ_DEBUG_0/1/2
        pushes = [g["seq_a"][i]] + ([g["seq_b"][i]] if g["valid_b"][i] else [])
        rows = run(exe, "arm", struct.pack("<ii", K, len(pushes)) + b"".join(seq_blob(s) for s in pushes), tmp_path)
        assert rows[0] == "status_before 99"  # NO_DATA until the first sequence starts
        assert len(rows) == K + 1
        tg = np.array([[int(x, 16) for x in r.split()[1:6]] for r in rows[1:]], dtype=np.uint32)
        np.testing.assert_array_equal(tg, g["trace"][:, 0:5, i], err_msg=f"arm {i} targets")
        st = [int(x) for x in rows[-1].split()[6:8]]
        if int(g["ids"][i]) in (1, 2):
            assert st[int(g["ids"][i]) - 1] == int(g["status"][i])
        if i == 2:  # SURVEY Appendix D, POS_CMD_SEQ_DEBUG_2
            f = tg.view(np.float32)
            np.testing.assert_array_equal(f[0], np.zeros(5, dtype=np.float32))
            np.testing.assert_array_equal(f[1], np.array([0, 120, -90, 0, 45], dtype=np.float32))
            np.testing.assert_array_equal(f[100], np.array([19.6000004, 61.2000008, -31.2000008, 44.0999985, -57.9000015], dtype=np.float32))
            np.testing.assert_array_equal(f[304], np.array([0, 120, -60, 0, 45], dtype=np.float32))
            # alone in the ring (Appendix D's run): PROCESSING from the first tick, DONE once the last segment has ended
            solo = run(exe, "arm", struct.pack("<ii", K, 1) + seq_blob(g["seq_a"][i]), tmp_path)
            done = [int(r.split()[6]) for r in solo[1:]]
            assert done[0] == 0 and 304 <= done.index(1) <= 306 and all(x == 1 for x in done[done.index(1):])
            np.testing.assert_array_equal(np.array([[int(x, 16) for x in r.split()[1:6]] for r in solo[1:306]], dtype=np.uint32), tg[:305])


def test_manager_shim_replays_golden(exe, tmp_path):
    g = np.load(os.path.join(GOLD, "rmt_golden.npz"))
    n, K = 40, 460
    inp = streams.rm_inputs(n, K, seed=0x5EED)  # [K, 3, n, 4]
    for i in (0, 7, 39):
        blob = np.ascontiguousarray(inp[:, :, i, :]).astype("<u4").tobytes()
        rows = run(exe, "rmt", blob, tmp_path)
        assert len(rows) == K
        got = np.array([[int(x, 16) for x in r.split()[1:6]] for r in rows], dtype=np.uint32)
        np.testing.assert_array_equal(got[:, :4], g["cmd"][:, i, :], err_msg=f"manager {i} message")
        np.testing.assert_array_equal(got[:, 4], g["abort"][:, i], err_msg=f"manager {i} fault word")
