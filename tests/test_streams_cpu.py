"""The v2 stream definitions (streams.py, 32-bit counter hash) that the device generators reproduce: counter-based
(any subset of robots / updates can be generated on its own), inside the documented ranges, and pinned to a few
literal values so that a silent change of the definition is caught on the CPU."""
import numpy as np

from roboken_fmskf_robot_controller_b200 import _cabi, streams


def test_hash_known_answers():
    assert int(streams.mix32(np.uint32(0))) == 0
    assert [int(x) for x in streams.mix32(np.array([1, 2, 0xFFFFFFFF], dtype=np.uint32))] == [0x688990C0, 0xD1132181, 0x6768824A]
    assert int(streams.h32(0x5EED, 20, 12345, 7)) == 0x9E789B82 and int(streams.sub32(np.uint32(0x12345678), 3)) == 0xA372FE14
    assert int(streams.lite32(np.uint32(0x12345678), 3)) == 0xEC8E46C0
    h = streams.h32(0x5EED, 20, np.array([0, 1, 1 << 24], dtype=np.uint64), np.array([0, 7, 99], dtype=np.uint64))
    assert h.dtype == np.uint32 and len(set(int(x) for x in h)) == 3
    # instance and index enter modulo 2^32
    assert int(streams.h32(1, 2, (1 << 32) + 5, 9)) == int(streams.h32(1, 2, 5, 9))


def test_subsets_equal_slices():
    n, first = 300, 10_000_000
    idx = np.array([0, 17, 299], dtype=np.uint64)
    g = idx + np.uint64(first)
    cmd = streams.vehicle_commands_v2(n, 8, 0x5EED, first)
    np.testing.assert_array_equal(streams.vehicle_commands_v2(0, 8, 0x5EED, inst=g), cmd[:, idx.astype(int)])
    regs, have = streams.imu_samples_v2(n, 30, 0x5EED, first)
    r2, h2 = streams.imu_samples_v2(0, 30, 0x5EED, inst=g)
    np.testing.assert_array_equal(r2, regs[:, :, idx.astype(int)])
    np.testing.assert_array_equal(h2, have[:, idx.astype(int)])
    r3, h3 = streams.imu_samples_v2(n, 10, 0x5EED, first, first_update=20)
    np.testing.assert_array_equal(r3, regs[20:])
    np.testing.assert_array_equal(h3, have[20:])
    np.testing.assert_array_equal(streams.arm_sequences_v2(0, 0x5EED, inst=g), streams.arm_sequences_v2(n, 0x5EED, first)[idx.astype(int)])
    np.testing.assert_array_equal(streams.vehicle_yaw_reg_v2(0, 12, 0x5EED, inst=g), streams.vehicle_yaw_reg_v2(n, 12, 0x5EED, first)[:, idx.astype(int)])


def test_distributions():
    n = 4096
    cmd = streams.vehicle_commands_v2(n, 16)
    speed = np.sqrt(cmd["vx"].astype(np.float64) ** 2 + cmd["vy"].astype(np.float64) ** 2)
    assert speed.max() <= 400.0 * (1 + 1e-6) and np.abs(cmd["vth"]).max() <= 2 * np.pi * (1 + 1e-6)
    stop = cmd["kind"] == _cabi.RK_CMD_STOP
    assert 0.09 < stop.mean() < 0.16 and (cmd["vx"][stop] == 0).all() and (cmd["kind"][~stop] == _cabi.RK_CMD_MOVE).all()
    regs, have = streams.imu_samples_v2(n, 64)
    q = regs[:, 12:16, :].astype(np.float64)
    assert np.abs(np.sqrt((q * q).sum(axis=1)) - 32767.0).max() < 1.5
    assert 0.975 < have.mean() < 0.995
    assert regs[:, :12, :].min() < -32000 and regs[:, :12, :].max() > 32000
    img = streams.arm_sequences_v2(n)
    ln = img[:, 1]
    assert ln.min() == 2 and ln.max() == 32 and (img[:, 0] == 1).all()
    wp = img[:, 4:].reshape(n, 32, 8)
    used = np.arange(32)[None, :] < ln[:, None]  # waypoints past the length are zero
    dd = np.diff(wp[:, :, 0].astype(np.int64), axis=1)
    assert (dd[used[:, 1:]] >= 0).all() and (wp[:, :, 6:] == 0).all() and (wp[~used] == 0).all()
    ang = wp[:, :, 1:6].view(np.float32)
    assert ang.min() >= -150.0 and ang.max() <= 150.0 and (ang * 64 == np.rint(ang * 64)).all()
    yaw = streams.vehicle_yaw_reg_v2(n, 50).astype(np.int64)
    d = np.diff(yaw, axis=0) & 0xFFFF
    assert (d == d[0]).all() and set(np.unique(np.minimum(d[0], 65536 - d[0]))) <= {182, 364, 546, 728, 910}


def test_pinned_values():
    """Literal outputs of the definition (seed 0x5EED, robot 0 / 12345): a change of the hash or of the float recipe shows
    up here before it shows up as a GPU mismatch."""
    cmd = streams.vehicle_commands_v2(1, 2)
    assert cmd["vx"][0, 0].tobytes().hex() == np.float32(308.5701).tobytes().hex() or abs(float(cmd["vx"][0, 0]) - 308.5701) < 1e-4
    regs, _ = streams.imu_samples_v2(1, 1)
    assert [int(x) for x in regs[0, :4, 0]] == [-31763, 24349, -30630, 22169]
    assert [int(x) for x in streams.vehicle_yaw_reg_v2(4, 2)[1]] == [5080, 24669, -26840, 14572]
    img = streams.arm_sequences_v2(1)
    assert [int(x) for x in img[0, :5]] == [1, 10, 0, 0, 554]
