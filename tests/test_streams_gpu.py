"""The on-device stream generators (rk_stream_*, csrc/rk_stream.cu) against their definition, the numpy `*_v2`
generators of streams.py: every block bit for bit, at odd sizes, non-zero first robot / first update and non-default
distribution parameters."""
import numpy as np
import pytest
import torch

from roboken_fmskf_robot_controller_b200 import layout, streams
from roboken_fmskf_robot_controller_b200.devstreams import DeviceStreams

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,first,seed", [(1, 0, 0x5EED), (1000, 0, 0x5EED), (4097, 123456789, 7), (2050, (1 << 24) - 1000, 0xFFFFFFFF)])
def test_device_streams_equal_numpy_definition(n, first, seed):
    ds = DeviceStreams(DEV, seed=seed, first=first, stop_every=5, drop_every=7, arm_min_len=3, arm_max_len=17, arm_seq_id=9,
                       arm_dt_zero_every=3, first_update=11)
    n_seg, n_yaw, n_upd = 9, 37, 23
    cmd = ds.vehicle_commands(torch.zeros((n_seg, n, 4), dtype=torch.int32, device=DEV))
    yaw = ds.vehicle_yaw_reg(torch.zeros((n_yaw, n), dtype=torch.int16, device=DEV))
    regs, have = ds.imu_samples(torch.zeros((n_upd, 2, n, 8), dtype=torch.int16, device=DEV), torch.zeros((n_upd, n), dtype=torch.uint8, device=DEV))
    seq = ds.arm_sequences(torch.zeros(layout.ACMD_SLOT_WORDS * n, dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    exp = streams.vehicle_commands_v2(n, n_seg, seed, first, stop_every=5)
    np.testing.assert_array_equal(cmd.cpu().numpy().view(np.uint32).reshape(n_seg, n, 4), exp.view(np.uint32).reshape(n_seg, n, 4))
    np.testing.assert_array_equal(yaw.cpu().numpy(), streams.vehicle_yaw_reg_v2(n, n_yaw, seed, first))
    eregs, ehave = streams.imu_samples_v2(n, n_upd, seed, first, drop_every=7, first_update=11)
    np.testing.assert_array_equal(regs.cpu().numpy(), streams.imu_cells(eregs))
    np.testing.assert_array_equal(have.cpu().numpy(), ehave)
    eseq = streams.arm_sequences_v2(n, seed, first, max_len=17, min_len=3, seq_id=9, dt_zero_every=3)
    np.testing.assert_array_equal(seq.cpu().numpy().view(np.uint32), layout.aos_to_soa(eseq))


def test_device_streams_defaults_and_no_have_block():
    n = 300
    ds = DeviceStreams(DEV)
    regs, _ = ds.imu_samples(torch.zeros((5, 2, n, 8), dtype=torch.int16, device=DEV), None)
    torch.cuda.synchronize()
    eregs, _ = streams.imu_samples_v2(n, 5)
    np.testing.assert_array_equal(regs.cpu().numpy(), streams.imu_cells(eregs))
    q = eregs[:, 12:16, :].astype(np.float64)
    assert np.abs(np.sqrt((q * q).sum(axis=1)) - 32767.0).max() < 1.5  # unit quaternions x 32767


def test_capped_generators_give_the_same_blocks():
    """RK_OPT_STREAM_CTAS only changes how the robots are spread over CTAs."""
    from roboken_fmskf_robot_controller_b200 import _cabi

    n, lib = 70001, _cabi.load()
    ref = None
    try:
        for cap in (0, 1, 3):
            lib.rk_set_option(_cabi.RK_OPT_STREAM_CTAS, cap)
            ds = DeviceStreams(DEV, seed=5, first=99)
            got = [ds.vehicle_commands(torch.zeros((3, n, 4), dtype=torch.int32, device=DEV)),
                   ds.imu_samples(torch.zeros((4, 2, n, 8), dtype=torch.int16, device=DEV), torch.zeros((4, n), dtype=torch.uint8, device=DEV))[0],
                   ds.arm_sequences(torch.zeros(layout.ACMD_SLOT_WORDS * n, dtype=torch.int32, device=DEV))]
            torch.cuda.synchronize()
            got = [g.cpu().numpy().copy() for g in got]
            if ref is None:
                ref = got
            for a, b in zip(got, ref):
                np.testing.assert_array_equal(a, b)
    finally:
        lib.rk_set_option(_cabi.RK_OPT_STREAM_CTAS, 0)
