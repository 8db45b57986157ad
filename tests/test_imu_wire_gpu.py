"""GPU parity: rk_imt_feed_bytes (the WIT serial codec on the device) vs the oracle, the golden fixture
made by the compiled vendor parser, and the register-level kernel.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from roboken_fmskf_robot_controller_b200 import layout, streams
from roboken_fmskf_robot_controller_b200.imu import ImuBatch
from test_imu_wire_cpu import parser_sreg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_feed(ib, cells, nbytes, do_init, want_yaw=False):
    K, _, n, _ = cells.shape
    out = torch.zeros((K, 4, n, 4), dtype=torch.float32, device=DEV)
    yaw = torch.zeros((K, n), dtype=torch.float32, device=DEV) if want_yaw else None
    ib.feed_bytes(torch.from_numpy(cells.view(np.int32)).to(DEV),
                  None if nbytes is None else torch.from_numpy(nbytes.view(np.int16)).to(DEV), out, yaw, do_init)
    torch.cuda.synchronize()
    o = out.cpu().numpy().view(np.uint32)
    return (o, yaw.cpu().numpy()) if want_yaw else o


def blocks(ib):
    return ib.state.cpu().numpy().view(np.uint32), ib.parser.cpu().numpy().view(np.uint32)


def test_wire_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "imu_wire_golden.npz"))
    n = 48
    ib = ImuBatch(n, DEV)
    out = gpu_feed(ib, g["cells"], g["nbytes"], True)
    st, ps = blocks(ib)
    np.testing.assert_array_equal(out, g["out"])
    np.testing.assert_array_equal(st, g["state"])
    np.testing.assert_array_equal(parser_sreg(ps, n), g["sreg"])


@pytest.mark.parametrize("n,K,ncells,seed,full", [(1, 6, 1, 1, False), (300, 40, 4, 2, False), (1031, 12, 2, 3, False),
                                                   (129, 7, 10, 4, False), (513, 20, 3, 5, True),
                                                   (2049, 30, 4, 6, True), (700, 25, 4, 7, False)])  # ncells == 4: the static-offset path
def test_wire_fuzz_vs_port(n, K, ncells, seed, full):
    wire, nb = streams.imu_wire_fuzz(n, K, ncells=ncells, seed=seed, full_slots=full)
    if full:
        nb = None
    a, pa = np.zeros(layout.IS_WORDS * n, dtype=np.uint32), np.zeros(layout.IP_WORDS * n, dtype=np.uint32)
    oa, ya = ol.imu_bytes_port(a, pa, n, wire, nb, want_out=True, want_yaw=True, do_init=True)
    ib = ImuBatch(n, DEV)
    out, yaw = gpu_feed(ib, wire, nb, True, want_yaw=True)
    st, ps = blocks(ib)
    np.testing.assert_array_equal(out, oa)
    np.testing.assert_array_equal(yaw.view(np.uint32), ya.view(np.uint32))
    np.testing.assert_array_equal(st, a)
    np.testing.assert_array_equal(ps, pa)  # window bytes, fill count, read index, pending flag, sReg: all of it
    # carry on from that parser state, no init, different traffic
    wire2, nb2 = streams.imu_wire_fuzz(n, K, ncells=ncells, seed=seed + 50, full_slots=full)
    if full:
        nb2 = None
    ob = ol.imu_bytes_port(a, pa, n, wire2, nb2, want_out=True)
    out2 = gpu_feed(ib, wire2, nb2, False)
    st, ps = blocks(ib)
    np.testing.assert_array_equal(out2, ob)
    np.testing.assert_array_equal(st, a)
    np.testing.assert_array_equal(ps, pa)


def test_wire_one_launch_equals_many():
    n, K, ncells = 257, 24, 2
    wire, nb = streams.imu_wire_fuzz(n, K, ncells=ncells, seed=77)
    one = ImuBatch(n, DEV)
    o1 = gpu_feed(one, wire, nb, True)
    many, outs, k0 = ImuBatch(n, DEV), [], 0
    for k1 in (1, 5, 6, 17, 24):
        outs.append(gpu_feed(many, np.ascontiguousarray(wire[k0:k1]), np.ascontiguousarray(nb[k0:k1]), k0 == 0))
        k0 = k1
    np.testing.assert_array_equal(np.concatenate(outs), o1)
    for x, y in zip(blocks(one), blocks(many)):
        np.testing.assert_array_equal(x, y)


def test_clean_wire_equals_register_kernel():
    """A healthy sensor's five frames through the byte codec == the same registers through rk_imt_update."""
    n, K = 5000, 16
    regs, _ = streams.imu_samples(n, K, seed=31)
    wire, nb = streams.imu_wire_clean(regs)
    a = ImuBatch(n, DEV)
    oa = torch.zeros((K, 4, n, 4), dtype=torch.float32, device=DEV)
    a.update(torch.from_numpy(streams.imu_cells(regs)).to(DEV), None, oa, True)
    b = ImuBatch(n, DEV)
    ob = gpu_feed(b, wire, nb, True)
    np.testing.assert_array_equal(oa.cpu().numpy().view(np.uint32), ob)
    np.testing.assert_array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    np.testing.assert_array_equal(parser_sreg(blocks(b)[1], n), regs[-1].T)


def test_feed_bytes_argument_errors():
    ib = ImuBatch(4, DEV)
    from roboken_fmskf_robot_controller_b200 import _cabi
    lib = _cabi.load()
    ib.parser = torch.zeros(layout.IP_WORDS * 4, dtype=torch.int32, device=DEV)
    assert lib.rk_imt_feed_bytes(ib.state.data_ptr(), ib.parser.data_ptr(), 4, 1, 2, None, None, None, None, 0, None) == 1
    assert lib.rk_imt_feed_bytes(ib.state.data_ptr(), None, 4, 1, 0, None, None, None, None, 0, None) == 1
    assert lib.rk_imt_feed_bytes(ib.state.data_ptr(), ib.parser.data_ptr(), 0, 1, 2, None, None, None, None, 0, None) == 0


def test_init_waits_for_a_late_first_quaternion_frame():
    """with the first quaternion frame up to three update slots late, the device neither latches q_init nor
    publishes before it -- same bits as the port (pinned against the compiled reference's blocking init() on the CPU) --
    in one launch and with the wait carried across a launch boundary in the parser block."""
    from test_imu_wire_cpu import late_quaternion_wire

    n, K = 300, 10
    cells, nb, late = late_quaternion_wire(n, K, 91)
    st, ps = np.zeros(layout.IS_WORDS * n, dtype=np.uint32), np.zeros(layout.IP_WORDS * n, dtype=np.uint32)
    exp = ol.imu_bytes_port(st, ps, n, cells, nb, want_out=True, do_init=True)
    ib = ImuBatch(n, DEV)
    out = gpu_feed(ib, cells, nb, True)
    gs, gp = blocks(ib)
    np.testing.assert_array_equal(out, exp)
    np.testing.assert_array_equal(gs, st)
    np.testing.assert_array_equal(gp, ps)
    ib2 = ImuBatch(n, DEV)
    o1 = gpu_feed(ib2, np.ascontiguousarray(cells[:2]), np.ascontiguousarray(nb[:2]), True)
    o2 = gpu_feed(ib2, np.ascontiguousarray(cells[2:]), np.ascontiguousarray(nb[2:]), False)
    np.testing.assert_array_equal(np.concatenate([o1, o2]), exp)
    gs2, gp2 = blocks(ib2)
    np.testing.assert_array_equal(gs2, st)
    np.testing.assert_array_equal(gp2, ps)
