"""GPU parity: rk_rmt_guard / rk_mymath_atan2f (the RobotManager guard) vs the oracle, the golden
fixture made by the compiled RobotManager task, and chained into rk_vdt_rollout's command layer.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams
from roboken_fmskf_robot_controller_b200.rmt import ManagerBatch, atan2f
from roboken_fmskf_robot_controller_b200.vehicle import VehicleBatch
from test_rmt_cpu import generated_tables, special_values

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_guard(mb, inp, want_abort=True):
    K, _, n, _ = inp.shape
    cmd = torch.zeros((K, n, 4), dtype=torch.int32, device=DEV)
    ab = torch.zeros((K, n), dtype=torch.int32, device=DEV) if want_abort else None
    mb.guard(torch.from_numpy(inp.view(np.int32)).to(DEV), cmd, ab)
    torch.cuda.synchronize()
    return cmd.cpu().numpy().view(np.uint32), (ab.cpu().numpy().view(np.uint32) if want_abort else None)


def test_guard_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rmt_golden.npz"))
    n, K = 40, 460
    mb = ManagerBatch(n, DEV)
    cmd, ab = gpu_guard(mb, streams.rm_inputs(n, K, seed=0x5EED))
    np.testing.assert_array_equal(cmd, g["cmd"])
    np.testing.assert_array_equal(ab, g["abort"])
    np.testing.assert_array_equal(mb.state.cpu().numpy().view(np.uint32), g["state"])
    out = atan2f(torch.from_numpy(g["atan_y"]).to(DEV), torch.from_numpy(g["atan_x"]).to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(out.view(np.uint32), g["atan_out"].view(np.uint32))


@pytest.mark.parametrize("n,K,seed", [(1, 30, 1), (1000, 640, 2), (4099, 250, 3)])
def test_guard_vs_port(n, K, seed):
    inp = streams.rm_inputs(n, K, seed=seed)
    a = np.zeros(layout.RS_WORDS * n, dtype=np.uint32)
    ca, aa = ol.rm_guard("port", a, n, inp)
    mb = ManagerBatch(n, DEV)
    cmd, ab = gpu_guard(mb, inp)
    np.testing.assert_array_equal(cmd, ca)
    np.testing.assert_array_equal(ab, aa)
    np.testing.assert_array_equal(mb.state.cpu().numpy().view(np.uint32), a)
    # resume with other parameters, no abort output
    mb.params = _cabi.RmtParams(9, 77, 650)
    inp2 = streams.rm_inputs(n, 64, seed=seed + 20)
    ca, _ = ol.rm_guard("port", a, n, inp2, params=mb.params)
    cmd, _ = gpu_guard(mb, inp2, want_abort=False)
    np.testing.assert_array_equal(cmd, ca)
    np.testing.assert_array_equal(mb.state.cpu().numpy().view(np.uint32), a)


def test_guard_one_launch_equals_many():
    n, K = 513, 420
    inp = streams.rm_inputs(n, K, seed=5)
    one, many = ManagerBatch(n, DEV), ManagerBatch(n, DEV)
    c1, a1 = gpu_guard(one, inp)
    cs, k0 = [], 0
    for k1 in (1, 100, 101, 300, 420):
        cs.append(gpu_guard(many, np.ascontiguousarray(inp[k0:k1]))[0])
        k0 = k1
    np.testing.assert_array_equal(np.concatenate(cs), c1)
    np.testing.assert_array_equal(one.state.cpu().numpy(), many.state.cpu().numpy())


def test_atan2f_vs_port_dense():
    rng = np.random.default_rng(12)
    d = generated_tables()[1]
    xs = np.concatenate([special_values(), d, np.nextafter(d, np.float32(np.inf)), np.nextafter(d, np.float32(-np.inf)),
                         (rng.standard_normal(200000) * 3).astype(np.float32), np.exp(rng.uniform(-25, 25, 200000)).astype(np.float32)])
    ys = rng.permutation(xs)
    ones = np.ones_like(xs)
    p = ol.port()
    for y, x in ((xs, ones), (ys, xs), (xs, -ones), (ys, np.zeros_like(xs))):
        got = atan2f(torch.from_numpy(y).to(DEV), torch.from_numpy(x).to(DEV)).cpu().numpy()
        idx = np.concatenate([np.arange(400), rng.integers(0, len(xs), 3000)])
        exp = np.array([p.orc_atan2f(float(y[k]), float(x[k])) for k in idx], dtype=np.float32)
        np.testing.assert_array_equal(got[idx].view(np.uint32), exp.view(np.uint32))
    # the whole array against numpy's arctan2: the table lerp is within 2e-4 of the true angle
    y, x = (rng.standard_normal(100000) * 100).astype(np.float32), (rng.standard_normal(100000) * 100).astype(np.float32)
    got = atan2f(torch.from_numpy(y).to(DEV), torch.from_numpy(x).to(DEV)).cpu().numpy()
    assert np.max(np.abs(got - np.arctan2(y, x))) < 2e-4


def test_guard_chained_into_vehicle_rollout():
    """Guarded rollout: the manager's output records drive rk_vdt_rollout's command layer (one manager cycle per
    20 ms segment) -- device chain == port chain, state and trace."""
    n, K, seg_len = 300, 60, 20
    steps = K * seg_len
    inp = streams.rm_inputs(n, K, seed=21, idle_every=1000)
    a = np.zeros(layout.RS_WORDS * n, dtype=np.uint32)
    ca, _ = ol.rm_guard("port", a, n, inp)
    yaw = streams.vehicle_yaw(n, steps // 10, 21)
    st = np.zeros(layout.VS_WORDS * n, dtype=np.uint32)
    ro = ol.HostRollout(n, steps, _cabi.RK_SENSOR_PLANT, ca.view(np.dtype([("vx", "<f4"), ("vy", "<f4"), ("vth", "<f4"), ("kind", "<i4")])).reshape(K, n),
                        seg_len, yaw, 10, trace=True, task_period=10)
    ol.run_port(st, n, ro)
    mb, vb = ManagerBatch(n, DEV), VehicleBatch(n, DEV)
    cmd = torch.zeros((K, n, 4), dtype=torch.int32, device=DEV)
    mb.guard(torch.from_numpy(inp.view(np.int32)).to(DEV), cmd)
    tr = torch.zeros((steps, 16, n), dtype=torch.int32, device=DEV)
    vb.rollout(steps, sensor_mode=_cabi.RK_SENSOR_PLANT, cmd=cmd, seg_len=seg_len, yaw=torch.from_numpy(yaw).to(DEV), yaw_period=10,
               trace=tr, task_period=10)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(tr.cpu().numpy().view(np.uint32), ro.trace)
    np.testing.assert_array_equal(vb.state.cpu().numpy().view(np.uint32), st)
    assert (ro.trace[:, 3:6, :].view(np.float32) != 0).any()  # the robots did move


def test_guard_argument_errors():
    lib = _cabi.load()
    mb = ManagerBatch(4, DEV)
    p = mb.params
    import ctypes as C
    assert lib.rk_rmt_guard(C.byref(p), mb.state.data_ptr(), 4, 1, None, None, None, None) == 1
    assert lib.rk_rmt_guard(C.byref(p), mb.state.data_ptr(), 0, 1, None, None, None, None) == 0
    assert lib.rk_rmt_guard(C.byref(p), mb.state.data_ptr() + 4, 4, 1, mb.state.data_ptr(), mb.state.data_ptr(), None, None) == 1


def test_manager_handle_equals_port():
    """rk_rmt_t: one routine_ros() block per call on host words == the port, cycle by cycle."""
    from roboken_fmskf_robot_controller_b200.rmt import Manager

    K = 260
    inp = streams.rm_inputs(1, K, seed=4, idle_every=1)  # index 0 is a mostly-silent robot: the watchdog fires
    st = np.zeros(layout.RS_WORDS, dtype=np.uint32)
    ca, aa = ol.rm_guard("port", st, 1, inp)
    m = Manager()
    for u in range(K):
        rec, ab = m.cycle(inp[u, :, 0, :].reshape(-1))
        assert rec == list(ca[u, 0]) and ab == aa[u, 0], u
    assert m.state() == list(st)
    m.close()
