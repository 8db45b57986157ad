"""GPU parity of the full controller tick (rk_tick_rollout: vehicle + IMU + arm coupled through the
IMU yaw) against the per-module oracles composed on the host, bit for bit."""
import numpy as np
import pytest
import torch

import oracle_lib as ol
from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams
from roboken_fmskf_robot_controller_b200.robot import RobotBatch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_inputs(n, steps, slow, seed, seg_len=125):
    n_seg = (steps + seg_len - 1) // seg_len
    n_slow = (steps + slow - 1) // slow
    cmd = streams.vehicle_commands(n, n_seg, seed)
    regs, have = streams.imu_samples(n, n_slow + 1, seed=seed, drop_every=16)
    seq = streams.arm_sequences(n, seed=seed, seq_id=3, max_len=12)
    return cmd, regs, have, seq


def gpu_full(n, steps, slow, cmd, regs, have, seq, trace=True, goal=None):
    rb = RobotBatch(n, DEV)
    rb.reset()
    # IMU boot: IMU_IF_WT901C::init() consumes the first sample; arm: one sequence pushed
    rb.imu.update(torch.from_numpy(streams.imu_cells(regs[:1])).to(DEV), None, None, do_init=True)
    rb.arm.push_cmdseq(torch.from_numpy(layout.aos_to_soa(seq).view(np.int32)).to(DEV))
    n_slow = (steps + slow - 1) // slow
    cmd_d = torch.from_numpy(cmd.view(np.int32).reshape(cmd.shape[0], n, 4)).to(DEV)
    regs_d = torch.from_numpy(streams.imu_cells(regs[1 : 1 + n_slow])).to(DEV)
    have_d = torch.from_numpy(np.ascontiguousarray(have[1 : 1 + n_slow])).to(DEV)
    yaw_d = torch.zeros((n_slow, n), dtype=torch.float32, device=DEV)
    vtr = torch.zeros((steps, 16, n), dtype=torch.int32, device=DEV) if trace else None
    atr = torch.zeros((n_slow, 16, n), dtype=torch.int32, device=DEV) if trace else None
    goal_d = cost_d = None
    if goal is not None:
        goal_d, cost_d = torch.from_numpy(goal).to(DEV), torch.zeros(n, dtype=torch.float32, device=DEV)
    rb.rollout(steps, slow, cmd=cmd_d, seg_len=125, regs=regs_d, have_quat=have_d, yaw=yaw_d, vdt_trace=vtr, adt_trace=atr,
               goal=goal_d, cost=cost_d)
    torch.cuda.synchronize()
    u = lambda t: None if t is None else t.cpu().numpy().view(np.uint32)
    return dict(v=u(rb.vehicle.state), i=u(rb.imu.state), a=u(rb.arm.state), vtr=u(vtr), atr=u(atr), yaw=yaw_d.cpu().numpy(),
                cost=None if cost_d is None else cost_d.cpu().numpy())


def oracle_full(kind, n, steps, slow, cmd, regs, have, seq, trace=True, goal=None):
    v = np.zeros(layout.VS_WORDS * n, dtype=np.uint32)
    i = np.zeros(layout.IS_WORDS * n, dtype=np.uint32)
    a = np.zeros(layout.AS_WORDS * n, dtype=np.uint32)
    t = np.zeros(layout.ACMD_WORDS * n, dtype=np.uint32)
    (ol.imu_port if kind == "port" else ol.imu_ref)(i, n, regs[:1], None, do_init=True)
    ol.arm_batch(kind, "init", a, t, n)
    ol.arm_batch(kind, "push", a, t, n, seq=layout.aos_to_soa(seq))
    n_slow = (steps + slow - 1) // slow
    vtr, atr, yaw, cost = ol.full_tick(kind, n, steps, slow, cmd, 125, np.ascontiguousarray(regs[1 : 1 + n_slow]),
                                       np.ascontiguousarray(have[1 : 1 + n_slow]), v, i, a, t, trace=trace, goal=goal, nthreads=8)
    return dict(v=v, i=i, a=a, vtr=vtr, atr=atr, yaw=yaw, cost=cost)


def compare(g, o):
    np.testing.assert_array_equal(g["yaw"].view(np.uint32), o["yaw"].view(np.uint32))
    for k in ("vtr", "atr", "v", "i", "a"):
        if g[k] is not None:
            np.testing.assert_array_equal(g[k], o[k], err_msg=k)


@pytest.mark.parametrize("n,steps,seed", [(300, 1000, 41), (1, 95, 42), (2050, 250, 43)])
def test_full_tick_vs_port(n, steps, seed):
    inp = make_inputs(n, steps, 10, seed)
    compare(gpu_full(n, steps, 10, *inp), oracle_full("port", n, steps, 10, *inp))


@pytest.mark.skipif(not (ol.have_ref("libref_arm.so") and ol.have_ref("libref_imu.so")), reason="oracle/_ref not present")
def test_full_tick_vs_compiled_reference():
    n, steps = 64, 600
    inp = make_inputs(n, steps, 10, 44)
    compare(gpu_full(n, steps, 10, *inp), oracle_full("ref", n, steps, 10, *inp))


def test_full_tick_cost_and_slice_invariance():
    """The G-GPU result equals the 1-GPU result slice by slice: a rollout of the second
    half of the instances alone gives the second half of the full rollout's costs and states."""
    n, steps = 512, 400
    cmd, regs, have, seq = make_inputs(n, steps, 10, 45)
    goal = np.zeros((n, 2), dtype=np.float32)
    full = gpu_full(n, steps, 10, cmd, regs, have, seq, trace=False, goal=goal)
    h = n // 2
    half = gpu_full(h, steps, 10, np.ascontiguousarray(cmd[:, h:]), np.ascontiguousarray(regs[:, :, h:]),
                    np.ascontiguousarray(have[:, h:]), seq[h:], trace=False, goal=goal[h:])
    np.testing.assert_array_equal(full["cost"][h:].view(np.uint32), half["cost"].view(np.uint32))
    np.testing.assert_array_equal(layout.soa_to_aos(full["v"], n, layout.VS_WORDS)[h:], layout.soa_to_aos(half["v"], h, layout.VS_WORDS))
    np.testing.assert_array_equal(layout.soa_to_aos(full["a"], n, layout.AS_WORDS)[h:], layout.soa_to_aos(half["a"], h, layout.AS_WORDS))
    o = oracle_full("port", n, steps, 10, cmd, regs, have, seq, trace=False, goal=goal)
    np.testing.assert_array_equal(full["cost"].view(np.uint32), o["cost"].view(np.uint32))
