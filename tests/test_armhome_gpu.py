"""GPU parity: rk_adh_mode_init / rk_adh_update (the arm's homing modes) vs the oracle port and the golden
fixture made by the compiled reference modes; then the handover to the positioning mode.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import oracle_lib as ol
from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams
from roboken_fmskf_robot_controller_b200.arm import ArmBatch, ArmHomingBatch
from test_armhome_cpu import feedback, run, start_states

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_run(mode, s0, n, K, now, chunks=None, want_trace=True):
    ab = ArmBatch(n, DEV)
    ab.load_state_soa(s0)
    hb = ArmHomingBatch(ab)
    hb.mode_init(mode)
    now_d = None if now is None else torch.from_numpy(now).to(DEV)
    tr = torch.zeros((K, 16, n), dtype=torch.int32, device=DEV) if want_trace else None
    k0 = 0
    for k1 in (chunks or (K,)):
        hb.update(k1 - k0, now=None if now_d is None else now_d[k0:k1].contiguous(), trace=None if tr is None else tr[k0:k1])
        k0 = k1
    torch.cuda.synchronize()
    return (ab.state.cpu().numpy().view(np.uint32), hb.hstate.cpu().numpy().view(np.uint32),
            tr.cpu().numpy().view(np.uint32) if want_trace else None), ab


def test_homing_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "armhome_golden.npz"))
    n, K = 24, 1300
    for mode, tag in ((_cabi.RK_ADH_MODE_INIT, "init"), (_cabi.RK_ADH_MODE_INIT_POS_MOVE, "ipm")):
        (st, hs, tr), _ = gpu_run(mode, start_states(n, seed=0x5EED + mode), n, K, feedback(n, K, seed=0x5EED))
        np.testing.assert_array_equal(st, g[tag + "_state"])
        np.testing.assert_array_equal(hs, g[tag + "_hstate"])
        np.testing.assert_array_equal(tr[::13], g[tag + "_trace"])


@pytest.mark.parametrize("mode", [_cabi.RK_ADH_MODE_INIT, _cabi.RK_ADH_MODE_INIT_POS_MOVE])
@pytest.mark.parametrize("n,fb,zero", [(1, False, True), (700, True, False), (1031, False, False)])
def test_homing_vs_port(mode, n, fb, zero):
    K = 1400
    s0 = start_states(n, seed=mode * 100 + n, zero=zero)
    now = feedback(n, K, seed=n) if fb else None
    exp = run("port", mode, s0, n, K, now)
    got, _ = gpu_run(mode, s0, n, K, now)
    for x, y, nm in zip(got, exp, ("state", "mode block", "trace")):
        np.testing.assert_array_equal(x, y, err_msg=nm)
    # resumed in several launches, no trace
    got2, _ = gpu_run(mode, s0, n, K, now, chunks=(1, 2, 103, 104, 610, 1400), want_trace=False)
    np.testing.assert_array_equal(got2[0], exp[0])
    np.testing.assert_array_equal(got2[1], exp[1])


def test_homing_then_positioning_sequence():
    """Power-on -> INIT homing until COMPLETED -> ADTModePositioningSeq on the homed joints (offsets, limits and flags as
    the homing left them): device == port, end to end."""
    n, K = 300, 1000
    s0 = start_states(n, 0, zero=True)
    seq = layout.aos_to_soa(streams.arm_sequences(n, seed=8, seq_id=3, max_len=5))
    # port
    st, hs = s0.copy(), np.zeros(layout.HS_WORDS * n, dtype=np.uint32)
    ol.arm_homing("port", "init", st, hs, n, mode=_cabi.RK_ADH_MODE_INIT)
    ol.arm_homing("port", "update", st, hs, n, K=K)
    assert ((layout.soa_to_aos(hs, n, layout.HS_WORDS)[:, 0] >> 9) & 1).all()
    a = layout.soa_to_aos(st, n, layout.AS_WORDS)
    a[:, layout.AS_FSM] = layout.ASTATE_STANDBY | layout.AS_FSM_FIRSTCALL  # ADTModePositioningSeq::doInit
    a[:, layout.AS_SEQ_IDX] = (layout.ACMD_SLOTS - 1) | ((layout.ACMD_SLOTS - 1) << 16)
    st = layout.aos_to_soa(a)
    tab = np.zeros(layout.ACMD_WORDS * n, dtype=np.uint32)
    ol.arm_batch("port", "push", st, tab, n, seq=seq)
    tr, _ = ol.arm_batch("port", "update", st, tab, n, K=400, trace=True)
    # device
    ab = ArmBatch(n, DEV)
    ab.load_state_soa(s0)
    hb = ArmHomingBatch(ab)
    hb.mode_init(_cabi.RK_ADH_MODE_INIT)
    hb.update(K)
    g = layout.soa_to_aos(ab.state.cpu().numpy().view(np.uint32), n, layout.AS_WORDS)
    g[:, layout.AS_FSM] = layout.ASTATE_STANDBY | layout.AS_FSM_FIRSTCALL
    g[:, layout.AS_SEQ_IDX] = (layout.ACMD_SLOTS - 1) | ((layout.ACMD_SLOTS - 1) << 16)
    ab.load_state_soa(layout.aos_to_soa(g))
    ab.push_cmdseq(torch.from_numpy(seq.view(np.int32)).to(DEV))
    trd = torch.zeros((400, 16, n), dtype=torch.int32, device=DEV)
    ab.update(400, trace=trd)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(trd.cpu().numpy().view(np.uint32), tr)
    np.testing.assert_array_equal(ab.state.cpu().numpy().view(np.uint32), st)


def test_homing_argument_errors():
    lib = _cabi.load()
    ab = ArmBatch(4, DEV)
    hb = ArmHomingBatch(ab)
    assert lib.rk_adh_mode_init(hb.hstate.data_ptr(), 4, 7, None) == 1
    assert lib.rk_adh_mode_init(None, 4, 1, None) == 1
    import ctypes as C
    assert lib.rk_adh_update(C.byref(ab.params), ab.state.data_ptr(), None, 4, 1, None, None, None) == 1
    assert lib.rk_adh_update(C.byref(ab.params), ab.state.data_ptr(), hb.hstate.data_ptr(), 4, -1, None, None, None) == 1


def test_arm_handle_homing_equals_port():
    """rk_adt_t: rk_adt_home_init / rk_adt_home_tick (host feedback, isCompleted) == the port, tick by tick."""
    from roboken_fmskf_robot_controller_b200.arm import Arm

    K = 1000
    now = feedback(1, K, seed=2)
    s0 = start_states(1, 0, zero=True)
    st, hs = s0.copy(), np.zeros(layout.HS_WORDS, dtype=np.uint32)
    ol.arm_homing("port", "init", st, hs, 1, mode=_cabi.RK_ADH_MODE_INIT)
    arm = Arm()
    arm.home_init(_cabi.RK_ADH_MODE_INIT)
    done_at = None
    for t in range(K):
        ol.arm_homing("port", "update", st, hs, 1, K=1, now=np.ascontiguousarray(now[t : t + 1]))
        done = arm.home_update(now[t, :, 0])
        assert done == bool(hs[0] & layout.AS_FSM_IS_COMP), t
        if done and done_at is None:
            done_at = t
        if t % 97 == 0 or t == K - 1:
            np.testing.assert_array_equal(arm.get_state(), st)
    assert done_at is not None and done_at > 600
    arm.close()
