#!/usr/bin/env python3
"""Generate tests/golden/{vdt,imu,arm}_golden.npz from oracle/_ref -- the reference's own sources
compiled unmodified for x86 (needs /root/reference or a prebuilt oracle/_ref).

    python tests/golden/make_golden.py [vdt] [imu] [arm]     (default: all)

The fixtures hold OUTPUTS only (decimated traces + final state blocks); inputs are
re-derived from seeds by tests/workloads.py, so the file stays small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib as ol  # noqa: E402
import workloads as wl  # noqa: E402
from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams  # noqa: E402


def run_ref(inp, sensor=_cabi.RK_SENSOR_PLANT, frames=None, state=None):
    n = inp["n"]
    st = np.zeros(layout.VS_WORDS * n, dtype=np.uint32) if state is None else state.copy()
    ro = ol.HostRollout(n, inp["steps"], sensor, inp.get("cmd"), inp.get("seg_len", 0), inp.get("yaw"),
                        inp.get("yaw_period", 0), frames=frames, trace=True)
    ol.run_ref(st, n, ro)
    return st, ro.trace


def c1_rows():
    return np.unique(np.concatenate([np.arange(10), np.arange(0, 10000, 50), [4999, 5000, 5001, 9999]]))


def arm_golden():
    """Arm: the reference's own POS_CMD_SEQ_DEBUG_0/1/2 (AD_mode_positioning_seq_debug_data.cpp)
    on arms 0..2 + seeded random sequences, through the compiled reference.  Inputs are stored
    too (they are small) so the fixture is self-contained."""
    n, K = 24, 420
    r = ol.ref("libref_arm.so")
    seq_a = streams.arm_sequences(n, seed=0x5EED, seq_id=1, max_len=6)
    seq_b = streams.arm_sequences(n, seed=0x5EED + 1, seq_id=2, max_len=4)
    for w in range(3):
        q = _cabi.AdtPosCmdSeq()
        r.ref_adt_debug_seq(w, q)
        seq_a[w] = ol.seq_struct_to_image(q)
        seq_a[w, 0] = 1
    valid_b = (np.arange(n) % 4 != 1).astype(np.uint8)
    st = np.zeros(layout.AS_WORDS * n, dtype=np.uint32)
    tb = np.zeros(layout.ACMD_WORDS * n, dtype=np.uint32)
    ol.arm_batch("ref", "init", st, tb, n)
    ol.arm_batch("ref", "push", st, tb, n, seq=layout.aos_to_soa(seq_a))
    ol.arm_batch("ref", "push", st, tb, n, seq=layout.aos_to_soa(seq_b), valid=valid_b)
    tr, _ = ol.arm_batch("ref", "update", st, tb, n, K=K, trace=True)
    ids = (np.arange(n) % 3).astype(np.uint32)
    _, status = ol.arm_batch("ref", "status", st, tb, n, ids=ids)
    path = os.path.join(HERE, "arm_golden.npz")
    np.savez_compressed(path, n=n, K=K, seq_a=seq_a, seq_b=seq_b, valid_b=valid_b, trace=tr, state=st, cmdtab=tb, ids=ids,
                        status=status)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    which = set(sys.argv[1:]) or {"vdt", "imu", "arm", "wire", "rmt", "armhome"}
    if "arm" in which:
        arm_golden()
    if "vdt" in which:
        vdt_golden()
    if "imu" in which:
        imu_golden()
    if "wire" in which:
        imu_wire_golden()
    if "rmt" in which:
        rmt_golden()
    if "armhome" in which:
        armhome_golden()


def vdt_golden():
    out = {}
    st, tr = run_ref(wl.c1_inputs())
    out["c1_rows"] = c1_rows()
    out["c1_trace"] = tr[out["c1_rows"]]
    out["c1_state"] = st
    inp = wl.plant_inputs(16, 1000, seed=0x5EED)
    st, tr = run_ref(inp)
    out["plant_trace"] = tr[::100]
    out["plant_last"] = tr[-1]
    out["plant_state"] = st
    inp = wl.plant_inputs(8, 300, seed=21)
    fr = streams.vehicle_frames(8, 300, seed=21)
    st, tr = run_ref(inp, sensor=_cabi.RK_SENSOR_STREAM, frames=fr)
    out["stream_trace"] = tr[::30]
    out["stream_state"] = st
    st0 = layout.aos_to_soa(wl.random_states(64, seed=9))
    inp = wl.plant_inputs(64, 24, seed=9, seg_len=6, yaw_period=3)
    st, tr = run_ref(inp, state=st0)
    out["rand_trace"] = tr[::6]
    out["rand_state"] = st
    path = os.path.join(HERE, "vdt_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def imu_golden():
    # IMU: compiled reference driven through WT901 serial frames
    n, K = 64, 32
    regs, have = streams.imu_samples(n, K, seed=0x5EED, drop_every=8)
    st = np.zeros(layout.IS_WORDS * n, dtype=np.uint32)
    o = ol.imu_ref(st, n, regs, have, want_out=True, do_init=True)
    path = os.path.join(HERE, "imu_golden.npz")
    np.savez_compressed(path, out=o, state=st)
    print("wrote", path, os.path.getsize(path), "bytes")


def imu_wire_golden():
    # WIT serial codec: adversarial byte streams through the unmodified vendor parser + IMU_IF_WT901C
    n, K, ncells = 48, 20, 2
    cells, nbytes = streams.imu_wire_fuzz(n, K, ncells=ncells, seed=0x5EED)
    st = np.zeros(layout.IS_WORDS * n, dtype=np.uint32)
    o, sreg = ol.imu_bytes_ref(st, n, cells, nbytes, want_out=True)
    path = os.path.join(HERE, "imu_wire_golden.npz")
    np.savez_compressed(path, cells=cells, nbytes=nbytes, out=o, state=st, sreg=sreg)
    print("wrote", path, os.path.getsize(path), "bytes")


def rmt_golden():
    # RobotManager guard: the unmodified RM_task_main.cpp routine_ros() + util_mymath.cpp arctangent
    n, K = 40, 460
    inp = streams.rm_inputs(n, K, seed=0x5EED)
    st = np.zeros(layout.RS_WORDS * n, dtype=np.uint32)
    cmd, ab = ol.rm_guard("ref", st, n, inp)
    rng = np.random.default_rng(7)
    y = np.concatenate([rng.standard_normal(300) * 300, [0, 0, 1, -1, 0.0, 5e5, -3, np.inf]]).astype(np.float32)
    x = np.concatenate([rng.standard_normal(300) * 300, [0, 1, 0, 0, -2.0, 1e-3, 1e7, 1.0]]).astype(np.float32)
    r = ol.ref("libref_rm.so")
    out = np.array([r.ref_rm_atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)
    path = os.path.join(HERE, "rmt_golden.npz")
    np.savez_compressed(path, cmd=cmd, abort=ab, state=st, atan_y=y, atan_x=x, atan_out=out)
    print("wrote", path, os.path.getsize(path), "bytes")


def armhome_golden():
    # arm homing modes: the unmodified ADTModeInitialize / ADTModeInitPosMove on the unmodified joints
    import test_armhome_cpu as th

    n, K = 24, 1300
    out = {}
    for mode, tag in ((_cabi.RK_ADH_MODE_INIT, "init"), (_cabi.RK_ADH_MODE_INIT_POS_MOVE, "ipm")):
        st, hs, tr = th.run("ref", mode, th.start_states(n, seed=0x5EED + mode), n, K, th.feedback(n, K, seed=0x5EED))
        out[tag + "_state"], out[tag + "_hstate"], out[tag + "_trace"] = st, hs, tr[::13]
    path = os.path.join(HERE, "armhome_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
