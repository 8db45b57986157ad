#!/usr/bin/env python3
"""Generate tests/golden/vdt_golden.npz from oracle/_ref -- the reference's own sources
compiled unmodified for x86 (needs /root/reference or a prebuilt oracle/_ref).

    python tests/golden/make_golden.py

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


def main():
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
    # IMU: compiled reference driven through WT901 serial frames
    n, K = 64, 32
    regs, have = streams.imu_samples(n, K, seed=0x5EED, drop_every=8)
    st = np.zeros(layout.IS_WORDS * n, dtype=np.uint32)
    o = ol.imu_ref(st, n, regs, have, want_out=True, do_init=True)
    path = os.path.join(HERE, "imu_golden.npz")
    np.savez_compressed(path, out=o, state=st)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
