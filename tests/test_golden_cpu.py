"""CPU: the plain-C oracle reproduces the committed golden fixtures (generated from the
unmodified reference by tests/golden/make_golden.py)."""
import os

import numpy as np

import oracle_lib as ol
import workloads as wl
from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "vdt_golden.npz"))


def _port(inp, sensor=_cabi.RK_SENSOR_PLANT, frames=None, state=None):
    n = inp["n"]
    st = np.zeros(layout.VS_WORDS * n, dtype=np.uint32) if state is None else state.copy()
    ro = ol.HostRollout(n, inp["steps"], sensor, inp.get("cmd"), inp.get("seg_len", 0), inp.get("yaw"),
                        inp.get("yaw_period", 0), frames=frames, trace=True)
    ol.run_port(st, n, ro)
    return st, ro.trace


def test_golden_c1():
    st, tr = _port(wl.c1_inputs())
    np.testing.assert_array_equal(tr[G["c1_rows"]], G["c1_trace"])
    np.testing.assert_array_equal(st, G["c1_state"])


def test_golden_plant():
    st, tr = _port(wl.plant_inputs(16, 1000, seed=0x5EED))
    np.testing.assert_array_equal(tr[::100], G["plant_trace"])
    np.testing.assert_array_equal(tr[-1], G["plant_last"])
    np.testing.assert_array_equal(st, G["plant_state"])


def test_golden_stream():
    fr = streams.vehicle_frames(8, 300, seed=21)
    st, tr = _port(wl.plant_inputs(8, 300, seed=21), sensor=_cabi.RK_SENSOR_STREAM, frames=fr)
    np.testing.assert_array_equal(tr[::30], G["stream_trace"])
    np.testing.assert_array_equal(st, G["stream_state"])


def test_golden_random_states():
    st0 = layout.aos_to_soa(wl.random_states(64, seed=9))
    st, tr = _port(wl.plant_inputs(64, 24, seed=9, seg_len=6, yaw_period=3), state=st0)
    np.testing.assert_array_equal(tr[::6], G["rand_trace"])
    np.testing.assert_array_equal(st, G["rand_state"])


def test_threads_do_not_change_results():
    inp = wl.plant_inputs(40, 300, seed=4)
    a, _ = _port(inp)
    n = inp["n"]
    st = np.zeros(layout.VS_WORDS * n, dtype=np.uint32)
    ro = ol.HostRollout(n, inp["steps"], _cabi.RK_SENSOR_PLANT, inp["cmd"], inp["seg_len"], inp["yaw"], inp["yaw_period"])
    ol.run_port(st, n, ro, nthreads=4)
    np.testing.assert_array_equal(st, a)
