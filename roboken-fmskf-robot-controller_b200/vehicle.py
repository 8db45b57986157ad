"""Host-side mirror of the vehicle entry points over the C-ABI, with torch owning the HBM.

torch is plumbing only (device buffers, streams); every number is produced by the CUDA
kernels behind librobotick_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi, layout


def _dev_index(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise _cabi.RobotickError("robotick state blocks live in GPU memory (device must be cuda:N)")
    return d.index if d.index is not None else torch.cuda.current_device()


class VehicleBatch:
    """N vehicles' state in HBM (SoA of 128-bit planes) + the fused rollout.

    The all-zero block is the firmware's power-on state (static zero-initialisation).
    """

    def __init__(self, n, device="cuda:0", params=None):
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.dev_index = _dev_index(device)
        self.params = params or _cabi.default_params()
        assert self.lib.rk_vdt_state_words() == layout.VS_WORDS
        with torch.cuda.device(self.dev_index):
            self.state = torch.zeros(layout.VS_WORDS * self.n, dtype=torch.int32, device=self.device)
        self._keep = []

    # ---- state access (host tooling / tests) ------------------------------------------
    def state_aos(self):
        return layout.soa_to_aos(self.state.cpu().numpy().view(np.uint32), self.n, layout.VS_WORDS)

    def load_state_aos(self, aos):
        soa = layout.aos_to_soa(aos).view(np.int32)
        self.state.copy_(torch.from_numpy(soa))

    def load_state_soa(self, soa_u32):
        self.state.copy_(torch.from_numpy(np.asarray(soa_u32).view(np.int32)))

    def reset(self):
        self.state.zero_()

    # ---- the hot path -------------------------------------------------------------------
    def make_args(self, steps, sensor_mode=_cabi.RK_SENSOR_PLANT, cmd=None, seg_len=0, yaw=None, yaw_period=0,
                  frames=None, trace=None, goal=None, cost=None, task_period=0, reset_state=False):
        """Device tensors -> rk_vdt_rollout_t.  cmd: int32/float32 [n_seg, n, 4] (rk_vdt_cmd_t
        records), yaw: float32 [n_yaw, n] radians or int16 [n_yaw, n] WT901C Yaw register counts, frames: int64 [steps, 4, n], trace: int32
        [steps, 16, n], goal: float32 [n, 2], cost: float32 [n]."""
        a = _cabi.VdtRollout()
        a.steps, a.sensor_mode = int(steps), int(sensor_mode)
        a.reset_state = 1 if reset_state else 0  # start from the power-on block instead of the contents of self.state
        a.task_period = int(task_period)  # > 0: VDT::main runs every task_period ticks (RK_CMD_MSG_* records, move-time auto-stop)
        if cmd is not None:
            assert cmd.is_cuda and cmd.is_contiguous() and cmd.shape[1] == self.n and cmd.element_size() * cmd.shape[-1] == 16
            a.d_cmd, a.n_seg, a.seg_len = cmd.data_ptr(), cmd.shape[0], int(seg_len)
        if yaw is not None:
            assert yaw.is_cuda and yaw.is_contiguous() and yaw.dtype in (torch.float32, torch.int16) and yaw.shape[1] == self.n
            a.n_yaw, a.yaw_period = yaw.shape[0], int(yaw_period)
            if yaw.dtype == torch.float32:
                a.d_yaw = yaw.data_ptr()  # radians
            else:
                a.d_yaw_reg = yaw.data_ptr()  # the WT901C Yaw register (180/32768 deg per count)
        if frames is not None:
            assert frames.is_cuda and frames.is_contiguous() and frames.dtype == torch.int64
            assert tuple(frames.shape) == (steps, 4, self.n)
            a.d_frames = frames.data_ptr()
        if trace is not None:
            assert trace.is_cuda and trace.is_contiguous() and tuple(trace.shape) == (steps, _cabi.RK_VDT_TRACE_WORDS, self.n)
            a.d_trace = trace.data_ptr()
        if goal is not None and cost is not None:
            assert goal.is_cuda and cost.is_cuda and goal.dtype == torch.float32 and cost.dtype == torch.float32
            a.d_goal, a.d_cost = goal.data_ptr(), cost.data_ptr()
        self._keep = [cmd, yaw, frames, trace, goal, cost]
        return a

    def rollout_args(self, args, stream=None):
        """rk_vdt_rollout() on the current (or given) torch stream; asynchronous."""
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_vdt_rollout(C.byref(self.params), self.state.data_ptr(), self.n, C.byref(args),
                                            C.c_void_p(st.cuda_stream)))

    def rollout(self, steps, **kw):
        stream = kw.pop("stream", None)
        self.rollout_args(self.make_args(steps, **kw), stream)

    # ---- batched setters ------------------------------------------------------------------
    def set_power(self, on=None):
        st = torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_vdt_set_power(self.state.data_ptr(), self.n, None if on is None else on.data_ptr(),
                                              C.c_void_p(st.cuda_stream)))

    def set_target_vel(self, v, a, j):
        """v, a, j: float32 device tensors [3, n]  (VEHICLE_CTRL::set_target_vel)"""
        st = torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_vdt_set_target_vel(C.byref(self.params), self.state.data_ptr(), self.n, v.data_ptr(),
                                                   a.data_ptr(), j.data_ptr(), C.c_void_p(st.cuda_stream)))

    def tx_frames(self, out=None):
        """CAN_CTRL::tx_routine for every vehicle: int64 [n] device tensor of C610 frames (four big-endian s16 currents)."""
        st = torch.cuda.current_stream(self.dev_index)
        if out is None:
            with torch.cuda.device(self.dev_index):
                out = torch.empty(self.n, dtype=torch.int64, device=self.device)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_vdt_tx_frames(self.state.data_ptr(), self.n, out.data_ptr(), C.c_void_p(st.cuda_stream)))
        return out

    def motor_rx(self, wheel, frames, usec=None):
        """frames: int64 [n] device tensor of 8-byte M2006 frames (MOTOR_IF_M2006::rx_callback)"""
        st = torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_vdt_motor_rx(C.byref(self.params), self.state.data_ptr(), self.n, int(wheel),
                                             frames.data_ptr(), None if usec is None else usec.data_ptr(),
                                             C.c_void_p(st.cuda_stream)))


class Vehicle:
    """Single-instance handle (rk_vdt_t): the drop-in for the static objects of
    VD_task_main.cpp:75-108 -- a batch of one on the same kernels."""

    def __init__(self, params=None):
        self.lib = _cabi.load()
        self.h = C.c_void_p()
        _cabi.check(self.lib.rk_vdt_create(C.byref(self.h), C.byref(params) if params else None))

    def close(self):
        if self.h:
            self.lib.rk_vdt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def start(self):
        _cabi.check(self.lib.rk_vdt_start(self.h))

    def stop(self):
        _cabi.check(self.lib.rk_vdt_stop(self.h))

    def update(self):
        _cabi.check(self.lib.rk_vdt_update(self.h))

    def set_target_vel(self, v, a, j):
        F3 = C.c_float * 3
        _cabi.check(self.lib.rk_vdt_set_target(self.h, F3(*v), F3(*a), F3(*j)))

    def set_now_yaw_world(self, yaw_rad):
        _cabi.check(self.lib.rk_vdt_set_yaw(self.h, C.c_float(yaw_rad)))

    def rx_callback(self, wheel, frame_bytes, usec_id):
        _cabi.check(self.lib.rk_vdt_rx(self.h, wheel, bytes(frame_bytes), usec_id))

    def _get3(self, fn):
        out = (C.c_float * 3)()
        _cabi.check(fn(self.h, out))
        return np.array(out[:], dtype=np.float32)

    def get_vehicle_pos_m_latest(self):
        return self._get3(self.lib.rk_vdt_get_pos)

    def get_vehicle_vel_mmps_latest(self):
        return self._get3(self.lib.rk_vdt_get_vel)

    def get_vehicle_vel_tgt_mmps_latest(self):
        return self._get3(self.lib.rk_vdt_get_vel_tgt)

    def get_rawCurr_tgt(self):
        out = (C.c_int16 * 4)()
        _cabi.check(self.lib.rk_vdt_get_raw_current(self.h, out))
        return list(out)

    def tx_routine(self):
        """CAN_CTRL::tx_routine (VD_can_controller.hpp:43-55): the 8 bytes of the C610 current frame (id 0x200)."""
        out = (C.c_uint8 * 8)()
        _cabi.check(self.lib.rk_vdt_get_tx_frame(self.h, out))
        return bytes(out)

    def get_rawAngleSum(self):
        out = (C.c_int64 * 4)()
        _cabi.check(self.lib.rk_vdt_get_angle_sum(self.h, out))
        return list(out)

    def get_state(self):
        out = (C.c_uint32 * layout.VS_WORDS)()
        _cabi.check(self.lib.rk_vdt_get_state(self.h, out))
        return np.array(out[:], dtype=np.uint32)

    def set_state(self, words):
        arr = (C.c_uint32 * layout.VS_WORDS)(*[int(x) for x in words])
        _cabi.check(self.lib.rk_vdt_set_state(self.h, arr))
