"""Full controller tick (BASELINE configs[4]): N robots = vehicle + IMU + arm, coupled through the
IMU yaw exactly as the firmware couples them (VD_task_main.cpp:368).  Thin host side over
rk_tick_rollout(); torch owns the HBM."""
import ctypes as C

import torch

from . import _cabi, layout
from .arm import ArmBatch
from .imu import ImuBatch
from .vehicle import VehicleBatch


class RobotBatch:
    def __init__(self, n, device="cuda:0", vdt_params=None, adt_params=None, arm_cmdtab=None):
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.vehicle = VehicleBatch(n, device, vdt_params)
        self.imu = ImuBatch(n, device)
        self.arm = ArmBatch(n, device, adt_params, cmdtab=arm_cmdtab)
        self.dev_index = self.vehicle.dev_index
        self._keep = []

    def reset(self):
        """Power-on state for all three sub-systems, then the arm's mode bring-up."""
        self.vehicle.state.zero_()
        self.imu.state.zero_()
        self.arm.state.zero_()
        self.arm.mode_init()

    def make_args(self, steps, slow_period, cmd, seg_len, regs, have_quat, yaw, goal=None, cost=None, vdt_trace=None,
                  adt_trace=None, reset_vehicle=False, yaw_reg=None, imu_desc=None):
        """cmd: [n_seg, n, 4] rk_vdt_cmd_t records; regs: int16 [n_slow, 2, n, 8] (streams.imu_cells); have_quat: uint8
        [n_slow, n] or None; yaw: float32 scratch, >= n words (receives the IMU yaw as it was at launch)."""
        n_slow = (steps + slow_period - 1) // slow_period
        # imu_desc: a DeviceStreams whose IMU stream the update draws in registers (regs may then be None; yaw_reg / have_quat are
        # the columns DeviceStreams.imu_columns wrote)
        assert imu_desc is not None or regs is not None
        if regs is not None:
            assert regs.is_cuda and regs.dtype == torch.int16 and tuple(regs.shape) == (n_slow, 2, self.n, 8) and regs.is_contiguous()
        if imu_desc is not None:
            assert yaw_reg is not None
        assert yaw.is_cuda and yaw.dtype == torch.float32 and yaw.numel() >= self.n
        a = _cabi.TickRollout()
        a.steps, a.slow_period = int(steps), int(slow_period)
        a.reset_vehicle = 1 if reset_vehicle else 0  # the vehicles start from the power-on block
        if cmd is not None:
            assert cmd.is_cuda and cmd.is_contiguous() and cmd.shape[1] == self.n
            a.d_cmd, a.n_seg, a.seg_len = cmd.data_ptr(), cmd.shape[0], int(seg_len)
        a.d_regs = None if regs is None else regs.data_ptr()
        a.d_imu_desc = None if imu_desc is None else imu_desc.dev.data_ptr()
        a.d_have_quat = None if have_quat is None else have_quat.data_ptr()
        a.d_yaw = yaw.data_ptr()
        if yaw_reg is not None:  # the Yaw register column of regs, int16 [n_slow, n]
            assert yaw_reg.is_cuda and yaw_reg.dtype == torch.int16 and tuple(yaw_reg.shape) == (n_slow, self.n) and yaw_reg.is_contiguous()
            a.d_yaw_reg = yaw_reg.data_ptr()
        if goal is not None and cost is not None:
            a.d_goal, a.d_cost = goal.data_ptr(), cost.data_ptr()
        if vdt_trace is not None:
            a.d_vdt_trace = vdt_trace.data_ptr()
        if adt_trace is not None:
            a.d_adt_trace = adt_trace.data_ptr()
        self._keep = [cmd, regs, have_quat, yaw, goal, cost, vdt_trace, adt_trace, yaw_reg, imu_desc]
        return a

    def rollout_args(self, args, stream=None):
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_tick_rollout(C.byref(self.vehicle.params), C.byref(self.arm.params),
                                             self.vehicle.state.data_ptr(), self.imu.state.data_ptr(),
                                             self.arm.state.data_ptr(), self.arm.cmdtab.data_ptr(), self.n, C.byref(args),
                                             C.c_void_p(st.cuda_stream)))

    def rollout(self, steps, slow_period, **kw):
        stream = kw.pop("stream", None)
        self.rollout_args(self.make_args(steps, slow_period, **kw), stream)
