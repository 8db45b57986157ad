"""Host side of the on-device stream generators (rk_stream_*, csrc/rk_stream.cu): the descriptor lives in pinned host
memory, travels to the device as 48 bytes, and the kernels expand it into the blocks the engine consumes.  The streams
themselves are defined in streams.py (`*_v2`)."""
import ctypes as C

import torch

from . import _cabi, layout


class DeviceStreams:
    def __init__(self, device="cuda:0", **fields):
        self.lib = _cabi.load()
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.host = torch.zeros(C.sizeof(_cabi.StreamDesc), dtype=torch.uint8).pin_memory()
        self.desc = _cabi.StreamDesc.from_address(self.host.data_ptr())  # a view of the pinned bytes
        self.lib.rk_stream_default_desc(C.byref(self.desc))
        for k, v in fields.items():
            setattr(self.desc, k, int(v))
        with torch.cuda.device(self.dev_index):
            self.dev = torch.zeros(C.sizeof(_cabi.StreamDesc), dtype=torch.uint8, device=self.device)
        self.upload()

    NBYTES = C.sizeof(_cabi.StreamDesc)

    def set(self, **fields):
        for k, v in fields.items():
            setattr(self.desc, k, int(v))
        return self

    def upload(self, stream=None):
        """H2D of the descriptor (48 bytes, pinned -> device), asynchronous on `stream`."""
        with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(self.dev_index)):
            self.dev.copy_(self.host, non_blocking=True)

    def _st(self, stream):
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        return C.c_void_p(st.cuda_stream)

    def vehicle_commands(self, out, stream=None):
        """out: int32 [n_seg, n, 4]"""
        n_seg, n = out.shape[0], out.shape[1]
        _cabi.check(self.lib.rk_stream_vehicle_commands(self.dev.data_ptr(), n, n_seg, out.data_ptr(), self._st(stream)))
        return out

    def vehicle_yaw_reg(self, out, stream=None):
        """out: int16 [n_yaw, n]"""
        _cabi.check(self.lib.rk_stream_vehicle_yaw_reg(self.dev.data_ptr(), out.shape[1], out.shape[0], out.data_ptr(), self._st(stream)))
        return out

    def imu_samples(self, regs, have=None, stream=None, yaw_reg=None):
        """regs: int16 [n_upd, 2, n, 8]; have: uint8 [n_upd, n] or None; yaw_reg: int16 [n_upd, n] or None (receives the Yaw
        register column, what rk_tick_rollout_t::d_yaw_reg takes)"""
        n_upd, n = regs.shape[0], regs.shape[2]
        assert yaw_reg is None or (yaw_reg.dtype == torch.int16 and tuple(yaw_reg.shape) == (n_upd, n) and yaw_reg.is_contiguous())
        _cabi.check(self.lib.rk_stream_imu_samples_yaw(self.dev.data_ptr(), n, n_upd, regs.data_ptr(),
                                                       None if have is None else have.data_ptr(),
                                                       None if yaw_reg is None else yaw_reg.data_ptr(), self._st(stream)))
        return regs, have

    def imu_columns(self, yaw_reg, have=None, stream=None):
        """Only the columns of the IMU stream a vehicle rollout reads: yaw_reg int16 [n_upd, n] (the Yaw register), have
        uint8 [n_upd, n] or None -- for RobotBatch.make_args(imu_desc=self), whose IMU update draws the samples itself."""
        n_upd, n = yaw_reg.shape
        assert yaw_reg.dtype == torch.int16 and yaw_reg.is_contiguous()
        _cabi.check(self.lib.rk_stream_imu_samples_yaw(self.dev.data_ptr(), n, n_upd, None, None if have is None else have.data_ptr(),
                                                       yaw_reg.data_ptr(), self._st(stream)))
        return yaw_reg, have

    def arm_sequences(self, out, stream=None):
        """out: int32 [ACMD_SLOT_WORDS * n] (65 planes of 128-bit cells)"""
        n = out.numel() // layout.ACMD_SLOT_WORDS
        _cabi.check(self.lib.rk_stream_arm_sequences(self.dev.data_ptr(), n, out.data_ptr(), self._st(stream)))
        return out
