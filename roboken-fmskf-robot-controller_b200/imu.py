"""Host-side mirror of the IMU entry points (src/Imu) over the C-ABI; torch owns the HBM."""
import ctypes as C

import numpy as np
import torch

from . import _cabi, layout


class ImuBatch:
    """N WT901C interfaces: q_init + readable Data page + is_error, SoA of 128-bit planes."""

    def __init__(self, n, device="cuda:0"):
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        assert self.lib.rk_imt_state_words() == layout.IS_WORDS
        with torch.cuda.device(self.dev_index):
            self.state = torch.zeros(layout.IS_WORDS * self.n, dtype=torch.int32, device=self.device)
        self.parser = None  # WIT serial parser block (RK_IP_*), allocated by the first feed_bytes()

    def update(self, regs, have_quat=None, out=None, do_init=False, stream=None):
        """regs: int16 [K, 2, n, 8] (two 128-bit cells per sample, streams.imu_cells); have_quat: uint8 [K, n] or None; out: float32 [K, 4, n, 4] or
        None (IMU_IF_WT901C::update / ::init when do_init)."""
        assert regs.is_cuda and regs.dtype == torch.int16 and regs.is_contiguous() and tuple(regs.shape[1:]) == (2, self.n, 8)
        K = regs.shape[0]
        if have_quat is not None:
            assert have_quat.is_cuda and have_quat.dtype == torch.uint8 and tuple(have_quat.shape) == (K, self.n)
        if out is not None:
            assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (K, 4, self.n, 4)
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_imt_update(self.state.data_ptr(), self.n, K, regs.data_ptr(),
                                           None if have_quat is None else have_quat.data_ptr(),
                                           None if out is None else out.data_ptr(), int(bool(do_init)),
                                           C.c_void_p(st.cuda_stream)))

    def feed_bytes(self, cells, nbytes=None, out=None, yaw_rad=None, do_init=False, stream=None):
        """cells: int32 [K, ncells, n, 4] serial bytes in 128-bit cells (wire order from the low byte up); nbytes:
        int16/uint16 [K, n] bytes really on the wire per update, or None (full slots).  Runs the vendor parser's
        state machine (lib/wt901c/wit_c_sdk.c:132-164), then IMU_IF_WT901C::update per update (::init for the
        first when do_init).  out: float32 [K, 4, n, 4] or None; yaw_rad: float32 [K, n] or None."""
        assert cells.is_cuda and cells.dtype in (torch.int32, torch.uint32) and cells.is_contiguous() and cells.dim() == 4
        assert cells.shape[2] == self.n and cells.shape[3] == 4
        K, ncells = int(cells.shape[0]), int(cells.shape[1])
        if nbytes is not None:
            assert nbytes.is_cuda and nbytes.dtype in (torch.int16, torch.uint16) and nbytes.is_contiguous()
            assert tuple(nbytes.shape) == (K, self.n)
        if out is not None:
            assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (K, 4, self.n, 4)
        if yaw_rad is not None:
            assert yaw_rad.is_cuda and yaw_rad.dtype == torch.float32 and tuple(yaw_rad.shape) == (K, self.n)
        if self.parser is None:
            assert self.lib.rk_imt_parser_words() == layout.IP_WORDS
            with torch.cuda.device(self.dev_index):
                self.parser = torch.zeros(layout.IP_WORDS * self.n, dtype=torch.int32, device=self.device)
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_imt_feed_bytes(self.state.data_ptr(), self.parser.data_ptr(), self.n, K, ncells, cells.data_ptr(),
                                               None if nbytes is None else nbytes.data_ptr(),
                                               None if out is None else out.data_ptr(),
                                               None if yaw_rad is None else yaw_rad.data_ptr(), int(bool(do_init)),
                                               C.c_void_p(st.cuda_stream)))

    def state_aos(self):
        return layout.soa_to_aos(self.state.cpu().numpy().view(np.uint32), self.n, layout.IS_WORDS)

    def load_state_soa(self, soa_u32):
        self.state.copy_(torch.from_numpy(np.asarray(soa_u32).view(np.int32)))


class Imu:
    """Single-instance handle (rk_imt_t) mirroring IMU_IF (imu_if_base.hpp:20-29)."""

    def __init__(self):
        self.lib = _cabi.load()
        self.h = C.c_void_p()
        _cabi.check(self.lib.rk_imt_create(C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.rk_imt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, regs):
        _cabi.check(self.lib.rk_imt_init(self.h, (C.c_int16 * 16)(*[int(x) for x in regs])))

    def update(self, regs, have_quat=True):
        _cabi.check(self.lib.rk_imt_update1(self.h, (C.c_int16 * 16)(*[int(x) for x in regs]), int(have_quat)))

    def getDataLatest(self):
        d, e = (C.c_float * 16)(), C.c_int()
        _cabi.check(self.lib.rk_imt_get(self.h, d, C.byref(e)))
        return np.array(d[:], dtype=np.float32), bool(e.value)

    def isError(self):
        return self.getDataLatest()[1]

    def getYawDate(self):
        y = C.c_float()
        _cabi.check(self.lib.rk_imt_get_yaw(self.h, C.byref(y)))
        return y.value
