"""Host-side mirror of the RobotManager guard (src/RobotManager/RM_task_main.cpp routine_ros) over the C-ABI."""
import ctypes as C

import torch

from . import _cabi, layout


class ManagerBatch:
    """N RobotManager instances: NOW_CMD_STATUS, IS_IGNORE_FLOOR_DETECTION, U32_MCN_NO_CMD_CNT, vdt_abort (one plane)."""

    def __init__(self, n, device="cuda:0", params=None):
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        assert self.lib.rk_rmt_state_words() == layout.RS_WORDS
        self.params = params or _cabi.RmtParams()
        if params is None:
            self.lib.rk_rmt_default_params(C.byref(self.params))
        with torch.cuda.device(self.dev_index):
            self.state = torch.zeros(layout.RS_WORDS * self.n, dtype=torch.int32, device=self.device)

    def guard(self, inp, cmd_out, abort_out=None, stream=None):
        """inp: int32 [K, 3, n, 4] (RK_RI_* records); cmd_out: int32 [K, n, 4] -- the rk_vdt_cmd_t records VDT receives,
        directly usable as VehicleBatch.rollout(cmd=...); abort_out: int32 [K, n] (vdt_abort.val) or None."""
        assert inp.is_cuda and inp.dtype == torch.int32 and inp.is_contiguous() and tuple(inp.shape[1:]) == (3, self.n, 4)
        K = int(inp.shape[0])
        assert cmd_out.is_cuda and cmd_out.dtype == torch.int32 and cmd_out.is_contiguous() and tuple(cmd_out.shape) == (K, self.n, 4)
        if abort_out is not None:
            assert abort_out.is_cuda and abort_out.dtype == torch.int32 and tuple(abort_out.shape) == (K, self.n)
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        _cabi.check(self.lib.rk_rmt_guard(C.byref(self.params), self.state.data_ptr(), self.n, K, inp.data_ptr(), cmd_out.data_ptr(),
                                          None if abort_out is None else abort_out.data_ptr(), C.c_void_p(st.cuda_stream)))


class Manager:
    """Single RobotManager (rk_rmt_t): the file-static state of RM_task_main.cpp behind one handle; one
    routine_ros() vehicle-management block per cycle()."""

    def __init__(self, params=None):
        self.lib = _cabi.load()
        self.h = C.c_void_p()
        _cabi.check(self.lib.rk_rmt_create(C.byref(self.h), None if params is None else C.byref(params)))

    def close(self):
        if self.h:
            self.lib.rk_rmt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cycle(self, record):
        """record: 12 uint32 words (RK_RI_*).  Returns (rk_vdt_cmd_t as 4 uint32 words, vdt_abort.val)."""
        rec = (C.c_uint32 * layout.RI_WORDS)(*[int(x) for x in record])
        out, ab = _cabi.VdtCmd(), C.c_uint32()
        _cabi.check(self.lib.rk_rmt_cycle(self.h, rec, C.byref(out), C.byref(ab)))
        return list((C.c_uint32 * 4).from_buffer_copy(out)), ab.value

    def state(self):
        w = (C.c_uint32 * layout.RS_WORDS)()
        _cabi.check(self.lib.rk_rmt_get_state(self.h, w))
        return list(w)


def atan2f(y, x, stream=None):
    """UTIL::mymath::atan2f (table arctangent, src/Utility/util_mymath.cpp:98-126) on float32 CUDA tensors."""
    lib = _cabi.load()
    assert y.is_cuda and x.is_cuda and y.dtype == torch.float32 and x.dtype == torch.float32 and y.shape == x.shape
    y, x = y.contiguous(), x.contiguous()
    out = torch.empty_like(y)
    dev = y.device.index
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    _cabi.check(lib.rk_set_device(dev))
    _cabi.check(lib.rk_mymath_atan2f(y.data_ptr(), x.data_ptr(), out.data_ptr(), y.numel(), C.c_void_p(st.cuda_stream)))
    return out
