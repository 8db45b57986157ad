"""Host-side mirror of the arm entry points (src/ArmDrive) over the C-ABI; torch owns the HBM.

ArmBatch  = N x {ADTModePositioningSeq + the seven joint objects of AD_task_main.cpp:108-116}
Arm       = one arm behind an rk_adt_t handle, method names as in the reference
            (push_cmdseq / update / get_q_cmdseq_status / get_tgt_deg).
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi, layout

PROCESSING, DONE, NO_DATA = 0, 1, 99  # ADTModePositioningSeq::CmdStatus  AD_mode_positioning_seq.hpp:36-40


class ArmBatch:
    def __init__(self, n, device="cuda:0", params=None, cmdtab=None):
        """cmdtab: an existing command-ring tensor to use instead of allocating one (rollout drivers
        that re-push every pass share one ring between batches they run back to back)."""
        self.lib = _cabi.load()
        self.n = int(n)
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.params = params or _cabi.default_arm_params()
        assert self.lib.rk_adt_state_words() == layout.AS_WORDS
        with torch.cuda.device(self.dev_index):
            self.state = torch.zeros(layout.AS_WORDS * self.n, dtype=torch.int32, device=self.device)
            if cmdtab is None:
                cmdtab = torch.zeros(layout.ACMD_WORDS * self.n, dtype=torch.int32, device=self.device)
            assert cmdtab.is_cuda and cmdtab.numel() == layout.ACMD_WORDS * self.n and cmdtab.element_size() == 4
            self.cmdtab = cmdtab

    def _st(self, stream):
        _cabi.check(self.lib.rk_set_device(self.dev_index))
        st = stream if stream is not None else torch.cuda.current_stream(self.dev_index)
        return C.c_void_p(st.cuda_stream)

    def mode_init(self, stream=None):
        """ADTModeBase::init() after a finished INIT mode (rk_adt_mode_init)."""
        _cabi.check(self.lib.rk_adt_mode_init(C.byref(self.params), self.state.data_ptr(), self.n, self._st(stream)))

    def push_cmdseq(self, seq_soa, valid=None, stream=None):
        """seq_soa: int32/uint32 device tensor, 260*n words in plane order; valid: uint8 [n] or None."""
        assert seq_soa.is_cuda and seq_soa.numel() == layout.ACMD_SLOT_WORDS * self.n and seq_soa.element_size() == 4
        if valid is not None:
            assert valid.is_cuda and valid.dtype == torch.uint8 and valid.numel() == self.n
        _cabi.check(self.lib.rk_adt_push_cmdseq(self.state.data_ptr(), self.cmdtab.data_ptr(), self.n, seq_soa.data_ptr(),
                                                None if valid is None else valid.data_ptr(), self._st(stream)))

    def update(self, K=1, trace=None, stream=None):
        """K fused 100 Hz ticks; trace: int32 device tensor [K, 16, n] or None."""
        if trace is not None:
            assert trace.is_cuda and trace.element_size() == 4 and trace.numel() == K * layout.ADT_TRACE_WORDS * self.n
        _cabi.check(self.lib.rk_adt_update(C.byref(self.params), self.state.data_ptr(), self.cmdtab.data_ptr(), self.n, int(K),
                                           None if trace is None else trace.data_ptr(), self._st(stream)))

    def bldc_rx(self, slot, frames, cmdid=None, cur=None, stream=None):
        """JointMyBldcServo::rx_callback for servo slot 0 DF_Left / 1 DF_Right / 2 P3; frames: int64 [n] device tensor,
        cmdid: int32 [n] or None (= status summary), cur: float32 [n] or None (receives fl_out_now_cur)."""
        assert frames.is_cuda and frames.dtype == torch.int64 and frames.numel() == self.n
        _cabi.check(self.lib.rk_adt_bldc_rx(C.byref(self.params), self.state.data_ptr(), self.n, int(slot), frames.data_ptr(),
                                            None if cmdid is None else cmdid.data_ptr(), None if cur is None else cur.data_ptr(),
                                            self._st(stream)))

    def mg_rx(self, frames, cur=None, stream=None):
        """JointMgServo::rx_callback; frames: int64 [n] device tensor, cur: float32 [n] or None."""
        assert frames.is_cuda and frames.dtype == torch.int64 and frames.numel() == self.n
        _cabi.check(self.lib.rk_adt_mg_rx(C.byref(self.params), self.state.data_ptr(), self.n, frames.data_ptr(),
                                          None if cur is None else cur.data_ptr(), self._st(stream)))

    def cmdseq_status(self, ids, stream=None):
        assert ids.is_cuda and ids.element_size() == 4 and ids.numel() == self.n
        out = torch.empty(self.n, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.rk_adt_cmdseq_status(self.state.data_ptr(), self.cmdtab.data_ptr(), self.n, ids.data_ptr(),
                                                  out.data_ptr(), self._st(stream)))
        return out

    def state_host(self):
        return self.state.cpu().numpy().view(np.uint32)

    def cmdtab_host(self):
        return self.cmdtab.cpu().numpy().view(np.uint32)

    def load_state_soa(self, soa_u32, tab_u32=None):
        self.state.copy_(torch.from_numpy(np.asarray(soa_u32).view(np.int32)))
        if tab_u32 is not None:
            self.cmdtab.copy_(torch.from_numpy(np.asarray(tab_u32).view(np.int32)))


class ArmPositioningBatch:
    """ADTModePositioning (the single-command mode behind REQ_MOVE_POS) for the arms of an ArmBatch:
    the joints are the ArmBatch's, the mode block (FIFO of <= 4 commands) lives here."""

    def __init__(self, arms):
        self.arms, self.lib, self.n = arms, arms.lib, arms.n
        assert self.lib.rk_adp_state_words() == layout.PS_WORDS
        with torch.cuda.device(arms.dev_index):
            self.pstate = torch.zeros(layout.PS_WORDS * self.n, dtype=torch.int32, device=arms.device)

    def mode_init(self, stream=None):
        _cabi.check(self.lib.rk_adp_mode_init(self.pstate.data_ptr(), self.n, self.arms._st(stream)))

    def push_cmd(self, cmd, valid=None, stream=None):
        """cmd: int32/uint32 device tensor [2, n, 4]: {id, dt_ms, tgt0, tgt1}, {tgt2, tgt3, tgt4, 0} per arm."""
        assert cmd.is_cuda and cmd.numel() == 8 * self.n and cmd.element_size() == 4 and cmd.is_contiguous()
        _cabi.check(self.lib.rk_adp_push_cmd(self.pstate.data_ptr(), self.n, cmd.data_ptr(),
                                             None if valid is None else valid.data_ptr(), self.arms._st(stream)))

    def update(self, K=1, trace=None, stream=None):
        _cabi.check(self.lib.rk_adp_update(C.byref(self.arms.params), self.arms.state.data_ptr(), self.pstate.data_ptr(), self.n,
                                           int(K), None if trace is None else trace.data_ptr(), self.arms._st(stream)))

    def cmd_status(self, ids, stream=None):
        out = torch.empty(self.n, dtype=torch.int32, device=self.arms.device)
        _cabi.check(self.lib.rk_adp_cmd_status(self.pstate.data_ptr(), self.n, ids.data_ptr(), out.data_ptr(), self.arms._st(stream)))
        return out


class ArmHomingBatch:
    """The homing modes ADTModeInitialize (mode RK_ADH_MODE_INIT) and ADTModeInitPosMove (RK_ADH_MODE_INIT_POS_MOVE) for
    the arms of an ArmBatch: the joints are the ArmBatch's, the mode block (state, wait counter, ramp directions) lives
    here.  isCompleted() of the firmware = (hstate word 0 >> 9) & 1."""

    def __init__(self, arms):
        self.arms, self.lib, self.n = arms, arms.lib, arms.n
        assert self.lib.rk_adh_state_words() == layout.HS_WORDS
        with torch.cuda.device(arms.dev_index):
            self.hstate = torch.zeros(layout.HS_WORDS * self.n, dtype=torch.int32, device=arms.device)

    def mode_init(self, mode, stream=None):
        _cabi.check(self.lib.rk_adh_mode_init(self.hstate.data_ptr(), self.n, int(mode), self.arms._st(stream)))

    def update(self, K=1, now=None, trace=None, stream=None):
        """now: float32 [K, 4, n] servo angles (P1, DF_Left, DF_Right, P3) reported before each tick, or None."""
        if now is not None:
            assert now.is_cuda and now.dtype == torch.float32 and now.is_contiguous() and tuple(now.shape) == (int(K), 4, self.n)
        _cabi.check(self.lib.rk_adh_update(C.byref(self.arms.params), self.arms.state.data_ptr(), self.hstate.data_ptr(), self.n, int(K),
                                           None if now is None else now.data_ptr(), None if trace is None else trace.data_ptr(),
                                           self.arms._st(stream)))


class Arm:
    """Single arm (rk_adt_t): the statics of AD_task_main.cpp:108-156 behind one handle."""

    def __init__(self, params=None):
        self.lib = _cabi.load()
        self.h = C.c_void_p()
        _cabi.check(self.lib.rk_adt_create(C.byref(self.h), None if params is None else C.byref(params)))

    def close(self):
        if self.h:
            self.lib.rk_adt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        _cabi.check(self.lib.rk_adt_init(self.h))

    def push_cmdseq(self, seq_id, waypoints):
        """waypoints: [(dt_ms, (5 deg)), ...] (<= 32)."""
        q = _cabi.AdtPosCmdSeq()
        q.id, q.len = int(seq_id), len(waypoints)
        for k, (dt, deg) in enumerate(waypoints):
            q.cmd[k].dt_ms = int(dt)
            q.cmd[k].tgt_deg[:] = [float(x) for x in deg]
        _cabi.check(self.lib.rk_adt_push(self.h, C.byref(q)))

    def update(self):
        _cabi.check(self.lib.rk_adt_tick(self.h))

    def home_init(self, mode):
        """set_next_mode(INIT / INIT_POS_MOVE) -> init() of that mode (RK_ADH_MODE_*)."""
        _cabi.check(self.lib.rk_adt_home_init(self.h, int(mode)))

    def home_update(self, servo_now_deg=None):
        """One loop body with the homing mode active; returns isCompleted()."""
        done = C.c_int()
        now = None if servo_now_deg is None else (C.c_float * 4)(*[float(x) for x in servo_now_deg])
        _cabi.check(self.lib.rk_adt_home_tick(self.h, now, C.byref(done)))
        return bool(done.value)

    def get_q_cmdseq_status(self, seq_id):
        s = C.c_int32()
        _cabi.check(self.lib.rk_adt_status(self.h, int(seq_id), C.byref(s)))
        return s.value

    def get_tgt_deg(self):
        out = (C.c_float * 5)()
        _cabi.check(self.lib.rk_adt_get_targets_deg(self.h, out))
        return np.array(out[:], dtype=np.float32)

    def get_state(self):
        w = (C.c_uint32 * layout.AS_WORDS)()
        _cabi.check(self.lib.rk_adt_get_state(self.h, w))
        return np.array(w[:], dtype=np.uint32)
