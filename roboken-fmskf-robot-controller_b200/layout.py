"""State-block layout of include/robotick.h ("SoA of 128-bit planes") for host code.

word w of instance i of an n-instance block lives at  u32[((w // 4) * n + i) * 4 + (w % 4)].
AoS views ([n, WORDS] uint32) are for tests and host tooling only.
"""
import numpy as np

# ---- vehicle (RK_VS_*) ----------------------------------------------------------------
VS_POS_X, VS_POS_Y, VS_POS_TH, VS_FLAGS = 0, 1, 2, 3
VS_VEL_X, VS_VEL_Y, VS_VEL_TH, VS_TGT_X = 4, 5, 6, 7
VS_TGT_Y, VS_TGT_TH, VS_MOVE_CNT = 8, 9, 10
VS_INTERP0 = 12
VS_CTRL0 = VS_INTERP0 + 3 * 12
VS_MOTOR0 = VS_CTRL0 + 4 * 8
VS_WORDS = VS_MOTOR0 + 4 * 8
VS_FLAG_POWER_ON = 1

VI_VEL_NOW, VI_ACL_NOW, VI_VEL_TGT, VI_ACL_MAX = 0, 1, 2, 3
VI_JERK_P, VI_JERK_M, VI_DT1, VI_DT2 = 4, 5, 6, 7
VI_DT3, VI_VEL_INI, VI_ACL_INI, VI_DT = 8, 9, 10, 11

VC_PREV_VAL, VC_INTEG, VC_LPF_Y, VC_LPF_X = 0, 1, 2, 3
VC_NOW_TGT, VC_NOW_ERR, VC_NOW_CTRL = 4, 5, 6

VM_SUM_LO, VM_SUM_HI, VM_PREV_LO, VM_PREV_HI = 0, 1, 2, 3
VM_ANG_RPM, VM_CUR_TGT, VM_USEC, VM_PLANT = 4, 5, 6, 7

assert VS_WORDS == 112

# ---- IMU (RK_IS_*) ---------------------------------------------------------------------
IS_QINIT, IS_DATA, IS_FLAGS, IS_WORDS = 0, 4, 20, 24
IS_D_ACCEL, IS_D_GYRO, IS_D_MAG, IS_D_ANGLE, IS_D_QUT = 0, 3, 6, 9, 12
IS_FLAG_ERROR = 1
IMT_REGS = 16
(REG_AX, REG_AY, REG_AZ, REG_GX, REG_GY, REG_GZ, REG_HX, REG_HY, REG_HZ, REG_ROLL, REG_PITCH, REG_YAW,
 REG_Q0, REG_Q1, REG_Q2, REG_Q3) = range(16)
# WIT serial parser block (RK_IP_*): 11-byte window + fill count, flags, the 16 tracked sReg words
IP_WINDOW, IP_FLAGS, IP_SREG, IP_WORDS = 0, 3, 4, 12

# ---- RobotManager guard (RK_RS_* / RK_RI_*) -------------------------------------------------
RS_CMD_STATUS, RS_IGNORE_FLOOR, RS_NO_CMD_CNT, RS_ABORT, RS_WORDS = 0, 1, 2, 3, 4
RI_KIND, RI_A, RI_B, RI_C, RI_X, RI_Y, RI_Z, RI_FLOOR, RI_WORDS = 0, 1, 2, 3, 4, 6, 8, 10, 12
RM_ABORT_WALL_XP, RM_ABORT_WALL_XM, RM_ABORT_WALL_YP, RM_ABORT_WALL_YM = 1 << 0, 1 << 1, 1 << 2, 1 << 3
RM_ABORT_FLOOR_XP, RM_ABORT_FLOOR_XM, RM_ABORT_FLOOR_YP, RM_ABORT_FLOOR_YM = 1 << 8, 1 << 9, 1 << 10, 1 << 11
RM_ABORT_CONT_TRANS = 1 << 16

# ---- arm (RK_AS_* / RK_ACMD_*) ------------------------------------------------------------
AJ_Y0, AJ_P1, AJ_DFL, AJ_DFR, AJ_P2, AJ_R0, AJ_P3, AJ_NUM = range(8)
AS_FSM, AS_SEQ_IDX, AS_CMD_IDX, AS_MOVE_CNT, AS_CYCLE, AS_TOTAL_MS, AS_NOW_DT = range(7)
AS_NOW_TGT, AS_MOVE_DEG, AS_DFV_P, AS_DFV_R, AS_JOINT0 = 8, 13, 18, 19, 20
AS_JFLAGS, AS_MG_PRE_TGT, AS_ICS_POS, AS_ICS_SERVO = 48, 49, 50, 51
AS_MG_TX, AS_BLDC_TX0, AS_MG_CTRL, AS_WORDS = 52, 56, 68, 76
AJ_OFS, AJ_RAW_TGT, AJ_CURLIM, AJ_RAW_NOW = range(4)
ASTATE_STANDBY, ASTATE_MOVE_START, ASTATE_MOVING, ASTATE_COMPLETED = range(4)
AS_FSM_FIRSTCALL, AS_FSM_IS_COMP = 0x100, 0x200
AJF_CONNECTED, AJF_TORQUE_ON, AJF_INITIALIZED, AJF_TORQUE_PREV = 1, 2, 4, 8
ACMD_SLOTS, ACMD_MAX_LEN, ACMD_SLOT_WORDS = 4, 32, 260
ACMD_WORDS = ACMD_SLOTS * ACMD_SLOT_WORDS
ADT_TRACE_WORDS = 16
# ADTModePositioning mode block (RK_PS_*)
PS_STATE, PS_MOVE_CNT, PS_CYCLE, PS_QSIZE, PS_PREV_ID0, PS_PREV_ID1 = 0, 1, 2, 3, 4, 5
PS_NOW_CMD, PS_MOVE_DEG, PS_QUEUE, PS_WORDS = 8, 16, 24, 56
# homing mode block (RK_HS_*)
HS_STATE, HS_WAIT_CNT, HS_VEL_DIR, HS_WORDS = 0, 1, 4, 12
ADT_AXIS = (AJ_Y0, AJ_P1, AJ_P2, AJ_R0, AJ_P3)  # mode axes J0..J4 (AD_task_main.cpp:148)


def soa_to_aos(block, n, words):
    """[planes, n, 4] SoA block (any uint32 array of words*n elements) -> [n, words]."""
    b = np.asarray(block, dtype=np.uint32).reshape(words // 4, n, 4)
    return np.ascontiguousarray(b.transpose(1, 0, 2)).reshape(n, words)


def aos_to_soa(aos):
    """[n, words] -> flat SoA block (uint32, words*n)."""
    a = np.asarray(aos, dtype=np.uint32)
    n, words = a.shape
    return np.ascontiguousarray(a.reshape(n, words // 4, 4).transpose(1, 0, 2)).reshape(-1)


def f32(words):
    return np.asarray(words, dtype=np.uint32).view(np.float32)


def s16_lo(words):
    return (np.asarray(words, dtype=np.uint32) & 0xFFFF).astype(np.uint16).view(np.int16)


def s16_hi(words):
    return (np.asarray(words, dtype=np.uint32) >> 16).astype(np.uint16).view(np.int16)


def vehicle_view(aos):
    """Decode an [n, VS_WORDS] AoS vehicle state into named numpy arrays (copies)."""
    a = np.asarray(aos, dtype=np.uint32)
    out = {
        "pos": f32(a[:, VS_POS_X : VS_POS_X + 3]),
        "power_on": (a[:, VS_FLAGS] & VS_FLAG_POWER_ON) != 0,
        "vel": f32(a[:, VS_VEL_X : VS_VEL_X + 3]),
        "vel_tgt": f32(a[:, VS_TGT_X : VS_TGT_X + 3]),
    }
    m = a[:, VS_MOTOR0:VS_WORDS].reshape(-1, 4, 8)
    lo = m[:, :, VM_SUM_LO].astype(np.uint64)
    hi = m[:, :, VM_SUM_HI].astype(np.uint64)
    out["angle_sum"] = ((hi << np.uint64(32)) | lo).view(np.int64)
    out["raw_angle"] = s16_lo(m[:, :, VM_ANG_RPM])
    out["raw_rpm"] = s16_hi(m[:, :, VM_ANG_RPM])
    out["raw_cur"] = s16_lo(m[:, :, VM_CUR_TGT])
    out["cur_tgt"] = s16_hi(m[:, :, VM_CUR_TGT])
    return out
