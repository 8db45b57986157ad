// rk_arm.cu -- 5-axis arm tick batched (src/ArmDrive): ADTModePositioningSeq::update() +
// the joint command packers, one ADT::main loop body per tick (AD_task_main.cpp:208-229).
//
// One thread per arm.  The 19 state planes (304 B) are loaded once with 128-bit loads, the
// live words stay in registers for the K fused ticks, and are stored once.  The command ring
// (4 slots x 65 planes per arm) lives in its own block and is touched only at a segment start
// (2 planes = one waypoint, + the slot header), so a 100 Hz tick costs no HBM traffic at all:
// the kernel is issue/latency bound, not bandwidth bound.
#include <string.h>

#include "rk_common.cuh"

namespace rk {

// (int32_t)(float) as the x86 build of the firmware source performs it: cvttss2si, which
// returns INT_MIN for NaN / out of range (F2I.TRUNC saturates instead).
RK_DEV int32_t f2i_x86(float f) {
  const int32_t r = __float2int_rz(f);
  return (fabsf(f) < 2147483648.0f) ? r : (int32_t)0x80000000u;
}

// IcsBaseClass::degPos100 / posDeg100   lib/IcsClass_V210/src/IcsBaseClass.cpp:105-137
// (|deg| <= 18000 so the products fit 32 bits; C division truncates toward zero)
RK_DEV int ics_degPos100(int deg) {
  if(deg > 18000 || deg < -18000) return -1;
  return (deg * 2963) / 10000 + 7500;
}
RK_DEV int ics_posDeg100(int pos) {
  const long long a   = (long long)pos - 7500;
  const int       deg = (int)((a * 1000) / 296);
  if(deg > 18000) return 0x7FFF;
  if(deg < -18000) return -0x7FFF;
  return deg;
}

struct Joint {
  float ofs, raw_tgt, curlim, raw_now; // JointBase  AD_joint_base.hpp:62-74
};

struct Arm {
  uint32_t fsm, exec, head, cmd_idx; // nowState | flags ; u16_seq_exec_idx_ ; u16_seq_write_head_ ; u8_nowcmd_idx_
  int32_t  move_cnt, cycle;
  uint32_t total_ms, now_dt;
  float    now_tgt[5], move_deg[5];
  float    dfv_p, dfv_r;
  Joint    j[RK_AJ_NUM];
  uint32_t jflags;
  float    mg_pre;
  uint32_t ics_pos, ics_servo;
  uint32_t mg_tx[3];
  uint32_t bldc[3][4];
  uint32_t mg_ctrl[8]; // reserved words carried through unchanged
  uint32_t rsv0, mg_tx3;
};

RK_DEV void load_arm(const uint4 *st, int64_t n, int64_t i, Arm &a) {
  uint32_t w[RK_AS_WORDS];
#pragma unroll
  for(int pl = 0; pl < RK_AS_WORDS / 4; pl++) {
    const uint4 v = ld_plane(st, n, pl, i);
    w[4 * pl] = v.x, w[4 * pl + 1] = v.y, w[4 * pl + 2] = v.z, w[4 * pl + 3] = v.w;
  }
  a.fsm = w[RK_AS_FSM], a.exec = w[RK_AS_SEQ_IDX] & 0xFFFFu, a.head = w[RK_AS_SEQ_IDX] >> 16, a.cmd_idx = w[RK_AS_CMD_IDX];
  a.move_cnt = (int32_t)w[RK_AS_MOVE_CNT], a.cycle = (int32_t)w[RK_AS_CYCLE];
  a.total_ms = w[RK_AS_TOTAL_MS], a.now_dt = w[RK_AS_NOW_DT], a.rsv0 = w[RK_AS_RSV0];
#pragma unroll
  for(int k = 0; k < 5; k++) a.now_tgt[k] = u2f(w[RK_AS_NOW_TGT + k]), a.move_deg[k] = u2f(w[RK_AS_MOVE_DEG + k]);
  a.dfv_p = u2f(w[RK_AS_DFV_P]), a.dfv_r = u2f(w[RK_AS_DFV_R]);
#pragma unroll
  for(int k = 0; k < RK_AJ_NUM; k++) {
    const uint32_t *q = w + RK_AS_JOINT0 + 4 * k;
    a.j[k].ofs = u2f(q[RK_AJ_OFS]), a.j[k].raw_tgt = u2f(q[RK_AJ_RAW_TGT]);
    a.j[k].curlim = u2f(q[RK_AJ_CURLIM]), a.j[k].raw_now = u2f(q[RK_AJ_RAW_NOW]);
  }
  a.jflags = w[RK_AS_JFLAGS], a.mg_pre = u2f(w[RK_AS_MG_PRE_TGT]), a.ics_pos = w[RK_AS_ICS_POS], a.ics_servo = w[RK_AS_ICS_SERVO];
  a.mg_tx[0] = w[RK_AS_MG_TX], a.mg_tx[1] = w[RK_AS_MG_TX + 1], a.mg_tx[2] = w[RK_AS_MG_TX + 2], a.mg_tx3 = w[RK_AS_MG_TX + 3];
#pragma unroll
  for(int s = 0; s < 3; s++)
#pragma unroll
    for(int k = 0; k < 4; k++) a.bldc[s][k] = w[RK_AS_BLDC_TX0 + 4 * s + k];
#pragma unroll
  for(int k = 0; k < 8; k++) a.mg_ctrl[k] = w[RK_AS_MG_CTRL + k];
}

RK_DEV void store_arm(uint4 *st, int64_t n, int64_t i, const Arm &a) {
  uint32_t w[RK_AS_WORDS];
  w[RK_AS_FSM] = a.fsm, w[RK_AS_SEQ_IDX] = a.exec | (a.head << 16), w[RK_AS_CMD_IDX] = a.cmd_idx;
  w[RK_AS_MOVE_CNT] = (uint32_t)a.move_cnt, w[RK_AS_CYCLE] = (uint32_t)a.cycle;
  w[RK_AS_TOTAL_MS] = a.total_ms, w[RK_AS_NOW_DT] = a.now_dt, w[RK_AS_RSV0] = a.rsv0;
#pragma unroll
  for(int k = 0; k < 5; k++) w[RK_AS_NOW_TGT + k] = f2u(a.now_tgt[k]), w[RK_AS_MOVE_DEG + k] = f2u(a.move_deg[k]);
  w[RK_AS_DFV_P] = f2u(a.dfv_p), w[RK_AS_DFV_R] = f2u(a.dfv_r);
#pragma unroll
  for(int k = 0; k < RK_AJ_NUM; k++) {
    uint32_t *q = w + RK_AS_JOINT0 + 4 * k;
    q[RK_AJ_OFS] = f2u(a.j[k].ofs), q[RK_AJ_RAW_TGT] = f2u(a.j[k].raw_tgt);
    q[RK_AJ_CURLIM] = f2u(a.j[k].curlim), q[RK_AJ_RAW_NOW] = f2u(a.j[k].raw_now);
  }
  w[RK_AS_JFLAGS] = a.jflags, w[RK_AS_MG_PRE_TGT] = f2u(a.mg_pre), w[RK_AS_ICS_POS] = a.ics_pos, w[RK_AS_ICS_SERVO] = a.ics_servo;
  w[RK_AS_MG_TX] = a.mg_tx[0], w[RK_AS_MG_TX + 1] = a.mg_tx[1], w[RK_AS_MG_TX + 2] = a.mg_tx[2], w[RK_AS_MG_TX + 3] = a.mg_tx3;
#pragma unroll
  for(int s = 0; s < 3; s++)
#pragma unroll
    for(int k = 0; k < 4; k++) w[RK_AS_BLDC_TX0 + 4 * s + k] = a.bldc[s][k];
#pragma unroll
  for(int k = 0; k < 8; k++) w[RK_AS_MG_CTRL + k] = a.mg_ctrl[k];
#pragma unroll
  for(int pl = 0; pl < RK_AS_WORDS / 4; pl++) st_plane(st, n, pl, i, make_uint4(w[4 * pl], w[4 * pl + 1], w[4 * pl + 2], w[4 * pl + 3]));
}

// mode axis J0..J4 -> joint object (AD_task_main.cpp:148)
RK_DEV constexpr int axis_joint(int ax) { return ax == 0 ? RK_AJ_Y0 : ax == 1 ? RK_AJ_P1 : ax == 2 ? RK_AJ_P2 : ax == 3 ? RK_AJ_R0 : RK_AJ_P3; }

// JointBase::get_tgt_deg  AD_joint_base.hpp:47
RK_DEV float get_tgt_deg(const Arm &a, int ax) { return fsub(a.j[axis_joint(ax)].raw_tgt, a.j[axis_joint(ax)].ofs); }

// ADTModePositioningSeq::update   AD_mode_positioning_seq.cpp:13-117
RK_DEV void mode_update(Arm &a, const rk_adt_params_t &p, const uint4 *__restrict__ tab, int64_t n, int64_t i) {
  uint32_t state = a.fsm & 0xFFu;
  if(state == RK_ASTATE_STANDBY) { // exec_standby :24-42
    a.fsm |= RK_AS_FSM_IS_COMP;
    if(a.exec != a.head) {
      a.exec = (a.exec + 1) & 0xFFFFu;
      a.exec = (a.exec >= RK_ACMD_SLOTS) ? 0u : a.exec;
      a.cmd_idx  = 0;
      a.total_ms = 0;
      state      = RK_ASTATE_MOVE_START;
      a.fsm &= ~RK_AS_FSM_FIRSTCALL;
    }
  }
  if(state == RK_ASTATE_MOVE_START) { // exec_move_start :48-83
    const int      base_pl = (int)(a.exec % RK_ACMD_SLOTS) * (RK_ACMD_SLOT_WORDS / 4);
    const uint32_t len     = __ldg(&tab[(int64_t)base_pl * n + i]).y & 0xFFu;
    const uint32_t idx     = a.cmd_idx & 0xFFu;
    if(idx >= len) {
      state = RK_ASTATE_STANDBY;
    } else {
      // now_cmd_ = cmd_seq_[exec].cmd_seq[idx]: one waypoint = two 128-bit planes
      const int   pl = base_pl + 1 + 2 * (int)(idx % RK_ACMD_MAX_LEN);
      const uint4 w0 = __ldg(&tab[(int64_t)pl * n + i]);
      const uint4 w1 = __ldg(&tab[(int64_t)(pl + 1) * n + i]);
      a.now_dt       = w0.x;
      a.now_tgt[0] = u2f(w0.y), a.now_tgt[1] = u2f(w0.z), a.now_tgt[2] = u2f(w0.w), a.now_tgt[3] = u2f(w1.x), a.now_tgt[4] = u2f(w1.y);
      int32_t cnt = f2i_x86(fdiv(fmul(__uint2float_rn(a.now_dt - a.total_ms), 0.001f), p.cycle_time_s));
      cnt         = (cnt <= 0) ? 1 : cnt;
      const float fc = (float)cnt;
#pragma unroll
      for(int k = 0; k < 5; k++) a.move_deg[k] = fdiv(fsub(a.now_tgt[k], get_tgt_deg(a, k)), fc);
      a.move_cnt = cnt;
      a.total_ms = a.now_dt;
      a.cycle    = 0;
      a.fsm &= ~RK_AS_FSM_IS_COMP;
      state = RK_ASTATE_MOVING;
    }
  }
  if(state == RK_ASTATE_MOVING) { // exec_moving :89-117
    const float rem = (float)(a.move_cnt - a.cycle);
    float       raw[5];
#pragma unroll
    for(int k = 0; k < 5; k++) {
      // JointBase::set_tgt_ang_deg :42 (and the first line of the DfGear overrides)
      raw[k]                        = fadd(fsub(a.now_tgt[k], fmul(a.move_deg[k], rem)), a.j[axis_joint(k)].ofs);
      a.j[axis_joint(k)].raw_tgt = raw[k];
    }
    // JointDfGearPitch/Roll::set_tgt_ang_deg -> JointDfGearVirtual::set_P/R_tgt_ang_deg
    // (AD_joint_dfgear.hpp:19-29,60-63,93-96): the Pitch call's left/right targets are overwritten
    // by the Roll call that follows it in the same tick, so only the final pair is formed.
    a.dfv_p                 = fmul(raw[2], p.gear_ratio[RK_AJ_P2]);
    a.dfv_r                 = fmul(raw[3], p.gear_ratio[RK_AJ_R0]);
    a.j[RK_AJ_DFL].raw_tgt = fadd(fsub(a.dfv_p, a.dfv_r), a.j[RK_AJ_DFL].ofs);
    a.j[RK_AJ_DFR].raw_tgt = fadd(-fadd(a.dfv_p, a.dfv_r), a.j[RK_AJ_DFR].ofs);
    if(a.move_cnt <= a.cycle) {
      a.cmd_idx = (a.cmd_idx + 1) & 0xFFu;
      state     = RK_ASTATE_MOVE_START;
    } else {
      a.cycle++;
    }
  }
  a.fsm = (a.fsm & ~0xFFu) | state;
}

RK_DEV uint32_t jflag(const Arm &a, int k) { return (a.jflags >> (4 * k)) & 0xFu; }
RK_DEV void     set_prev(Arm &a, int k, bool on) {
  a.jflags = (a.jflags & ~((uint32_t)RK_AJF_TORQUE_PREV << (4 * k))) | ((on ? (uint32_t)RK_AJF_TORQUE_PREV : 0u) << (4 * k));
}

// JointMgServo::update -> subproc_posctrl   AD_joint_mg_servo.cpp:50-73,136-149.  The
// torque-control branches (:104-134; joint not initialised or torque off) are not part of this
// path: no MG frame is produced there (valid word = 0).
RK_DEV void mg_update(Arm &a, const rk_adt_params_t &p) {
  const uint32_t b  = jflag(a, RK_AJ_P1);
  const bool     on = (b & RK_AJF_TORQUE_ON) != 0, prev = (b & RK_AJF_TORQUE_PREV) != 0, ini = (b & RK_AJF_INITIALIZED) != 0;
  const float    tgt = a.j[RK_AJ_P1].raw_tgt;
  a.mg_tx[2]         = 0;
  if(prev && !on) {
  } else if(!ini && on) {
  } else if(on) {
    const float    v  = fabsf(fmul(fdiv(fsub(tgt, a.mg_pre), p.ctrl_time_s[RK_AJ_P1]), -10.0f));
    const uint32_t vl = (uint32_t)f2i_x86((v > 1800.0f) ? 1800.0f : v) & 0xFFFFu;
    a.mg_tx[0]        = 0xA4u | (vl << 16);
    a.mg_tx[1]        = (uint32_t)f2i_x86(fmul(tgt, -100.0f * 10.0f));
    a.mg_tx[2]        = 1;
  }
  set_prev(a, RK_AJ_P1, on);
  a.mg_pre = tgt;
}

// JointMyBldcServo::update   AD_joint_mybldc_servo.cpp:7-36 ; slot 0..2 = DF_Left, DF_Right, P3
template <int SLOT>
RK_DEV void bldc_update(Arm &a, const rk_adt_params_t &p) {
  constexpr int  k  = SLOT == 0 ? RK_AJ_DFL : SLOT == 1 ? RK_AJ_DFR : RK_AJ_P3;
  const uint32_t b  = jflag(a, k);
  const bool     on = (b & RK_AJF_TORQUE_ON) != 0, prev = (b & RK_AJF_TORQUE_PREV) != 0;
  uint32_t      *q  = a.bldc[SLOT];
  if(!on) {
    q[0] = 0, q[1] = 0, q[2] = 0x8002u;
  } else if(!prev) {
    q[0] = 0, q[1] = 0, q[2] = 0x8001u;
  } else {
    const int32_t  ang = f2i_x86(fmul(fmul(fmul(a.j[k].raw_tgt, p.gear_ratio[k]), p.motor_dir[k]), 65536.0f));
    const uint32_t ms  = (uint32_t)f2i_x86(fmul(p.ctrl_time_s[k], 1000.0f)) & 0xFFFFu;
    const uint32_t cl  = (uint32_t)f2i_x86(fmul(a.j[k].curlim, 256.0f)) & 0xFFFFu;
    q[0] = (uint32_t)ang, q[1] = ms | (cl << 16), q[2] = 0x8010u;
  }
  q[3] = 1;
  set_prev(a, k, on);
}

// JointIcsServo::update   AD_joint_ics_servo.cpp:5-29, the UART replaced by an ideal servo that
// answers with the commanded position (a free command answers with the last one).
RK_DEV void ics_update(Arm &a, const rk_adt_params_t &p) {
  const uint32_t b = jflag(a, RK_AJ_Y0);
  if(!(b & RK_AJF_CONNECTED)) return;
  const int tgt_pos = ics_degPos100(f2i_x86(fmul(fmul(a.j[RK_AJ_Y0].raw_tgt, p.motor_dir[RK_AJ_Y0]), 100.0f)));
  if(tgt_pos == -1) return;
  int now_pos;
  if(b & RK_AJF_TORQUE_ON) {
    if(tgt_pos > 11500 || tgt_pos < 3500) { // IcsBaseClass::setPos range check: nothing is sent
      now_pos = -1;
    } else {
      a.ics_pos   = (uint32_t)tgt_pos;
      a.ics_servo = (uint32_t)(tgt_pos - 7500);
      now_pos     = tgt_pos;
    }
  } else {
    a.ics_pos = 0xFFFFFFFFu;
    now_pos   = (int)(int32_t)a.ics_servo + 7500;
  }
  a.j[RK_AJ_Y0].raw_now = fmul(fmul((float)ics_posDeg100(now_pos), 0.01f), p.motor_dir[RK_AJ_Y0]);
}

RK_DEV uint32_t bldc_id_byte(uint32_t id) { return (id & 0xFFu) | ((id & 0x8000u) ? 0x80u : 0u); }

RK_DEV void arm_tick(Arm &a, const rk_adt_params_t &p, const uint4 *__restrict__ tab, int64_t n, int64_t i) {
  mode_update(a, p, tab, n, i);
  mg_update(a, p);
  bldc_update<0>(a, p);
  bldc_update<1>(a, p);
  bldc_update<2>(a, p);
  ics_update(a, p);
}

template <bool TRACE>
__global__ void __launch_bounds__(128)
adt_update_kernel(const rk_adt_params_t p, uint4 *__restrict__ state, const uint4 *__restrict__ tab, int64_t n, int K,
                  uint32_t *__restrict__ trace) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Arm a;
  load_arm(state, n, i, a);
  for(int t = 0; t < K; t++) {
    arm_tick(a, p, tab, n, i);
    if(TRACE) {
      uint32_t *tr = trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i;
#pragma unroll
      for(int k = 0; k < 5; k++) tr[(int64_t)k * n] = f2u(get_tgt_deg(a, k));
      tr[5 * n] = a.mg_tx[0] >> 16, tr[6 * n] = a.mg_tx[1];
#pragma unroll
      for(int s = 0; s < 3; s++) tr[(int64_t)(7 + s) * n] = a.bldc[s][0];
      tr[10 * n] = a.ics_pos;
      tr[11 * n] = a.fsm & 0xFFu;
      tr[12 * n] = a.cmd_idx;
      tr[13 * n] = bldc_id_byte(a.bldc[0][2]) | (bldc_id_byte(a.bldc[1][2]) << 8) | (bldc_id_byte(a.bldc[2][2]) << 16);
      tr[14 * n] = 0u, tr[15 * n] = 0u;
    }
  }
  store_arm(state, n, i, a);
}

// prepare_task() + a finished INIT mode + ADTModeBase::init() -> doInit()
// (AD_task_main.cpp:170-193, AD_mode_initialize.cpp:58-63,133-135, AD_mode_positioning_seq.cpp:5-11)
__global__ void __launch_bounds__(128) adt_mode_init_kernel(const rk_adt_params_t p, uint4 *__restrict__ state, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Arm a;
  load_arm(state, n, i, a);
  // JointIcsServo::init  AD_joint_ics_servo.cpp:35-56: free command, target := present position
  const float now = fmul(fmul((float)ics_posDeg100((int)(int32_t)a.ics_servo + 7500), 0.01f), p.motor_dir[RK_AJ_Y0]);
  a.j[RK_AJ_Y0].raw_now = now, a.j[RK_AJ_Y0].raw_tgt = now;
  a.ics_pos             = 0xFFFFFFFFu;
  uint32_t f            = a.jflags;
  const uint32_t all    = RK_AJF_CONNECTED | RK_AJF_TORQUE_ON | RK_AJF_INITIALIZED;
  f = (f & ~(0xFu << (4 * RK_AJ_Y0))) | (all << (4 * RK_AJ_Y0));
  f = (f & ~(0xFu << (4 * RK_AJ_P1))) | (all << (4 * RK_AJ_P1)); // JointMgServo::init clears torque_on_prev
  f |= (uint32_t)RK_AJF_TORQUE_ON << (4 * RK_AJ_DFL);            // JointDfGearPitch::set_torque_on -> both motors
  f |= (uint32_t)RK_AJF_TORQUE_ON << (4 * RK_AJ_DFR);
  f |= (uint32_t)RK_AJF_INITIALIZED << (4 * RK_AJ_P2);
  f |= (uint32_t)RK_AJF_INITIALIZED << (4 * RK_AJ_R0);
  f |= (uint32_t)(RK_AJF_TORQUE_ON | RK_AJF_INITIALIZED) << (4 * RK_AJ_P3);
  a.jflags = f;
  a.j[RK_AJ_Y0].curlim  = p.curlim_default_A[RK_AJ_Y0];
  a.j[RK_AJ_P1].curlim  = p.curlim_default_A[RK_AJ_P1];
  a.j[RK_AJ_DFL].curlim = p.curlim_default_A[RK_AJ_R0]; // J3 (Roll) is the last to set both motors
  a.j[RK_AJ_DFR].curlim = p.curlim_default_A[RK_AJ_R0];
  a.j[RK_AJ_P3].curlim  = p.curlim_default_A[RK_AJ_P3];
  a.fsm  = RK_ASTATE_STANDBY | RK_AS_FSM_FIRSTCALL;
  a.exec = RK_ACMD_SLOTS - 1, a.head = RK_ACMD_SLOTS - 1;
  store_arm(state, n, i, a);
}

// ADTModePositioningSeq::push_cmdseq   AD_mode_positioning_seq.cpp:124-137
__global__ void __launch_bounds__(128)
adt_push_kernel(uint4 *__restrict__ state, uint4 *__restrict__ tab, int64_t n, const uint4 *__restrict__ seq, const uint8_t *__restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  if(valid && !valid[i]) return;
  uint4          s0   = ld_plane(state, n, 0, i);
  const uint32_t exec = s0.y & 0xFFFFu, head = s0.y >> 16;
  uint32_t       nw   = (head + 1) & 0xFFFFu;
  nw                  = (nw >= RK_ACMD_SLOTS) ? 0u : nw;
  if(nw == exec) return; // ring full: dropped silently
  const int base = (int)nw * (RK_ACMD_SLOT_WORDS / 4);
  for(int pl = 0; pl < RK_ACMD_SLOT_WORDS / 4; pl++) {
    uint4 v = __ldcs(seq + (int64_t)pl * n + i);
    if(pl == 0) v.y &= 0xFFu; // u8_cmd_seq_len
    tab[(int64_t)(base + pl) * n + i] = v;
  }
  s0.y = exec | (nw << 16);
  st_plane(state, n, 0, i, s0);
}

// ADTModePositioningSeq::get_q_cmdseq_status   AD_mode_positioning_seq.cpp:146-184
__global__ void __launch_bounds__(128)
adt_status_kernel(const uint4 *__restrict__ state, const uint4 *__restrict__ tab, int64_t n, const uint32_t *__restrict__ ids, int32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  const uint4    s0   = ld_plane(state, n, 0, i);
  const uint32_t exec = s0.y & 0xFFFFu, head = s0.y >> 16, id = ids[i];
  int32_t        sts  = 99;
  if((s0.x & RK_AS_FSM_FIRSTCALL) && id == 0) {
    out[i] = 99;
    return;
  }
#pragma unroll
  for(uint32_t s = 0; s < RK_ACMD_SLOTS; s++) {
    if(__ldg(&tab[(int64_t)s * (RK_ACMD_SLOT_WORDS / 4) * n + i]).x != id) continue;
    if(exec == head) {
      const uint32_t len = __ldg(&tab[(int64_t)(exec % RK_ACMD_SLOTS) * (RK_ACMD_SLOT_WORDS / 4) * n + i]).y & 0xFFu;
      sts                = ((s0.z & 0xFFu) >= len) ? 1 : 0;
    } else if(exec < head) {
      sts = (exec <= s && s <= head) ? 0 : 1;
    } else {
      sts = ((exec <= s && s < RK_ACMD_SLOTS) || s <= head) ? 0 : 1;
    }
  }
  out[i] = sts;
}

// JointBase::get_tgt_deg() of the five mode axes (AD_joint_base.hpp:47)
__global__ void adt_targets_kernel(const uint4 *__restrict__ state, int64_t n, int64_t i, float *__restrict__ out5) {
  if(threadIdx.x != 0 || blockIdx.x != 0) return;
  Arm a;
  load_arm(state, n, i, a);
#pragma unroll
  for(int k = 0; k < 5; k++) out5[k] = get_tgt_deg(a, k);
}

} // namespace rk

using namespace rk;

extern "C" {

void rk_adt_default_params(rk_adt_params_t *p) {
  if(!p) return;
  // AD_task_main.cpp:38-107 in RK_AJ_* order (Y0, P1, DF_Left, DF_Right, P2, R0, P3)
  const float gear[RK_AJ_NUM] = {1.0f, 1.0f, 1.0f, 1.0f, 24.0f / 7.0f, 48.0f / 7.0f, 48.0f / 19.0f};
  const float dir[RK_AJ_NUM]  = {-1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, -1.0f};
  const float cl[RK_AJ_NUM]   = {3.0f, 0.7f, 0.5f, 0.5f, 1.0f, 1.0f, 0.8f};
  for(int k = 0; k < RK_AJ_NUM; k++) p->ctrl_time_s[k] = 0.01f, p->gear_ratio[k] = gear[k], p->motor_dir[k] = dir[k], p->curlim_default_A[k] = cl[k];
  p->cycle_time_s = 0.01f; // AD_task_main.cpp:149
}

size_t rk_adt_state_words(void) { return RK_AS_WORDS; }
size_t rk_adt_state_bytes(int64_t n) { return n < 0 ? 0 : (size_t)n * RK_AS_WORDS * 4u; }
size_t rk_adt_cmdtab_bytes(int64_t n) { return n < 0 ? 0 : (size_t)n * RK_ACMD_WORDS * 4u; }

static int adt_check(const char *who, const void *a, const void *b, int64_t n) {
  if(n < 0) {
    set_error("%s: n < 0", who);
    return RK_ERR_ARG;
  }
  if(!a || ((uintptr_t)a & 15u) || ((uintptr_t)b & 15u)) {
    set_error("%s: blocks must be non-NULL and 16-byte aligned", who);
    return RK_ERR_ARG;
  }
  return require_device();
}
static unsigned adt_grid(int64_t n) { return (unsigned)((n + 127) / 128); }

int rk_adt_mode_init(const rk_adt_params_t *p, void *d_state, int64_t n, void *stream) {
  if(n == 0) return RK_OK;
  if(!p) {
    set_error("rk_adt_mode_init: params NULL");
    return RK_ERR_ARG;
  }
  if(int rc = adt_check("rk_adt_mode_init", d_state, nullptr, n)) return rc;
  adt_mode_init_kernel<<<adt_grid(n), 128, 0, (cudaStream_t)stream>>>(*p, (uint4 *)d_state, n);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

int rk_adt_push_cmdseq(void *d_state, void *d_cmdtab, int64_t n, const void *d_seq, const uint8_t *d_valid, void *stream) {
  if(n == 0) return RK_OK;
  if(!d_cmdtab || !d_seq || ((uintptr_t)d_seq & 15u)) {
    set_error("rk_adt_push_cmdseq: d_cmdtab / d_seq must be non-NULL and 16-byte aligned");
    return RK_ERR_ARG;
  }
  if(int rc = adt_check("rk_adt_push_cmdseq", d_state, d_cmdtab, n)) return rc;
  adt_push_kernel<<<adt_grid(n), 128, 0, (cudaStream_t)stream>>>((uint4 *)d_state, (uint4 *)d_cmdtab, n, (const uint4 *)d_seq, d_valid);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

int rk_adt_update(const rk_adt_params_t *p, void *d_state, const void *d_cmdtab, int64_t n, int32_t K, uint32_t *d_trace, void *stream) {
  if(n == 0 || K == 0) return RK_OK;
  if(!p || K < 0 || !d_cmdtab) {
    set_error("rk_adt_update: params / cmdtab NULL or K < 0");
    return RK_ERR_ARG;
  }
  if(int rc = adt_check("rk_adt_update", d_state, d_cmdtab, n)) return rc;
  if(d_trace)
    adt_update_kernel<true><<<adt_grid(n), 128, 0, (cudaStream_t)stream>>>(*p, (uint4 *)d_state, (const uint4 *)d_cmdtab, n, K, d_trace);
  else
    adt_update_kernel<false><<<adt_grid(n), 128, 0, (cudaStream_t)stream>>>(*p, (uint4 *)d_state, (const uint4 *)d_cmdtab, n, K, nullptr);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

int rk_adt_cmdseq_status(const void *d_state, const void *d_cmdtab, int64_t n, const uint32_t *d_id, int32_t *d_status, void *stream) {
  if(n == 0) return RK_OK;
  if(!d_cmdtab || !d_id || !d_status) {
    set_error("rk_adt_cmdseq_status: NULL argument");
    return RK_ERR_ARG;
  }
  if(int rc = adt_check("rk_adt_cmdseq_status", d_state, d_cmdtab, n)) return rc;
  adt_status_kernel<<<adt_grid(n), 128, 0, (cudaStream_t)stream>>>((const uint4 *)d_state, (const uint4 *)d_cmdtab, n, d_id, d_status);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

// ---- single-instance handle: a batch of one over the same kernels ---------------------------
struct rk_adt {
  rk_adt_params_t p;
  uint32_t       *d_state, *d_tab, *d_seq, *d_misc; // d_misc: [0] id, [1] status, [2..6] targets
  uint32_t       *h_stage;                         // pinned, RK_ACMD_SLOT_WORDS words
  cudaStream_t    st;
};

int rk_adt_create(rk_adt_t **out, const rk_adt_params_t *p) {
  if(!out) return RK_ERR_ARG;
  *out = nullptr;
  if(int rc = require_device()) return rc;
  rk_adt *h = new rk_adt();
  memset(h, 0, sizeof(*h));
  if(p) h->p = *p;
  else rk_adt_default_params(&h->p);
  cudaError_t e = cudaMalloc((void **)&h->d_state, RK_AS_WORDS * 4);
  if(e == cudaSuccess) e = cudaMalloc((void **)&h->d_tab, RK_ACMD_WORDS * 4);
  if(e == cudaSuccess) e = cudaMalloc((void **)&h->d_seq, RK_ACMD_SLOT_WORDS * 4);
  if(e == cudaSuccess) e = cudaMalloc((void **)&h->d_misc, 32);
  if(e == cudaSuccess) e = cudaMallocHost((void **)&h->h_stage, RK_ACMD_SLOT_WORDS * 4);
  if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if(e == cudaSuccess) e = cudaMemsetAsync(h->d_state, 0, RK_AS_WORDS * 4, h->st);
  if(e == cudaSuccess) e = cudaMemsetAsync(h->d_tab, 0, RK_ACMD_WORDS * 4, h->st);
  if(e != cudaSuccess) {
    int rc = cuda_fail(e, "rk_adt_create");
    rk_adt_destroy(h);
    return rc;
  }
  *out = h;
  return RK_OK;
}
void rk_adt_destroy(rk_adt_t *h) {
  if(!h) return;
  if(h->st) {
    cudaStreamSynchronize(h->st);
    cudaStreamDestroy(h->st);
  }
  if(h->d_state) cudaFree(h->d_state);
  if(h->d_tab) cudaFree(h->d_tab);
  if(h->d_seq) cudaFree(h->d_seq);
  if(h->d_misc) cudaFree(h->d_misc);
  if(h->h_stage) cudaFreeHost(h->h_stage);
  delete h;
}
int rk_adt_init(rk_adt_t *h) {
  if(!h) return RK_ERR_ARG;
  return rk_adt_mode_init(&h->p, h->d_state, 1, h->st);
}
int rk_adt_push(rk_adt_t *h, const rk_adt_poscmdseq_t *seq) {
  if(!h || !seq) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st));
  uint32_t *w = h->h_stage;
  memset(w, 0, RK_ACMD_SLOT_WORDS * 4);
  w[0] = seq->id, w[1] = seq->len;
  for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
    w[4 + 8 * k] = seq->cmd[k].dt_ms;
    memcpy(&w[4 + 8 * k + 1], seq->cmd[k].tgt_deg, 20);
  }
  RK_CUDA(cudaMemcpyAsync(h->d_seq, w, RK_ACMD_SLOT_WORDS * 4, cudaMemcpyHostToDevice, h->st));
  return rk_adt_push_cmdseq(h->d_state, h->d_tab, 1, h->d_seq, nullptr, h->st);
}
int rk_adt_tick(rk_adt_t *h) {
  if(!h) return RK_ERR_ARG;
  return rk_adt_update(&h->p, h->d_state, h->d_tab, 1, 1, nullptr, h->st);
}
int rk_adt_status(rk_adt_t *h, uint32_t id, int32_t *status) {
  if(!h || !status) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st));
  h->h_stage[0] = id;
  RK_CUDA(cudaMemcpyAsync(h->d_misc, h->h_stage, 4, cudaMemcpyHostToDevice, h->st));
  if(int rc = rk_adt_cmdseq_status(h->d_state, h->d_tab, 1, h->d_misc, (int32_t *)(h->d_misc + 1), h->st)) return rc;
  RK_CUDA(cudaMemcpyAsync(h->h_stage + 1, h->d_misc + 1, 4, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  *status = (int32_t)h->h_stage[1];
  return RK_OK;
}
int rk_adt_get_state(rk_adt_t *h, uint32_t words[RK_AS_WORDS]) {
  if(!h || !words) return RK_ERR_ARG;
  RK_CUDA(cudaMemcpyAsync(h->h_stage, h->d_state, RK_AS_WORDS * 4, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(words, h->h_stage, RK_AS_WORDS * 4);
  return RK_OK;
}
int rk_adt_set_state(rk_adt_t *h, const uint32_t words[RK_AS_WORDS]) {
  if(!h || !words) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(h->h_stage, words, RK_AS_WORDS * 4);
  RK_CUDA(cudaMemcpyAsync(h->d_state, h->h_stage, RK_AS_WORDS * 4, cudaMemcpyHostToDevice, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  return RK_OK;
}
int rk_adt_get_targets_deg(rk_adt_t *h, float out[5]) {
  if(!h || !out) return RK_ERR_ARG;
  adt_targets_kernel<<<1, 32, 0, h->st>>>((const uint4 *)h->d_state, 1, 0, (float *)(h->d_misc + 2));
  RK_CUDA(cudaGetLastError());
  RK_CUDA(cudaMemcpyAsync(h->h_stage, h->d_misc + 2, 20, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(out, h->h_stage, 20);
  return RK_OK;
}
}
