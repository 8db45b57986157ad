/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * The reference's WHOLE vehicle task, unmodified: src/VehicleDrive/VD_task_main.cpp (the 100 Hz
 * message loop VDT::main with its speed limiters and move-time auto-stop, :119-151,165-322, the
 * 1 kHz ISR can_tx_routine_intr :366-372, the object wiring :75-108) is included below where it
 * lies -- its objects have internal linkage -- together with VD_vehicle_controller.cpp,
 * VD_motor_if_m2006.cpp and util_mymath.cpp compiled by oracle/Makefile.  FreeRTOS, FlexCAN,
 * the MPU6500 SPI link and the IntervalTimer are the stubs of oracle/stubs/.
 *
 * Scheduling model (the same as rk_vdt_rollout with task_period = 10): VDT::main iteration k
 * runs before ISR tick 10*k; vTaskDelayUntil() (stub, below) runs the ten ISR ticks of the
 * period that just ended -- integer motor plant -> MOTOR_IF_M2006::rx_callback x4 ->
 * can_tx_routine_intr() -- and ends the loop by throwing when the rollout is complete.
 * IMT::get_status_now_yaw() returns the scripted IMU yaw (degrees).
 */
#define ORACLE_MICROS_EXTERN
#include <stdlib.h>
#include <string.h>

#include <Arduino.h>
#include <FreeRTOS_TEENSY4.h>
#include <message_buffer.h>

#include "robotick.h"

/* ---- harness state ------------------------------------------------------------------------ */
namespace {
struct TaskRun {
  int64_t                 n = 0, i = 0;
  const rk_vdt_rollout_t *a = nullptr;
  int                     tick = 0;      /* next ISR tick to run */
  int                     period_k = -1; /* VDT::main iteration being executed */
  float                   yaw_deg = 0.0f;
  uint32_t                micros_now = 0;
  int32_t                 plant_rpm[4] = {0, 0, 0, 0}, plant_ang[4] = {0, 0, 0, 0};
} g_run;
struct LoopDone {};
void run_isr_tick();
} // namespace

uint32_t micros() { return g_run.micros_now; }
uint32_t get_gptimer_cnt() { return 0; }
HardwareSerial Serial6;
HardwareSerial Serial7;
namespace DEBUG {
char EXT_PRINT_BUF[1024];
void print(char *, uint32_t) {}
void record_proc_load(uint8_t, uint8_t) {}
} // namespace DEBUG
namespace LGT {
void push_buffer(char *, uint32_t) {}
} // namespace LGT
namespace IMT { /* src/Imu/imu_task_main.hpp:51 -- scripted */
float get_status_now_yaw() { return g_run.yaw_deg; }
} // namespace IMT

TickType_t xTaskGetTickCount() { return 0; }
MessageBufferHandle_t xMessageBufferCreate(size_t) { return (MessageBufferHandle_t)&g_run; }
size_t xMessageBufferSend(MessageBufferHandle_t, const void *, size_t, uint32_t) { return 0; }

/* the reference task, as it is */
#include "VehicleDrive/VD_task_main.cpp"

namespace {
using namespace VDT;

/* one message slot per command segment: a rk_vdt_cmd_t of kind RK_CMD_MSG_* carries a VDT::MSG_REQ */
bool next_message(MSG_REQ *m) {
  const rk_vdt_rollout_t *a = g_run.a;
  const int               t = g_run.period_k * a->task_period;
  if(!a->d_cmd || a->seg_len <= 0 || (t % a->seg_len) != 0 || (t / a->seg_len) >= a->n_seg) return false;
  const rk_vdt_cmd_t &c    = a->d_cmd[(int64_t)(t / a->seg_len) * g_run.n + g_run.i];
  const uint32_t      kind = (uint32_t)c.kind & 0xFFu, time_ms = (uint32_t)c.kind >> 8;
  memset(m, 0, sizeof(*m));
  if(kind == RK_CMD_MSG_MOVE_DIR) {
    uint32_t u[2];
    memcpy(u, &c.vx, 8);
    m->move_dir.cmn.MsgId   = REQ_MOVE_DIR;
    m->move_dir.u32_cmd     = u[0];
    m->move_dir.u32_speed   = u[1];
    m->move_dir.u32_time_ms = time_ms;
    return true;
  }
  if(kind == RK_CMD_MSG_MOVE_CONT_DIR) {
    m->move_cont_dir.cmn.MsgId       = REQ_MOVE_CONT_DIR;
    m->move_cont_dir.fl_vel_x_mmps   = c.vx;
    m->move_cont_dir.fl_vel_y_mmps   = c.vy;
    m->move_cont_dir.fl_vel_th_radps = c.vth;
    m->move_cont_dir.u32_time_ms     = time_ms;
    return true;
  }
  if(kind == RK_CMD_MSG_UNKNOWN) { /* an id VDT::main ignores (the countdown still runs) */
    m->common.MsgId = MSG_UNKNOWN;
    return true;
  }
  return false;
}

MOTOR_IF_M2006 *const MOTORS[4] = {&FL_motor, &BL_motor, &BR_motor, &FR_motor};
UTIL::FF_PI_D *const  CTRLS[4]  = {&FL_m_ctrl, &BL_m_ctrl, &BR_m_ctrl, &FR_m_ctrl};
UTIL::VelInterpConstJerk *const INTERPS[3] = {&VelIntpConstJerk_Xdir, &VelIntpConstJerk_Ydir, &VelIntpConstJerk_Tdir};

inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
inline uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

void run_isr_tick() {
  const rk_vdt_rollout_t *a = g_run.a;
  const int               t = g_run.tick;
  if(a->d_yaw && a->yaw_period > 0 && (t % a->yaw_period) == 0 && (t / a->yaw_period) < a->n_yaw)
    g_run.yaw_deg = a->d_yaw[(int64_t)(t / a->yaw_period) * g_run.n + g_run.i];
  g_run.micros_now = (uint32_t)((t + 1) * 1000);
  for(int k = 0; k < 4; k++) { /* integer motor plant (robotick.h RK_SENSOR_PLANT) -> CAN mailbox callback */
    int32_t cur = MOTORS[k]->get_rawCurr_tgt(), rpm = g_run.plant_rpm[k], ang = g_run.plant_ang[k];
    rpm += ((cur * 4 - rpm) >> 4);
    ang = (ang + rpm * 8192 / 60000) & 8191;
    g_run.plant_rpm[k] = rpm, g_run.plant_ang[k] = ang;
    CAN_message_t msg;
    msg.id = 0x201 + k;
    const uint8_t f[8] = {(uint8_t)(ang >> 8), (uint8_t)ang, (uint8_t)(rpm >> 8), (uint8_t)rpm, (uint8_t)(cur >> 8), (uint8_t)cur, 0, 0};
    memcpy(msg.buf, f, 8);
    M_CAN.handler[k](msg); /* CAN_CTRL<CAN1>::mbK_callback -> rx_callback(msg.buf, micros() & 0x7FFF) */
  }
  canTxTimer.isr(); /* can_tx_routine_intr(): set_now_yaw_world(deg2rad(IMU yaw)); update(); M_CAN.tx_routine() */
  if(a->d_trace) {
    uint32_t *tr = a->d_trace + (int64_t)t * RK_VDT_TRACE_WORDS * g_run.n + g_run.i;
    float     px, py, pr, vx, vy, vr;
    get_status_now_vehicle_pos_world(px, py, pr);
    get_status_now_vehicle_vel(vx, vy, vr);
    Direction g;
    vhclCtrl.get_vehicle_vel_tgt_mmps_latest(g);
    const float f[9] = {px, py, pr, vx, vy, vr, g.x, g.y, g.th};
    for(int j = 0; j < 9; j++) tr[(int64_t)j * g_run.n] = f2u(f[j]);
    /* the currents as the C610 frame carries them on the wire (VD_can_controller.hpp:43-55) */
    for(int k = 0; k < 4; k++)
      tr[(int64_t)(9 + k) * g_run.n] = (uint32_t)(int32_t)(int16_t)((M_CAN.last_tx.buf[2 * k] << 8) | M_CAN.last_tx.buf[2 * k + 1]);
    tr[(int64_t)13 * g_run.n] = U32_MOVE_TIME_CNT_ORDER;
    /* ... and the frame itself, as CAN_CTRL<CAN1>::tx_routine() handed it to write() (words 14-15, wire order from the low byte) */
    for(int j = 0; j < 2; j++) {
      const uint8_t *b = M_CAN.last_tx.buf + 4 * j;
      tr[(int64_t)(14 + j) * g_run.n] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    }
  }
  g_run.tick++;
}

/* power-on state of every object of the task (static zero-initialisation + constructors) */
void reset_task_objects() {
  for(int k = 0; k < 4; k++) {
    MOTOR_IF_M2006 *m = MOTORS[k];
    m->s64_rawAngleSum = 0, m->s16_rawCurr_tgt = 0, m->status_head = 0;
    memset(m->status_buf, 0, sizeof(m->status_buf));
    CTRLS[k]->reset();
    vhclCtrl.s64_rawAngleSumPrev[k] = 0;
  }
  for(int a = 0; a < 3; a++) {
    INTERPS[a]->vel_now_ = 0.0f, INTERPS[a]->acl_now_ = 0.0f, INTERPS[a]->u8_now_use_ = 0;
    memset(INTERPS[a]->sts, 0, sizeof(INTERPS[a]->sts));
  }
  vhclCtrl.now_vhcl_pos_m_ = Direction{0, 0, 0}, vhclCtrl.now_vhcl_vel_mmps = Direction{0, 0, 0};
  vhclCtrl.now_vhcl_vel_tgt_mmps = Direction{0, 0, 0}, vhclCtrl.isPowerOn = false;
  U32_MOVE_TIME_CNT_ORDER = 0;
}

void export_task_state(uint32_t *w) {
  memset(w, 0, sizeof(uint32_t) * RK_VS_WORDS);
  w[RK_VS_POS_X] = f2u(vhclCtrl.now_vhcl_pos_m_.x), w[RK_VS_POS_Y] = f2u(vhclCtrl.now_vhcl_pos_m_.y);
  w[RK_VS_POS_TH] = f2u(vhclCtrl.now_vhcl_pos_m_.th), w[RK_VS_FLAGS] = vhclCtrl.isPowerOn ? RK_VS_FLAG_POWER_ON : 0u;
  w[RK_VS_VEL_X] = f2u(vhclCtrl.now_vhcl_vel_mmps.x), w[RK_VS_VEL_Y] = f2u(vhclCtrl.now_vhcl_vel_mmps.y);
  w[RK_VS_VEL_TH] = f2u(vhclCtrl.now_vhcl_vel_mmps.th), w[RK_VS_TGT_X] = f2u(vhclCtrl.now_vhcl_vel_tgt_mmps.x);
  w[RK_VS_TGT_Y] = f2u(vhclCtrl.now_vhcl_vel_tgt_mmps.y), w[RK_VS_TGT_TH] = f2u(vhclCtrl.now_vhcl_vel_tgt_mmps.th);
  w[RK_VS_MOVE_CNT] = U32_MOVE_TIME_CNT_ORDER;
  for(int a = 0; a < 3; a++) {
    uint32_t *q  = w + RK_VS_INTERP0 + 12 * a;
    auto     &it = *INTERPS[a];
    auto     &p  = it.sts[it.u8_now_use_];
    q[RK_VI_VEL_NOW] = f2u(it.vel_now_), q[RK_VI_ACL_NOW] = f2u(it.acl_now_), q[RK_VI_VEL_TGT] = f2u(p.vel_tgt_);
    q[RK_VI_ACL_MAX] = f2u(p.acl_max_), q[RK_VI_JERK_P] = f2u(p.jerk_p_), q[RK_VI_JERK_M] = f2u(p.jerk_m_);
    q[RK_VI_DT1] = f2u(p.dt1_), q[RK_VI_DT2] = f2u(p.dt2_), q[RK_VI_DT3] = f2u(p.dt3_);
    q[RK_VI_VEL_INI] = f2u(p.vel_ini_), q[RK_VI_ACL_INI] = f2u(p.acl_ini_), q[RK_VI_DT] = f2u(p.dt_);
  }
  for(int k = 0; k < 4; k++) {
    uint32_t *q = w + RK_VS_CTRL0 + 8 * k;
    auto     &c = *CTRLS[k];
    q[RK_VC_PREV_VAL] = f2u(c.prev_val_), q[RK_VC_INTEG] = f2u(c.Integ_), q[RK_VC_LPF_Y] = f2u(c.velLpf_.prev_Y_);
    q[RK_VC_LPF_X] = f2u(c.velLpf_.prev_X_), q[RK_VC_NOW_TGT] = f2u(c.now_tgt_), q[RK_VC_NOW_ERR] = f2u(c.now_error_);
    q[RK_VC_NOW_CTRL] = f2u(c.now_ctrl_);
    uint32_t *r  = w + RK_VS_MOTOR0 + 8 * k;
    auto     &m  = *MOTORS[k];
    auto     &st = m.status_buf[m.status_head];
    uint64_t  sum = (uint64_t)m.s64_rawAngleSum, prev = (uint64_t)vhclCtrl.s64_rawAngleSumPrev[k];
    r[RK_VM_SUM_LO] = (uint32_t)sum, r[RK_VM_SUM_HI] = (uint32_t)(sum >> 32);
    r[RK_VM_PREV_LO] = (uint32_t)prev, r[RK_VM_PREV_HI] = (uint32_t)(prev >> 32);
    r[RK_VM_ANG_RPM] = pack16(st.s16_rawAngle, st.s16_rawSpeedRpm), r[RK_VM_CUR_TGT] = pack16(st.s16_rawCurr, m.s16_rawCurr_tgt);
    r[RK_VM_USEC] = pack16(st.s16_microsec_id, m.status_head), r[RK_VM_PLANT] = pack16(g_run.plant_ang[k], g_run.plant_rpm[k]);
  }
}
inline uint32_t &soa(uint32_t *blk, int64_t n, int64_t i, int w) { return blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }
} // namespace

void vTaskDelayUntil(TickType_t *, TickType_t) {
  const rk_vdt_rollout_t *a = g_run.a;
  if(a->task_period <= 0) throw LoopDone(); /* the loop below would never advance */
  if(g_run.period_k >= 0) { /* the ISR ticks of the period VDT::main just prepared */
    for(int j = 0; j < a->task_period && g_run.tick < a->steps; j++) run_isr_tick();
  }
  if(g_run.tick >= a->steps) throw LoopDone();
  g_run.period_k++;
}
size_t xMessageBufferReceive(MessageBufferHandle_t, void *dst, size_t bytes, uint32_t) {
  MSG_REQ m;
  if(bytes != sizeof(MSG_REQ) || !next_message(&m)) return 0;
  memcpy(dst, &m, sizeof(m));
  return sizeof(m);
}

extern "C" {
/* rk_vdt_rollout()'s contract on HOST arrays for RK_SENSOR_PLANT + task_period > 0, instances [i0, i1), each from
 * the power-on state (the task's objects are file-static: one instance at a time); final states into `state`.
 * d_yaw is in RADIANS like the library's; it is converted back to the degrees the IMU task reports with the
 * inverse the tests guarantee to be exact (whole degrees). */
void ref_vdt_task_rollout(uint32_t *state, int64_t n, int64_t i0, int64_t i1, const rk_vdt_rollout_t *args, const float *yaw_deg) {
  static bool prepared = false;
  if(!prepared) {
    prepare_task(); /* VD_task_main.cpp:153-163 */
    prepared = true;
  }
  rk_vdt_rollout_t a = *args;
  a.d_yaw            = yaw_deg; /* the ISR reads degrees and applies mymath::deg2rad itself (:368) */
  for(int64_t i = i0; i < i1; i++) {
    reset_task_objects();
    g_run = TaskRun();
    g_run.n = n, g_run.i = i, g_run.a = &a;
    try {
      VDT::main(nullptr);
    } catch(const LoopDone &) {
    }
    if(state) {
      uint32_t w[RK_VS_WORDS];
      export_task_state(w);
      for(int k = 0; k < RK_VS_WORDS; k++) soa(state, n, i, k) = w[k];
    }
  }
}
}
