/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * C-ABI harness around the UNMODIFIED reference arm sources
 *   src/ArmDrive/AD_mode_positioning_seq.{hpp,cpp}, AD_mode_positioning.{hpp,cpp}, AD_mode_base.hpp, AD_joint_base.hpp,
 *   AD_joint_dfgear.hpp, AD_joint_mybldc_servo.{hpp,cpp}, AD_joint_mg_servo.{hpp,cpp},
 *   AD_joint_ics_servo.{hpp,cpp}, lib/IcsClass_V210/src/IcsBaseClass.{h,cpp}
 * compiled where they lie (oracle/Makefile -> oracle/_ref/libref_arm.so).  Joint objects are
 * wired exactly as AD_task_main.cpp:38-116,148-149; one tick is ADT::main's loop body
 * (:208-229) with the CAN controllers' tx_routine() (AD_can_controller_mybldc.hpp:43-52,
 * AD_can_controller_mg.hpp:43-55) reduced to "take the frame".  ADTModeBase::P_JOINT_ is a
 * class static, so ONE arm is live at a time; batches run instance after instance.
 * The Kondo ICS UART is the ideal servo of oracle/stubs/IcsHardSerialClass.h.
 */
#include <new>
#include <stdlib.h>
#include <string.h>

#include "ArmDrive/AD_joint_dfgear.hpp"
#include "ArmDrive/AD_joint_ics_servo.hpp"
#include "ArmDrive/AD_joint_mg_servo.hpp"
#include "ArmDrive/AD_joint_mybldc_servo.hpp"
#include "ArmDrive/AD_mode_initialize.hpp"
#include "ArmDrive/AD_mode_initpos_move.hpp"
#include "ArmDrive/AD_mode_positioning.hpp"
#include "ArmDrive/AD_mode_positioning_seq.hpp"
/* the three in-tree command sequences (POS_CMD_SEQ_DEBUG_0/1/2 have internal linkage, so the
 * reference translation unit is included where it lies -- nothing is copied) */
#include "ArmDrive/AD_mode_positioning_seq_debug_data.cpp"

#include "robotick.h"

HardwareSerial Serial6;
HardwareSerial Serial7;
uint32_t       get_gptimer_cnt() { return 0; }
namespace DEBUG {
char EXT_PRINT_BUF[1024];
void print(char *, uint32_t) {}
void record_proc_load(uint8_t, uint8_t) {}
} // namespace DEBUG
namespace LGT {
void push_buffer(char *, uint32_t) {}
} // namespace LGT

/* IcsBaseClass.h:144-147 declares this virtual and asks for it to be "written externally"; it is
 * the class's key function, so the vtable lives wherever it is defined.  The stub subclass
 * overrides it; this base version is never called. */
bool IcsBaseClass::synchronize(byte *, byte, byte *, byte) { return false; }

namespace ADT {
/* defined in AD_task_main.cpp:148-149 (not compiled: FreeRTOS / FlexCAN) */
JointBase *ADTModeBase::P_JOINT_[JointAxis::J_NUM] = {nullptr, nullptr, nullptr, nullptr, nullptr};
float      ADTModeBase::FL_CYCLE_TIME_S            = 0.01f;
} // namespace ADT

namespace {
using namespace ADT;

/* AD_task_main.cpp:38-107 */
const JointBase::ConstParams CP[RK_AJ_NUM] = {
    {0.01f, 1.0f, -1.0f, 3.0f, -45.0f, 15.0f, 1.0f, 0.0f},          /* j_Y0 */
    {0.01f, 1.0f, 1.0f, 0.7f, 150.0f, 30.0f, 0.15f, 145.0f},        /* j_P1 */
    {0.01f, 1.0f, 1.0f, 0.5f, 0.0f, 10.0f, 0.5f, 0.0f},             /* j_DF_Left */
    {0.01f, 1.0f, 1.0f, 0.5f, 0.0f, 10.0f, 0.5f, 0.0f},             /* j_DF_Right */
    {0.01f, 24.0f / 7.0f, 1.0f, 1.0f, 0.0f, 30.0f, 1.0f, -90.0f},   /* j_DF_Pt */
    {0.01f, 48.0f / 7.0f, 1.0f, 1.0f, 0.0f, 30.0f, 1.0f, 0.0f},     /* j_DF_Rl */
    {0.01f, 48.0f / 19.0f, -1.0f, 0.8f, -90.0f, -60.0f, 0.5f, 0.0f} /* j_P3 */
};

struct ArmSet {
  JointBase::ConstParams cp[RK_AJ_NUM];
  IcsHardSerialClass     ics;
  JointIcsServo          j_Y0;
  JointMgServo           j_P1;
  JointMyBldcServo       j_DFL, j_DFR;
  JointDfGearVirtual     dfv;
  JointDfGearPitch       j_P2;
  JointDfGearRoll        j_R0;
  JointMyBldcServo       j_P3;
  ADTModePositioningSeq  posseq;
  ADTModePositioning     pos; /* the single-command mode (REQ_MOVE_POS) on the same joints */
  ADTModeInitialize      m_init;    /* homing: INIT           (AD_task_main.cpp:152) */
  ADTModeInitPosMove     m_initpos; /* homing: INIT_POS_MOVE  (:153) */
  /* what the CAN tx routines / the UART took this tick */
  uint8_t  mg_tx[8];
  int      mg_valid;
  uint8_t  bldc_tx[3][8];
  uint32_t bldc_id[3];
  int      bldc_valid[3];

  ArmSet()
      : cp{CP[0], CP[1], CP[2], CP[3], CP[4], CP[5], CP[6]}, j_Y0(cp[0]), j_P1(cp[1]), j_DFL(cp[2], 1), j_DFR(cp[3], 2),
        dfv(j_DFL, j_DFR), j_P2(cp[4], dfv), j_R0(cp[5], dfv), j_P3(cp[6], 3) {}

  JointBase *jb(int k) {
    JointBase *t[RK_AJ_NUM] = {&j_Y0, &j_P1, &j_DFL, &j_DFR, &j_P2, &j_R0, &j_P3};
    return t[k];
  }
  JointMyBldcServo *bl(int k) {
    JointMyBldcServo *t[3] = {&j_DFL, &j_DFR, &j_P3};
    return t[k];
  }
  void bind() { /* AD_task_main.cpp:148 */
    ADTModeBase::P_JOINT_[0] = &j_Y0, ADTModeBase::P_JOINT_[1] = &j_P1, ADTModeBase::P_JOINT_[2] = &j_P2;
    ADTModeBase::P_JOINT_[3] = &j_R0, ADTModeBase::P_JOINT_[4] = &j_P3;
  }
};

ArmSet *make() {
  void   *mem = calloc(1, sizeof(ArmSet));
  ArmSet *s   = new(mem) ArmSet();
  s->bind();
  return s;
}

/* prepare_task() (AD_task_main.cpp:170-193) + what a completed INIT mode leaves behind
 * (AD_mode_initialize.cpp:58-63,133-135) + set_next_mode(POSITIONING_SEQ) -> init() (:316-320) */
void bringup(ArmSet *s) {
  s->bind();
  s->j_Y0.doinit(&s->ics, 0);
  s->j_Y0.set_torque_on(false);
  s->j_P1.set_torque_on(false);
  s->j_P1.init();
  for(int i = 0; i < JointAxis::J_NUM; i++) {
    JointBase *j = ADTModeBase::P_JOINT_[i];
    j->set_torque_on(true);
    j->set_initilized(true);
    j->set_curlim_A(j->get_curlim_default_A());
  }
  s->posseq.init();
}

/* ADT::main loop body  AD_task_main.cpp:208-229 ; which = the active mode object */
void tick(ArmSet *s, int which = 0) {
  if(which == 0) s->posseq.update();
  else if(which == 1) s->pos.update();
  else if(which == 2) s->m_init.update();
  else s->m_initpos.update();
  s->j_P1.update();
  s->j_DFL.update();
  s->j_DFR.update();
  s->j_P3.update();
  /* MG_CAN.tx_routine(): tx1 if updated, else tx2 */
  uint8_t tmp[8];
  s->mg_valid = s->j_P1.get_cantx1_data(s->mg_tx) ? 1 : 0;
  if(!s->mg_valid) s->j_P1.get_cantx2_data(tmp);
  for(int k = 0; k < 3; k++) s->bldc_valid[k] = s->bl(k)->get_cantx_data(s->bldc_tx[k], s->bldc_id[k]) ? 1 : 0;
  s->j_Y0.update();
}

inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint32_t ld32(const uint8_t *p) {
  uint32_t u;
  memcpy(&u, p, 4);
  return u;
}

int ics_pos_word(ArmSet *s) { /* position word of the last setPos frame, -1 when the last frame was not one */
  const byte *t = s->ics.last_tx;
  if((t[0] & 0xE0) != 0x80 || (t[1] == 0 && t[2] == 0)) return -1;
  return ((int)t[1] << 7) | t[2];
}

void export_state(ArmSet *s, uint32_t *w) {
  memset(w, 0, 4 * RK_AS_WORDS);
  auto &m = s->posseq;
  w[RK_AS_FSM] = (uint32_t)m.nowState | (m.isModeFirstCall ? RK_AS_FSM_FIRSTCALL : 0u) | (m.is_comp ? RK_AS_FSM_IS_COMP : 0u);
  w[RK_AS_SEQ_IDX]  = (uint32_t)m.u16_seq_exec_idx_ | ((uint32_t)m.u16_seq_write_head_ << 16);
  w[RK_AS_CMD_IDX]  = m.u8_nowcmd_idx_;
  w[RK_AS_MOVE_CNT] = (uint32_t)m.s32_move_cnt_;
  w[RK_AS_CYCLE]    = (uint32_t)m.s32_cycle_counter_;
  w[RK_AS_TOTAL_MS] = m.u32_total_move_ms_;
  w[RK_AS_NOW_DT]   = m.now_cmd_.u32_dt_ms;
  for(int j = 0; j < 5; j++) w[RK_AS_NOW_TGT + j] = f2u(m.now_cmd_.fl_tgt_pos_deg[j]), w[RK_AS_MOVE_DEG + j] = f2u(m.fl_move_deg_[j]);
  w[RK_AS_DFV_P] = f2u(s->dfv.fl_rawP_tgt_deg_);
  w[RK_AS_DFV_R] = f2u(s->dfv.fl_rawR_tgt_deg_);
  uint32_t fl = 0;
  for(int k = 0; k < RK_AJ_NUM; k++) {
    JointBase *j = s->jb(k);
    uint32_t  *q = w + RK_AS_JOINT0 + 4 * k;
    q[RK_AJ_OFS] = f2u(j->fl_out_ofs_deg), q[RK_AJ_RAW_TGT] = f2u(j->fl_raw_tgt_deg);
    q[RK_AJ_CURLIM] = f2u(j->fl_curlim_A), q[RK_AJ_RAW_NOW] = f2u(j->fl_raw_now_deg);
    uint32_t b = (j->is_connected ? RK_AJF_CONNECTED : 0u) | (j->is_torque_on ? RK_AJF_TORQUE_ON : 0u) |
                 (j->is_initialized ? RK_AJF_INITIALIZED : 0u);
    if(k == RK_AJ_P1) b |= s->j_P1.is_torque_on_prev ? RK_AJF_TORQUE_PREV : 0u;
    if(k == RK_AJ_DFL) b |= s->j_DFL.is_torque_on_prev ? RK_AJF_TORQUE_PREV : 0u;
    if(k == RK_AJ_DFR) b |= s->j_DFR.is_torque_on_prev ? RK_AJF_TORQUE_PREV : 0u;
    if(k == RK_AJ_P3) b |= s->j_P3.is_torque_on_prev ? RK_AJF_TORQUE_PREV : 0u;
    fl |= b << (4 * k);
  }
  w[RK_AS_JFLAGS]     = fl;
  w[RK_AS_MG_PRE_TGT] = f2u(s->j_P1.fl_pre_raw_tgt_deg);
  w[RK_AS_ICS_POS]    = (uint32_t)ics_pos_word(s);
  w[RK_AS_ICS_SERVO]  = (uint32_t)((((int)s->ics.pos_h << 7) | s->ics.pos_l) - 7500);
  w[RK_AS_MG_TX] = ld32(s->mg_tx), w[RK_AS_MG_TX + 1] = ld32(s->mg_tx + 4), w[RK_AS_MG_TX + 2] = (uint32_t)s->mg_valid;
  { /* JointMgServo::pos_ctrl_ (UTIL::PI_D); word 7 = "InitGain applied" (the only gain set the firmware ever loads) */
    auto     &c = s->j_P1.pos_ctrl_;
    uint32_t *q = w + RK_AS_MG_CTRL;
    q[0] = f2u(c.prev_val_), q[1] = f2u(c.Integ_), q[2] = f2u(c.velLpf_.now_Y_), q[3] = f2u(c.velLpf_.prev_X_);
    q[4] = f2u(c.now_tgt_), q[5] = f2u(c.now_error_), q[6] = f2u(c.now_ctrl_), q[7] = (c.Pgain_ != 0.0f) ? 1u : 0u;
  }
  for(int k = 0; k < 3; k++) {
    uint32_t *q = w + RK_AS_BLDC_TX0 + 4 * k;
    q[0] = ld32(s->bldc_tx[k]), q[1] = ld32(s->bldc_tx[k] + 4), q[2] = s->bl(k)->u32_txcmdid, q[3] = (uint32_t)s->bldc_valid[k];
  }
}

void import_state(ArmSet *s, const uint32_t *w) {
  auto &m           = s->posseq;
  m.nowState        = (ADTModePositioningSeq::State)(w[RK_AS_FSM] & 0xFF);
  m.isModeFirstCall = (w[RK_AS_FSM] & RK_AS_FSM_FIRSTCALL) != 0;
  m.is_comp         = (w[RK_AS_FSM] & RK_AS_FSM_IS_COMP) != 0;
  m.u16_seq_exec_idx_   = (uint16_t)(w[RK_AS_SEQ_IDX] & 0xFFFF);
  m.u16_seq_write_head_ = (uint16_t)(w[RK_AS_SEQ_IDX] >> 16);
  m.u8_nowcmd_idx_      = (uint8_t)w[RK_AS_CMD_IDX];
  m.s32_move_cnt_       = (int32_t)w[RK_AS_MOVE_CNT];
  m.s32_cycle_counter_  = (int32_t)w[RK_AS_CYCLE];
  m.u32_total_move_ms_  = w[RK_AS_TOTAL_MS];
  m.now_cmd_.u32_dt_ms  = w[RK_AS_NOW_DT];
  for(int j = 0; j < 5; j++) m.now_cmd_.fl_tgt_pos_deg[j] = u2f(w[RK_AS_NOW_TGT + j]), m.fl_move_deg_[j] = u2f(w[RK_AS_MOVE_DEG + j]);
  s->dfv.fl_rawP_tgt_deg_ = u2f(w[RK_AS_DFV_P]);
  s->dfv.fl_rawR_tgt_deg_ = u2f(w[RK_AS_DFV_R]);
  for(int k = 0; k < RK_AJ_NUM; k++) {
    JointBase      *j = s->jb(k);
    const uint32_t *q = w + RK_AS_JOINT0 + 4 * k;
    j->fl_out_ofs_deg = u2f(q[RK_AJ_OFS]), j->fl_raw_tgt_deg = u2f(q[RK_AJ_RAW_TGT]);
    j->fl_curlim_A = u2f(q[RK_AJ_CURLIM]), j->fl_raw_now_deg = u2f(q[RK_AJ_RAW_NOW]);
    uint32_t b        = (w[RK_AS_JFLAGS] >> (4 * k)) & 0xF;
    j->is_connected   = (b & RK_AJF_CONNECTED) != 0;
    j->is_torque_on   = (b & RK_AJF_TORQUE_ON) != 0;
    j->is_initialized = (b & RK_AJF_INITIALIZED) != 0;
    bool prev         = (b & RK_AJF_TORQUE_PREV) != 0;
    if(k == RK_AJ_P1) s->j_P1.is_torque_on_prev = prev;
    if(k == RK_AJ_DFL) s->j_DFL.is_torque_on_prev = prev;
    if(k == RK_AJ_DFR) s->j_DFR.is_torque_on_prev = prev;
    if(k == RK_AJ_P3) s->j_P3.is_torque_on_prev = prev;
  }
  s->j_P1.fl_pre_raw_tgt_deg = u2f(w[RK_AS_MG_PRE_TGT]);
  {
    auto           &c = s->j_P1.pos_ctrl_;
    const uint32_t *q = w + RK_AS_MG_CTRL;
    c.prev_val_ = c.now_val_ = u2f(q[0]), c.Integ_ = u2f(q[1]);
    c.velLpf_.now_Y_ = c.velLpf_.prev_Y_ = u2f(q[2]), c.velLpf_.prev_X_ = u2f(q[3]);
    c.now_tgt_ = u2f(q[4]), c.now_error_ = c.prev_error_ = u2f(q[5]), c.now_ctrl_ = u2f(q[6]);
    c.Pgain_ = q[7] ? 0.01f : 0.0f, c.Igain_ = 0.0f, c.Dgain_ = 0.0f, c.I_limit_ = 0.0f;
  }
  int sp                     = (int)(int32_t)w[RK_AS_ICS_SERVO] + 7500;
  s->ics.pos_h = (byte)((sp >> 7) & 0x7F), s->ics.pos_l = (byte)(sp & 0x7F);
  s->j_Y0.p_ics_serial = &s->ics;
  s->j_Y0.u8_id        = 0;
  int pw               = (int)(int32_t)w[RK_AS_ICS_POS];
  if(pw >= 0) s->ics.last_tx[0] = 0x80, s->ics.last_tx[1] = (byte)((pw >> 7) & 0x7F), s->ics.last_tx[2] = (byte)(pw & 0x7F);
  else s->ics.last_tx[0] = 0, s->ics.last_tx[1] = 0, s->ics.last_tx[2] = 0;
  memcpy(s->mg_tx, &w[RK_AS_MG_TX], 8);
  s->mg_valid = (int)w[RK_AS_MG_TX + 2];
  for(int k = 0; k < 3; k++) {
    const uint32_t *q = w + RK_AS_BLDC_TX0 + 4 * k;
    memcpy(s->bldc_tx[k], q, 8);
    memcpy(s->bl(k)->txmsg.u8_data, q, 8);
    s->bl(k)->u32_txcmdid = q[2];
    s->bldc_id[k]         = q[2];
    s->bldc_valid[k]      = (int)q[3];
  }
}

void export_pstate(ArmSet *s, uint32_t *w) {
  memset(w, 0, 4 * RK_PS_WORDS);
  auto &m = s->pos;
  w[RK_PS_STATE]    = (uint32_t)m.nowState | (m.is_comp ? RK_AS_FSM_IS_COMP : 0u);
  w[RK_PS_MOVE_CNT] = m.u32_move_cnt_, w[RK_PS_CYCLE] = m.u32_cycle_counter_, w[RK_PS_QSIZE] = (uint32_t)m.cmd_q_.size();
  w[RK_PS_PREV_ID0] = m.u32_prev_cmd_id_[0], w[RK_PS_PREV_ID1] = m.u32_prev_cmd_id_[1];
  w[RK_PS_NOW_CMD] = m.now_cmd_.u32_id, w[RK_PS_NOW_CMD + 1] = m.now_cmd_.u32_dt_ms;
  for(int j = 0; j < 5; j++) w[RK_PS_NOW_CMD + 2 + j] = f2u(m.now_cmd_.fl_tgt_pos_deg[j]), w[RK_PS_MOVE_DEG + j] = f2u(m.fl_move_deg_[j]);
  int e = 0;
  for(auto it = m.cmd_q_.cbegin(); it != m.cmd_q_.cend() && e < 4; ++it, ++e) {
    uint32_t *q = w + RK_PS_QUEUE + 8 * e;
    q[0] = it->u32_id, q[1] = it->u32_dt_ms;
    for(int j = 0; j < 5; j++) q[2 + j] = f2u(it->fl_tgt_pos_deg[j]);
  }
}
void import_pstate(ArmSet *s, const uint32_t *w) {
  auto &m    = s->pos;
  m.nowState = (ADTModePositioning::State)(w[RK_PS_STATE] & 0xFF);
  m.is_comp  = (w[RK_PS_STATE] & RK_AS_FSM_IS_COMP) != 0;
  m.u32_move_cnt_ = w[RK_PS_MOVE_CNT], m.u32_cycle_counter_ = w[RK_PS_CYCLE];
  m.u32_prev_cmd_id_[0] = w[RK_PS_PREV_ID0], m.u32_prev_cmd_id_[1] = w[RK_PS_PREV_ID1];
  m.now_cmd_.u32_id = w[RK_PS_NOW_CMD], m.now_cmd_.u32_dt_ms = w[RK_PS_NOW_CMD + 1];
  for(int j = 0; j < 5; j++) m.now_cmd_.fl_tgt_pos_deg[j] = u2f(w[RK_PS_NOW_CMD + 2 + j]), m.fl_move_deg_[j] = u2f(w[RK_PS_MOVE_DEG + j]);
  m.cmd_q_.clear();
  for(uint32_t e = 0; e < w[RK_PS_QSIZE] && e < 4; e++) {
    ADTModePositioning::PosCmd c;
    const uint32_t            *q = w + RK_PS_QUEUE + 8 * e;
    c.u32_id = q[0], c.u32_dt_ms = q[1];
    for(int j = 0; j < 5; j++) c.fl_tgt_pos_deg[j] = u2f(q[2 + j]);
    m.cmd_q_.push_back(c);
  }
}

void export_hstate(ArmSet *s, int mode, uint32_t *w) {
  memset(w, 0, 4 * RK_HS_WORDS);
  if(mode == RK_ADH_MODE_INIT) {
    auto &m = s->m_init;
    w[RK_HS_STATE] = (uint32_t)m.nowState | (m.is_comp ? RK_AS_FSM_IS_COMP : 0u) | ((uint32_t)mode << 16);
    w[RK_HS_WAIT_CNT] = m.u16_wait_cnt_;
  } else {
    auto &m = s->m_initpos;
    w[RK_HS_STATE] = (uint32_t)m.nowState | (m.is_comp ? RK_AS_FSM_IS_COMP : 0u) | ((uint32_t)mode << 16);
    w[RK_HS_WAIT_CNT] = m.u16_wait_cnt_;
    for(int j = 0; j < 5; j++) w[RK_HS_VEL_DIR + j] = f2u(m.fl_move_vel_dir_[j]);
  }
}
int import_hstate(ArmSet *s, const uint32_t *w) {
  const int mode = (int)(w[RK_HS_STATE] >> 16);
  if(mode == RK_ADH_MODE_INIT) {
    auto &m = s->m_init;
    m.nowState = (ADTModeInitialize::State)(w[RK_HS_STATE] & 0xFF), m.is_comp = (w[RK_HS_STATE] & RK_AS_FSM_IS_COMP) != 0;
    m.u16_wait_cnt_ = (uint16_t)w[RK_HS_WAIT_CNT];
  } else {
    auto &m = s->m_initpos;
    m.nowState = (ADTModeInitPosMove::State)(w[RK_HS_STATE] & 0xFF), m.is_comp = (w[RK_HS_STATE] & RK_AS_FSM_IS_COMP) != 0;
    m.u16_wait_cnt_ = (uint16_t)w[RK_HS_WAIT_CNT];
    for(int j = 0; j < 5; j++) m.fl_move_vel_dir_[j] = u2f(w[RK_HS_VEL_DIR + j]);
  }
  return mode;
}

inline uint32_t &soa(uint32_t *blk, int64_t n, int64_t i, int w) { return blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }

void load_cmdtab(ArmSet *s, const uint32_t *tab, int64_t n, int64_t i) {
  for(int sl = 0; sl < RK_ACMD_SLOTS; sl++) {
    auto &q          = s->posseq.cmd_seq_[sl];
    int   b          = sl * RK_ACMD_SLOT_WORDS;
    q.u32_id         = soa((uint32_t *)tab, n, i, b + 0);
    q.u8_cmd_seq_len = (uint8_t)soa((uint32_t *)tab, n, i, b + 1);
    for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
      q.cmd_seq[k].u32_dt_ms = soa((uint32_t *)tab, n, i, b + 4 + 8 * k);
      for(int j = 0; j < 5; j++) q.cmd_seq[k].fl_tgt_pos_deg[j] = u2f(soa((uint32_t *)tab, n, i, b + 4 + 8 * k + 1 + j));
    }
  }
}
void store_cmdtab(ArmSet *s, uint32_t *tab, int64_t n, int64_t i) {
  for(int sl = 0; sl < RK_ACMD_SLOTS; sl++) {
    auto &q = s->posseq.cmd_seq_[sl];
    int   b = sl * RK_ACMD_SLOT_WORDS;
    soa(tab, n, i, b + 0) = q.u32_id, soa(tab, n, i, b + 1) = q.u8_cmd_seq_len, soa(tab, n, i, b + 2) = 0, soa(tab, n, i, b + 3) = 0;
    for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
      soa(tab, n, i, b + 4 + 8 * k) = q.cmd_seq[k].u32_dt_ms;
      for(int j = 0; j < 5; j++) soa(tab, n, i, b + 4 + 8 * k + 1 + j) = f2u(q.cmd_seq[k].fl_tgt_pos_deg[j]);
      soa(tab, n, i, b + 4 + 8 * k + 6) = 0, soa(tab, n, i, b + 4 + 8 * k + 7) = 0;
    }
  }
}

uint32_t bldc_id_byte(uint32_t id) { return (id & 0xFF) | ((id & 0x8000) ? 0x80u : 0u); }

void trace_row(ArmSet *s, uint32_t *tr, int64_t n, int which = 0) {
  for(int j = 0; j < 5; j++) tr[(int64_t)j * n] = f2u(ADTModeBase::P_JOINT_[j]->get_tgt_deg());
  uint16_t vl;
  int32_t  ang;
  memcpy(&vl, s->mg_tx + 2, 2), memcpy(&ang, s->mg_tx + 4, 4);
  tr[5 * n] = vl, tr[6 * n] = (uint32_t)ang;
  for(int k = 0; k < 3; k++) tr[(int64_t)(7 + k) * n] = ld32(s->bldc_tx[k]);
  tr[10 * n] = (uint32_t)ics_pos_word(s);
  tr[11 * n] = which == 0 ? (uint32_t)s->posseq.nowState : which == 1 ? (uint32_t)s->pos.nowState
               : which == 2 ? (uint32_t)s->m_init.nowState : (uint32_t)s->m_initpos.nowState;
  tr[12 * n] = which == 0 ? (uint32_t)s->posseq.u8_nowcmd_idx_ : which == 1 ? (uint32_t)s->pos.cmd_q_.size()
               : which == 2 ? (uint32_t)s->m_init.u16_wait_cnt_ : (uint32_t)s->m_initpos.u16_wait_cnt_;
  tr[13 * n] = bldc_id_byte(s->bl(0)->u32_txcmdid) | (bldc_id_byte(s->bl(1)->u32_txcmdid) << 8) | (bldc_id_byte(s->bl(2)->u32_txcmdid) << 16);
  tr[14 * n] = 0, tr[15 * n] = 0;
}

} // namespace

extern "C" {

void *ref_adt_create(void) { return make(); }
void  ref_adt_destroy(void *h) {
  if(!h) return;
  ((ArmSet *)h)->~ArmSet();
  free(h);
}
void ref_adt_bringup(void *h) { bringup((ArmSet *)h); }
int  ref_adt_push(void *h, const rk_adt_poscmdseq_t *seq) {
  ArmSet                          *s = (ArmSet *)h;
  ADTModePositioningSeq::PosCmdSeq q;
  memset(&q, 0, sizeof(q));
  q.u32_id = seq->id, q.u8_cmd_seq_len = seq->len;
  for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
    q.cmd_seq[k].u32_dt_ms = seq->cmd[k].dt_ms;
    for(int j = 0; j < 5; j++) q.cmd_seq[k].fl_tgt_pos_deg[j] = seq->cmd[k].tgt_deg[j];
  }
  uint16_t before = s->posseq.u16_seq_write_head_;
  s->bind();
  s->posseq.push_cmdseq(q);
  return s->posseq.u16_seq_write_head_ != before;
}
void ref_adt_tick(void *h) {
  ((ArmSet *)h)->bind();
  tick((ArmSet *)h);
}
int  ref_adt_status(void *h, uint32_t id) { return ((ArmSet *)h)->posseq.get_q_cmdseq_status(id); }
void ref_adt_targets(void *h, float out[5]) {
  ((ArmSet *)h)->bind();
  for(int j = 0; j < 5; j++) out[j] = ADTModeBase::P_JOINT_[j]->get_tgt_deg();
}
/* AD_mode_positioning_seq_debug_data.cpp:5-64 */
int ref_adt_debug_seq(int which, rk_adt_poscmdseq_t *out) {
  const ADTModePositioningSeq::PosCmdSeq *q = which == 0 ? &POS_CMD_SEQ_DEBUG_0 : which == 1 ? &POS_CMD_SEQ_DEBUG_1 : which == 2 ? &POS_CMD_SEQ_DEBUG_2 : nullptr;
  if(!q) return -1;
  memset(out, 0, sizeof(*out));
  out->id = q->u32_id, out->len = q->u8_cmd_seq_len;
  for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
    out->cmd[k].dt_ms = q->cmd_seq[k].u32_dt_ms;
    for(int j = 0; j < 5; j++) out->cmd[k].tgt_deg[j] = q->cmd_seq[k].fl_tgt_pos_deg[j];
  }
  return get_poscmdseq_debug() == q ? 1 : 0;
}
void ref_adt_export(void *h, uint32_t *w) { export_state((ArmSet *)h, w); }
void ref_adt_import(void *h, const uint32_t *w) { import_state((ArmSet *)h, w); }
void ref_adt_trace_row(void *h, uint32_t *row16) { trace_row((ArmSet *)h, row16, 1); }

/* Batch driver on HOST arrays, same contracts as the rk_adt_* batch calls.
 * op: 0 = bring-up (rk_adt_mode_init), 1 = push d_seq (per-instance slot image, valid mask),
 *     2 = K ticks (+trace), 3 = status query into status[] for ids[] */
void ref_adt_batch(int op, uint32_t *state, uint32_t *cmdtab, int64_t n, int64_t i0, int64_t i1, int K,
                   const uint32_t *seq, const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status) {
  for(int64_t i = i0; i < i1; i++) {
    ArmSet  *s = make();
    uint32_t w[RK_AS_WORDS];
    for(int k = 0; k < RK_AS_WORDS; k++) w[k] = soa(state, n, i, k);
    import_state(s, w);
    if(cmdtab) load_cmdtab(s, cmdtab, n, i);
    if(op == 0) {
      bringup(s);
    } else if(op == 1) {
      if(!valid || valid[i]) {
        ADTModePositioningSeq::PosCmdSeq q;
        memset(&q, 0, sizeof(q));
        q.u32_id         = soa((uint32_t *)seq, n, i, 0);
        q.u8_cmd_seq_len = (uint8_t)soa((uint32_t *)seq, n, i, 1);
        for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
          q.cmd_seq[k].u32_dt_ms = soa((uint32_t *)seq, n, i, 4 + 8 * k);
          for(int j = 0; j < 5; j++) q.cmd_seq[k].fl_tgt_pos_deg[j] = u2f(soa((uint32_t *)seq, n, i, 4 + 8 * k + 1 + j));
        }
        s->posseq.push_cmdseq(q);
      }
    } else if(op == 2) {
      for(int t = 0; t < K; t++) {
        tick(s);
        if(trace) trace_row(s, trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i, n);
      }
    } else if(op == 3) {
      status[i] = s->posseq.get_q_cmdseq_status(ids[i]);
    }
    if(op != 3) {
      export_state(s, w);
      for(int k = 0; k < RK_AS_WORDS; k++) soa(state, n, i, k) = w[k];
      if(cmdtab && op == 1) store_cmdtab(s, cmdtab, n, i);
    }
    s->~ArmSet();
    free(s);
  }
}

/* Servo feedback frames through the reference's own CAN rx callbacks (arm side), same contract as
 * rk_adt_bldc_rx / rk_adt_mg_rx: kind 0..2 = JointMyBldcServo::rx_callback of DF_Left / DF_Right / P3
 * (AD_joint_mybldc_servo.cpp:45-70; cmdid NULL = CMD_ID_RES_STATUS_SUMMARY), kind 3 = JointMgServo::rx_callback
 * (AD_joint_mg_servo.cpp:75-92).  cur[i] receives fl_out_now_cur when the frame carries a current, else stays. */
void ref_adt_rx_batch(int kind, uint32_t *state, int64_t n, int64_t i0, int64_t i1, const uint64_t *frames, const uint32_t *cmdid,
                      float *cur) {
  for(int64_t i = i0; i < i1; i++) {
    ArmSet  *s = make();
    uint32_t w[RK_AS_WORDS];
    for(int k = 0; k < RK_AS_WORDS; k++) w[k] = soa(state, n, i, k);
    import_state(s, w);
    const uint32_t SENT = 0x7FC12345u; /* a NaN no arithmetic produces: "not written" */
    JointBase     *j    = (kind == 3) ? (JointBase *)&s->j_P1 : (JointBase *)s->bl(kind);
    memcpy(&j->fl_out_now_cur, &SENT, 4);
    uint64_t f = frames[i];
    if(kind == 3) s->j_P1.rx_callback((JointMgServo::MgMsgRx *)&f, 0);
    else s->bl(kind)->rx_callback(cmdid ? cmdid[i] : (uint32_t)CMD_ID_RES_STATUS_SUMMARY, (RES_MESSAGE *)&f, 0);
    uint32_t cb;
    memcpy(&cb, &j->fl_out_now_cur, 4);
    if(cur && cb != SENT) memcpy(&cur[i], &cb, 4);
    export_state(s, w);
    for(int k = 0; k < RK_AS_WORDS; k++) soa(state, n, i, k) = w[k];
    s->~ArmSet();
    free(s);
  }
}

/* ADTModePositioning batch driver on HOST arrays, same contracts as the rk_adp_* calls:
 * op 0 = rk_adp_mode_init, 1 = rk_adp_push_cmd (cmd: two planes per arm), 2 = K ticks (+trace), 3 = status */
void ref_adp_batch(int op, uint32_t *state, uint32_t *pstate, int64_t n, int64_t i0, int64_t i1, int K, const uint32_t *cmd,
                   const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status) {
  for(int64_t i = i0; i < i1; i++) {
    ArmSet  *s = make();
    uint32_t w[RK_AS_WORDS], pw[RK_PS_WORDS];
    for(int k = 0; k < RK_AS_WORDS; k++) w[k] = soa(state, n, i, k);
    for(int k = 0; k < RK_PS_WORDS; k++) pw[k] = soa(pstate, n, i, k);
    import_state(s, w);
    import_pstate(s, pw);
    if(op == 0) {
      s->pos.init();
    } else if(op == 1) {
      if(!valid || valid[i]) {
        ADTModePositioning::PosCmd c;
        c.u32_id = soa((uint32_t *)cmd, n, i, 0), c.u32_dt_ms = soa((uint32_t *)cmd, n, i, 1);
        for(int j = 0; j < 5; j++) c.fl_tgt_pos_deg[j] = u2f(soa((uint32_t *)cmd, n, i, 2 + j));
        s->pos.push_cmd(c);
      }
    } else if(op == 2) {
      for(int t = 0; t < K; t++) {
        tick(s, 1);
        if(trace) trace_row(s, trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i, n, 1);
      }
    } else if(op == 3) {
      status[i] = s->pos.get_q_cmd_status(ids[i]);
    }
    if(op != 3) {
      export_state(s, w);
      export_pstate(s, pw);
      if(op == 2)
        for(int k = 0; k < RK_AS_WORDS; k++) soa(state, n, i, k) = w[k];
      for(int k = 0; k < RK_PS_WORDS; k++) soa(pstate, n, i, k) = pw[k];
    }
    s->~ArmSet();
    free(s);
  }
}

/* Homing modes on HOST arrays, same contracts as rk_adh_mode_init (op 0, mode in K) / rk_adh_update (op 2):
 * the servo feedback stream `now` ([K][4][n]: P1, DF_Left, DF_Right, P3) lands where the CAN rx callbacks store it */
void ref_adh_batch(int op, uint32_t *state, uint32_t *hstate, int64_t n, int64_t i0, int64_t i1, int K, const float *now,
                   uint32_t *trace) {
  for(int64_t i = i0; i < i1; i++) {
    ArmSet  *s = make();
    uint32_t w[RK_AS_WORDS], hw[RK_HS_WORDS];
    for(int k = 0; k < RK_AS_WORDS; k++) w[k] = soa(state, n, i, k);
    for(int k = 0; k < RK_HS_WORDS; k++) hw[k] = soa(hstate, n, i, k);
    import_state(s, w);
    int mode = import_hstate(s, hw);
    if(op == 0) {
      mode = K;
      if(mode == RK_ADH_MODE_INIT) s->m_init.init();
      else s->m_initpos.init();
    } else {
      JointBase *fb[4] = {&s->j_P1, &s->j_DFL, &s->j_DFR, &s->j_P3};
      for(int t = 0; t < K; t++) {
        if(now)
          for(int k = 0; k < 4; k++) fb[k]->fl_raw_now_deg = now[((int64_t)t * 4 + k) * n + i];
        tick(s, mode == RK_ADH_MODE_INIT ? 2 : 3);
        if(trace) trace_row(s, trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i, n, mode == RK_ADH_MODE_INIT ? 2 : 3);
      }
      export_state(s, w);
      for(int k = 0; k < RK_AS_WORDS; k++) soa(state, n, i, k) = w[k];
    }
    export_hstate(s, mode, hw);
    for(int k = 0; k < RK_HS_WORDS; k++) soa(hstate, n, i, k) = hw[k];
    s->~ArmSet();
    free(s);
  }
}
}
