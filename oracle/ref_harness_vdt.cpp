/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * C-ABI harness around the UNMODIFIED reference vehicle sources
 *   src/VehicleDrive/VD_vehicle_controller.{hpp,cpp}, VD_motor_if_m2006.{hpp,cpp},
 *   src/Utility/util_{controller,iir,vel_interp,mymath}.hpp (+ util_mymath.cpp)
 * compiled where they lie under /root/reference by oracle/Makefile into oracle/_ref/.
 * Built with -fno-access-control so the harness can export/import the private state in the
 * word order of include/robotick.h; the reference translation units themselves are compiled
 * untouched.
 *
 * Object wiring follows VD_task_main.cpp:75-108 and prepare_task() :157-160.  The reference
 * relies on zero-initialised static storage (finding 6): every object
 * set is placement-constructed into calloc()ed memory, which is exactly static
 * initialisation (zero-fill, then constructors).
 */
#include <new>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

#include "VehicleDrive/VD_vehicle_controller.hpp"
#include "Utility/util_mymath.hpp"

#include "robotick.h"

HardwareSerial Serial6;
HardwareSerial Serial7;
uint32_t       get_gptimer_cnt() { return 0; }

/* global_config.hpp declares these; nothing on the path calls them (macros compiled out). */
namespace DEBUG {
char EXT_PRINT_BUF[1024];
void print(char *, uint32_t) {}
void record_proc_load(uint8_t, uint8_t) {}
} // namespace DEBUG
namespace LGT {
void push_buffer(char *, uint32_t) {}
} // namespace LGT

namespace {

using VDT::Direction;
using VDT::MOTOR_IF_M2006;
using VDT::VEHICLE_CTRL;

struct VehicleSet {
  MOTOR_IF_M2006           motor[4];
  UTIL::FF_PI_D            ctrl[4];
  UTIL::VelInterpConstJerk interp[3];
  VEHICLE_CTRL::Parts      parts;
  VEHICLE_CTRL             vhcl;
  /* synthetic plant (robotick.h RK_VM_PLANT), motor frame */
  int32_t plant_rpm[4];
  int32_t plant_ang[4];

  /* only ever placement-constructed into zeroed memory */
  VehicleSet()
      : motor{MOTOR_IF_M2006(1), MOTOR_IF_M2006(1), MOTOR_IF_M2006(-1), MOTOR_IF_M2006(-1)},
        ctrl{UTIL::FF_PI_D(100.0f, 0.0075f, 0.02f, 0.01f, 0.0f, 0.5f, 10.0f),
             UTIL::FF_PI_D(100.0f, 0.0075f, 0.02f, 0.01f, 0.0f, 0.5f, 10.0f),
             UTIL::FF_PI_D(100.0f, 0.0075f, 0.02f, 0.01f, 0.0f, 0.5f, 10.0f),
             UTIL::FF_PI_D(100.0f, 0.0075f, 0.02f, 0.01f, 0.0f, 0.5f, 10.0f)},
        interp{UTIL::VelInterpConstJerk(1.0f / 1000.0f), UTIL::VelInterpConstJerk(1.0f / 1000.0f),
               UTIL::VelInterpConstJerk(1.0f / 1000.0f)},
        parts{nullptr,
              {&interp[0], &interp[1], &interp[2]},
              {&motor[0], &motor[1], &motor[2], &motor[3]},
              {&ctrl[0], &ctrl[1], &ctrl[2], &ctrl[3]}},
        vhcl(parts) {
    for(int w = 0; w < 4; w++) ctrl[w].set_FF_limit(1.0f);
  }
};

/* NOTE: copying a MOTOR_IF_M2006 / FF_PI_D temporary into the array element copies every
 * member; members without initialisers are copied from the (indeterminate) temporary.  To
 * keep the "zero-initialised static" contract we re-zero exactly those members after
 * construction -- they are the ones SURVEY.md finding 6 lists. */
void zero_uninitialised(VehicleSet *s) {
  for(int w = 0; w < 4; w++) {
    s->motor[w].s16_rawCurr_tgt = 0;
    memset(s->motor[w].status_buf, 0, sizeof(s->motor[w].status_buf));
  }
  for(int a = 0; a < 3; a++) {
    s->interp[a].vel_now_ = 0.0f;
    s->interp[a].acl_now_ = 0.0f;
    memset(s->interp[a].sts, 0, sizeof(s->interp[a].sts));
  }
  s->vhcl.now_vhcl_pos_m_       = Direction{0, 0, 0};
  s->vhcl.now_vhcl_vel_mmps     = Direction{0, 0, 0};
  s->vhcl.now_vhcl_vel_tgt_mmps = Direction{0, 0, 0};
  memset(s->vhcl.s64_rawAngleSumPrev, 0, sizeof(s->vhcl.s64_rawAngleSumPrev));
  s->vhcl.isPowerOn = false;
  memset(&s->vhcl.now_imu_data_, 0, sizeof(s->vhcl.now_imu_data_));
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
inline uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
inline int      lo16(uint32_t w) { return (int16_t)(w & 0xFFFFu); }
inline int      hi16(uint32_t w) { return (int16_t)(w >> 16); }

void export_state(VehicleSet *s, uint32_t *w) {
  memset(w, 0, sizeof(uint32_t) * RK_VS_WORDS);
  w[RK_VS_POS_X]  = f2u(s->vhcl.now_vhcl_pos_m_.x);
  w[RK_VS_POS_Y]  = f2u(s->vhcl.now_vhcl_pos_m_.y);
  w[RK_VS_POS_TH] = f2u(s->vhcl.now_vhcl_pos_m_.th);
  w[RK_VS_FLAGS]  = s->vhcl.isPowerOn ? RK_VS_FLAG_POWER_ON : 0u;
  w[RK_VS_VEL_X]  = f2u(s->vhcl.now_vhcl_vel_mmps.x);
  w[RK_VS_VEL_Y]  = f2u(s->vhcl.now_vhcl_vel_mmps.y);
  w[RK_VS_VEL_TH] = f2u(s->vhcl.now_vhcl_vel_mmps.th);
  w[RK_VS_TGT_X]  = f2u(s->vhcl.now_vhcl_vel_tgt_mmps.x);
  w[RK_VS_TGT_Y]  = f2u(s->vhcl.now_vhcl_vel_tgt_mmps.y);
  w[RK_VS_TGT_TH] = f2u(s->vhcl.now_vhcl_vel_tgt_mmps.th);
  for(int a = 0; a < 3; a++) {
    uint32_t *q  = w + RK_VS_INTERP0 + 12 * a;
    auto     &it = s->interp[a];
    auto     &p  = it.sts[it.u8_now_use_];
    q[RK_VI_VEL_NOW] = f2u(it.vel_now_);
    q[RK_VI_ACL_NOW] = f2u(it.acl_now_);
    q[RK_VI_VEL_TGT] = f2u(p.vel_tgt_);
    q[RK_VI_ACL_MAX] = f2u(p.acl_max_);
    q[RK_VI_JERK_P]  = f2u(p.jerk_p_);
    q[RK_VI_JERK_M]  = f2u(p.jerk_m_);
    q[RK_VI_DT1]     = f2u(p.dt1_);
    q[RK_VI_DT2]     = f2u(p.dt2_);
    q[RK_VI_DT3]     = f2u(p.dt3_);
    q[RK_VI_VEL_INI] = f2u(p.vel_ini_);
    q[RK_VI_ACL_INI] = f2u(p.acl_ini_);
    q[RK_VI_DT]      = f2u(p.dt_);
  }
  for(int k = 0; k < 4; k++) {
    uint32_t *q = w + RK_VS_CTRL0 + 8 * k;
    auto     &c = s->ctrl[k];
    q[RK_VC_PREV_VAL] = f2u(c.prev_val_);
    q[RK_VC_INTEG]    = f2u(c.Integ_);
    q[RK_VC_LPF_Y]    = f2u(c.velLpf_.prev_Y_);
    q[RK_VC_LPF_X]    = f2u(c.velLpf_.prev_X_);
    q[RK_VC_NOW_TGT]  = f2u(c.now_tgt_);
    q[RK_VC_NOW_ERR]  = f2u(c.now_error_);
    q[RK_VC_NOW_CTRL] = f2u(c.now_ctrl_);
  }
  for(int k = 0; k < 4; k++) {
    uint32_t *q  = w + RK_VS_MOTOR0 + 8 * k;
    auto     &m  = s->motor[k];
    auto     &st = m.status_buf[m.status_head];
    uint64_t  sum = (uint64_t)m.s64_rawAngleSum, prev = (uint64_t)s->vhcl.s64_rawAngleSumPrev[k];
    q[RK_VM_SUM_LO]  = (uint32_t)sum;
    q[RK_VM_SUM_HI]  = (uint32_t)(sum >> 32);
    q[RK_VM_PREV_LO] = (uint32_t)prev;
    q[RK_VM_PREV_HI] = (uint32_t)(prev >> 32);
    q[RK_VM_ANG_RPM] = pack16(st.s16_rawAngle, st.s16_rawSpeedRpm);
    q[RK_VM_CUR_TGT] = pack16(st.s16_rawCurr, m.s16_rawCurr_tgt);
    q[RK_VM_USEC]    = pack16(st.s16_microsec_id, m.status_head);
    q[RK_VM_PLANT]   = pack16(s->plant_ang[k], s->plant_rpm[k]);
  }
}

void import_state(VehicleSet *s, const uint32_t *w) {
  s->vhcl.now_vhcl_pos_m_       = Direction{u2f(w[RK_VS_POS_X]), u2f(w[RK_VS_POS_Y]), u2f(w[RK_VS_POS_TH])};
  s->vhcl.isPowerOn             = (w[RK_VS_FLAGS] & RK_VS_FLAG_POWER_ON) != 0;
  s->vhcl.now_vhcl_vel_mmps     = Direction{u2f(w[RK_VS_VEL_X]), u2f(w[RK_VS_VEL_Y]), u2f(w[RK_VS_VEL_TH])};
  s->vhcl.now_vhcl_vel_tgt_mmps = Direction{u2f(w[RK_VS_TGT_X]), u2f(w[RK_VS_TGT_Y]), u2f(w[RK_VS_TGT_TH])};
  for(int a = 0; a < 3; a++) {
    const uint32_t *q  = w + RK_VS_INTERP0 + 12 * a;
    auto           &it = s->interp[a];
    memset(it.sts, 0, sizeof(it.sts));
    it.u8_now_use_ = 0;
    auto &p        = it.sts[0];
    it.vel_now_    = u2f(q[RK_VI_VEL_NOW]);
    it.acl_now_    = u2f(q[RK_VI_ACL_NOW]);
    p.vel_tgt_     = u2f(q[RK_VI_VEL_TGT]);
    p.acl_max_     = u2f(q[RK_VI_ACL_MAX]);
    p.jerk_p_      = u2f(q[RK_VI_JERK_P]);
    p.jerk_m_      = u2f(q[RK_VI_JERK_M]);
    p.dt1_         = u2f(q[RK_VI_DT1]);
    p.dt2_         = u2f(q[RK_VI_DT2]);
    p.dt3_         = u2f(q[RK_VI_DT3]);
    p.vel_ini_     = u2f(q[RK_VI_VEL_INI]);
    p.acl_ini_     = u2f(q[RK_VI_ACL_INI]);
    p.dt_          = u2f(q[RK_VI_DT]);
  }
  for(int k = 0; k < 4; k++) {
    const uint32_t *q = w + RK_VS_CTRL0 + 8 * k;
    auto           &c = s->ctrl[k];
    c.prev_val_ = c.now_val_ = u2f(q[RK_VC_PREV_VAL]);
    c.Integ_                 = u2f(q[RK_VC_INTEG]);
    c.velLpf_.prev_Y_ = c.velLpf_.now_Y_ = u2f(q[RK_VC_LPF_Y]);
    c.velLpf_.prev_X_                    = u2f(q[RK_VC_LPF_X]);
    c.now_tgt_                           = u2f(q[RK_VC_NOW_TGT]);
    c.now_error_ = c.prev_error_ = u2f(q[RK_VC_NOW_ERR]);
    c.now_ctrl_                  = u2f(q[RK_VC_NOW_CTRL]);
  }
  for(int k = 0; k < 4; k++) {
    const uint32_t *q = w + RK_VS_MOTOR0 + 8 * k;
    auto           &m = s->motor[k];
    m.s64_rawAngleSum            = (int64_t)(((uint64_t)q[RK_VM_SUM_HI] << 32) | q[RK_VM_SUM_LO]);
    s->vhcl.s64_rawAngleSumPrev[k] = (int64_t)(((uint64_t)q[RK_VM_PREV_HI] << 32) | q[RK_VM_PREV_LO]);
    memset(m.status_buf, 0, sizeof(m.status_buf));
    m.status_head        = (uint8_t)(hi16(q[RK_VM_USEC]) % 3);
    auto &st             = m.status_buf[m.status_head];
    st.s16_rawAngle      = (int16_t)lo16(q[RK_VM_ANG_RPM]);
    st.s16_rawSpeedRpm   = (int16_t)hi16(q[RK_VM_ANG_RPM]);
    st.s16_rawCurr       = (int16_t)lo16(q[RK_VM_CUR_TGT]);
    m.s16_rawCurr_tgt    = (int16_t)hi16(q[RK_VM_CUR_TGT]);
    st.s16_microsec_id   = (int16_t)lo16(q[RK_VM_USEC]);
    s->plant_ang[k]      = lo16(q[RK_VM_PLANT]);
    s->plant_rpm[k]      = hi16(q[RK_VM_PLANT]);
  }
}

/* The synthetic plant of robotick.h (RK_SENSOR_PLANT): first-order integer motor model in
 * the motor's own frame, emitting the 8-byte C610 feedback frame of
 * MOTOR_IF_M2006::CanMsgRx (VD_motor_if_m2006.hpp:13-21). */
inline uint64_t plant_step(VehicleSet *s, int k) {
  int32_t cur = s->motor[k].get_rawCurr_tgt();
  int32_t rpm = s->plant_rpm[k];
  int32_t ang = s->plant_ang[k];
  rpm += ((cur * 4 - rpm) >> 4);
  ang = (ang + rpm * 8192 / 60000) & 8191;
  s->plant_rpm[k] = rpm;
  s->plant_ang[k] = ang;
  uint8_t f[8]    = {(uint8_t)(ang >> 8), (uint8_t)ang, (uint8_t)(rpm >> 8), (uint8_t)rpm,
                     (uint8_t)(cur >> 8), (uint8_t)cur, 0, 0};
  uint64_t v;
  memcpy(&v, f, 8);
  return v;
}

inline void apply_cmd(VehicleSet *s, const rk_vdt_cmd_t &c) {
  /* constants: VD_task_main.cpp:29-48 */
  static Direction A_MOVE = {1000.0f, 1000.0f, 30.0f}, J_MOVE = {10000.0f, 10000.0f, 300.0f};
  static Direction A_STOP = {2000.0f, 2000.0f, 70.0f}, J_STOP = {30000.0f, 30000.0f, 1000.0f};
  if(c.kind == RK_CMD_NONE) return;
  Direction v = {c.vx, c.vy, c.vth};
  s->vhcl.start();
  if(c.kind == RK_CMD_STOP)
    s->vhcl.set_target_vel(v, A_STOP, J_STOP);
  else
    s->vhcl.set_target_vel(v, A_MOVE, J_MOVE);
}

void rollout_one(VehicleSet *s, int64_t n, int64_t i, const rk_vdt_rollout_t *a) {
  for(int t = 0; t < a->steps; t++) {
    if(a->d_cmd && a->seg_len > 0 && (t % a->seg_len) == 0 && (t / a->seg_len) < a->n_seg)
      apply_cmd(s, a->d_cmd[(int64_t)(t / a->seg_len) * n + i]);
    if((a->d_yaw || a->d_yaw_reg) && a->yaw_period > 0 && (t % a->yaw_period) == 0 && (t / a->yaw_period) < a->n_yaw) {
      const int64_t yi = (int64_t)(t / a->yaw_period) * n + i;
      /* from the Yaw register: what IMU_IF_WT901C::updateData (imu_if_wt901c.cpp:100) stores, then the ISR's deg2rad */
      s->vhcl.set_now_yaw_world(a->d_yaw ? a->d_yaw[yi]
                                         : UTIL::mymath::deg2rad(static_cast<float>(a->d_yaw_reg[yi]) / 32768.0f * 180.0f));
    }
    int16_t us = (int16_t)(((t + 1) * 1000) & 0x7FFF);
    if(a->sensor_mode == RK_SENSOR_PLANT) {
      for(int k = 0; k < 4; k++) {
        uint64_t f = plant_step(s, k);
        s->motor[k].rx_callback((MOTOR_IF_M2006::CanMsgRx *)&f, us);
      }
    } else if(a->sensor_mode == RK_SENSOR_STREAM) {
      for(int k = 0; k < 4; k++) {
        uint64_t f = a->d_frames[((int64_t)t * 4 + k) * n + i];
        s->motor[k].rx_callback((MOTOR_IF_M2006::CanMsgRx *)&f, us);
      }
    }
    s->vhcl.update();
    if(a->d_trace) {
      uint32_t *tr = a->d_trace + (int64_t)t * RK_VDT_TRACE_WORDS * n + i;
      Direction p, v, g;
      s->vhcl.get_vehicle_pos_m_latest(p);
      s->vhcl.get_vehicle_vel_mmps_latest(v);
      s->vhcl.get_vehicle_vel_tgt_mmps_latest(g);
      float f[9] = {p.x, p.y, p.th, v.x, v.y, v.th, g.x, g.y, g.th};
      for(int j = 0; j < 9; j++) tr[(int64_t)j * n] = f2u(f[j]);
      for(int k = 0; k < 4; k++) tr[(int64_t)(9 + k) * n] = (uint32_t)(int32_t)s->motor[k].get_rawCurr_tgt();
      tr[(int64_t)13 * n] = 0;
      { /* the C610 frame: VD_can_controller.hpp:43-55 needs FlexCAN and a static motor table, so the reference's own
         * tx_routine() runs in the whole-task harness (libref_vdt_task.so, where tests/test_vdt_task_cpu.py pins these
         * words); here the same bytes are laid out from get_rawCurr_tgt() */
        uint8_t b[8];
        for(int k = 0; k < 4; k++) {
          b[2 * k]     = (uint8_t)(s->motor[k].get_rawCurr_tgt() >> 8);
          b[2 * k + 1] = (uint8_t)(s->motor[k].get_rawCurr_tgt() & 0x00FF);
        }
        for(int j = 0; j < 2; j++)
          tr[(int64_t)(14 + j) * n] = (uint32_t)b[4 * j] | ((uint32_t)b[4 * j + 1] << 8) | ((uint32_t)b[4 * j + 2] << 16) | ((uint32_t)b[4 * j + 3] << 24);
      }
    }
  }
  if(a->d_cost && a->d_goal) {
    float dx = s->vhcl.now_vhcl_pos_m_.x - a->d_goal[2 * i], dy = s->vhcl.now_vhcl_pos_m_.y - a->d_goal[2 * i + 1];
    a->d_cost[i] = dx * dx + dy * dy;
  }
}

VehicleSet *make_set() {
  void *mem = calloc(1, sizeof(VehicleSet));
  if(!mem) return nullptr;
  VehicleSet *s = new(mem) VehicleSet();
  zero_uninitialised(s);
  return s;
}

inline uint32_t &soa(uint32_t *blk, int64_t n, int64_t i, int w) { return blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }

} // namespace

extern "C" {

void *ref_vdt_create(void) { return make_set(); }
void  ref_vdt_destroy(void *h) {
  if(!h) return;
  ((VehicleSet *)h)->~VehicleSet();
  free(h);
}
void ref_vdt_start(void *h) { ((VehicleSet *)h)->vhcl.start(); }
void ref_vdt_stop(void *h) { ((VehicleSet *)h)->vhcl.stop(); }
void ref_vdt_set_target(void *h, const float v[3], const float a[3], const float j[3]) {
  Direction dv = {v[0], v[1], v[2]}, da = {a[0], a[1], a[2]}, dj = {j[0], j[1], j[2]};
  ((VehicleSet *)h)->vhcl.set_target_vel(dv, da, dj);
}
void ref_vdt_set_yaw(void *h, float yaw_rad) { ((VehicleSet *)h)->vhcl.set_now_yaw_world(yaw_rad); }
void ref_vdt_rx(void *h, int wheel, const uint8_t frame[8], int16_t usec_id) {
  MOTOR_IF_M2006::CanMsgRx m;
  memcpy(&m, frame, 8);
  ((VehicleSet *)h)->motor[wheel].rx_callback(&m, usec_id);
}
void ref_vdt_update(void *h) { ((VehicleSet *)h)->vhcl.update(); }
void ref_vdt_export(void *h, uint32_t *words) { export_state((VehicleSet *)h, words); }
void ref_vdt_import(void *h, const uint32_t *words) { import_state((VehicleSet *)h, words); }
void ref_vdt_get(void *h, float pos[3], float vel[3], float tgt[3], int16_t cur[4]) {
  VehicleSet *s = (VehicleSet *)h;
  Direction   p, v, g;
  s->vhcl.get_vehicle_pos_m_latest(p);
  s->vhcl.get_vehicle_vel_mmps_latest(v);
  s->vhcl.get_vehicle_vel_tgt_mmps_latest(g);
  pos[0] = p.x, pos[1] = p.y, pos[2] = p.th;
  vel[0] = v.x, vel[1] = v.y, vel[2] = v.th;
  tgt[0] = g.x, tgt[1] = g.y, tgt[2] = g.th;
  for(int k = 0; k < 4; k++) cur[k] = s->motor[k].get_rawCurr_tgt();
}

/* Same contract as rk_vdt_rollout() but on HOST arrays (same SoA indexing, pitch n), for
 * instances [i0, i1), on nthreads host threads.  state may be NULL (power-on state, result
 * discarded -- throughput runs). */
void ref_vdt_rollout(uint32_t *state, int64_t n, int64_t i0, int64_t i1, const rk_vdt_rollout_t *args, int nthreads) {
  if(nthreads < 1) nthreads = 1;
  auto work = [&](int tid) {
    VehicleSet *s = make_set();
    uint32_t    w[RK_VS_WORDS];
    for(int64_t i = i0 + tid; i < i1; i += nthreads) {
      if(state) {
        for(int k = 0; k < RK_VS_WORDS; k++) w[k] = soa(state, n, i, k);
      } else {
        memset(w, 0, sizeof(w));
      }
      import_state(s, w);
      rollout_one(s, n, i, args);
      if(state) {
        export_state(s, w);
        for(int k = 0; k < RK_VS_WORDS; k++) soa(state, n, i, k) = w[k];
      }
    }
    s->~VehicleSet();
    free(s);
  };
  if(nthreads == 1) {
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for(int t = 0; t < nthreads; t++) th.emplace_back(work, t);
  for(auto &t : th) t.join();
}

/* mymath probes (SURVEY.md Appendix D known answers) */
float ref_normalize_rad_0to2pi(float x) { return UTIL::mymath::normalize_rad_0to2pi(x); }
float ref_normalize_deg_0to360(float x) { return UTIL::mymath::normalize_deg_0to360(x); }
float ref_atan2f(float y, float x) { return UTIL::mymath::atan2f(y, x); }
float ref_atanf(float x) { return UTIL::mymath::atanf(x); }
float ref_sinf(float x) { return UTIL::mymath::sinf(x); }
float ref_cosf(float x) { return UTIL::mymath::cosf(x); }
float ref_sqrtf(float x) { return UTIL::mymath::sqrtf(x); }
}
