// rk_vehicle.cuh -- device-side vehicle tick: VEHICLE_CTRL::update() and the src/Utility /
// MOTOR_IF_M2006 pieces it calls, one thread per robot, state in registers.
//
// Reference (paths relative to the reference tree):
//   src/VehicleDrive/VD_vehicle_controller.cpp:6-99,101-105,113-118,126-130
//   src/VehicleDrive/VD_motor_if_m2006.hpp:36-37,44-47,57 ; VD_motor_if_m2006.cpp:32-72
//   src/Utility/util_vel_interp.hpp:53-143 ; util_controller.hpp:90-124,156-165 ; util_iir.hpp:39-45
// Evaluation order is the C++ left-to-right order of those expressions, one rounding per
// operation.
#pragma once
#include "rk_common.cuh"
#include <limits.h>

#include "rk_math.cuh"

namespace rk {

// VD_motor_if_m2006.hpp:76-82 (constexpr float arithmetic, folded identically by nvcc's
// front end: single IEEE operations on literals)
#define RK_RPM_TO_RADPS (2.0f * 3.1415926f / 60.0f)
#define RK_AMPERE_TO_RAW_CURR 1000.0f
#define RK_GEAR_RATIO 36.0f
#define RK_GEAR_RATIO_INV (1.0f / 36.0f)
#define RK_OUT_RAD_PER_RAW_ANGLE (2.0f * 3.1415926f / 8191.0f)

struct Interp { // VelInterpConstJerk: vel_now_/acl_now_ + active StatusBuf page
  float vel, acl, vel_tgt, acl_max, jerk_p, jerk_m, dt1, dt2, dt3, vel_ini, acl_ini, dt;
};
struct Ctrl { // FF_PI_D
  float prev_val, integ, lpf_y, lpf_x, now_tgt, now_err, now_ctrl;
};
struct Motor { // MOTOR_IF_M2006 head Status + sums + synthetic plant
  int64_t sum, prev;
  int32_t ang, rpm, cur, cur_tgt, usec, head, p_ang, p_rpm;
};
struct Veh {
  float    pos[3], vel[3], tgt[3];
  uint32_t flags;
  uint32_t move_cnt, rsv1; // VDT::U32_MOVE_TIME_CNT_ORDER (VD_task_main.cpp:115); reserved word, carried through
  Interp   it[3];
  Ctrl     c[4];
  Motor    m[4];
};

// Values derived from rk_vdt_params_t once per thread, with the same float operations the
// reference constructors perform (util_controller.hpp:10,90-92).
struct Derived {
  float A1, B0, B1; // IIR1 coefficients of PI_D::velLpf_
  float ki_dt;      // Igain_ * dt_
  float s2l;        // SQRTF2 * WHEEL_L_MM
};
RK_DEV Derived derive(const rk_vdt_params_t &p) {
  Derived d;
  float   two_f = fmul(2.0f, p.ctrl_freq);
  float   den   = fadd(two_f, p.lpf_freq);
  d.A1          = fdiv(fsub(two_f, p.lpf_freq), den);
  d.B0          = fdiv(p.lpf_freq, den);
  d.B1          = d.B0;
  d.ki_dt       = fmul(p.ki, fdiv(1.0f, p.ctrl_freq));
  d.s2l         = fmul(p.sqrtf2, p.wheel_l_mm);
  return d;
}

// ---- state block <-> registers --------------------------------------------------------
RK_DEV void load_interp(const uint4 *blk, int64_t n, int64_t i, int a, Interp &t) {
  uint4 q0 = ld_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 0, i);
  uint4 q1 = ld_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 1, i);
  uint4 q2 = ld_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 2, i);
  t.vel = u2f(q0.x), t.acl = u2f(q0.y), t.vel_tgt = u2f(q0.z), t.acl_max = u2f(q0.w);
  t.jerk_p = u2f(q1.x), t.jerk_m = u2f(q1.y), t.dt1 = u2f(q1.z), t.dt2 = u2f(q1.w);
  t.dt3 = u2f(q2.x), t.vel_ini = u2f(q2.y), t.acl_ini = u2f(q2.z), t.dt = u2f(q2.w);
}
RK_DEV void store_interp(uint4 *blk, int64_t n, int64_t i, int a, const Interp &t) {
  st_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 0, i, make_uint4(f2u(t.vel), f2u(t.acl), f2u(t.vel_tgt), f2u(t.acl_max)));
  st_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 1, i, make_uint4(f2u(t.jerk_p), f2u(t.jerk_m), f2u(t.dt1), f2u(t.dt2)));
  st_plane(blk, n, RK_VS_INTERP0 / 4 + 3 * a + 2, i, make_uint4(f2u(t.dt3), f2u(t.vel_ini), f2u(t.acl_ini), f2u(t.dt)));
}
RK_DEV void load_ctrl(const uint4 *blk, int64_t n, int64_t i, int w, Ctrl &c) {
  uint4 q0 = ld_plane(blk, n, RK_VS_CTRL0 / 4 + 2 * w + 0, i);
  uint4 q1 = ld_plane(blk, n, RK_VS_CTRL0 / 4 + 2 * w + 1, i);
  c.prev_val = u2f(q0.x), c.integ = u2f(q0.y), c.lpf_y = u2f(q0.z), c.lpf_x = u2f(q0.w);
  c.now_tgt = u2f(q1.x), c.now_err = u2f(q1.y), c.now_ctrl = u2f(q1.z);
}
RK_DEV void store_ctrl(uint4 *blk, int64_t n, int64_t i, int w, const Ctrl &c) {
  st_plane(blk, n, RK_VS_CTRL0 / 4 + 2 * w + 0, i, make_uint4(f2u(c.prev_val), f2u(c.integ), f2u(c.lpf_y), f2u(c.lpf_x)));
  st_plane(blk, n, RK_VS_CTRL0 / 4 + 2 * w + 1, i, make_uint4(f2u(c.now_tgt), f2u(c.now_err), f2u(c.now_ctrl), 0u));
}
RK_DEV void load_motor(const uint4 *blk, int64_t n, int64_t i, int w, Motor &m) {
  uint4 q0 = ld_plane(blk, n, RK_VS_MOTOR0 / 4 + 2 * w + 0, i);
  uint4 q1 = ld_plane(blk, n, RK_VS_MOTOR0 / 4 + 2 * w + 1, i);
  m.sum     = (int64_t)(((uint64_t)q0.y << 32) | q0.x);
  m.prev    = (int64_t)(((uint64_t)q0.w << 32) | q0.z);
  m.ang     = lo16(q1.x);
  m.rpm     = hi16(q1.x);
  m.cur     = lo16(q1.y);
  m.cur_tgt = hi16(q1.y);
  m.usec    = lo16(q1.z);
  m.head    = hi16(q1.z) % 3;
  m.p_ang   = lo16(q1.w);
  m.p_rpm   = hi16(q1.w);
}
RK_DEV void store_motor(uint4 *blk, int64_t n, int64_t i, int w, const Motor &m) {
  uint64_t s = (uint64_t)m.sum, p = (uint64_t)m.prev;
  st_plane(blk, n, RK_VS_MOTOR0 / 4 + 2 * w + 0, i,
           make_uint4((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)p, (uint32_t)(p >> 32)));
  st_plane(blk, n, RK_VS_MOTOR0 / 4 + 2 * w + 1, i,
           make_uint4(pack16(m.ang, m.rpm), pack16(m.cur, m.cur_tgt), pack16(m.usec, m.head), pack16(m.p_ang, m.p_rpm)));
}
RK_DEV void load_veh(const uint4 *blk, int64_t n, int64_t i, Veh &v) {
  uint4 q0 = ld_plane(blk, n, 0, i), q1 = ld_plane(blk, n, 1, i), q2 = ld_plane(blk, n, 2, i);
  v.pos[0] = u2f(q0.x), v.pos[1] = u2f(q0.y), v.pos[2] = u2f(q0.z), v.flags = q0.w;
  v.vel[0] = u2f(q1.x), v.vel[1] = u2f(q1.y), v.vel[2] = u2f(q1.z), v.tgt[0] = u2f(q1.w);
  v.tgt[1] = u2f(q2.x), v.tgt[2] = u2f(q2.y), v.move_cnt = q2.z, v.rsv1 = q2.w;
#pragma unroll
  for(int a = 0; a < 3; a++) load_interp(blk, n, i, a, v.it[a]);
#pragma unroll
  for(int w = 0; w < 4; w++) load_ctrl(blk, n, i, w, v.c[w]);
#pragma unroll
  for(int w = 0; w < 4; w++) load_motor(blk, n, i, w, v.m[w]);
}
RK_DEV void store_veh(uint4 *blk, int64_t n, int64_t i, const Veh &v) {
  st_plane(blk, n, 0, i, make_uint4(f2u(v.pos[0]), f2u(v.pos[1]), f2u(v.pos[2]), v.flags));
  st_plane(blk, n, 1, i, make_uint4(f2u(v.vel[0]), f2u(v.vel[1]), f2u(v.vel[2]), f2u(v.tgt[0])));
  st_plane(blk, n, 2, i, make_uint4(f2u(v.tgt[1]), f2u(v.tgt[2]), v.move_cnt, v.rsv1));
#pragma unroll
  for(int a = 0; a < 3; a++) store_interp(blk, n, i, a, v.it[a]);
#pragma unroll
  for(int w = 0; w < 4; w++) store_ctrl(blk, n, i, w, v.c[w]);
#pragma unroll
  for(int w = 0; w < 4; w++) store_motor(blk, n, i, w, v.m[w]);
}

// ---- VelInterpConstJerk ---------------------------------------------------------------
// set_target_params  util_vel_interp.hpp:53-108.  Writes the inactive page and flips to it;
// only the page that becomes active is state, so it is written in place.
RK_DEV void interp_set(Interp &t, float v_t, float a_m, float jrk) {
  float vel_tgt = v_t, acl_max = a_m, vel_ini = t.vel, acl_ini = t.acl;
  if(fsub(vel_tgt, vel_ini) < 0.0f) acl_max = -a_m;
  float jerk_m = (acl_max >= 0.0f) ? -jrk : jrk;
  float jm_inv = fdiv(1.0f, jerk_m);
  float jerk_p = (fsub(acl_max, acl_ini) >= 0.0f) ? jrk : -jrk;
  float jp_inv = fdiv(1.0f, jerk_p);
  float dt1    = fmul(fsub(acl_max, acl_ini), jp_inv);
  float dt3    = fmul(acl_max, -jm_inv);
  //   1.0f / acl_max * (vel_tgt - vel_ini - acl_ini*dt1*0.5f - acl_max*(dt1+dt3)*0.5f)
  float inner = fsub(fsub(fsub(vel_tgt, vel_ini), fmul(fmul(acl_ini, dt1), 0.5f)),
                     fmul(fmul(acl_max, fadd(dt1, dt3)), 0.5f));
  float dt2   = fmul(fdiv(1.0f, acl_max), inner);
  if(dt2 < 0.0f) {
    float q     = fmul(acl_ini, jp_inv);
    float sq_in = fadd(fmul(fmul(q, q), 0.5f), fmul(fsub(vel_tgt, vel_ini), jp_inv));
    float sq    = arm_sqrt(sq_in);
    dt1         = fsub(sq, fmul(acl_ini, jp_inv));
    acl_max     = fadd(acl_ini, fmul(jerk_p, dt1));
    dt2         = 0.0f;
    dt3         = fmul(acl_max, -jm_inv);
  }
  dt1 = (dt1 < 0.0f) ? 0.0f : dt1;
  dt3 = (dt3 < 0.0f) ? 0.0f : dt3;
  t.vel_tgt = vel_tgt, t.acl_max = acl_max, t.jerk_p = jerk_p, t.jerk_m = jerk_m;
  t.dt1 = dt1, t.dt2 = dt2, t.dt3 = dt3, t.vel_ini = vel_ini, t.acl_ini = acl_ini, t.dt = 0.0f;
}
// set_target_params with the three reciprocals it forms -- 1 / jerk_m, 1 / jerk_p, 1 / acl_max, each +-1 / a launch
// constant -- taken from the host's IEEE divisions: 1 / (-x) == -(1 / x) exactly, zeros and infinities included.
RK_DEV void interp_set_rcp(Interp &t, float v_t, float a_m, float jrk, float ra, float rj) {
  float vel_tgt = v_t, acl_max = a_m, am_inv = ra, vel_ini = t.vel, acl_ini = t.acl;
  if(fsub(vel_tgt, vel_ini) < 0.0f) acl_max = -a_m, am_inv = -ra;
  const bool  mneg   = acl_max >= 0.0f;
  float       jerk_m = mneg ? -jrk : jrk;
  const float jm_inv = mneg ? -rj : rj;
  const bool  ppos   = fsub(acl_max, acl_ini) >= 0.0f;
  float       jerk_p = ppos ? jrk : -jrk;
  const float jp_inv = ppos ? rj : -rj;
  float dt1    = fmul(fsub(acl_max, acl_ini), jp_inv);
  float dt3    = fmul(acl_max, -jm_inv);
  float inner = fsub(fsub(fsub(vel_tgt, vel_ini), fmul(fmul(acl_ini, dt1), 0.5f)),
                     fmul(fmul(acl_max, fadd(dt1, dt3)), 0.5f));
  float dt2   = fmul(am_inv, inner);
  if(dt2 < 0.0f) {
    float q     = fmul(acl_ini, jp_inv);
    float sq_in = fadd(fmul(fmul(q, q), 0.5f), fmul(fsub(vel_tgt, vel_ini), jp_inv));
    float sq    = arm_sqrt(sq_in);
    dt1         = fsub(sq, fmul(acl_ini, jp_inv));
    acl_max     = fadd(acl_ini, fmul(jerk_p, dt1));
    dt2         = 0.0f;
    dt3         = fmul(acl_max, -jm_inv);
  }
  dt1 = (dt1 < 0.0f) ? 0.0f : dt1;
  dt3 = (dt3 < 0.0f) ? 0.0f : dt3;
  t.vel_tgt = vel_tgt, t.acl_max = acl_max, t.jerk_p = jerk_p, t.jerk_m = jerk_m;
  t.dt1 = dt1, t.dt2 = dt2, t.dt3 = dt3, t.vel_ini = vel_ini, t.acl_ini = acl_ini, t.dt = 0.0f;
}
// update  util_vel_interp.hpp:110-136
RK_DEV float interp_update(Interp &t, float ts) {
  float t1 = fadd(t.dt1, ts);
  float t2 = fadd(fadd(t.dt1, t.dt2), ts);
  float t3 = fadd(fadd(fadd(t.dt1, t.dt2), t.dt3), ts);
  if(t.dt <= t1) {
    t.acl = fadd(t.acl_ini, fmul(t.jerk_p, t.dt));
    t.vel = fadd(t.vel_ini, fmul(fmul(fadd(t.acl_ini, t.acl), t.dt), 0.5f));
    t.dt  = fadd(t.dt, ts);
  } else if(t.dt <= t2) {
    t.acl = t.acl_max;
    t.vel = fadd(t.vel, fmul(t.acl, ts));
    t.dt  = fadd(t.dt, ts);
  } else if(t.dt <= t3) {
    t.acl = fadd(t.acl_max, fmul(t.jerk_m, fsub(fsub(t.dt, t.dt1), t.dt2)));
    t.vel = fadd(t.vel, fmul(t.acl, ts));
    t.dt  = fadd(t.dt, ts);
  } else {
    t.acl = 0.0f;
    t.vel = t.vel_tgt;
  }
  return t.vel;
}
// reset  util_vel_interp.hpp:138-143
RK_DEV void interp_reset(Interp &t) {
  t.vel = t.acl = t.vel_tgt = t.acl_max = t.jerk_p = t.jerk_m = 0.0f;
  t.dt1 = t.dt2 = t.dt3 = t.vel_ini = t.acl_ini = t.dt = 0.0f;
}

// ---- FF_PI_D::update -> PI_D::update -> IIR1::update ----------------------------------
// util_controller.hpp:159-165, :94-110 ; util_iir.hpp:39-45
RK_DEV float ctrl_update(Ctrl &c, const rk_vdt_params_t &p, const Derived &d, float now_val) {
  float err = fsub(c.now_tgt, now_val);
  float x   = fmul(fsub(now_val, c.prev_val), p.ctrl_freq);
  float y   = fadd(fadd(fmul(d.A1, c.lpf_y), fmul(d.B0, x)), fmul(d.B1, c.lpf_x));
  c.lpf_y   = y;
  c.lpf_x   = x;
  float I   = fadd(c.integ, fmul(d.ki_dt, err));
  I         = (I >= p.i_limit) ? p.i_limit : ((I <= -p.i_limit) ? -p.i_limit : I);
  c.integ   = I;
  float u   = fsub(fadd(fmul(p.kp, err), I), fmul(p.kd, y));
  c.prev_val = now_val;
  c.now_err  = err;
  float ff   = fmul(c.now_tgt, p.kff);
  ff         = (ff >= p.ff_limit) ? p.ff_limit : ((ff <= -p.ff_limit) ? -p.ff_limit : ff);
  u          = fadd(u, ff);
  c.now_ctrl = u;
  return u;
}
// PI_D::reset  util_controller.hpp:112-124
RK_DEV void ctrl_reset(Ctrl &c) { c.prev_val = c.integ = c.lpf_y = c.lpf_x = c.now_tgt = c.now_err = c.now_ctrl = 0.0f; }

// ---- MOTOR_IF_M2006 ---------------------------------------------------------------------
// set_CurrA_tgt -> set_rawCurr_tgt -> sat_curr  VD_motor_if_m2006.hpp:36-37,57
RK_DEV void motor_set_curr(Motor &m, int dir, int lim, float amp) {
  int32_t t = f2s16(fmul(amp, RK_AMPERE_TO_RAW_CURR));
  int32_t c = sext16(t * dir);
  int32_t l = sext16(lim);
  m.cur_tgt = (c > l) ? l : ((c < -l) ? sext16(-l) : c);
}
// rx_callback  VD_motor_if_m2006.cpp:32-72, integer part.  (flt_SpeedRadPS and
// flt_dltOutAngle_rad are never consumed: VD_vehicle_controller.cpp:20-24 takes `#if 1`.)
// frame bytes: [0..1] angle BE, [2..3] speed BE, [4..5] current BE (little-endian uint64 load).
RK_DEV void motor_rx(Motor &m, int dir, uint64_t frame, int32_t usec) {
  uint32_t lo = (uint32_t)frame, hi = (uint32_t)(frame >> 32);
  int32_t  a  = sext16((int32_t)(__byte_perm(lo, 0, 0x4401)));   // (b0<<8)|b1
  int32_t  r  = sext16((int32_t)(__byte_perm(lo, 0, 0x4423)));   // (b2<<8)|b3
  int32_t  c  = sext16((int32_t)(__byte_perm(hi, 0, 0x4401)));   // (b4<<8)|b5
  int32_t  raw_ang = (dir == 1) ? a : sext16(8192 - a);
  int32_t  d       = sext16(raw_ang - m.ang);
  d                = (d > 4096) ? sext16(d - 8192) : ((d < -4096) ? sext16(d + 8192) : d);
  m.sum            = m.sum + (int64_t)d;
  m.ang            = raw_ang;
  m.rpm            = sext16(r * dir);
  m.cur            = sext16(c * dir);
  m.usec           = sext16(usec);
  m.head           = (m.head + 1 >= 3) ? 0 : m.head + 1;
}

// The synthetic plant of RK_SENSOR_PLANT (robotick.h): first-order integer motor model in
// the motor's own frame; emits the C610 feedback frame.  Not part of the reference.
RK_DEV uint64_t plant_frame(Motor &m) {
  int32_t cur = m.cur_tgt, rpm = m.p_rpm, ang = m.p_ang;
  rpm += ((cur * 4 - rpm) >> 4);
  ang     = (ang + rpm * 8192 / 60000) & 8191;
  m.p_rpm = rpm, m.p_ang = ang;
  uint32_t lo = __byte_perm((uint32_t)ang, (uint32_t)rpm, 0x4501); // b0=ang>>8 b1=ang b2=rpm>>8 b3=rpm
  uint32_t hi = __byte_perm((uint32_t)cur, 0, 0x4401);             // b4=cur>>8 b5=cur
  return ((uint64_t)hi << 32) | lo;
}

// ---- kinematics  VD_vehicle_controller.cpp:113-118,126-130 -----------------------------
RK_DEV void fk_xy(const rk_vdt_params_t &p, const float M[4], float &x, float &y) {
  x = fmul(fmul(fadd(fadd(fadd(M[0], M[1]), M[2]), M[3]), 0.25f), p.wheel_radius_mm);
  y = fmul(fmul(fadd(fsub(fadd(-M[0], M[1]), M[2]), M[3]), 0.25f), p.wheel_radius_mm);
}
RK_DEV float fk_th(const rk_vdt_params_t &p, const float M[4]) {
  float s = fadd(fadd(fsub(-M[0], M[1]), M[2]), M[3]);
  return fmul(fdiv(fdiv(fmul(s, 0.25f), p.sqrtf2), p.wheel_l_mm), p.wheel_radius_mm);
}
RK_DEV void ik(const rk_vdt_params_t &p, const Derived &d, const float V[3], float M[4]) {
  float T   = fmul(fmul(d.s2l, V[2]), 4.0f);
  float xmy = fsub(V[0], V[1]), xpy = fadd(V[0], V[1]);
  M[0]      = fdiv(fsub(xmy, T), p.wheel_radius_mm);
  M[1]      = fdiv(fsub(xpy, T), p.wheel_radius_mm);
  M[2]      = fdiv(fadd(xmy, T), p.wheel_radius_mm);
  M[3]      = fdiv(fadd(xpy, T), p.wheel_radius_mm);
}

// cos/sin of the world yaw, hoisted: pos.th only changes through set_now_yaw_world(), so
// VD_vehicle_controller.cpp:47-49 is re-evaluated only when the yaw word changes.
RK_DEV void yaw_trig(const float *s_tab, float th, float &c, float &s) {
  float rad = normalize_rad_0to2pi(th);
  c         = arm_cos(s_tab, rad);
  s         = arm_sin(s_tab, rad);
}

// ---- VEHICLE_CTRL::update()  VD_vehicle_controller.cpp:6-99 ------------------------------
RK_DEV void veh_update(Veh &v, const rk_vdt_params_t &p, const Derived &d, float cth, float sth) {
  float Mvel[4], Mrad[4], Mtgt[4];
#pragma unroll
  for(int k = 0; k < 4; k++) Mvel[k] = fmul(fmul((float)v.m[k].rpm, RK_RPM_TO_RADPS), RK_GEAR_RATIO_INV);
  fk_xy(p, Mvel, v.vel[0], v.vel[1]);
  v.vel[2] = fk_th(p, Mvel);
#pragma unroll
  for(int k = 0; k < 4; k++) {
    double dd   = (double)(v.m[k].sum - v.m[k].prev);
    Mrad[k]     = __double2float_rn(__dmul_rn(__dmul_rn(dd, (double)RK_OUT_RAD_PER_RAW_ANGLE), (double)RK_GEAR_RATIO_INV));
    v.m[k].prev = v.m[k].sum;
  }
  float lx, ly;
  fk_xy(p, Mrad, lx, ly); // .th of the odometry increment is never used (:45-51)
  v.pos[0] = fadd(v.pos[0], fmul(fsub(fmul(lx, cth), fmul(ly, sth)), 0.001f));
  v.pos[1] = fadd(v.pos[1], fmul(fadd(fmul(lx, sth), fmul(ly, cth)), 0.001f));
#pragma unroll
  for(int a = 0; a < 3; a++) v.tgt[a] = interp_update(v.it[a], p.ts);
  ik(p, d, v.tgt, Mtgt);
  if(v.flags & RK_VS_FLAG_POWER_ON) {
#pragma unroll
    for(int k = 0; k < 4; k++) {
      v.c[k].now_tgt = fmul(Mtgt[k], RK_GEAR_RATIO);
      float u        = ctrl_update(v.c[k], p, d, fmul(Mvel[k], RK_GEAR_RATIO));
      motor_set_curr(v.m[k], p.motor_dir[k], p.raw_curr_lim, u);
    }
  } else {
#pragma unroll
    for(int a = 0; a < 3; a++) interp_reset(v.it[a]);
#pragma unroll
    for(int k = 0; k < 4; k++) {
      ctrl_reset(v.c[k]);
      motor_set_curr(v.m[k], p.motor_dir[k], p.raw_curr_lim, 0.0f);
    }
  }
}

// VEHICLE_CTRL::set_target_vel  VD_vehicle_controller.cpp:101-105
RK_DEV void veh_set_target(Veh &v, const float vv[3], const float a[3], const float j[3]) {
#pragma unroll
  for(int k = 0; k < 3; k++) interp_set(v.it[k], vv[k], a[k], j[k]);
}

// ---- VDT::main, the 100 Hz command layer (VD_task_main.cpp:119-151,165-322) ---------------------
// speed_limit :119-125 / rot_speed_limit :144-151
RK_DEV float vdt_speed_limit(const rk_vdt_params_t &p, uint32_t spd) {
  if(spd == 0u) return p.default_speed_mmps;
  const float f = __uint2float_rn(spd);
  return (f > p.limit_speed_mmps) ? p.limit_speed_mmps : f;
}
RK_DEV float vdt_rot_speed_limit(const rk_vdt_params_t &p, uint32_t spd) {
  if(spd == 0u) return p.default_rot_radps;
  const float f = __double2float_rn(__dmul_rn((double)__uint2float_rn(spd), 0.1)); // `(float)u32_spd * 0.1`: a double product
  return (f > p.limit_rot_radps) ? p.limit_rot_radps : f;
}
// one received MSG_REQ (the switch at :178-296); cq = {vx | u32_cmd, vy | u32_speed, vth, kind | time_ms << 8}
RK_DEV void vdt_task_message(Veh &v, const rk_vdt_params_t &p, uint4 cq) {
  const uint32_t kind = cq.w & 0xFFu, time_ms = cq.w >> 8;
  float          mv[3] = {0.0f, 0.0f, 0.0f};
  if(kind == RK_CMD_MSG_MOVE_DIR) {
    v.move_cnt = time_ms * p.task_freq_hz / 1000u + 1u;
    bool stop  = false;
    switch(cq.x) {
    case RK_DIR_GO_FORWARD: mv[0] = vdt_speed_limit(p, cq.y); break;
    case RK_DIR_GO_BACK: mv[0] = -vdt_speed_limit(p, cq.y); break;
    case RK_DIR_GO_RIGHT: mv[1] = -vdt_speed_limit(p, cq.y); break;
    case RK_DIR_GO_LEFT: mv[1] = vdt_speed_limit(p, cq.y); break;
    case RK_DIR_GO_RIGHT_FORWARD:
    case RK_DIR_GO_LEFT_FORWARD:
    case RK_DIR_GO_RIGHT_BACK:
    case RK_DIR_GO_LEFT_BACK: {
      // (+-(float)speed * sqrtf(2)) * 0.5f: the sign commutes with both products
      const float diag = fmul(fmul(vdt_speed_limit(p, cq.y), 1.41421354f /* sqrtf(2) */), 0.5f);
      mv[0] = (cq.x == RK_DIR_GO_RIGHT_FORWARD || cq.x == RK_DIR_GO_LEFT_FORWARD) ? diag : -diag;
      mv[1] = (cq.x == RK_DIR_GO_LEFT_FORWARD || cq.x == RK_DIR_GO_LEFT_BACK) ? diag : -diag;
    } break;
    case RK_DIR_ROT_RIGHT: mv[2] = -vdt_rot_speed_limit(p, cq.y); break;
    case RK_DIR_ROT_LEFT: mv[2] = vdt_rot_speed_limit(p, cq.y); break;
    default: stop = true; break; // MOVE_STOP and any other code
    }
    v.flags |= RK_VS_FLAG_POWER_ON;
    if(stop) veh_set_target(v, mv, p.accel_stop, p.jerk_stop);
    else veh_set_target(v, mv, p.accel_move, p.jerk_move);
  } else if(kind == RK_CMD_MSG_MOVE_CONT_DIR) {
    v.move_cnt     = time_ms * p.task_freq_hz / 1000u + 1u;
    const float vx = u2f(cq.x), vy = u2f(cq.y), vth = u2f(cq.z);
    const float len = arm_sqrt(fadd(fmul(vx, vx), fmul(vy, vy))); // speed_limit_xy :127-137
    const float lim = (len > p.limit_speed_mmps) ? p.limit_speed_mmps : len;
    if(len != 0.0f) mv[0] = fdiv(fmul(vx, lim), len), mv[1] = fdiv(fmul(vy, lim), len);
    mv[2] = (vth > p.limit_rot_radps) ? p.limit_rot_radps : ((vth < -p.limit_rot_radps) ? -p.limit_rot_radps : vth);
    v.flags |= RK_VS_FLAG_POWER_ON;
    veh_set_target(v, mv, p.accel_move, p.jerk_move);
  }
}
// the move-time countdown that ends every VDT::main iteration  :298-316
RK_DEV void vdt_task_countdown(Veh &v, const rk_vdt_params_t &p) {
  if(v.move_cnt > 1u) {
    v.move_cnt--;
  } else if(v.move_cnt == 1u) {
    const float z[3] = {0.0f, 0.0f, 0.0f};
    v.flags |= RK_VS_FLAG_POWER_ON;
    veh_set_target(v, z, p.accel_stop, p.jerk_stop);
    v.move_cnt = 0u;
  }
}

// Command / task events scheduled at tick t (commands BEFORE the tick, as rk_vdt_rollout_t documents).
struct Sched {
  int next_cmd, seg, next_task;
};
RK_DEV void sched_init(Sched &s, const rk_vdt_rollout_t &a) {
  const bool has_cmd = a.d_cmd != nullptr && a.seg_len > 0 && a.n_seg > 0;
  s.next_cmd = has_cmd ? 0 : INT_MAX, s.seg = 0;
  s.next_task = (a.task_period > 0) ? 0 : INT_MAX;
}
struct CmdRcp { // host-side RN(1 / x) of the command layer's acceleration / jerk constants (interp_set_rcp)
  float ra_move[3], rj_move[3], ra_stop[3], rj_stop[3];
};
RK_DEV void sched_events(Veh &v, const rk_vdt_params_t &p, const rk_vdt_rollout_t &a, int64_t n, int64_t i, int t, Sched &s,
                         const CmdRcp *rc = nullptr) {
  uint4 cq   = make_uint4(0u, 0u, 0u, 0u);
  bool  have = false;
  if(t == s.next_cmd) {
    cq   = __ldcs(reinterpret_cast<const uint4 *>(a.d_cmd) + (int64_t)s.seg * n + i);
    have = true;
    const uint32_t kind = cq.w & 0xFFu;
    if(kind == RK_CMD_MOVE || kind == RK_CMD_STOP) { // VDT::main -> start(); set_target_vel()   VD_task_main.cpp:280-281,294-295
      const float vv[3] = {u2f(cq.x), u2f(cq.y), u2f(cq.z)};
      v.flags |= RK_VS_FLAG_POWER_ON;
      if(rc) {
        const bool stop = kind == RK_CMD_STOP;
#pragma unroll
        for(int k = 0; k < 3; k++)
          interp_set_rcp(v.it[k], vv[k], stop ? p.accel_stop[k] : p.accel_move[k], stop ? p.jerk_stop[k] : p.jerk_move[k],
                         stop ? rc->ra_stop[k] : rc->ra_move[k], stop ? rc->rj_stop[k] : rc->rj_move[k]);
      } else if(kind == RK_CMD_STOP) veh_set_target(v, vv, p.accel_stop, p.jerk_stop);
      else veh_set_target(v, vv, p.accel_move, p.jerk_move);
    }
    s.seg++;
    s.next_cmd = (s.seg < a.n_seg) ? s.next_cmd + a.seg_len : INT_MAX;
  }
  if(t == s.next_task) { // one VDT::main iteration: message (if any), then the countdown
    if(have) vdt_task_message(v, p, cq);
    vdt_task_countdown(v, p);
    s.next_task += a.task_period;
  }
}
// tick at which the countdown will issue its stop, given the state after the events of the current tick
RK_DEV int sched_fire_tick(const Veh &v, const rk_vdt_rollout_t &a, const Sched &s) {
  if(s.next_task == INT_MAX || v.move_cnt == 0u) return INT_MAX;
  const long long f = (long long)s.next_task + (long long)(v.move_cnt - 1u) * (long long)a.task_period;
  return f > (long long)INT_MAX ? INT_MAX : (int)f;
}
// after jumping from a tick < t_now to t_now without visiting the task boundaries in between: each of them
// only decremented the countdown (the jump never crosses sched_fire_tick)
RK_DEV void sched_skip_to(Veh &v, const rk_vdt_rollout_t &a, Sched &s, int t_now) {
  if(s.next_task == INT_MAX || t_now <= s.next_task) return;
  const int passed = (t_now - 1 - s.next_task) / a.task_period + 1;
  if(v.move_cnt > 0u) v.move_cnt -= (uint32_t)passed;
  s.next_task += passed * a.task_period;
}
// the countdown as it stands after tick t's events, for the trace (word 13)
RK_DEV uint32_t sched_cnt_at(uint32_t move_cnt, const rk_vdt_rollout_t &a, const Sched &s, int t) {
  if(a.task_period <= 0) return 0u;
  if(move_cnt == 0u || t < s.next_task) return move_cnt;
  return move_cnt - (uint32_t)((t - s.next_task) / a.task_period + 1);
}

} // namespace rk
