/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * Plain-C restatement of the reference's numeric control tick.  Every function cites the
 * reference file:line it follows (paths relative to the reference tree).  It is compiled
 * with the pinned flags of oracle/Makefile (-O2 -ffp-contract=off, SSE float evaluation),
 * and is PINNED by tests/test_oracle_pin.py against
 *   (a) oracle/_ref -- the reference's own sources compiled unmodified for x86 -- on seeded
 *       streams, state word for state word, bit-exact; and
 *   (b) the golden fixtures in tests/golden/ (generated from oracle/_ref by
 *       tests/golden/make_golden.py) which include the SURVEY.md Appendix D probe values.
 * The reference itself ships no tests or golden vectors (test/README is a placeholder).
 *
 * One boundary stays UNPINNED: CMSIS-DSP arm_sin_f32/arm_cos_f32 (un-vendored, un-pinned
 * dependency; restated from the published algorithm -- see oracle/cmsis_shim.c).  It only
 * influences pos.x / pos.y.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 */
#include "robotick_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* constants: VD_motor_if_m2006.hpp:76-82, util_mymath.hpp + CMSIS PI                    */
#define ORC_PI 3.14159265358979f
static const float RPM_TO_RADPS          = 2.0f * 3.1415926f / 60.0f;
static const float AMPERE_TO_RAW_CURR    = 1000.0f;
static const float GEAR_RATIO            = 36.0f;
static const float GEAR_RATIO_INV        = 1.0f / 36.0f;
static const float OUT_RAD_PER_RAW_ANGLE = 2.0f * 3.1415926f / 8191.0f;

typedef struct {
  float vel, acl, vel_tgt, acl_max, jerk_p, jerk_m, dt1, dt2, dt3, vel_ini, acl_ini, dt;
} interp_t;
typedef struct {
  float prev_val, integ, lpf_y, lpf_x, now_tgt, now_err, now_ctrl;
} ctrl_t;
typedef struct {
  int64_t sum, prev;
  int16_t ang, rpm, cur, cur_tgt, usec;
  uint8_t head;
  int32_t p_ang, p_rpm;
} motor_t;
typedef struct {
  float    pos[3], vel[3], tgt[3];
  int      power;
  uint32_t move_cnt; /* VDT::U32_MOVE_TIME_CNT_ORDER */
  interp_t it[3];
  ctrl_t   c[4];
  motor_t  m[4];
} veh_t;

static uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
static int      lo16(uint32_t w) { return (int16_t)(w & 0xFFFFu); }
static int      hi16(uint32_t w) { return (int16_t)(w >> 16); }

static void unpack(veh_t *v, const uint32_t *w) {
  int a, k;
  for(a = 0; a < 3; a++) {
    v->pos[a] = u2f(w[RK_VS_POS_X + a]);
    v->vel[a] = u2f(w[RK_VS_VEL_X + a]);
    v->tgt[a] = u2f(w[RK_VS_TGT_X + a]);
  }
  v->power = (w[RK_VS_FLAGS] & RK_VS_FLAG_POWER_ON) != 0;
  v->move_cnt = w[RK_VS_MOVE_CNT];
  for(a = 0; a < 3; a++) {
    const uint32_t *q = w + RK_VS_INTERP0 + 12 * a;
    interp_t       *t = &v->it[a];
    t->vel = u2f(q[RK_VI_VEL_NOW]), t->acl = u2f(q[RK_VI_ACL_NOW]);
    t->vel_tgt = u2f(q[RK_VI_VEL_TGT]), t->acl_max = u2f(q[RK_VI_ACL_MAX]);
    t->jerk_p = u2f(q[RK_VI_JERK_P]), t->jerk_m = u2f(q[RK_VI_JERK_M]);
    t->dt1 = u2f(q[RK_VI_DT1]), t->dt2 = u2f(q[RK_VI_DT2]), t->dt3 = u2f(q[RK_VI_DT3]);
    t->vel_ini = u2f(q[RK_VI_VEL_INI]), t->acl_ini = u2f(q[RK_VI_ACL_INI]), t->dt = u2f(q[RK_VI_DT]);
  }
  for(k = 0; k < 4; k++) {
    const uint32_t *q = w + RK_VS_CTRL0 + 8 * k;
    ctrl_t         *c = &v->c[k];
    c->prev_val = u2f(q[RK_VC_PREV_VAL]), c->integ = u2f(q[RK_VC_INTEG]);
    c->lpf_y = u2f(q[RK_VC_LPF_Y]), c->lpf_x = u2f(q[RK_VC_LPF_X]);
    c->now_tgt = u2f(q[RK_VC_NOW_TGT]), c->now_err = u2f(q[RK_VC_NOW_ERR]), c->now_ctrl = u2f(q[RK_VC_NOW_CTRL]);
  }
  for(k = 0; k < 4; k++) {
    const uint32_t *q = w + RK_VS_MOTOR0 + 8 * k;
    motor_t        *m = &v->m[k];
    m->sum     = (int64_t)(((uint64_t)q[RK_VM_SUM_HI] << 32) | q[RK_VM_SUM_LO]);
    m->prev    = (int64_t)(((uint64_t)q[RK_VM_PREV_HI] << 32) | q[RK_VM_PREV_LO]);
    m->ang     = (int16_t)lo16(q[RK_VM_ANG_RPM]);
    m->rpm     = (int16_t)hi16(q[RK_VM_ANG_RPM]);
    m->cur     = (int16_t)lo16(q[RK_VM_CUR_TGT]);
    m->cur_tgt = (int16_t)hi16(q[RK_VM_CUR_TGT]);
    m->usec    = (int16_t)lo16(q[RK_VM_USEC]);
    m->head    = (uint8_t)(hi16(q[RK_VM_USEC]) % 3);
    m->p_ang   = lo16(q[RK_VM_PLANT]);
    m->p_rpm   = hi16(q[RK_VM_PLANT]);
  }
}

static void pack(const veh_t *v, uint32_t *w) {
  int a, k;
  memset(w, 0, sizeof(uint32_t) * RK_VS_WORDS);
  for(a = 0; a < 3; a++) {
    w[RK_VS_POS_X + a] = f2u(v->pos[a]);
    w[RK_VS_VEL_X + a] = f2u(v->vel[a]);
    w[RK_VS_TGT_X + a] = f2u(v->tgt[a]);
  }
  w[RK_VS_FLAGS] = v->power ? RK_VS_FLAG_POWER_ON : 0u;
  w[RK_VS_MOVE_CNT] = v->move_cnt;
  for(a = 0; a < 3; a++) {
    uint32_t       *q = w + RK_VS_INTERP0 + 12 * a;
    const interp_t *t = &v->it[a];
    q[RK_VI_VEL_NOW] = f2u(t->vel), q[RK_VI_ACL_NOW] = f2u(t->acl);
    q[RK_VI_VEL_TGT] = f2u(t->vel_tgt), q[RK_VI_ACL_MAX] = f2u(t->acl_max);
    q[RK_VI_JERK_P] = f2u(t->jerk_p), q[RK_VI_JERK_M] = f2u(t->jerk_m);
    q[RK_VI_DT1] = f2u(t->dt1), q[RK_VI_DT2] = f2u(t->dt2), q[RK_VI_DT3] = f2u(t->dt3);
    q[RK_VI_VEL_INI] = f2u(t->vel_ini), q[RK_VI_ACL_INI] = f2u(t->acl_ini), q[RK_VI_DT] = f2u(t->dt);
  }
  for(k = 0; k < 4; k++) {
    uint32_t     *q = w + RK_VS_CTRL0 + 8 * k;
    const ctrl_t *c = &v->c[k];
    q[RK_VC_PREV_VAL] = f2u(c->prev_val), q[RK_VC_INTEG] = f2u(c->integ);
    q[RK_VC_LPF_Y] = f2u(c->lpf_y), q[RK_VC_LPF_X] = f2u(c->lpf_x);
    q[RK_VC_NOW_TGT] = f2u(c->now_tgt), q[RK_VC_NOW_ERR] = f2u(c->now_err), q[RK_VC_NOW_CTRL] = f2u(c->now_ctrl);
  }
  for(k = 0; k < 4; k++) {
    uint32_t      *q = w + RK_VS_MOTOR0 + 8 * k;
    const motor_t *m = &v->m[k];
    q[RK_VM_SUM_LO]  = (uint32_t)(uint64_t)m->sum;
    q[RK_VM_SUM_HI]  = (uint32_t)((uint64_t)m->sum >> 32);
    q[RK_VM_PREV_LO] = (uint32_t)(uint64_t)m->prev;
    q[RK_VM_PREV_HI] = (uint32_t)((uint64_t)m->prev >> 32);
    q[RK_VM_ANG_RPM] = pack16(m->ang, m->rpm);
    q[RK_VM_CUR_TGT] = pack16(m->cur, m->cur_tgt);
    q[RK_VM_USEC]    = pack16(m->usec, m->head);
    q[RK_VM_PLANT]   = pack16(m->p_ang, m->p_rpm);
  }
}

/* ------------------------------------------------------------------------------------ */
/* CMSIS-DSP sin/cos restatement (un-vendored dependency; see oracle/cmsis_shim.c)       */
static const float sin_table[513] = {
#include "cmsis_sin_table.inc"
};
static float table_lerp(float in) {
  int32_t  n = (int32_t)in;
  float    findex, fract;
  uint16_t index;
  if(in < 0.0f) n--;
  in     = in - (float)n;
  findex = 512.0f * in;
  index  = (uint16_t)findex;
  if(index >= 512) {
    index = 0;
    findex -= 512.0f;
  }
  fract = findex - (float)index;
  return (1.0f - fract) * sin_table[index] + fract * sin_table[index + 1];
}
float orc_sin(float x) { return table_lerp(x * 0.159154943092f); }
float orc_cos(float x) { return table_lerp(x * 0.159154943092f + 0.25f); }
/* arm_sqrt_f32 on an FPU core: IEEE sqrt, negative -> 0 (util_vel_interp.hpp:90) */
static float orc_sqrt(float x) { return (x >= 0.0f) ? sqrtf(x) : 0.0f; }

/* util_mymath.hpp:18-25 */
float orc_normalize_rad_0to2pi(float d) {
  if(d < 0.0f || d >= 2.0f * ORC_PI) {
    int mod = (int)(d / (2.0f * ORC_PI));
    d -= (mod * 2.0f * ORC_PI);
    if(d < 0.0f) d = d + 2.0f * ORC_PI;
  }
  return d;
}
/* util_mymath.hpp:27-34 */
float orc_normalize_deg_0to360(float d) {
  if(d < 0.0f || d >= 360.0f) {
    int mod = (int)(d / (360.0f));
    d -= (mod * 360.0f);
    if(d < 0.0f) d = d + 360.0f;
  }
  return d;
}

/* ------------------------------------------------------------------------------------ */
/* VelInterpConstJerk::set_target_params   util_vel_interp.hpp:53-108 (writes the other page
 * and flips; only the page that becomes active is state)                                 */
static void interp_set(interp_t *t, float v_t, float a_m, float jrk) {
  float vel_tgt = v_t, acl_max = a_m, vel_ini = t->vel, acl_ini = t->acl;
  float jerk_m, jerk_p, jm_inv, jp_inv, dt1, dt2, dt3;
  if((vel_tgt - vel_ini) < 0) acl_max = -a_m;
  jerk_m = (acl_max >= 0) ? -jrk : jrk;
  jm_inv = 1.0f / jerk_m;
  jerk_p = (acl_max - acl_ini >= 0) ? jrk : -jrk;
  jp_inv = 1.0f / jerk_p;
  dt1    = (acl_max - acl_ini) * jp_inv;
  dt3    = acl_max * (-jm_inv);
  dt2    = 1.0f / acl_max * (vel_tgt - vel_ini - acl_ini * dt1 * 0.5f - acl_max * (dt1 + dt3) * 0.5f);
  if(dt2 < 0.0f) {
    float sq_in = (acl_ini * jp_inv) * (acl_ini * jp_inv) * 0.5f + (vel_tgt - vel_ini) * jp_inv;
    float sq    = orc_sqrt(sq_in);
    dt1         = sq - acl_ini * jp_inv;
    acl_max     = acl_ini + jerk_p * dt1;
    dt2         = 0.0f;
    dt3         = acl_max * (-jm_inv);
  }
  dt1 = (dt1 < 0.0f) ? 0.0f : dt1;
  dt3 = (dt3 < 0.0f) ? 0.0f : dt3;
  t->vel_tgt = vel_tgt, t->acl_max = acl_max, t->jerk_p = jerk_p, t->jerk_m = jerk_m;
  t->dt1 = dt1, t->dt2 = dt2, t->dt3 = dt3, t->vel_ini = vel_ini, t->acl_ini = acl_ini, t->dt = 0.0f;
}
/* VelInterpConstJerk::update   util_vel_interp.hpp:110-136 */
static float interp_update(interp_t *t, float ts) {
  if(t->dt <= t->dt1 + ts) {
    t->acl = t->acl_ini + t->jerk_p * t->dt;
    t->vel = t->vel_ini + (t->acl_ini + t->acl) * t->dt * 0.5f;
    t->dt  = t->dt + ts;
  } else if(t->dt <= t->dt1 + t->dt2 + ts) {
    t->acl = t->acl_max;
    t->vel = t->vel + t->acl * ts;
    t->dt  = t->dt + ts;
  } else if(t->dt <= t->dt1 + t->dt2 + t->dt3 + ts) {
    t->acl = t->acl_max + t->jerk_m * (t->dt - t->dt1 - t->dt2);
    t->vel = t->vel + t->acl * ts;
    t->dt  = t->dt + ts;
  } else {
    t->acl = 0.0f;
    t->vel = t->vel_tgt;
  }
  return t->vel;
}
/* VelInterpConstJerk::reset   util_vel_interp.hpp:138-143 */
static void interp_reset(interp_t *t) { memset(t, 0, sizeof(*t)); }

/* FF_PI_D::update -> PI_D::update -> IIR1::update
 * util_controller.hpp:159-165, :94-110, ctor :90-92 ; util_iir.hpp:39-45 */
static float ctrl_update(ctrl_t *c, const rk_vdt_params_t *p, float nowval) {
  const float freq = p->ctrl_freq, dt = 1.0f / p->ctrl_freq, lpf = p->lpf_freq;
  const float A1 = (2.0f * freq - lpf) / (2.0f * freq + lpf);
  const float B0 = lpf / (2.0f * freq + lpf), B1 = lpf / (2.0f * freq + lpf);
  float       now_val = nowval, err, x, y, ctrl, ff;
  err       = c->now_tgt - now_val;
  x         = (now_val - c->prev_val) * freq;
  y         = A1 * c->lpf_y + B0 * x + B1 * c->lpf_x;
  c->lpf_y  = y;
  c->lpf_x  = x;
  c->integ += p->ki * dt * err;
  c->integ  = (c->integ >= p->i_limit) ? p->i_limit : ((c->integ <= -p->i_limit) ? -p->i_limit : c->integ);
  ctrl      = p->kp * err + c->integ - p->kd * y;
  c->prev_val = now_val;
  c->now_err  = err;
  ff          = c->now_tgt * p->kff;
  ff          = (ff >= p->ff_limit) ? p->ff_limit : ((ff <= -p->ff_limit) ? -p->ff_limit : ff);
  ctrl        = ctrl + ff;
  c->now_ctrl = ctrl;
  return ctrl;
}
/* PI_D::reset  util_controller.hpp:112-124 */
static void ctrl_reset(ctrl_t *c) { memset(c, 0, sizeof(*c)); }

/* MOTOR_IF_M2006::set_CurrA_tgt -> set_rawCurr_tgt -> sat_curr  VD_motor_if_m2006.hpp:36-37,57
 * (int16_t)(float) is done as x86 does it: cvttss2si to 32 bits, keep the low 16. */
static void motor_set_curr(motor_t *m, int dir, int lim, float amp) {
  int16_t t  = (int16_t)(int32_t)(amp * AMPERE_TO_RAW_CURR);
  int16_t c  = (int16_t)(t * dir);
  int16_t l  = (int16_t)lim;
  m->cur_tgt = (c > l) ? l : ((c < -l) ? (int16_t)-l : c);
}

/* MOTOR_IF_M2006::rx_callback  VD_motor_if_m2006.cpp:32-72 (integer part; the float fields
 * flt_SpeedRadPS / flt_dltOutAngle_rad are dead: VD_vehicle_controller.cpp:20-24 `#if 1`) */
static void motor_rx(motor_t *m, int dir, const uint8_t f[8], int16_t usec) {
  uint8_t w = (uint8_t)(m->head + 1);
  int16_t a, raw_ang, d;
  if(w >= 3) w = 0;
  a = (int16_t)((f[0] << 8) | f[1]);
  if(dir == 1)
    raw_ang = a;
  else
    raw_ang = (int16_t)(8192 - a);
  d       = (int16_t)(raw_ang - m->ang);
  d       = (d > 4096) ? (int16_t)(d - 8192) : ((d < -4096) ? (int16_t)(d + 8192) : d);
  m->sum  = m->sum + d;
  m->ang  = raw_ang;
  m->rpm  = (int16_t)((int16_t)((f[2] << 8) | f[3]) * dir);
  m->cur  = (int16_t)((int16_t)((f[4] << 8) | f[5]) * dir);
  m->usec = usec;
  m->head = w;
}

/* VEHICLE_CTRL::conv_Mdir_to_Vdir  VD_vehicle_controller.cpp:126-130 */
static void fk(const rk_vdt_params_t *p, const float M[4], float V[3]) {
  const float R = p->wheel_radius_mm, L = p->wheel_l_mm, S2 = p->sqrtf2;
  V[0] = (M[0] + M[1] + M[2] + M[3]) * 0.25f * R;
  V[1] = (-M[0] + M[1] - M[2] + M[3]) * 0.25f * R;
  V[2] = (-M[0] - M[1] + M[2] + M[3]) * 0.25f / S2 / L * R;
}
/* VEHICLE_CTRL::conv_Vdir_to_Mdir  VD_vehicle_controller.cpp:113-118 */
static void ik(const rk_vdt_params_t *p, const float V[3], float M[4]) {
  const float R = p->wheel_radius_mm, L = p->wheel_l_mm, S2 = p->sqrtf2;
  M[0] = (V[0] - V[1] - S2 * L * V[2] * 4.0f) / R;
  M[1] = (V[0] + V[1] - S2 * L * V[2] * 4.0f) / R;
  M[2] = (V[0] - V[1] + S2 * L * V[2] * 4.0f) / R;
  M[3] = (V[0] + V[1] + S2 * L * V[2] * 4.0f) / R;
}

/* VEHICLE_CTRL::update  VD_vehicle_controller.cpp:6-99 */
static void veh_update(veh_t *v, const rk_vdt_params_t *p) {
  float Mvel[4], Mrad[4], loc[3], Mtgt[4], rad, c, s;
  int   k;
  for(k = 0; k < 4; k++) Mvel[k] = (float)v->m[k].rpm * RPM_TO_RADPS * GEAR_RATIO_INV;
  fk(p, Mvel, v->vel);
  for(k = 0; k < 4; k++) {
    Mrad[k]      = (float)((double)(v->m[k].sum - v->m[k].prev) * OUT_RAD_PER_RAW_ANGLE * GEAR_RATIO_INV);
    v->m[k].prev = v->m[k].sum;
  }
  fk(p, Mrad, loc);
  rad       = orc_normalize_rad_0to2pi(v->pos[2]);
  c         = orc_cos(rad);
  s         = orc_sin(rad);
  v->pos[0] = v->pos[0] + (loc[0] * c - loc[1] * s) * 0.001f;
  v->pos[1] = v->pos[1] + (loc[0] * s + loc[1] * c) * 0.001f;
  for(k = 0; k < 3; k++) v->tgt[k] = interp_update(&v->it[k], p->ts);
  ik(p, v->tgt, Mtgt);
  if(v->power) {
    for(k = 0; k < 4; k++) v->c[k].now_tgt = Mtgt[k] * GEAR_RATIO;
    for(k = 0; k < 4; k++) motor_set_curr(&v->m[k], p->motor_dir[k], p->raw_curr_lim, ctrl_update(&v->c[k], p, Mvel[k] * GEAR_RATIO));
  } else {
    for(k = 0; k < 3; k++) interp_reset(&v->it[k]);
    for(k = 0; k < 4; k++) ctrl_reset(&v->c[k]);
    for(k = 0; k < 4; k++) motor_set_curr(&v->m[k], p->motor_dir[k], p->raw_curr_lim, 0.0f);
  }
}

/* synthetic plant of robotick.h (RK_SENSOR_PLANT) -- not part of the reference */
static void plant_frame(motor_t *m, uint8_t f[8]) {
  int32_t cur = m->cur_tgt, rpm = m->p_rpm, ang = m->p_ang;
  rpm += ((cur * 4 - rpm) >> 4);
  ang      = (ang + rpm * 8192 / 60000) & 8191;
  m->p_rpm = rpm, m->p_ang = ang;
  f[0] = (uint8_t)(ang >> 8), f[1] = (uint8_t)ang, f[2] = (uint8_t)(rpm >> 8), f[3] = (uint8_t)rpm;
  f[4] = (uint8_t)(cur >> 8), f[5] = (uint8_t)cur, f[6] = 0, f[7] = 0;
}

/* VEHICLE_CTRL::set_target_vel  VD_vehicle_controller.cpp:101-105 */
static void veh_set_target(veh_t *v, const float vv[3], const float a[3], const float j[3]) {
  int k;
  for(k = 0; k < 3; k++) interp_set(&v->it[k], vv[k], a[k], j[k]);
}

/* VDT::main's limiters  VD_task_main.cpp:119-151 */
static float vdt_speed_limit(const rk_vdt_params_t *p, uint32_t spd) {
  if(spd == 0) return p->default_speed_mmps;
  return ((float)spd > p->limit_speed_mmps) ? p->limit_speed_mmps : (float)spd;
}
static float vdt_rot_speed_limit(const rk_vdt_params_t *p, uint32_t spd) {
  float fl;
  if(spd == 0) return p->default_rot_radps;
  fl = (float)((double)(float)spd * 0.1); /* `(float)u32_spd * 0.1` is a double product, :146 */
  return (fl > p->limit_rot_radps) ? p->limit_rot_radps : fl;
}
/* one received MSG_REQ: the switch of VDT::main  VD_task_main.cpp:178-296 */
static void vdt_task_message(veh_t *v, const rk_vdt_params_t *p, const rk_vdt_cmd_t *c) {
  uint32_t kind = (uint32_t)c->kind & 0xFFu, time_ms = (uint32_t)c->kind >> 8;
  float    mv[3] = {0.0f, 0.0f, 0.0f};
  const float *acl = p->accel_move, *jrk = p->jerk_move;
  if(kind == RK_CMD_MSG_MOVE_DIR) {
    uint32_t u[2];
    float    speed, diag;
    memcpy(u, &c->vx, 8);
    v->move_cnt = time_ms * p->task_freq_hz / 1000u + 1u;
    switch(u[0]) {
    case RK_DIR_GO_FORWARD: mv[0] = vdt_speed_limit(p, u[1]); break;
    case RK_DIR_GO_BACK: mv[0] = -vdt_speed_limit(p, u[1]); break;
    case RK_DIR_GO_RIGHT: mv[1] = -vdt_speed_limit(p, u[1]); break;
    case RK_DIR_GO_LEFT: mv[1] = vdt_speed_limit(p, u[1]); break;
    case RK_DIR_GO_RIGHT_FORWARD:
    case RK_DIR_GO_LEFT_FORWARD:
    case RK_DIR_GO_RIGHT_BACK:
    case RK_DIR_GO_LEFT_BACK:
      speed = vdt_speed_limit(p, u[1]);
      diag  = speed * sqrtf(2) * 0.5f; /* (+-(float)speed * sqrtf(2)) * 0.5f: the sign commutes with both products */
      mv[0] = (u[0] == RK_DIR_GO_RIGHT_FORWARD || u[0] == RK_DIR_GO_LEFT_FORWARD) ? diag : -diag;
      mv[1] = (u[0] == RK_DIR_GO_LEFT_FORWARD || u[0] == RK_DIR_GO_LEFT_BACK) ? diag : -diag;
      break;
    case RK_DIR_ROT_RIGHT: mv[2] = -vdt_rot_speed_limit(p, u[1]); break;
    case RK_DIR_ROT_LEFT: mv[2] = vdt_rot_speed_limit(p, u[1]); break;
    default: acl = p->accel_stop, jrk = p->jerk_stop; break; /* MOVE_STOP and anything else */
    }
    v->power = 1;
    veh_set_target(v, mv, acl, jrk);
  } else if(kind == RK_CMD_MSG_MOVE_CONT_DIR) {
    float len, lim;
    v->move_cnt = time_ms * p->task_freq_hz / 1000u + 1u;
    len = orc_sqrt(c->vx * c->vx + c->vy * c->vy); /* speed_limit_xy :127-137 */
    lim = (len > p->limit_speed_mmps) ? p->limit_speed_mmps : len;
    if(len == 0) {
      mv[0] = 0, mv[1] = 0;
    } else {
      mv[0] = c->vx * lim / len, mv[1] = c->vy * lim / len;
    }
    mv[2] = (c->vth > p->limit_rot_radps) ? p->limit_rot_radps : ((c->vth < -p->limit_rot_radps) ? -p->limit_rot_radps : c->vth);
    v->power = 1;
    veh_set_target(v, mv, p->accel_move, p->jerk_move);
  }
}
/* the move-time countdown at the end of every VDT::main iteration  :298-316 */
static void vdt_task_countdown(veh_t *v, const rk_vdt_params_t *p) {
  if(v->move_cnt > 1) {
    v->move_cnt--;
  } else if(v->move_cnt == 1) {
    float z[3] = {0.0f, 0.0f, 0.0f};
    v->power   = 1;
    veh_set_target(v, z, p->accel_stop, p->jerk_stop);
    v->move_cnt = 0;
  }
}

static void rollout_one(veh_t *v, const rk_vdt_params_t *p, int64_t n, int64_t i, const rk_vdt_rollout_t *a) {
  int t, k, j;
  for(t = 0; t < a->steps; t++) {
    int16_t us;
    const rk_vdt_cmd_t *c = NULL;
    if(a->d_cmd && a->seg_len > 0 && (t % a->seg_len) == 0 && (t / a->seg_len) < a->n_seg) c = &a->d_cmd[(int64_t)(t / a->seg_len) * n + i];
    if(c && ((uint32_t)c->kind & 0xFFu) != RK_CMD_NONE && ((uint32_t)c->kind & 0xFFu) < RK_CMD_MSG_MOVE_DIR) {
      float vv[3] = {c->vx, c->vy, c->vth};
      v->power    = 1; /* VEHICLE_CTRL::start()  VD_vehicle_controller.hpp:54 */
      if(c->kind == RK_CMD_STOP)
        veh_set_target(v, vv, p->accel_stop, p->jerk_stop);
      else
        veh_set_target(v, vv, p->accel_move, p->jerk_move);
    }
    if(a->task_period > 0 && (t % a->task_period) == 0) { /* one VDT::main iteration */
      if(c) vdt_task_message(v, p, c);
      vdt_task_countdown(v, p);
    }
    if((a->d_yaw || a->d_yaw_reg) && a->yaw_period > 0 && (t % a->yaw_period) == 0 && (t / a->yaw_period) < a->n_yaw) {
      const int64_t yi = (int64_t)(t / a->yaw_period) * n + i;
      /* set_now_yaw_world :57; from the Yaw register: imu_if_wt901c.cpp:100, getYawDate :160, deg2rad VD_task_main.cpp:368 */
      v->pos[2] = a->d_yaw ? a->d_yaw[yi] : ((float)a->d_yaw_reg[yi] / 32768.0f * 180.0f) * (ORC_PI / 180.0f);
    }
    us = (int16_t)(((t + 1) * 1000) & 0x7FFF);
    if(a->sensor_mode == RK_SENSOR_PLANT) {
      for(k = 0; k < 4; k++) {
        uint8_t f[8];
        plant_frame(&v->m[k], f);
        motor_rx(&v->m[k], p->motor_dir[k], f, us);
      }
    } else if(a->sensor_mode == RK_SENSOR_STREAM) {
      for(k = 0; k < 4; k++) {
        uint8_t f[8];
        memcpy(f, &a->d_frames[((int64_t)t * 4 + k) * n + i], 8);
        motor_rx(&v->m[k], p->motor_dir[k], f, us);
      }
    }
    veh_update(v, p);
    if(a->d_trace) {
      uint32_t *tr = a->d_trace + (int64_t)t * RK_VDT_TRACE_WORDS * n + i;
      for(j = 0; j < 3; j++) {
        tr[(int64_t)j * n]       = f2u(v->pos[j]);
        tr[(int64_t)(3 + j) * n] = f2u(v->vel[j]);
        tr[(int64_t)(6 + j) * n] = f2u(v->tgt[j]);
      }
      for(k = 0; k < 4; k++) tr[(int64_t)(9 + k) * n] = (uint32_t)(int32_t)v->m[k].cur_tgt;
      tr[(int64_t)13 * n] = (a->task_period > 0) ? v->move_cnt : 0u;
      { /* CAN_CTRL::tx_routine  VD_can_controller.hpp:43-55: the C610 frame (id 0x200), four currents big-endian */
        uint8_t b[8];
        for(k = 0; k < 4; k++) b[2 * k] = (uint8_t)(v->m[k].cur_tgt >> 8), b[2 * k + 1] = (uint8_t)(v->m[k].cur_tgt & 0x00FF);
        tr[(int64_t)14 * n] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        tr[(int64_t)15 * n] = (uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24);
      }
    }
  }
  if(a->d_cost && a->d_goal) {
    float dx = v->pos[0] - a->d_goal[2 * i], dy = v->pos[1] - a->d_goal[2 * i + 1];
    a->d_cost[i] = dx * dx + dy * dy;
  }
}

static uint32_t *soa(uint32_t *blk, int64_t n, int64_t i, int w) { return &blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }

typedef struct {
  const rk_vdt_params_t  *p;
  uint32_t               *state;
  int64_t                 n, i0, i1;
  const rk_vdt_rollout_t *args;
  int                     tid, nthreads;
} job_t;

static void *job_main(void *arg) {
  job_t   *jb = (job_t *)arg;
  uint32_t w[RK_VS_WORDS];
  veh_t    v;
  int64_t  i;
  int      k;
  for(i = jb->i0 + jb->tid; i < jb->i1; i += jb->nthreads) {
    if(jb->state) {
      for(k = 0; k < RK_VS_WORDS; k++) w[k] = *soa(jb->state, jb->n, i, k);
    } else {
      memset(w, 0, sizeof(w));
    }
    unpack(&v, w);
    rollout_one(&v, jb->p, jb->n, i, jb->args);
    if(jb->state) {
      pack(&v, w);
      for(k = 0; k < RK_VS_WORDS; k++) *soa(jb->state, jb->n, i, k) = w[k];
    }
  }
  return NULL;
}

void orc_vdt_rollout(const rk_vdt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1,
                     const rk_vdt_rollout_t *args, int nthreads) {
  job_t     jobs[256];
  pthread_t th[256];
  int       t;
  if(nthreads < 1) nthreads = 1;
  if(nthreads > 256) nthreads = 256;
  for(t = 0; t < nthreads; t++) {
    jobs[t].p = p, jobs[t].state = state, jobs[t].n = n, jobs[t].i0 = i0, jobs[t].i1 = i1;
    jobs[t].args = args, jobs[t].tid = t, jobs[t].nthreads = nthreads;
  }
  if(nthreads == 1) {
    job_main(&jobs[0]);
    return;
  }
  for(t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, job_main, &jobs[t]);
  for(t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

void orc_vdt_set_target(const rk_vdt_params_t *p, uint32_t *words, const float v[3], const float a[3], const float j[3]) {
  veh_t s;
  (void)p;
  unpack(&s, words);
  veh_set_target(&s, v, a, j);
  pack(&s, words);
}
void orc_vdt_rx(const rk_vdt_params_t *p, uint32_t *words, int wheel, const uint8_t frame[8], int16_t usec_id) {
  veh_t s;
  unpack(&s, words);
  motor_rx(&s.m[wheel], p->motor_dir[wheel], frame, usec_id);
  pack(&s, words);
}
void orc_vdt_update(const rk_vdt_params_t *p, uint32_t *words) {
  veh_t s;
  unpack(&s, words);
  veh_update(&s, p);
  pack(&s, words);
}

/* ------------------------------------------------------------------------------------ */
/* IMU_IF_WT901C::updateData   src/Imu/imu_if_wt901c.cpp:91-129                           */
static void imu_update_data(const float qi[4], const int16_t r[16], float d[16]) {
  float a[3], g[3], m[3], e[3], q[4];
  int   i;
  for(i = 0; i < 3; i++) {
    a[i] = (float)r[RK_IMT_REG_AX + i] / 32768.0f * 16.0f;
    g[i] = (float)r[RK_IMT_REG_GX + i] / 32768.0f * 2000.0f;
    m[i] = (float)r[RK_IMT_REG_HX + i];
    e[i] = (float)r[RK_IMT_REG_ROLL + i] / 32768.0f * 180.0f;
  }
  for(i = 0; i < 4; i++) q[i] = r[RK_IMT_REG_Q0 + i] / 32768.0f;
  d[0] = a[0], d[1] = -a[1], d[2] = -a[2];
  d[3] = g[0], d[4] = -g[1], d[5] = -g[2];
  d[6] = m[0], d[7] = -m[1], d[8] = -m[2];
  d[9]  = orc_normalize_deg_0to360(e[0]) - 180.0f;
  d[10] = e[1];
  d[11] = e[2];
  d[14] = -(qi[3] * q[0] + qi[2] * q[1] - qi[1] * q[2] - qi[0] * q[3]);
  d[13] = (-qi[2] * q[0] + qi[3] * q[1] + qi[0] * q[2] - qi[1] * q[3]);
  d[12] = -(qi[1] * q[0] - qi[0] * q[1] + qi[3] * q[2] - qi[2] * q[3]);
  d[15] = (qi[0] * q[0] + qi[1] * q[1] + qi[2] * q[2] + qi[3] * q[3]);
}

void orc_imt_update(uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const int16_t *regs,
                    const uint8_t *have_quat, uint32_t *out, int do_init) {
  int64_t i;
  int     u, k;
  for(i = i0; i < i1; i++) {
    float    qi[4] = {0, 0, 0, 0}, d[16];
    uint32_t flags = 0;
    memset(d, 0, sizeof(d));
    if(state) {
      for(k = 0; k < 4; k++) qi[k] = u2f(*soa(state, n, i, RK_IS_QINIT + k));
      for(k = 0; k < 16; k++) d[k] = u2f(*soa(state, n, i, RK_IS_DATA + k));
      flags = *soa(state, n, i, RK_IS_FLAGS);
    }
    for(u = 0; u < K; u++) {
      int16_t r[16];
      int     hq = have_quat ? have_quat[(int64_t)u * n + i] : 1;
      for(k = 0; k < 16; k++) r[k] = regs[(((int64_t)u * 2 + k / 8) * n + i) * 8 + k % 8]; /* two 128-bit cells per sample */
      if(do_init && u == 0) { /* IMU_IF_WT901C::init  :63-77 (getDataImmediately -> updateData, then latch) */
        imu_update_data(qi, r, d);
        for(k = 0; k < 4; k++) qi[k] = r[RK_IMT_REG_Q0 + k] / 32768.0f;
      } else if(hq) { /* ::update  :83-89 */
        flags &= ~RK_IS_FLAG_ERROR;
        imu_update_data(qi, r, d);
      } else {
        flags |= RK_IS_FLAG_ERROR;
      }
      if(out)
        for(k = 0; k < 16; k++) out[(((int64_t)u * 4 + k / 4) * n + i) * 4 + (k % 4)] = f2u(d[k]);
    }
    if(state) {
      for(k = 0; k < 4; k++) *soa(state, n, i, RK_IS_QINIT + k) = f2u(qi[k]);
      for(k = 0; k < 16; k++) *soa(state, n, i, RK_IS_DATA + k) = f2u(d[k]);
      *soa(state, n, i, RK_IS_FLAGS) = flags;
      for(k = RK_IS_FLAGS + 1; k < RK_IS_WORDS; k++) *soa(state, n, i, k) = 0;
    }
  }
}

/* ------------------------------------------------------------------------------------ */
/* Arm: ADTModePositioningSeq + joint command packers (src/ArmDrive), on AoS state words   */
/* (int32_t)(float) as the x86 build performs it: cvttss2si, out of range / NaN -> INT_MIN   */
static int32_t f2i_x86(float f) {
  if(!(fabsf(f) < 2147483648.0f)) return (int32_t)0x80000000u;
  return (int32_t)f;
}
/* IcsBaseClass::degPos100 / posDeg100   lib/IcsClass_V210/src/IcsBaseClass.cpp:105-137 */
static int ics_degPos100(int deg) {
  long long a;
  if(deg > 18000 || deg < -18000) return -1;
  a = ((long long)deg * 2963) / 10000;
  return (int)a + 7500;
}
static int ics_posDeg100(int pos) {
  long long a   = (long long)pos - 7500;
  int       deg = (int)((a * 1000) / 296);
  if(deg > 18000) return 0x7FFF;
  if(deg < -18000) return -0x7FFF;
  return deg;
}
#define AJ(w, k, f) ((w)[RK_AS_JOINT0 + 4 * (k) + (f)])
static float adt_get_tgt_deg(const uint32_t *w, int k) { /* JointBase::get_tgt_deg  AD_joint_base.hpp:47 */
  return u2f(AJ(w, k, RK_AJ_RAW_TGT)) - u2f(AJ(w, k, RK_AJ_OFS));
}
static const int ADT_AXIS[5] = {RK_AJ_Y0, RK_AJ_P1, RK_AJ_P2, RK_AJ_R0, RK_AJ_P3}; /* AD_task_main.cpp:148 */

/* JointBase::set_tgt_ang_deg (:42) and the DfGear overrides (AD_joint_dfgear.hpp:14-37,60-63,93-96) */
static void adt_set_tgt(const rk_adt_params_t *p, uint32_t *w, int axis, float tgt) {
  int   k   = ADT_AXIS[axis];
  float raw = tgt + u2f(AJ(w, k, RK_AJ_OFS));
  AJ(w, k, RK_AJ_RAW_TGT) = f2u(raw);
  if(k == RK_AJ_P2 || k == RK_AJ_R0) {
    float P, R;
    if(k == RK_AJ_P2) w[RK_AS_DFV_P] = f2u(raw * p->gear_ratio[k]);
    else w[RK_AS_DFV_R] = f2u(raw * p->gear_ratio[k]);
    P = u2f(w[RK_AS_DFV_P]), R = u2f(w[RK_AS_DFV_R]);
    AJ(w, RK_AJ_DFL, RK_AJ_RAW_TGT) = f2u((P - R) + u2f(AJ(w, RK_AJ_DFL, RK_AJ_OFS)));
    AJ(w, RK_AJ_DFR, RK_AJ_RAW_TGT) = f2u(-(P + R) + u2f(AJ(w, RK_AJ_DFR, RK_AJ_OFS)));
  }
}

/* ADTModePositioningSeq::update   AD_mode_positioning_seq.cpp:13-117 */
static void adt_mode_update(const rk_adt_params_t *p, uint32_t *w, const uint32_t *tab, int64_t n, int64_t i) {
  uint32_t fsm   = w[RK_AS_FSM];
  uint32_t state = fsm & 0xFFu;
  uint32_t exec = w[RK_AS_SEQ_IDX] & 0xFFFFu, head = w[RK_AS_SEQ_IDX] >> 16;
  int      j;
  if(state == RK_ASTATE_STANDBY) { /* exec_standby :24-42 */
    fsm |= RK_AS_FSM_IS_COMP;
    if(exec != head) {
      exec = (exec + 1) & 0xFFFFu;
      exec = (exec >= RK_ACMD_SLOTS) ? 0 : exec;
      w[RK_AS_CMD_IDX]  = 0;
      w[RK_AS_TOTAL_MS] = 0;
      state             = RK_ASTATE_MOVE_START;
      fsm &= ~RK_AS_FSM_FIRSTCALL;
    }
  }
  if(state == RK_ASTATE_MOVE_START) { /* exec_move_start :48-83 */
    /* cmd_seq_[exec] with exec > 3 is out of bounds in the reference; slots wrap here */
    int      base = (int)(exec % RK_ACMD_SLOTS) * RK_ACMD_SLOT_WORDS;
    uint32_t len  = *soa((uint32_t *)tab, n, i, base + 1) & 0xFFu;
    uint32_t idx  = w[RK_AS_CMD_IDX] & 0xFFu;
    if(idx >= len) {
      state = RK_ASTATE_STANDBY;
    } else {
      int32_t cnt;
      float   fc;
      int     wb = base + 4 + 8 * (int)(idx % RK_ACMD_MAX_LEN);
      w[RK_AS_NOW_DT] = *soa((uint32_t *)tab, n, i, wb);
      for(j = 0; j < 5; j++) w[RK_AS_NOW_TGT + j] = *soa((uint32_t *)tab, n, i, wb + 1 + j);
      cnt = f2i_x86((float)(uint32_t)(w[RK_AS_NOW_DT] - w[RK_AS_TOTAL_MS]) * 0.001f / p->cycle_time_s);
      cnt = (cnt <= 0) ? 1 : cnt;
      fc  = (float)cnt;
      for(j = 0; j < 5; j++) w[RK_AS_MOVE_DEG + j] = f2u((u2f(w[RK_AS_NOW_TGT + j]) - adt_get_tgt_deg(w, ADT_AXIS[j])) / fc);
      w[RK_AS_MOVE_CNT] = (uint32_t)cnt;
      w[RK_AS_TOTAL_MS] = w[RK_AS_NOW_DT];
      w[RK_AS_CYCLE]    = 0;
      fsm &= ~RK_AS_FSM_IS_COMP;
      state = RK_ASTATE_MOVING;
    }
  }
  if(state == RK_ASTATE_MOVING) { /* exec_moving :89-117 */
    int32_t cnt = (int32_t)w[RK_AS_MOVE_CNT], cyc = (int32_t)w[RK_AS_CYCLE];
    float   rem = (float)(cnt - cyc);
    for(j = 0; j < 5; j++) adt_set_tgt(p, w, j, u2f(w[RK_AS_NOW_TGT + j]) - u2f(w[RK_AS_MOVE_DEG + j]) * rem);
    if(cnt <= cyc) {
      w[RK_AS_CMD_IDX] = (w[RK_AS_CMD_IDX] + 1) & 0xFFu;
      state            = RK_ASTATE_MOVE_START;
    } else {
      w[RK_AS_CYCLE] = (uint32_t)(cyc + 1);
    }
  }
  w[RK_AS_FSM]     = (fsm & ~0xFFu) | state;
  w[RK_AS_SEQ_IDX] = exec | (head << 16);
}

static uint32_t jflag(const uint32_t *w, int k) { return (w[RK_AS_JFLAGS] >> (4 * k)) & 0xFu; }
static void     set_jflag(uint32_t *w, int k, uint32_t b) { w[RK_AS_JFLAGS] = (w[RK_AS_JFLAGS] & ~(0xFu << (4 * k))) | (b << (4 * k)); }

/* JointMgServo::subproc_torquectrl   AD_joint_mg_servo.cpp:104-134 with UTIL::PI_D pos_ctrl_
 * (util_controller.hpp:86-153; constructed as PI_D(1/ctrl_time, 0, 0, 0, 0, 10), AD_joint_mg_servo.hpp:69)
 * and the double-precision current -> raw map (AD_joint_mg_servo.hpp:120-136). */
enum { MGC_PREV_VAL = 0, MGC_INTEG, MGC_LPF_Y, MGC_LPF_X, MGC_NOW_TGT, MGC_NOW_ERR, MGC_NOW_CTRL, MGC_GAINSET };
static int32_t d2i_x86(double d) { /* cvttsd2si */
  if(!(fabs(d) < 2147483648.0)) return (int32_t)0x80000000u;
  return (int32_t)d;
}
static void adt_mg_torquectrl(const rk_adt_params_t *p, uint32_t *w, int ini) {
  uint32_t *c    = w + RK_AS_MG_CTRL;
  float     freq = 1.0f / p->ctrl_time_s[RK_AJ_P1], dt = 1.0f / freq, lpf = 10.0f;
  float     A1 = (2.0f * freq - lpf) / (2.0f * freq + lpf), B0 = lpf / (2.0f * freq + lpf), B1 = B0;
  float     pg = c[MGC_GAINSET] ? 0.01f : 0.0f, ig = 0.0f, dg = 0.0f, ilim = 0.0f; /* InitGain  .cpp:25-31 */
  float     tgt = u2f(AJ(w, RK_AJ_P1, RK_AJ_RAW_TGT)), now = u2f(AJ(w, RK_AJ_P1, RK_AJ_RAW_NOW));
  float     curlim = u2f(AJ(w, RK_AJ_P1, RK_AJ_CURLIM));
  float     err, x, y, integ, iq;
  double    raw;
  const double C_A = 0.0000057204, C_B = -0.0000485371;
  int32_t   s;
  err   = tgt - now;
  x     = (now - u2f(c[MGC_PREV_VAL])) * freq;
  y     = A1 * u2f(c[MGC_LPF_Y]) + B0 * x + B1 * u2f(c[MGC_LPF_X]);
  integ = u2f(c[MGC_INTEG]);
  integ += ig * dt * err;
  integ = (integ >= ilim) ? ilim : ((integ <= -ilim) ? -ilim : integ);
  iq    = pg * err + integ - dg * y;
  c[MGC_PREV_VAL] = f2u(now), c[MGC_INTEG] = f2u(integ), c[MGC_LPF_Y] = f2u(y), c[MGC_LPF_X] = f2u(x);
  c[MGC_NOW_TGT] = f2u(tgt), c[MGC_NOW_ERR] = f2u(err), c[MGC_NOW_CTRL] = f2u(iq);
  if(ini) iq -= 0.05f * orc_sin((now - u2f(AJ(w, RK_AJ_P1, RK_AJ_OFS))) * (ORC_PI / 180.0f));
  iq = (iq > curlim) ? curlim : ((iq < -curlim) ? -curlim : iq);
  if((double)iq >= 0) raw = (-C_B + orc_sqrt((float)(C_B * C_B + 4.0 * C_A * (double)iq))) / (2.0 * C_A);
  else raw = (C_B - orc_sqrt((float)(C_B * C_B - 4.0 * C_A * (double)iq))) / (2.0 * C_A);
  s = (int32_t)(int16_t)d2i_x86(-1.0f * raw);
  s = (s > 450) ? 450 : ((s < -450) ? -450 : s);
  w[RK_AS_MG_TX]     = 0xA1u;
  w[RK_AS_MG_TX + 1] = (uint32_t)s & 0xFFFFu;
  w[RK_AS_MG_TX + 2] = 1;
}

/* JointMgServo::update   AD_joint_mg_servo.cpp:50-73 */
static void adt_mg_update(const rk_adt_params_t *p, uint32_t *w) {
  uint32_t b    = jflag(w, RK_AJ_P1);
  int      on   = (b & RK_AJF_TORQUE_ON) != 0, prev = (b & RK_AJF_TORQUE_PREV) != 0, ini = (b & RK_AJF_INITIALIZED) != 0;
  float    tgt  = u2f(AJ(w, RK_AJ_P1, RK_AJ_RAW_TGT));
  int      k;
  w[RK_AS_MG_TX + 2] = 0;
  if(prev && !on) { /* pos_ctrl_.reset()  util_controller.hpp:112-124: everything but the gains */
    for(k = MGC_PREV_VAL; k <= MGC_NOW_CTRL; k++) w[RK_AS_MG_CTRL + k] = 0;
  } else if(!ini && on) {
    adt_mg_torquectrl(p, w, ini);
  } else if(on) { /* subproc_posctrl :136-149 */
    float    v   = fabsf((tgt - u2f(w[RK_AS_MG_PRE_TGT])) / p->ctrl_time_s[RK_AJ_P1] * -10.0f);
    uint32_t vl  = (uint32_t)f2i_x86((v > 1800) ? 1800 : v) & 0xFFFFu;
    int32_t  ang = f2i_x86(tgt * (-100.0f * 10.0f));
    w[RK_AS_MG_TX]     = 0xA4u | (vl << 16);
    w[RK_AS_MG_TX + 1] = (uint32_t)ang;
    w[RK_AS_MG_TX + 2] = 1;
  } else { /* torque off: set_myctrl_gain_params(InitGain) -- set_VelLpf_CutOff resets the IIR -- then torque control */
    w[RK_AS_MG_CTRL + MGC_GAINSET] = 1;
    w[RK_AS_MG_CTRL + MGC_LPF_Y] = 0, w[RK_AS_MG_CTRL + MGC_LPF_X] = 0;
    adt_mg_torquectrl(p, w, ini);
  }
  set_jflag(w, RK_AJ_P1, (b & ~RK_AJF_TORQUE_PREV) | (on ? RK_AJF_TORQUE_PREV : 0u));
  w[RK_AS_MG_PRE_TGT] = f2u(tgt);
}

/* JointMyBldcServo::update   AD_joint_mybldc_servo.cpp:7-36 ; slot 0..2 = DF_Left, DF_Right, P3 */
static void adt_bldc_update(const rk_adt_params_t *p, uint32_t *w, int slot) {
  static const int JK[3] = {RK_AJ_DFL, RK_AJ_DFR, RK_AJ_P3};
  int       k  = JK[slot];
  uint32_t  b  = jflag(w, k);
  int       on = (b & RK_AJF_TORQUE_ON) != 0, prev = (b & RK_AJF_TORQUE_PREV) != 0;
  uint32_t *q  = w + RK_AS_BLDC_TX0 + 4 * slot;
  if(!on) {
    q[0] = 0, q[1] = 0, q[2] = 0x8002u;
  } else if(!prev) {
    q[0] = 0, q[1] = 0, q[2] = 0x8001u;
  } else {
    int32_t  a  = f2i_x86(u2f(AJ(w, k, RK_AJ_RAW_TGT)) * p->gear_ratio[k] * p->motor_dir[k] * 65536.0f);
    uint32_t ms = (uint32_t)f2i_x86(p->ctrl_time_s[k] * 1000.0f) & 0xFFFFu;
    uint32_t cl = (uint32_t)f2i_x86(u2f(AJ(w, k, RK_AJ_CURLIM)) * 256.0f) & 0xFFFFu;
    q[0] = (uint32_t)a, q[1] = ms | (cl << 16), q[2] = 0x8010u;
  }
  q[3] = 1;
  set_jflag(w, k, (b & ~RK_AJF_TORQUE_PREV) | (on ? RK_AJF_TORQUE_PREV : 0u));
}

/* JointIcsServo::update   AD_joint_ics_servo.cpp:5-29 over the ideal servo of
 * oracle/stubs/IcsHardSerialClass.h (echoes the commanded position) */
static void adt_ics_update(const rk_adt_params_t *p, uint32_t *w) {
  uint32_t b = jflag(w, RK_AJ_Y0);
  int      tgt_pos, now_pos;
  if(!(b & RK_AJF_CONNECTED)) return;
  tgt_pos = ics_degPos100(f2i_x86(u2f(AJ(w, RK_AJ_Y0, RK_AJ_RAW_TGT)) * p->motor_dir[RK_AJ_Y0] * 100.0f));
  if(tgt_pos == -1) return;
  if(b & RK_AJF_TORQUE_ON) {
    if(tgt_pos > 11500 || tgt_pos < 3500) { /* IcsBaseClass::setPos range check: nothing transmitted */
      now_pos = -1;
    } else {
      w[RK_AS_ICS_POS]   = (uint32_t)tgt_pos;
      w[RK_AS_ICS_SERVO] = (uint32_t)(tgt_pos - 7500);
      now_pos            = tgt_pos;
    }
  } else { /* setFree: {0x80 + id, 0, 0} */
    w[RK_AS_ICS_POS] = (uint32_t)-1;
    now_pos          = (int)(int32_t)w[RK_AS_ICS_SERVO] + 7500;
  }
  AJ(w, RK_AJ_Y0, RK_AJ_RAW_NOW) = f2u((float)ics_posDeg100(now_pos) * 0.01f * p->motor_dir[RK_AJ_Y0]);
}

/* one ADT::main loop body   AD_task_main.cpp:208-229 */
static void adt_tick(const rk_adt_params_t *p, uint32_t *w, const uint32_t *tab, int64_t n, int64_t i) {
  adt_mode_update(p, w, tab, n, i);
  adt_mg_update(p, w);
  adt_bldc_update(p, w, 0);
  adt_bldc_update(p, w, 1);
  adt_bldc_update(p, w, 2);
  adt_ics_update(p, w);
}

/* prepare_task + finished INIT mode + ADTModeBase::init (see oracle/ref_harness_arm.cpp bringup()) */
static void adt_bringup(const rk_adt_params_t *p, uint32_t *w) {
  float now = (float)ics_posDeg100((int)(int32_t)w[RK_AS_ICS_SERVO] + 7500) * 0.01f * p->motor_dir[RK_AJ_Y0];
  AJ(w, RK_AJ_Y0, RK_AJ_RAW_NOW) = f2u(now);
  AJ(w, RK_AJ_Y0, RK_AJ_RAW_TGT) = f2u(now);
  w[RK_AS_ICS_POS]               = (uint32_t)-1;
  set_jflag(w, RK_AJ_Y0, RK_AJF_CONNECTED | RK_AJF_TORQUE_ON | RK_AJF_INITIALIZED);
  set_jflag(w, RK_AJ_P1, RK_AJF_CONNECTED | RK_AJF_TORQUE_ON | RK_AJF_INITIALIZED);
  set_jflag(w, RK_AJ_DFL, jflag(w, RK_AJ_DFL) | RK_AJF_TORQUE_ON);
  set_jflag(w, RK_AJ_DFR, jflag(w, RK_AJ_DFR) | RK_AJF_TORQUE_ON);
  set_jflag(w, RK_AJ_P2, jflag(w, RK_AJ_P2) | RK_AJF_INITIALIZED);
  set_jflag(w, RK_AJ_R0, jflag(w, RK_AJ_R0) | RK_AJF_INITIALIZED);
  set_jflag(w, RK_AJ_P3, jflag(w, RK_AJ_P3) | RK_AJF_TORQUE_ON | RK_AJF_INITIALIZED);
  AJ(w, RK_AJ_Y0, RK_AJ_CURLIM)  = f2u(p->curlim_default_A[RK_AJ_Y0]);
  AJ(w, RK_AJ_P1, RK_AJ_CURLIM)  = f2u(p->curlim_default_A[RK_AJ_P1]);
  AJ(w, RK_AJ_DFL, RK_AJ_CURLIM) = f2u(p->curlim_default_A[RK_AJ_R0]);
  AJ(w, RK_AJ_DFR, RK_AJ_CURLIM) = f2u(p->curlim_default_A[RK_AJ_R0]);
  AJ(w, RK_AJ_P3, RK_AJ_CURLIM)  = f2u(p->curlim_default_A[RK_AJ_P3]);
  w[RK_AS_MG_CTRL + 7] = 1; /* JointMgServo::init(): set_myctrl_gain_params(InitGain), IIR reset */
  w[RK_AS_MG_CTRL + 2] = 0, w[RK_AS_MG_CTRL + 3] = 0;
  w[RK_AS_FSM]     = RK_ASTATE_STANDBY | RK_AS_FSM_FIRSTCALL;
  w[RK_AS_SEQ_IDX] = (RK_ACMD_SLOTS - 1) | ((uint32_t)(RK_ACMD_SLOTS - 1) << 16);
}

/* ADTModePositioningSeq::push_cmdseq   AD_mode_positioning_seq.cpp:124-137 */
static void adt_push(uint32_t *w, uint32_t *tab, int64_t n, int64_t i, const uint32_t *seq) {
  uint32_t exec = w[RK_AS_SEQ_IDX] & 0xFFFFu, head = w[RK_AS_SEQ_IDX] >> 16;
  uint32_t nw   = (head + 1) & 0xFFFFu;
  int      k;
  nw = (nw >= RK_ACMD_SLOTS) ? 0 : nw;
  if(nw == exec) return;
  for(k = 0; k < RK_ACMD_SLOT_WORDS; k++) *soa(tab, n, i, (int)nw * RK_ACMD_SLOT_WORDS + k) = *soa((uint32_t *)seq, n, i, k);
  *soa(tab, n, i, (int)nw * RK_ACMD_SLOT_WORDS + 1) &= 0xFFu;
  w[RK_AS_SEQ_IDX] = exec | (nw << 16);
}

/* ADTModePositioningSeq::get_q_cmdseq_status   AD_mode_positioning_seq.cpp:146-184 */
static int32_t adt_status(const uint32_t *w, const uint32_t *tab, int64_t n, int64_t i, uint32_t id) {
  uint32_t exec = w[RK_AS_SEQ_IDX] & 0xFFFFu, head = w[RK_AS_SEQ_IDX] >> 16;
  int32_t  sts  = 99;
  uint32_t s;
  if((w[RK_AS_FSM] & RK_AS_FSM_FIRSTCALL) && id == 0) return 99;
  for(s = 0; s < RK_ACMD_SLOTS; s++) {
    if(*soa((uint32_t *)tab, n, i, (int)s * RK_ACMD_SLOT_WORDS) != id) continue;
    if(exec == head) {
      uint32_t len = *soa((uint32_t *)tab, n, i, (int)(exec % RK_ACMD_SLOTS) * RK_ACMD_SLOT_WORDS + 1) & 0xFFu;
      sts          = ((w[RK_AS_CMD_IDX] & 0xFFu) >= len) ? 1 : 0;
    } else if(exec < head) {
      sts = (exec <= s && s <= head) ? 0 : 1;
    } else {
      sts = ((exec <= s && s < RK_ACMD_SLOTS) || s <= head) ? 0 : 1;
    }
  }
  return sts;
}

static uint32_t bldc_id_byte(uint32_t id) { return (id & 0xFFu) | ((id & 0x8000u) ? 0x80u : 0u); }

void orc_adt_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *cmdtab, int64_t n, int64_t i0, int64_t i1,
                   int K, const uint32_t *seq, const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status) {
  int64_t i;
  int     k, t, j;
  for(i = i0; i < i1; i++) {
    uint32_t w[RK_AS_WORDS];
    for(k = 0; k < RK_AS_WORDS; k++) w[k] = *soa(state, n, i, k);
    if(op == 0) {
      adt_bringup(p, w);
    } else if(op == 1) {
      if(!valid || valid[i]) adt_push(w, cmdtab, n, i, seq);
    } else if(op == 2) {
      for(t = 0; t < K; t++) {
        adt_tick(p, w, cmdtab, n, i);
        if(trace) {
          uint32_t *tr = trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i;
          for(j = 0; j < 5; j++) tr[(int64_t)j * n] = f2u(adt_get_tgt_deg(w, ADT_AXIS[j]));
          tr[5 * n] = w[RK_AS_MG_TX] >> 16, tr[6 * n] = w[RK_AS_MG_TX + 1];
          for(j = 0; j < 3; j++) tr[(int64_t)(7 + j) * n] = w[RK_AS_BLDC_TX0 + 4 * j];
          tr[10 * n] = w[RK_AS_ICS_POS];
          tr[11 * n] = w[RK_AS_FSM] & 0xFFu;
          tr[12 * n] = w[RK_AS_CMD_IDX];
          tr[13 * n] = bldc_id_byte(w[RK_AS_BLDC_TX0 + 2]) | (bldc_id_byte(w[RK_AS_BLDC_TX0 + 6]) << 8) |
                       (bldc_id_byte(w[RK_AS_BLDC_TX0 + 10]) << 16);
          tr[14 * n] = 0, tr[15 * n] = 0;
        }
      }
    } else if(op == 3) {
      status[i] = adt_status(w, cmdtab, n, i, ids[i]);
    }
    if(op != 3)
      for(k = 0; k < RK_AS_WORDS; k++) *soa(state, n, i, k) = w[k];
  }
}

/* ------------------------------------------------------------------------------------ */
/* Servo feedback (arm side): the CAN rx callbacks of the arm's servos       */
static int16_t le16(const uint8_t *b) { return (int16_t)((uint16_t)b[0] | ((uint16_t)b[1] << 8)); }
/* JointMyBldcServo::rx_callback -> rx_summary_status  AD_joint_mybldc_servo.cpp:45-70, frame layout
 * RES_STATUS_SUMMAY AD_joint_mybldc_servo.hpp:49-60: [0] flags, [1] mode, [2..3] s16_out_ang_deg_Q4, [4] s8_motor_curr_A_Q4 */
static void adt_bldc_rx(const rk_adt_params_t *p, uint32_t *w, int slot, uint32_t cmdid, const uint8_t f[8], float *cur) {
  static const int JK[3] = {RK_AJ_DFL, RK_AJ_DFR, RK_AJ_P3};
  const int        k    = JK[slot];
  float            now;
  if(cmdid != 0x1000u) return; /* CMD_ID_RES_STATUS_SUMMARY; every other id is ignored (:49-57) */
  now = (float)le16(f + 2) / 16.0f / p->gear_ratio[k] * p->motor_dir[k];
  AJ(w, k, RK_AJ_RAW_NOW) = f2u(now);
  if(cur) *cur = (float)(int8_t)f[4] / 16.0f * p->motor_dir[k];
  if(!((w[RK_AS_JFLAGS] >> (4 * k)) & RK_AJF_TORQUE_ON)) AJ(w, k, RK_AJ_RAW_TGT) = f2u(now); /* :69 */
}
/* JointMgServo::rx_callback  AD_joint_mg_servo.cpp:75-92; conv_raw_to_current AD_joint_mg_servo.hpp:120-128.
 * 0x92 (multi-turn angle): the firmware ORs `u8_ang[i] << (i * 8)` for i = 0..6 with the byte promoted to a 32-bit int,
 * so shifts of 32 and more are undefined in C++.  On the Cortex-M7 a register shift by >= 32 yields 0 and the i = 3
 * term sign-extends into the 64-bit accumulator: the angle is the sign-extended low 32 bits.  That is what is
 * restated here; the x86 build of the reference agrees whenever bytes 5..7 of the frame are zero (x86 masks the
 * shift count to 5 bits instead), which is where tests pin it. */
static void adt_mg_rx(uint32_t *w, const uint8_t f[8], float *cur) {
  if(f[0] == 0x92) {
    const int32_t lo   = (int32_t)((uint32_t)f[1] | ((uint32_t)f[2] << 8) | ((uint32_t)f[3] << 16) | ((uint32_t)f[4] << 24));
    const int64_t ang  = (int64_t)((uint64_t)(int64_t)lo << 8);
    const double  DB   = -1.0f / 100.0f / 10.0f / 256.0f; /* DB_ANG_RAW_TO_DEG :17: float arithmetic, then widened */
    const float   now  = (float)((double)ang * DB);
    AJ(w, RK_AJ_P1, RK_AJ_RAW_NOW) = f2u(now);
    if(!((w[RK_AS_JFLAGS] >> (4 * RK_AJ_P1)) & RK_AJF_TORQUE_ON)) AJ(w, RK_AJ_P1, RK_AJ_RAW_TGT) = f2u(now);
  } else if(f[0] == 0x9C || f[0] == 0xA1) {
    const double C_A = 0.0000057204, C_B = -0.0000485371, raw = (double)le16(f + 2);
    double       c;
    if(raw >= 0) c = C_A * raw * raw + C_B * raw;
    else c = -(C_A * raw * raw - C_B * raw);
    if(cur) *cur = -1.0f * (float)c; /* FL_CURR_DIR */
  }
}
void orc_adt_rx_batch(int kind, const rk_adt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1, const uint64_t *frames,
                      const uint32_t *cmdid, float *cur) {
  int64_t i;
  int     k;
  for(i = i0; i < i1; i++) {
    uint32_t w[RK_AS_WORDS];
    uint8_t  f[8];
    for(k = 0; k < RK_AS_WORDS; k++) w[k] = *soa(state, n, i, k);
    memcpy(f, &frames[i], 8);
    if(kind == 3) adt_mg_rx(w, f, cur ? &cur[i] : NULL);
    else adt_bldc_rx(p, w, kind, cmdid ? cmdid[i] : 0x1000u, f, cur ? &cur[i] : NULL);
    for(k = 0; k < RK_AS_WORDS; k++) *soa(state, n, i, k) = w[k];
  }
}

/* ------------------------------------------------------------------------------------ */
/* ADTModePositioning   src/ArmDrive/AD_mode_positioning.cpp (single-command FIFO mode)    */
static float adt_get_now_deg(const rk_adt_params_t *p, const uint32_t *w, int axis) {
  /* JointBase::get_now_deg (AD_joint_base.hpp:48); DfGear overrides (AD_joint_dfgear.hpp:76-77,98) */
  float ln = u2f(AJ(w, RK_AJ_DFL, RK_AJ_RAW_NOW)) - u2f(AJ(w, RK_AJ_DFL, RK_AJ_OFS));
  float rn = u2f(AJ(w, RK_AJ_DFR, RK_AJ_RAW_NOW)) - u2f(AJ(w, RK_AJ_DFR, RK_AJ_OFS));
  int   k  = ADT_AXIS[axis];
  if(k == RK_AJ_P2) return (ln - rn) * 0.5f / p->gear_ratio[k] - u2f(AJ(w, k, RK_AJ_OFS));
  if(k == RK_AJ_R0) return -(ln + rn) * 0.5f / p->gear_ratio[k] - u2f(AJ(w, k, RK_AJ_OFS));
  return u2f(AJ(w, k, RK_AJ_RAW_NOW)) - u2f(AJ(w, k, RK_AJ_OFS));
}

static void adp_mode_update(const rk_adt_params_t *p, uint32_t *w, uint32_t *pw) {
  uint32_t state = pw[RK_PS_STATE] & 0xFFu, flags = pw[RK_PS_STATE] & ~0xFFu;
  int      j, e;
  if(state == 0) { /* exec_standby :27-58 */
    flags |= RK_AS_FSM_IS_COMP;
    if(pw[RK_PS_QSIZE] > 0) {
      uint32_t cnt, qs = pw[RK_PS_QSIZE] > 4 ? 4 : pw[RK_PS_QSIZE];
      float    fc;
      for(j = 0; j < 8; j++) pw[RK_PS_NOW_CMD + j] = pw[RK_PS_QUEUE + j];
      for(e = 1; e < (int)qs; e++)
        for(j = 0; j < 8; j++) pw[RK_PS_QUEUE + 8 * (e - 1) + j] = pw[RK_PS_QUEUE + 8 * e + j];
      for(j = 0; j < 8; j++) pw[RK_PS_QUEUE + 8 * (qs - 1) + j] = 0; /* entries past size() are kept zero */
      pw[RK_PS_QSIZE] = qs - 1;
      cnt = (uint32_t)f2i_x86((float)pw[RK_PS_NOW_CMD + 1] * 0.001f / p->cycle_time_s);
      cnt = (cnt == 0) ? 1 : cnt;
      fc  = (float)cnt;
      for(j = 0; j < 5; j++) pw[RK_PS_MOVE_DEG + j] = f2u((u2f(pw[RK_PS_NOW_CMD + 2 + j]) - adt_get_now_deg(p, w, j)) / fc);
      pw[RK_PS_MOVE_CNT] = cnt;
      pw[RK_PS_CYCLE]    = 0;
      flags &= ~RK_AS_FSM_IS_COMP;
      state = 1;
    }
  } else if(state == 1) { /* exec_moving :64-110 */
    uint32_t cnt = pw[RK_PS_MOVE_CNT], cyc = pw[RK_PS_CYCLE];
    float    rem = (float)(uint32_t)(cnt - cyc);
    for(j = 0; j < 5; j++) adt_set_tgt(p, w, j, u2f(pw[RK_PS_NOW_CMD + 2 + j]) - u2f(pw[RK_PS_MOVE_DEG + j]) * rem);
    if(cnt <= cyc) {
      pw[RK_PS_PREV_ID1] = pw[RK_PS_PREV_ID0];
      pw[RK_PS_PREV_ID0] = pw[RK_PS_NOW_CMD];
      state              = 0;
    } else {
      pw[RK_PS_CYCLE] = cyc + 1;
    }
  }
  pw[RK_PS_STATE] = flags | state;
}

void orc_adp_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *pstate, int64_t n, int64_t i0, int64_t i1, int K,
                   const uint32_t *cmd, const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status) {
  int64_t i;
  int     k, t, j, e;
  for(i = i0; i < i1; i++) {
    uint32_t w[RK_AS_WORDS], pw[RK_PS_WORDS];
    for(k = 0; k < RK_AS_WORDS; k++) w[k] = *soa(state, n, i, k);
    for(k = 0; k < RK_PS_WORDS; k++) pw[k] = *soa(pstate, n, i, k);
    if(op == 0) { /* ADTModeBase::init -> doInit */
      pw[RK_PS_STATE] = 0;
    } else if(op == 1) { /* push_cmd :118-124 */
      if(!valid || valid[i]) {
        uint32_t qs = pw[RK_PS_QSIZE] > 4 ? 4 : pw[RK_PS_QSIZE];
        if(qs >= 4) {
          for(e = 1; e < 4; e++)
            for(j = 0; j < 8; j++) pw[RK_PS_QUEUE + 8 * (e - 1) + j] = pw[RK_PS_QUEUE + 8 * e + j];
          qs = 3;
        }
        for(j = 0; j < 7; j++) pw[RK_PS_QUEUE + 8 * qs + j] = *soa((uint32_t *)cmd, n, i, j);
        pw[RK_PS_QUEUE + 8 * qs + 7] = 0;
        pw[RK_PS_QSIZE]              = qs + 1;
      }
    } else if(op == 2) {
      for(t = 0; t < K; t++) {
        adp_mode_update(p, w, pw);
        adt_mg_update(p, w);
        adt_bldc_update(p, w, 0);
        adt_bldc_update(p, w, 1);
        adt_bldc_update(p, w, 2);
        adt_ics_update(p, w);
        if(trace) {
          uint32_t *tr = trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i;
          for(j = 0; j < 5; j++) tr[(int64_t)j * n] = f2u(adt_get_tgt_deg(w, ADT_AXIS[j]));
          tr[5 * n] = w[RK_AS_MG_TX] >> 16, tr[6 * n] = w[RK_AS_MG_TX + 1];
          for(j = 0; j < 3; j++) tr[(int64_t)(7 + j) * n] = w[RK_AS_BLDC_TX0 + 4 * j];
          tr[10 * n] = w[RK_AS_ICS_POS];
          tr[11 * n] = pw[RK_PS_STATE] & 0xFFu;
          tr[12 * n] = pw[RK_PS_QSIZE];
          tr[13 * n] = bldc_id_byte(w[RK_AS_BLDC_TX0 + 2]) | (bldc_id_byte(w[RK_AS_BLDC_TX0 + 6]) << 8) |
                       (bldc_id_byte(w[RK_AS_BLDC_TX0 + 10]) << 16);
          tr[14 * n] = 0, tr[15 * n] = 0;
        }
      }
    } else if(op == 3) { /* get_q_cmd_status :134-148 */
      int32_t  sts = 0x63;
      uint32_t qs  = pw[RK_PS_QSIZE] > 4 ? 4 : pw[RK_PS_QSIZE];
      for(e = 0; e < (int)qs; e++)
        if(pw[RK_PS_QUEUE + 8 * e] == ids[i]) sts = 0;
      if(pw[RK_PS_PREV_ID0] == ids[i] || pw[RK_PS_PREV_ID1] == ids[i]) sts = 1;
      status[i] = sts;
    }
    if(op != 3) {
      if(op == 2)
        for(k = 0; k < RK_AS_WORDS; k++) *soa(state, n, i, k) = w[k];
      for(k = 0; k < RK_PS_WORDS; k++) *soa(pstate, n, i, k) = pw[k];
    }
  }
}


/* ------------------------------------------------------------------------------------ */
/* Homing modes: ADTModeInitialize (src/ArmDrive/AD_mode_initialize.cpp) and              */
/* ADTModeInitPosMove (src/ArmDrive/AD_mode_initpos_move.cpp) on the same joints           */
static float adt_absf(float x) { return (x < 0) ? -x : x; } /* mymath::absf  util_mymath.hpp:40 */
/* set_torque_on: JointBase :39; DfGearPitch/Roll forward to both motors (AD_joint_dfgear.hpp:50-53) */
static void adh_set_torque_on(uint32_t *w, int axis, int on) {
  int k = ADT_AXIS[axis];
  if(k == RK_AJ_P2 || k == RK_AJ_R0) {
    set_jflag(w, RK_AJ_DFL, (jflag(w, RK_AJ_DFL) & ~RK_AJF_TORQUE_ON) | (on ? RK_AJF_TORQUE_ON : 0u));
    set_jflag(w, RK_AJ_DFR, (jflag(w, RK_AJ_DFR) & ~RK_AJF_TORQUE_ON) | (on ? RK_AJF_TORQUE_ON : 0u));
  } else {
    set_jflag(w, k, (jflag(w, k) & ~RK_AJF_TORQUE_ON) | (on ? RK_AJF_TORQUE_ON : 0u));
  }
}
static void adh_set_initialized(uint32_t *w, int axis, int on) { /* JointBase :40 (not overridden) */
  int k = ADT_AXIS[axis];
  set_jflag(w, k, (jflag(w, k) & ~RK_AJF_INITIALIZED) | (on ? RK_AJF_INITIALIZED : 0u));
}
static void adh_set_curlim(uint32_t *w, int axis, float lim) { /* JointBase :43; DfGear forwards (:55-58) */
  int k = ADT_AXIS[axis];
  if(k == RK_AJ_P2 || k == RK_AJ_R0) AJ(w, RK_AJ_DFL, RK_AJ_CURLIM) = f2u(lim), AJ(w, RK_AJ_DFR, RK_AJ_CURLIM) = f2u(lim);
  else AJ(w, k, RK_AJ_CURLIM) = f2u(lim);
}
static void adh_mech_reset_pos(const rk_adt_params_t *p, uint32_t *w, int axis) { /* JointBase :36-38; DfGear :65-71,100-106 */
  int k = ADT_AXIS[axis];
  if(k == RK_AJ_P2) { /* the pitch joint resets both motors, the roll joint does not */
    AJ(w, RK_AJ_DFL, RK_AJ_OFS) = f2u(u2f(AJ(w, RK_AJ_DFL, RK_AJ_RAW_NOW)) - p->mechend_pos_deg[RK_AJ_DFL]);
    AJ(w, RK_AJ_DFR, RK_AJ_OFS) = f2u(u2f(AJ(w, RK_AJ_DFR, RK_AJ_RAW_NOW)) - p->mechend_pos_deg[RK_AJ_DFR]);
  }
  AJ(w, k, RK_AJ_OFS) = f2u(u2f(AJ(w, k, RK_AJ_RAW_NOW)) - p->mechend_pos_deg[k]);
}
static void adh_joint_init(const rk_adt_params_t *p, uint32_t *w, int axis) { /* JointBase::init() and its overrides */
  int k = ADT_AXIS[axis];
  if(k == RK_AJ_Y0) { /* JointIcsServo::init  AD_joint_ics_servo.cpp:35-55 over the ideal servo (setFree answers the last position) */
    float now = (float)ics_posDeg100((int)(int32_t)w[RK_AS_ICS_SERVO] + 7500) * 0.01f * p->motor_dir[RK_AJ_Y0];
    AJ(w, k, RK_AJ_RAW_NOW) = f2u(now), AJ(w, k, RK_AJ_RAW_TGT) = f2u(now);
    w[RK_AS_ICS_POS] = (uint32_t)-1;
    set_jflag(w, k, jflag(w, k) | RK_AJF_CONNECTED);
  } else if(k == RK_AJ_P1) { /* JointMgServo::init  AD_joint_mg_servo.cpp:38-48 */
    set_jflag(w, k, (jflag(w, k) & ~(RK_AJF_TORQUE_PREV | RK_AJF_TORQUE_ON)) | RK_AJF_CONNECTED);
    w[RK_AS_MG_CTRL + MGC_GAINSET] = 1; /* set_myctrl_gain_params(InitGain): gains, and set_VelLpf_CutOff resets the IIR */
    w[RK_AS_MG_CTRL + MGC_LPF_Y] = 0, w[RK_AS_MG_CTRL + MGC_LPF_X] = 0;
  }
}
/* exec_move_initpos of either mode (AD_mode_initialize.cpp:113-143 / AD_mode_initpos_move.cpp:70-95) for one axis;
 * returns whether the axis has arrived */
static int adh_ramp_axis(const rk_adt_params_t *p, uint32_t *w, int axis, float dir_vel) {
  int   k = ADT_AXIS[axis], arrived;
  float initpos = p->initpos_deg[k], nowpos = adt_get_tgt_deg(w, k);
  float vel = dir_vel * adt_absf(p->vel_init_degps[k]);
  float tgtpos = nowpos + vel * p->cycle_time_s;
  arrived = ((vel > 0) && (tgtpos > initpos)) || ((vel < 0) && (tgtpos < initpos));
  if(arrived) tgtpos = initpos;
  adt_set_tgt(p, w, axis, tgtpos);
  adh_set_curlim(w, axis, p->curlim_default_A[k]);
  return arrived;
}
static void adh_move_mechend(const rk_adt_params_t *p, uint32_t *w, int axis) { /* ax_move_mechend :150-167 */
  int   k = ADT_AXIS[axis];
  float vel = p->vel_init_degps[k], nowpos = adt_get_now_deg(p, w, axis), tgtpos = adt_get_tgt_deg(w, k);
  if(adt_absf(nowpos - tgtpos) > 45.0f) adt_set_tgt(p, w, axis, tgtpos); /* gone too far: hold */
  else adt_set_tgt(p, w, axis, tgtpos + vel * p->cycle_time_s);
  adh_set_curlim(w, axis, p->curlim_init_A[k]);
}
static void adh_mode_update(const rk_adt_params_t *p, uint32_t *w, uint32_t *hw) {
  uint32_t state = hw[RK_HS_STATE] & 0xFFu, comp = hw[RK_HS_STATE] & RK_AS_FSM_IS_COMP, mode = hw[RK_HS_STATE] >> 16;
  uint32_t cnt = hw[RK_HS_WAIT_CNT] & 0xFFFFu;
  int      ax, all;
  if(mode == RK_ADH_MODE_INIT) {
    switch(state) {
    case 0: /* exec_init :43-50 */
      for(ax = 0; ax < 5; ax++) adh_joint_init(p, w, ax), adh_set_initialized(w, ax, 0);
      state = 1;
      break;
    case 1: /* exec_torqueon :56-72 */
      if(cnt == 0) {
        for(ax = 0; ax < 5; ax++) adh_set_torque_on(w, ax, 1);
        cnt++;
      } else if(cnt == 100) state = 2, cnt = 0;
      else cnt++;
      break;
    case 2: /* exec_move_mechend :79-94 */
      if(cnt < 500) {
        adh_move_mechend(p, w, 1);
        adh_move_mechend(p, w, 4);
        cnt++;
      } else if(cnt == 500) state = 3, cnt = 0;
      break;
    case 3: /* exec_resetangle :100-109 + ax_reset_angle :174-179 */
      for(ax = 1; ax < 5; ax++) {
        adh_mech_reset_pos(p, w, ax);
        adt_set_tgt(p, w, ax, adt_get_now_deg(p, w, ax));
      }
      state = 4;
      break;
    case 4: /* exec_move_initpos :115-143 */
      all = 1;
      for(ax = 0; ax < 5; ax++) {
        int   k = ADT_AXIS[ax];
        float d = p->initpos_deg[k] - adt_get_tgt_deg(w, k);
        adh_set_initialized(w, ax, 1); /* before set_tgt_ang_deg in the reference too; neither reads the other */
        all &= adh_ramp_axis(p, w, ax, (d >= 0.0f) ? 1.0f : -1.0f);
      }
      if(all) state = 5;
      break;
    case 5: comp = RK_AS_FSM_IS_COMP; break;
    default: break;
    }
  } else if(mode == RK_ADH_MODE_INIT_POS_MOVE) {
    switch(state) {
    case 0: /* exec_init :37-45 */
      for(ax = 0; ax < 5; ax++) {
        int k = ADT_AXIS[ax];
        adt_set_tgt(p, w, ax, adt_get_now_deg(p, w, ax));
        hw[RK_HS_VEL_DIR + ax] = f2u((p->initpos_deg[k] >= adt_get_now_deg(p, w, ax)) ? 1.0f : -1.0f);
      }
      state = 1;
      break;
    case 1: /* exec_torqueon :51-67 */
      if(cnt == 0) {
        for(ax = 0; ax < 5; ax++) adh_set_torque_on(w, ax, 1);
        cnt++;
      } else if(cnt == 100) state = 2, cnt = 0;
      else cnt++;
      break;
    case 2: /* exec_move_initpos :73-95 */
      all = 1;
      for(ax = 0; ax < 5; ax++) all &= adh_ramp_axis(p, w, ax, u2f(hw[RK_HS_VEL_DIR + ax]));
      if(all) state = 3;
      break;
    case 3: comp = RK_AS_FSM_IS_COMP; break;
    default: break;
    }
  }
  hw[RK_HS_STATE]    = state | comp | (mode << 16);
  hw[RK_HS_WAIT_CNT] = cnt;
}

/* op 0 = rk_adh_mode_init(mode = K), op 2 = K ticks (+ feedback stream, + trace) */
void orc_adh_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *hstate, int64_t n, int64_t i0, int64_t i1, int K,
                   const float *now, uint32_t *trace) {
  int64_t i;
  int     k, t;
  for(i = i0; i < i1; i++) {
    uint32_t w[RK_AS_WORDS], hw[RK_HS_WORDS];
    for(k = 0; k < RK_AS_WORDS; k++) w[k] = *soa(state, n, i, k);
    for(k = 0; k < RK_HS_WORDS; k++) hw[k] = *soa(hstate, n, i, k);
    if(op == 0) { /* ADTModeBase::init(): is_comp = false; doInit(): nowState = INIT, u16_wait_cnt_ = 0, flags / directions zeroed */
      memset(hw, 0, sizeof(hw));
      hw[RK_HS_STATE] = (uint32_t)K << 16;
    } else {
      for(t = 0; t < K; t++) {
        if(now) {
          static const int JK[4] = {RK_AJ_P1, RK_AJ_DFL, RK_AJ_DFR, RK_AJ_P3};
          for(k = 0; k < 4; k++) AJ(w, JK[k], RK_AJ_RAW_NOW) = f2u(now[((int64_t)t * 4 + k) * n + i]);
        }
        adh_mode_update(p, w, hw);
        adt_mg_update(p, w);
        adt_bldc_update(p, w, 0);
        adt_bldc_update(p, w, 1);
        adt_bldc_update(p, w, 2);
        adt_ics_update(p, w);
        if(trace) {
          uint32_t *tr = trace + (int64_t)t * RK_ADT_TRACE_WORDS * n + i;
          for(k = 0; k < 5; k++) tr[(int64_t)k * n] = f2u(adt_get_tgt_deg(w, ADT_AXIS[k]));
          tr[5 * n] = w[RK_AS_MG_TX] >> 16, tr[6 * n] = w[RK_AS_MG_TX + 1];
          for(k = 0; k < 3; k++) tr[(int64_t)(7 + k) * n] = w[RK_AS_BLDC_TX0 + 4 * k];
          tr[10 * n] = w[RK_AS_ICS_POS];
          tr[11 * n] = hw[RK_HS_STATE] & 0xFFu, tr[12 * n] = hw[RK_HS_WAIT_CNT];
          tr[13 * n] = bldc_id_byte(w[RK_AS_BLDC_TX0 + 2]) | (bldc_id_byte(w[RK_AS_BLDC_TX0 + 6]) << 8) | (bldc_id_byte(w[RK_AS_BLDC_TX0 + 10]) << 16);
          tr[14 * n] = 0, tr[15 * n] = 0;
        }
      }
      for(k = 0; k < RK_AS_WORDS; k++) *soa(state, n, i, k) = w[k];
    }
    for(k = 0; k < RK_HS_WORDS; k++) *soa(hstate, n, i, k) = hw[k];
  }
}


/* ------------------------------------------------------------------------------------ */
/* WIT serial codec: WitSerialDataIn / CopeWitData (lib/wt901c/wit_c_sdk.c:77-164) + the update flags   */
/* of SensorDataUpdata (imu_if_wt901c.cpp:24-46) + IMU_IF_WT901C::update / init on the parsed registers */
typedef struct {
  uint8_t  buf[11], cnt;
  uint32_t flags; /* bit 0 QUAT_UPDATE; bits 8-15 s_uiReadRegIndex */
  int16_t  reg[16];
} wit_t;
static int wit_reg_slot(uint32_t r) { /* tracked registers -> RK_IMT_REG_* slot, else -1 */
  if(r >= 0x34 && r <= 0x3f) return (int)(r - 0x34);
  if(r >= 0x51 && r <= 0x54) return (int)(r - 0x51) + 12;
  return -1;
}
static void wit_write_regs(wit_t *w, uint32_t reg, uint32_t len, const uint16_t *val) {
  uint32_t k;
  for(k = 0; k < len; k++) {
    int slot = wit_reg_slot(reg + k);
    if(slot >= 0) w->reg[slot] = (int16_t)val[k];
    if(reg + k == 0x54) w->flags |= 1u; /* q3 -> QUAT_UPDATE */
  }
}
static void wit_byte(wit_t *w, uint8_t b) {
  uint16_t d[4];
  uint8_t  sum = 0, type;
  uint32_t reg1 = 0, reg2 = 0, len1 = 4, len2 = 0;
  int      k;
  w->buf[w->cnt++] = b;
  if(w->buf[0] != 0x55) {
    w->cnt--;
    memmove(w->buf, w->buf + 1, w->cnt);
    return;
  }
  if(w->cnt < 11) return;
  for(k = 0; k < 10; k++) sum += w->buf[k];
  if(sum != w->buf[10]) {
    w->cnt--;
    memmove(w->buf, w->buf + 1, w->cnt);
    return;
  }
  for(k = 0; k < 4; k++) d[k] = (uint16_t)(((uint16_t)w->buf[3 + 2 * k] << 8) | w->buf[2 + 2 * k]);
  type   = w->buf[1];
  w->cnt = 0;
  switch(type) { /* CopeWitData :85-113 */
  case 0x51: reg1 = 0x34, len1 = 3, reg2 = 0x40, len2 = 1; break; /* WIT_ACC: AX.., TEMP */
  case 0x53: reg1 = 0x3d, len1 = 3, reg2 = 0x2e, len2 = 1; break; /* WIT_ANGLE: Roll.., VERSION */
  case 0x50: reg1 = 0x30; break;                                  /* WIT_TIME */
  case 0x52: reg1 = 0x37, len1 = 3; break;                        /* WIT_GYRO */
  case 0x54: reg1 = 0x3a, len1 = 3; break;                        /* WIT_MAGNETIC */
  case 0x55: reg1 = 0x41; break;                                  /* WIT_DPORT */
  case 0x56: reg1 = 0x45; break;                                  /* WIT_PRESS */
  case 0x57: reg1 = 0x49; break;                                  /* WIT_GPS */
  case 0x58: reg1 = 0x4d; break;                                  /* WIT_VELOCITY */
  case 0x59: reg1 = 0x51; break;                                  /* WIT_QUATER */
  case 0x5A: reg1 = 0x55; break;                                  /* WIT_GSA */
  case 0x5F: reg1 = (w->flags >> 8) & 0xFFu; break;               /* WIT_REGVALUE: s_uiReadRegIndex */
  default: return;
  }
  wit_write_regs(w, reg1, len1, d);
  if(len2) wit_write_regs(w, reg2, len2, d + 3);
}

void orc_imt_feed_bytes(uint32_t *state, uint32_t *parser, int64_t n, int64_t i0, int64_t i1, int K, int ncells,
                        const uint32_t *cells, const uint16_t *nbytes, uint32_t *out, float *yaw_rad, int do_init) {
  int64_t i;
  int     u, k, b;
  for(i = i0; i < i1; i++) {
    float    qi[4], d[16];
    uint32_t flags = *soa(state, n, i, RK_IS_FLAGS);
    wit_t    w;
    for(k = 0; k < 4; k++) qi[k] = u2f(*soa(state, n, i, RK_IS_QINIT + k));
    for(k = 0; k < 16; k++) d[k] = u2f(*soa(state, n, i, RK_IS_DATA + k));
    for(k = 0; k < 12; k++) {
      uint32_t word = *soa(parser, n, i, RK_IP_WINDOW + k / 4);
      uint8_t  by   = (uint8_t)(word >> (8 * (k % 4)));
      if(k < 11) w.buf[k] = by;
      else w.cnt = by > 10 ? 10 : by;
    }
    w.flags = *soa(parser, n, i, RK_IP_FLAGS);
    for(k = 0; k < 16; k++) w.reg[k] = (int16_t)(*soa(parser, n, i, RK_IP_SREG + k / 2) >> (16 * (k % 2)));
    for(u = 0; u < K; u++) {
      int init = do_init && u == 0;
      if(init) { /* WitInit: s_uiWitDataCnt = 0 ; WitReadReg(q0, 4): s_uiReadRegIndex = q0 */
        w.cnt   = 0;
        w.flags = (w.flags & ~0xFF00u) | (0x51u << 8) | RK_IP_FLAG_INIT_PENDING;
      }
      {
        int nb = 16 * ncells;
        if(nbytes && nbytes[(int64_t)u * n + i] < nb) nb = nbytes[(int64_t)u * n + i];
        for(b = 0; b < nb; b++) { /* byte b of the update: cell b/16, word (b%16)/4, from the low byte up */
          uint32_t word = cells[(((int64_t)u * ncells + b / 16) * n + i) * 4 + (b % 16) / 4];
          wit_byte(&w, (uint8_t)(word >> (8 * (b % 4))));
        }
      }
      {
        int hq = (w.flags & 1u) != 0;
        if(hq) w.flags &= ~0xFFu; /* s_cDataUpdate = 0 */
        if(w.flags & RK_IP_FLAG_INIT_PENDING) {
          /* init() -> getDataImmediately (imu_if_wt901c.cpp:63-77,149-158) spins in isComComp() until a quaternion frame
           * has arrived: an update slot without one is still part of that wait (nothing is published); the slot that
           * brings it completes init: updateData, q_init latched */
          if(hq) {
            imu_update_data(qi, w.reg, d);
            for(k = 0; k < 4; k++) qi[k] = w.reg[RK_IMT_REG_Q0 + k] / 32768.0f;
            w.flags &= ~RK_IP_FLAG_INIT_PENDING;
          }
        } else if(hq) {
          flags &= ~RK_IS_FLAG_ERROR;
          imu_update_data(qi, w.reg, d);
        } else {
          flags |= RK_IS_FLAG_ERROR;
        }
      }
      if(out)
        for(k = 0; k < 16; k++) out[(((int64_t)u * 4 + k / 4) * n + i) * 4 + (k % 4)] = f2u(d[k]);
      if(yaw_rad) yaw_rad[(int64_t)u * n + i] = d[RK_IS_D_ANGLE + 2] * (ORC_PI / 180.0f);
    }
    for(k = 0; k < 4; k++) *soa(state, n, i, RK_IS_QINIT + k) = f2u(qi[k]);
    for(k = 0; k < 16; k++) *soa(state, n, i, RK_IS_DATA + k) = f2u(d[k]);
    *soa(state, n, i, RK_IS_FLAGS) = flags;
    {
      uint32_t pw[RK_IP_WORDS];
      memset(pw, 0, sizeof(pw));
      for(k = 0; k < (int)w.cnt; k++) pw[k / 4] |= (uint32_t)w.buf[k] << (8 * (k % 4)); /* bytes past the count are kept zero */
      pw[2] |= (uint32_t)w.cnt << 24;
      pw[RK_IP_FLAGS] = w.flags;
      for(k = 0; k < 16; k++) pw[RK_IP_SREG + k / 2] |= ((uint32_t)(uint16_t)w.reg[k]) << (16 * (k % 2));
      for(k = 0; k < RK_IP_WORDS; k++) *soa(parser, n, i, k) = pw[k];
    }
  }
}


/* ------------------------------------------------------------------------------------ */
/* RobotManager guard: the vehicle-management block of routine_ros()                      */
/* (src/RobotManager/RM_task_main.cpp:484-767) + UTIL::mymath::atanf / atan2f                */
/* (src/Utility/util_mymath.cpp:98-126; tables restated by tools/gen_atan_table.py)          */
#include "atan_table.inc"
static const uint32_t orc_atan_table_bits[]   = {RK_ATAN_TABLE_BITS};
static const uint32_t orc_atan_delimit_bits[] = {RK_ATAN_DELIMIT_BITS};
static const uint32_t orc_atan_width_bits[]   = {RK_ATAN_WIDTH_BITS};
#define orc_atan_table(i) u2f(orc_atan_table_bits[i])
#define orc_atan_delimit(i) u2f(orc_atan_delimit_bits[i])
#define orc_atan_width(i) u2f(orc_atan_width_bits[i])

float orc_atanf(float x) { /* util_mymath.cpp:98-115 */
  int   i, index_int;
  float index, index_dec;
  if(x < 0) return -orc_atanf(-x);
  if(x == 0.0) return 0.0f;
  for(i = 1; i <= 26; i++) {
    if(x <= orc_atan_delimit(i)) {
      index     = 24 * (i - 1) + ((x - orc_atan_delimit(i - 1)) / orc_atan_width(i - 1));
      index_int = (int)index;
      index_dec = index - (float)index_int;
      return orc_atan_table(index_int) + index_dec * ((orc_atan_table(index_int + 1) - orc_atan_table(index_int)));
    }
  }
  return orc_atan_table(577 - 1); /* TABLE_SIXE_ATAN is 577 although the table has 625 entries: :6,114 */
}
float orc_atan2f(float y, float x) { /* :117-126 */
  if(x > 0.0) return orc_atanf(y / x);
  if(y >= 0.0 && x < 0.0) return orc_atanf(y / x) + ORC_PI;
  if(y < 0.0 && x < 0.0) return orc_atanf(y / x) - ORC_PI;
  if(y > 0.0 && x == 0.0) return (float)(ORC_PI / 2.0);
  if(y < 0.0 && x == 0.0) return (float)(-ORC_PI / 2.0);
  return 0.0f;
}

enum { RM_REQ_MOVE_DIR = 1, RM_REQ_MOVE_CONT_DIR = 2 }; /* VDT::MSG_ID  VD_task_main.hpp:8-12 */
enum { RM_FLOOR = 1, RM_WALL = 2 };                     /* FD_task_main.hpp:20-22 */
typedef struct {
  uint32_t id, cmd, time_ms, speed; /* MSG_ReqMoveDir */
  float    vx, vy, vth;             /* MSG_ReqMoveContDir (time_ms shared) */
} rm_msg_t;
static double rm_double(const uint32_t *w) {
  uint64_t u = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  double   d;
  memcpy(&d, &u, 8);
  return d;
}
static void rm_stop(rm_msg_t *m) { /* the MOVE_STOP every veto writes  :584-589 */
  m->id = RM_REQ_MOVE_DIR, m->cmd = RK_DIR_MOVE_STOP, m->time_ms = 1, m->speed = 0;
}
static rk_vdt_cmd_t rm_record(const rm_msg_t *m) {
  rk_vdt_cmd_t c;
  memset(&c, 0, sizeof(c));
  if(m->id == RM_REQ_MOVE_DIR) {
    memcpy(&c.vx, &m->cmd, 4), memcpy(&c.vy, &m->speed, 4);
    c.kind = (int32_t)(RK_CMD_MSG_MOVE_DIR | (m->time_ms << 8));
  } else {
    c.vx = m->vx, c.vy = m->vy, c.vth = m->vth;
    c.kind = (int32_t)(RK_CMD_MSG_MOVE_CONT_DIR | (m->time_ms << 8));
  }
  return c;
}

void orc_rmt_guard(const rk_rmt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const uint32_t *in,
                   rk_vdt_cmd_t *cmd_out, uint32_t *abort_out) {
  int64_t i;
  int     u, k;
  for(i = i0; i < i1; i++) {
    uint32_t cmd_status = *soa(state, n, i, RK_RS_CMD_STATUS), ignore = *soa(state, n, i, RK_RS_IGNORE_FLOOR);
    uint32_t no_cmd = *soa(state, n, i, RK_RS_NO_CMD_CNT), abort_v = *soa(state, n, i, RK_RS_ABORT);
    for(u = 0; u < K; u++) {
      uint32_t w[RK_RI_WORDS];
      rm_msg_t buf, msg;
      int      updated = 0, exist = 0, sent = 0, nf = 0, nw = 0;
      uint8_t  rF, lF, rB, lB, right, left, fwd, back, fl[8];
      rk_vdt_cmd_t rec;
      for(k = 0; k < RK_RI_WORDS; k++) w[k] = in[(((int64_t)u * 3 + k / 4) * n + i) * 4 + (k % 4)];
      memset(&buf, 0, sizeof(buf));
      /* rclc_executor_spin_some(): the subscription callbacks  :159-248 */
      switch(w[RK_RI_KIND]) {
      case RK_ROS_MECANUM_CMD: buf.id = RM_REQ_MOVE_DIR, buf.cmd = w[RK_RI_A], buf.time_ms = w[RK_RI_B], buf.speed = w[RK_RI_C], updated = 1; break;
      case RK_ROS_MECANUM_CONT:
        buf.id = RM_REQ_MOVE_CONT_DIR, buf.vx = (float)rm_double(w + RK_RI_X), buf.vy = (float)rm_double(w + RK_RI_Y);
        buf.vth = (float)rm_double(w + RK_RI_Z), buf.time_ms = w[RK_RI_A], updated = 1;
        break;
      case RK_ROS_CMD_VEL:
        buf.id = RM_REQ_MOVE_CONT_DIR, buf.vx = (float)(rm_double(w + RK_RI_X) * 1000.0), buf.vy = (float)(rm_double(w + RK_RI_Y) * 1000.0);
        buf.vth = (float)rm_double(w + RK_RI_Z), buf.time_ms = 500, updated = 1;
        break;
      case RK_ROS_COMMAND: /* a Command always stops the vehicle, then switches the manager's mode  :162-201 */
        rm_stop(&buf), updated = 1;
        cmd_status = w[RK_RI_A];
        switch(w[RK_RI_A]) {
        case 0: case 1: case 2: case 4: break;         /* RELAX, MOVE_READY, MOVE_START, INIT (arm / gimbal requests: out of scope) */
        case 10: ignore = !ignore; break;              /* SWITCH_FLOOR_SENSOR */
        default: cmd_status = 0xFF; break;             /* QUIT_PG and the rest: UNKNOWN_CMD */
        }
        break;
      default: break;
      }
      /* :484-505 */
      if(updated) exist = 1, abort_v = 0, msg = buf;
      else memset(&msg, 0, sizeof(msg)), msg.id = RM_REQ_MOVE_DIR;
      for(k = 0; k < 8; k++) fl[k] = (uint8_t)(w[RK_RI_FLOOR + k / 4] >> (8 * (k % 4)));
      for(k = 0; k < 8; k++) { /* :507-530 */
        if(fl[k] == 0) nf++;
        else if(fl[k] == RM_WALL) nw++;
      }
      if(nf >= 5 || nw >= 5 || ignore)
        for(k = 0; k < 8; k++) fl[k] = RM_FLOOR; /* :532-542 */
      rF = fl[0], lF = fl[1], rB = fl[2], lB = fl[3], right = fl[4], left = fl[5], fwd = fl[6], back = fl[7];
      if(cmd_status == 2) { /* MOVE_START: leave a wall (an opponent)  :546-577 */
        uint32_t dir = 0;
        if(fwd == RM_WALL) dir = RK_DIR_GO_BACK, abort_v |= 1u << 0;
        else if(back == RM_WALL) dir = RK_DIR_GO_FORWARD, abort_v |= 1u << 1;
        else if(left == RM_WALL) dir = RK_DIR_GO_RIGHT, abort_v |= 1u << 2;
        else if(right == RM_WALL) dir = RK_DIR_GO_LEFT, abort_v |= 1u << 3;
        if(dir) msg.id = RM_REQ_MOVE_DIR, msg.cmd = dir, msg.time_ms = p->wall_leave_time_ms, msg.speed = p->wall_leave_speed_mmps, exist = 1;
      }
      if(msg.id == RM_REQ_MOVE_DIR) { /* :581-673 */
        uint8_t  need = RM_FLOOR;
        uint32_t bits = 0;
        switch(msg.cmd) {
        case RK_DIR_GO_FORWARD: need = fwd, bits = 1u << 8; break;
        case RK_DIR_GO_BACK: need = back, bits = 1u << 9; break;
        case RK_DIR_GO_RIGHT: need = right, bits = 1u << 11; break;
        case RK_DIR_GO_LEFT: need = left, bits = 1u << 10; break;
        case RK_DIR_GO_RIGHT_FORWARD: need = rF, bits = (1u << 8) | (1u << 11); break;
        case RK_DIR_GO_LEFT_FORWARD: need = lF, bits = (1u << 8) | (1u << 10); break;
        case RK_DIR_GO_RIGHT_BACK: need = rB, bits = (1u << 9) | (1u << 11); break;
        case RK_DIR_GO_LEFT_BACK: need = lB, bits = (1u << 9) | (1u << 10); break;
        default: break;
        }
        if(need != RM_FLOOR) rm_stop(&msg), exist = 1, abort_v |= bits;
      } else if(msg.id == RM_REQ_MOVE_CONT_DIR) { /* :674-749 */
        const float ax = msg.vx < 0 ? -msg.vx : msg.vx, ay = msg.vy < 0 ? -msg.vy : msg.vy;
        if(!(ax < 0.01f && ay < 0.01f)) {
          const float vph = orc_atan2f(msg.vy, msg.vx);
          int         veto = 0;
          if(fwd != RM_FLOOR && (-3.1415f * 0.33f < vph && vph <= +3.1415f * 0.33f)) veto = 1;
          if(back != RM_FLOOR && (+3.1415f * 0.66f < vph || vph <= -3.1415f * 0.66f)) veto = 1;
          if(left != RM_FLOOR && (+3.1415f * 0.16f < vph && vph <= +3.1415f * 0.84f)) veto = 1;
          if(right != RM_FLOOR && (-3.1415f * 0.84f < vph && vph <= -3.1415f * 0.16f)) veto = 1;
          if(rB != RM_FLOOR && (+3.1415f * 0.92f < vph || vph <= -3.1415f * 0.42f)) veto = 1;
          if(rF != RM_FLOOR && (-3.1415f * 0.58f < vph && vph <= +3.1415f * 0.08f)) veto = 1;
          if(lF != RM_FLOOR && (-3.1415f * 0.08f < vph && vph <= +3.1415f * 0.58f)) veto = 1;
          if(lB != RM_FLOOR && (+3.1415f * 0.42f < vph || vph <= -3.1415f * 0.92f)) veto = 1;
          if(veto) msg.vx = 0, msg.vy = 0, abort_v |= 1u << 16;
        }
      }
      memset(&rec, 0, sizeof(rec));
      if(exist) no_cmd = 0, rec = rm_record(&msg), sent = 1; /* :752-757 */
      else no_cmd++;
      if(no_cmd > p->no_cmd_stop_thre) { /* :760-767 */
        rm_stop(&msg), rec = rm_record(&msg), sent = 1;
        no_cmd = 0;
      }
      (void)sent;
      cmd_out[(int64_t)u * n + i] = rec;
      if(abort_out) abort_out[(int64_t)u * n + i] = abort_v;
    }
    *soa(state, n, i, RK_RS_CMD_STATUS) = cmd_status, *soa(state, n, i, RK_RS_IGNORE_FLOOR) = ignore;
    *soa(state, n, i, RK_RS_NO_CMD_CNT) = no_cmd, *soa(state, n, i, RK_RS_ABORT) = abort_v;
  }
}
