// rk_vehicle_fast.cuh -- the issue-optimised closed-loop tick (RK_SENSOR_PLANT).
//
// Same results as rk_vehicle.cuh (the direct transcription of the firmware), word for word,
// but restructured for the B200 FP32 issue roofline:
//   * plant -> CAN frame -> rx_callback collapses algebraically: with |rpm| < 30000 the
//     per-tick angle step |dang| < 4096, so the +-4096 unwrap of rx_callback
//     (VD_motor_if_m2006.cpp:66-69) returns exactly dir*dang; no byte frame is built and the
//     s64 angle sum is carried as a 32-bit per-launch delta;
//   * every x / c with a launch-constant divisor uses q = x*rcp; e = fma(q,c,-x);
//     q' = fma(-(q*c-x),rcp,q), which rk_exact.cu proves bit-identical to IEEE x / c by
//     exhaustive search over all 2^32 inputs before the fast path is enabled;
//   * the odometry product (double)d * OUT_RAD_PER_RAW_ANGLE * GEAR_RATIO_INV -> float
//     (VD_vehicle_controller.cpp:37-41) is fma(d, K_hi, d*K_lo) in FP32, verified against the
//     FP64 expression for every |d| <= 8192 (rk_exact.cu);
//   * the three VelInterpConstJerk updates and all clamps are branch-free selects / FMNMX;
//   * motor directions are compile-time (+1,+1,-1,-1 = VD_task_main.cpp:75-78), so negations
//     fold into operand modifiers.
// Preconditions are checked per thread (fast_ok) and per command (cmd_ok); a thread that
// fails one runs the transcription instead, so results never depend on which path ran.
#pragma once
#include "rk_vehicle.cuh"

namespace rk {

// Launch-constant values prepared (and verified) on the host side once per parameter set.
struct FastConsts {
  float rcp_r, rcp_s2, rcp_l; // RN(1/c)
  float k_hi, k_lo;           // split of (double)OUT_RAD_PER_RAW_ANGLE * (double)GEAR_RATIO_INV
  float A1, B0, ki_dt, s2l;   // = Derived (computed on device by derive(), copied here)
  float neg_i_limit, neg_ff_limit;
  float2 khi_pm, klo_pm;      // {k_hi, -k_hi}, {k_lo, -k_lo}: lane constants of the packed tick, pinned in registers
};

// exact x / c for launch-constant c > 0 (see header comment); rcp = RN(1/c).
// e = q*c - x is exact; writing the correction as fma(-e, rcp, q) (rather than r = x - q*c,
// fma(r, rcp, q)) also returns the IEEE sign for x = +-0, with no fix-up instruction.
RK_DEV float div_const(float x, float c, float rcp) {
  const float q = fmul(x, rcp);
  const float e = __fmaf_rn(q, c, -x);
  return __fmaf_rn(-e, rcp, q);
}

// rpm * 8192 / 60000 (C truncating division) for |rpm| <= 32768: one signed high multiply by
// ceil(2^32 * 8192 / 60000) gives the floor; adding the sign bit turns it into truncation.
// Verified for every int16 by rk_exact.cu.
RK_DEV int32_t plant_dang(int32_t rpm) { return __mulhi(rpm, 586406202) + (int32_t)((uint32_t)rpm >> 31); }

struct FastInterp { // VelInterpConstJerk with the phase thresholds hoisted out of the tick
  float vel, acl, dt;
  float t1, t2, t3;                                         // dt1+ts, (dt1+dt2)+ts, ((dt1+dt2)+dt3)+ts
  float vel_tgt, acl_max, jerk_p, jerk_m, dt1, dt2, dt3, vel_ini, acl_ini;
};
struct FastWheel {
  int32_t rpm, cur;  // plant rpm and s16_rawCurr_tgt, motor frame
  int32_t dsum;      // sum over this chunk of the WHEEL-frame angle steps (= the s16 deltas rx_callback adds)
  float   prev_val, integ, lpf_y, lpf_x; // FF_PI_D state
  float   b0x;                           // B0 * lpf_x, carried so the product is formed once (B1 == B0)
};
struct FastVeh {
  float      px, py;
  FastInterp it[3];
  FastWheel  w[4];
};

RK_DEV void fast_interp_load(FastInterp &f, const Interp &t, float ts) {
  f.vel = t.vel, f.acl = t.acl, f.dt = t.dt;
  f.vel_tgt = t.vel_tgt, f.acl_max = t.acl_max, f.jerk_p = t.jerk_p, f.jerk_m = t.jerk_m;
  f.dt1 = t.dt1, f.dt2 = t.dt2, f.dt3 = t.dt3, f.vel_ini = t.vel_ini, f.acl_ini = t.acl_ini;
  f.t1 = fadd(t.dt1, ts);
  f.t2 = fadd(fadd(t.dt1, t.dt2), ts);
  f.t3 = fadd(fadd(fadd(t.dt1, t.dt2), t.dt3), ts);
  // keep the thresholds in registers: without this ptxas rematerialises the six adds per
  // interpolator every tick
  asm volatile("" : "+f"(f.t1), "+f"(f.t2), "+f"(f.t3));
}
RK_DEV void fast_interp_store(const FastInterp &f, Interp &t) {
  t.vel = f.vel, t.acl = f.acl, t.dt = f.dt;
  t.vel_tgt = f.vel_tgt, t.acl_max = f.acl_max, t.jerk_p = f.jerk_p, t.jerk_m = f.jerk_m;
  t.dt1 = f.dt1, t.dt2 = f.dt2, t.dt3 = f.dt3, t.vel_ini = f.vel_ini, t.acl_ini = f.acl_ini;
}

// util_vel_interp.hpp:110-136 as selects.  The if / else-if chain picks the FIRST true
// condition, which the nested selects reproduce for any ordering of the thresholds.
RK_DEV float fast_interp_update(FastInterp &f, float ts) {
  const bool  p1  = f.dt <= f.t1, p2 = f.dt <= f.t2, p3 = f.dt <= f.t3;
  const float a1  = fadd(f.acl_ini, fmul(f.jerk_p, f.dt));
  const float v1  = fadd(f.vel_ini, fmul(fmul(fadd(f.acl_ini, a1), f.dt), 0.5f));
  const float a3  = fadd(f.acl_max, fmul(f.jerk_m, fsub(fsub(f.dt, f.dt1), f.dt2)));
  const float a23 = p2 ? f.acl_max : a3;
  const float v23 = fadd(f.vel, fmul(a23, ts));
  const bool  p23 = p2 || p3;
  f.acl           = p1 ? a1 : (p23 ? a23 : 0.0f);
  f.vel           = p1 ? v1 : (p23 ? v23 : f.vel_tgt);
  f.dt            = (p1 || p23) ? fadd(f.dt, ts) : f.dt;
  return f.vel;
}

// (x >= lim) ? lim : ((x <= -lim) ? -lim : x) for lim > 0 and non-NaN x  (util_controller.hpp:101,162)
RK_DEV float clamp_sym(float x, float lim, float neg_lim) { return fminf(fmaxf(x, neg_lim), lim); }

// Which motor state can enter the fast path (see header comment).
template <int DIR> RK_DEV bool fast_motor_ok(const Motor &m, int lim) {
  const int raw = (DIR == 1) ? m.p_ang : 8192 - m.p_ang;
  bool      ok  = (m.sum == m.prev);
  ok &= (m.p_ang >= 0) && (m.p_ang <= 8191);
  ok &= (m.p_rpm >= -4 * lim) && (m.p_rpm <= 4 * lim);
  ok &= (m.cur_tgt >= -lim) && (m.cur_tgt <= lim);
  ok &= (m.ang == raw) || (m.ang == 0 && raw == 8192);
  return ok;
}
RK_DEV bool finite_bounded(float x) { return fabsf(x) <= 1.0e9f; } // false for NaN / Inf

// RK_SENSOR_STREAM: the plant fields are not used; the status words are whatever the last frame said
RK_DEV bool fast_motor_ok_stream(const Motor &m, int lim) { return (m.sum == m.prev) && (m.cur_tgt >= -lim) && (m.cur_tgt <= lim); }

template <int D0, int D1, int D2, int D3, bool STREAM = false>
RK_DEV bool fast_ok(const Veh &v, const rk_vdt_params_t &p) {
  bool ok = (v.flags & RK_VS_FLAG_POWER_ON) != 0;
  if(STREAM) {
#pragma unroll
    for(int k = 0; k < 4; k++) ok &= fast_motor_ok_stream(v.m[k], p.raw_curr_lim);
  } else {
    ok &= fast_motor_ok<D0>(v.m[0], p.raw_curr_lim) && fast_motor_ok<D1>(v.m[1], p.raw_curr_lim);
    ok &= fast_motor_ok<D2>(v.m[2], p.raw_curr_lim) && fast_motor_ok<D3>(v.m[3], p.raw_curr_lim);
  }
  ok &= finite_bounded(v.pos[0]) && finite_bounded(v.pos[1]);
#pragma unroll
  for(int a = 0; a < 3; a++) {
    const Interp &t = v.it[a];
    ok &= finite_bounded(t.vel) && finite_bounded(t.acl) && finite_bounded(t.vel_tgt) && finite_bounded(t.acl_max);
    ok &= finite_bounded(t.jerk_p) && finite_bounded(t.jerk_m) && finite_bounded(t.dt1) && finite_bounded(t.dt2);
    ok &= finite_bounded(t.dt3) && finite_bounded(t.vel_ini) && finite_bounded(t.acl_ini) && finite_bounded(t.dt);
  }
#pragma unroll
  for(int k = 0; k < 4; k++) {
    const Ctrl &c = v.c[k];
    ok &= finite_bounded(c.prev_val) && finite_bounded(c.integ) && finite_bounded(c.lpf_y) && finite_bounded(c.lpf_x);
  }
  return ok;
}

template <int D0, int D1, int D2, int D3>
RK_DEV void to_fast(const Veh &v, FastVeh &f, float ts, float b0) {
  f.px = v.pos[0], f.py = v.pos[1];
#pragma unroll
  for(int a = 0; a < 3; a++) fast_interp_load(f.it[a], v.it[a], ts);
#pragma unroll
  for(int k = 0; k < 4; k++) {
    f.w[k].rpm = v.m[k].p_rpm, f.w[k].cur = v.m[k].cur_tgt, f.w[k].dsum = 0;
    f.w[k].prev_val = v.c[k].prev_val, f.w[k].integ = v.c[k].integ, f.w[k].lpf_y = v.c[k].lpf_y, f.w[k].lpf_x = v.c[k].lpf_x;
    f.w[k].b0x = fmul(b0, v.c[k].lpf_x);
  }
}

// Back to the transcription's state after `nticks` fast ticks.  Fields the next
// veh_update()/motor_rx() overwrite unconditionally (status current, now_tgt/err/ctrl, vel,
// tgt) are left as they were: the launch always ends with transcription ticks.
template <int DIR> RK_DEV void from_fast_motor(Motor &m, const FastWheel &w, int nticks) {
  m.sum += (int64_t)w.dsum;
  m.prev  = m.sum;
  m.p_ang = (m.p_ang + DIR * w.dsum) & 8191;
  m.p_rpm = w.rpm;
  m.ang   = (DIR == 1) ? m.p_ang : 8192 - m.p_ang;
  m.rpm   = DIR * w.rpm;
  m.cur_tgt = w.cur;
  m.head    = (m.head + nticks) % 3;
}
template <int D0, int D1, int D2, int D3>
RK_DEV void from_fast(Veh &v, const FastVeh &f, int nticks) {
  if(nticks <= 0) return;
  v.pos[0] = f.px, v.pos[1] = f.py;
#pragma unroll
  for(int a = 0; a < 3; a++) fast_interp_store(f.it[a], v.it[a]);
#pragma unroll
  for(int k = 0; k < 4; k++) {
    v.c[k].prev_val = f.w[k].prev_val, v.c[k].integ = f.w[k].integ, v.c[k].lpf_y = f.w[k].lpf_y, v.c[k].lpf_x = f.w[k].lpf_x;
  }
  from_fast_motor<D0>(v.m[0], f.w[0], nticks);
  from_fast_motor<D1>(v.m[1], f.w[1], nticks);
  from_fast_motor<D2>(v.m[2], f.w[2], nticks);
  from_fast_motor<D3>(v.m[3], f.w[3], nticks);
}

// One wheel of one tick: plant step, rx_callback (collapsed), wheel speed.  Returns the
// wheel-frame speed Mvel (VD_vehicle_controller.cpp:21-24) and odometry increment Mrad (:37-41).
template <int DIR>
RK_DEV void fast_wheel_sense(FastWheel &w, const FastConsts &fc, float &mvel, float &mrad) {
  w.rpm += ((w.cur * 4 - w.rpm) >> 4);
  // rx_callback applies the direction to the INTEGER fields (VD_motor_if_m2006.cpp:40-46), so
  // negate before converting: (float)(-0) is +0 exactly as in the transcription.  The
  // truncating division is odd-symmetric, so the wheel-frame step is plant_dang(-rpm).
  const int32_t rw   = (DIR == 1) ? w.rpm : -w.rpm;
  const int32_t dang = plant_dang(rw);
  w.dsum += dang;
  mvel           = fmul(fmul((float)rw, RK_RPM_TO_RADPS), RK_GEAR_RATIO_INV);
  const float df = (float)dang;
  mrad           = __fmaf_rn(df, fc.k_hi, fmul(df, fc.k_lo));
}

// FF_PI_D::update + set_CurrA_tgt for one wheel (util_controller.hpp:94-110,159-165;
// VD_motor_if_m2006.hpp:36-37,57).  FFSAT: ff_limit == 1.0f, so the feed-forward clamp
// clamp(t, -1, 1) is sat(t) - sat(-t) with sat = clamp to [0,1] -- two FMUL.SAT on the FMA
// pipe instead of two half-rate FMNMX.  (It returns +0 for t = -0; u only feeds the integer
// conversion here, so the sign of a zero feed-forward is immaterial.)
template <int DIR, bool FFSAT>
RK_DEV void fast_wheel_ctrl(FastWheel &w, const rk_vdt_params_t &p, const FastConsts &fc, float mtgt, float mvel) {
  const float tgt = fmul(mtgt, RK_GEAR_RATIO);
  const float now = fmul(mvel, RK_GEAR_RATIO);
  const float err = fsub(tgt, now);
  const float x   = fmul(fsub(now, w.prev_val), p.ctrl_freq);
  const float b0x = fmul(fc.B0, x);
  const float y   = fadd(fadd(fmul(fc.A1, w.lpf_y), b0x), w.b0x); // B1 * prev_X_ with B1 == B0
  w.lpf_y         = y;
  w.lpf_x         = x;
  w.b0x           = b0x;
  w.integ         = clamp_sym(fadd(w.integ, fmul(fc.ki_dt, err)), p.i_limit, fc.neg_i_limit);
  float u         = fsub(fadd(fmul(p.kp, err), w.integ), fmul(p.kd, y));
  w.prev_val      = now;
  float ff;
  if(FFSAT)
    ff = fsub(__saturatef(fmul(tgt, p.kff)), __saturatef(fmul(tgt, -p.kff)));
  else
    ff = clamp_sym(fmul(tgt, p.kff), p.ff_limit, fc.neg_ff_limit);
  u         = fadd(u, ff);
  int32_t t = __float2int_rz(fmul(u, RK_AMPERE_TO_RAW_CURR));
  t         = sext16(DIR > 0 ? t : -t);
  w.cur     = min(max(t, -p.raw_curr_lim), p.raw_curr_lim);
}

// One fast tick = plant + rx_callback x4 + VEHICLE_CTRL::update()
template <int D0, int D1, int D2, int D3, bool FFSAT>
RK_DEV void fast_tick(FastVeh &f, const rk_vdt_params_t &p, const FastConsts &fc, float cth, float sth,
                      float vel[3], float tgt[3]) {
  float Mvel[4], Mrad[4], Mtgt[4];
  fast_wheel_sense<D0>(f.w[0], fc, Mvel[0], Mrad[0]);
  fast_wheel_sense<D1>(f.w[1], fc, Mvel[1], Mrad[1]);
  fast_wheel_sense<D2>(f.w[2], fc, Mvel[2], Mrad[2]);
  fast_wheel_sense<D3>(f.w[3], fc, Mvel[3], Mrad[3]);
  // conv_Mdir_to_Vdir  VD_vehicle_controller.cpp:126-130
  const float R = p.wheel_radius_mm;
  vel[0] = fmul(fmul(fadd(fadd(fadd(Mvel[0], Mvel[1]), Mvel[2]), Mvel[3]), 0.25f), R);
  vel[1] = fmul(fmul(fadd(fsub(fadd(-Mvel[0], Mvel[1]), Mvel[2]), Mvel[3]), 0.25f), R);
  {
    const float s = fmul(fadd(fadd(fsub(-Mvel[0], Mvel[1]), Mvel[2]), Mvel[3]), 0.25f);
    vel[2]        = fmul(div_const(div_const(s, p.sqrtf2, fc.rcp_s2), p.wheel_l_mm, fc.rcp_l), R);
  }
  const float lx = fmul(fmul(fadd(fadd(fadd(Mrad[0], Mrad[1]), Mrad[2]), Mrad[3]), 0.25f), R);
  const float ly = fmul(fmul(fadd(fsub(fadd(-Mrad[0], Mrad[1]), Mrad[2]), Mrad[3]), 0.25f), R);
  f.px           = fadd(f.px, fmul(fsub(fmul(lx, cth), fmul(ly, sth)), 0.001f));
  f.py           = fadd(f.py, fmul(fadd(fmul(lx, sth), fmul(ly, cth)), 0.001f));
#pragma unroll
  for(int a = 0; a < 3; a++) tgt[a] = fast_interp_update(f.it[a], p.ts);
  // conv_Vdir_to_Mdir  :113-118
  const float T   = fmul(fmul(fc.s2l, tgt[2]), 4.0f);
  const float xmy = fsub(tgt[0], tgt[1]), xpy = fadd(tgt[0], tgt[1]);
  Mtgt[0]         = div_const(fsub(xmy, T), R, fc.rcp_r);
  Mtgt[1]         = div_const(fsub(xpy, T), R, fc.rcp_r);
  Mtgt[2]         = div_const(fadd(xmy, T), R, fc.rcp_r);
  Mtgt[3]         = div_const(fadd(xpy, T), R, fc.rcp_r);
  fast_wheel_ctrl<D0, FFSAT>(f.w[0], p, fc, Mtgt[0], Mvel[0]);
  fast_wheel_ctrl<D1, FFSAT>(f.w[1], p, fc, Mtgt[1], Mvel[1]);
  fast_wheel_ctrl<D2, FFSAT>(f.w[2], p, fc, Mtgt[2], Mvel[2]);
  fast_wheel_ctrl<D3, FFSAT>(f.w[3], p, fc, Mtgt[3], Mvel[3]);
}

// set_target_params on the fast state (shares interp_set with the transcription)
RK_DEV void fast_interp_set(FastInterp &f, float v_t, float a_m, float jrk, float ts) {
  Interp t;
  t.vel = f.vel, t.acl = f.acl;
  interp_set(t, v_t, a_m, jrk);
  fast_interp_load(f, t, ts);
}

} // namespace rk
