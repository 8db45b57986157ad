// rk_vehicle_fast2.cuh -- the closed-loop fast tick of rk_vehicle_fast.cuh, re-expressed on
// Blackwell's packed FP32 instructions (FADD2 / FFMA2, sm_100).
//
// Why: the tick is bound by the SM's ISSUE rate (one warp-instruction per clock per
// sub-partition), not by the FP32 lanes themselves -- bit parity forbids FMA contraction, so
// every algorithmic flop is an instruction, and a third of the instruction stream is
// integer / select / convert work competing for the same issue slots.  A packed instruction
// retires two IEEE-rounded FP32 operations for ONE issue slot (it occupies the FMA pipe for
// two cycles, so the lane throughput is unchanged; measured with tools/ubench/f32x2_probe.cu).
// Pairing used here -- chosen so that no lane shuffles (MOV) are needed:
//   * forward kinematics: lanes = {wheel speed path, odometry-increment path} (the two calls
//     of conv_Mdir_to_Vdir, VD_vehicle_controller.cpp:26,42, run the same formula);
//   * odometry rotation: lanes = {x, y};
//   * jerk-limited targets: lanes = {x interpolator, y interpolator} (theta stays scalar);
//   * inverse kinematics and the four FF_PI_D loops: lanes = {FL, BL} and {BR, FR}.
// Results are bit-identical to the scalar code: each lane performs the same single-rounding
// IEEE operation (packed ops are .rn, denormals preserved).
//
// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false, which
// would break parity.  Every packed multiply is therefore written as fma(a, b, nz) with nz an
// OPAQUE -0.0f (derived from a kernel parameter at run time): a*b + (-0) is exactly RN(a*b)
// including the sign of a zero product, and an FMA cannot be contracted with the add that
// follows it.  cuobjdump shows FFMA2 ... Rnz.F32 followed by separate FADD2s.
#pragma once
#include "rk_vehicle_fast.cuh"

namespace rk {

RK_DEV float2 bc2(float s) { return make_float2(s, s); }
RK_DEV float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
RK_DEV float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
RK_DEV float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, neg2(b)); } // a - b == a + (-b) exactly
RK_DEV float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
RK_DEV float2 mul2(float2 a, float2 b, float nz) { return __ffma2_rn(a, b, make_float2(nz, nz)); }

// exact x / c per lane (div_const of rk_vehicle_fast.cuh)
RK_DEV float2 div_const2(float2 x, float c, float rcp, float nz) {
  const float2 q = mul2(x, bc2(rcp), nz);
  const float2 e = fma2(q, bc2(c), neg2(x));
  return fma2(neg2(e), bc2(rcp), q);
}

struct FastInterp2 { // x and y VelInterpConstJerk, lane-paired
  float2 vel, acl, dt;
  float2 t1, t2, t3;
  float2 vel_tgt, acl_max, jerk_p, jerk_m, dt1, dt2, vel_ini, acl_ini;
  float  dt3x, dt3y; // only needed to store the state back
};
struct FastWheel2 { // two wheels of equal direction, lane-paired
  int32_t rpm[2], cur[2], dsum[2];
  float2  prev_val, integ, lpf_y, lpf_x, b0x;
};
struct FastVeh2 {
  float2      p; // pos x, y
  FastInterp2 xy;
  FastInterp  th;
  FastWheel2  w01, w23;
};

RK_DEV void fast_interp2_load(FastInterp2 &f, const Interp &a, const Interp &b, float ts) {
  FastInterp fa, fb;
  fast_interp_load(fa, a, ts);
  fast_interp_load(fb, b, ts);
  f.vel = make_float2(fa.vel, fb.vel), f.acl = make_float2(fa.acl, fb.acl), f.dt = make_float2(fa.dt, fb.dt);
  f.t1 = make_float2(fa.t1, fb.t1), f.t2 = make_float2(fa.t2, fb.t2), f.t3 = make_float2(fa.t3, fb.t3);
  f.vel_tgt = make_float2(fa.vel_tgt, fb.vel_tgt), f.acl_max = make_float2(fa.acl_max, fb.acl_max);
  f.jerk_p = make_float2(fa.jerk_p, fb.jerk_p), f.jerk_m = make_float2(fa.jerk_m, fb.jerk_m);
  f.dt1 = make_float2(fa.dt1, fb.dt1), f.dt2 = make_float2(fa.dt2, fb.dt2);
  f.vel_ini = make_float2(fa.vel_ini, fb.vel_ini), f.acl_ini = make_float2(fa.acl_ini, fb.acl_ini);
  f.dt3x = fa.dt3, f.dt3y = fb.dt3;
}
RK_DEV void fast_interp2_store(const FastInterp2 &f, Interp &a, Interp &b) {
  a.vel = f.vel.x, a.acl = f.acl.x, a.dt = f.dt.x, a.vel_tgt = f.vel_tgt.x, a.acl_max = f.acl_max.x, a.jerk_p = f.jerk_p.x;
  a.jerk_m = f.jerk_m.x, a.dt1 = f.dt1.x, a.dt2 = f.dt2.x, a.dt3 = f.dt3x, a.vel_ini = f.vel_ini.x, a.acl_ini = f.acl_ini.x;
  b.vel = f.vel.y, b.acl = f.acl.y, b.dt = f.dt.y, b.vel_tgt = f.vel_tgt.y, b.acl_max = f.acl_max.y, b.jerk_p = f.jerk_p.y;
  b.jerk_m = f.jerk_m.y, b.dt1 = f.dt1.y, b.dt2 = f.dt2.y, b.dt3 = f.dt3y, b.vel_ini = f.vel_ini.y, b.acl_ini = f.acl_ini.y;
}

// fast_interp_update on two interpolators at once: arithmetic packed, selects per lane
RK_DEV float2 fast_interp2_update(FastInterp2 &f, float ts, float nz) {
  const float2 a1  = add2(f.acl_ini, mul2(f.jerk_p, f.dt, nz));
  const float2 v1  = add2(f.vel_ini, mul2(mul2(add2(f.acl_ini, a1), f.dt, nz), bc2(0.5f), nz));
  const float2 a3  = add2(f.acl_max, mul2(f.jerk_m, sub2(sub2(f.dt, f.dt1), f.dt2), nz));
  const bool   p1x = f.dt.x <= f.t1.x, p2x = f.dt.x <= f.t2.x, p3x = f.dt.x <= f.t3.x;
  const bool   p1y = f.dt.y <= f.t1.y, p2y = f.dt.y <= f.t2.y, p3y = f.dt.y <= f.t3.y;
  const float2 a23 = make_float2(p2x ? f.acl_max.x : a3.x, p2y ? f.acl_max.y : a3.y);
  const float2 v23 = add2(f.vel, mul2(a23, bc2(ts), nz));
  const float2 dtn = add2(f.dt, bc2(ts));
  const bool   p23x = p2x || p3x, p23y = p2y || p3y;
  f.acl = make_float2(p1x ? a1.x : (p23x ? a23.x : 0.0f), p1y ? a1.y : (p23y ? a23.y : 0.0f));
  f.vel = make_float2(p1x ? v1.x : (p23x ? v23.x : f.vel_tgt.x), p1y ? v1.y : (p23y ? v23.y : f.vel_tgt.y));
  f.dt  = make_float2((p1x || p23x) ? dtn.x : f.dt.x, (p1y || p23y) ? dtn.y : f.dt.y);
  return f.vel;
}

template <int DA, int DB>
RK_DEV void to_fast_wheel2(FastWheel2 &w, const Veh &v, int a, int b, float b0) {
  w.rpm[0] = v.m[a].p_rpm, w.rpm[1] = v.m[b].p_rpm, w.cur[0] = v.m[a].cur_tgt, w.cur[1] = v.m[b].cur_tgt;
  w.dsum[0] = 0, w.dsum[1] = 0;
  w.prev_val = make_float2(v.c[a].prev_val, v.c[b].prev_val), w.integ = make_float2(v.c[a].integ, v.c[b].integ);
  w.lpf_y = make_float2(v.c[a].lpf_y, v.c[b].lpf_y), w.lpf_x = make_float2(v.c[a].lpf_x, v.c[b].lpf_x);
  w.b0x = make_float2(fmul(b0, v.c[a].lpf_x), fmul(b0, v.c[b].lpf_x));
}
RK_DEV void to_fast2(const Veh &v, FastVeh2 &f, float ts, float b0) {
  f.p = make_float2(v.pos[0], v.pos[1]);
  fast_interp2_load(f.xy, v.it[0], v.it[1], ts);
  fast_interp_load(f.th, v.it[2], ts);
  to_fast_wheel2<1, 1>(f.w01, v, 0, 1, b0);
  to_fast_wheel2<-1, -1>(f.w23, v, 2, 3, b0);
}
RK_DEV FastWheel lane_wheel(const FastWheel2 &w, int l) {
  FastWheel s;
  s.rpm = w.rpm[l], s.cur = w.cur[l], s.dsum = w.dsum[l];
  s.prev_val = l ? w.prev_val.y : w.prev_val.x, s.integ = l ? w.integ.y : w.integ.x;
  s.lpf_y = l ? w.lpf_y.y : w.lpf_y.x, s.lpf_x = l ? w.lpf_x.y : w.lpf_x.x, s.b0x = l ? w.b0x.y : w.b0x.x;
  return s;
}
RK_DEV void from_fast2_common(Veh &v, const FastVeh2 &f, FastWheel w[4]) {
  v.pos[0] = f.p.x, v.pos[1] = f.p.y;
  fast_interp2_store(f.xy, v.it[0], v.it[1]);
  fast_interp_store(f.th, v.it[2]);
  w[0] = lane_wheel(f.w01, 0), w[1] = lane_wheel(f.w01, 1), w[2] = lane_wheel(f.w23, 0), w[3] = lane_wheel(f.w23, 1);
#pragma unroll
  for(int k = 0; k < 4; k++) v.c[k].prev_val = w[k].prev_val, v.c[k].integ = w[k].integ, v.c[k].lpf_y = w[k].lpf_y, v.c[k].lpf_x = w[k].lpf_x;
}
template <int D0, int D1, int D2, int D3>
RK_DEV void from_fast2(Veh &v, const FastVeh2 &f, int nticks) {
  if(nticks <= 0) return;
  FastWheel w[4];
  from_fast2_common(v, f, w);
  from_fast_motor<D0>(v.m[0], w[0], nticks);
  from_fast_motor<D1>(v.m[1], w[1], nticks);
  from_fast_motor<D2>(v.m[2], w[2], nticks);
  from_fast_motor<D3>(v.m[3], w[3], nticks);
}

// ---- RK_SENSOR_STREAM on the fast tick: the wheel feedback comes from recorded C610 frames, so rx_callback
// (VD_motor_if_m2006.cpp:32-72) runs in full -- big-endian fields, direction, the +-4096 unwrap -- instead of the
// collapsed plant.  The odometry product fma(d, K_hi, d*K_lo) is proven for every step the unwrap can produce
// (|d| <= 2^17, rk_exact.cu), so no frame can leave the fast path's domain.
struct StreamSense {
  int32_t ang[4]; // head Status s16_rawAngle (direction applied): the one status word the next frame's unwrap needs.
                  // s16_rawSpeedRpm / s16_rawCurr are rewritten by the next rx_callback before anything reads them, and
                  // every launch ends on a transcription tick, so they are not carried through the fast ticks.
};
RK_DEV void stream_sense_load(StreamSense &ss, const Veh &v) {
#pragma unroll
  for(int k = 0; k < 4; k++) ss.ang[k] = v.m[k].ang;
}
template <int DIR>
RK_DEV float2 fast_wheel_rx2(StreamSense &ss, int k, int32_t &dsum, uint64_t frame, const FastConsts &fc) {
  const uint32_t lo = (uint32_t)frame;
  const int32_t  a = sext16((int32_t)(__byte_perm(lo, 0, 0x4401))); // (b0<<8)|b1
  const int32_t  r = sext16((int32_t)(__byte_perm(lo, 0, 0x4423))); // (b2<<8)|b3
  const int32_t  raw_ang = (DIR == 1) ? a : sext16(8192 - a);
  int32_t        d       = sext16(raw_ang - ss.ang[k]);
  d                      = (d > 4096) ? sext16(d - 8192) : ((d < -4096) ? sext16(d + 8192) : d);
  dsum += d;
  ss.ang[k] = raw_ang;
  const float df = (float)d;
  return make_float2(fmul(fmul((float)sext16(r * DIR), RK_RPM_TO_RADPS), RK_GEAR_RATIO_INV), __fmaf_rn(df, fc.k_hi, fmul(df, fc.k_lo)));
}
RK_DEV void from_fast2_stream(Veh &v, const FastVeh2 &f, const StreamSense &ss, int nticks) {
  if(nticks <= 0) return;
  FastWheel w[4];
  from_fast2_common(v, f, w);
#pragma unroll
  for(int k = 0; k < 4; k++) {
    Motor &m = v.m[k];
    m.sum += (int64_t)w[k].dsum;
    m.prev = m.sum;
    m.ang = ss.ang[k]; // rpm / cur: see StreamSense
    m.cur_tgt = w[k].cur;
    m.head    = (m.head + nticks) % 3;
  }
}

// plant step + collapsed rx_callback for one wheel; returns {Mvel, Mrad} as one lane pair
template <int DIR>
RK_DEV float2 fast_wheel_sense2(int32_t &rpm, int32_t cur, int32_t &dsum, const FastConsts &fc) {
  rpm += ((cur * 4 - rpm) >> 4);
  const int32_t rw   = (DIR == 1) ? rpm : -rpm;
  const int32_t dang = plant_dang(rw);
  dsum += dang;
  const float df = (float)dang;
  return make_float2(fmul(fmul((float)rw, RK_RPM_TO_RADPS), RK_GEAR_RATIO_INV), __fmaf_rn(df, fc.k_hi, fmul(df, fc.k_lo)));
}

// FF_PI_D::update + set_CurrA_tgt for a pair of wheels of direction DIR (see fast_wheel_ctrl)
template <int DIR, bool FFSAT>
RK_DEV void fast_wheel_ctrl2(FastWheel2 &w, const rk_vdt_params_t &p, const FastConsts &fc, float2 mtgt, float mvel_a,
                             float mvel_b, float nz) {
  const float2 tgt = mul2(mtgt, bc2(RK_GEAR_RATIO), nz);
  const float2 now = make_float2(fmul(mvel_a, RK_GEAR_RATIO), fmul(mvel_b, RK_GEAR_RATIO));
  const float2 err = sub2(tgt, now);
  const float2 x   = mul2(sub2(now, w.prev_val), bc2(p.ctrl_freq), nz);
  const float2 b0x = mul2(x, bc2(fc.B0), nz);
  const float2 y   = add2(add2(mul2(w.lpf_y, bc2(fc.A1), nz), b0x), w.b0x);
  w.lpf_y = y, w.lpf_x = x, w.b0x = b0x;
  const float2 ig = add2(w.integ, mul2(err, bc2(fc.ki_dt), nz));
  w.integ         = make_float2(clamp_sym(ig.x, p.i_limit, fc.neg_i_limit), clamp_sym(ig.y, p.i_limit, fc.neg_i_limit));
  float2 u        = sub2(add2(mul2(err, bc2(p.kp), nz), w.integ), mul2(y, bc2(p.kd), nz));
  w.prev_val      = now;
  float2 ff;
  if(FFSAT) {
    const float2 sp = make_float2(__saturatef(fmul(tgt.x, p.kff)), __saturatef(fmul(tgt.y, p.kff)));
    const float2 sn = make_float2(__saturatef(fmul(tgt.x, -p.kff)), __saturatef(fmul(tgt.y, -p.kff)));
    ff              = sub2(sp, sn);
  } else {
    const float2 m = mul2(tgt, bc2(p.kff), nz);
    ff             = make_float2(clamp_sym(m.x, p.ff_limit, fc.neg_ff_limit), clamp_sym(m.y, p.ff_limit, fc.neg_ff_limit));
  }
  u               = add2(u, ff);
  const float2 tq = mul2(u, bc2(RK_AMPERE_TO_RAW_CURR), nz);
  int32_t      ta = __float2int_rz(tq.x), tb = __float2int_rz(tq.y);
  ta = sext16(DIR > 0 ? ta : -ta), tb = sext16(DIR > 0 ? tb : -tb);
  w.cur[0] = min(max(ta, -p.raw_curr_lim), p.raw_curr_lim);
  w.cur[1] = min(max(tb, -p.raw_curr_lim), p.raw_curr_lim);
}

// One packed fast tick.  cs = {cos, sin}(yaw), sc = {sin, cos}(yaw).
template <bool FFSAT>
RK_DEV void fast_tick2_core(FastVeh2 &f, const rk_vdt_params_t &p, const FastConsts &fc, float2 cs, float2 sc, float nz,
                            float vel[3], float tgt[3], float2 m0, float2 m1, float2 m2, float2 m3);
template <bool FFSAT>
RK_DEV void fast_tick2(FastVeh2 &f, const rk_vdt_params_t &p, const FastConsts &fc, float2 cs, float2 sc, float nz,
                       float vel[3], float tgt[3]) {
  const float2 m0 = fast_wheel_sense2<1>(f.w01.rpm[0], f.w01.cur[0], f.w01.dsum[0], fc);
  const float2 m1 = fast_wheel_sense2<1>(f.w01.rpm[1], f.w01.cur[1], f.w01.dsum[1], fc);
  const float2 m2 = fast_wheel_sense2<-1>(f.w23.rpm[0], f.w23.cur[0], f.w23.dsum[0], fc);
  const float2 m3 = fast_wheel_sense2<-1>(f.w23.rpm[1], f.w23.cur[1], f.w23.dsum[1], fc);
  fast_tick2_core<FFSAT>(f, p, fc, cs, sc, nz, vel, tgt, m0, m1, m2, m3);
}
// the same tick fed by four recorded frames (RK_SENSOR_STREAM)
template <bool FFSAT>
RK_DEV void fast_tick2_stream(FastVeh2 &f, StreamSense &ss, const uint64_t fr[4], const rk_vdt_params_t &p, const FastConsts &fc,
                              float2 cs, float2 sc, float nz, float vel[3], float tgt[3]) {
  const float2 m0 = fast_wheel_rx2<1>(ss, 0, f.w01.dsum[0], fr[0], fc);
  const float2 m1 = fast_wheel_rx2<1>(ss, 1, f.w01.dsum[1], fr[1], fc);
  const float2 m2 = fast_wheel_rx2<-1>(ss, 2, f.w23.dsum[0], fr[2], fc);
  const float2 m3 = fast_wheel_rx2<-1>(ss, 3, f.w23.dsum[1], fr[3], fc);
  fast_tick2_core<FFSAT>(f, p, fc, cs, sc, nz, vel, tgt, m0, m1, m2, m3);
}
template <bool FFSAT>
RK_DEV void fast_tick2_core(FastVeh2 &f, const rk_vdt_params_t &p, const FastConsts &fc, float2 cs, float2 sc, float nz,
                            float vel[3], float tgt[3], float2 m0, float2 m1, float2 m2, float2 m3) {
  // conv_Mdir_to_Vdir on both paths at once  VD_vehicle_controller.cpp:126-130 (called at :26 and :42)
  // (sum * 0.25f) * R == sum * (0.25f * R) bit for bit: scaling by 2^-2 is exact in both places as long as
  // nothing underflows, and a sum of wheel speeds / angle steps is 0 or >= 2^-40 in magnitude (they derive
  // from int16 rpm and integer encoder steps).  The same scaling commutes with the two exact divisions.
  const float  R = p.wheel_radius_mm, qR = fmul(0.25f, R);
  const float2 vx = mul2(add2(add2(add2(m0, m1), m2), m3), bc2(qR), nz);       // {vel.x, local dx}
  const float2 vy = mul2(add2(sub2(add2(neg2(m0), m1), m2), m3), bc2(qR), nz); // {vel.y, local dy}
  {
    const float s = fadd(fadd(fsub(-m0.x, m1.x), m2.x), m3.x);
    vel[2]        = fmul(div_const(div_const(s, p.sqrtf2, fc.rcp_s2), p.wheel_l_mm, fc.rcp_l), qR);
  }
  vel[0] = vx.x, vel[1] = vy.x;
  // pos += (R(yaw) * local) * 0.001   :50-51 ; {lx*c - ly*s, lx*s + ly*c}
  const float2 a = mul2(bc2(vx.y), cs, nz);
  const float2 b = mul2(bc2(vy.y), sc, nz);
  f.p            = add2(f.p, mul2(__fadd2_rn(a, make_float2(-b.x, b.y)), bc2(0.001f), nz));
  const float2 txy = fast_interp2_update(f.xy, p.ts, nz);
  tgt[0] = txy.x, tgt[1] = txy.y;
  tgt[2] = fast_interp_update(f.th, p.ts);
  // conv_Vdir_to_Mdir  :113-118
  const float  T  = fmul(fmul(fc.s2l, tgt[2]), 4.0f); // (not folded: a denormal-range target would round differently)
  const float2 xy = make_float2(fsub(tgt[0], tgt[1]), fadd(tgt[0], tgt[1]));
  const float2 M01 = div_const2(sub2(xy, bc2(T)), R, fc.rcp_r, nz);
  const float2 M23 = div_const2(add2(xy, bc2(T)), R, fc.rcp_r, nz);
  fast_wheel_ctrl2<1, FFSAT>(f.w01, p, fc, M01, m0.x, m1.x, nz);
  fast_wheel_ctrl2<-1, FFSAT>(f.w23, p, fc, M23, m2.x, m3.x, nz);
}

} // namespace rk
