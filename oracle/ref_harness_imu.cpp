/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * C-ABI harness around the UNMODIFIED reference IMU sources
 *   src/Imu/imu_if_wt901c.{hpp,cpp}, src/Imu/imu_if_base.hpp, lib/wt901c/wit_c_sdk.{c,h}, REG.h
 * compiled where they lie under /root/reference (oracle/Makefile -> oracle/_ref/libref_imu.so).
 * Sensor samples enter exactly as on the robot: as WT901 serial frames (0x55, type, 4 x int16
 * little-endian, 8-bit checksum -- lib/wt901c/wit_c_sdk.c:132-164) pushed into the fake
 * Serial6 FIFO, drained by IMU_IF_WT901C::isComComp() through WitSerialDataIn().
 *
 * The WIT SDK keeps its register file and parser state in globals (sReg[], s_cDataUpdate),
 * so this library holds ONE live IMU at a time; batches run instance after instance.
 */
#include <new>
#include <stdlib.h>
#include <string.h>

#include "Imu/imu_if_wt901c.hpp"
extern "C" {
#include <wit_c_sdk.h>
}
#include "robotick.h"

HardwareSerial Serial6;
HardwareSerial Serial7;
/* isComComp() reads the timer on entry (imu_if_wt901c.cpp:133).  While init() spins in getDataImmediately() waiting for
 * its first quaternion frame, that read is where the harness lets time pass: with the UART drained, the bytes of the next
 * update slot "arrive" (see ref_imt_bytes_rollout).  A stream that ends before the frame would spin forever: thrown out. */
namespace {
struct InitFeed {
  bool            active = false;
  const uint32_t *cells  = nullptr;
  const uint16_t *nbytes = nullptr;
  int64_t         n = 0, i = 0;
  int             K = 0, ncells = 0, next_u = 0;
} g_feed;
struct InitStarved {};
void feed_slot(int u) {
  int nb = 16 * g_feed.ncells;
  if(g_feed.nbytes && g_feed.nbytes[(int64_t)u * g_feed.n + g_feed.i] < nb) nb = g_feed.nbytes[(int64_t)u * g_feed.n + g_feed.i];
  for(int c = 0; 16 * c < nb; c++) {
    const uint32_t *cell = g_feed.cells + (((int64_t)u * g_feed.ncells + c) * g_feed.n + g_feed.i) * 4;
    Serial6.feed((const uint8_t *)cell, nb - 16 * c < 16 ? nb - 16 * c : 16);
  }
}
} // namespace
uint32_t get_gptimer_cnt() {
  if(g_feed.active && Serial6.available() == 0) {
    if(g_feed.next_u >= g_feed.K) throw InitStarved();
    feed_slot(g_feed.next_u++);
  }
  return 0;
}
namespace DEBUG {
char EXT_PRINT_BUF[1024];
void print(char *, uint32_t) {}
void record_proc_load(uint8_t, uint8_t) {}
} // namespace DEBUG
namespace LGT {
void push_buffer(char *, uint32_t) {}
} // namespace LGT

namespace {

using IMT::IMU_IF_WT901C;

void push_frame(uint8_t type, const int16_t w[4]) {
  uint8_t f[11];
  f[0] = 0x55, f[1] = type;
  for(int k = 0; k < 4; k++) f[2 + 2 * k] = (uint8_t)(w[k] & 0xFF), f[3 + 2 * k] = (uint8_t)((w[k] >> 8) & 0xFF);
  uint8_t s = 0;
  for(int k = 0; k < 10; k++) s += f[k];
  f[10] = s;
  Serial6.feed(f, 11);
}

/* regs: AX AY AZ GX GY GZ HX HY HZ Roll Pitch Yaw q0 q1 q2 q3 (robotick.h RK_IMT_REG_*) */
void push_sample(const int16_t r[16], int have_quat) {
  const int16_t acc[4] = {r[0], r[1], r[2], 0}, gyr[4] = {r[3], r[4], r[5], 0}, mag[4] = {r[6], r[7], r[8], 0};
  const int16_t ang[4] = {r[9], r[10], r[11], 0}, qut[4] = {r[12], r[13], r[14], r[15]};
  push_frame(WIT_ACC, acc);
  push_frame(WIT_GYRO, gyr);
  push_frame(WIT_ANGLE, ang);
  push_frame(WIT_MAGNETIC, mag);
  if(have_quat) push_frame(WIT_QUATER, qut);
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

void export_state(IMU_IF_WT901C *m, uint32_t *w) {
  memset(w, 0, 4 * RK_IS_WORDS);
  for(int k = 0; k < 4; k++) w[RK_IS_QINIT + k] = f2u(m->q_init[k]);
  const IMT::IMU_IF::Data &d = m->d_buf[m->u8_d_buf_read_page];
  memcpy(&w[RK_IS_DATA], &d, 16 * 4);
  w[RK_IS_FLAGS] = m->is_error ? RK_IS_FLAG_ERROR : 0u;
}
void import_state(IMU_IF_WT901C *m, const uint32_t *w) {
  for(int k = 0; k < 4; k++) m->q_init[k] = u2f(w[RK_IS_QINIT + k]);
  memset(m->d_buf, 0, sizeof(m->d_buf));
  m->u8_d_buf_read_page = 0;
  memcpy(&m->d_buf[0], &w[RK_IS_DATA], 16 * 4);
  m->is_error = (w[RK_IS_FLAGS] & RK_IS_FLAG_ERROR) != 0;
}

IMU_IF_WT901C *make() {
  void *mem = calloc(1, sizeof(IMU_IF_WT901C));
  return new(mem) IMU_IF_WT901C();
}

inline uint32_t &soa(uint32_t *blk, int64_t n, int64_t i, int w) { return blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }

} // namespace

extern "C" {

void *ref_imt_create(void) {
  memset(sReg, 0, sizeof(int16_t) * REGSIZE);
  return make();
}
void ref_imt_destroy(void *h) {
  if(!h) return;
  ((IMU_IF_WT901C *)h)->~IMU_IF_WT901C();
  free(h);
}
/* IMU_IF_WT901C::init()  imu_if_wt901c.cpp:63-77 with the boot-time sample already on the wire */
void ref_imt_init(void *h, const int16_t regs[16]) {
  push_sample(regs, 1);
  ((IMU_IF_WT901C *)h)->init();
}
/* IMU_IF_WT901C::update()  :83-89 */
void ref_imt_update(void *h, const int16_t regs[16], int have_quat) {
  push_sample(regs, have_quat);
  ((IMU_IF_WT901C *)h)->update();
}
float ref_imt_yaw(void *h) { return ((IMU_IF_WT901C *)h)->getYawDate(); }
int   ref_imt_is_error(void *h) { return ((IMU_IF_WT901C *)h)->isError() ? 1 : 0; }
void  ref_imt_get(void *h, float out[16]) {
  IMT::IMU_IF::Data d;
  ((IMU_IF_WT901C *)h)->getDataLatest(d);
  memcpy(out, &d, 64);
}
void ref_imt_export(void *h, uint32_t *words) { export_state((IMU_IF_WT901C *)h, words); }
void ref_imt_import(void *h, const uint32_t *words) { import_state((IMU_IF_WT901C *)h, words); }

/* Same contract as rk_imt_update() on HOST arrays: K updates of instances [i0,i1).
 * regs: int16 in two 128-bit cells per sample (register r of sample u, instance i at ((u*2 + r/8)*n + i)*8 + r%8);
 * have_quat: uint8 [K][n] or NULL; out: Data planes [K][4][n] float4 or NULL; do_init: run init() with the
 * first sample instead of update(). */
void ref_imt_rollout(uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const int16_t *regs,
                     const uint8_t *have_quat, uint32_t *out, int do_init) {
  for(int64_t i = i0; i < i1; i++) {
    memset(sReg, 0, sizeof(int16_t) * REGSIZE);
    IMU_IF_WT901C *m = make();
    uint32_t       w[RK_IS_WORDS];
    if(state) {
      for(int k = 0; k < RK_IS_WORDS; k++) w[k] = soa(state, n, i, k);
      import_state(m, w);
    }
    for(int u = 0; u < K; u++) {
      int16_t r[16];
      for(int k = 0; k < 16; k++) r[k] = regs[(((int64_t)u * 2 + k / 8) * n + i) * 8 + k % 8];
      int hq = have_quat ? have_quat[(int64_t)u * n + i] : 1;
      if(do_init && u == 0) {
        push_sample(r, 1);
        m->init();
      } else {
        push_sample(r, hq);
        m->update();
      }
      if(out) {
        IMT::IMU_IF::Data d;
        m->getDataLatest(d);
        const uint32_t *dw = (const uint32_t *)&d;
        for(int k = 0; k < 16; k++) out[(((int64_t)u * 4 + k / 4) * n + i) * 4 + (k % 4)] = dw[k];
      }
    }
    if(state) {
      export_state(m, w);
      for(int k = 0; k < RK_IS_WORDS; k++) soa(state, n, i, k) = w[k];
    }
    m->~IMU_IF_WT901C();
    free(m);
  }
}

/* Byte-stream path (rk_imt_feed_bytes): the serial bytes go to the UART stub unframed and the vendor parser
 * (wit_c_sdk.c, compiled as it is) does the rest.  The parser's window and flags are file-static in the
 * reference, so every instance is replayed from power-on: update 0 is init(), the following K-1 are update().
 * cells / nbytes: the rk_imt_feed_bytes wire layout (128-bit cells; nbytes NULL = full slots).  Outputs as
 * ref_imt_rollout; final sReg (16 tracked registers) into sreg_out[16*i ..]. */
void ref_imt_bytes_rollout(uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, int ncells, const uint32_t *cells,
                           const uint16_t *nbytes, uint32_t *out, int16_t *sreg_out) {
  for(int64_t i = i0; i < i1; i++) {
    memset(sReg, 0, sizeof(int16_t) * REGSIZE);
    IMU_IF_WT901C *m = make();
    g_feed.cells = cells, g_feed.nbytes = nbytes, g_feed.n = n, g_feed.i = i, g_feed.K = K, g_feed.ncells = ncells;
    auto publish = [&](int u) {
      if(!out) return;
      IMT::IMU_IF::Data d;
      m->getDataLatest(d);
      const uint32_t *dw = (const uint32_t *)&d;
      for(int k = 0; k < 16; k++) out[(((int64_t)u * 4 + k / 4) * n + i) * 4 + (k % 4)] = dw[k];
    };
    int u = 0;
    if(K > 0) { /* update slot 0 is init(); it takes as many further slots as its wait for a quaternion frame needs */
      for(int v = 0; v < K; v++) publish(v); /* slots swallowed by the wait publish nothing new: the power-on page */
      feed_slot(0);
      g_feed.next_u = 1, g_feed.active = true;
      bool done = true;
      try {
        m->init();
      } catch(const InitStarved &) {
        done = false; /* the stream ended inside the wait */
      }
      g_feed.active = false;
      u = g_feed.next_u;
      if(done) publish(u - 1);
    }
    for(; u < K; u++) {
      feed_slot(u);
      m->update();
      publish(u);
    }
    if(state) {
      uint32_t w[RK_IS_WORDS];
      export_state(m, w);
      for(int k = 0; k < RK_IS_WORDS; k++) soa(state, n, i, k) = w[k];
    }
    if(sreg_out) {
      for(int k = 0; k < 12; k++) sreg_out[16 * i + k] = sReg[AX + k];
      for(int k = 0; k < 4; k++) sreg_out[16 * i + 12 + k] = sReg[q0 + k];
    }
    m->~IMU_IF_WT901C();
    free(m);
  }
}
}
