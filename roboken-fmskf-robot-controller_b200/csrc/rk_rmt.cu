// rk_rmt.cu -- the RobotManager's vehicle-management block (routine_ros(), src/RobotManager/RM_task_main.cpp:484-767)
// batched: ROS command in, floor / wall sensors in, the message VDT receives out.
//
// Streaming, HBM-bound: one thread per robot, the four manager state words in registers across the K fused
// cycles; per cycle three 128-bit loads (48 B record) and one 128-bit + one 32-bit store.  The table arctangent
// UTIL::mymath::atanf (src/Utility/util_mymath.cpp:98-115) is staged in shared memory once per CTA (per-lane
// indices differ, a __constant__ bank would serialise).
#include <string.h>

#include "rk_common.cuh"
#include "rk_math.cuh"

namespace rk {

#include "atan_table.inc"
__device__ const uint32_t g_atan_table[626]  = {RK_ATAN_TABLE_BITS};
__device__ const uint32_t g_atan_delimit[27] = {RK_ATAN_DELIMIT_BITS};
__device__ const uint32_t g_atan_width[26]   = {RK_ATAN_WIDTH_BITS};

struct AtanTab {
  float table[626], delimit[27], width[26];
};
RK_DEV void stage_atan(AtanTab &s) {
  for(int k = threadIdx.x; k < 626; k += blockDim.x) s.table[k] = u2f(g_atan_table[k]);
  for(int k = threadIdx.x; k < 27; k += blockDim.x) s.delimit[k] = u2f(g_atan_delimit[k]);
  for(int k = threadIdx.x; k < 26; k += blockDim.x) s.width[k] = u2f(g_atan_width[k]);
  __syncthreads();
}
// mymath::atanf :98-115 (the recursion on negative x unrolled: atanf(-x) = -atanf(x))
RK_DEV float my_atanf(const AtanTab &s, float x) {
  const bool neg = x < 0.0f;
  if(neg) x = -x;
  float r;
  if(x == 0.0f) {
    r = 0.0f;
  } else if(x <= s.delimit[26]) { // first i in 1..26 with x <= delimit[i]
    int i = 1, hi = 26; // the delimiters ascend, so the firmware's linear scan is a lower bound: 5 halvings of 26
#pragma unroll
    for(int st = 0; st < 5; st++) {
      const int  mid = (i + hi) >> 1;
      const bool le  = x <= s.delimit[mid];
      hi = le ? mid : hi, i = le ? i : mid + 1;
    }
    const float index = fadd((float)(24 * (i - 1)), fdiv(fsub(x, s.delimit[i - 1]), s.width[i - 1]));
    const int   ii    = __float2int_rz(index);
    const float dec   = fsub(index, (float)ii);
    r = fadd(s.table[ii], fmul(dec, fsub(s.table[ii + 1], s.table[ii])));
  } else {
    r = s.table[577 - 1]; // TABLE_SIXE_ATAN is 577 although the table has 625 entries (:6,114); NaN lands here too
  }
  return neg ? -r : r;
}
// mymath::atan2f :117-126
// The three quadrant cases share ONE evaluation of atanf(y / x): lanes of a warp that sit in different quadrants would
// otherwise run the table search three times over.
RK_DEV float my_atan2f(const AtanTab &s, float y, float x) {
  const bool q1 = x > 0.0f, q2 = y >= 0.0f && x < 0.0f, q3 = y < 0.0f && x < 0.0f;
  if(q1 || q2 || q3) {
    const float a = my_atanf(s, fdiv(y, x));
    return q1 ? a : (q2 ? fadd(a, RK_PI) : fsub(a, RK_PI));
  }
  if(y > 0.0f && x == 0.0f) return (float)((double)RK_PI / 2.0);
  if(y < 0.0f && x == 0.0f) return (float)(-(double)RK_PI / 2.0);
  return 0.0f;
}

__global__ void __launch_bounds__(128) atan2f_kernel(const float *__restrict__ y, const float *__restrict__ x, float *__restrict__ out, int64_t n) {
  __shared__ AtanTab s;
  stage_atan(s);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = my_atan2f(s, y[i], x[i]);
}

RK_DEV double u2d(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

enum : uint32_t { MSG_NONE = 0u, MSG_DIR = 1u, MSG_CONT = 2u }; // VDT::MSG_ID  VD_task_main.hpp:8-12
enum : uint32_t { FLOOR = 1u, WALL = 2u };                      // FD_task_main.hpp:20-22

__global__ void __launch_bounds__(128)
rmt_guard_kernel(const rk_rmt_params_t p, uint4 *__restrict__ state, int64_t n, int K, const uint4 *__restrict__ in,
                 uint4 *__restrict__ cmd_out, uint32_t *__restrict__ abort_out) {
  __shared__ AtanTab s;
  stage_atan(s);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  uint4    st = __ldcs(state + i);
  uint32_t cmd_status = st.x, ignore = st.y, no_cmd = st.z, abort_v = st.w;
  const uint4 *src = in + i;
  uint4 c0 = K > 0 ? __ldcs(src) : make_uint4(0u, 0u, 0u, 0u), c1 = K > 0 ? __ldcs(src + n) : c0, c2 = K > 0 ? __ldcs(src + 2 * n) : c0;
  for(int u = 0; u < K; u++) {
    const uint4 a = c0, b = c1, c = c2; // words 0-3, 4-7, 8-11
    if(u + 1 < K) { // next record in flight while this one is decided
      const uint4 *nx = src + (int64_t)(u + 1) * 3 * n;
      c0 = __ldcs(nx), c1 = __ldcs(nx + n), c2 = __ldcs(nx + 2 * n);
    }
    // rclc_executor_spin_some(): the subscription callbacks  :159-248.  The record kind differs from lane to lane,
    // so everything but the double-precision conversions and the arctangent is selects, not branches.
    const uint32_t kind = a.x;
    const bool k_cmd = kind == RK_ROS_MECANUM_CMD, k_cont = kind == RK_ROS_MECANUM_CONT, k_vel = kind == RK_ROS_CMD_VEL;
    const bool k_command = kind == RK_ROS_COMMAND; // a Command always stops the vehicle, then switches the manager's mode :162-201
    const bool updated = k_cmd || k_cont || k_vel || k_command;
    uint32_t   id = (k_cont || k_vel) ? MSG_CONT : MSG_DIR; // not updated: the idle message MOVE_DIR / MOVE_STOP / 0 / 0
    uint32_t   m_cmd = k_cmd ? a.y : (uint32_t)RK_DIR_MOVE_STOP, m_speed = k_cmd ? a.w : 0u;
    uint32_t   m_time = k_cmd ? a.z : (k_cont ? a.y : (k_vel ? 500u : (k_command ? 1u : 0u)));
    float      vx = 0.0f, vy = 0.0f, vth = 0.0f;
    if(k_cont || k_vel) {
      double dx = u2d(b.x, b.y), dy = u2d(b.z, b.w);
      if(k_vel) dx = __dmul_rn(dx, 1000.0), dy = __dmul_rn(dy, 1000.0);
      vx = __double2float_rn(dx), vy = __double2float_rn(dy), vth = __double2float_rn(u2d(c.x, c.y));
    }
    {
      const bool known = a.y == 0u || a.y == 1u || a.y == 2u || a.y == 4u || a.y == 10u; // the rest: UNKNOWN_CMD
      cmd_status = k_command ? (known ? a.y : 0xFFu) : cmd_status;
      ignore     = (k_command && a.y == 10u) ? (ignore ? 0u : 1u) : ignore; // SWITCH_FLOOR_SENSOR
    }
    bool exist = updated; // :484-505
    abort_v    = updated ? 0u : abort_v;
    // floor sensors :507-542 -- bytes rForward, lForward, rBack, lBack | right, left, forward, back
    uint32_t f0 = c.z, f1 = c.w;
    {
      const uint32_t nf = __popc(__vcmpeq4(f0, 0u) & 0x01010101u) + __popc(__vcmpeq4(f1, 0u) & 0x01010101u);
      const uint32_t nw = __popc(__vcmpeq4(f0, 0x02020202u) & 0x01010101u) + __popc(__vcmpeq4(f1, 0x02020202u) & 0x01010101u);
      const bool     all_floor = nf >= 5u || nw >= 5u || ignore != 0u;
      f0 = all_floor ? 0x01010101u : f0, f1 = all_floor ? 0x01010101u : f1;
    }
    const uint32_t rF = f0 & 0xFFu, lF = (f0 >> 8) & 0xFFu, rB = (f0 >> 16) & 0xFFu, lB = f0 >> 24;
    const uint32_t right = f1 & 0xFFu, left = (f1 >> 8) & 0xFFu, fwd = (f1 >> 16) & 0xFFu, back = f1 >> 24;
    { // MOVE_START: leave a wall (an opponent)  :546-577 -- the first of forward, back, left, right that sees one
      const bool     w_f = fwd == WALL, w_b = back == WALL, w_l = left == WALL, w_r = right == WALL;
      const uint32_t dir = w_f ? RK_DIR_GO_BACK : (w_b ? RK_DIR_GO_FORWARD : (w_l ? RK_DIR_GO_RIGHT : (w_r ? RK_DIR_GO_LEFT : 0u)));
      const uint32_t bit = w_f ? 1u : (w_b ? 2u : (w_l ? 4u : (w_r ? 8u : 0u)));
      const bool     leave = cmd_status == 2u && dir != 0u;
      abort_v |= leave ? bit : 0u;
      id = leave ? (uint32_t)MSG_DIR : id, m_cmd = leave ? dir : m_cmd, m_time = leave ? p.wall_leave_time_ms : m_time;
      m_speed = leave ? p.wall_leave_speed_mmps : m_speed, exist = exist || leave;
    }
    { // direction commands: the sensor a direction needs and the abort bits it raises  :581-673
      // GO_FORWARD..GO_LEFT_BACK (1..8) -> sensor byte index 6,7,4,5,0,1,2,3 ; fllr_abort bits 8..11 as a nibble
      const uint32_t k    = (m_cmd - 1u) & 7u;
      const uint32_t sidx = (0x32105476u >> (4u * k)) & 7u;
      const uint32_t need = (((sidx & 4u) ? f1 : f0) >> (8u * (sidx & 3u))) & 0xFFu;
      const uint32_t bits = ((0x6A594821u >> (4u * k)) & 0xFu) << 8;
      const bool     veto = id == MSG_DIR && m_cmd >= 1u && m_cmd <= 8u && need != FLOOR;
      m_cmd = veto ? (uint32_t)RK_DIR_MOVE_STOP : m_cmd, m_time = veto ? 1u : m_time, m_speed = veto ? 0u : m_speed;
      exist = exist || veto, abort_v |= veto ? bits : 0u;
    }
    if(id == MSG_CONT) { // REQ_MOVE_CONT_DIR :674-749
      const float ax = vx < 0.0f ? -vx : vx, ay = vy < 0.0f ? -vy : vy;
      if(!(ax < 0.01f && ay < 0.01f)) {
        const float vph = my_atan2f(s, vy, vx);
        const float P   = 3.1415f;
        bool        veto = false;
        veto |= fwd != FLOOR && (fmul(-P, 0.33f) < vph && vph <= fmul(P, 0.33f));
        veto |= back != FLOOR && (fmul(P, 0.66f) < vph || vph <= fmul(-P, 0.66f));
        veto |= left != FLOOR && (fmul(P, 0.16f) < vph && vph <= fmul(P, 0.84f));
        veto |= right != FLOOR && (fmul(-P, 0.84f) < vph && vph <= fmul(-P, 0.16f));
        veto |= rB != FLOOR && (fmul(P, 0.92f) < vph || vph <= fmul(-P, 0.42f));
        veto |= rF != FLOOR && (fmul(-P, 0.58f) < vph && vph <= fmul(P, 0.08f));
        veto |= lF != FLOOR && (fmul(-P, 0.08f) < vph && vph <= fmul(P, 0.58f));
        veto |= lB != FLOOR && (fmul(P, 0.42f) < vph || vph <= fmul(-P, 0.92f));
        if(veto) vx = 0.0f, vy = 0.0f, abort_v |= 1u << 16;
      }
    }
    bool sent = exist; // :752-767
    if(exist) no_cmd = 0u;
    else no_cmd++;
    if(no_cmd > p.no_cmd_stop_thre) {
      id = MSG_DIR, m_cmd = RK_DIR_MOVE_STOP, m_time = 1u, m_speed = 0u, sent = true;
      no_cmd = 0u;
    }
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if(sent) {
      if(id == MSG_DIR) rec = make_uint4(m_cmd, m_speed, 0u, (uint32_t)RK_CMD_MSG_MOVE_DIR | (m_time << 8));
      else rec = make_uint4(f2u(vx), f2u(vy), f2u(vth), (uint32_t)RK_CMD_MSG_MOVE_CONT_DIR | (m_time << 8));
    }
    __stcs(cmd_out + (int64_t)u * n + i, rec);
    if(abort_out) __stcs(abort_out + (int64_t)u * n + i, abort_v);
  }
  __stcs(state + i, make_uint4(cmd_status, ignore, no_cmd, abort_v));
}

} // namespace rk

using namespace rk;

extern "C" {

void rk_rmt_default_params(rk_rmt_params_t *p) { // RM_task_main.cpp:62-64
  if(!p) return;
  p->no_cmd_stop_thre = 200u, p->wall_leave_time_ms = 200u, p->wall_leave_speed_mmps = 100u;
}
size_t rk_rmt_state_words(void) { return RK_RS_WORDS; }
size_t rk_rmt_state_bytes(int64_t n) { return n < 0 ? 0 : (size_t)n * RK_RS_WORDS * 4u; }

int rk_rmt_guard(const rk_rmt_params_t *p, void *d_state, int64_t n, int32_t K, const void *d_in, rk_vdt_cmd_t *d_cmd_out,
                 uint32_t *d_abort_out, void *stream) {
  if(n == 0 || K == 0) return RK_OK;
  if(n < 0 || K < 0 || !d_state || !d_in || !d_cmd_out) {
    set_error("rk_rmt_guard: bad n / K or NULL state / in / cmd_out");
    return RK_ERR_ARG;
  }
  if(((uintptr_t)d_state & 15u) || ((uintptr_t)d_in & 15u) || ((uintptr_t)d_cmd_out & 15u) || ((uintptr_t)d_abort_out & 3u)) {
    set_error("rk_rmt_guard: d_state / d_in / d_cmd_out must be 16-byte aligned");
    return RK_ERR_ARG;
  }
  rk_rmt_params_t q;
  if(p) q = *p;
  else rk_rmt_default_params(&q);
  if(int rc = require_device()) return rc;
  rmt_guard_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(q, (uint4 *)d_state, n, K, (const uint4 *)d_in,
                                                                                  (uint4 *)d_cmd_out, d_abort_out);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

struct rk_rmt {
  rk_rmt_params_t p;
  uint32_t       *d_buf; // [0..3] state plane, [4..15] input record, [16..19] output record, [20] abort
  uint32_t       *h_stage;
  cudaStream_t    st;
};
int rk_rmt_create(rk_rmt_t **out, const rk_rmt_params_t *p) {
  if(!out) return RK_ERR_ARG;
  *out = nullptr;
  if(int rc = require_device()) return rc;
  rk_rmt *h = new rk_rmt();
  memset(h, 0, sizeof(*h));
  if(p) h->p = *p;
  else rk_rmt_default_params(&h->p);
  cudaError_t e = cudaMalloc((void **)&h->d_buf, 24 * 4);
  if(e == cudaSuccess) e = cudaMallocHost((void **)&h->h_stage, 24 * 4);
  if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if(e == cudaSuccess) e = cudaMemsetAsync(h->d_buf, 0, 24 * 4, h->st);
  if(e != cudaSuccess) {
    int rc = cuda_fail(e, "rk_rmt_create");
    rk_rmt_destroy(h);
    return rc;
  }
  *out = h;
  return RK_OK;
}
void rk_rmt_destroy(rk_rmt_t *h) {
  if(!h) return;
  if(h->st) {
    cudaStreamSynchronize(h->st);
    cudaStreamDestroy(h->st);
  }
  if(h->d_buf) cudaFree(h->d_buf);
  if(h->h_stage) cudaFreeHost(h->h_stage);
  delete h;
}
int rk_rmt_cycle(rk_rmt_t *h, const uint32_t in[RK_RI_WORDS], rk_vdt_cmd_t *out, uint32_t *abort_val) {
  if(!h || !in) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(h->h_stage + 4, in, RK_RI_WORDS * 4); // n = 1: the three cells of a record are consecutive
  RK_CUDA(cudaMemcpyAsync(h->d_buf + 4, h->h_stage + 4, RK_RI_WORDS * 4, cudaMemcpyHostToDevice, h->st));
  if(int rc = rk_rmt_guard(&h->p, h->d_buf, 1, 1, h->d_buf + 4, (rk_vdt_cmd_t *)(h->d_buf + 16), h->d_buf + 20, h->st)) return rc;
  RK_CUDA(cudaMemcpyAsync(h->h_stage + 16, h->d_buf + 16, 5 * 4, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  if(out) memcpy(out, h->h_stage + 16, 16);
  if(abort_val) *abort_val = h->h_stage[20];
  return RK_OK;
}
int rk_rmt_get_state(rk_rmt_t *h, uint32_t words[RK_RS_WORDS]) {
  if(!h || !words) return RK_ERR_ARG;
  RK_CUDA(cudaMemcpyAsync(h->h_stage, h->d_buf, RK_RS_WORDS * 4, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(words, h->h_stage, RK_RS_WORDS * 4);
  return RK_OK;
}

int rk_mymath_atan2f(const float *d_y, const float *d_x, float *d_out, int64_t n, void *stream) {
  if(n == 0) return RK_OK;
  if(n < 0 || !d_y || !d_x || !d_out) {
    set_error("rk_mymath_atan2f: bad n or NULL array");
    return RK_ERR_ARG;
  }
  if(int rc = require_device()) return rc;
  atan2f_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_y, d_x, d_out, n);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}
}
