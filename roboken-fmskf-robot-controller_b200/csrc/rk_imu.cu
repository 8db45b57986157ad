// rk_imu.cu -- IMU_IF_WT901C::update()/updateData()/init() batched (src/Imu/imu_if_wt901c.cpp).
//
// Streaming, HBM-bound form: one thread per IMU, q_init and the readable Data page held in
// registers across the K fused updates; per update 32 B of registers in (16 coalesced int16
// planes) and, when the caller wants every sample's output, 64 B out (four 128-bit stores).
#include <string.h>

#include "rk_common.cuh"
#include "rk_math.cuh"

namespace rk {

struct ImuData {
  float d[16]; // IMU_IF::Data in struct order (imu_if_base.hpp:12-18)
};

// IMU_IF_WT901C::updateData  imu_if_wt901c.cpp:91-129.  x / 32768.0f is an exact scaling
// (2^-15), written as a multiply.
RK_DEV void imu_update_data(const float qi[4], const int r[16], ImuData &o) {
  const float S = 1.0f / 32768.0f;
  float       a[3], g[3], m[3], e[3], q[4];
#pragma unroll
  for(int i = 0; i < 3; i++) {
    a[i] = fmul(fmul((float)r[RK_IMT_REG_AX + i], S), 16.0f);
    g[i] = fmul(fmul((float)r[RK_IMT_REG_GX + i], S), 2000.0f);
    m[i] = (float)r[RK_IMT_REG_HX + i];
    e[i] = fmul(fmul((float)r[RK_IMT_REG_ROLL + i], S), 180.0f);
  }
#pragma unroll
  for(int i = 0; i < 4; i++) q[i] = fmul((float)r[RK_IMT_REG_Q0 + i], S);
  o.d[0] = a[0], o.d[1] = -a[1], o.d[2] = -a[2];
  o.d[3] = g[0], o.d[4] = -g[1], o.d[5] = -g[2];
  o.d[6] = m[0], o.d[7] = -m[1], o.d[8] = -m[2];
  o.d[9]  = fsub(normalize_deg_0to360(e[0]), 180.0f);
  o.d[10] = e[1];
  o.d[11] = e[2];
  // rows copied literally from :123-126 (left-to-right, one rounding per operation)
  o.d[14] = -fsub(fsub(fadd(fmul(qi[3], q[0]), fmul(qi[2], q[1])), fmul(qi[1], q[2])), fmul(qi[0], q[3]));
  o.d[13] = fsub(fadd(fadd(fmul(-qi[2], q[0]), fmul(qi[3], q[1])), fmul(qi[0], q[2])), fmul(qi[1], q[3]));
  o.d[12] = -fsub(fadd(fsub(fmul(qi[1], q[0]), fmul(qi[0], q[1])), fmul(qi[3], q[2])), fmul(qi[2], q[3]));
  o.d[15] = fadd(fadd(fadd(fmul(qi[0], q[0]), fmul(qi[1], q[1])), fmul(qi[2], q[2])), fmul(qi[3], q[3]));
}

__global__ void __launch_bounds__(256)
imt_update_kernel(uint4 *__restrict__ state, int64_t n, int K, const int16_t *__restrict__ regs,
                  const uint8_t *__restrict__ have_quat, float4 *__restrict__ out, float *__restrict__ yaw_rad, int do_init) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n) return;
  float   qi[4];
  ImuData cur;
  {
    const uint4 q = ld_plane(state, n, 0, i);
    qi[0] = u2f(q.x), qi[1] = u2f(q.y), qi[2] = u2f(q.z), qi[3] = u2f(q.w);
#pragma unroll
    for(int pl = 0; pl < 4; pl++) {
      const uint4 v = ld_plane(state, n, 1 + pl, i);
      cur.d[4 * pl] = u2f(v.x), cur.d[4 * pl + 1] = u2f(v.y), cur.d[4 * pl + 2] = u2f(v.z), cur.d[4 * pl + 3] = u2f(v.w);
    }
  }
  uint32_t flags = ld_plane(state, n, 5, i).x;
  for(int u = 0; u < K; u++) {
    int r[16];
#pragma unroll
    for(int k = 0; k < 16; k++) r[k] = (int)__ldcs(regs + ((int64_t)u * 16 + k) * n + i);
    const bool hq = have_quat ? (__ldcs(have_quat + (int64_t)u * n + i) != 0) : true;
    if(do_init && u == 0) { // IMU_IF_WT901C::init  :63-77
      imu_update_data(qi, r, cur);
      const float S = 1.0f / 32768.0f;
#pragma unroll
      for(int k = 0; k < 4; k++) qi[k] = fmul((float)r[RK_IMT_REG_Q0 + k], S);
    } else if(hq) { // ::update  :83-89
      flags &= ~RK_IS_FLAG_ERROR;
      imu_update_data(qi, r, cur);
    } else {
      flags |= RK_IS_FLAG_ERROR; // previous page stays readable
    }
    // what the vehicle ISR reads each tick: mymath::deg2rad(IMT::get_status_now_yaw())
    // (VD_task_main.cpp:368, imu_task_main.cpp:102-104, util_mymath.hpp:13,16)
    if(yaw_rad) __stcs(yaw_rad + (int64_t)u * n + i, fmul(cur.d[RK_IS_D_ANGLE + 2], RK_DEG2RAD));
    if(out) {
#pragma unroll
      for(int pl = 0; pl < 4; pl++)
        __stcs(out + ((int64_t)u * 4 + pl) * n + i, make_float4(cur.d[4 * pl], cur.d[4 * pl + 1], cur.d[4 * pl + 2], cur.d[4 * pl + 3]));
    }
  }
  st_plane(state, n, 0, i, make_uint4(f2u(qi[0]), f2u(qi[1]), f2u(qi[2]), f2u(qi[3])));
#pragma unroll
  for(int pl = 0; pl < 4; pl++)
    st_plane(state, n, 1 + pl, i, make_uint4(f2u(cur.d[4 * pl]), f2u(cur.d[4 * pl + 1]), f2u(cur.d[4 * pl + 2]), f2u(cur.d[4 * pl + 3])));
  st_plane(state, n, 5, i, make_uint4(flags, 0u, 0u, 0u));
}

} // namespace rk

using namespace rk;

extern "C" {

size_t rk_imt_state_words(void) { return RK_IS_WORDS; }
size_t rk_imt_state_bytes(int64_t n) { return n < 0 ? 0 : (size_t)n * RK_IS_WORDS * 4u; }

int rk_imt_update(void *d_state, int64_t n, int32_t K, const int16_t *d_regs, const uint8_t *d_have_quat, float *d_out,
                  int do_init, void *stream) {
  return rk_imt_update_yaw(d_state, n, K, d_regs, d_have_quat, d_out, nullptr, do_init, stream);
}

int rk_imt_update_yaw(void *d_state, int64_t n, int32_t K, const int16_t *d_regs, const uint8_t *d_have_quat, float *d_out,
                      float *d_yaw_rad, int do_init, void *stream) {
  if(n == 0 || K == 0) return RK_OK;
  if(n < 0 || K < 0 || !d_regs) {
    set_error("rk_imt_update: bad n/K/regs");
    return RK_ERR_ARG;
  }
  if(!d_state || ((uintptr_t)d_state & 15u) || ((uintptr_t)d_out & 15u)) {
    set_error("rk_imt_update: d_state/d_out must be 16-byte aligned");
    return RK_ERR_ARG;
  }
  if(int rc = require_device()) return rc;
  imt_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint4 *)d_state, n, K, d_regs, d_have_quat,
                                                                                  (float4 *)d_out, d_yaw_rad, do_init);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

struct rk_imt {
  uint32_t    *d_state;
  int16_t     *d_regs;
  uint8_t     *d_zero; // one zero byte (have_quat = 0)
  uint32_t    *h_stage; // pinned: RK_IS_WORDS words state + 16 int16
  cudaStream_t st;
};

int rk_imt_create(rk_imt_t **out) {
  if(!out) return RK_ERR_ARG;
  *out = nullptr;
  if(int rc = require_device()) return rc;
  rk_imt     *h = new rk_imt();
  cudaError_t e = cudaMalloc((void **)&h->d_state, RK_IS_WORDS * 4);
  if(e == cudaSuccess) e = cudaMalloc((void **)&h->d_regs, 16 * sizeof(int16_t));
  if(e == cudaSuccess) e = cudaMalloc((void **)&h->d_zero, 16);
  if(e == cudaSuccess) e = cudaMemset(h->d_zero, 0, 16);
  if(e == cudaSuccess) e = cudaMallocHost((void **)&h->h_stage, RK_IS_WORDS * 4 + 64);
  if(e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if(e == cudaSuccess) e = cudaMemsetAsync(h->d_state, 0, RK_IS_WORDS * 4, h->st);
  if(e != cudaSuccess) {
    int rc = cuda_fail(e, "rk_imt_create");
    rk_imt_destroy(h);
    return rc;
  }
  *out = h;
  return RK_OK;
}
void rk_imt_destroy(rk_imt_t *h) {
  if(!h) return;
  if(h->st) {
    cudaStreamSynchronize(h->st);
    cudaStreamDestroy(h->st);
  }
  if(h->d_state) cudaFree(h->d_state);
  if(h->d_regs) cudaFree(h->d_regs);
  if(h->d_zero) cudaFree(h->d_zero);
  if(h->h_stage) cudaFreeHost(h->h_stage);
  delete h;
}
static int imt_step(rk_imt_t *h, const int16_t regs[16], int have_quat, int do_init) {
  if(!h || !regs) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st)); // staging buffer reuse
  int16_t *stage = (int16_t *)(h->h_stage + RK_IS_WORDS);
  memcpy(stage, regs, 32);
  RK_CUDA(cudaMemcpyAsync(h->d_regs, stage, 32, cudaMemcpyHostToDevice, h->st));
  // no quaternion frame since the last call: pass a zero have_quat byte (is_error = true)
  const uint8_t *d_flag = (!have_quat && !do_init) ? h->d_zero : nullptr;
  return rk_imt_update(h->d_state, 1, 1, h->d_regs, d_flag, nullptr, do_init, h->st);
}
int rk_imt_init(rk_imt_t *h, const int16_t regs[RK_IMT_REGS]) { return imt_step(h, regs, 1, 1); }
int rk_imt_update1(rk_imt_t *h, const int16_t regs[RK_IMT_REGS], int have_quat) { return imt_step(h, regs, have_quat, 0); }
int rk_imt_get_state(rk_imt_t *h, uint32_t words[RK_IS_WORDS]) {
  if(!h || !words) return RK_ERR_ARG;
  RK_CUDA(cudaMemcpyAsync(h->h_stage, h->d_state, RK_IS_WORDS * 4, cudaMemcpyDeviceToHost, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(words, h->h_stage, RK_IS_WORDS * 4);
  return RK_OK;
}
int rk_imt_set_state(rk_imt_t *h, const uint32_t words[RK_IS_WORDS]) {
  if(!h || !words) return RK_ERR_ARG;
  RK_CUDA(cudaStreamSynchronize(h->st));
  memcpy(h->h_stage, words, RK_IS_WORDS * 4);
  RK_CUDA(cudaMemcpyAsync(h->d_state, h->h_stage, RK_IS_WORDS * 4, cudaMemcpyHostToDevice, h->st));
  RK_CUDA(cudaStreamSynchronize(h->st));
  return RK_OK;
}
int rk_imt_get(rk_imt_t *h, float data[16], int *is_error) {
  uint32_t w[RK_IS_WORDS];
  if(!h || !data) return RK_ERR_ARG;
  if(int rc = rk_imt_get_state(h, w)) return rc;
  memcpy(data, &w[RK_IS_DATA], 64);
  if(is_error) *is_error = (w[RK_IS_FLAGS] & RK_IS_FLAG_ERROR) ? 1 : 0;
  return RK_OK;
}
int rk_imt_get_yaw(rk_imt_t *h, float *yaw_deg) {
  float d[16];
  if(!yaw_deg) return RK_ERR_ARG;
  if(int rc = rk_imt_get(h, d, nullptr)) return rc;
  *yaw_deg = d[RK_IS_D_ANGLE + 2];
  return RK_OK;
}
}
