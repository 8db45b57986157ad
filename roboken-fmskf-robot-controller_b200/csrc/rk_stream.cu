// rk_stream.cu -- the seeded synthetic command / sensor streams of SURVEY.md section 8d, generated on the device.
//
// A rollout engine fed over PCIe is bound by the host link: the full tick consumes 4.5 KB of tables per robot and
// launch (3.2 KB of WT901 register snapshots, 1 KB of arm waypoints, commands), 75 GB per 2^24-robot pass.  A
// sampling planner does not ship its samples from the host -- it ships the distribution.  These kernels expand a
// 48-byte descriptor (seed, global index of the batch's first robot, distribution parameters) into exactly the
// blocks rk_vdt_rollout / rk_imt_update / rk_adt_push_cmdseq consume, one thread per robot, every store a full
// 128-bit cell.  The definition is the numpy code in streams.py (`*_v2`, 32-bit counter hash); every value here is
// the same integer arithmetic or the same single IEEE operation in the same order, and tests/test_streams_gpu.py
// compares the blocks bit for bit.
#include <math.h>

#include "rk_stream.cuh"

#ifndef RK_STREAM_BLOCK
#define RK_STREAM_BLOCK 256
#endif

namespace rk {

// streams.vehicle_commands_v2: [n_seg][n] rk_vdt_cmd_t
__global__ void __launch_bounds__(RK_STREAM_BLOCK)
stream_vehicle_commands_kernel(const rk_stream_desc_t *__restrict__ dd, int64_t n, int n_seg, uint4 *__restrict__ cmd) {
  const rk_stream_desc_t d = *dd;
  const float four_pi = (float)(4.0 * M_PI), two_pi = (float)(2.0 * M_PI), rl = (float)(6.0 * M_PI);
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
  const uint32_t px = h32_prefix(d.seed, 1u, (uint64_t)(d.first + i));
  for(int s = 0; s < n_seg; s++) {
    const uint32_t b = h32_idx(px, (uint32_t)s);
    float vx  = fsub(fmul(u01_32(sub32(b, 0u)), 800.0f), 400.0f);
    float vy  = fsub(fmul(u01_32(sub32(b, 1u)), 800.0f), 400.0f);
    float vth = fsub(fmul(u01_32(sub32(b, 2u)), four_pi), two_pi);
    // speed_limit_xy (VD_task_main.cpp:127-137): sqrt, clamp, x * lim / len
    const float ln  = fsqrt(fadd(fmul(vx, vx), fmul(vy, vy)));
    const float lim = fminf(ln, 400.0f);
    if(ln != 0.0f) vx = fdiv(fmul(vx, lim), ln), vy = fdiv(fmul(vy, lim), ln);
    else vx = 0.0f, vy = 0.0f;
    vth             = fminf(fmaxf(vth, -rl), rl);
    const bool stop = d.stop_every != 0u && (sub32(b, 3u) % d.stop_every) == 0u;
    uint4      q;
    q.x = stop ? 0u : f2u(vx), q.y = stop ? 0u : f2u(vy), q.z = stop ? 0u : f2u(vth);
    q.w = stop ? (uint32_t)RK_CMD_STOP : (uint32_t)RK_CMD_MOVE;
    __stcs(cmd + (int64_t)s * n + i, q);
  }
  }
}

// streams.vehicle_yaw_reg_v2: int16 [n_yaw][n]; a thread serves two neighbouring vehicles so that stores are 32-bit
__global__ void __launch_bounds__(256)
stream_vehicle_yaw_reg_kernel(const rk_stream_desc_t *__restrict__ dd, int64_t n, int n_yaw, int16_t *__restrict__ out) {
  const int64_t i2 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if(i2 >= n) return;
  const rk_stream_desc_t d = *dd;
  uint32_t phase[2], step[2];
#pragma unroll
  for(int k = 0; k < 2; k++) {
    const uint32_t b    = h32_idx(h32_prefix(d.seed, 5u, (uint64_t)(d.first + i2 + k)), 0u);
    phase[k]            = sub32(b, 0u) & 0xFFFFu;
    const uint32_t rate = ((sub32(b, 1u) % 5u) + 1u) * 182u;
    step[k]             = (sub32(b, 2u) & 1u) ? (0u - rate) : rate; // modulo 2^16 in the end
  }
  const bool pair = (i2 + 1 < n) && ((n & 1) == 0);
  for(int y = 0; y < n_yaw; y++) {
    const uint32_t r0 = (phase[0] + step[0] * (uint32_t)y) & 0xFFFFu, r1 = (phase[1] + step[1] * (uint32_t)y) & 0xFFFFu;
    int16_t       *row = out + (int64_t)y * n + i2;
    if(pair) {
      *reinterpret_cast<uint32_t *>(row) = r0 | (r1 << 16);
    } else {
      row[0] = (int16_t)r0;
      if(i2 + 1 < n) row[1] = (int16_t)r1;
    }
  }
}

// streams.imu_samples_v2 in the cell layout rk_imt_update takes: int16 [n_upd][2][n][8] + have_quat uint8 [n_upd][n]
__global__ void __launch_bounds__(RK_STREAM_BLOCK)
stream_imu_samples_kernel(const rk_stream_desc_t *__restrict__ dd, int64_t n, int n_upd, uint4 *__restrict__ cells,
                          uint8_t *__restrict__ have, int16_t *__restrict__ yaw_reg) {
  const rk_stream_desc_t d = *dd;
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
  const uint32_t px = h32_prefix(d.seed, 20u, (uint64_t)(d.first + i));
  for(int u = 0; u < n_upd; u++) {
    if(cells) {
      uint4 c0, c1;
      bool  hv;
      stream_imu_sample(px, d.first_update + (uint32_t)u, d.drop_every, c0, c1, hv);
      __stcs(cells + ((int64_t)u * 2 + 0) * n + i, c0);
      __stcs(cells + ((int64_t)u * 2 + 1) * n + i, c1);
      if(yaw_reg) yaw_reg[(int64_t)u * n + i] = (int16_t)(c1.y >> 16); // register RK_IMT_REG_YAW again, as a 2-byte column
      if(have) have[(int64_t)u * n + i] = hv ? 1u : 0u;
    } else { // columns only: the IMU update draws the samples itself (rk_tick_rollout_t::d_imu_desc)
      int16_t y;
      bool    hv;
      stream_imu_yaw(px, d.first_update + (uint32_t)u, d.drop_every, y, hv);
      if(yaw_reg) yaw_reg[(int64_t)u * n + i] = y;
      if(have) have[(int64_t)u * n + i] = hv ? 1u : 0u;
    }
  }
  }
}

// streams.arm_sequences_v2 as the SoA slot image rk_adt_push_cmdseq takes: 65 planes of uint4 [n]
__global__ void __launch_bounds__(RK_STREAM_BLOCK)
stream_arm_sequences_kernel(const rk_stream_desc_t *__restrict__ dd, int64_t n, uint4 *__restrict__ img) {
  const rk_stream_desc_t d = *dd;
  for(int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
  const uint64_t inst = (uint64_t)(d.first + i);
  const uint32_t         b0   = h32_idx(h32_prefix(d.seed, 30u, inst), 0u);
  const uint32_t         ln   = (sub32(b0, 0u) % (d.arm_max_len - d.arm_min_len + 1u)) + d.arm_min_len;
  const bool             z    = d.arm_dt_zero_every != 0u && (sub32(b0, 1u) % d.arm_dt_zero_every) == 0u;
  __stcs(img + i, make_uint4(d.arm_seq_id, ln, 0u, 0u));
  const uint32_t px = h32_prefix(d.seed, 31u, inst);
  uint32_t       dt = 0u;
  for(int k = 0; k < RK_ACMD_MAX_LEN; k++) {
    uint4 w0 = make_uint4(0u, 0u, 0u, 0u), w1 = w0; // waypoints past the sequence length are zero
    if((uint32_t)k < ln) {
      const uint32_t b   = h32_idx(px, (uint32_t)k);
      uint32_t       inc = (lite32(b, 0u) % 991u) + 10u;
      if((lite32(b, 1u) % 8u) == 0u) inc = 0u;
      if(z && k == 0) inc = 0u;
      dt += inc;
      uint32_t a[5];
#pragma unroll
      for(int j = 0; j < 5; j++) a[j] = f2u(fmul((float)((int32_t)(lite32(b, 2u + (uint32_t)j) % (300u * 64u + 1u)) - 150 * 64), 1.0f / 64.0f));
      w0 = make_uint4(dt, a[0], a[1], a[2]), w1 = make_uint4(a[3], a[4], 0u, 0u);
    }
    __stcs(img + (int64_t)(1 + 2 * k) * n + i, w0);
    __stcs(img + (int64_t)(2 + 2 * k) * n + i, w1);
  }
  }
}

// rk_set_option(RK_OPT_STREAM_CTAS, k): the generators run on at most k CTAs per SM (0 = one thread per robot, full grid);
// a planner that expands the next batch's streams while the current rollout runs keeps them out of the rollout's way
static int g_stream_ctas_per_sm = 0;
int        stream_set_ctas(int v) {
  if(v < -100000 || v > 32) return RK_ERR_ARG; // v < 0: -v CTAs in total (a trickle: less than one CTA per SM)
  g_stream_ctas_per_sm = v;
  return RK_OK;
}
static unsigned stream_grid(int64_t n) {
  unsigned g = (unsigned)((n + RK_STREAM_BLOCK - 1) / RK_STREAM_BLOCK);
  if(g_stream_ctas_per_sm > 0) {
    int dev = 0, sms = 148;
    if(cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned cap = (unsigned)(g_stream_ctas_per_sm * sms);
    if(g > cap) g = cap;
  } else if(g_stream_ctas_per_sm < 0) {
    const unsigned cap = (unsigned)(-g_stream_ctas_per_sm);
    if(g > cap) g = cap;
  }
  return g;
}

static int stream_check(const char *who, const void *d_desc, const void *out, int64_t n) {
  if(n < 0 || !d_desc || ((uintptr_t)d_desc & 7u) || !out || ((uintptr_t)out & 15u)) {
    set_error("%s: n < 0, or descriptor / output NULL or misaligned (descriptor 8, blocks 16 bytes)", who);
    return RK_ERR_ARG;
  }
  return require_device();
}

} // namespace rk

using namespace rk;

extern "C" {

void rk_stream_default_desc(rk_stream_desc_t *d) {
  if(!d) return;
  d->seed = 0x5EEDu, d->first_update = 0u, d->first = 0;
  d->stop_every = 8u, d->drop_every = 64u;
  d->arm_min_len = 2u, d->arm_max_len = 32u, d->arm_seq_id = 1u, d->arm_dt_zero_every = 4u;
}

int rk_stream_vehicle_commands(const rk_stream_desc_t *d_desc, int64_t n, int32_t n_seg, rk_vdt_cmd_t *d_cmd, void *stream) {
  if(n == 0 || n_seg <= 0) return RK_OK;
  if(int rc = stream_check("rk_stream_vehicle_commands", d_desc, d_cmd, n)) return rc;
  stream_vehicle_commands_kernel<<<stream_grid(n), RK_STREAM_BLOCK, 0, (cudaStream_t)stream>>>(d_desc, n, n_seg, (uint4 *)d_cmd);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

int rk_stream_vehicle_yaw_reg(const rk_stream_desc_t *d_desc, int64_t n, int32_t n_yaw, int16_t *d_yaw_reg, void *stream) {
  if(n == 0 || n_yaw <= 0) return RK_OK;
  if(int rc = stream_check("rk_stream_vehicle_yaw_reg", d_desc, d_yaw_reg, n)) return rc;
  const int64_t pairs = (n + 1) / 2;
  stream_vehicle_yaw_reg_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_desc, n, n_yaw, d_yaw_reg);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

int rk_stream_imu_samples_yaw(const rk_stream_desc_t *d_desc, int64_t n, int32_t n_upd, int16_t *d_regs, uint8_t *d_have_quat,
                              int16_t *d_yaw_reg, void *stream) {
  if(n == 0 || n_upd <= 0) return RK_OK;
  // d_regs NULL: only the columns are written (the IMU update then draws the samples itself, rk_tick_rollout_t::d_imu_desc)
  if(int rc = stream_check("rk_stream_imu_samples", d_desc, d_regs ? (const void *)d_regs : (const void *)d_yaw_reg, n)) return rc;
  if((uintptr_t)d_yaw_reg & 1u) {
    set_error("rk_stream_imu_samples_yaw: d_yaw_reg must be 2-byte aligned");
    return RK_ERR_ARG;
  }
  stream_imu_samples_kernel<<<stream_grid(n), RK_STREAM_BLOCK, 0, (cudaStream_t)stream>>>(d_desc, n, n_upd, (uint4 *)d_regs, d_have_quat, d_yaw_reg);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}
int rk_stream_imu_samples(const rk_stream_desc_t *d_desc, int64_t n, int32_t n_upd, int16_t *d_regs, uint8_t *d_have_quat, void *stream) {
  return rk_stream_imu_samples_yaw(d_desc, n, n_upd, d_regs, d_have_quat, nullptr, stream);
}

int rk_stream_arm_sequences(const rk_stream_desc_t *d_desc, int64_t n, void *d_seq, void *stream) {
  if(n == 0) return RK_OK;
  if(int rc = stream_check("rk_stream_arm_sequences", d_desc, d_seq, n)) return rc;
  stream_arm_sequences_kernel<<<stream_grid(n), RK_STREAM_BLOCK, 0, (cudaStream_t)stream>>>(d_desc, n, (uint4 *)d_seq);
  RK_CUDA(cudaGetLastError());
  return RK_OK;
}

} // extern "C"
