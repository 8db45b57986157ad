// rk_tick.cu -- the full controller tick (vehicle + IMU + arm) as one asynchronous call.
//
// The firmware's three tasks share exactly one value: the vehicle ISR reads
// deg2rad(IMT::get_status_now_yaw()) before each VEHICLE_CTRL::update() (VD_task_main.cpp:368).
// A monolithic one-thread-per-robot kernel would carry 53 planes of state (848 B) per thread
// and run the issue-bound vehicle tick at a fraction of its occupancy, so the coupling is
// expressed through HBM instead: the IMU kernel (HBM-bound, 6 planes) emits the 4-byte yaw
// stream, the vehicle rollout (issue-bound, 28 planes) consumes it, and the arm kernel
// (19 planes, no coupling) runs concurrently on a side stream -- the bandwidth-bound and the
// issue-bound kernels overlap on the same SMs.  Extra traffic: 8 B per robot per slow tick.
#include <mutex>

#include "rk_common.cuh"

namespace rk {
struct TickStreams {
  cudaStream_t side = nullptr;
  cudaEvent_t  fork = nullptr, join = nullptr;
};
static std::mutex  g_tick_mu;
static TickStreams g_tick[64]; // per device

static int tick_streams(TickStreams **out) {
  int dev = 0;
  RK_CUDA(cudaGetDevice(&dev));
  if(dev < 0 || dev >= 64) {
    set_error("rk_tick_rollout: device index %d out of range", dev);
    return RK_ERR_ARG;
  }
  std::lock_guard<std::mutex> lk(g_tick_mu);
  TickStreams                &t = g_tick[dev];
  if(!t.side) {
    RK_CUDA(cudaStreamCreateWithFlags(&t.side, cudaStreamNonBlocking));
    RK_CUDA(cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming));
    RK_CUDA(cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming));
  }
  *out = &t;
  return RK_OK;
}
} // namespace rk

using namespace rk;

extern "C" int rk_tick_rollout(const rk_vdt_params_t *vp, const rk_adt_params_t *ap, void *d_vdt_state, void *d_imt_state,
                               void *d_adt_state, const void *d_adt_cmdtab, int64_t n, const rk_tick_rollout_t *a, void *stream) {
  if(!vp || !ap || !a) {
    set_error("rk_tick_rollout: NULL params/args");
    return RK_ERR_ARG;
  }
  if(n == 0 || a->steps == 0) return RK_OK;
  if(n < 0 || a->steps < 0 || a->slow_period <= 0 || !a->d_regs || !a->d_yaw) {
    set_error("rk_tick_rollout: bad n / steps / slow_period, or d_regs / d_yaw NULL");
    return RK_ERR_ARG;
  }
  if(int rc = require_device()) return rc;
  TickStreams *ts = nullptr;
  if(int rc = tick_streams(&ts)) return rc;
  cudaStream_t  st     = (cudaStream_t)stream;
  const int32_t n_slow = (a->steps + a->slow_period - 1) / a->slow_period;

  // arm: n_slow ticks on the side stream, forked from / joined to the caller's stream
  RK_CUDA(cudaEventRecord(ts->fork, st));
  RK_CUDA(cudaStreamWaitEvent(ts->side, ts->fork, 0));
  int rc = rk_adt_update(ap, d_adt_state, d_adt_cmdtab, n, n_slow, a->d_adt_trace, ts->side);
  RK_CUDA(cudaEventRecord(ts->join, ts->side));
  if(rc == RK_OK) {
    // IMU: n_slow updates, emitting deg2rad(yaw) after each
    rc = rk_imt_update_yaw(d_imt_state, n, n_slow, a->d_regs, a->d_have_quat, nullptr, a->d_yaw, 0, st);
  }
  if(rc == RK_OK) {
    rk_vdt_rollout_t v = {};
    v.steps = a->steps, v.sensor_mode = RK_SENSOR_PLANT;
    v.d_cmd = a->d_cmd, v.n_seg = a->n_seg, v.seg_len = a->seg_len;
    v.d_yaw = a->d_yaw, v.n_yaw = n_slow, v.yaw_period = a->slow_period;
    v.d_trace = a->d_vdt_trace, v.d_goal = a->d_goal, v.d_cost = a->d_cost;
    rc = rk_vdt_rollout(vp, d_vdt_state, n, &v, st);
  }
  RK_CUDA(cudaStreamWaitEvent(st, ts->join, 0));
  return rc;
}
