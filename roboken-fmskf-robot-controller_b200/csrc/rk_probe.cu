// rk_probe.cu -- measurement probes used by bench.py for the roofline denominators that
// MEASURED_PEAKS.json does not carry (it has HBM and bf16 tensor numbers only):
//   rk_probe_fp32(): dense FP32 FFMA throughput (the "FP32 peak" BASELINE.json's metric is a
//   percentage of) and dense non-fused FP32 (FMUL/FADD) issue throughput -- the reachable
//   ceiling of a bit-parity kernel that may not contract a*b+c.
#include "rk_common.cuh"

namespace rk {

template <bool FUSED>
__global__ void __launch_bounds__(256) probe_fp32_kernel(float *out, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f + seed * 1e-9f, c = seed * 1e-7f;
#pragma unroll 1
  for(int i = 0; i < iters; i++) {
#pragma unroll
    for(int u = 0; u < 8; u++) {
      if(FUSED) {
        a0 = __fmaf_rn(a0, m, c), a1 = __fmaf_rn(a1, m, c), a2 = __fmaf_rn(a2, m, c), a3 = __fmaf_rn(a3, m, c);
        a4 = __fmaf_rn(a4, m, c), a5 = __fmaf_rn(a5, m, c), a6 = __fmaf_rn(a6, m, c), a7 = __fmaf_rn(a7, m, c);
      } else { // alternate FMUL / FADD, one flop per issued instruction
        a0 = __fmul_rn(a0, m), a1 = __fadd_rn(a1, c), a2 = __fmul_rn(a2, m), a3 = __fadd_rn(a3, c);
        a4 = __fmul_rn(a4, m), a5 = __fadd_rn(a5, c), a6 = __fmul_rn(a6, m), a7 = __fadd_rn(a7, c);
      }
    }
  }
  float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if(s == 12345.678f) out[0] = s; // keeps the chains alive; practically never true
}

} // namespace rk

extern "C" {

/* Launches the probe on `stream`; the caller times it with CUDA events.  Returns the number
 * of floating-point operations the launch performs through *flops (FFMA = 2). */
int rk_probe_fp32(int fused, int blocks, int iters, float *d_out, double *flops, void *stream) {
  if(int rc = rk::require_device()) return rc;
  if(blocks <= 0 || iters <= 0 || !d_out) {
    rk::set_error("rk_probe_fp32: bad arguments");
    return RK_ERR_ARG;
  }
  if(fused)
    rk::probe_fp32_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters, 1.0f);
  else
    rk::probe_fp32_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_out, iters, 1.0f);
  RK_CUDA(cudaGetLastError());
  if(flops) *flops = (double)blocks * 256.0 * (double)iters * 64.0 * (fused ? 2.0 : 1.0);
  return RK_OK;
}
}
