// Microbenchmark: does packed FP32 (FADD2/FMUL2/FFMA2, sm_100) relieve the issue bottleneck of a
// non-contractable (bit-parity) FP32 kernel?   nvcc -gencode arch=compute_100a,code=sm_100a -O3
// Each variant runs CH independent chains per thread, 8 warps/SMSP resident.
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
  float2 a[8];
  int    q[8];
#pragma unroll
  for(int i = 0; i < 8; i++) a[i] = make_float2(seed + threadIdx.x + i, seed + 2.f * i), q[i] = threadIdx.x * (i + 1);
  const float2 m = make_float2(0.999f + seed * 1e-9f, 0.998f), c = make_float2(seed * 1e-7f, seed * 2e-7f);
  const int    qi = (int)seed + 3;
#pragma unroll 1
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(int u = 0; u < 4; u++) {
#pragma unroll
      for(int i = 0; i < 8; i++) {
        if(V == 0) { // scalar, one flop per instruction: 2 instr per element pair
          a[i].x = (i & 1) ? __fmul_rn(a[i].x, m.x) : __fadd_rn(a[i].x, c.x);
          a[i].y = (i & 1) ? __fmul_rn(a[i].y, m.y) : __fadd_rn(a[i].y, c.y);
        } else if(V == 1) { // packed non-fused
          a[i] = (i & 1) ? __fmul2_rn(a[i], m) : __fadd2_rn(a[i], c);
        } else if(V == 2) { // packed fused
          a[i] = __ffma2_rn(a[i], m, c);
        } else if(V == 3) { // scalar fused
          a[i].x = __fmaf_rn(a[i].x, m.x, c.x), a[i].y = __fmaf_rn(a[i].y, m.y, c.y);
        } else if(V == 4) { // packed non-fused + one integer op each (IMAD-free: add/xor)
          a[i] = (i & 1) ? __fmul2_rn(a[i], m) : __fadd2_rn(a[i], c);
          q[i] = (q[i] + qi) ^ it;
        } else if(V == 5) { // scalar non-fused pair + one integer op
          a[i].x = (i & 1) ? __fmul_rn(a[i].x, m.x) : __fadd_rn(a[i].x, c.x);
          a[i].y = (i & 1) ? __fmul_rn(a[i].y, m.y) : __fadd_rn(a[i].y, c.y);
          q[i] = (q[i] + qi) ^ it;
        } else if(V == 6) { // packed + two integer ops
          a[i] = (i & 1) ? __fmul2_rn(a[i], m) : __fadd2_rn(a[i], c);
          q[i] = (q[i] + qi) ^ it;
          q[i] = (q[i] << 1) - q[(i + 1) & 7];
        }
      }
    }
  }
  float s = 0;
  int   z = 0;
#pragma unroll
  for(int i = 0; i < 8; i++) s += a[i].x + a[i].y, z += q[i];
  if(s == 12345.678f || z == 0x7fffffff) out[0] = s + z;
}

template <int V>
void run(const char *name, float *d, int sms, double flop_per_elem, double instr_per_elem) {
  const int blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for(int w = 0; w < 2; w++) k<V><<<blocks, 256>>>(d, iters, 1.0f);
  cudaEventRecord(e0);
  for(int r = 0; r < 5; r++) k<V><<<blocks, 256>>>(d, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double elems = (double)blocks * 256 * iters * 4 * 8; // float2 element updates
  printf("%-44s %8.3f ms  %7.2f TFLOP/s  %7.2f Tinstr/s (thread)  %.3f warp-instr/clk/SMSP @1.965GHz\n", name, ms,
         elems * flop_per_elem / ms * 1e-9, elems * instr_per_elem / ms * 1e-9,
         elems * instr_per_elem / 32.0 / (ms * 1e-3) / (sms * 4.0) / 1.965e9);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s  SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  float *d;
  cudaMalloc(&d, 64);
  int sms = p.multiProcessorCount;
  run<0>("scalar FMUL/FADD (2 instr / pair)", d, sms, 2, 2);
  run<1>("packed FMUL2/FADD2 (1 instr / pair)", d, sms, 2, 1);
  run<2>("packed FFMA2", d, sms, 4, 1);
  run<3>("scalar FFMA x2", d, sms, 4, 2);
  run<4>("packed non-fused + 2 int ops", d, sms, 2, 3);
  run<5>("scalar non-fused pair + 2 int ops", d, sms, 2, 4);
  run<6>("packed non-fused + 4 int ops", d, sms, 2, 5);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
