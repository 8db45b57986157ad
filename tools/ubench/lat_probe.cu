// Dependent-issue latency of scalar vs packed FP32 ops and of the ALU-pipe ops the tick uses (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void k(float *out, long long *cyc, float seed, int iters) {
  float2 a = make_float2(seed + threadIdx.x, seed * 2.f), m = make_float2(0.999f, 0.998f), c = make_float2(seed * 1e-7f, seed * 2e-7f);
  float  nz = seed * -0.0f;
  int    q  = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for(int i = 0; i < iters; i++) {
#pragma unroll
    for(int u = 0; u < 32; u++) {
      if(V == 0) a.x = __fadd_rn(a.x, c.x);
      if(V == 1) a = __fadd2_rn(a, c);
      if(V == 2) a = __ffma2_rn(a, m, make_float2(nz, nz));
      if(V == 3) a.x = __fmaf_rn(a.x, m.x, c.x);
      if(V == 4) a.x = fminf(a.x, c.x + (float)u); // FMNMX chain
      if(V == 5) a.x = (a.x <= m.x) ? a.x * 1.0001f : c.x; // FSETP+FSEL(+FMUL)
      if(V == 6) q = max(q + 3, u);                         // int ALU
      if(V == 7) a.x = (float)(__float2int_rz(a.x) + 1);    // F2I + I2FP (+IADD)
    }
  }
  long long t1 = clock64();
  if(threadIdx.x == 0) cyc[0] = t1 - t0;
  if(a.x + a.y + q == 12345.678f) out[0] = a.x;
}
template <int V> void run(const char *n, float *d, long long *dc, int ops) {
  k<V><<<1, 32>>>(d, dc, 1.0f, 1000);
  k<V><<<1, 32>>>(d, dc, 1.0f, 1000);
  long long c;
  cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s %.2f cycles / iteration-step (%d dependent ops each)\n", n, (double)c / (1000.0 * 32), ops);
}
int main() {
  float *d; long long *dc;
  cudaMalloc(&d, 64); cudaMalloc(&dc, 8);
  run<0>("FADD dependent", d, dc, 1);
  run<1>("FADD2 dependent", d, dc, 1);
  run<2>("FFMA2 dependent", d, dc, 1);
  run<3>("FFMA dependent", d, dc, 1);
  run<4>("FMNMX(+FADD indep) dependent", d, dc, 1);
  run<5>("FSETP+FSEL+FMUL dependent", d, dc, 3);
  run<6>("IADD+VIMNMX dependent", d, dc, 2);
  run<7>("F2I+IADD+I2FP dependent", d, dc, 3);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
