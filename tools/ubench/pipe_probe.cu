// Pipe-rate probe for sm_100: which of the tick's instruction classes share an issue port / a pipe, and at what rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe_probe tools/ubench/pipe_probe.cu && ./pipe_probe
// 8 CTAs x 256 threads per SM, 8 independent chains per thread; prints warp-instructions per clock per SM sub-partition
// for the instruction group of each variant (count the SASS of the loop body with cuobjdump to confirm the opcodes).
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed) {
  float2 a[8];
  int    q[8];
  float  s[8];
#pragma unroll
  for(int i = 0; i < 8; i++) a[i] = make_float2(seed + threadIdx.x + i, seed + 2.f * i), q[i] = threadIdx.x * (i + 1), s[i] = seed * i + threadIdx.x;
  const float2 c  = make_float2(seed * 1e-3f, seed * 2e-3f);
  const int    qi = (int)seed + 3;
  const float  lo = -seed * 1e6f, hi = seed * 1e6f;
#pragma unroll 1
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(int u = 0; u < 4; u++) {
#pragma unroll
      for(int i = 0; i < 8; i++) {
        if(V == 0) { // IADD3 only
          q[i] += q[(i + 1) & 7];
        } else if(V == 1) { // LOP3 only
          q[i] ^= q[(i + 3) & 7];
        } else if(V == 2) { // FMNMX only
          asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(hi));
          asm volatile("max.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(lo));
        } else if(V == 3) { // scalar FADD + IADD3
          s[i] = __fadd_rn(s[i], c.x);
          q[i] += q[(i + 1) & 7];
        } else if(V == 4) { // packed FADD2 + IADD3
          a[i] = __fadd2_rn(a[i], c);
          q[i] += q[(i + 1) & 7];
        } else if(V == 5) { // packed FADD2 + scalar FADD
          a[i] = __fadd2_rn(a[i], c);
          s[i] = __fadd_rn(s[i], c.x);
        } else if(V == 6) { // I2FP only
          s[i] = __fadd_rn(s[i], 0.f);
          asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(s[i]) : "r"(q[i] + it));
        } else if(V == 7) { // F2I only
          asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(s[i]));
          s[i] = __int_as_float(q[i] + it);
        } else if(V == 8) { // FSETP + FSEL
          s[i] = (s[i] <= a[i].x) ? a[i].y : s[i];
        } else if(V == 9) { // packed FADD2 + FMNMX
          a[i] = __fadd2_rn(a[i], c);
          if(u & 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(hi));
          else asm volatile("max.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(lo));
        } else if(V == 10) { // scalar FADD x2 + IADD3 (the unpacked equivalent of V4)
          a[i].x = __fadd_rn(a[i].x, c.x), a[i].y = __fadd_rn(a[i].y, c.y);
          q[i] += q[(i + 1) & 7];
        } else if(V == 11) { // IMAD (FMA pipe) + IADD3 (ALU)
          asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(qi), "r"(it));
          q[(i + 1) & 7] += q[(i + 2) & 7];
        }
      }
    }
  }
  float z = 0;
#pragma unroll
  for(int i = 0; i < 8; i++) z += a[i].x + a[i].y + s[i] + (float)q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = z;
}

template <int V>
void run(const char *name, float *d, int sms, double instr_per_elem, double mhz) {
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
  const double groups = (double)iters * 4 * 8 * 8 /*warps per CTA*/ * blocks; // warp-level groups
  const double cyc    = ms * 1e-3 * mhz * 1e6;                              // cycles
  printf("%-44s %8.3f ms  %6.3f cycles/group/SMSP  %.3f warp-instr/clk/SMSP (nominal %g instr/group)\n", name, ms,
         cyc / (groups / (sms * 4.0)), instr_per_elem * groups / (sms * 4.0) / cyc, instr_per_elem);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  printf("%s  SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, mhz);
  float *d;
  cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  const int    sms = p.multiProcessorCount;
  const double f   = mhz / 1000.0;
  run<0>("IADD3", d, sms, 1, f);
  run<1>("LOP3", d, sms, 1, f);
  run<2>("FMNMX x2", d, sms, 2, f);
  run<3>("FADD + IADD3", d, sms, 2, f);
  run<4>("FADD2 + IADD3", d, sms, 2, f);
  run<5>("FADD2 + FADD", d, sms, 2, f);
  run<6>("FADD + I2FP (+IADD)", d, sms, 3, f);
  run<7>("F2I (+IADD)", d, sms, 2, f);
  run<8>("FSETP + FSEL", d, sms, 2, f);
  run<9>("FADD2 + FMNMX", d, sms, 2, f);
  run<10>("FADD x2 + IADD3", d, sms, 3, f);
  run<11>("IMAD + IADD3", d, sms, 2, f);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
