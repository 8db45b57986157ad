/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * Restatement of the three CMSIS-DSP functions the reference's hot path calls
 * (src/Utility/util_mymath.hpp:44-45,50-54; src/Utility/util_vel_interp.hpp:90).
 * CMSIS-DSP is not under /root/reference and not version-pinned (platformio.ini:15-17,
 * "platform = teensy"), so this follows the PUBLISHED algorithm of arm_sin_f32.c /
 * arm_cos_f32.c (FAST_MATH_TABLE_SIZE = 512, linear interpolation) and of arm_sqrt_f32
 * on an FPU core (VSQRT.F32: IEEE correctly rounded; negative input -> 0 + error).
 * PARITY UNPINNED at this boundary.
 *
 * The two published variants of the index wrap (older "& 0x1ff", newer "if(index>=512)")
 * and of the floor test ("x < 0" vs "in < 0") agree for every input the hot path can
 * produce: VEHICLE_CTRL::update() (VD_vehicle_controller.cpp:47-49) only passes
 * normalize_rad_0to2pi() results, i.e. x in [0, 2*pi), for which in in [0, 1.25).
 */
#include "arm_math.h"
#include <math.h>

#define FAST_MATH_TABLE_SIZE 512

static const float sinTable_f32[FAST_MATH_TABLE_SIZE + 1] = {
#include "cmsis_sin_table.inc"
};

#ifdef ORACLE_TRIG_LIBM

float32_t arm_sin_f32(float32_t x) { return sinf(x); }
float32_t arm_cos_f32(float32_t x) { return cosf(x); }

#else

static float table_lerp(float in) {
  int32_t  n;
  float    findex, fract, a, b;
  uint16_t index;

  n = (int32_t)in;
  if(in < 0.0f) n--;
  in     = in - (float)n;
  findex = (float)FAST_MATH_TABLE_SIZE * in;
  index  = (uint16_t)findex;
  if(index >= FAST_MATH_TABLE_SIZE) {
    index = 0;
    findex -= (float)FAST_MATH_TABLE_SIZE;
  }
  fract = findex - (float)index;
  a     = sinTable_f32[index];
  b     = sinTable_f32[index + 1];
  return (1.0f - fract) * a + fract * b;
}

float32_t arm_sin_f32(float32_t x) {
  float in = x * 0.159154943092f;
  return table_lerp(in);
}

float32_t arm_cos_f32(float32_t x) {
  float in = x * 0.159154943092f + 0.25f;
  return table_lerp(in);
}

#endif

arm_status arm_sqrt_f32(float32_t in, float32_t *pOut) {
  if(in >= 0.0f) {
    *pOut = sqrtf(in);
    return ARM_MATH_SUCCESS;
  }
  *pOut = 0.0f;
  return ARM_MATH_ARGUMENT_ERROR;
}
