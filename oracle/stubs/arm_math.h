/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * Stand-in for CMSIS-DSP "arm_math.h".  CMSIS-DSP is an UN-VENDORED, UN-PINNED third-party
 * dependency of the reference (Teensyduino core, platformio.ini:15-17); call sites:
 * src/Utility/util_mymath.hpp:44,45,52 and src/Utility/util_vel_interp.hpp:90.
 * The three functions the hot path reaches are restated in oracle/cmsis_shim.c from the
 * published CMSIS-DSP algorithm (512-interval table + linear interpolation for sin/cos,
 * VSQRT semantics for sqrt).  PARITY UNPINNED at this boundary: no golden vector in the
 * reference exercises it.  -DORACLE_TRIG_LIBM switches sin/cos to libm (the shim the
 * SURVEY.md Appendix D probe values were produced with).
 */
#ifndef ORACLE_STUB_ARM_MATH_H_
#define ORACLE_STUB_ARM_MATH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PI
#define PI 3.14159265358979f
#endif

typedef float float32_t;

typedef enum {
  ARM_MATH_SUCCESS        = 0,
  ARM_MATH_ARGUMENT_ERROR = -1
} arm_status;

float32_t  arm_sin_f32(float32_t x);
float32_t  arm_cos_f32(float32_t x);
arm_status arm_sqrt_f32(float32_t in, float32_t *pOut);

#ifdef __cplusplus
}
#endif

#endif
