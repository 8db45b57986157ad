/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * Plain-C restatement of the reference's numeric control tick; see robotick_oracle.c.
 * Pinned against oracle/_ref (the unmodified reference compiled for x86) by
 * tests/test_oracle_pin.py and against the golden fixtures in tests/golden/.
 */
#ifndef ROBOTICK_ORACLE_H_
#define ROBOTICK_ORACLE_H_

#include "robotick.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Same contract as rk_vdt_rollout() on HOST arrays (same SoA indexing with pitch n), for
 * instances [i0, i1) on nthreads host threads.  state == NULL: power-on state, results
 * discarded (throughput runs). */
void orc_vdt_rollout(const rk_vdt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1,
                     const rk_vdt_rollout_t *args, int nthreads);
/* VEHICLE_CTRL::set_target_vel on one AoS state (RK_VS_WORDS words) */
void orc_vdt_set_target(const rk_vdt_params_t *p, uint32_t *words, const float v[3], const float a[3],
                        const float j[3]);
/* MOTOR_IF_M2006::rx_callback on one AoS state */
void orc_vdt_rx(const rk_vdt_params_t *p, uint32_t *words, int wheel, const uint8_t frame[8], int16_t usec_id);
/* VEHICLE_CTRL::update on one AoS state */
void orc_vdt_update(const rk_vdt_params_t *p, uint32_t *words);

/* Same contract as rk_imt_update() on HOST arrays, instances [i0, i1). */
void orc_imt_update(uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const int16_t *regs,
                    const uint8_t *have_quat, uint32_t *out, int do_init);

/* Arm batch driver on HOST arrays, same contracts as the rk_adt_* batch calls (see
 * oracle/ref_harness_arm.cpp ref_adt_batch): op 0 = rk_adt_mode_init, 1 = rk_adt_push_cmdseq,
 * 2 = rk_adt_update (K ticks, optional trace), 3 = rk_adt_cmdseq_status. */
void orc_adt_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *cmdtab, int64_t n, int64_t i0, int64_t i1,
                   int K, const uint32_t *seq, const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status);

/* ADTModePositioning batch driver (ref_adp_batch's contract): op 0 init, 1 push_cmd, 2 K ticks, 3 status */
void orc_adp_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *pstate, int64_t n, int64_t i0, int64_t i1, int K,
                   const uint32_t *cmd, const uint8_t *valid, uint32_t *trace, const uint32_t *ids, int32_t *status);

/* rk_adh_* on HOST arrays: op 0 = mode init (mode in K), 2 = K ticks */
void orc_adh_batch(int op, const rk_adt_params_t *p, uint32_t *state, uint32_t *hstate, int64_t n, int64_t i0, int64_t i1, int K,
                   const float *now, uint32_t *trace);

/* rk_adt_bldc_rx() (kind 0..2) / rk_adt_mg_rx() (kind 3) on HOST arrays */
void orc_adt_rx_batch(int kind, const rk_adt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1, const uint64_t *frames,
                      const uint32_t *cmdid, float *cur);

/* rk_imt_feed_bytes() on HOST arrays */
void orc_imt_feed_bytes(uint32_t *state, uint32_t *parser, int64_t n, int64_t i0, int64_t i1, int K, int ncells,
                        const uint32_t *cells, const uint16_t *nbytes, uint32_t *out, float *yaw_rad, int do_init);

/* rk_rmt_guard() on HOST arrays; UTIL::mymath::atanf / atan2f */
void orc_rmt_guard(const rk_rmt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const uint32_t *in,
                   rk_vdt_cmd_t *cmd_out, uint32_t *abort_out);
float orc_atanf(float x);
float orc_atan2f(float y, float x);

float orc_sin(float x);
float orc_cos(float x);
float orc_normalize_rad_0to2pi(float x);
float orc_normalize_deg_0to360(float x);

#ifdef __cplusplus
}
#endif
#endif
