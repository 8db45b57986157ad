/* TEST INFRASTRUCTURE (oracle).  FreeRTOS is out of scope (BASELINE.json north_star);
 * the only symbols the compiled reference files need are a delay and the GPT counter. */
#ifndef ORACLE_STUB_FREERTOS_TEENSY4_H_
#define ORACLE_STUB_FREERTOS_TEENSY4_H_
#include <stdint.h>
static inline void vTaskDelay(uint32_t) {}
uint32_t           get_gptimer_cnt();
#endif
