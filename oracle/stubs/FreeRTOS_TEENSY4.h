/* TEST INFRASTRUCTURE (oracle).  FreeRTOS is out of scope (BASELINE.json north_star);
 * the only symbols the compiled reference files need are a delay and the GPT counter. */
#ifndef ORACLE_STUB_FREERTOS_TEENSY4_H_
#define ORACLE_STUB_FREERTOS_TEENSY4_H_
#include <stdint.h>
static inline void vTaskDelay(uint32_t) {}
uint32_t           get_gptimer_cnt();
/* task-shell harness (oracle/ref_harness_vdt_task.cpp): the 100 Hz loop of VDT::main is driven by the harness,
 * which runs the 1 kHz ISR ticks of one task period inside vTaskDelayUntil() and ends the loop by throwing */
typedef uint32_t TickType_t;
#define configTICK_RATE_HZ 1000
TickType_t xTaskGetTickCount();
void       vTaskDelayUntil(TickType_t *last, TickType_t inc);
#endif
