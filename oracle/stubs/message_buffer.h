/* TEST INFRASTRUCTURE (oracle).  FreeRTOS message buffers, reduced to what the task shells of
 * the reference call (VD_task_main.cpp:111,154,178,397): the harness owns the queue. */
#ifndef ORACLE_STUB_MESSAGE_BUFFER_H_
#define ORACLE_STUB_MESSAGE_BUFFER_H_
#include <stddef.h>
#include <stdint.h>
typedef void *MessageBufferHandle_t;
MessageBufferHandle_t xMessageBufferCreate(size_t bytes);
size_t                xMessageBufferReceive(MessageBufferHandle_t h, void *dst, size_t bytes, uint32_t ticks_to_wait);
size_t                xMessageBufferSend(MessageBufferHandle_t h, const void *src, size_t bytes, uint32_t ticks_to_wait);
#endif
