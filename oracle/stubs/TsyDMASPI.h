/* TEST INFRASTRUCTURE (oracle).  The MPU6500 SPI link of VD_task_main.cpp:58-73 is never exercised
 * (VEHICLE_CTRL::update takes its yaw from set_now_yaw_world); the calls only have to compile. */
#ifndef ORACLE_STUB_TSYDMASPI_H_
#define ORACLE_STUB_TSYDMASPI_H_
#include <stddef.h>
#include <stdint.h>
#define MSBFIRST 1
#define SPI_MODE3 3
struct SPISettings {
  SPISettings(uint32_t, int, int) {}
};
struct TsyDMASPIStub {
  void   begin(uint8_t, SPISettings, bool) {}
  void   queue(uint8_t *, uint8_t *, size_t, uint8_t) {}
  size_t remained() { return 0; }
};
static TsyDMASPIStub TsyDMASPI0;
#endif
