/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * Minimal stand-in for the Teensyduino <Arduino.h> so that the UNMODIFIED reference
 * sources (src/VehicleDrive, src/Imu, src/ArmDrive, src/Utility, lib/wt901c,
 * lib/IcsClass_V210) compile for x86 under oracle/Makefile.  Only the handful of
 * names those translation units touch are provided; everything is a host fake.
 */
#ifndef ORACLE_STUB_ARDUINO_H_
#define ORACLE_STUB_ARDUINO_H_

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef uint8_t byte;
typedef bool    boolean;

#ifndef HIGH
#define HIGH 1
#define LOW 0
#define OUTPUT 1
#define INPUT 0
#endif

static inline void     pinMode(int, int) {}
static inline void     digitalWrite(int, int) {}
/* micros(): a harness that needs a running clock defines ORACLE_MICROS_EXTERN before including this header */
#ifdef ORACLE_MICROS_EXTERN
uint32_t micros();
#else
static inline uint32_t micros() { return 0; }
#endif
static inline uint32_t millis() { return 0; }
static inline void     delay(uint32_t) {}

/* A byte FIFO the harness fills; IMU_IF_WT901C drains it through available()/read(). */
class HardwareSerial {
public:
  void   begin(long) {}
  size_t write(const uint8_t *p, size_t n) {
    tx_count += n;
    (void)p;
    return n;
  }
  void flush() {}
  void clear() {}
  int  available() { return (int)(n_ - pos_); }
  int  read() { return (pos_ < n_) ? buf_[pos_++] : -1; }

  /* harness side */
  void feed(const uint8_t *p, size_t n) {
    if(pos_ == n_) pos_ = n_ = 0;
    if(n > sizeof(buf_) - n_) n = sizeof(buf_) - n_;
    memcpy(buf_ + n_, p, n);
    n_ += n;
  }
  size_t tx_count = 0;

private:
  uint8_t buf_[512];
  size_t  n_ = 0, pos_ = 0;
};

/* Teensy IntervalTimer: keeps the callback so the harness can fire the 1 kHz ISR itself */
class IntervalTimer {
public:
  bool begin(void (*fn)(), uint32_t period) {
    isr = fn, arg = period;
    return true;
  }
  void (*isr)() = nullptr;
  uint32_t arg  = 0;
};

extern HardwareSerial Serial6;
extern HardwareSerial Serial7;

#endif
