/* TEST INFRASTRUCTURE (oracle).  Replaces lib/IcsClass_V210/src/IcsHardSerialClass.h (which
 * needs a real Arduino UART) with an "ideal servo" on the same IcsBaseClass: every 3-byte
 * position/free command is answered with the commanded position (or the last one for a
 * free command), so JointIcsServo::update() (src/ArmDrive/AD_joint_ics_servo.cpp:5-29)
 * runs unmodified and its integer command is observable through last_tx[]. */
#ifndef ORACLE_STUB_ICS_HARD_SERIAL_H_
#define ORACLE_STUB_ICS_HARD_SERIAL_H_
#include <Arduino.h>
#include <IcsBaseClass.h>

class IcsHardSerialClass : public IcsBaseClass {
public:
  IcsHardSerialClass() {}
  bool begin() { return true; }
  bool synchronize(byte *txBuf, byte txLen, byte *rxBuf, byte rxLen) override {
    for(int i = 0; i < 4; i++) last_tx[i] = (i < txLen) ? txBuf[i] : 0;
    n_tx++;
    if((txBuf[0] & 0xE0) == 0x80) { /* position / free command */
      if(txBuf[1] != 0 || txBuf[2] != 0) {
        pos_h = txBuf[1];
        pos_l = txBuf[2];
      }
      if(rxLen >= 3) {
        rxBuf[0] = txBuf[0] & 0x7F;
        rxBuf[1] = pos_h;
        rxBuf[2] = pos_l;
      }
    } else {
      for(int i = 0; i < rxLen; i++) rxBuf[i] = (i < txLen) ? txBuf[i] : 0;
    }
    return true;
  }
  byte     last_tx[4] = {0, 0, 0, 0};
  uint32_t n_tx       = 0;
  byte     pos_h = (7500 >> 7) & 0x7F, pos_l = 7500 & 0x7F; /* neutral */
};
#endif
