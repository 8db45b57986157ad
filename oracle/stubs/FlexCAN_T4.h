/* TEST INFRASTRUCTURE (oracle).  Shape of the FlexCAN_T4 API that CAN_CTRL<CAN1>
 * (VD_can_controller.hpp:15-58) uses; write() keeps the last frame so the harness can read the
 * C610 current command the ISR transmits. */
#ifndef ORACLE_STUB_FLEXCAN_T4_H_
#define ORACLE_STUB_FLEXCAN_T4_H_
#include <stdint.h>
#include <string.h>
enum CAN_DEV_TABLE { CAN1 = 1, CAN2, CAN3 };
enum FLEXCAN_RXQUEUE_TABLE { RX_SIZE_256 = 256 };
enum FLEXCAN_TXQUEUE_TABLE { TX_SIZE_16 = 16 };
enum FLEXCAN_MAILBOX { MB0 = 0, MB1, MB2, MB3, MB4, MB5, MB6, MB7 };
enum FLEXCAN_RXTX { TX, RX };
enum FLEXCAN_IDE { NONE = 0, EXT = 1, RTR = 2, STD = 3, INACTIVE };
enum FLEXCAN_FLTEN { ACCEPT_ALL = 0, REJECT_ALL = 1 };
struct CAN_message_t {
  uint32_t id = 0;
  uint8_t  len = 8;
  uint8_t  buf[8] = {0};
};
typedef void (*_MB_ptr)(const CAN_message_t &msg);
template <CAN_DEV_TABLE _bus, FLEXCAN_RXQUEUE_TABLE _rx, FLEXCAN_TXQUEUE_TABLE _tx> class FlexCAN_T4 {
public:
  void begin() {}
  void setBaudRate(uint32_t) {}
  void setMaxMB(uint8_t) {}
  void setMB(FLEXCAN_MAILBOX, FLEXCAN_RXTX, FLEXCAN_IDE) {}
  void setMBFilter(FLEXCAN_FLTEN) {}
  void enableMBInterrupts() {}
  void onReceive(FLEXCAN_MAILBOX mb, _MB_ptr fn) { handler[(int)mb] = fn; }
  bool setMBUserFilter(FLEXCAN_MAILBOX, uint32_t, uint32_t) { return true; }
  int  write(const CAN_message_t &m) {
    last_tx = m;
    n_tx++;
    return 1;
  }
  _MB_ptr       handler[8] = {nullptr};
  CAN_message_t last_tx;
  uint32_t      n_tx = 0;
};
#endif
