// vdt_shim.hpp -- what a firmware maintainer compiles INSTEAD of src/VehicleDrive/VD_vehicle_controller.{hpp,cpp} and
// VD_motor_if_m2006.{hpp,cpp}: the same class names, member names, argument meaning and (void) error behaviour,
// implemented over the C-ABI of librobotick_b200.so (include/robotick.h).  VD_task_main.cpp's statics
// (`static VEHICLE_CTRL vhclCtrl`, the four `MOTOR_IF_M2006`) keep compiling against these.
#pragma once
#include <stdint.h>

#include "robotick.h"

namespace VDT {

struct Direction { // VD_vehicle_controller.hpp:14-18
  float x, y, th;
};

class VEHICLE_CTRL;

class MOTOR_IF_M2006 { // VD_motor_if_m2006.hpp:11-83
public:
  union CanMsgRx { // :13-21: the C610 feedback frame as FlexCAN hands it over
    uint8_t u8_data[8];
  };
  MOTOR_IF_M2006(VEHICLE_CTRL &v, int wheel) : v_(v), w_(wheel) {}
  void    rx_callback(CanMsgRx *m, int16_t usec_id); // VD_motor_if_m2006.cpp:32-72
  int16_t get_rawCurr_tgt();                         // .hpp:52
  int64_t get_rawAngleSum();                         // .hpp:42

private:
  VEHICLE_CTRL &v_;
  int           w_;
};

class VEHICLE_CTRL { // VD_vehicle_controller.hpp:21-87
public:
  VEHICLE_CTRL() { rk_vdt_create(&h_, nullptr); } // nullptr: the constants of VD_task_main.cpp:22-48,75-108,157-160
  ~VEHICLE_CTRL() { rk_vdt_destroy(h_); }
  void update() { rk_vdt_update(h_); } // :52  (void: errors are dropped, as in the firmware; rk_last_error() keeps the text)
  void start() { rk_vdt_start(h_); }   // :54
  void stop() { rk_vdt_stop(h_); }     // :55
  void set_target_vel(Direction &v, Direction &a, Direction &j) { // :56
    const float vv[3] = {v.x, v.y, v.th}, aa[3] = {a.x, a.y, a.th}, jj[3] = {j.x, j.y, j.th};
    rk_vdt_set_target(h_, vv, aa, jj);
  }
  void set_now_yaw_world(float yaw) { rk_vdt_set_yaw(h_, yaw); }                               // :57
  void get_vehicle_pos_m_latest(Direction &d) { rk_vdt_get_pos(h_, &d.x); }                    // :59
  void get_vehicle_vel_mmps_latest(Direction &d) { rk_vdt_get_vel(h_, &d.x); }                 // :60
  void get_vehicle_vel_tgt_mmps_latest(Direction &d) { rk_vdt_get_vel_tgt(h_, &d.x); }         // :61
  bool ok() const { return h_ != nullptr; }
  rk_vdt_t *h_ = nullptr;
};

inline void MOTOR_IF_M2006::rx_callback(CanMsgRx *m, int16_t usec_id) { rk_vdt_rx(v_.h_, w_, m->u8_data, usec_id); }
inline int16_t MOTOR_IF_M2006::get_rawCurr_tgt() {
  int16_t c[4] = {0, 0, 0, 0};
  rk_vdt_get_raw_current(v_.h_, c);
  return c[w_];
}
inline int64_t MOTOR_IF_M2006::get_rawAngleSum() {
  int64_t s[4] = {0, 0, 0, 0};
  rk_vdt_get_angle_sum(v_.h_, s);
  return s[w_];
}

} // namespace VDT
