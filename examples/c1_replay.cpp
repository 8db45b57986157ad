// c1_replay.cpp -- BASELINE.json configs[0] through the firmware-side shim: one mecanum vehicle, 1 kHz tick, 10 s of
// command replay (MOVE (200, 100, 1.0) at tick 0, STOP at tick 5000; yaw_deg = ((tick / 10) mod 360) - 180), wheel
// feedback from the integer motor plant of include/robotick.h, driven exactly as VD_task_main.cpp drives its statics:
// CAN mailbox callbacks -> rx_callback x4, then the 1 kHz ISR: set_now_yaw_world(deg2rad(yaw)); update(); tx_routine.
//   g++ -std=c++17 -Iinclude examples/c1_replay.cpp -o c1_replay roboken-fmskf-robot-controller_b200/librobotick_b200.so
// Prints the rows tests/test_example_shim_gpu.py compares with the golden trace of the compiled reference.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vdt_shim.hpp"

int main(int argc, char **argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 10000;
  static VDT::VEHICLE_CTRL vhclCtrl; // the firmware keeps these in static storage (VD_task_main.cpp:75-108)
  if(!vhclCtrl.ok()) {
    fprintf(stderr, "c1_replay: %s\n", rk_last_error());
    return 2;
  }
  static VDT::MOTOR_IF_M2006 motors[4] = {{vhclCtrl, 0}, {vhclCtrl, 1}, {vhclCtrl, 2}, {vhclCtrl, 3}};
  VDT::Direction a_move = {1000.0f, 1000.0f, 30.0f}, j_move = {10000.0f, 10000.0f, 300.0f}; // C_ACCEL_MAX_MOVE / C_JERK_MAX_MOVE
  VDT::Direction a_stop = {2000.0f, 2000.0f, 70.0f}, j_stop = {30000.0f, 30000.0f, 1000.0f}; // ..._STOP
  const float    deg2rad = 3.14159265358979f / 180.0f; // mymath::const_deg2rad
  int32_t        rpm[4] = {0, 0, 0, 0}, ang[4] = {0, 0, 0, 0};
  for(int t = 0; t < steps; t++) {
    if(t == 0) {
      VDT::Direction v = {200.0f, 100.0f, 1.0f};
      vhclCtrl.start();
      vhclCtrl.set_target_vel(v, a_move, j_move);
    }
    if(t == 5000) {
      VDT::Direction v = {0.0f, 0.0f, 0.0f};
      vhclCtrl.set_target_vel(v, a_stop, j_stop);
    }
    for(int k = 0; k < 4; k++) { // integer motor plant -> C610 feedback frame -> mailbox callback
      const int32_t cur = motors[k].get_rawCurr_tgt();
      rpm[k] += ((cur * 4 - rpm[k]) >> 4);
      ang[k] = (ang[k] + rpm[k] * 8192 / 60000) & 8191;
      VDT::MOTOR_IF_M2006::CanMsgRx m;
      const uint8_t f[8] = {(uint8_t)(ang[k] >> 8), (uint8_t)ang[k], (uint8_t)(rpm[k] >> 8), (uint8_t)rpm[k], (uint8_t)(cur >> 8), (uint8_t)cur, 0, 0};
      memcpy(m.u8_data, f, 8);
      motors[k].rx_callback(&m, (int16_t)(((t + 1) * 1000) & 0x7FFF));
    }
    if(t % 10 == 0) vhclCtrl.set_now_yaw_world((float)(((t / 10) % 360) - 180) * deg2rad); // the IMU task runs at 100 Hz
    vhclCtrl.update();
    if(t < 10 || t % 50 == 0 || t == 4999 || t == 5001 || t == steps - 1) {
      VDT::Direction p, v, g;
      vhclCtrl.get_vehicle_pos_m_latest(p), vhclCtrl.get_vehicle_vel_mmps_latest(v), vhclCtrl.get_vehicle_vel_tgt_mmps_latest(g);
      uint32_t w[9];
      memcpy(w, &p, 12), memcpy(w + 3, &v, 12), memcpy(w + 6, &g, 12);
      printf("%d", t);
      for(int j = 0; j < 9; j++) printf(" %08x", w[j]);
      for(int k = 0; k < 4; k++) printf(" %d", (int)motors[k].get_rawCurr_tgt());
      printf("\n");
    }
  }
  return 0;
}
