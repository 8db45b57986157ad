// rmt_shim.hpp -- the vehicle-management block of RMT's routine_ros() (src/RobotManager/RM_task_main.cpp:484-767) and the
// four subscription callbacks that feed it (:159-248), with the firmware's names, over the C-ABI.  The micro-ROS node,
// its publishers and the agent FSM stay with the maintainer; what changes hands is the message buffer.
#pragma once
#include <stdint.h>
#include <string.h>

#include "robotick.h"

namespace RMT {

struct Info_FloorDetect { // FDT::Info_FloorDetect  FD_task_main.hpp:24-33 ; 0 none, 1 floor, 2 wall
  uint8_t rForward, lForward, rBack, lBack, right, left, forward, back;
};

class VehicleManager {
public:
  VehicleManager() { rk_rmt_create(&h_, nullptr); } // nullptr: U32_MCN_* of RM_task_main.cpp:62-64
  ~VehicleManager() { rk_rmt_destroy(h_); }
  // the callbacks only fill the message buffer (vdt_msg_buf_), as in the firmware
  void sb_mecanumCmd_callback(uint32_t cmd, uint32_t time, uint32_t speed) { // :206-218
    clear();
    in_[RK_RI_KIND] = RK_ROS_MECANUM_CMD, in_[RK_RI_A] = cmd, in_[RK_RI_B] = time, in_[RK_RI_C] = speed;
  }
  void sb_mecanumContOdr_callback(double lin_x, double lin_y, double ang_z, uint32_t time_ms) { // :220-233
    clear();
    in_[RK_RI_KIND] = RK_ROS_MECANUM_CONT, in_[RK_RI_A] = time_ms;
    put(RK_RI_X, lin_x), put(RK_RI_Y, lin_y), put(RK_RI_Z, ang_z);
  }
  void sb_mecanumCmdVel_callback(double lin_x, double lin_y, double ang_z) { // :235-248
    clear();
    in_[RK_RI_KIND] = RK_ROS_CMD_VEL;
    put(RK_RI_X, lin_x), put(RK_RI_Y, lin_y), put(RK_RI_Z, ang_z);
  }
  void sb_cmd_callback(uint32_t command) { // :159-204
    clear();
    in_[RK_RI_KIND] = RK_ROS_COMMAND, in_[RK_RI_A] = command;
  }
  // one routine_ros() cycle: floor sensors in, the message VDT::send_req_msg() would get out (kind 0: nothing sent);
  // returns vdt_abort.val (VehicleInfo.fault, :828)
  uint32_t routine_ros(const Info_FloorDetect &fd, rk_vdt_cmd_t &to_vdt) {
    memcpy(&in_[RK_RI_FLOOR], &fd, 8);
    uint32_t fault = 0;
    rk_rmt_cycle(h_, in_, &to_vdt, &fault);
    clear(); // the buffer is consumed
    return fault;
  }
  bool ok() const { return h_ != nullptr; }

private:
  void clear() { memset(in_, 0, sizeof(in_)); }
  void put(int w, double v) { memcpy(&in_[w], &v, 8); }
  rk_rmt_t *h_ = nullptr;
  uint32_t  in_[RK_RI_WORDS] = {0};
};

} // namespace RMT
