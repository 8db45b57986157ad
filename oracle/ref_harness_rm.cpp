/* TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
 *
 * The reference's RobotManager task, unmodified: src/RobotManager/RM_task_main.cpp is included below where it
 * lies (routine_ros() and its state are file-static), together with src/Utility/util_mymath.cpp (the table
 * arctangent; its tables have internal linkage too).  micro-ROS is represented by the message / service
 * STRUCT headers the reference vendors under lib/micro_ros_arduino/src (declarations only) and by the link
 * stubs at the bottom of this file; Ethernet, FreeRTOS and the other tasks (VDT, ADT, CGT, IMT, FDT) are
 * stubs that script the inputs and capture what the manager sends.
 *
 * One manager cycle = one routine_ros() call: the first rclc_executor_spin_some() of the cycle delivers the
 * scripted ROS message through the reference's own subscription callback, FDT::get_now_FDinfo() returns the
 * scripted floor sensors, VDT::send_req_msg() captures the outgoing vehicle message.
 */
#include <stdlib.h>
#include <string.h>

#include <Arduino.h>
#include <FreeRTOS_TEENSY4.h>
#include <message_buffer.h>

#include "robotick.h"

/* ---- what Teensyduino / NativeEthernet would provide ------------------------------------------ */
static inline int digitalRead(int) { return 0; }
struct IPAddress {
  IPAddress(int, int, int, int) {}
  IPAddress() {}
  operator uint32_t() const { return 0; }
};
struct EthernetStub {
  int       begin(byte *, unsigned long = 0, unsigned long = 0) { return 1; }
  void      begin(byte *, IPAddress) {}
  IPAddress localIP() { return IPAddress(); }
} Ethernet;
static uint32_t HW_OCOTP_MAC0 = 0, HW_OCOTP_MAC1 = 0;
static inline void set_microros_native_ethernet_udp_transports(byte *, IPAddress, IPAddress, uint16_t) {}

HardwareSerial Serial6;
HardwareSerial Serial7;
uint32_t       get_gptimer_cnt() { return 0; }
namespace DEBUG {
char EXT_PRINT_BUF[1024];
void print(char *, uint32_t) {}
void record_proc_load(uint8_t, uint8_t) {}
} // namespace DEBUG
namespace LGT {
void push_buffer(char *, uint32_t) {}
} // namespace LGT
TickType_t xTaskGetTickCount() { return 0; }
void       vTaskDelayUntil(TickType_t *, TickType_t) {}

/* the reference, as it is */
#include "Utility/util_mymath.cpp"
#include "RobotManager/RM_task_main.cpp"

/* ---- harness state ----------------------------------------------------------------------------- */
namespace {
struct Cycle {
  const uint32_t *in = nullptr; /* RK_RI_* words of this cycle */
  bool            delivered = false;
  bool            sent = false;
  VDT::MSG_REQ    out;
} g_cyc;
double rm_double(const uint32_t *w) {
  uint64_t u = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  double   d;
  memcpy(&d, &u, 8);
  return d;
}
} // namespace

/* ---- the other tasks ---------------------------------------------------------------------------- */
namespace FDT {
void get_now_FDinfo(Info_FloorDetect &f) {
  const uint32_t *w = g_cyc.in + RK_RI_FLOOR;
  uint8_t         b[8];
  memcpy(b, w, 8);
  f.u8_rForward = b[0], f.u8_lForward = b[1], f.u8_rBack = b[2], f.u8_lBack = b[3];
  f.u8_right = b[4], f.u8_left = b[5], f.u8_forward = b[6], f.u8_back = b[7];
}
float get_now_walldist(SENSOR_DIR) { return 0.0f; }
} // namespace FDT
namespace VDT {
void send_req_msg(MSG_REQ *m) { g_cyc.out = *m, g_cyc.sent = true; }
void get_status_now_vehicle_pos_world(float &x, float &y, float &r) { x = y = r = 0.0f; }
void get_status_now_vehicle_vel_world(float &x, float &y, float &r) { x = y = r = 0.0f; }
void get_status_now_vehicle_vel(float &x, float &y, float &r) { x = y = r = 0.0f; }
} // namespace VDT
namespace ADT {
void     send_req_msg(MSG_REQ *) {}
uint32_t get_status_timeangle_proc(uint32_t) { return 99; }
uint32_t get_status_movepos_proc(uint32_t) { return 99; }
void     get_arm_angle_rad(float *a) { memset(a, 0, 5 * sizeof(float)); }
} // namespace ADT
namespace CGT {
void  send_req_msg(MSG_REQ *) {}
float get_pitch_angle_deg() { return 0.0f; }
} // namespace CGT
namespace IMT {
void  get_status_now_imu(imu_data &d) { memset(&d, 0, sizeof(d)); }
float get_status_now_yaw() { return 0.0f; }
} // namespace IMT

/* ---- micro-ROS link stubs ------------------------------------------------------------------------- */
extern "C" {
rcl_ret_t rclc_executor_spin_some(rclc_executor_t *e, const uint64_t) {
  if(e != &executor || g_cyc.delivered) return RCL_RET_OK;
  g_cyc.delivered   = true;
  const uint32_t *w = g_cyc.in;
  switch(w[RK_RI_KIND]) {
  case RK_ROS_MECANUM_CMD: {
    interfaces__msg__MecanumCommand m;
    memset(&m, 0, sizeof(m));
    m.cmd = w[RK_RI_A], m.time = w[RK_RI_B], m.speed = w[RK_RI_C];
    sb_mecanumCmd_callback(&m);
  } break;
  case RK_ROS_MECANUM_CONT: {
    interfaces__msg__MecanumContOrder m;
    memset(&m, 0, sizeof(m));
    m.speed.linear.x = rm_double(w + RK_RI_X), m.speed.linear.y = rm_double(w + RK_RI_Y), m.speed.angular.z = rm_double(w + RK_RI_Z);
    m.time_ms = w[RK_RI_A];
    sb_mecanumContOdr_callback(&m);
  } break;
  case RK_ROS_CMD_VEL: {
    geometry_msgs__msg__Twist m;
    memset(&m, 0, sizeof(m));
    m.linear.x = rm_double(w + RK_RI_X), m.linear.y = rm_double(w + RK_RI_Y), m.angular.z = rm_double(w + RK_RI_Z);
    sb_mecanumCmdVel_callback(&m);
  } break;
  case RK_ROS_COMMAND: {
    interfaces__msg__Command m;
    memset(&m, 0, sizeof(m));
    m.command = w[RK_RI_A];
    sb_cmd_callback(&m);
  } break;
  default: break;
  }
  return RCL_RET_OK;
}
rcl_ret_t rcl_publish(const rcl_publisher_t *, const void *, rmw_publisher_allocation_t *) { return RCL_RET_OK; }
}

namespace {
inline uint32_t &soa(uint32_t *blk, int64_t n, int64_t i, int w) { return blk[((int64_t)(w / 4) * n + i) * 4 + (w % 4)]; }
rk_vdt_cmd_t record_of(const VDT::MSG_REQ &m) { /* the RK_CMD_MSG_* record rk_vdt_rollout's command layer takes */
  rk_vdt_cmd_t c;
  memset(&c, 0, sizeof(c));
  if(m.common.MsgId == VDT::MSG_ID::REQ_MOVE_DIR) {
    memcpy(&c.vx, &m.move_dir.u32_cmd, 4), memcpy(&c.vy, &m.move_dir.u32_speed, 4);
    c.kind = (int32_t)(RK_CMD_MSG_MOVE_DIR | (m.move_dir.u32_time_ms << 8));
  } else {
    c.vx = m.move_cont_dir.fl_vel_x_mmps, c.vy = m.move_cont_dir.fl_vel_y_mmps, c.vth = m.move_cont_dir.fl_vel_th_radps;
    c.kind = (int32_t)(RK_CMD_MSG_MOVE_CONT_DIR | (m.move_cont_dir.u32_time_ms << 8));
  }
  return c;
}
} // namespace

extern "C" {
/* rk_rmt_guard()'s contract on HOST arrays, instances [i0, i1) (the manager's state is file-static: one at a time) */
void ref_rmt_guard(const rk_rmt_params_t *p, uint32_t *state, int64_t n, int64_t i0, int64_t i1, int K, const uint32_t *in,
                   rk_vdt_cmd_t *cmd_out, uint32_t *abort_out) {
  U32_MCN_NO_CMD_STOP_THRE = p->no_cmd_stop_thre, U32_MCN_WALL_LEAVE_TIME_MS = p->wall_leave_time_ms;
  U32_MCN_WALL_LEAVE_SPEED_MMPS = p->wall_leave_speed_mmps;
  /* the one piece of create_microros_entities() (:457-458) the cycle depends on: the ArmInfo publish buffer */
  msg_pb_armInfo.servo.theta.data = &fl_ArmAngThetaBuffer[0], msg_pb_armInfo.servo.theta.capacity = U8_ARMANGLE_BUF_LEN;
  for(int64_t i = i0; i < i1; i++) {
    NOW_CMD_STATUS            = (CmdStatus)soa(state, n, i, RK_RS_CMD_STATUS);
    IS_IGNORE_FLOOR_DETECTION = soa(state, n, i, RK_RS_IGNORE_FLOOR) != 0;
    U32_MCN_NO_CMD_CNT        = soa(state, n, i, RK_RS_NO_CMD_CNT);
    vdt_abort.val             = soa(state, n, i, RK_RS_ABORT);
    IS_MCN_CMD_UPDATED = false, U8_VDT_MSG_BUF_WRITE = 0, U8_PUB_PHASE = 0;
    memset(vdt_msg_buf_, 0, sizeof(vdt_msg_buf_));
    for(int u = 0; u < K; u++) {
      uint32_t w[RK_RI_WORDS];
      for(int k = 0; k < RK_RI_WORDS; k++) w[k] = in[(((int64_t)u * 3 + k / 4) * n + i) * 4 + (k % 4)];
      g_cyc    = Cycle();
      g_cyc.in = w;
      RMT::routine_ros();
      rk_vdt_cmd_t rec;
      memset(&rec, 0, sizeof(rec));
      if(g_cyc.sent) rec = record_of(g_cyc.out);
      cmd_out[(int64_t)u * n + i] = rec;
      if(abort_out) abort_out[(int64_t)u * n + i] = vdt_abort.val;
    }
    soa(state, n, i, RK_RS_CMD_STATUS) = (uint32_t)NOW_CMD_STATUS, soa(state, n, i, RK_RS_IGNORE_FLOOR) = IS_IGNORE_FLOOR_DETECTION ? 1u : 0u;
    soa(state, n, i, RK_RS_NO_CMD_CNT) = U32_MCN_NO_CMD_CNT, soa(state, n, i, RK_RS_ABORT) = vdt_abort.val;
  }
}
float ref_rm_atanf(float x) { return UTIL::mymath::atanf(x); }
float ref_rm_atan2f(float y, float x) { return UTIL::mymath::atan2f(y, x); }
/* the reference's three arctangent arrays, for pinning tools/gen_atan_table.py */
int ref_rm_atan_tables(float *table, float *delimit, float *width) {
  const int nt = (int)(sizeof(UTIL::mymath::atan_table) / sizeof(float));
  if(table) memcpy(table, UTIL::mymath::atan_table, sizeof(UTIL::mymath::atan_table));
  if(delimit) memcpy(delimit, UTIL::mymath::atan_table_delimit_val, sizeof(UTIL::mymath::atan_table_delimit_val));
  if(width) memcpy(width, UTIL::mymath::atan_table_width, sizeof(UTIL::mymath::atan_table_width));
  return nt;
}
}
