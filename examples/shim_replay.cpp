// shim_replay.cpp -- the IMU, arm and manager shims of INTEGRATION.md as compiled code, driven the way the firmware's
// task loops drive the classes they replace, on recorded inputs (a binary file written by tests/test_example_shims_gpu.py
// from the seeded streams / the golden fixtures).  Prints one line of hex words per tick for the test to compare with
// the trace of the reference compiled for x86.
//   g++ -std=c++17 -Iinclude -Iexamples examples/shim_replay.cpp -o shim_replay roboken-fmskf-robot-controller_b200/librobotick_b200.so
//   shim_replay imu|arm|rmt <input.bin>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "adt_shim.hpp"
#include "imt_shim.hpp"
#include "rmt_shim.hpp"

static std::vector<uint8_t> slurp(const char *path) {
  std::vector<uint8_t> b;
  FILE                *f = fopen(path, "rb");
  if(!f) return b;
  uint8_t tmp[4096];
  size_t  k;
  while((k = fread(tmp, 1, sizeof(tmp), f)) > 0) b.insert(b.end(), tmp, tmp + k);
  fclose(f);
  return b;
}
static void put_words(const void *p, int n) {
  uint32_t w;
  for(int k = 0; k < n; k++) {
    memcpy(&w, (const uint8_t *)p + 4 * k, 4);
    printf(" %08x", w);
  }
}

// input: K records of {int16 regs[16]; int32 quat_frame}; record 0 is consumed by init() (IMT::main, imu_task_main.cpp:36-47)
static int replay_imu(const std::vector<uint8_t> &in) {
  static IMT::IMU_IF_WT901C imu_if; // static storage, as imu_task_main.cpp:25
  if(!imu_if.ok()) return 2;
  const size_t rec = 36, K = in.size() / rec;
  for(size_t u = 0; u < K; u++) {
    int16_t regs[16];
    int32_t quat;
    memcpy(regs, &in[u * rec], 32), memcpy(&quat, &in[u * rec + 32], 4);
    imu_if.on_registers(regs, quat != 0);
    if(u == 0) imu_if.init();
    else imu_if.update();
    IMT::IMU_IF::Data d;
    imu_if.getDataLatest(d);
    const float yaw = imu_if.getYawDate();
    printf("%zu", u);
    put_words(&d, 16);
    put_words(&yaw, 1);
    printf(" %d\n", imu_if.isError() ? 1 : 0);
  }
  return 0;
}

// input: int32 K; int32 n_push; n_push x rk_adt_poscmdseq_t pushed before the first tick (ADT::main, AD_task_main.cpp:199-229)
static int replay_arm(const std::vector<uint8_t> &in) {
  static ADT::ADTModePositioningSeq mode;
  if(!mode.ok() || in.size() < 8) return 2;
  int32_t K, n_push;
  memcpy(&K, &in[0], 4), memcpy(&n_push, &in[4], 4);
  mode.init();
  printf("status_before %d\n", (int)mode.get_q_cmdseq_status(1));
  for(int k = 0; k < n_push; k++) {
    ADT::ADTModePositioningSeq::PosCmdSeq q;
    memcpy(&q, &in[8 + (size_t)k * sizeof(q)], sizeof(q));
    mode.push_cmdseq(q);
  }
  for(int t = 0; t < K; t++) {
    mode.update();
    float tgt[5];
    mode.get_tgt_deg(tgt);
    printf("%d", t);
    put_words(tgt, 5);
    printf(" %d %d\n", (int)mode.get_q_cmdseq_status(1), (int)mode.get_q_cmdseq_status(2));
  }
  return 0;
}

// input: K records of RK_RI_WORDS words as streams.rm_inputs lays one robot out (the ROS message of the cycle, if any, + the
// floor sensors); the shim's callbacks are called as the executor would call them, then routine_ros()
static int replay_rmt(const std::vector<uint8_t> &in) {
  static RMT::VehicleManager mgr;
  if(!mgr.ok()) return 2;
  const size_t rec = RK_RI_WORDS * 4, K = in.size() / rec;
  for(size_t u = 0; u < K; u++) {
    uint32_t w[RK_RI_WORDS];
    memcpy(w, &in[u * rec], rec);
    double x, y, z;
    memcpy(&x, &w[RK_RI_X], 8), memcpy(&y, &w[RK_RI_Y], 8), memcpy(&z, &w[RK_RI_Z], 8);
    switch(w[RK_RI_KIND]) {
    case RK_ROS_MECANUM_CMD: mgr.sb_mecanumCmd_callback(w[RK_RI_A], w[RK_RI_B], w[RK_RI_C]); break;
    case RK_ROS_MECANUM_CONT: mgr.sb_mecanumContOdr_callback(x, y, z, w[RK_RI_A]); break;
    case RK_ROS_CMD_VEL: mgr.sb_mecanumCmdVel_callback(x, y, z); break;
    case RK_ROS_COMMAND: mgr.sb_cmd_callback(w[RK_RI_A]); break;
    default: break;
    }
    RMT::Info_FloorDetect fd;
    memcpy(&fd, &w[RK_RI_FLOOR], 8);
    rk_vdt_cmd_t   msg;
    const uint32_t fault = mgr.routine_ros(fd, msg);
    printf("%zu", u);
    put_words(&msg, 4);
    printf(" %08x\n", fault);
  }
  return 0;
}

int main(int argc, char **argv) {
  if(argc < 3) {
    fprintf(stderr, "usage: shim_replay imu|arm|rmt <input.bin>\n");
    return 1;
  }
  const std::vector<uint8_t> in = slurp(argv[2]);
  int                        rc = 1;
  if(!strcmp(argv[1], "imu")) rc = replay_imu(in);
  else if(!strcmp(argv[1], "arm")) rc = replay_arm(in);
  else if(!strcmp(argv[1], "rmt")) rc = replay_rmt(in);
  if(rc == 2) fprintf(stderr, "shim_replay: %s\n", rk_last_error());
  return rc;
}
