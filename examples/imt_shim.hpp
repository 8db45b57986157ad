// imt_shim.hpp -- what a firmware maintainer compiles INSTEAD of src/Imu/imu_if_wt901c.{hpp,cpp}: the IMU_IF vtable of
// src/Imu/imu_if_base.hpp:20-29 with the same member names and argument meaning, over the C-ABI of librobotick_b200.so.
// The vendor parser (lib/wt901c) stays where it is: its SensorDataUpdata callback (imu_if_wt901c.cpp:23-48) keeps filling
// sReg[]; the shim hands the snapshot over instead of scaling it on the MCU.
#pragma once
#include <stdint.h>
#include <string.h>

#include "robotick.h"

namespace IMT {

class IMU_IF { // imu_if_base.hpp:10-30
public:
  struct Data { // :12-18 -- 16 contiguous floats, the order rk_imt_get() returns
    float accel[3], gyro[3], mag[3], angle[3], qut[4];
  };
  virtual ~IMU_IF() {}
  virtual void  init()                  = 0;
  virtual void  update()                = 0;
  virtual bool  isComComp()             = 0;
  virtual void  getDataLatest(Data &d)  = 0;
  virtual float getYawDate()            = 0;
  virtual bool  isError()               = 0;
};

class IMU_IF_WT901C : public IMU_IF { // imu_if_wt901c.hpp
public:
  IMU_IF_WT901C() { rk_imt_create(&h_); }
  ~IMU_IF_WT901C() override { rk_imt_destroy(h_); }
  // what the vendor parser's callback maintains: the 16 registers updateData() reads, in RK_IMT_REG_* order (sReg[AX..Yaw],
  // sReg[q0..q3]), and "a quaternion frame arrived since the last isComComp()" (QUAT_UPDATE, imu_if_wt901c.cpp:44,138-141)
  void on_registers(const int16_t regs[RK_IMT_REGS], bool quat_frame) {
    memcpy(sreg_, regs, sizeof(sreg_));
    quat_update_ = quat_update_ || quat_frame;
  }
  void init() override { // .cpp:63-77: blocking first read, updateData(), latch q_init
    rk_imt_init(h_, sreg_);
    quat_update_ = false;
  }
  bool isComComp() override { // .cpp:132-143
    const bool r = quat_update_;
    quat_update_ = false;
    return r;
  }
  void update() override { rk_imt_update1(h_, sreg_, isComComp() ? 1 : 0); } // .cpp:83-89
  void getDataLatest(Data &d) override {                                      // :145-147
    int e = 0;
    rk_imt_get(h_, d.accel, &e);
  }
  float getYawDate() override { // :160-162
    float y = 0.0f;
    rk_imt_get_yaw(h_, &y);
    return y;
  }
  bool isError() override { // :164
    Data d;
    int  e = 0;
    rk_imt_get(h_, d.accel, &e);
    return e != 0;
  }
  bool ok() const { return h_ != nullptr; }

private:
  rk_imt_t *h_ = nullptr;
  int16_t   sreg_[RK_IMT_REGS] = {0};
  bool      quat_update_        = false;
};

} // namespace IMT
