// adt_shim.hpp -- what a firmware maintainer compiles INSTEAD of src/ArmDrive/AD_mode_positioning_seq.{hpp,cpp} and the
// joint classes it drives: the ADTModeBase vtable (AD_mode_base.hpp:17-25), push_cmdseq / get_q_cmdseq_status
// (AD_mode_positioning_seq.hpp:33-40) and the servo rx callbacks, same names and argument meaning, over the C-ABI.
#pragma once
#include <stdint.h>

#include "robotick.h"

namespace ADT {

class ADTModeBase { // AD_mode_base.hpp:12-31
public:
  virtual ~ADTModeBase() {}
  void init() { // :19-22
    is_comp = false;
    doInit();
  }
  virtual void doInit()        = 0;
  virtual void update()        = 0;
  virtual void end()           = 0;
  virtual bool isCompleted() { return is_comp; }

protected:
  bool is_comp = false;
};

class ADTModePositioningSeq : public ADTModeBase { // AD_mode_positioning_seq.hpp
public:
  struct PosCmd { // :15-18 == rk_adt_poscmd_t (24 B)
    uint32_t u32_dt_ms;
    float    fl_tgt_pos_deg[5];
  };
  struct PosCmdSeq { // :20-24 == rk_adt_poscmdseq_t (776 B)
    uint32_t u32_id;
    uint8_t  u8_cmd_seq_len;
    PosCmd   cmd_seq[32];
  };
  enum { PROCESSING = 0, DONE = 1, NO_DATA = 99 }; // :36-40

  ADTModePositioningSeq() { rk_adt_create(&h_, nullptr); } // nullptr: the joint constants of AD_task_main.cpp:38-116
  ~ADTModePositioningSeq() override { rk_adt_destroy(h_); }
  void doInit() override { rk_adt_init(h_); }  // .cpp:5-11 (+ the flags a finished INIT mode leaves, AD_mode_initialize.cpp:133-135)
  void update() override { rk_adt_tick(h_); }  // .cpp:13-18 + the five joint update() calls of one ADT::main loop body (:213-228)
  void end() override {}
  void push_cmdseq(PosCmdSeq &c) { // .cpp:124-137; a full ring drops the sequence silently
    static_assert(sizeof(PosCmdSeq) == sizeof(rk_adt_poscmdseq_t), "layout");
    rk_adt_push(h_, reinterpret_cast<const rk_adt_poscmdseq_t *>(&c));
  }
  int32_t get_q_cmdseq_status(uint32_t id) { // .cpp:146-184
    int32_t s = NO_DATA;
    rk_adt_status(h_, id, &s);
    return s;
  }
  // JointBase::get_tgt_deg() of J0..J4 (AD_joint_base.hpp:47 / DfGear overrides)
  void get_tgt_deg(float out[5]) { rk_adt_get_targets_deg(h_, out); }
  // CAN rx: JointMyBldcServo::rx_callback(cmdid, msg) for DF_Left / DF_Right / P3 (slot 0..2), JointMgServo::rx_callback(msg)
  void bldc_rx_callback(int slot, uint32_t cmdid, const uint8_t msg[8]) { rk_adt_rx(h_, slot, cmdid, msg, nullptr); }
  void mg_rx_callback(const uint8_t msg[8]) { rk_adt_rx(h_, 3, 0, msg, nullptr); }
  bool ok() const { return h_ != nullptr; }
  rk_adt_t *h_ = nullptr;
};

} // namespace ADT
