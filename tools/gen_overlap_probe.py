#!/usr/bin/env python3
"""What does expanding the next chunk's streams cost beside a running rk_tick_rollout?  One chunk: rollout alone, the three
generators alone (uncapped / capped grid), both together (generators on a high-priority stream)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from roboken_fmskf_robot_controller_b200 import _cabi, layout  # noqa: E402
from roboken_fmskf_robot_controller_b200.devstreams import DeviceStreams  # noqa: E402
from roboken_fmskf_robot_controller_b200.robot import RobotBatch  # noqa: E402

dev = torch.device("cuda", 0)
lib = _cabi.load()
n, T, slow = 1 << 20, 1000, 10
n_seg, n_slow = 8, 100
ds = DeviceStreams(dev, first_update=1)


def tables():
    return dict(cmd=torch.empty((n_seg, n, 4), dtype=torch.int32, device=dev), regs=torch.empty((n_slow, 2, n, 8), dtype=torch.int16, device=dev),
                have=torch.empty((n_slow, n), dtype=torch.uint8, device=dev), seq=torch.empty(layout.ACMD_SLOT_WORDS * n, dtype=torch.int32, device=dev))


def generate(t, stream, which="cia"):
    with torch.cuda.stream(stream):
        if "c" in which:
            ds.vehicle_commands(t["cmd"], stream=stream)
        if "i" in which:
            ds.imu_samples(t["regs"], t["have"], stream=stream)
        if "a" in which:
            ds.arm_sequences(t["seq"], stream=stream)


ta, tb = tables(), tables()
main = torch.cuda.current_stream(dev)
generate(ta, main), generate(tb, main)
rb = RobotBatch(n, dev)
rb.imu.update(ta["regs"][:1].contiguous(), None, None, do_init=True)
args = rb.make_args(T, slow, cmd=ta["cmd"], seg_len=125, regs=ta["regs"], have_quat=ta["have"], yaw=torch.zeros(n, dtype=torch.float32, device=dev))
roll_s, gen_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps):
        fn()
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def roll():
    roll_s.wait_stream(main)
    rb.arm.mode_init(stream=roll_s)
    rb.arm.push_cmdseq(ta["seq"], stream=roll_s)
    rb.rollout_args(args, stream=roll_s)
    main.wait_stream(roll_s)


def gen(which="cia"):
    gen_s.wait_stream(main)
    generate(tb, gen_s, which)
    main.wait_stream(gen_s)


def both(which="cia"):
    roll_s.wait_stream(main), gen_s.wait_stream(main)
    rb.arm.mode_init(stream=roll_s)
    rb.arm.push_cmdseq(ta["seq"], stream=roll_s)
    rb.rollout_args(args, stream=roll_s)
    generate(tb, gen_s, which)
    main.wait_stream(roll_s), main.wait_stream(gen_s)


t_roll = timed(roll)
print(f"rollout alone                      {t_roll:8.3f} ms")
for cap in (0, 2):
    lib.rk_set_option(_cabi.RK_OPT_STREAM_CTAS, cap)
    for which in ("cia", "i", "a", "c"):
        tg = timed(lambda: gen(which))
        tb_ = timed(lambda: both(which))
        print(f"cap {cap} generators {which:4s} alone {tg:7.3f} ms   rollout + generators {tb_:8.3f} ms   cost beside the rollout {tb_ - t_roll:7.3f} ms")
