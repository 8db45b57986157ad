#!/usr/bin/env python3
"""Where does the end-to-end step spend its time?  Variants of bench.py's e2e loop on one GPU:
   full | no H2D | no zero_ | no D2H, each timed with CUDA events over 20 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import roboken_fmskf_robot_controller_b200 as rk
from roboken_fmskf_robot_controller_b200 import _cabi, streams
from roboken_fmskf_robot_controller_b200.vehicle import VehicleBatch

lib = rk.load()
dev = torch.device("cuda", 0)
n, T, K = 1 << 20, 1000, 20
cmd_h = torch.from_numpy(streams.vehicle_commands(n, 8, 0x5EED, 0).view(np.int32).reshape(8, n, 4)).pin_memory()
yaw_h = torch.from_numpy(streams.vehicle_yaw_reg(n, 100, 0x5EED, 0)).pin_memory()
goal_d = torch.zeros((n, 2), dtype=torch.float32, device=dev)
vb = VehicleBatch(n, dev)
copy_s, back_s, comp_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
bufs = []
for b in range(2):
    c, y = cmd_h.to(dev), yaw_h.to(dev)
    co = torch.zeros(n, dtype=torch.float32, device=dev)
    bufs.append(dict(cmd=c, yaw=y, cost=co, cost_h=torch.empty(n, dtype=torch.float32).pin_memory(),
                     args=vb.make_args(T, sensor_mode=_cabi.RK_SENSOR_PLANT, cmd=c, seg_len=125, yaw=y, yaw_period=10, goal=goal_d, cost=co),
                     up=torch.cuda.Event(), done=torch.cuda.Event(), down=torch.cuda.Event()))

def step(s, h2d=True, zero=True, d2h=True):
    b = bufs[s % 2]
    with torch.cuda.stream(copy_s):
        copy_s.wait_event(b["done"])
        if h2d:
            b["cmd"].copy_(cmd_h, non_blocking=True)
            b["yaw"].copy_(yaw_h, non_blocking=True)
        b["up"].record(copy_s)
    with torch.cuda.stream(comp_s):
        comp_s.wait_event(b["up"])
        comp_s.wait_event(b["down"])
        if zero:
            vb.state.zero_()
        vb.rollout_args(b["args"], stream=comp_s)
        b["done"].record(comp_s)
    with torch.cuda.stream(back_s):
        back_s.wait_event(b["done"])
        if d2h:
            b["cost_h"].copy_(b["cost"], non_blocking=True)
        b["down"].record(back_s)

for name, kw in (("full", {}), ("no_h2d", dict(h2d=False)), ("no_zero", dict(zero=False)), ("no_d2h", dict(d2h=False)),
                 ("kernel_only", dict(h2d=False, zero=False, d2h=False))):
    for s in range(4):
        step(s, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in (copy_s, back_s, comp_s):
        st.wait_stream(torch.cuda.current_stream())
    for s in range(K):
        step(s, **kw)
    for st in (copy_s, back_s, comp_s):
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:12s} {e0.elapsed_time(e1) / K:.3f} ms/step", flush=True)
# raw H2D time
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    bufs[0]["cmd"].copy_(cmd_h, non_blocking=True)
    bufs[0]["yaw"].copy_(yaw_h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print(f"h2d alone    {e0.elapsed_time(e1) / 10:.3f} ms per step's tables ({(cmd_h.numel() * 4 + yaw_h.numel() * 2) / 1e6:.0f} MB)")
