#!/usr/bin/env python3
"""Kernel-level measurements of the two slow-rate modules (BASELINE configs[2] and [3]) -- one JSON
line each, CUDA-event timed with inputs resident in HBM, roofline against MEASURED_PEAKS.json.

    python tools/bench_modules.py [--n 1048576] [--imu-updates 64] [--arm-ticks 1000] [--reps 10] [--only imu|wire|guard|stream|arm]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from roboken_fmskf_robot_controller_b200 import layout, streams  # noqa: E402
from roboken_fmskf_robot_controller_b200.arm import ArmBatch  # noqa: E402
from roboken_fmskf_robot_controller_b200.imu import ImuBatch  # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bench_imu(a, dev):
    n, K = a.n, a.imu_updates
    uniq = min(n, 1 << 16)
    regs, have = streams.imu_samples(uniq, K, seed=3, drop_every=64)
    regs_d = torch.from_numpy(np.tile(streams.imu_cells(regs), (1, 1, n // uniq, 1))).to(dev)
    have_d = torch.from_numpy(np.tile(have, (1, n // uniq))).to(dev)
    ib = ImuBatch(n, dev)
    ib.update(regs_d[:1].contiguous(), None, None, do_init=True)
    out = torch.empty((K, 4, n, 4), dtype=torch.float32, device=dev)
    peak, src = hbm_peak()
    for mode in ("full_output", "state_only"):
        ms = timed(lambda: ib.update(regs_d, have_d, out if mode == "full_output" else None), a.reps)
        per = 32 + 1 + (64 if mode == "full_output" else 0)
        nbytes = n * (K * per + 2 * 96)
        print(json.dumps({"kernel": "rk::imt_update_kernel", "workload": f"configs[2]: {n} IMUs x {K} fused updates, {mode}",
                          "updates_per_s": n * K / (ms * 1e-3), "ms_per_launch": ms,
                          "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "peak_source": src,
                                       "algorithmic_bytes_per_update": per, "algorithmic_bytes_per_launch": nbytes}}), flush=True)


def bench_wire(a, dev):
    """rk_imt_feed_bytes: 55 serial bytes (five WIT frames) per update in four 128-bit cells through the parser."""
    n, K, ncells = a.n, a.imu_updates, 4
    uniq = min(n, 1 << 14)
    regs, _ = streams.imu_samples(uniq, K, seed=3)
    cells, nb = streams.imu_wire_clean(regs, ncells=ncells)
    cells_d = torch.from_numpy(np.tile(cells, (1, 1, n // uniq, 1)).view(np.int32)).to(dev)
    nb_d = torch.from_numpy(np.tile(nb, (1, n // uniq)).view(np.int16)).to(dev)
    ib = ImuBatch(n, dev)
    ib.feed_bytes(cells_d[:1].contiguous(), nb_d[:1].contiguous(), None, None, do_init=True)
    out = torch.empty((K, 4, n, 4), dtype=torch.float32, device=dev)
    peak, src = hbm_peak()
    for mode in ("full_output", "state_only"):
        ms = timed(lambda: ib.feed_bytes(cells_d, nb_d, out if mode == "full_output" else None, None), a.reps)
        per = ncells * 16 + 2 + (64 if mode == "full_output" else 0)
        nbytes = n * (K * per + 2 * (96 + 48))
        print(json.dumps({"kernel": "rk::imt_feed_bytes_kernel", "workload": f"8f-3: {n} IMUs x {K} updates x 55 wire bytes in {ncells} cells, {mode}",
                          "updates_per_s": n * K / (ms * 1e-3), "wire_bytes_per_s": n * K * 55 / (ms * 1e-3), "ms_per_launch": ms,
                          "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "peak_source": src,
                                       "algorithmic_bytes_per_update": per, "algorithmic_bytes_per_launch": nbytes}}), flush=True)


def bench_guard(a, dev):
    """rk_rmt_guard: 48 B command + floor record in, 16 B vehicle message + 4 B abort word out per manager cycle."""
    from roboken_fmskf_robot_controller_b200.rmt import ManagerBatch

    n, K = a.n, a.imu_updates
    uniq = min(n, 1 << 14)
    inp = streams.rm_inputs(uniq, K, seed=3)
    inp_d = torch.from_numpy(np.tile(inp, (1, 1, n // uniq, 1)).view(np.int32)).to(dev)
    mb = ManagerBatch(n, dev)
    cmd = torch.empty((K, n, 4), dtype=torch.int32, device=dev)
    ab = torch.empty((K, n), dtype=torch.int32, device=dev)
    peak, src = hbm_peak()
    ms = timed(lambda: mb.guard(inp_d, cmd, ab), a.reps)
    per = 48 + 16 + 4
    nbytes = n * (K * per + 2 * 16)
    print(json.dumps({"kernel": "rk::rmt_guard_kernel", "workload": f"8f-2: {n} managers x {K} cycles (command + 8 floor sensors -> guarded vehicle message)",
                      "cycles_per_s": n * K / (ms * 1e-3), "ms_per_launch": ms,
                      "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "peak_source": src,
                                   "algorithmic_bytes_per_cycle": per, "algorithmic_bytes_per_launch": nbytes}}), flush=True)


def bench_stream(a, dev):
    """rk_vdt_rollout in RK_SENSOR_STREAM mode: 32 B of M2006 feedback frames per vehicle tick from HBM."""
    from roboken_fmskf_robot_controller_b200 import _cabi
    from roboken_fmskf_robot_controller_b200.vehicle import VehicleBatch

    n, T = a.n, 200
    uniq = min(n, 1 << 12)
    fr = streams.vehicle_frames(uniq, T, seed=3)
    fr_d = torch.from_numpy(np.tile(fr, (1, 1, n // uniq)).view(np.int64)).to(dev)
    cmd = streams.vehicle_commands(n, 2, seed=3)
    cmd_d = torch.from_numpy(cmd.view(np.int32).reshape(2, n, 4)).to(dev)
    yaw_d = torch.from_numpy(streams.vehicle_yaw(n, T // 10, seed=3)).to(dev)
    vb = VehicleBatch(n, dev)
    args = vb.make_args(T, sensor_mode=_cabi.RK_SENSOR_STREAM, cmd=cmd_d, seg_len=100, yaw=yaw_d, yaw_period=10, frames=fr_d)
    peak, src = hbm_peak()
    ms = timed(lambda: vb.rollout_args(args), a.reps)
    nbytes = n * (T * 32 + 2 * 448)
    print(json.dumps({"kernel": "rk::vdt_rollout_stream_fast_kernel", "workload": f"configs[1] streamed sensors: {n} vehicles x {T} ticks, 32 B of CAN frames per tick",
                      "steps_per_s": n * T / (ms * 1e-3), "ms_per_launch": ms,
                      "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "peak_source": src,
                                   "algorithmic_bytes_per_tick": 32, "algorithmic_bytes_per_launch": nbytes,
                                   "fp32_tflops": 183 * n * T / (ms * 1e-3) / 1e12}}), flush=True)


def bench_arm(a, dev):
    n, K = a.n, a.arm_ticks
    seq = torch.from_numpy(layout.aos_to_soa(streams.arm_sequences(n, seed=0xC4, seq_id=9, max_len=32)).view(np.int32)).to(dev)
    ab = ArmBatch(n, dev)

    def one():
        ab.mode_init()
        ab.push_cmdseq(seq)
        ab.update(K)

    ms_all = timed(one, a.reps)
    ms_setup = timed(lambda: (ab.mode_init(), ab.push_cmdseq(seq)), a.reps)
    ms = ms_all - ms_setup
    print(json.dumps({"kernel": "rk::adt_update_kernel<false>", "workload": f"configs[3]: {n} arms x {K} fused 100 Hz ticks, one PosCmdSeq each",
                      "arm_ticks_per_s": n * K / (ms * 1e-3), "ms_per_launch": ms, "ms_init_plus_push": ms_setup,
                      "roofline": {"bound": "issue", "achieved": 46 * n * K / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                   "algorithmic_flop_per_tick": 46, "note": "no per-tick HBM traffic; see DESIGN.md 3.4"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--imu-updates", type=int, default=64)
    ap.add_argument("--arm-ticks", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    if a.only in ("", "imu"):
        bench_imu(a, dev)
    if a.only in ("", "wire"):
        bench_wire(a, dev)
    if a.only in ("", "guard"):
        bench_guard(a, dev)
    if a.only in ("", "stream"):
        bench_stream(a, dev)
    if a.only in ("", "arm"):
        bench_arm(a, dev)


if __name__ == "__main__":
    main()
