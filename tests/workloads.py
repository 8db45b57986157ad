"""TEST INFRASTRUCTURE: canned workloads shared by the CPU pin tests, the golden-fixture
generator and the GPU parity tests."""
import numpy as np

from roboken_fmskf_robot_controller_b200 import _cabi, layout, streams

CMD_DT = np.dtype([("vx", "<f4"), ("vy", "<f4"), ("vth", "<f4"), ("kind", "<i4")])


def c1_inputs():
    """BASELINE.json configs[0] / SURVEY.md section 8d C1: 1 vehicle, 10 000 ticks; MOVE
    (200,100,1.0) at tick 0, STOP at tick 5000; yaw_deg = ((tick/10) mod 360) - 180."""
    cmd = np.zeros((2, 1), dtype=CMD_DT)
    cmd[0, 0] = (200.0, 100.0, 1.0, _cabi.RK_CMD_MOVE)
    cmd[1, 0] = (0.0, 0.0, 0.0, _cabi.RK_CMD_STOP)
    k = np.arange(1000)
    yaw = (((k % 360) - 180).astype(np.float32) * streams.DEG2RAD).astype(np.float32).reshape(1000, 1)
    return dict(n=1, steps=10000, cmd=cmd, seg_len=5000, yaw=yaw, yaw_period=10)


# SURVEY.md Appendix D (probe of the unmodified reference; libm-sin shim for pos)
APPENDIX_D = {
    0: ((0, 0, 0), (0, 0, 0), (0, 0, 0, 0)),
    1: ((0.00500000035, 0.00500000035, 0.000150000007), (0, 0, 0), (0, 0, 0, 0)),
    2: ((0.0200000014, 0.0200000014, 0.000600000028), (0, 0, 0), (-1, 0, -1, -2)),
    1000: ((200, 100, 1), (201.394623, 98.8019943, 4.0522871), (62, 519, -408, -855)),
    5000: ((200, 100, 1), (201.176453, 98.8019943, 4.06408024), (62, 518, -408, -857)),
    6000: ((0, 0, 0), (-0.354519993, -0.0272707716, -0.00147409493), (2, 1, 2, 1)),
    9999: ((0, 0, 0), (0, 0, 0), (1, 1, 3, 3)),
}
APPENDIX_D_POS_LIBM = (0.0650224909, -0.232292712)


def plant_inputs(n, steps, seed=0x5EED, seg_len=125, yaw_period=10, first=0):
    """C2-style closed-loop workload."""
    n_seg = (steps + seg_len - 1) // seg_len
    n_yaw = (steps + yaw_period - 1) // yaw_period
    return dict(
        n=n,
        steps=steps,
        cmd=streams.vehicle_commands(n, n_seg, seed, first),
        seg_len=seg_len,
        yaw=streams.vehicle_yaw(n, n_yaw, seed, first),
        yaw_period=yaw_period,
    )


def random_states(n, seed=1):
    """Plausible random vehicle states (finite, in-range) as AoS [n, VS_WORDS] uint32."""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, layout.VS_WORDS), dtype=np.uint32)

    def setf(col, lo, hi):
        a[:, col] = rng.uniform(lo, hi, n).astype(np.float32).view(np.uint32)

    setf(layout.VS_POS_X, -5, 5)
    setf(layout.VS_POS_Y, -5, 5)
    setf(layout.VS_POS_TH, -10, 10)
    a[:, layout.VS_FLAGS] = rng.integers(0, 2, n)
    for ax in range(3):
        b = layout.VS_INTERP0 + 12 * ax
        sc = 400.0 if ax < 2 else 19.0
        setf(b + layout.VI_VEL_NOW, -sc, sc)
        setf(b + layout.VI_ACL_NOW, -2 * sc, 2 * sc)
        setf(b + layout.VI_VEL_TGT, -sc, sc)
        setf(b + layout.VI_ACL_MAX, -2 * sc, 2 * sc)
        setf(b + layout.VI_JERK_P, -20 * sc, 20 * sc)
        setf(b + layout.VI_JERK_M, -20 * sc, 20 * sc)
        setf(b + layout.VI_DT1, 0, 0.2)
        setf(b + layout.VI_DT2, 0, 0.3)
        setf(b + layout.VI_DT3, 0, 0.2)
        setf(b + layout.VI_VEL_INI, -sc, sc)
        setf(b + layout.VI_ACL_INI, -2 * sc, 2 * sc)
        setf(b + layout.VI_DT, 0, 0.8)
    for w in range(4):
        b = layout.VS_CTRL0 + 8 * w
        setf(b + layout.VC_PREV_VAL, -1500, 1500)
        setf(b + layout.VC_INTEG, -0.5, 0.5)
        setf(b + layout.VC_LPF_Y, -1000, 1000)
        setf(b + layout.VC_LPF_X, -1000, 1000)
        setf(b + layout.VC_NOW_TGT, -1500, 1500)
        m = layout.VS_MOTOR0 + 8 * w
        s = rng.integers(-(1 << 40), 1 << 40, n, dtype=np.int64)
        d = rng.integers(-20000, 20000, n, dtype=np.int64)
        a[:, m + layout.VM_SUM_LO] = (s & 0xFFFFFFFF).astype(np.uint32)
        a[:, m + layout.VM_SUM_HI] = ((s >> 32) & 0xFFFFFFFF).astype(np.uint32)
        pv = s - d
        a[:, m + layout.VM_PREV_LO] = (pv & 0xFFFFFFFF).astype(np.uint32)
        a[:, m + layout.VM_PREV_HI] = ((pv >> 32) & 0xFFFFFFFF).astype(np.uint32)
        ang = rng.integers(0, 8192, n).astype(np.uint32)
        rpm = rng.integers(-12000, 12001, n).astype(np.int64)
        cur = rng.integers(-3000, 3001, n).astype(np.int64)
        tgt = rng.integers(-3000, 3001, n).astype(np.int64)
        dirw = 1 if w < 2 else -1
        raw_ang = ang if dirw == 1 else (8192 - ang)
        a[:, m + layout.VM_ANG_RPM] = (raw_ang & 0xFFFF) | (((rpm * dirw) & 0xFFFF).astype(np.uint32) << 16)
        a[:, m + layout.VM_CUR_TGT] = (cur & 0xFFFF).astype(np.uint32) | ((tgt & 0xFFFF).astype(np.uint32) << 16)
        a[:, m + layout.VM_USEC] = rng.integers(0, 0x7FFF, n).astype(np.uint32) | (
            rng.integers(0, 3, n).astype(np.uint32) << 16)
        a[:, m + layout.VM_PLANT] = ang | ((rpm & 0xFFFF).astype(np.uint32) << 16)
    return a
