"""GPU parity: ADTModePositioning kernels (rk_adp_*) vs the oracle port / compiled reference."""
import numpy as np
import pytest
import torch

import oracle_lib as ol
from roboken_fmskf_robot_controller_b200 import layout
from roboken_fmskf_robot_controller_b200.arm import ArmBatch, ArmPositioningBatch
from test_armpos_cpu import bringup_state, pos_script, run_pos_script

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def gpu_pos_script(n, script, astate):
    ab = ArmBatch(n, DEV)
    ab.load_state_soa(astate)
    pb = ArmPositioningBatch(ab)
    outs = []
    for step in script:
        if step[0] == "init":
            pb.mode_init()
        elif step[0] == "push":
            pb.push_cmd(torch.from_numpy(step[1].view(np.int32)).to(DEV), None if step[2] is None else torch.from_numpy(step[2]).to(DEV))
        elif step[0] == "update":
            tr = torch.zeros((step[1], layout.ADT_TRACE_WORDS, n), dtype=torch.int32, device=DEV)
            pb.update(step[1], tr)
            outs.append(tr.cpu().numpy().view(np.uint32))
        elif step[0] == "status":
            outs.append(pb.cmd_status(torch.from_numpy(np.asarray(step[1], dtype=np.uint32).view(np.int32)).to(DEV)).cpu().numpy())
        torch.cuda.synchronize()
        outs += [ab.state_host().copy(), pb.pstate.cpu().numpy().view(np.uint32)]
    return outs


@pytest.mark.parametrize("n,seed", [(1, 1), (300, 2), (4100, 3)])
def test_positioning_mode_vs_port(n, seed):
    st0 = bringup_state(n, seed)
    sc = pos_script(n, seed)
    got, exp = gpu_pos_script(n, sc, st0), run_pos_script("port", n, sc, st0)[2]
    assert len(got) == len(exp)
    for k, (x, y) in enumerate(zip(got, exp)):
        np.testing.assert_array_equal(x, y, err_msg=f"output {k}")


@pytest.mark.skipif(not ol.have_ref("libref_arm.so"), reason="oracle/_ref/libref_arm.so not present")
def test_positioning_mode_vs_compiled_reference():
    n = 48
    st0 = bringup_state(n, 9)
    sc = pos_script(n, 9)
    got, exp = gpu_pos_script(n, sc, st0), run_pos_script("ref", n, sc, st0)[2]
    for k, (x, y) in enumerate(zip(got, exp)):
        np.testing.assert_array_equal(x, y, err_msg=f"output {k}")
