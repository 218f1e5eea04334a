import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


GOLDEN_CASES = ["unsafe_td1", "unsafe_td3", "unsafe_td2_mixed", "hss_td3", "hss_td3_mixed", "mass_td1",
                "mass_td3_srew", "mass_td3_mixed", "mass_td2_mixed_mrew",
                # lateral_control = steer_vel, and the 5x5-observation env id merge-multi-agent-v05
                "steervel_unsafe_td2", "steervel_hss_td3_mixed", "steervel_mass_td2", "v05_steervel_unsafe_td1",
                # traffic_type = av: one shielded CAV among HDVs
                "mass_td3_av"]
V0_CASES = ["v0_unsafe_td1", "v0_unsafe_td2_mixed"]
# env merge-multi-agent-hdv-v1 (MergeEnvLCHDV, traffic_type = hdv): every vehicle observed, nobody controlled
HDV_CASES = ["hdv_td3"]
HDV_TIE_CASES = ["ties_hdv_td3"]      # the same env with x / y / speed snapped to grid values before every step
# x positions and speeds snapped to integers before every policy step: exact ties in x, s and the |ds| sort keys
# (the reference's stable-sort / "<=" tie rules); every step is its own (pre-state, post-state) pair
# (ties_y_*: y snapped to a 0.5 m grid as well - vehicles exactly between two lanes, closest-lane argmin ties)
# (ties_half_*: x on a 0.5 m grid - vehicles exactly on the strict after_end thresholds 217.5 / 317.5 / 417.5)
TIE_CASES = ["ties_mass_td3", "ties_hss_td3_mixed", "ties_y_mass_td3_mixed", "ties_half_mass_td3_mixed"]
V0_TIE_CASES = ["ties_v0_unsafe_td2_mixed"]      # the same snapping on env merge-multi-agent-v0 (teacher-forced only)
# baseline supervisors priority / dmc on env v0: the policy's actions, the supervised actions the env executed and the
# supervisor's random draws (the supervisor itself is not built: DESIGN.md section 8)
SUPERVISED_CASES = ["priority_v0_td3_mixed", "dmc_v0_td3_mixed"]
