"""Device entry point of the baseline supervisors (mm_supervise): the kernel wrapper around csrc/supervisor_core.h must
return, for every step of the reference fixtures, the action tuple the reference's `safety_supervisor` /
`safety_layer_dmc` handed to _simulate (same scenes, same policy tuples, same np.random.rand() draws).

Both cases (1 269 steps) pass on a B200 with the committed build (big helpers as device calls; round-1 driver run:
XPASS x 2), so this is a regular, strict test.  The step path itself calls the supervisors when safety_guarantee is
"priority" / "dmc" (tests/test_gpu_parity.py::test_supervised_step_*)."""
import numpy as np
import pytest

from conftest import SUPERVISED_CASES
from helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", SUPERVISED_CASES)
def test_device_supervisor_reproduces_the_reference_tuples(name):
    import torch
    import marl_mass_b200 as mm
    import oracle as orc
    from test_gpu_parity import env_config
    g, cfg = load_golden(name)
    rows = g["row_of_step"]
    T = len(rows)
    env = mm.MergeEnvBatched(T, dict(env_config(cfg), safety_guarantee="none"))   # the supervisor alone: mm_supervise
    try:
        env.set_state(orc.state_from_golden(g, rows))
        draws = np.zeros((T, 32))
        draws[:, :16] = np.nan_to_num(g["rand_draws"])
        got, used = env.supervise(torch.from_numpy(np.ascontiguousarray(g["act"])).cuda(), cfg["safety_guarantee"],
                                  torch.from_numpy(draws).cuda(), return_used=True)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
        assert np.array_equal(used.cpu().numpy(), np.sum(~np.isnan(g["rand_draws"]), axis=1))
        live = np.arange(12)[None, :] < g["st_n_cav"][rows][:, None]
        bad = np.argwhere((got != g["new_act"]) & live)
        assert len(bad) == 0, (len(bad), bad[:5].tolist())
        assert np.array_equal(got[~live], g["act"][~live])
    finally:
        env.close()


@pytest.mark.parametrize("cap", [0, 7])
def test_dmc_with_a_short_task_list(cap):
    """The dmc supervisor defers predicted collisions to a task list (supervisor.cu, two kernels); collisions that do not
    fit are evaluated in place by the scan.  MM_SUP_TASK_CAP shrinks the list (0: one kernel, everything in place; 7:
    seven deferred, the rest in place) - the tuples must not change.  The knob is read once per process: subprocess."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    if "MM_SUP_TASK_CAP" in os.environ:
        pytest.skip("inner run")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_zz_supervisor_gpu.py"), "-q", "-x", "-m", "gpu",
                        "-k", "reproduces_the_reference_tuples and dmc"],
                       env=dict(os.environ, MM_SUP_TASK_CAP=str(cap)), capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0 and "1 passed" in p.stdout, (p.stdout[-1500:], p.stderr[-1500:])
