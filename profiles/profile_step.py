"""Small driver for ncu: N auto-reset steps of one workload (python profiles/profile_step.py [workload] [envs] [steps])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import marl_mass_b200 as mm  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mass_td3"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cfg = dict(mm.DEFAULT_CONFIG, **WORKLOADS[name]["cfg"])
env = mm.MergeEnvBatched(E, cfg)
env.reset(seed=1)
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(steps):
    a = torch.randint(0, 5, (E, mm.MAXV), generator=gen, device="cuda", dtype=torch.int8)
    env.step(a, auto_reset=True)
torch.cuda.synchronize()
print("ok", env.stats()["agent_steps"])
