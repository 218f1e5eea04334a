"""Per-step device time right after spawn vs in steady state (python profiles/time_steps.py workload envs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import marl_mass_b200 as mm
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "mass_td3"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
cfg = dict(mm.DEFAULT_CONFIG, **WORKLOADS[name]["cfg"])
env = mm.MergeEnvBatched(E, cfg)
env.reset(seed=1)
gen = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.randint(0, 5, (E, mm.MAXV), generator=gen, device="cuda", dtype=torch.int8) for _ in range(4)]
times = []
for t in range(130):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(acts[t % 4], auto_reset=True); b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
import numpy as _np
print(name, E, "steady(100..129) mean %.3f" % _np.mean(times[100:130]), "ms/step at t=0,1,2,5,10,20,40,60,80,99,100,110,129:", [round(times[i], 3) for i in (0, 1, 2, 5, 10, 20, 40, 60, 80, 99, 100, 110, 129)])
