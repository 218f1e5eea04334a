"""Step time of the 3- vs 4-CTAs-per-SM builds at a given batch: python profiles/time_variants.py workload envs"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import marl_mass_b200 as mm
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "mass_td3"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
cfg = dict(mm.DEFAULT_CONFIG, **WORKLOADS[name]["cfg"])
gen = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.randint(0, 5, (E, mm.MAXV), generator=gen, device="cuda", dtype=torch.int8) for _ in range(4)]
for variant in (3, 4, 0):
    mm.set_step_variant(variant)
    env = mm.MergeEnvBatched(E, cfg)
    env.reset(seed=1)
    times = []
    for t in range(100):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(acts[t % 4], auto_reset=True); b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    print("%s %d variant %d: mean over the episode %.3f ms/step (t=5: %.3f, t=60: %.3f)" % (name, E, variant, np.mean(times[1:]), times[5], times[60]))
    env.close()
mm.set_step_variant(0)
