"""A saved soak scene (soak_repro.py; e.g. tests/golden/ulp_tie_scene.npz 1) advanced by ONE sub-step on the GPU and by the oracle: prints x of every vehicle
on both sides (policy_frequency = simulation_frequency makes a policy step a single sub-step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "profiles")):
    sys.path.insert(0, p)
import numpy as np
import torch
import marl_mass_b200 as mm
import oracle as orc
from soak_cases import CASES

d = np.load(sys.argv[1])
ci = int(sys.argv[2])
shield, traffic, td, reward, lateral = CASES[ci]
cfg = dict(mm.DEFAULT_CONFIG, safety_guarantee=shield, lateral_control=lateral, traffic_type=traffic, traffic_density=td,
           agent_reward=reward, HEADWAY_TIME=0.5, cbf_eta=0.03125, HIGH_SPEED_REWARD=4, HEADWAY_COST=1, MERGING_LANE_COST=8)
cfg["policy_frequency"] = cfg["simulation_frequency"]
cfg["duration"] = 10   # 150 one-sub-step policy steps: inside the 255-step limit
E = 128
env = mm.MergeEnvBatched(E, cfg, record_diag=True)
env.reset(seed=1)
st = env.get_state()
for k in d.files:
    if k.startswith("pre_"):
        st[k[4:]][:] = np.array(d[k])[None]
env.set_state(st)
a = np.repeat(d["actions"][None].astype(np.int8), E, axis=0)
ocfg = orc.make_config(cfg)
orc.step(ocfg, st, a, n_threads=1)
env.step(torch.from_numpy(a).cuda())
post = env.get_state()
n = int(d["pre_n_veh"])
for i in range(n):
    print(i, repr(float(post["x"][0][i])), repr(float(st["x"][0][i])), "same" if post["x"][0][i] == st["x"][0][i] else "DIFFERENT",
          "heading", repr(float(post["heading"][0][i])), repr(float(st["heading"][0][i])))
