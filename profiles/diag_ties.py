"""Diagnostic: teacher-forced CUDA step on a golden case, print every shield-record / state difference.
python profiles/diag_ties.py ties_mass_td3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import marl_mass_b200 as mm
import oracle as orc
from helpers import load_golden, SH_I, SH_F, I32_FIELDS, F64_FIELDS, used_mask
from test_gpu_parity import env_config

np.set_printoptions(linewidth=220, precision=6, suppress=True)
name = sys.argv[1]
g, cfg = load_golden(name)
rows = g["row_of_step"]
T = len(rows)
env = mm.MergeEnvBatched(T, env_config(cfg), record_diag=True)
pre = orc.state_from_golden(g, rows)
env.set_state(pre)
env.step(torch.from_numpy(np.ascontiguousarray(g["act"])).cuda())
post = env.get_state()
want = orc.state_from_golden(g, rows + 1)
diag = env.shield_diag()
ran = g["sh_ran"] == 1
shown = 0
for k in SH_I:
    bad = np.argwhere((diag[k] != g["sh_" + k]) & ran)
    for t, s, v in bad[:12]:
        print("shield", k, "step", t, "sub", s, "veh", v, "cuda", diag[k][t, s, v], "ref", g["sh_" + k][t, s, v],
              "| roles cuda", [int(diag[r][t, s, v]) for r in ("leader", "front_adj", "rear_adj")],
              "ref", [int(g["sh_" + r][t, s, v]) for r in ("leader", "front_adj", "rear_adj")],
              "margin", diag["lc_margin"][t, s, v])
for k in SH_F:
    err = np.abs(diag[k] - g["sh_" + k]) * ran
    bad = np.argwhere(err > 1e-9)
    for t, s, v in bad[:12]:
        print("shield", k, "step", t, "sub", s, "veh", v, "cuda", diag[k][t, s, v], "ref", g["sh_" + k][t, s, v])
m = used_mask(want)
for k in I32_FIELDS:
    bad = np.argwhere((post[k] != want[k]) & m)
    for t, v in bad[:8]:
        print("state", k, "step", t, "veh", v, "cuda", post[k][t, v], "ref", want[k][t, v])
for k in F64_FIELDS:
    err = np.abs(post[k] - want[k]) * m
    bad = np.argwhere(err > 1e-9)
    for t, v in bad[:8]:
        print("state", k, "step", t, "veh", v, "cuda", post[k][t, v], "ref", want[k][t, v])
print("done", name)
