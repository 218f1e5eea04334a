"""One case of one soak round, stopping at the first shield-record mismatch and saving that scene:
python profiles/soak_repro.py <round> <case index> [snap] -> gpurun_out/soak_repro_<round>_<case>.npz
(the pre-step state of the env, its actions, the oracle's and the kernel's shield records; MM_LIB_PATH selects a build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "profiles")):
    sys.path.insert(0, p)
import numpy as np
import torch
import marl_mass_b200 as mm
import oracle as orc
from helpers import used_mask, LC_BOUNDARY_EPS, SH_I
from soak_cases import CASES

rnd, ci = int(sys.argv[1]), int(sys.argv[2])
snap = len(sys.argv) > 3 and sys.argv[3] == "snap"
E, T = int(os.environ.get("MM_SOAK_ENVS", "4096")), 100
mm.set_step_variant((0, 4, 7, 3)[rnd % 4])
shield, traffic, td, reward, lateral = CASES[ci]
cfg = dict(mm.DEFAULT_CONFIG, safety_guarantee=shield, lateral_control=lateral, traffic_type=traffic, traffic_density=td,
           agent_reward=reward, HEADWAY_TIME=0.5, cbf_eta=0.03125, HIGH_SPEED_REWARD=4, HEADWAY_COST=1, MERGING_LANE_COST=8)
env = mm.MergeEnvBatched(E, cfg, record_diag=True)
env.reset(seed=900000 + 1000 * rnd + ci)
st = env.get_state()
ocfg = orc.make_config(cfg)
rng = np.random.RandomState(100 * rnd + ci)
alive, clean = np.ones(E, bool), np.ones(E, bool)
n_bad = 0
for t in range(T):
    if snap:
        um = used_mask(st)
        st["x"] = np.where(um, np.round(st["x"] * 2) / 2, st["x"])
        st["y"] = np.where(um, np.round(st["y"] * 2) / 2, st["y"])
        st["speed"] = np.where(um, np.round(st["speed"] / 2.5) * 2.5 if t % 2 else np.round(st["speed"]), st["speed"])
        h1 = um & (st["hist_len"] >= 1)
        st["rec1_x"] = np.where(h1, st["x"], st["rec1_x"])
        st["rec1_vx"] = np.where(h1, st["speed"] * np.cos(st["heading"]), st["rec1_vx"])
        env.set_state(st)
    pre = {k: np.array(v, copy=True) for k, v in st.items()}
    a = rng.randint(0, 5, size=(E, 12)).astype(np.int8)
    want = orc.step(ocfg, st, a, n_threads=16)
    env.step(torch.from_numpy(a).cuda())
    diag = env.shield_diag()
    sel = alive & clean
    ran = (want["sh_ran"] == 1) & sel[:, None, None]
    bnd = ran & (diag["lc_margin"] < LC_BOUNDARY_EPS)
    for k in SH_I:
        bad = (diag[k] != want["sh_" + k]) & ran & ~bnd
        if bad.any():
            e = int(np.argwhere(bad)[0][0])
            print("build", env.step_build(), "step", t, "field", k, "first", np.argwhere(bad)[:4].tolist())
            out = os.path.join(ROOT, "gpurun_out", "soak_repro_%d_%d.npz" % (rnd, ci))
            np.savez(out, t=t, env=e, actions=a[e], **{"pre_" + kk: vv[e] for kk, vv in pre.items()},
                     **{"want_" + kk: np.asarray(want["sh_" + kk])[e] for kk in SH_I},
                     **{"got_" + kk: diag[kk][e] for kk in SH_I}, got_lc_margin=diag["lc_margin"][e])
            print("saved", out)
            n_bad += 1
            break
    if n_bad:
        break
    alive &= want["done"] == 0
    clean &= ~((st["speed"] < 3.0) & used_mask(st)).any(axis=1)
print("mismatch" if n_bad else "no mismatch", "round", rnd, "case", CASES[ci], "build", env.step_build())
