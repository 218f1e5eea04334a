import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import marl_mass_b200 as mm, oracle as orc
from helpers import F64_FIELDS, I32_FIELDS, rel_err, used_mask
shield, traffic, td, reward = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
E, T = 2048, 100
cfg = dict(mm.DEFAULT_CONFIG, safety_guarantee=shield, traffic_type=traffic, traffic_density=td, agent_reward=reward,
           HEADWAY_TIME=0.5, cbf_eta=0.03125, HIGH_SPEED_REWARD=4, HEADWAY_COST=1, MERGING_LANE_COST=8)
env = mm.MergeEnvBatched(E, cfg, record_diag=True)
env.reset(seed=1234 + td)
st = env.get_state()
ocfg = orc.make_config(cfg)
rng = np.random.RandomState(7)
alive = np.ones(E, bool)
for t in range(T):
    a = rng.randint(0, 5, size=(E, 12)).astype(np.int8)
    want = orc.step(ocfg, st, a, n_threads=8)
    env.step(torch.from_numpy(a).cuda())
    post = env.get_state()
    m = used_mask(st) & alive[:, None]
    worst = (0, None)
    for k in F64_FIELDS:
        if k == "rec1_x": continue
        err = rel_err(post[k], st[k]) * m
        if err.max() > worst[0]:
            e, i = np.unravel_index(err.argmax(), err.shape)
            worst = (float(err.max()), (k, int(e), int(i)))
    nd = sum(int(((post[k] != st[k]) & m).sum()) for k in I32_FIELDS)
    k, e, i = worst[1]
    print("t=%3d worst %.2e %-10s env %4d slot %2d kind %d speed %.4f x %.2f lane %d crashed %d | discrete mismatches %d alive %d" % (
        t, worst[0], k, e, i, st["kind"][e, i], st["speed"][e, i], st["x"][e, i], st["lane"][e, i], st["crashed"][e, i], nd, alive.sum()))
    alive &= want["done"] == 0
