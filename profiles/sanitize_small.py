"""Small end-to-end run for compute-sanitizer (one tool per gpurun call): device steps with auto-reset, both host paths,
the fused actor draw, the return kernel, get/set_state, the QP kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import marl_mass_b200 as mm
from marl_mass_b200 import rollout

for traffic, diag in (("cav", False), ("mixed", True)):
    E = 1000   # not a multiple of the tile
    cfg = dict(mm.DEFAULT_CONFIG, safety_guarantee="cbf-cav", traffic_density=3, traffic_type=traffic, HEADWAY_TIME=0.5,
               cbf_eta=0.03125)
    env = mm.MergeEnvBatched(E, cfg, record_diag=diag)
    obs, _ = env.reset(seed=1)
    v = env.buffers()
    actor = rollout.ActorNetwork().cuda()
    for t in range(12):
        a = rollout.actor_sample(actor, obs, v["n_agents"], seed=1, step=t)
        obs, r, d, info = env.step(a, auto_reset=True)
    st = env.get_state()
    env.set_state(st)
    out = env.alloc_host_out()
    rag = env.alloc_host_out(ragged=True)
    acts = np.random.RandomState(0).randint(0, 5, size=(E, 12)).astype(np.int8)
    env.step_host(acts, auto_reset=True, out=out)
    env.step_host_ragged(acts, auto_reset=True, out=rag)
    if diag:
        env.shield_diag()
    R = torch.rand(7, E, 12, device="cuda")
    D = (torch.rand(7, E, device="cuda") < 0.1)
    rollout.discounted_returns(R, D, torch.rand(E, 12, device="cuda"), 0.99)
    torch.cuda.synchronize()
    print(traffic, "ok", env.stats()["agent_steps"])
    env.close()
n = 4096
g = torch.Generator(device="cuda").manual_seed(0)
q = [torch.rand(n, generator=g, device="cuda", dtype=torch.float64) for _ in range(5)]
mm.shield_qp(q[0], q[1], q[2], (q[3] < 0.3).to(torch.uint8), q[3] - 1, q[4] + 1)
torch.cuda.synchronize()
print("done")
