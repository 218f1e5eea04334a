"""Small driver for ncu on the actor kernels: python profiles/profile_actor.py [envs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import marl_mass_b200 as mm
from marl_mass_b200 import rollout
E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
obs = (torch.rand(E, 12, mm.NS, device="cuda") * 2 - 1).contiguous()
n_ag = torch.randint(7, 12, (E,), device="cuda", dtype=torch.int32)
actor, gi = rollout.ActorNetwork().cuda(), rollout.ActorCriticNetwork().cuda()
for t in range(4):
    rollout.actor_sample(actor, obs, n_ag, seed=1, step=t)
    rollout.policy_sample(gi, obs, n_ag, seed=1, step=t)
torch.cuda.synchronize()
print("ok")
