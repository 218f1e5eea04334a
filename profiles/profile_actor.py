"""ncu driver for the fused actor kernel: python profiles/profile_actor.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import marl_mass_b200 as mm
from marl_mass_b200 import rollout
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 786432
torch.manual_seed(0)
actor = rollout.ActorNetwork().cuda()
obs = (torch.rand(rows, mm.NS, device="cuda") * 2 - 1).contiguous()
for t in range(4):
    a = rollout.actor_sample(actor, obs, None, seed=1, step=t)
torch.cuda.synchronize()
print("ok", int(a.sum()))
