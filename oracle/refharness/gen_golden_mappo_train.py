"""Golden vectors for the MAPPO learner's update, made by running the reference's own code here:
ONE `MAPPO.train` update (marl/mappo.py:161-206) on a 24-sample batch of one agent - actor and critic parameters before
and after (RMSprop, MAX_GRAD_NORM clip, clip_param 0.2), with older target networks so that the ratio is not 1.

    python oracle/refharness/gen_golden_mappo_train.py     ->  tests/golden/mappo_train_step.npz

TEST INFRASTRUCTURE ONLY: /root/reference does not exist on the GPU box; the fixture travels instead."""
import os
import sys
from copy import deepcopy
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

ref_loader.load()            # installs the import stubs (gym, pygame, matplotlib, cvxopt) and imports highway_env
sys.path.insert(0, os.path.join(ref_loader.REFERENCE_ROOT, "marl"))
sys.path.insert(0, ref_loader.REFERENCE_ROOT)
import torch  # noqa: E402
from torch.optim import RMSprop  # noqa: E402
from marl.single_agent.Model_common import ActorNetwork, CriticNetwork  # noqa: E402
from marl.mappo import MAPPO  # noqa: E402


def main():
    rng = np.random.RandomState(29)
    torch.manual_seed(9)
    actor = ActorNetwork(30, 128, 5, torch.nn.functional.log_softmax)
    critic = CriticNetwork(30, 5, 128, 1)
    actor_t, critic_t = deepcopy(actor), deepcopy(critic)
    with torch.no_grad():
        for p in list(actor_t.parameters()) + list(critic_t.parameters()):
            p.add_(0.05 * torch.randn_like(p))
    N = 24
    states = rng.uniform(-1.2, 1.2, size=(N, 1, 30)).astype(np.float32)
    actions = np.eye(5, dtype=np.float32)[rng.randint(0, 5, size=N)].reshape(N, 1, 5)
    returns = rng.uniform(-2, 2, size=(N, 1)).astype(np.float32)
    data = {}
    for name, net in (("actor", actor), ("critic", critic), ("actor_t", actor_t), ("critic_t", critic_t)):
        data.update({"%s_%s" % (name, k.replace(".", "_")): v.numpy().copy() for k, v in net.state_dict().items()})
    m = object.__new__(MAPPO)
    m.n_episodes, m.episodes_before_train, m.batch_size, m.use_cuda = 3, 1, N, False
    m.n_agents, m.state_dim, m.action_dim = 1, 30, 5
    m.clip_param, m.critic_loss, m.max_grad_norm, m.target_update_steps, m.target_tau = 0.2, "mse", 5.0, 1000, 1.0
    m.actor, m.critic, m.actor_target, m.critic_target = actor, critic, actor_t, critic_t
    m.actor_optimizer = RMSprop(actor.parameters(), lr=5e-4)
    m.critic_optimizer = RMSprop(critic.parameters(), lr=5e-4)
    m.memory = SimpleNamespace(sample=lambda n: SimpleNamespace(states=states.tolist(), actions=actions.tolist(),
                                                                  rewards=returns.tolist()))
    m.train()
    for name, net in (("actor_after", actor), ("critic_after", critic)):
        data.update({"%s_%s" % (name, k.replace(".", "_")): v.numpy().copy() for k, v in net.state_dict().items()})
    data.update(train_states=states[:, 0], train_actions=actions[:, 0], train_returns=returns, lr=5e-4, clip_param=0.2,
                max_grad_norm=5.0)
    out = os.path.join(HERE, "..", "..", "tests", "golden", "mappo_train_step.npz")
    np.savez_compressed(out, **data)
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    main()
