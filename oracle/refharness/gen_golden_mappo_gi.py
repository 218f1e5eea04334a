"""Golden vectors for the shared-network caller, made by running the reference's own code here:

  * ActorCriticNetwork (marl/single_agent/Model_gi.py:137-220, state_split=True as MAPPO_GI builds it) on random
    observation rows: log-probabilities without and with an action mask, and state values;
  * ONE `MAPPO_GI.train` update of the shared branch (marl/mappo_gi.py:232-349) on a 24-sample batch of one agent:
    the network's parameters before and after (RMSprop at actor_lr, MAX_GRAD_NORM clip, clip_param 0.2).

    python oracle/refharness/gen_golden_mappo_gi.py     ->  tests/golden/mappo_gi_caller.npz

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
from marl.single_agent.Model_gi import ActorCriticNetwork  # noqa: E402
from marl.mappo_gi import MAPPO_GI  # noqa: E402


def main():
    rng = np.random.RandomState(23)
    torch.manual_seed(5)
    net = ActorCriticNetwork(30, 5, 128, 1, state_split=True)
    obs = torch.from_numpy(rng.uniform(-1.2, 1.2, size=(256, 30)).astype(np.float32))
    obs[::9] = 0.0
    mask = (rng.uniform(size=(256, 5)) < 0.7).astype(np.int64)
    mask[:, 1] = 1                                   # IDLE is always available (abstract.py:219-240)
    with torch.no_grad():
        logp = net(obs)
        logp_masked = net(obs, action_mask=torch.from_numpy(mask))
        value = net(obs, out_type="v")
    data = dict(obs=obs.numpy(), mask=mask, logp=logp.numpy(), logp_masked=logp_masked.numpy(), value=value.numpy())
    data.update({"w_" + k.replace(".", "_"): v.numpy().copy() for k, v in net.state_dict().items()})

    # one shared-branch update through the reference's own train()
    N = 24
    states = rng.uniform(-1.2, 1.2, size=(N, 1, 30)).astype(np.float32)
    actions = np.eye(5, dtype=np.float32)[rng.randint(0, 5, size=N)].reshape(N, 1, 5)
    returns = rng.uniform(-2, 2, size=(N, 1)).astype(np.float32)
    target = deepcopy(net)
    with torch.no_grad():                            # an older policy: the ratio must not be 1 everywhere
        for p in target.parameters():
            p.add_(0.05 * torch.randn_like(p))
    m = object.__new__(MAPPO_GI)
    m.n_episodes, m.episodes_before_train, m.batch_size, m.use_cuda = 3, 1, N, False
    m.n_agents, m.state_dim, m.action_dim, m.shared_network = 1, 30, 5, True
    m.clip_param, m.critic_loss, m.max_grad_norm, m.target_update_steps, m.target_tau = 0.2, "mse", 5.0, 1000, 1.0
    m.policy, m.policy_target = net, target
    m.policy_optimizer = RMSprop(net.parameters(), lr=5e-4)
    m.memory = SimpleNamespace(sample=lambda n: SimpleNamespace(states=states.tolist(), actions=actions.tolist(),
                                                                  rewards=returns.tolist()))
    data.update({"t_" + k.replace(".", "_"): v.numpy().copy() for k, v in target.state_dict().items()})
    m.train()
    data.update({"after_" + k.replace(".", "_"): v.numpy().copy() for k, v in net.state_dict().items()})
    data.update(train_states=states[:, 0], train_actions=actions[:, 0], train_returns=returns, lr=5e-4, clip_param=0.2,
                max_grad_norm=5.0)
    out = os.path.join(HERE, "..", "..", "tests", "golden", "mappo_gi_caller.npz")
    np.savez_compressed(out, **data)
    print("wrote", os.path.normpath(out), "max |dW| after the update:",
          max(float(np.abs(data["after_" + k[2:]] - data[k]).max()) for k in data if k.startswith("w_")))


if __name__ == "__main__":
    main()
