"""Golden vectors for the caller-side kernels, made by running the reference's own code here:

  * MAPPO._discount_reward (marl/mappo.py:364-370) on random reward columns with a bootstrap value;
  * ActorNetwork (marl/single_agent/Model_common.py:5-23) with log_softmax output on random observation rows.

    python oracle/refharness/gen_golden_mappo.py     ->  tests/golden/mappo_caller.npz

TEST INFRASTRUCTURE ONLY: /root/reference does not exist on the GPU box; the fixture travels instead."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402  (installs the import stubs)

ref_loader.load() if hasattr(ref_loader, "load") else None
sys.path.insert(0, os.path.join(ref_loader.REFERENCE_ROOT, "marl"))
sys.path.insert(0, ref_loader.REFERENCE_ROOT)
import torch  # noqa: E402
from single_agent.Model_common import ActorNetwork  # noqa: E402
import mappo as ref_mappo  # noqa: E402


class _Self(object):
    reward_gamma = 0.99


def main():
    rng = np.random.RandomState(7)
    # discounted returns: 6 columns of different lengths, as MAPPO.interact calls it per agent
    cols = []
    for T in (1, 2, 7, 33, 100, 100):
        r = rng.uniform(-3, 1, size=T)
        fv = float(rng.uniform(-2, 2)) if T != 33 else 0.0
        cols.append((r, fv, ref_mappo.MAPPO._discount_reward(_Self(), r.copy(), fv)))
    T_max = 100
    rewards = np.zeros((T_max, len(cols)))
    returns = np.zeros((T_max, len(cols)))
    lengths = np.array([len(c[0]) for c in cols])
    finals = np.array([c[1] for c in cols])
    for k, (r, fv, d) in enumerate(cols):
        rewards[:len(r), k] = r
        returns[:len(r), k] = d
    # actor
    torch.manual_seed(11)
    actor = ActorNetwork(30, 128, 5, torch.nn.functional.log_softmax)
    obs = torch.from_numpy(rng.uniform(-1.2, 1.2, size=(512, 30)).astype(np.float32))
    obs[::7] = 0.0           # rows of absent agents are all-zero in the env buffer
    with torch.no_grad():
        logp = actor(obs)
        if logp.dim() == 2 and abs(float(logp.exp().sum(1).mean()) - 1.0) > 1e-3:
            raise SystemExit("log_softmax axis surprise")
    out = os.path.join(HERE, "..", "..", "tests", "golden", "mappo_caller.npz")
    np.savez_compressed(out, rewards=rewards, returns=returns, lengths=lengths, finals=finals, gamma=0.99,
                        obs=obs.numpy(), logp=logp.numpy(),
                        **{"w_" + k.replace(".", "_"): v.numpy() for k, v in actor.state_dict().items()})
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    main()
