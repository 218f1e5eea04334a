"""Batched MAPPO rollout on top of `MergeEnvBatched` (BASELINE configs[3]).

The reference's `MAPPO.interact` (marl/mappo.py:102-158) steps ONE env, one agent at a time through a shared
actor; here the same data flow runs for E envs x up to 11 agents per step with no host synchronisation:

    obs [E,12,30] --actor--> log-probs --multinomial--> actions [E,12] int8 --env.step(auto_reset)--> ...

Kept from the reference: the networks (marl/single_agent/Model_common.py:5-40: actor 30-128-128-5 log-softmax,
critic on (state, one-hot action)), reward selection `regionalR | global_R` (mappo.py:122-125), reward scaling,
the discounted return (`_discount_reward`, mappo.py:364-370, final value 0 at episode end, critic bootstrap
otherwise) and the PPO-clip / critic losses of `MAPPO.train` (mappo.py:161-206) — batched over every live
(env, agent, t) sample instead of looped per agent.  The learner itself stays out of scope for the hot-path work;
this module is the caller that lets BASELINE configs[3] (policy + env on device) be measured.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import MAXV, NA, NS


class ActorNetwork(nn.Module):
    """marl/single_agent/Model_common.py:5-23 with output_act = log_softmax."""

    def __init__(self, state_dim=NS, hidden_size=128, output_size=NA):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_size)
        self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.fc3 = nn.Linear(hidden_size, output_size)

    def forward(self, state):
        out = torch.relu(self.fc1(state))
        out = torch.relu(self.fc2(out))
        return torch.log_softmax(self.fc3(out), dim=-1)


class CriticNetwork(nn.Module):
    """marl/single_agent/Model_common.py:26-41."""

    def __init__(self, state_dim=NS, action_dim=NA, hidden_size=128):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_size)
        self.fc2 = nn.Linear(hidden_size + action_dim, hidden_size)
        self.fc3 = nn.Linear(hidden_size, 1)

    def forward(self, state, action_one_hot):
        out = torch.relu(self.fc1(state))
        out = torch.relu(self.fc2(torch.cat([out, action_one_hot], -1)))
        return self.fc3(out)


class ActorCriticNetwork(nn.Module):
    """marl/single_agent/Model_gi.py:137-220: the shared-trunk network of `MAPPO_GI` (mappo_gi.py:142-148 builds it
    with `state_split=True` and hidden size = critic_hidden_size).  The split reads the observation as five vehicles
    of FIVE features - presence | x, y | vx, vy - through fixed column lists that stop at column 24, also when the env
    delivers 5 x 6 = 30 columns (merge-multi-agent-v1): kept as the reference has it.  out_type "p": log-softmax over
    the actions, with `action_mask` the masked logits are -1e8 (Model_gi.py:207-213); "v": the state value."""
    COLS1 = (0, 5, 10, 15, 20)
    COLS2 = (1, 2, 6, 7, 11, 12, 16, 17, 21, 22)
    COLS3 = (3, 4, 8, 9, 13, 14, 18, 19, 23, 24)

    def __init__(self, state_dim=NS, action_dim=NA, hidden_size=128, critic_output_size=1, state_split=True):
        super().__init__()
        self.state_split = bool(state_split)
        if self.state_split:
            self.fc11 = nn.Linear(5, hidden_size // 4)
            self.fc12 = nn.Linear(10, hidden_size // 2)
            self.fc13 = nn.Linear(10, hidden_size // 2)
            self.fc2 = nn.Linear(hidden_size // 4 + hidden_size // 2 + hidden_size // 2, hidden_size)
        else:
            self.fc1 = nn.Linear(state_dim, hidden_size)
            self.fc2 = nn.Linear(hidden_size, hidden_size)
        self.actor_linear = nn.Linear(hidden_size, action_dim)
        self.critic_linear = nn.Linear(hidden_size, critic_output_size)

    def trunk(self, state):
        if self.state_split:
            out = torch.cat([torch.relu(self.fc11(state[:, self.COLS1])), torch.relu(self.fc12(state[:, self.COLS2])),
                             torch.relu(self.fc13(state[:, self.COLS3]))], 1)
        else:
            out = torch.relu(self.fc1(state))
        return torch.relu(self.fc2(out))

    def policy_head(self, hidden, action_mask=None):
        logits = self.actor_linear(hidden)
        if action_mask is not None:
            logits = torch.where(action_mask == 0, torch.full_like(logits, -1e8), logits)
            return torch.log_softmax(logits + 1e-8, dim=1)
        return torch.log_softmax(logits, dim=1)

    def forward(self, state, action_mask=None, out_type="p"):
        hidden = self.trunk(state)
        return self.policy_head(hidden, action_mask) if out_type == "p" else self.critic_linear(hidden)


def mappo_losses(actor, actor_target, critic, critic_target, states, actions_one_hot, returns, clip_param=0.2,
                 critic_loss="mse", pairwise=False):
    """The two losses of one `MAPPO.train` update (mappo.py:170-199): the PPO-clip surrogate with the advantage
    `return - Q_target(s, a)` for the actor, the regression of Q(s, a) on the return for the critic.  They share no
    parameters, so stepping the actor first (as the reference does) or both at once is the same update.
    `pairwise`: see `shared_network_loss`.  -> (actor_loss, critic_loss)"""
    with torch.no_grad():
        adv = returns - critic_target(states, actions_one_hot)
        old_logp = (actor_target(states) * actions_one_hot).sum(1)
    logp = (actor(states) * actions_one_hot).sum(1)
    ratio = torch.exp(logp - old_logp)
    if not pairwise:
        adv = adv.squeeze(1)
    actor_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv).mean()
    values = critic(states, actions_one_hot)
    if critic_loss == "huber":
        c_loss = nn.functional.smooth_l1_loss(values, returns)
    else:
        c_loss = nn.functional.mse_loss(values, returns)
    return actor_loss, c_loss


def shared_network_loss(policy, policy_target, states, actions_one_hot, returns, clip_param=0.2, critic_loss="mse",
                        pairwise=False):
    """The loss of one `MAPPO_GI.train` update with the shared network (mappo_gi.py:307-343): PPO-clip surrogate with
    the advantage `return - V(s)` (value detached) plus the value regression, summed.  -> (loss, actor_loss, critic_loss)

    `pairwise`: the reference multiplies `ratio` [N] by `advantages` [N, 1] (mappo_gi.py:313-326, mappo.py:173-183), which
    broadcasts to an [N, N] table of every ratio against every advantage before min / mean.  With its batches of 100
    samples that is what it trains on; with the 1e5..1e6-sample minibatches of the batched learner the table cannot
    exist, so the default is the per-sample PPO surrogate (ratio_i * adv_i) the formula stands for.  pairwise=True
    reproduces the reference's table (small batches; used to pin this function against the reference's arithmetic)."""
    hidden = policy.trunk(states)
    logp = (policy.policy_head(hidden) * actions_one_hot).sum(1)
    values = policy.critic_linear(hidden)
    adv = returns - values.detach()
    with torch.no_grad():
        old_logp = (policy_target(states) * actions_one_hot).sum(1)
    ratio = torch.exp(logp - old_logp)
    if not pairwise:
        adv = adv.squeeze(1)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv
    actor_loss = -torch.min(surr1, surr2).mean()
    if critic_loss == "huber":
        c_loss = nn.functional.smooth_l1_loss(values, returns)
    else:
        c_loss = nn.functional.mse_loss(values, returns)
    return actor_loss + c_loss, actor_loss, c_loss


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def discounted_returns(rewards, dones, final_value, gamma):
    """R_t = r_t + gamma * R_{t+1}, restarted after a terminal step (mappo.py:364-370 run per episode), one CUDA
    kernel for the whole rollout (mm_discounted_returns).

    rewards [T, E, A] f32 cuda, dones [T, E] (non-zero where step t ended the episode of env e), final_value [E, A]
    = bootstrap for the step after the last one (ignored where the last step was terminal) or None."""
    T, E, A = rewards.shape
    rewards = rewards.contiguous().float()
    d8 = (dones != 0).to(torch.uint8).contiguous()
    assert rewards.is_cuda and d8.shape == (T, E)
    fv = None if final_value is None else final_value.contiguous().float()
    out = torch.empty_like(rewards)
    _lib.check(_lib.lib().mm_discounted_returns(_ptr(rewards), _ptr(d8), _ptr(fv), C.c_float(gamma), T, C.c_int64(E * A),
                                                A, _ptr(out), _stream()))
    return out


def set_actor_impl(name):
    """'tcgen05' (default: fp16 operands, a warpgroup per 128-row tile, weights once per SM), 'mma' (the warp-level
    mma.sync TF32 kernel, kept as an independent cross-check), 'tcgen05_tf32' (the first tcgen05 kernel: TF32 operands,
    one CTA per SM) or 'tcgen05_cta' (fp16 operands, two CTAs per SM: the default of the first half of round 2)."""
    _lib.check(_lib.lib().mm_set_actor_impl({"tcgen05": 0, "mma": 1, "tcgen05_tf32": 2, "tcgen05_cta": 3}[name]))


_DENSE_CACHE = {}


def _split_first_layer(policy):
    """The three first-layer blocks of the state_split network (Model_gi.py:137-176: fc11 over the presence columns,
    fc12 over x / y, fc13 over vx / vy) as ONE 30 -> 160 layer: weight [160][30], zero outside the blocks, bias [160].
    Cached per module until a parameter changes (torch bumps `_version` on every in-place update)."""
    ps = (policy.fc11.weight, policy.fc11.bias, policy.fc12.weight, policy.fc12.bias, policy.fc13.weight, policy.fc13.bias)
    key = tuple((p.data_ptr(), p._version) for p in ps)
    hit = _DENSE_CACHE.get(id(policy))
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    h4, h2 = policy.fc11.out_features, policy.fc12.out_features
    w = torch.zeros((h4 + 2 * h2, NS), dtype=torch.float32, device=ps[0].device)
    w[:h4, list(policy.COLS1)] = ps[0].detach().float()
    w[h4:h4 + h2, list(policy.COLS2)] = ps[2].detach().float()
    w[h4 + h2:, list(policy.COLS3)] = ps[4].detach().float()
    b = torch.cat([ps[1].detach().float(), ps[3].detach().float(), ps[5].detach().float()]).contiguous()
    _DENSE_CACHE[id(policy)] = (key, w, b)
    return w, b


def policy_sample(policy, obs, n_agents=None, seed=0, step=0, want_logp=False, action_mask=None, want_value=False,
                  out_actions=None, obs_copy=None, live_out=None):
    """Fused forward + exploration draw of the MAPPO_GI shared network (mm_actor_sample_mlp with h1 = 160): `policy` is
    an ActorCriticNetwork with state_split and hidden size 128; -> actions int8 [...] (+ log-probabilities [..., 5] with
    want_logp, + V(s) [...] from critic_linear with want_value).  Other arguments as for `actor_sample`."""
    assert policy.state_split and policy.fc2.out_features == 128 and policy.fc2.in_features == 160, \
        "the fused kernel covers the reference's shared network (hidden size 128, state_split)"
    rows = obs.numel() // NS
    obs = obs.contiguous()
    assert obs.is_cuda and obs.dtype == torch.float32
    w1, b1 = _split_first_layer(policy)
    w = [policy.fc2.weight, policy.fc2.bias, policy.actor_linear.weight, policy.actor_linear.bias,
         policy.critic_linear.weight, policy.critic_linear.bias]
    w = [t.detach().contiguous().float() for t in w]
    actions = torch.empty(obs.shape[:-1], dtype=torch.int8, device=obs.device) if out_actions is None else out_actions
    assert actions.dtype == torch.int8 and actions.is_contiguous() and actions.numel() == rows
    _check_record_slots(rows, obs_copy, live_out)
    logp = torch.empty(obs.shape[:-1] + (NA,), dtype=torch.float32, device=obs.device) if want_logp else None
    values = torch.empty(obs.shape[:-1], dtype=torch.float32, device=obs.device) if want_value else None
    if n_agents is not None:
        assert n_agents.dtype == torch.int32 and n_agents.is_cuda and n_agents.numel() * MAXV == rows
    if action_mask is not None:
        assert action_mask.dtype == torch.uint8 and action_mask.is_cuda and action_mask.numel() == rows
        action_mask = action_mask.contiguous()
    _lib.check(_lib.lib().mm_actor_sample_mlp(_ptr(obs), _ptr(n_agents), C.c_int64(rows), 160, _ptr(w1), _ptr(b1), _ptr(w[0]),
                                              _ptr(w[1]), _ptr(w[2]), _ptr(w[3]), _ptr(w[4]), _ptr(w[5]), C.c_uint64(seed),
                                              C.c_uint64(step), _ptr(action_mask), _ptr(actions), _ptr(logp), _ptr(None),
                                              _ptr(values), _ptr(obs_copy), _ptr(live_out), _stream()))
    out = (actions,) + ((logp,) if want_logp else ()) + ((values,) if want_value else ())
    return out if len(out) > 1 else actions


def actor_sample(actor, obs, n_agents=None, seed=0, step=0, want_logp=False, action_mask=None, out_actions=None,
                 obs_copy=None, live_out=None):
    """Fused actor forward + exploration draw (mm_actor_sample): obs [..., 30] f32 cuda -> actions int8 [...].

    `actor` is an ActorNetwork (or any module with fc1/fc2/fc3 Linear layers 30-128-128-5); its parameters are read
    in place.  n_agents [E] int32 marks the live rows when obs is the env's [E, 12, 30] buffer.  With want_logp the
    log-probabilities of all five actions are returned too ([..., 5] f32).  action_mask: uint8 bit masks, one per row
    (`env.buffers()["action_mask"]`, bit k = action k available) for the MAPPO_GI actor's invalid-action masking.
    Rollout-buffer appends fused into the same launch (MAPPO.interact, mappo.py:117-131): out_actions (int8, slot t of
    the action buffer) receives the actions, obs_copy (f32, slot t of the state buffer) the rows they were drawn from,
    live_out (uint8) 1 where the row's agent exists."""
    rows = obs.numel() // NS
    obs = obs.contiguous()
    assert obs.is_cuda and obs.dtype == torch.float32
    w = [actor.fc1.weight, actor.fc1.bias, actor.fc2.weight, actor.fc2.bias, actor.fc3.weight, actor.fc3.bias]
    assert tuple(w[0].shape) == (128, NS) and tuple(w[2].shape) == (128, 128) and tuple(w[4].shape) == (NA, 128)
    w = [t.detach().contiguous().float() for t in w]
    actions = torch.empty(obs.shape[:-1], dtype=torch.int8, device=obs.device) if out_actions is None else out_actions
    assert actions.dtype == torch.int8 and actions.is_contiguous() and actions.numel() == rows
    logp = torch.empty(obs.shape[:-1] + (NA,), dtype=torch.float32, device=obs.device) if want_logp else None
    if n_agents is not None:
        assert n_agents.dtype == torch.int32 and n_agents.is_cuda and n_agents.numel() * MAXV == rows
    if action_mask is not None:
        assert action_mask.dtype == torch.uint8 and action_mask.is_cuda and action_mask.numel() == rows
        action_mask = action_mask.contiguous()
    if obs_copy is None and live_out is None:
        _lib.check(_lib.lib().mm_actor_sample(_ptr(obs), _ptr(n_agents), C.c_int64(rows), *[_ptr(t) for t in w],
                                              C.c_uint64(seed), C.c_uint64(step), _ptr(action_mask), _ptr(actions), _ptr(logp),
                                              _ptr(None), _stream()))
    else:
        _check_record_slots(rows, obs_copy, live_out)
        _lib.check(_lib.lib().mm_actor_sample_mlp(_ptr(obs), _ptr(n_agents), C.c_int64(rows), 128, *[_ptr(t) for t in w],
                                                  _ptr(None), _ptr(None), C.c_uint64(seed), C.c_uint64(step), _ptr(action_mask),
                                                  _ptr(actions), _ptr(logp), _ptr(None), _ptr(None), _ptr(obs_copy), _ptr(live_out),
                                                  _stream()))
    return (actions, logp) if want_logp else actions


def _check_record_slots(rows, obs_copy, live_out):
    if obs_copy is not None:
        assert obs_copy.is_cuda and obs_copy.dtype == torch.float32 and obs_copy.is_contiguous() and obs_copy.numel() == rows * NS
    if live_out is not None:
        assert live_out.is_cuda and live_out.dtype == torch.uint8 and live_out.is_contiguous() and live_out.numel() == rows


class BatchedMAPPORollout(object):
    def __init__(self, env, actor=None, critic=None, roll_out_n_steps=100, reward_gamma=0.99, reward_scale=20.0,
                 reward_type="regionalR", clip_param=0.2, actor_lr=5e-4, critic_lr=5e-4, max_grad_norm=5.0, seed=0,
                 fused=True, obs=None, reset_seed=None):
        """obs: the observation of an env the caller has already reset (train.py resets with a rank-distinct spawn
        seed and hands its first observation over).  Without it the first collect() spawns the scenes itself with
        `reset_seed` (default: derived from `seed`, which the data-parallel driver makes rank-distinct) - never with the
        env's config seed, which is the same on every rank and would make the shards replicas of one another."""
        self.env = env
        dev = torch.device("cuda", env.device)
        self.actor = (actor or ActorNetwork()).to(dev)
        self.critic = (critic or CriticNetwork()).to(dev)
        self.actor_target = ActorNetwork().to(dev)
        self.critic_target = CriticNetwork().to(dev)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.T = int(roll_out_n_steps)
        self.gamma, self.reward_scale, self.reward_type = float(reward_gamma), float(reward_scale), reward_type
        self.clip_param, self.max_grad_norm = clip_param, max_grad_norm
        self.actor_opt = torch.optim.RMSprop(self.actor.parameters(), lr=actor_lr)   # mappo.py:88-90
        self.critic_opt = torch.optim.RMSprop(self.critic.parameters(), lr=critic_lr)
        self.dev = dev
        self.obs = obs
        self.reset_seed = (0x5EED0000 + int(seed)) if reset_seed is None else int(reset_seed)
        E = env.n_envs
        self._slot = torch.arange(MAXV, device=dev)[None, :]
        self.buf = None
        self.E = E
        self.seed, self._draws = int(seed), 0
        self.fused = bool(fused)

    @torch.no_grad()
    def act_fused(self, obs, n_agents):
        """The draw of `act_torch` through the fused CUDA kernel (TF32 actor + inverse-CDF sampling, one launch)."""
        self._draws += 1
        a = actor_sample(self.actor, obs, n_agents, seed=self.seed, step=self._draws)
        return a, self._slot < n_agents[:, None]

    @torch.no_grad()
    def _record_act(self, obs, n_agents, a_slot, s_slot, l_slot):
        self._draws += 1
        return actor_sample(self.actor, obs, n_agents, seed=self.seed, step=self._draws, out_actions=a_slot, obs_copy=s_slot,
                            live_out=l_slot)

    @torch.no_grad()
    def act_torch(self, obs, n_agents):
        """Plain torch fp32 actor + torch.multinomial: the numerics reference of `act_fused`."""
        logp = self.actor(obs.reshape(-1, NS))                               # [E*12, 5]
        a = torch.multinomial(logp.exp(), 1).view(obs.shape[0], MAXV)           # exploration_action, mappo.py:225-230
        live = self._slot < n_agents[:, None]
        return torch.where(live, a, torch.ones_like(a)).to(torch.int8), live

    def _act(self, obs, n_agents):
        return self.act_fused(obs, n_agents) if self.fused else self.act_torch(obs, n_agents)

    @torch.no_grad()
    def collect(self):
        """One rollout of T policy steps for every env; returns the batch dict (device tensors)."""
        env, E, T = self.env, self.E, self.T
        v = env.buffers()
        if self.obs is None:
            self.obs, _ = env.reset(seed=self.reset_seed)
        S = torch.empty((T, E, MAXV, NS), device=self.dev)
        A8 = torch.empty((T, E, MAXV), dtype=torch.int8, device=self.dev)
        L8 = torch.empty((T, E, MAXV), dtype=torch.uint8, device=self.dev)
        R = torch.empty((T, E, MAXV), device=self.dev)
        D = torch.empty((T, E), device=self.dev)
        for t in range(T):
            if self.fused:
                # the draw kernel appends to the rollout buffer itself: state S[t], action A8[t], live mask L8[t]
                a = self._record_act(self.obs, v["n_agents"], A8[t], S[t], L8[t])
            else:
                n_agents = v["n_agents"].clone()
                S[t].copy_(self.obs)
                a, live = self._act(self.obs, n_agents)
                A8[t].copy_(a)
                L8[t].copy_(live)
            self.obs, reward, done, info = env.step(a, auto_reset=True)
            if self.reward_type == "regionalR":
                R[t].copy_(info["regional_rewards"])
            else:
                R[t].copy_(reward[:, None].expand(E, MAXV))
            D[t].copy_(done)
        A, L = A8.long(), L8.bool()
        if self.reward_scale > 0:
            R /= self.reward_scale
        # bootstrap where the last step did not end the episode (mappo.py:148-150)
        final_value = self._final_value(self.obs, v["n_agents"])
        returns = discounted_returns(R, D, final_value, self.gamma)
        self.buf = dict(states=S, actions=A, returns=returns, live=L, dones=D, rewards=R)
        return self.buf

    def _final_value(self, obs, n_agents):
        a_fin, _ = self._act(obs, n_agents)
        onehot = torch.nn.functional.one_hot(a_fin.long(), NA).float()
        return self.critic(obs.reshape(-1, NS), onehot.view(-1, NA)).view(self.E, MAXV)

    def networks(self):
        """every module a replica must hold identically (sync_parameters, the replica check of train.py)"""
        return [self.actor, self.critic, self.actor_target, self.critic_target]

    def sample_actions(self, obs, n_agents, seed, step):
        """MAPPO.action (mappo.py:231-236): a draw from the softmax, also at evaluation time."""
        return actor_sample(self.actor, obs.contiguous(), n_agents, seed=seed, step=step)

    # ---- data parallelism over env shards: one process per GPU, each with its own envs; the only
    # traffic is the gradient all-reduce of the two small networks (~42 k parameters each) per minibatch
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def sync_parameters(self, src=0):
        """Broadcast rank `src`'s networks (call once after construction under torchrun)."""
        import torch.distributed as dist
        if self._world() > 1:
            for m in self.networks():
                for t in list(m.parameters()) + list(m.buffers()):
                    dist.broadcast(t.data, src)

    def _allreduce_grads(self, params):
        import torch.distributed as dist
        w = self._world()
        if w == 1:
            return
        params = [q for q in params if q.grad is not None]
        flat = torch.cat([q.grad.reshape(-1) for q in params])
        dist.all_reduce(flat)
        flat /= w
        o = 0
        for q in params:
            q.grad.copy_(flat[o:o + q.numel()].view_as(q.grad))
            o += q.numel()

    def update(self, minibatch=1 << 18, epochs=1):
        """PPO-clip actor update and critic regression on the last rollout (mappo.py:161-206, `mappo_losses`), all
        agents at once, with the per-sample surrogate (the reference's [N] x [N, 1] product is an N x N table).
        Under torch.distributed every rank runs the same number of minibatches and the gradients are averaged."""
        b = self.buf
        live = b["live"].reshape(-1)
        idx = live.nonzero(as_tuple=False).squeeze(1)
        S = b["states"].reshape(-1, NS)
        A = torch.nn.functional.one_hot(b["actions"].reshape(-1), NA).float()
        G = b["returns"].reshape(-1, 1)
        stats = {}
        n_max = torch.tensor([idx.numel()], device=self.dev)
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
        n_mb = max(1, -(-int(n_max) // int(minibatch)))       # same count on every rank
        for _ in range(epochs):
            perm = idx[torch.randperm(idx.numel(), device=self.dev)]
            for j in torch.tensor_split(perm, n_mb):
                actor_loss, critic_loss = mappo_losses(self.actor, self.actor_target, self.critic, self.critic_target,
                                                       S[j], A[j], G[j], self.clip_param)
                self.actor_opt.zero_grad(set_to_none=True)
                actor_loss.backward()
                self._allreduce_grads(list(self.actor.parameters()))
                nn.utils.clip_grad_norm_(self.actor.parameters(), self.max_grad_norm)
                self.actor_opt.step()
                self.critic_opt.zero_grad(set_to_none=True)
                critic_loss.backward()
                self._allreduce_grads(list(self.critic.parameters()))
                nn.utils.clip_grad_norm_(self.critic.parameters(), self.max_grad_norm)
                self.critic_opt.step()
                stats = {"actor_loss": float(actor_loss.detach()), "critic_loss": float(critic_loss.detach()),
                         "samples": int(idx.numel())}
        self.actor_target.load_state_dict(self.actor.state_dict())     # TARGET_TAU = 1.0 in every shipped ini
        self.critic_target.load_state_dict(self.critic.state_dict())
        return stats


class BatchedMAPPOGIRollout(BatchedMAPPORollout):
    """`MAPPO_GI` with `shared_network = True` (marl/mappo_gi.py; the 6 `*-shared*.ini` configs, run_mappo.py:232-277):
    ONE ActorCriticNetwork (Model_gi.py:137-220, state_split) gives the action distribution and the state value, one
    RMSprop optimiser at `actor_lr` (mappo_gi.py:150-156) steps the summed loss (`shared_network_loss`), the bootstrap
    value is V(final state) (mappo_gi.py:396-404).  Rollout, reward selection, scaling and returns are the base class's.
    As in the reference the policy is evaluated WITHOUT an action mask everywhere (`self.policy(state)`,
    mappo_gi.py:309,359): `action_masking` only changes what env.reset / step report.  With the reference's hidden size
    (128) the action draw and the bootstrap value come from the fused kernel (`policy_sample`: mm_actor_sample_mlp with
    the three first-layer blocks as one 30 -> 160 layer); other sizes run as torch layers."""

    def __init__(self, env, policy=None, hidden_size=128, **kw):
        net = policy or ActorCriticNetwork(NS, NA, hidden_size, 1, state_split=True)
        kw["fused"] = bool(kw.get("fused", True)) and net.state_split and net.fc2.in_features == 160 and \
            net.fc2.out_features == 128
        dev = torch.device("cuda", env.device)
        super().__init__(env, **kw)
        self.policy = net.to(dev)
        self.policy_target = ActorCriticNetwork(NS, NA, hidden_size, 1, state_split=self.policy.state_split).to(dev)
        self.policy_target.load_state_dict(self.policy.state_dict())
        self.policy_opt = torch.optim.RMSprop(self.policy.parameters(), lr=self.actor_opt.param_groups[0]["lr"])
        self.actor = self.critic = self.actor_target = self.critic_target = None     # not part of this learner
        self.actor_opt = self.critic_opt = None

    @torch.no_grad()
    def act_torch(self, obs, n_agents):
        logp = self.policy(obs.reshape(-1, NS))
        a = torch.multinomial(logp.exp(), 1).view(obs.shape[0], MAXV)           # exploration_action, mappo_gi.py:368-373
        live = self._slot < n_agents[:, None]
        return torch.where(live, a, torch.ones_like(a)).to(torch.int8), live

    @torch.no_grad()
    def act_fused(self, obs, n_agents):
        self._draws += 1
        a = policy_sample(self.policy, obs, n_agents, seed=self.seed, step=self._draws)
        return a, self._slot < n_agents[:, None]

    @torch.no_grad()
    def _record_act(self, obs, n_agents, a_slot, s_slot, l_slot):
        self._draws += 1
        return policy_sample(self.policy, obs, n_agents, seed=self.seed, step=self._draws, out_actions=a_slot, obs_copy=s_slot,
                             live_out=l_slot)

    def _final_value(self, obs, n_agents):
        if self.fused:     # V(s) from the same kernel family: critic_linear as a sixth output column
            return policy_sample(self.policy, obs, n_agents, seed=self.seed, step=0, want_value=True)[1].view(self.E, MAXV)
        return self.policy(obs.reshape(-1, NS), out_type="v").view(self.E, MAXV)

    def networks(self):
        return [self.policy, self.policy_target]

    def sample_actions(self, obs, n_agents, seed, step):
        if self.fused:
            return policy_sample(self.policy, obs.contiguous(), n_agents, seed=seed, step=step)
        return self.act_torch(obs, n_agents)[0]

    def update(self, minibatch=1 << 18, epochs=1):
        """One pass of `MAPPO_GI.train` (shared branch, mappo_gi.py:307-349) over the last rollout, all agents at once;
        under torch.distributed the gradients are averaged over the ranks."""
        b = self.buf
        idx = b["live"].reshape(-1).nonzero(as_tuple=False).squeeze(1)
        S = b["states"].reshape(-1, NS)
        A = torch.nn.functional.one_hot(b["actions"].reshape(-1), NA).float()
        G = b["returns"].reshape(-1, 1)
        stats = {}
        n_max = torch.tensor([idx.numel()], device=self.dev)
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
        n_mb = max(1, -(-int(n_max) // int(minibatch)))       # same count on every rank
        params = list(self.policy.parameters())
        for _ in range(epochs):
            perm = idx[torch.randperm(idx.numel(), device=self.dev)]
            for j in torch.tensor_split(perm, n_mb):
                loss, a_loss, c_loss = shared_network_loss(self.policy, self.policy_target, S[j], A[j], G[j], self.clip_param)
                self.policy_opt.zero_grad(set_to_none=True)
                loss.backward()
                self._allreduce_grads(params)
                nn.utils.clip_grad_norm_(params, self.max_grad_norm)
                self.policy_opt.step()
                stats = {"actor_loss": float(a_loss.detach()), "critic_loss": float(c_loss.detach()),
                         "samples": int(idx.numel())}
        self.policy_target.load_state_dict(self.policy.state_dict())   # TARGET_TAU = 1.0 in every shipped ini
        return stats
