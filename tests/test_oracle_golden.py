"""CPU suite, part 1: pin the oracle (oracle/merge_oracle.c) against the golden vectors that
oracle/refharness/gen_golden.py froze from the UNMODIFIED reference env + shield.

Teacher forcing: every policy step starts from the reference's own pre-step state, gets the
reference's action tuple, and must reproduce the reference's post-step state, MergeEnv.step outputs and
per-sub-step shield record.  Discrete outputs bit-exact; continuous within 1e-9 relative (observed: 1e-13).
"""
import numpy as np
import pytest

import oracle as orc
from conftest import GOLDEN_CASES, HDV_TIE_CASES, SUPERVISED_CASES, TIE_CASES, V0_CASES, V0_TIE_CASES
from helpers import (ENV_FIELDS, F64_FIELDS, LC_BOUNDARY_EPS, OUT_F, OUT_I, SH_F, SH_I, compare_states,
                     load_golden, obs25, rel_err)

TOL = 1e-9
ALL = (orc.F64_FIELDS, orc.I32_FIELDS, orc.ENV_FIELDS)


def golden_state(g, rows):
    return orc.state_from_golden(g, rows)


@pytest.mark.parametrize("name", GOLDEN_CASES + TIE_CASES)
def test_teacher_forced_step(name):
    g, cfg = load_golden(name)
    rows = g["row_of_step"]
    st = golden_state(g, rows)
    out = orc.step(cfg, st, g["act"], n_threads=4)
    want = golden_state(g, rows + 1)
    ran = g["sh_ran"] == 1
    assert np.array_equal(out["sh_ran"], g["sh_ran"])
    boundary = ran & (out["sh_lc_margin"] < LC_BOUNDARY_EPS)
    # DESIGN.md "Veto tie rule": a veto test that is 0 in exact arithmetic and +-1e-15 in float64.  Where the reference's
    # rounding noise fell on the other side, the step's state legitimately differs (the lane change is cancelled or
    # not): those steps leave the state / output comparison, and there must be next to none of them
    flipped = (boundary & (out["sh_is_lc_safe"] != g["sh_is_lc_safe"])).any(axis=(1, 2))
    assert flipped.sum() <= 1, np.where(flipped)[0]
    keep = np.where(~flipped)[0]
    compare_states({k: st[k][keep] for k in st}, {k: want[k][keep] for k in want}, TOL, name)
    for k in OUT_I:
        assert np.array_equal(out[k][keep], g[k][keep]), k
    for k in OUT_F:
        got = obs25(out[k]) if (k == "obs" and cfg["n_s"] == 25) else out[k]
        assert rel_err(got[keep], g[k][keep]).max() <= TOL, k
    for k in SH_I:
        bad = (out["sh_" + k] != g["sh_" + k]) & ran & ~boundary
        assert not bad.any(), (k, np.argwhere(bad)[:5].tolist())
    for k in SH_F:
        assert (rel_err(out["sh_" + k], g["sh_" + k]) * (ran & ~boundary)).max() <= TOL, k
    # the excluded boundary set must stay tiny (<= 1% of solves)
    assert boundary.sum() <= 0.01 * max(ran.sum(), 1) + 1


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_free_running_episodes(name):
    """From each episode's first state, step the oracle alone with the recorded actions to the end of the
    episode: it must terminate on the same step and land on the reference's final state."""
    g, cfg = load_golden(name)
    ep = g["ep_start"]
    rows = g["row_of_step"]
    for j in range(len(ep) - 1):
        first, last = ep[j], ep[j + 1] - 1
        steps = np.where((rows >= first) & (rows < last))[0]
        st = golden_state(g, [first])
        done = 0
        for t in steps:
            assert not done
            out = orc.step(cfg, st, g["act"][t:t + 1])
            done = int(out["done"][0])
            assert done == int(g["done"][t])
            assert rel_err(out["reward"], g["reward"][t]).max() <= 1e-7
        assert done == 1
        compare_states(st, golden_state(g, [last]), 1e-7, "%s episode %d" % (name, j))


@pytest.mark.parametrize("name", [c for c in GOLDEN_CASES if "unsafe" not in c])
def test_qp_known_answers(name):
    """Every QP the reference posed (G, h as built by cbf.py:288-322/374-422) -> same minimiser and active set."""
    g, _ = load_golden(name)
    n = len(g["qp_a"])
    assert n > 1000
    for i in range(0, n, 7):
        u, act = orc.qp(g["qp_a"][i], g["qp_c_lead"][i], g["qp_c_adj"][i], g["qp_has_adj"][i] > 0,
                        g["qp_lo"][i], g["qp_hi"][i])
        assert u == g["qp_u"][i] and act == int(g["qp_active"][i])


def test_qp_cases_by_hand():
    # inactive: 0 inside the box and below the barrier limit
    assert orc.qp(0.0667, 3.0, 0.0, False, -0.8, 0.4) == (0.0, 0)
    # barrier row active: u = c/a
    u, act = orc.qp(0.05, -0.01, 0.0, False, -0.8, 0.4)
    assert u == -0.01 / 0.05 and act == 1
    # lower bound + slack: the limit is below the box
    u, act = orc.qp(0.05, -1.0, 0.0, False, -0.8, 0.4)
    assert u == -0.8 and act == (1 | 4 | 16)
    # adjacent row tighter than the lead row
    u, act = orc.qp(0.05, -0.01, -0.02, True, -0.8, 0.4)
    assert u == -0.02 / 0.05 and act == 8
    # box entirely below zero
    u, act = orc.qp(0.05, 3.0, 0.0, False, -0.8, -0.1)
    assert u == -0.1 and act == 2


def test_observe_matches_reset_observation():
    """reset() returns the observation of the spawned scene: oracle observe on a golden pre-state equals the
    obs the reference returned for the previous step's post-state."""
    g, cfg = load_golden("mass_td3_mixed")
    rows = g["row_of_step"]
    st = golden_state(g, rows + 1)
    obs = orc.observe(st)
    assert rel_err(obs, g["obs"]).max() <= TOL


def test_fixture_inventory():
    """Fixtures cover crashes, vetoes, every QP active-set class that occurs, HDVs and both shields."""
    seen_active = set()
    crashed = vetoes = hdv = 0
    for name in GOLDEN_CASES:
        g, cfg = load_golden(name)
        seen_active |= set(np.unique(g["sh_active"][g["sh_ran"] == 1]).tolist())
        crashed += int(g["st_crashed"].max())
        vetoes += int(((g["sh_ran"] == 1) & (g["sh_is_lc_safe"] == 0)).sum())
        hdv += int((g["st_kind"] == 2).sum() > 0)
        for k in F64_FIELDS:
            assert np.isfinite(g["st_" + k]).all()
        for k in ENV_FIELDS:
            assert (g["st_" + k] >= 0).all()
    assert {0, 1, 8}.issubset(seen_active), seen_active
    assert crashed >= 3 and vetoes > 1000 and hdv >= 4
    # the tie fixtures really hold ties: vehicles sharing x, and egos with two others at the same |dx|
    for name in TIE_CASES:
        g, _ = load_golden(name)
        rows = g["row_of_step"]
        same_x = same_key = 0
        for r in rows:
            xs = g["st_x"][r, :g["st_n_veh"][r]]
            same_x += len(xs) - len(np.unique(xs))
            for i in range(len(xs)):
                d = np.delete(np.abs(xs - xs[i]), i)
                same_key += len(d) - len(np.unique(d))
        assert same_x >= 5 and same_key >= 100, (name, same_x, same_key)


@pytest.mark.parametrize("name", V0_CASES + V0_TIE_CASES)
def test_v0_env_teacher_forced_and_free_running(name):
    """BASELINE configs[0] (test-configs_marl-cav-unsafe.ini, env merge-multi-agent-v0): MDPVehicle CAVs without the
    [-12.5, 6] acceleration clip, plain IDMVehicle HDVs, 5x5 Kinematics observation."""
    g, cfg = load_golden(name)
    assert cfg["n_s"] == 25
    rows = g["row_of_step"]
    st = golden_state(g, rows)
    out = orc.step(cfg, st, g["act"], n_threads=4)
    compare_states(st, golden_state(g, rows + 1), TOL, name, v0=True)
    for k in OUT_I:
        assert np.array_equal(out[k], g[k]), k
    for k in OUT_F:
        got = obs25(out[k]) if k == "obs" else out[k]
        assert rel_err(got, g[k]).max() <= TOL, k
    assert not out["sh_ran"].any()
    if name in V0_TIE_CASES:        # snapped before every step: no chain to run freely
        return
    ep = g["ep_start"]
    for j in range(len(ep) - 1):
        steps = np.where((rows >= ep[j]) & (rows < ep[j + 1] - 1))[0]
        s1 = golden_state(g, [ep[j]])
        for t in steps:
            o = orc.step(cfg, s1, g["act"][t:t + 1])
            assert int(o["done"][0]) == int(g["done"][t])
        compare_states(s1, golden_state(g, [ep[j + 1] - 1]), 1e-7, "%s episode %d" % (name, j), v0=True)


def test_caller_side_restatements_match_the_reference_vectors():
    """oracle.mappo_discount / oracle.actor_log_probs against MAPPO._discount_reward and ActorNetwork outputs frozen
    from the reference (oracle/refharness/gen_golden_mappo.py)."""
    import os
    import oracle
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "mappo_caller.npz"))
    for k, T in enumerate(g["lengths"]):
        got = oracle.mappo_discount(g["rewards"][:T, k:k + 1], np.zeros((T, 1)), g["finals"][k:k + 1], float(g["gamma"]))
        assert np.abs(got[:, 0] - g["returns"][:T, k]).max() < 1e-12
    w = {k[2:]: g[k] for k in g.files if k.startswith("w_")}
    assert np.abs(oracle.actor_log_probs(w, g["obs"]) - g["logp"]).max() < 5e-6   # reference runs the net in fp32
    # an episode boundary inside the segment restarts the sum (MAPPO.interact pushes one segment per episode)
    r = np.arange(6, dtype=np.float64).reshape(6, 1)
    d = np.array([0, 0, 1, 0, 0, 0]).reshape(6, 1)
    got = oracle.mappo_discount(r, d, np.array([10.0]), 0.5)
    a = oracle.mappo_discount(r[:3], np.zeros((3, 1)), np.array([0.0]), 0.5)
    b = oracle.mappo_discount(r[3:], np.zeros((3, 1)), np.array([10.0]), 0.5)
    assert np.allclose(got[:, 0], np.concatenate([a[:, 0], b[:, 0]]))


@pytest.mark.parametrize("name", ["hdv_td3"] + HDV_TIE_CASES)
def test_hdv_env_teacher_forced_and_free_running(name):
    """MergeEnvLCHDV (merge-multi-agent-hdv-v1, test-idm-td3.ini): IDM / MOBIL only; one observation row and one reward
    term per vehicle, reward = their mean, done on any crash or at the horizon, min headway over every vehicle."""
    g, cfg = load_golden(name)
    rows = g["row_of_step"]
    st = golden_state(g, rows)
    out = orc.step(cfg, st, g["act"], n_threads=4)
    want = golden_state(g, rows + 1)
    compare_states(st, want, TOL, "hdv teacher-forced")
    nv = want["n_veh"]
    m = np.arange(12)[None, :] < nv[:, None]
    assert np.array_equal(out["done"], g["done"]) and (out["agents_dones"] == 0).all() and (out["regional_rewards"] == 0).all()
    assert rel_err(out["obs"], g["obs"]).max() <= TOL and (out["obs"][~m] == 0).all()
    for k in ("reward", "average_speed", "traffic_speed", "min_headway", "merge_percent"):
        assert rel_err(out[k], g[k]).max() <= TOL, k
    assert rel_err(out["agents_rewards"], g["agents_rewards"]).max() <= TOL
    assert np.array_equal(out["average_speed"], out["traffic_speed"]) and (out["sh_ran"] == 0).all()
    if name in HDV_TIE_CASES:       # snapped before every step: no chain to run freely
        return
    assert rel_err(orc.observe(golden_state(g, g["ep_start"][:-1]), env_hdv=True)[:, :, 0].sum(1), nv[g["ep_start"][:-1]].astype(float)).max() == 0
    # free running: whole episodes from the reference's reset states
    ep = g["ep_start"]
    st = golden_state(g, ep[:-1])
    first = np.searchsorted(rows, ep[:-1])
    for t in range(100):
        out = orc.step(cfg, st, g["act"][first + np.minimum(t, 99)], n_threads=2)
    assert out["done"].all() and (st["steps"] == 100).all()
    compare_states(st, golden_state(g, ep[1:] - 1), 1e-6, "hdv free-running")


@pytest.mark.parametrize("name", SUPERVISED_CASES)
def test_supervised_actions_drive_the_unshielded_v0_dynamics(name):
    """safety_guarantee = priority | dmc (central_layer.py, decentralised_dmc.py) only replaces the meta-action tuple
    before _simulate (abstract.py:459-467): stepping the un-shielded v0 env with the supervised tuple the reference
    executed (`new_act`) reproduces its post-state and outputs.  The fixtures also carry the policy's tuple and the
    supervisor's random draws for the supervisor kernel that is not built yet."""
    g, cfg = load_golden(name)
    assert cfg["safety_guarantee"] in ("priority", "dmc") and cfg["n_s"] == 25
    changed = (g["act"] != g["new_act"]).any(axis=1)
    assert changed.sum() >= 20 and (~np.isnan(g["rand_draws"])).sum(1).min() >= g["st_n_cav"][g["row_of_step"]].min()
    rows = g["row_of_step"]
    st = golden_state(g, rows)
    out = orc.step(dict(cfg, safety_guarantee="none"), st, g["new_act"], n_threads=4)
    compare_states(st, golden_state(g, rows + 1), TOL, name, v0=True)
    for k in OUT_I:
        assert np.array_equal(out[k], g[k]), k
    for k in OUT_F:
        got = obs25(out[k]) if k == "obs" else out[k]
        assert rel_err(got, g[k]).max() <= TOL, k
    # and the policy's own tuple would NOT have: the supervisor's replacements matter
    st2 = golden_state(g, rows[changed])
    orc.step(dict(cfg, safety_guarantee="none"), st2, g["act"][changed], n_threads=4)
    want = golden_state(g, rows[changed] + 1)
    assert np.abs(st2["speed"] - want["speed"]).max() > 1e-3 or np.abs(st2["y"] - want["y"]).max() > 1e-3


def test_qp_closed_form_minimises_the_reference_objective():
    """The closed form against a direct minimisation of the QP the reference hands to cvxopt
    (cbf.py:110-135): min 1/2 (u0^2 + u1^2 + 1e18 s^2) s.t. a u0 - s <= c_i, lo <= u0 <= hi.  u1 = 0 and
    s = max(0, max_i(a u0 - c_i)) at the optimum, which leaves a convex 1-D problem solved here by ternary search."""
    from fractions import Fraction
    rng = np.random.RandomState(5)
    n_checked = n_active = 0
    for _ in range(1000):
        a = float(rng.choice([1.0, -1.0]) * rng.uniform(0.01, 0.08)) if rng.rand() > 0.05 else 0.0
        c_lead = float(rng.normal(0.05, 0.15))
        has_adj = bool(rng.rand() < 0.5)
        c_adj = float(rng.normal(0.05, 0.15)) if has_adj else 0.0
        lo = float(-12.5 / 15 + rng.uniform(-0.2, 0.2))
        hi = float(6.0 / 15 + rng.uniform(-0.2, 0.2))
        if rng.rand() < 0.1:
            lo, hi = lo + 1.0, hi + 1.0           # box entirely above zero
        cs = [c_lead] + ([c_adj] if has_adj else [])

        def f(u):           # exact rational arithmetic: the 1e18 weight swamps 1/2 u^2 in float64
            u = Fraction(u)
            s = max(Fraction(0), max(Fraction(a) * u - Fraction(c) for c in cs))
            return u * u / 2 + Fraction(10 ** 18, 2) * s * s
        l, h = lo, hi
        for _ in range(120):
            m1, m2 = l + (h - l) / 3, h - (h - l) / 3
            if f(m1) <= f(m2):
                h = m2
            else:
                l = m1
        u_num = 0.5 * (l + h)
        u, act = orc.qp(a, c_lead, c_adj, has_adj, lo, hi)
        assert lo <= u <= hi
        assert f(u) <= f(u_num) * (1 + Fraction(1, 10 ** 9)) + Fraction(1, 10 ** 18), (a, cs, lo, hi, u, u_num)
        assert abs(u - u_num) <= 1e-6, (a, cs, lo, hi, u, u_num)
        n_checked += 1
        n_active += int(act != 0)
    assert n_checked == 1000 and 100 < n_active < 900


def test_oracle_threads_do_not_change_results():
    """The multi-threaded oracle is the CPU baseline bench.py times: any thread count gives the bits of one thread."""
    import copy
    import marl_mass_b200 as mm
    cfg = dict(mm.DEFAULT_CONFIG, safety_guarantee="cbf-cav", traffic_type="mixed", traffic_density=3, HEADWAY_TIME=0.5,
               cbf_eta=0.03125)
    st1 = mm.spawn.spawn_state(list(range(300)), 3, "mixed")
    st7 = copy.deepcopy(st1)
    ocfg = orc.make_config(cfg)
    rng = np.random.RandomState(3)
    for t in range(12):
        a = rng.randint(0, 5, size=(300, 12)).astype(np.int8)
        o1 = orc.step(ocfg, st1, a, n_threads=1)
        o7 = orc.step(ocfg, st7, a, n_threads=7)
        for k in o1:
            assert np.array_equal(o1[k], o7[k], equal_nan=True), (t, k)
    for k in st1:
        assert np.array_equal(st1[k], st7[k]), k


@pytest.mark.parametrize("name", SUPERVISED_CASES)
def test_supervisor_restatements_match_the_reference(name):
    """oracle/supervisor.py (central_layer.py / decentralised_dmc.py + mdp_controller.py + idm_controller.py restated as
    pure functions of scene, action tuple and random draws) reproduces the tuple the reference's `safety_supervisor` /
    `safety_layer_dmc` handed to _simulate on every step of the fixtures, including the ~25 % of steps on which an
    action was replaced."""
    import supervisor as sup
    g, cfg = load_golden(name)
    fn = sup.priority_supervisor if cfg["safety_guarantee"] == "priority" else sup.dmc_supervisor
    rows = g["row_of_step"]
    st = golden_state(g, rows)
    replaced = 0
    for t in range(len(rows)):
        n = int(st["n_cav"][t])
        want = [int(x) for x in g["new_act"][t, :n]]
        got = fn(st, t, g["act"][t], g["rand_draws"][t], cfg["HEADWAY_TIME"])
        assert got == want, (t, [int(x) for x in g["act"][t, :n]], got, want)
        replaced += int(want != [int(x) for x in g["act"][t, :n]])
    assert replaced >= 100
