#!/usr/bin/env python
"""bench.py — shielded agent-steps/s of the batched merge env + CBF shield on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--envs E] [--impl reference]

A "step" is one policy step (3 physics sub-steps, 3 shield solves per CAV, observation, rewards) of every env of
the workload, with finished envs re-spawned on the device (auto-reset), under i.i.d. uniform meta-actions.
Prints ONE JSON line (rank 0).  Keys: see the task contract; in short
  value      whole-job agent-steps/s with actions already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the host-buffer C-ABI call mm_step_host (pinned host actions in, obs/reward/
             done/regional rewards out every step, copies inside the timed region)
  roofline   step kernel alone: algorithmic bytes/launch / mean launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (C float64 port of the reference path) on this box's host cores, bounded sample
`--impl reference` times that CPU port as the reference arm (the Python reference cannot travel to the GPU box;
its in-container throughput is recorded in BASELINE.md / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

MASS = dict(safety_guarantee="cbf-cav", HEADWAY_TIME=0.5, cbf_eta=0.03125)
WORKLOADS = {
    # BASELINE.json configs[4] scenario (the one the 1/2/4/8-GPU metric and the 1e8 target are quoted on):
    # MASS shield, hard density (td3), all-CAV (7-11 CAVs per env); one 2^20-env shard per GPU.
    "mass_td3": dict(cfg=dict(MASS, traffic_density=3, traffic_type="cav"), envs=1 << 20,
                     desc="MASS cbf-cav hard-density merge (BASELINE configs[4]: marl_cav-heading-t_headway-cbf-cav, "
                          "traffic_density=3, all-CAV 7-11 agents/env)"),
    # opt-in device-spawn variant: vehicle counts drawn per 128-env tile (include/marl_mass_b200.h couple_counts);
    # NOT the default workload - reported separately (profiles/README.md)
    "mass_td3_coupled": dict(cfg=dict(MASS, traffic_density=3, traffic_type="cav", couple_vehicle_counts=True), envs=1 << 20,
                             desc="MASS cbf-cav hard density, all-CAV, vehicle counts coupled per 128-env tile (opt-in spawn "
                                  "variant: every env keeps the reference's law, envs of a tile share n_CAV)"),
    "mass_td3_mixed": dict(cfg=dict(MASS, traffic_density=3, traffic_type="mixed"), envs=1 << 20,
                           desc="MASS cbf-cav-mixed hard density (4-6 CAV + 3-5 IDM/MOBIL HDV per env)"),
    "mass_td3_srew": dict(cfg=dict(MASS, traffic_density=3, traffic_type="cav", agent_reward="srew",
                                   HIGH_SPEED_REWARD=4, HEADWAY_COST=1, MERGING_LANE_COST=8), envs=65536,
                          desc="MASS td3 srew (BASELINE configs[3] env side), 65536 envs"),
    "mass_td1": dict(cfg=dict(MASS, traffic_density=1, traffic_type="cav"), envs=65536,
                     desc="MASS cbf-cav td1 (BASELINE configs[2]), 65536 envs"),
    "hss_td3": dict(cfg=dict(safety_guarantee="cbf-avs_cint", HEADWAY_TIME=0.5, cbf_eta=0.03125, traffic_density=3,
                             traffic_type="cav"), envs=4096,
                    desc="HSS cbf-avs_cint td3 (BASELINE configs[1]), 4096 envs"),
    "unsafe_td1": dict(cfg=dict(safety_guarantee="none", HEADWAY_TIME=1.2, traffic_density=1, traffic_type="cav"),
                       envs=65536, desc="no shield, td1 (BASELINE configs[0] LC-env sibling)"),
}

# DESIGN.md "Algorithmic bytes": per vehicle 84 B state read + 124 B state written, per agent 120 B obs + 8 B
# rewards + 1 B done + 1 B action, per env 33 B of scalars
BYTES_PER_VEHICLE = 84 + 124
BYTES_PER_AGENT = 120 + 8 + 1 + 1
BYTES_PER_ENV = 33


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for row in self.samples:
            if len(row) < 6:
                continue
            try:
                sm.append(float(row[0]))
                mx = float(row[1])
            except ValueError:
                continue
            for n, v in zip(names, row[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_throughput(cfg, budget_s, sample_envs, warmup_steps=1, fixed_steps=None):
    """Time the CPU oracle (C port of the reference path) with all host threads on a bounded sample.
    Start states come from the host seed-exact spawn; envs are re-spawned on the host when done."""
    import oracle
    import marl_mass_b200.spawn as spawn
    from marl_mass_b200 import DEFAULT_CONFIG
    full = dict(DEFAULT_CONFIG, **cfg)
    cores = host_cores()
    ocfg = oracle.make_config(full)
    E = sample_envs
    base = spawn.spawn_state(range(64), full["traffic_density"], full["traffic_type"])
    st = {k: np.ascontiguousarray(np.tile(v, (E // 64 + 1,) + (1,) * (v.ndim - 1))[:E]) for k, v in base.items()}
    fresh = {k: v.copy() for k, v in st.items()}
    rng = np.random.RandomState(0)
    out = oracle.empty_out(E)
    agent_steps, t_total, n = 0, 0.0, 0
    per_step = []
    while True:
        a = rng.randint(0, 5, size=(E, 12)).astype(np.int8)
        n_live = int(st["n_cav"].sum())
        t0 = time.perf_counter()
        oracle.step(ocfg, st, a, out=out, n_threads=cores)
        dt = time.perf_counter() - t0
        done = out["done"] != 0
        if done.any():  # host-side re-spawn (untimed, like the reference's env.reset between episodes)
            for k in st:
                st[k][done] = fresh[k][done]
        n += 1
        if n > warmup_steps:
            agent_steps += n_live
            t_total += dt
            per_step.append(dt)
        if fixed_steps is not None:
            if n >= warmup_steps + fixed_steps:
                break
        elif t_total >= budget_s:
            break
    return agent_steps / t_total, cores, len(per_step), E, float(np.mean(per_step)) * 1e3


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(wl["envs"], 32768)
    v, cores, n_steps, E, ms = cpu_port_throughput(wl["cfg"], None, sample, warmup_steps=args.warmup,
                                                   fixed_steps=args.steps)
    line = {
        "impl": "reference", "metric": "shielded agent-steps/s", "value": v, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "envs_per_step": E,
                   "note": "CPU arm: C float64 port of the reference env+shield (oracle/), all host threads; the "
                           "Python reference itself cannot travel to the GPU box (see BASELINE.md for its "
                           "in-container throughput)"},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d envs x %d policy steps" % (E, n_steps)},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def claim_stdout():
    """Keep stdout to the ONE JSON line: everything any library prints to fd 1 from here on (NCCL's version banner,
    torchrun notices, ...) goes to stderr; `emit` writes the line to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mass_td3", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer pass (default min(steps, 20))")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--policy", action="store_true",
                    help="BASELINE configs[3]: sample the actions from the MAPPO actor (30-128-128-5) on the device "
                         "inside the timed region instead of reading pre-drawn uniform actions")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.envs:
        wl["envs"] = args.envs
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    import marl_mass_b200 as mm
    from marl_mass_b200 import dist as mmd

    rank, world, local_rank = mmd.rank_world()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if "MM_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["MM_NCCL_DEBUG"]
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    numa_cores = mmd.bind_to_gpu(dev) if world > 1 else 0   # host buffers on the GPU's own NUMA node
    E = wl["envs"]
    cfg = dict(mm.DEFAULT_CONFIG, **wl["cfg"])
    env = mm.MergeEnvBatched(E, cfg, device=dev, record_diag=False)
    env.reset(seed=mmd.rank_seed(1, rank))
    K, W = args.steps, args.warmup

    # action pool resident in HBM (i.i.d. uniform{0..4}); cycled so every step reads a different 12 MB slab
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    pool = [torch.randint(0, 5, (E, mm.MAXV), generator=gen, device="cuda", dtype=torch.int8) for _ in range(8)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Prologue (untimed): stagger the episode phases.  All envs are spawned together, so without this every env
    # would hit the 100-step horizon on the same step; re-spawning the envs with (index % T == j) at prologue
    # step j spreads the phases uniformly, which is the steady state of a long rollout (1/T of the envs re-spawn
    # per step inside the timed region).
    T = int(cfg["duration"] * cfg["policy_frequency"])
    # Stagger by 128-env tile, not by env: in a real rollout every env of the batch starts together and only the
    # rare crash de-synchronises one, so neighbouring envs share their episode phase; tile-granular staggering
    # keeps that property while making every timed step the average over all phases of an episode.
    idx = (torch.arange(E, device="cuda", dtype=torch.int32) // 128) % T
    for j in range(T):
        env.step(pool[j % 8], auto_reset=True)
        env.reset(seed=mmd.rank_seed(3 + j, rank), mask=(idx == j).to(torch.uint8))
    for t in range(W):
        env.step(pool[t % 8], auto_reset=True)
    barrier()
    env.stats(reset=True)
    l0 = env.kernel_launches()

    sampler = ClockSampler(dev)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    policy = None
    if args.policy:
        from marl_mass_b200.rollout import BatchedMAPPORollout
        torch.manual_seed(rank)
        policy = BatchedMAPPORollout(env, roll_out_n_steps=1)
        vbuf = env.buffers()
        for t in range(3):
            env.step(policy._act(vbuf["obs"], vbuf["n_agents"])[0], auto_reset=True)
    barrier()
    ev0.record()
    if policy is None:
        for t in range(K):
            env.step(pool[t % 8], auto_reset=True)
    else:
        for t in range(K):
            env.step(policy._act(vbuf["obs"], vbuf["n_agents"])[0], auto_reset=True)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = env.kernel_launches() - l0
    stats = env.stats(reset=True)

    # roofline pass: the step kernel alone, one event pair per launch (re-spawn outside the pair)
    kern_ms = []
    agent_steps_k, veh_steps_k = 0.0, 0.0
    v = env.buffers()
    for t in range(min(K, 30)):
        n_ag = float(v["n_agents"].sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step(pool[t % 8], auto_reset=False)
        b.record()
        torch.cuda.synchronize()
        kern_ms.append(a.elapsed_time(b))
        agent_steps_k += n_ag
        env.reset(seed=mmd.rank_seed(2, rank), mask=v["done"])
    torch.cuda.synchronize()
    clocks = sampler.stop()

    # CBF-QP alone (the metric's "CBF-QP solves/s"): 2^26 synthetic solves, ~10 % of rows active
    nq = 1 << 26
    g2 = torch.Generator(device="cuda").manual_seed(7)
    dtq = 1.0 / 15
    qa = dtq * torch.cos(torch.rand(nq, generator=g2, device="cuda", dtype=torch.float64) * 0.6 - 0.3)
    qcl = torch.randn(nq, generator=g2, device="cuda", dtype=torch.float64) * 3 + 2
    qca = torch.randn(nq, generator=g2, device="cuda", dtype=torch.float64) * 3 + 2
    qha = (torch.rand(nq, generator=g2, device="cuda") < 0.3).to(torch.uint8)
    qlo = torch.full((nq,), -12.5 * dtq, device="cuda", dtype=torch.float64) + 1e-3 * torch.rand(nq, generator=g2, device="cuda", dtype=torch.float64)
    qhi = qlo + 18.5 * dtq
    for _ in range(3):
        mm.shield_qp(qa, qcl, qca, qha, qlo, qhi)
    qp_runs = []
    for _ in range(4):   # best of 4 bursts of 10 launches (HBM-bound kernel timed alone -> burst peak applies)
        qe0, qe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        qe0.record()
        for _ in range(10):
            qu, qact = mm.shield_qp(qa, qcl, qca, qha, qlo, qhi)
        qe1.record()
        torch.cuda.synchronize()
        qp_runs.append(qe0.elapsed_time(qe1) / 10)
    qp_ms = min(qp_runs)
    qp_active = float((qact != 0).float().mean())
    del qa, qcl, qca, qha, qlo, qhi, qu, qact

    # e2e pass: host buffers through mm_step_host
    K2 = args.e2e_steps or min(K, 20)
    host_out = env.alloc_host_out(pinned=True)
    host_act = [torch.randint(0, 5, (E, mm.MAXV), dtype=torch.int8).pin_memory().numpy() for _ in range(4)]
    env.step_host(host_act[0], auto_reset=True, out=host_out)
    env.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    for t in range(K2):
        env.step_host(host_act[t % 4], auto_reset=True, out=host_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_stats = env.stats(reset=True)
    del host_out
    # the same with the observations packed as the reference returns them (live agents' rows only): mm_step_host_ragged
    rag_out = env.alloc_host_out(pinned=True, ragged=True)
    env.step_host_ragged(host_act[0], auto_reset=True, out=rag_out)
    env.stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    for t in range(K2):
        env.step_host_ragged(host_act[t % 4], auto_reset=True, out=rag_out)
    torch.cuda.synchronize()
    rag_s = time.perf_counter() - t0
    rag_stats = env.stats(reset=True)
    rag_rows = float(rag_out["n_agents"].sum())

    # fold over ranks: time = max, work = sum
    tot = mmd.all_reduce_stats(stats)
    e2e_tot = mmd.all_reduce_stats(e2e_stats)
    rag_tot = mmd.all_reduce_stats(rag_stats)
    if world > 1:
        tmax = torch.tensor([ms_total, e2e_s, rag_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, rag_s = float(tmax[0]), float(tmax[1]), float(tmax[2])
    value = tot["agent_steps"] / (ms_total * 1e-3)
    dense_value = e2e_tot["agent_steps"] / e2e_s
    e2e_value = rag_tot["agent_steps"] / rag_s

    mean_agents = stats["agent_steps"] / max(stats["env_steps"], 1.0)
    mean_hdv = 0.0 if cfg["traffic_type"] == "cav" else {1: 2.0, 2: 3.0, 3: 4.0}[int(cfg["traffic_density"])]
    veh_per_env = mean_agents + mean_hdv
    kms = float(np.mean(kern_ms))
    bytes_per_launch = E * (veh_per_env * BYTES_PER_VEHICLE + mean_agents * BYTES_PER_AGENT + BYTES_PER_ENV)
    peak, peak_src = load_peaks()
    achieved = bytes_per_launch / (kms * 1e-3) / 1e9
    # DRAM traffic of the step kernel from the committed ncu capture (profiles/r1_traffic.json), scaled per env
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if tj.get("workload") == args.workload:
            traffic = tj["dram_bytes"] * (E / float(tj["envs"]))
    except Exception:
        traffic = None

    line = {
        "metric": "shielded agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "envs_per_gpu": E, "mean_agents_per_env": round(mean_agents, 3),
                   "actions": "i.i.d. uniform{0..4}, resident in HBM", "auto_reset": True,
                   "episode_phases": "staggered uniformly over the 100-step episode, per 128-env tile (untimed prologue)",
                   "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (E * 3.0e-6),
                   "step_kernel_build": "automatic (3 CTAs/SM; 4 CTAs/SM for grids of 3..8 CTAs per SM whose wave "
                                        "structure favours it, include/marl_mass_b200.h mm_set_step_variant)",
                   "shield_solves_per_s": tot["shield_solves"] / (ms_total * 1e-3),
                   "shield_active_frac": tot["shield_active"] / max(tot["shield_solves"], 1.0),
                   "lane_change_veto_frac": tot["lane_change_vetoes"] / max(tot["shield_solves"], 1.0),
                   "crashed_episode_frac": tot["crashed_episodes"] / max(tot["episodes"], 1.0)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": E * mm.MAXV,
                "d2h_bytes_per_step": int(rag_rows * mm.NS * 4 + E * (8 + 4 + 1 + mm.MAXV * 4 + 4)), "steps": K2,
                "api": "mm_step_host_ragged (pinned host buffers, 64Ki-env chunks round-robin on 4 streams; observation "
                       "rows of the live agents only, as the reference returns them: packed on the device, exact-size copies)",
                "host_cores_bound": numa_cores,
                "dense": {"value": dense_value, "d2h_bytes_per_step": E * (mm.MAXV * mm.NS * 4 + 4 + 1 + mm.MAXV * 4 + 4),
                          "api": "mm_step_host (dense obs [E,12,30] by cudaMemcpyAsync)"}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "step_kernel<false, false>", "kernel_ms": kms,
                     "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                     "note": "issue/latency-bound f64 kernel (SURVEY.md 8d): HBM fraction is reported as asked; "
                             "see profiles/ for pipe utilisation"},
    }
    line["qp_microbench"] = {"solves_per_s": nq / (qp_ms * 1e-3), "n": nq, "ms": qp_ms, "active_frac": qp_active,
                             "ms_mean": float(np.mean(qp_runs)), "bytes_per_solve": 50, "achieved_GBps": nq * 50 / (qp_ms * 1e-3) / 1e9,
                             "hbm_frac": nq * 50 / (qp_ms * 1e-3) / 1e9 / peak, "dtype": "f64"}
    if args.policy:
        line["config"]["actions"] = ("sampled on device from the MAPPO actor 30-128-128-5 inside the timed region "
                                     "(mm_actor_sample: fused TF32 forward + inverse-CDF draw, one launch per step)")
    if rank == 0 and world == 1 and not args.skip_cpu:
        v_cpu, cores, n_steps, e_cpu, _ = cpu_port_throughput(wl["cfg"], args.cpu_seconds, min(E, 16384))
        line["cpu_baseline"] = {"value": v_cpu, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                "sample": "%d envs x %d policy steps (~%.0f s of CPU work)" % (e_cpu, n_steps, args.cpu_seconds)}
    if rank == 0:
        emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
