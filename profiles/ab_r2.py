"""Round-2 A/B timing: python profiles/ab_r2.py workload envs name[:variant] ...

Each entry is timed in its own process: `name` = "main" (the default library) or marl-mass_b200/_build/variants/lib_<name>.so,
`variant` = mm_set_step_variant value (0 automatic, 3 / 4 generic builds forced, 5 automatic without the specialised builds).
Prints the mean device time per policy step over the WHOLE second episode (steps 100..199 after a synchronous spawn: every
phase of an episode with equal weight, the same mix as bench.py's staggered steady state) and a few single steps."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import marl_mass_b200 as mm
from bench import WORKLOADS
name, E, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = dict(mm.DEFAULT_CONFIG, **WORKLOADS[name]["cfg"])
mm.set_step_variant(variant)
env = mm.MergeEnvBatched(E, cfg)
env.reset(seed=1)
gen = torch.Generator(device="cuda").manual_seed(0)
acts = [torch.randint(0, 5, (E, mm.MAXV), generator=gen, device="cuda", dtype=torch.int8) for _ in range(4)]
times = []
for t in range(200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(acts[t %% 4], auto_reset=True); b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
s = env.stats()
print("RESULT build %%d episode-mean %%.4f ms/step (agent-steps/s %%.3e) at t=101,105,120,160,199: %%s" %% (
    env.step_build(), np.mean(times[100:200]), s["agent_steps"] / 200 / (np.mean(times[100:200]) * 1e-3) if False else
    float(env.buffers()["n_agents"].sum()) / (np.mean(times[100:200]) * 1e-3), [round(times[i], 3) for i in (101, 105, 120, 160, 199)]))
''' % ROOT
name, envs, entries = sys.argv[1], sys.argv[2], sys.argv[3:]
for rep in range(2):
    for ent in entries:
        lib, _, variant = ent.partition(":")
        env = dict(os.environ)
        if lib != "main":
            env["MM_LIB_PATH"] = os.path.join(ROOT, "marl-mass_b200", "_build", "variants", "lib_%s.so" % lib)
        out = subprocess.run([sys.executable, "-c", CHILD, name, envs, variant or "0"], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        print("%-10s %-9s %-12s %s" % (name, envs, ent, line[0][7:] if line else "FAILED " + out.stderr[-600:]), flush=True)
