"""A/B timing of experiment builds: python profiles/ab_time.py workload envs variantA variantB ...

Each variant is marl-mass_b200/_build/variants/lib_<name>.so ("main" = the default library), timed in its own
process by profiles/time_steps.py; prints the mean of the steady-state steps 100..129 and a few single steps."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, envs, variants = sys.argv[1], sys.argv[2], sys.argv[3:]
for rep in range(2):
    for v in variants:
        env = dict(os.environ)
        if v != "main":
            env["MM_LIB_PATH"] = os.path.join(ROOT, "marl-mass_b200", "_build", "variants", "lib_%s.so" % v)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "time_steps.py"), name, envs], env=env,
                             capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if "ms/step" in l]
        print("%-24s %s" % (v, line[0] if line else "FAILED " + out.stderr[-400:]), flush=True)
