"""Import stub for `gym` (absent from this image) — just enough surface for
/root/reference/highway_env to import and run head-less.  Test infrastructure
only: used by oracle/refharness to drive the UNMODIFIED reference when golden
fixtures are generated; never imported by the product path."""
from . import spaces, logger, envs  # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


def make(env_id, **kwargs):
    return envs.registration.make(env_id, **kwargs)
