"""Import stub for `pygame`: only the names the reference touches at import time."""


class Surface(object):
    def __init__(self, *a, **k):
        pass


SurfaceType = Surface
SRCALPHA = 0
KEYDOWN = KEYUP = QUIT = 0
K_l = K_o = K_m = K_k = K_RIGHT = K_LEFT = K_DOWN = K_UP = 0


class event(object):
    class EventType(object):
        pass

    @staticmethod
    def get():
        return []
