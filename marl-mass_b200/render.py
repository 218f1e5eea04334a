"""CPU-side frame of one merge scene, fed from `get_state()` (replaces the pygame viewer).

The reference's `env.render(mode="rgb_array")` (abstract.py:512-556, graphics.py:92-145) returns the simulation surface
as uint8 `[screen_height, screen_width, 3]`; MAPPO.evaluation only feeds it to a video recorder (mappo.py:292-322).
This module rasterises the same view with numpy: the window is centred on the fixed world point (310, 4)
(graphics.py:142-144) with `centering_position` and `scaling` [px/m] from the env config (merge_env_v1.py:42-45),
lanes drawn from the fixed merge network (road.py geometry), vehicles as 5 x 2 m rectangles
turned by their heading, coloured as the reference colours them (vehicle/graphics.py: controlled green, others blue,
crashed red), the obstacle at (420, 4).  Not pixel-identical to pygame's anti-aliased drawing - same geometry.
"""
import numpy as np

from .spawn import KIND_CAV, LANE_LEN, LANE_SX, LANE_SY, L_KB0

GREY, WHITE = (100, 100, 100), (255, 255, 255)
GREEN, BLUE, RED, OBSTACLE = (50, 200, 0), (100, 200, 255), (255, 100, 100), (200, 200, 0)
VEH_LENGTH, VEH_WIDTH, LANE_WIDTH = 5.0, 2.0, 4.0
WINDOW_CENTRE = (310.0, 4.0)
OBSTACLE_POS = (420.0, 4.0)


def _lane_offset(lane, s):
    """lateral position of the lane centre at longitudinal s (straight lanes: 0; kb0: the sine of lane.py:196-210)"""
    if lane == L_KB0:
        return 3.25 * np.sin(np.pi / 100.0 * s + np.pi / 2)
    return np.zeros_like(s)


def _fill_rect(img, origin, scaling, cx, cy, heading, length, width, colour):
    """fill the rectangle centred at world (cx, cy), turned by heading, into img"""
    H, W, _ = img.shape
    half = 0.5 * np.hypot(length, width)
    x0 = max(int(np.floor((cx - half - origin[0]) * scaling)), 0)
    x1 = min(int(np.ceil((cx + half - origin[0]) * scaling)) + 1, W)
    y0 = max(int(np.floor((cy - half - origin[1]) * scaling)), 0)
    y1 = min(int(np.ceil((cy + half - origin[1]) * scaling)) + 1, H)
    if x0 >= x1 or y0 >= y1:
        return
    px = (np.arange(x0, x1) + 0.5) / scaling + origin[0] - cx
    py = (np.arange(y0, y1) + 0.5) / scaling + origin[1] - cy
    gx, gy = np.meshgrid(px, py)
    c, s = np.cos(heading), np.sin(heading)
    lon = c * gx + s * gy
    lat = -s * gx + c * gy
    inside = (np.abs(lon) <= length / 2) & (np.abs(lat) <= width / 2)
    img[y0:y1, x0:x1][inside] = colour


def render_scene(state, index=0, config=None):
    """uint8 [screen_height, screen_width, 3] frame of env `index` of an env-major state dict (`get_state()`)."""
    config = config or {}
    W = int(config.get("screen_width", 600))
    H = int(config.get("screen_height", 120))
    scaling = float(config.get("scaling", 3))
    cpos = config.get("centering_position", [0.3, 0.5])
    origin = (WINDOW_CENTRE[0] - cpos[0] * W / scaling, WINDOW_CENTRE[1] - cpos[1] * H / scaling)
    img = np.empty((H, W, 3), np.uint8)
    img[:] = GREY
    # lane borders: both edges of every lane, one world sample per pixel column
    xs = (np.arange(W) + 0.5) / scaling + origin[0]
    for lane in range(6):
        s = xs - LANE_SX[lane]
        on = (s >= 0) & (s <= LANE_LEN[lane])
        centre = LANE_SY[lane] + _lane_offset(lane, s)
        for edge in (-LANE_WIDTH / 2, LANE_WIDTH / 2):
            rows = np.floor((centre + edge - origin[1]) * scaling).astype(int)
            ok = on & (rows >= 0) & (rows < H)
            img[rows[ok], np.arange(W)[ok]] = WHITE
    _fill_rect(img, origin, scaling, OBSTACLE_POS[0], OBSTACLE_POS[1], 0.0, 2.0, 2.0, OBSTACLE)
    n = int(state["n_veh"][index])
    for i in range(n):
        crashed = bool(state["crashed"][index, i])
        colour = RED if crashed else (GREEN if int(state["kind"][index, i]) == KIND_CAV else BLUE)
        _fill_rect(img, origin, scaling, float(state["x"][index, i]), float(state["y"][index, i]),
                   float(state["heading"][index, i]), VEH_LENGTH, VEH_WIDTH, colour)
    return img


def save_png(path, img):
    """write a uint8 [H, W, 3] frame as a PNG (zlib only: no imaging library in the image)"""
    import struct
    import zlib
    H, W, _ = img.shape
    raw = b"".join(b"\x00" + np.ascontiguousarray(img[r]).tobytes() for r in range(H))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
