"""Stand-in for `cvxopt` (conda pin cvxopt=1.2.7, marl_cav.yml:14 — NOT vendored in
/root/reference and not installable here: no network).

Only the call sites on the shield path are covered
(highway_env/vehicle/safety/cbf.py:2-3,44,47,125-126,133-135,140):
`matrix(ndarray, tc="d")` and `solvers.qp(P, q, G, h, A, b)`.

`solvers.qp` restates the published problem   min 1/2 x'Px + q'x  s.t. Gx <= h
for the only shape the reference ever builds (cbf.py:288-322, 374-422):
x = (u0, u1, s), P = diag(1, 1, 1e18), q = 0 and rows

    [ a, 0, -1] x <= c_lead          (barrier row, slack s)
    [ 1, 0,  0] x <= hi
    [-1, 0,  0] x <= -lo
    [ a, 0, -1] x <= c_adj           (optional, MASS with constrain_adj)

Its unique minimiser (slack weight 1e18 treated as "slack only when infeasible")
is the 1-D clamp documented in SURVEY.md §8a-Q.  PARITY UNPINNED against real
cvxopt: an interior-point solve differs by O(1e-7) and `status` cannot be
reproduced.  Test infrastructure only.
"""
import numpy as np
from . import solvers  # noqa: F401


def matrix(x, size=None, tc="d"):
    return np.array(x, dtype=np.float64)
