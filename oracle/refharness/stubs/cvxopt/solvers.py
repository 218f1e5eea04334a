"""Closed-form stand-in for cvxopt.solvers.qp on the CBF-QP shape (see package docstring)."""
import numpy as np

options = {}

# Every solve appends one record here; the fixture generator drains it to freeze
# (a, c_lead, c_adj, lo, hi) -> (u, active) known-answer vectors.
trace = []

ACT_LEAD, ACT_UPPER, ACT_LOWER, ACT_ADJ, ACT_SLACK = 1, 2, 4, 8, 16


def solve_clamp(a, c_lead, lo, hi, c_adj=None):
    """min 1/2 u^2 (+ slack) s.t. a*u - s <= c_k, lo <= u <= hi.  Returns (u, s, active)."""
    active = 0
    u = 0.0
    if u > hi:
        u, active = hi, ACT_UPPER
    if u < lo:
        u, active = lo, ACT_LOWER
    if a > 0.0:
        lim, row = c_lead / a, ACT_LEAD
        if c_adj is not None and c_adj / a < lim:
            lim, row = c_adj / a, ACT_ADJ
        if u > lim:
            if lim >= lo:
                u, active = lim, row
            else:
                u, active = lo, row | ACT_LOWER | ACT_SLACK
    elif a < 0.0:
        lim, row = c_lead / a, ACT_LEAD
        if c_adj is not None and c_adj / a > lim:
            lim, row = c_adj / a, ACT_ADJ
        if u < lim:
            if lim <= hi:
                u, active = lim, row
            else:
                u, active = hi, row | ACT_UPPER | ACT_SLACK
    else:
        c = c_lead if c_adj is None else min(c_lead, c_adj)
        if c < 0.0:
            active |= ACT_SLACK | ACT_LEAD
    c = c_lead if c_adj is None else min(c_lead, c_adj)
    s = max(0.0, a * u - c)
    return u, s, active


def qp(P, q, G, h, A=None, b=None, **kwargs):
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64).ravel()
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).ravel()
    if not (P.shape == (3, 3) and np.all(q == 0.0) and G.shape[1] == 3 and G.shape[0] in (3, 4)
            and A is None and b is None):
        raise NotImplementedError("cvxopt stand-in: unexpected QP shape")
    if not (G[1, 0] == 1.0 and G[2, 0] == -1.0 and G[0, 2] == -1.0 and np.all(G[:, 1] == 0.0)):
        raise NotImplementedError("cvxopt stand-in: unexpected G structure")
    a = float(G[0, 0])
    c_lead = float(h[0])
    hi = float(h[1])
    lo = -float(h[2])
    c_adj = None
    if G.shape[0] == 4:
        if not (G[3, 0] == G[0, 0] and G[3, 2] == -1.0):
            raise NotImplementedError("cvxopt stand-in: unexpected adjacent row")
        c_adj = float(h[3])
    u, s, active = solve_clamp(a, c_lead, lo, hi, c_adj)
    trace.append(dict(a=a, c_lead=c_lead, c_adj=c_adj, lo=lo, hi=hi, u=u, s=s, active=active))
    return {"x": np.array([u, 0.0, s], dtype=np.float64), "status": "optimal", "active": active}
