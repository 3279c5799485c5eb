"""Stand-in for the third-party ``lap`` package (C++ Jonker-Volgenant; not in the mount, not installable offline), used ONLY by
tests/golden/make_golden.py to run the unmodified reference adapter on the CPU.  Published semantics of
``lapjv(cost, extend_cost=True, cost_limit=L)``: the optimal assignment of the cost matrix extended to (n+m) x (n+m) with
L/2 'stay unassigned' blocks; x[i] = assigned column or -1, y[j] = assigned row or -1."""
import numpy as np
from scipy.optimize import linear_sum_assignment


def lapjv(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    ext = np.full((n + m, n + m), cost_limit / 2.0)
    ext[n:, m:] = 0.0
    ext[:n, :m] = cost
    rows, cols = linear_sum_assignment(ext)
    x = np.full(n, -1, dtype=int)
    y = np.full(m, -1, dtype=int)
    total = 0.0
    for r, c in zip(rows, cols):
        if r < n and c < m:
            x[r], y[c] = c, r
            total += cost[r, c]
    return total, x, y
