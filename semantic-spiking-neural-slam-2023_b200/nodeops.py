"""Node callables that the backend lowers to device ops instead of calling Python.

The reference places Python callables *inside* the per-step loop (SURVEY.md F10, §8a rows
7-9): an identity lambda (``pathintegration.py:167``), the grid clean-up
(``slam.py:213-215,270``; ``slam_view.py:203-205,259``) and the gated correction
(``slam.py:233-237``; ``slam_view.py:225-229``).  The classes here are explicit,
introspectable equivalents; :func:`recognize` additionally identifies the reference's
own anonymous closures (unmodified source) by owner attribute + closure cells and checks
the match numerically before trusting it.  Anything unrecognised makes the lowering fail
loudly — there is no host-callback path.
"""
from __future__ import annotations

import numpy as np


class Identity:
    kind = "identity"

    def __call__(self, t, x):
        return x


class GridCleanup:
    """``x -> S[argmax_g S[g].x]`` (first maximum wins; ``x = 0`` gives row 0)."""
    kind = "cleanup"

    def __init__(self, sample_ssps):
        self.sample_ssps = np.ascontiguousarray(sample_ssps, dtype=np.float64)

    def __call__(self, t, x):
        return self.sample_ssps[np.argmax(self.sample_ssps @ x)]


class GatedCorrection:
    """``x = [p ; q ; flag]`` -> ``rate*(p-q)`` if ``|flag| <= atol`` and ``p.q > thres`` else 0."""
    kind = "gate"

    def __init__(self, d, shift_rate, update_thres, atol=1e-3):
        self.d, self.shift_rate, self.update_thres, self.atol = int(d), float(shift_rate), float(update_thres), float(atol)

    def __call__(self, t, x):
        d = self.d
        p, q = x[:d], x[d:2 * d]
        # np.allclose(flag, 0, atol=1e-3) == |flag| <= atol + rtol*0
        if abs(x[-1]) <= self.atol and float(np.sum(p * q)) > self.update_thres:
            return self.shift_rate * (p - q)
        return np.zeros(d)


def _closure_vars(fn):
    code = getattr(fn, "__code__", None)
    cells = getattr(fn, "__closure__", None) or ()
    if code is None:
        return {}
    out = {}
    for name, cell in zip(code.co_freevars, cells):
        try:
            out[name] = cell.cell_contents
        except ValueError:
            pass
    return out


def _agrees(fn, op, size_in, rng, n=6, scale=1.0):
    for i in range(n):
        x = rng.standard_normal(size_in) * scale
        if i % 2 == 0 and op.kind == "gate":
            x[-1] = 0.0  # exercise the open-gate branch
            x[op.d:2 * op.d] = x[:op.d] + 0.05 * x[op.d:2 * op.d]
        a = np.asarray(fn(0.001 * (i + 1), x.copy()), dtype=np.float64).reshape(-1)
        b = np.asarray(op(0.001 * (i + 1), x.copy()), dtype=np.float64).reshape(-1)
        if a.shape != b.shape or not np.allclose(a, b, rtol=1e-12, atol=1e-12):
            return False
    return True


def recognize(node, owners):
    """Return a device-op object for ``node.output`` or ``None``.

    ``owners`` are the networks of the model; the reference keeps the constants its
    closures need as attributes of the owning network (``slam.sample_ssps``) or in
    closure cells (``d``, ``shift_rate``, ``update_thres``)."""
    fn = node.output
    if isinstance(fn, (Identity, GridCleanup, GatedCorrection)):
        return fn
    if not callable(fn) or node.size_in == 0:
        return None
    rng = np.random.default_rng(12345)
    if node.size_out == node.size_in and _agrees(fn, Identity(), node.size_in, rng):
        return Identity()
    cv = _closure_vars(fn)
    # gate: closure cells d / shift_rate / update_thres
    if {"d", "shift_rate", "update_thres"} <= set(cv) and node.size_in == 2 * int(cv["d"]) + 1:
        op = GatedCorrection(cv["d"], cv["shift_rate"], cv["update_thres"])
        if _agrees(fn, op, node.size_in, rng, scale=0.3):
            return op
    # clean-up: the lambda closes over clean_up_fun, which closes over sample_ssps
    cands = []
    inner = cv.get("clean_up_fun")
    if inner is not None:
        s = _closure_vars(inner).get("sample_ssps")
        if s is not None:
            cands.append(s)
    for net in owners:
        s = getattr(net, "sample_ssps", None)
        if s is not None and (getattr(net, "gridcells", None) is node or getattr(net, "cleanup", None) is node):
            cands.append(s)
    for s in cands:
        s = np.asarray(s)
        if s.ndim == 2 and s.shape[1] == node.size_in == node.size_out:
            op = GridCleanup(s)
            if _agrees(fn, op, node.size_in, rng):
                return op
    return None
