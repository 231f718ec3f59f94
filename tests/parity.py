"""Comparison rule shared by all parity tests:
|new - ref| <= rtol*|ref| + atol_scale*max|ref| per array (the second term covers exact zeros)."""
import numpy as np


def mismatch(new, ref, rtol, atol_scale=1e-12, atol_abs=0.0):
    new = np.asarray(new, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if new.shape != ref.shape:
        return 'shape %s != %s' % (new.shape, ref.shape)
    if ref.size == 0:
        return None
    scale = np.max(np.abs(ref))
    tol = rtol * np.abs(ref) + atol_scale * scale + atol_abs
    err = np.abs(new - ref)
    bad = ~(err <= tol)
    if np.any(bad):
        k = np.argmax(np.where(bad, err / np.maximum(tol, 1e-300), 0))
        return 'max violation at flat index %d: new=%r ref=%r (|ref|max=%.3e, %d/%d bad)' % (
            k, new.ravel()[k], ref.ravel()[k], scale, int(bad.sum()), bad.size)
    return None


def assert_state_close(new_state, ref_state, rtol, keys=None, skip=(), atol_scale=1e-12, atol_abs=None):
    problems = []
    for key in (keys or sorted(ref_state)):
        if any(key.endswith(s) for s in skip):
            continue
        if key not in new_state:
            problems.append('%s: missing' % key)
            continue
        aa = 0.0 if atol_abs is None else atol_abs.get(key.split('.')[-1], 0.0)
        msg = mismatch(new_state[key], ref_state[key], rtol, atol_scale, aa)
        if msg:
            problems.append('%s: %s' % (key, msg))
    assert not problems, '\n'.join(problems[:40])


# sum_n phi^2 is pure roundoff (|sin(i pi)|^2 ~ 1e-32) for one-sample regions whose point sits on the
# interval edge; both sides are "zero" at the scale of phi^2 <= 1/L
ATOL_ABS = {'d': 1e-20}


def mask_degenerate(state, ref):
    """Regions whose samples all sit on the edge of the basis interval (x = +-L, e.g. one-sample regions:
    L = |x|, BasisInterval.py:15-16) have phi == 0 in exact arithmetic: y_tilde, A and the fi-mode B /
    kappa of such a region are pure roundoff in the reference itself.  They are excluded."""
    state, ref = dict(state), dict(ref)
    for key in [k for k in ref if k.endswith('.d')]:
        layer = key[:-2]
        dead = np.max(ref[key], axis=1) < 1e-20
        if not np.any(dead):
            continue
        for name in ('ytil', 'A', 'B', 'kappa'):
            k2 = layer + '.' + name
            if k2 in ref and ref[k2].shape[0] == dead.shape[0]:
                for d in (state, ref):
                    d[k2] = np.array(d[k2])
                    d[k2][dead] = 0.0
    return state, ref


# a shared bias of an upper fi layer is a sum of residuals that cancel to roundoff (|y| ~ 1): compare it at that scale
ATOL_ABS_SHARED = {'d': 1e-20, 'bias_mean': 1e-12}


def compare(state, ref, rtol=1e-6, atol_abs=None):
    state, ref = mask_degenerate(state, ref)
    assert_state_close(state, ref, rtol, skip=('kappa',), atol_abs=ATOL_ABS if atol_abs is None else atol_abs)
    for key in ref:
        if key.endswith('kappa'):
            assert mismatch(state[key], ref[key], rtol, atol_scale=1e-12) is None, key


def expand_shared(ref_state, like):
    """Goldens of the shared noise / bias variants hold the reference's shapes (a scalar, one (dy,) vector per
    layer); the oracle and the device store a shared posterior once per region.  Broadcast the golden."""
    out = dict(ref_state)
    for k, v in ref_state.items():
        if k in like and np.shape(v) != np.shape(like[k]):
            out[k] = np.broadcast_to(np.asarray(v), np.shape(like[k])).copy()
    return out
