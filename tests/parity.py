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
