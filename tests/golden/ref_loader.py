"""Import the UNMODIFIED reference (jtaghia/ciMRGP) from /root/reference/src.

Only used by the golden-fixture generator (tests/golden/make_golden.py) and by
optional CPU tests that are skipped when /root/reference is absent (it does not
exist on the GPU box).  Two import shims, applied before importing and touching
no reference file:
  1. scipy.misc.logsumexp alias  (Stats.py:4 imports it from scipy.misc)
  2. stub GPy / gpflow modules   (RegressionInput.py:4-5, pulled in by Inputs.py:1)
"""
import os
import sys
import types
import warnings

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


def load():
    if not available():
        raise RuntimeError("reference sources not present at %s" % REF_SRC)
    import scipy.misc
    import scipy.special
    scipy.misc.logsumexp = scipy.special.logsumexp
    for name in ("GPy", "gpflow"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    warnings.filterwarnings("ignore")
    mods = {}
    for name in ("IndexSetGenerator", "KernelClass", "MRGP", "BasisInterval", "Posteriors", "Stats",
                 "Priors", "CommonDensities", "computeRealBinghamConstant", "SanityCheck", "LatentOutputs"):
        mods[name] = __import__(name)
    return types.SimpleNamespace(**mods)
