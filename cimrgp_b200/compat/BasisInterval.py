from cimrgp_b200.BasisInterval import *  # noqa: F401,F403  (bare-name drop-in for the reference's src/BasisInterval.py)
