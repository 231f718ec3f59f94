from cimrgp_b200.MRGP import *  # noqa: F401,F403  (bare-name drop-in for the reference's src/MRGP.py)
