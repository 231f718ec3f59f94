from cimrgp_b200.RegressionInput import *  # noqa: F401,F403  (bare-name drop-in for the reference's src/RegressionInput.py)
