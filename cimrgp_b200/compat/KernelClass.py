from cimrgp_b200.KernelClass import *  # noqa: F401,F403  (bare-name drop-in for the reference's src/KernelClass.py)
