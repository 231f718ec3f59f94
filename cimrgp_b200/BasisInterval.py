"""Basis intervals: drop-in for the reference's src/BasisInterval.py constructor and static interval rule.
The per-sweep interval optimisation (`learn`, BasisInterval.py:18-134) runs on the device: the model reads
`use_prior` and `opt_interval_factor` from the object of each layer and hands them to
mrgp_set_adaptive_intervals (include/cimrgp.h); the bounded scalar search of all regions of a layer runs in
lock-step as CUDA kernels inside the sweep."""
import numpy as np


class BasisInterval(object):
    def __init__(self, use_prior=True, opt_interval_factor=(1., 1.2)):
        self.basis_interval = None
        self.basis_function_obj = None
        self.spectral_density_obj = None
        self.use_prior = use_prior
        self.opt_interval_factor = opt_interval_factor

    def max_input_range_by_factor_of(self, inputs, factor):
        # BasisInterval.py:15-16
        return factor * np.max(np.abs(inputs), axis=0)
