"""cimrgp_b200: the ciMRGP / fiMRGP variational-inference hot path on one or more B200s.

Module names mirror the reference's src/ (MRGP, IndexSetGenerator, KernelClass, BasisInterval); the
arithmetic lives in libcimrgp.so (csrc/, C ABI in include/cimrgp.h).  To use the package as a drop-in for
scripts that `import MRGP` etc. by bare name, put `cimrgp_b200/compat` on sys.path."""
from .BasisInterval import BasisInterval
from .IndexSetGenerator import IndexSetUniform
from .KernelClass import LaplacianEigenpairs, MaternKernel

__all__ = ['MultiResolutionGaussianProcess', 'IndexSetUniform', 'LaplacianEigenpairs', 'MaternKernel', 'BasisInterval',
           'SeriesBatch']


def __getattr__(name):
    if name == 'MultiResolutionGaussianProcess':
        from .MRGP import MultiResolutionGaussianProcess
        return MultiResolutionGaussianProcess
    if name == 'SeriesBatch':
        from .batch import SeriesBatch
        return SeriesBatch
    raise AttributeError(name)
