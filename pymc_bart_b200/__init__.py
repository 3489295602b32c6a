"""pymc_bart_b200 — B200-native PGBART behind pymc-bart's BART / PGBART API.

Public names mirror pymc_bart/__init__.py:20-45 for the hot path: ``BART`` and the step
class the reference imports from bartrs (``PGBART``), plus the prediction glue.  The
matplotlib analytics of pymc_bart/utils.py are out of scope.
"""
from .bart import BART, BARTRV  # noqa: F401
from .utils import PosteriorSampler, _decode_vi, _encode_vi, _get_posterior_sampler, _sample_posterior  # noqa: F401

try:   # what `import bartrs` does for the reference (pymc_bart/__init__.py:15-18): register PGBART with PyMC when it is there
    import pymc as _pm  # noqa: F401
except ImportError:
    _pm = None
if _pm is not None:
    from . import pymc_adapter  # noqa: F401  (appends the step class to pm.STEP_METHODS)

__version__ = "0.2.0"
__all__ = ["BART", "PGBART", "PosteriorSampler", "sample", "sample_joint", "compute_variable_importance", "get_variable_inclusion"]


def __getattr__(name):  # PGBART / sample import torch lazily
    if name == "PGBART":
        from .pgbart import PGBART

        return PGBART
    if name in ("sample", "sample_joint"):
        from . import sampling

        return getattr(sampling, name)
    if name in ("compute_variable_importance", "get_variable_inclusion", "vi_to_kulprit"):
        from . import importance

        return getattr(importance, name)
    raise AttributeError(name)
