"""pymc_bart_b200 — B200-native PGBART behind pymc-bart's BART / PGBART API.

Public names mirror pymc_bart/__init__.py:20-45 for the hot path: ``BART`` and the step
class the reference imports from bartrs (``PGBART``), plus the prediction glue.  The
matplotlib analytics of pymc_bart/utils.py are out of scope.
"""
from .bart import BART, BARTRV  # noqa: F401
from .utils import PosteriorSampler, _decode_vi, _encode_vi, _get_posterior_sampler, _sample_posterior  # noqa: F401

__version__ = "0.1.0"
__all__ = ["BART", "PGBART", "PosteriorSampler", "sample"]


def __getattr__(name):  # PGBART / sample import torch lazily
    if name == "PGBART":
        from .pgbart import PGBART

        return PGBART
    if name == "sample":
        from .sampling import sample

        return sample
    raise AttributeError(name)
