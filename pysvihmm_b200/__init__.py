"""pysvihmm_b200: B200-native local E-step engine for SVI in Bayesian HMMs.

Drop-in for the hot path of dillonalaird/pysvihmm (hmmsgd_metaobs.VBHMM.infer's per-minibatch
E-step + natural-gradient step, hmmbatchcd's batch variant) behind the reference's class surface:

    from pysvihmm_b200 import hmmsgd_metaobs, hmmbatchcd
    hmm = hmmsgd_metaobs.VBHMM(obs, prior_init, prior_tran, prior_emit, metaobs_half=255, mb_sz=256)
    hmm.infer()

All arithmetic of the path runs in hand-written sm_100a CUDA kernels behind the C ABI in
include/svihmm.h (pysvihmm_b200/lib/libsvihmm.so); there is no CPU fallback.
"""
from . import _lib
from ._lib import SvihmmError

__all__ = ["_lib", "SvihmmError"]
__version__ = "0.1.0"
