"""smash_paper_b200 -- B200-native implementation of the SMASH mapping+binning hot path.

Product path = CUDA (smash_paper_b200/csrc) behind the C ABI in include/smash_b200.h.
There is no CPU fallback: importing `smash_paper_b200.api` and calling into it without the
compiled library / without a GPU raises.
"""
__version__ = "0.1.0"
