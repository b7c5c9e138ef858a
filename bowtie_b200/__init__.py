"""bowtie_b200 — B200-native FM-index backward-search path of Bowtie 1.

The Python layer is plumbing over the C ABI in include/bowtie_b200.h (libbowtie_b200.so):
it loads the library, marshals numpy / torch buffers, and fails loudly when the CUDA library
or a GPU is missing.  There is no CPU search path in this package.
"""
from .api import (BT_HIT_HDR_WORDS, OVF_HITS, OVF_MM, Context, Index, Policy, Stats, build_index, build_index_text, build_library, decode_hits, lib_path,
                  load_library)

__all__ = ["Index", "Context", "Policy", "Stats", "load_library", "build_library", "build_index", "build_index_text", "lib_path", "decode_hits",
           "BT_HIT_HDR_WORDS", "OVF_HITS", "OVF_MM"]
