"""triple_accel_b200 -- B200 (sm_100a) batched edit-distance engine with the `triple_accel` crate's interface.

The product is libtriple_accel_b200.so (C ABI in include/triple_accel_b200.h, CUDA sources in csrc/).  This
package is the Python host mirror of the crate's public names used by the tests and the bench.
"""
from .api import (Edit, EditCosts, EditType, Engine, LEVENSHTEIN_COSTS, Match, RDAMERAU_COSTS, SearchType, TA_NONE,
                  TripleAccelError, default_engine, hamming, hamming_batch, hamming_search, hamming_search_batch,
                  hamming_search_simd, hamming_search_simd_with_opts, levenshtein, levenshtein_exp,
                  levenshtein_exp_batch, levenshtein_exp_with_opts, levenshtein_k_batch, levenshtein_search,
                  levenshtein_search_batch, levenshtein_search_simd, levenshtein_search_simd_with_opts,
                  levenshtein_simd_k, levenshtein_simd_k_str, levenshtein_simd_k_with_opts, pack, rdamerau, rdamerau_exp)
from .api import (alloc_str, fill_str, hamming_naive, hamming_search_naive, hamming_search_naive_with_opts,  # noqa: E402
                  hamming_simd_movemask, hamming_simd_parallel, hamming_words_64, hamming_words_128,
                  levenshtein_naive, levenshtein_naive_k, levenshtein_naive_k_with_opts,
                  levenshtein_naive_with_opts, levenshtein_search_naive, levenshtein_search_naive_with_opts,
                  levenstein_naive_str)

__all__ = [n for n in dir() if not n.startswith("_")]
