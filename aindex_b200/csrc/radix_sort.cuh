// radix_sort.cuh -- internal interface of radix_sort.cu (hand-written 64-bit LSD radix sort + run-length encode).
#pragma once
#include "aix_internal.cuh"

namespace aix {

// Stable sort of n 64-bit keys on bits [begin_bit, end_bit).  `alt` is a spare buffer of n keys; *sorted is set to
// whichever of the two buffers holds the result.  Synchronises `st` before returning (its scratch is freed).
int radix_sort_u64(aix_ctx *ctx, cudaStream_t st, uint64_t *keys, uint64_t *alt, uint64_t n, int begin_bit, int end_bit,
                   uint64_t **sorted);

// Run-length encoding of a sorted array: uniq[r], counts[r] (saturating u32) for the *n_runs distinct keys.
// uniq / counts must hold n entries in the worst case and must not alias `sorted`.
int rle_u64(aix_ctx *ctx, cudaStream_t st, const uint64_t *sorted, uint64_t n, uint64_t *uniq, uint32_t *counts,
            uint64_t *n_runs);

// Stable partition of n keys into n_ranges <= 16 key ranges (bound[r] = first key of range r; bound[0] is taken as 0):
// out = the keys grouped by range, input order kept inside a range; counts[r] (host) = keys in range r.
// Synchronises `st`.
int partition_by_range(aix_ctx *ctx, cudaStream_t st, const uint64_t *keys, uint64_t *out, uint64_t n, const uint64_t *bound,
                       int n_ranges, uint64_t *counts);

}  // namespace aix
