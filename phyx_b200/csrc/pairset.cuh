// phyx_b200 — device hash set of (body1, body2) pair keys: the manifold cache's index.
//
// Reference structure replaced: Collider::manifoldMap, a DenseHashSet<std::pair<unsigned, unsigned>>
// (src/Collider.h:58, src/base/DenseHash.h) — same observable behaviour (keys are in SWEEP order,
// not canonical: SURVEY App. B3) without the tombstone bug (App. B2): the table is rebuilt from the
// live manifolds every step, so it never holds tombstones.
#pragma once

namespace phyx
{

constexpr unsigned long long kEmptyPair = ~0ull;

__device__ __forceinline__ unsigned long long pair_key(unsigned a, unsigned b) { return (static_cast<unsigned long long>(a) << 32) | b; }

__device__ __forceinline__ size_t pair_slot(unsigned long long k, size_t mask)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return size_t(k) & mask;
}

__device__ __forceinline__ bool pair_contains(const unsigned long long* __restrict__ table, size_t mask, unsigned long long key)
{
    for (size_t i = pair_slot(key, mask);; i = (i + 1) & mask)
    {
        unsigned long long v = table[i];
        if (v == key) return true;
        if (v == kEmptyPair) return false;
    }
}

// the same, continuing a probe sequence at slot i (the caller has looked at the slots before it)
__device__ __forceinline__ bool pair_contains_from(const unsigned long long* __restrict__ table, size_t mask, unsigned long long key, size_t i)
{
    for (;; i = (i + 1) & mask)
    {
        unsigned long long v = table[i];
        if (v == key) return true;
        if (v == kEmptyPair) return false;
    }
}

__device__ __forceinline__ void pair_insert(unsigned long long* table, size_t mask, unsigned long long key)
{
    for (size_t i = pair_slot(key, mask);; i = (i + 1) & mask)
    {
        unsigned long long prev = atomicCAS(&table[i], kEmptyPair, key);
        if (prev == kEmptyPair || prev == key) return;
    }
}

} // namespace phyx
