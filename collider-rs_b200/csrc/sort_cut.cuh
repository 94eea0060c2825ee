// The pure functions of the event-queue sort (sort.cuh): the order-preserving image of an event time, the 128-bit order
// key, the bucket cuts of the one-shot sort and the compact event record.  No device intrinsics: this header is also
// compiled for the host *only by tests/host_sort_check.cpp*, which checks on seeded key sets that every cut is monotone in
// the key order (bucket order == key order is all the sort relies on) and that keys and records decode back to the events
// they were made from.  The product library never calls the host instantiation.
#pragma once
#include <stdint.h>
#include <string.h>

#include "narrowphase.cuh"  // CB_HD

namespace cb200 {

constexpr uint32_t kCollideBit = 0x80000000u;  // in PairEvent.b_kind: 1 = Collide, 0 = Separate
constexpr uint64_t kPairClass = 0x8000000000000000ull;

CB_HD unsigned long long cut_min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
CB_HD int bit_length64(unsigned long long v) {  // 0 for 0, else 1 + index of the highest set bit
#if defined(__CUDA_ARCH__)
    return v ? 64 - __clzll((long long)v) : 0;
#else
    int n = 0;
    while (v) {
        ++n;
        v >>= 1;
    }
    return n;
#endif
}

// Order-preserving u64 image of an f64 event time and its inverse.
CB_HD uint64_t time_key(double t) {
    t = t + 0.0;  // -0.0 -> +0.0 (round-to-nearest; every other value unchanged): the two zeros are one time in the reference's order (N64's Ord is partial_cmp, float.rs:36-42)
#if defined(__CUDA_ARCH__)
    const uint64_t b = (uint64_t)__double_as_longlong(t);
#else
    uint64_t b;
    memcpy(&b, &t, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
CB_HD double key_time(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double t;
    memcpy(&t, &b, 8);
    return t;
#endif
}

struct alignas(16) Key128 {
    unsigned long long lo;  // tie: class | hi << 32 | kind << 31 | lo     (solitaire: slot)
    unsigned long long hi;  // time_key(time)
};

struct BsRange {
    unsigned long long lo, hi;  // min / max of a key field
};

CB_HD unsigned long long bs_field(const Key128& k, int field) { return field ? k.hi : k.lo; }
// Bucket of a field value: monotone non-decreasing in the value (that is all the sort needs).  Times are cut linearly in
// the time itself — the bit pattern of a double is linear only inside one binade, and a step starting at T = 0 spans all of
// them; the tie word is an integer and is cut on its leading bits.
CB_HD uint32_t bs_bucket(unsigned long long t, const BsRange& r, int log2_buckets, int field) {
    if (field == 1) {
        const double lo = key_time(r.lo), hi = key_time(r.hi), x = key_time(t);
        const double scaled = (x - lo) * ((double)(1u << log2_buckets) / (hi - lo));  // hi > lo here; finite (times < 1e50)
        const uint32_t last = (1u << log2_buckets) - 1u;
        return scaled >= (double)last ? last : (uint32_t)scaled;
    }
    const unsigned long long span = r.hi - r.lo;
    const int bits = bit_length64(span);  // span < 2^bits
    const int shift = bits > log2_buckets ? bits - log2_buckets : 0;
    return (uint32_t)((t - r.lo) >> shift);
}

// One-shot variant: two bucket regions, see sort.cuh.
CB_HD uint32_t bs2_bucket(const Key128& k, const BsRange& r, int log2_buckets, const BsRange& lab) {
    const uint32_t nb = 1u << log2_buckets;
    if (k.hi == r.lo) {
        const uint32_t half = nb >> 1;
        const bool pair = (k.lo & kPairClass) != 0;
        const unsigned long long label = pair ? ((k.lo >> 32) & 0x7fffffffull) : cut_min(k.lo, 0xffffffffull);
        const unsigned long long bound = lab.hi - lab.lo + 1ull;                      // (lab.lo <= label <= lab.hi)
        const unsigned long long rel = cut_min(label - cut_min(label, lab.lo), bound - 1ull);
        return (pair ? half : 0u) + (uint32_t)((rel * half) / bound);
    }
    return nb + bs_bucket(k.hi, r, log2_buckets, 1);  // (r.hi > r.lo here)
}

CB_HD int bs2_log2_buckets(uint32_t nt, int log2_cap) {
    int l = 6;
    while ((1u << l) * 8u < nt && l < log2_cap) ++l;  // ~8 keys per bucket (a warp sorts up to kBsMaxBucket)
    return l;
}

// The compact form (cb200_event16): labels instead of ids, the kind in the two spare top bits.
struct Event16 {
    double time;
    uint32_t a, b;
};
CB_HD Event16 decode_event16(const Key128& k, int add_mode) {
    Event16 e;
    e.time = key_time(k.hi);
    if (k.lo & kPairClass) {
        const uint32_t hi = (uint32_t)((k.lo >> 32) & 0x7fffffffu);
        const uint32_t lo = add_mode ? (uint32_t)((k.lo & 0xffffffffu) >> 1) : (uint32_t)(k.lo & 0x7fffffffu);
        const bool collide = (add_mode ? (k.lo & 1ull) : (k.lo & kCollideBit)) != 0;
        e.a = hi;
        e.b = lo | (collide ? 0x80000000u : 0u);
    } else {
        e.a = (uint32_t)k.lo | 0x80000000u;  // Reiterate
        e.b = 0u;
    }
    return e;
}

}  // namespace cb200
