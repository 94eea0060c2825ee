// Window drain: the pair events of a bulk step that fall inside the step's own window [T, E] processed on the device, the
// way Collider::next / process_event / process_collision (collider.rs:113-197) would process them for a caller that does not
// react to them — and the overlap set (HitboxInfo.overlaps, collider.rs:463) kept on the device from step to step.
//
// Without a reaction (no set_hitbox_vel between two bulk steps) the chain of a pair is independent of every other pair:
//   Collide(a, b) at t   -> overlaps += {a, b}; Separate(a, b) queued at t + separate_time(a@t, b@t, padding)
//   Separate(a, b) at t  -> overlaps -= {a, b}; Collide(a, b) queued at t + collide_time(a@t, b@t)
// so one thread follows one pair: every follow-up event with time <= E is appended to the step's event list (it is an event
// the reference would deliver before E), and the pair's state at E decides whether it joins or leaves the overlap set.  The
// overlap list is then compacted and the hash table, the per-hitbox bits and the pair Bloom bits are rebuilt from it — all
// inside the step's CUDA graph; the host learns the new pair count with the step result.
// Exact as long as no Reiterate event falls inside the window (a Reiterate re-bins its hitbox and creates events with new
// neighbours): the step reports how many did (k_drain_commit), and bench.py / the tests check that it is zero.
#pragma once
#include "single.cuh"

namespace cb200 {

constexpr unsigned long long kOvDead = 0x8000000000000000ull;  // bit 63 of a table key: the pair separated (labels are < 2^31, so live keys never carry it)
constexpr uint32_t kDrainGo = 1u, kDrainOverflow = 2u, kDrainLongChain = 4u;
constexpr int kDrainMaxChain = 16;

// The tracked overlap set between two rebuilds:
//   table     open addressing (kernels.cuh OverlapSet).  A pair that separates keeps its slot with bit 63 set (dead): lookups
//             compare against the live key and so miss it; a pair that collides again revives its slot
//   list      every pair that has been in the set since the last rebuild, once; the solve stage walks it and skips the dead
//             ones by a table lookup (SolveArgs.ov_live_check).  New pairs are appended
//   bits      per-hitbox / per-pair Bloom bits only ever gain bits between rebuilds (a stale bit costs a lookup, nothing else)
// The host rebuilds everything compactly (engine.cu compact_overlaps) when the dead pairs or the table load grow too much.
struct OvDyn {
    uint32_t n_pairs;    // length of the list (read by the solve stage through SolveArgs.n_ov_dev)
    uint32_t n_live;     // pairs overlapping now
    uint32_t n_add;      // this step: pairs that joined
    uint32_t n_removed;  // this step: pairs that left
    uint32_t flags;      // kDrain*
    uint32_t done_blocks;
    unsigned long long n_followups;
};

struct DrainArgs {
    StateCols s;
    const uint32_t* row_of;
    uint32_t n_rows, n_labels;
    const StepDyn* dyn;
    double padding;
    PairEvent* events;
    uint32_t cap_events;
    Counters* ctr;
    unsigned long long* table;  // overlap table (kernels.cuh OverlapSet), 4 keys per bucket
    uint32_t bucket_mask;
    uint32_t* bits;
    uint32_t* pair_bits;
    uint32_t pair_mask;
    uint2* pairs;       // the list the solve stage walks
    uint32_t cap_pairs;        // capacity of the list
    uint32_t cap_table_pairs;  // list length up to which new pairs also get a table slot (the table has 2 slots per such pair)
    OvDyn* ov;
    StepResultDev* result;  // mapped host memory
    // the step is complete (will not be re-run with larger buffers) iff ...
    uint32_t cap_entries, seg_cap, cap_surv, is_tile;
    const uint32_t* work_count;
};

// Will the host accept this step as it is?  (If a buffer overflowed the step is re-run — and drained then.)
__device__ __forceinline__ bool drain_go(const DrainArgs& a) {
    const Counters* c = a.ctr;
    bool go = c->err_flags == 0u && c->n_first_events <= a.cap_events;
    if (a.is_tile) {
        go = go && c->tile_fail == 0u && c->n_surv <= a.cap_surv;
    } else {
        go = go && !c->overflow_entries && c->n_entries <= a.cap_entries;
        for (int l = 0; l < kWorkLists; ++l) go = go && a.work_count[l] <= a.seg_cap;
    }
    return go;
}

// The pair separated: mark its table slot dead.
__device__ __forceinline__ bool overlap_kill(unsigned long long* table, uint32_t bucket_mask, unsigned long long key) {
    uint32_t b = (uint32_t)mix64(key) & bucket_mask;
    for (;;) {
        for (int k = 0; k < 4; ++k) {
            const unsigned long long q = table[4 * (size_t)b + k];
            if (q == key) {
                table[4 * (size_t)b + k] = key | kOvDead;
                return true;
            }
            if (q == kKeyMax) return false;
        }
        b = (b + 1) & bucket_mask;
    }
}
// The pair overlaps now.  Its dead slot is revived (it is in the list already), or it is appended to the list and — while
// the table has room for it (the list length counts the slots ever claimed, so the table stays at most half full and every
// probe sequence ends) — given a slot.  A pair that found no room is in the list only: the host sees the list outgrow the
// table (ov_pairs_now > cap_table_pairs) and rebuilds everything larger before the next step; the compaction keeps listed
// pairs that are absent from the table (k_ov_compact).
__device__ __forceinline__ void overlap_join(const DrainArgs& a, unsigned long long key, uint32_t hi, uint32_t lo) {
    uint32_t b = (uint32_t)mix64(key) & a.bucket_mask;
    bool listed = false;
    for (;;) {
        for (int k = 0; k < 4; ++k) {
            unsigned long long* slot = &a.table[4 * (size_t)b + k];
            const unsigned long long q = *slot;
            if (q == (key | kOvDead)) {
                *slot = key;  // (one thread per pair: no race on this slot)
                return;
            }
            if (q == key) return;
            if (q == kKeyMax) {
                if (!listed) {
                    const uint32_t pos = atomicAdd(&a.ov->n_pairs, 1u);
                    if (pos < a.cap_pairs) a.pairs[pos] = make_uint2(hi, lo);
                    else atomicOr(&a.ov->flags, kDrainOverflow);
                    listed = true;
                    if (pos >= a.cap_table_pairs) return;  // no room in the table: listed only
                }
                const unsigned long long prev = atomicCAS(slot, kKeyMax, key);
                if (prev == kKeyMax) return;
                // another pair took the slot in the meantime: keep probing from the next slot
            }
        }
        b = (b + 1) & a.bucket_mask;
    }
}

// One thread per pair event of the step: follow the pair's chain up to E, then update the set.  The last block to finish
// publishes the counters with the step result.
__global__ void __launch_bounds__(kThreads) k_drain_chains(const DrainArgs a) {
    __shared__ uint32_t s_last;
    const bool go = drain_go(a);
    const uint32_t n_first = go ? (uint32_t)a.ctr->n_first_events : 0u;
    const double E = a.dyn->end_scalar;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_first; i += gridDim.x * blockDim.x) {
        const PairEvent e0 = a.events[i];
        if (!(e0.time <= E)) continue;
        const uint32_t hi = e0.a, lo = e0.b_kind & ~kCollideBit;
        bool collide = (e0.b_kind & kCollideBit) != 0u;
        const bool first_collide = collide;
        if (hi >= a.n_labels || lo >= a.n_labels) continue;  // (labels index row_of)
        const uint32_t ra = a.row_of[hi], rb = a.row_of[lo];
        if (ra >= a.n_rows || rb >= a.n_rows) continue;
        double t = e0.time;
        bool overlapping = false;
        int it = 0;
        for (; it < kDrainMaxChain; ++it) {
            uint32_t err = 0;
            const DurHb A = row_at_time(a.s, ra, t, &err), B = row_at_time(a.s, rb, t, &err);
            double delay;
            if (collide) {
                overlapping = true;  // process_collision (collider.rs:177-197)
                delay = separate_time(A, B, a.padding);
            } else {
                overlapping = false;  // Separate (collider.rs:139-158)
                delay = collide_time(A, B);
            }
            if (delay != delay) err |= kFNan;
            if (err) report_error(a.ctr, err, hi);
            collide = !collide;
            const double next = t + delay;
            if (!(next <= E) || !(next < kHighTime)) break;
            // a follow-up event the reference would deliver inside the window
            const unsigned long long pos = atomicAdd(&a.ctr->n_pair_events, 1ull);
            if (pos < a.cap_events) a.events[pos] = PairEvent{next, hi, lo | (collide ? kCollideBit : 0u)};
            else atomicOr(&a.ov->flags, kDrainOverflow);
            atomicAdd(&a.ov->n_followups, 1ull);
            t = next;
        }
        if (it == kDrainMaxChain) atomicOr(&a.ov->flags, kDrainLongChain);
        const unsigned long long key = ((unsigned long long)hi << 32) | lo;
        if (first_collide && overlapping) {
            atomicOr(a.bits + (hi >> 5), 1u << (hi & 31u));
            atomicOr(a.bits + (lo >> 5), 1u << (lo & 31u));
            const uint32_t pb = pair_bit_index(mix64(key));
            atomicOr(a.pair_bits + ((pb >> 5) & a.pair_mask), 1u << (pb & 31u));
            overlap_join(a, key, hi, lo);
            atomicAdd(&a.ov->n_add, 1u);
        } else if (!first_collide && !overlapping) {
            if (overlap_kill(a.table, a.bucket_mask, key)) atomicAdd(&a.ov->n_removed, 1u);
        }
    }
    // ---- last block: commit ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&a.ov->done_blocks, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    volatile OvDyn* ov = a.ov;
    const uint32_t n_add = ov->n_add, n_removed = ov->n_removed;
    const uint32_t live = ov->n_live + n_add - n_removed;
    const uint32_t flags = ov->flags | (go ? kDrainGo : 0u);
    if (go) {
        volatile Counters* c = a.ctr;
        a.result->ctr.n_pair_events = c->n_pair_events;  // (the follow-ups are part of the step's event list)
        a.result->ctr.err_flags = c->err_flags;
        a.result->ctr.err_slot = c->err_slot;
    }
    a.result->ov_pairs_now = ov->n_pairs;
    a.result->ov_live_now = live;
    a.result->drain_flags = flags;
    a.result->n_followups = ov->n_followups;
    a.result->n_removed = n_removed;
    a.result->n_added = n_add;
    __threadfence_system();
    // ready for the next step
    ov->n_live = live;
    ov->n_add = 0u;
    ov->n_removed = 0u;
    ov->n_followups = 0ull;
    ov->flags = 0u;
    ov->done_blocks = 0u;
}

// Host-triggered compaction (engine.cu compact_overlaps): the pairs of the list that are not marked dead in the table — live,
// or listed without a slot (overlap_join) — in list order.
__global__ void __launch_bounds__(kThreads) k_ov_compact(const uint2* __restrict__ pairs, uint32_t n_pairs, const unsigned long long* __restrict__ table,
                                                         uint32_t bucket_mask, uint2* out, uint32_t* n_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const uint2 pr = pairs[i];
        const unsigned long long key = ((unsigned long long)pr.x << 32) | pr.y;
        bool dead = false;
        uint32_t b = (uint32_t)mix64(key) & bucket_mask;
        for (bool done = false; !done;) {
            for (int k = 0; k < 4 && !done; ++k) {
                const unsigned long long q = table[4 * (size_t)b + k];
                if (q == (key | kOvDead)) {
                    dead = true;
                    done = true;
                } else if (q == key || q == kKeyMax) {
                    done = true;
                }
            }
            b = (b + 1) & bucket_mask;
        }
        if (dead) continue;
        out[warp_agg_claim32(n_out)] = pr;
    }
}

}  // namespace cb200
