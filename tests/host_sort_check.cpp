// TEST-ONLY pre-GPU check: compiles the pure functions of the product's event-queue sort (collider-rs_b200/csrc/sort_cut.cuh)
// for the host.  The one-shot bucket sort is correct iff bucket order == key order, i.e. every cut is monotone non-decreasing
// along the canonical key order (time, class, hi id, kind, lo id) — checked here on seeded key sets shaped like a bulk step's
// queue (a heavy tie at the smallest time, labels from a sub-range of the id space as on a sharded rank, a second heavy tie at
// a later time).  Also: the time image is order preserving with -0.0 == +0.0, keys and the 16-byte records decode back to
// the events they were built from.  Built and run by tests/test_sort_host.py; never linked into the product library.
#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../collider-rs_b200/csrc/sort_cut.cuh"

using namespace cb200;

static uint64_t sm_state;
static uint64_t splitmix64() {
    uint64_t z = (sm_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double uni(double lo, double hi) { return lo + (hi - lo) * ((splitmix64() >> 11) * (1.0 / 9007199254740992.0)); }

static int failures = 0;
#define CHECK(cond, ...)                       \
    do {                                       \
        if (!(cond)) {                         \
            if (failures < 20) {               \
                printf("FAIL %s: ", #cond);    \
                printf(__VA_ARGS__);           \
                printf("\n");                  \
            }                                  \
            ++failures;                        \
        }                                      \
    } while (0)

struct Ev {
    double t;
    int kind;  // 0 Reiterate, 1 Collide, 2 Separate
    uint32_t a, b;
};
// what k_keys_solitaire / keys_pairs_body build (one device: label = id rank; add_mode off)
static Key128 key_of(const Ev& e) {
    if (e.kind == 0) return Key128{(unsigned long long)e.a, time_key(e.t)};
    return Key128{kPairClass | ((unsigned long long)e.a << 32) | (e.kind == 1 ? kCollideBit : 0u) | e.b, time_key(e.t)};
}
static bool key_lt(const Key128& x, const Key128& y) { return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo); }

static void check_time_key() {
    const double specials[] = {-INFINITY, -1e300, -2.5, -1.0, -5e-324, -0.0, 0.0, 5e-324, 1e-300, 0.1, 1.0, 2.5, 1e49, 1e50, 1e300, INFINITY};
    const int n = sizeof(specials) / sizeof(specials[0]);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double x = specials[i], y = specials[j];
            CHECK((x < y) == (time_key(x) < time_key(y)), "%g %g", x, y);
            CHECK((x == y) == (time_key(x) == time_key(y)), "%g %g", x, y);  // (-0.0 == +0.0: one key)
        }
    for (int it = 0; it < 200000; ++it) {
        const double x = uni(-10, 10) * std::pow(10.0, (double)(splitmix64() % 40) - 20.0), y = uni(-10, 10);
        CHECK((x < y) == (time_key(x) < time_key(y)), "%a %a", x, y);
        const double back = key_time(time_key(x));
        CHECK(back == x, "%a -> %a", x, back);
    }
    CHECK(!std::signbit(key_time(time_key(-0.0))), "the image of -0.0 decodes to +0.0");
}

// One synthetic queue: n events in [T, T + window], `tie0` of them exactly at T, `tie1` at one later time, labels in [id_lo, id_hi].
static void check_cuts(uint32_t n, double T, double window, uint32_t tie0, uint32_t tie1, uint32_t id_lo, uint32_t id_hi, int pairs_only) {
    std::vector<Ev> ev(n);
    const double t1 = T + 0.37 * window;
    for (uint32_t i = 0; i < n; ++i) {
        Ev e;
        e.t = i < tie0 ? T : (i < tie0 + tie1 ? t1 : T + uni(0.0, window));
        e.kind = pairs_only ? 1 + (int)(splitmix64() % 2) : (int)(splitmix64() % 3);
        e.a = id_lo + (uint32_t)(splitmix64() % (uint64_t)(id_hi - id_lo + 1));
        e.b = e.kind ? id_lo + (uint32_t)(splitmix64() % (uint64_t)(id_hi - id_lo + 1)) : 0u;
        if (e.kind && e.b > e.a) std::swap(e.a, e.b);  // pair events carry (hi, lo)
        ev[i] = e;
    }
    std::vector<Key128> keys(n);
    BsRange tr{~0ull, 0ull}, lr{~0ull, 0ull};  // what the key kernels reduce: time keys; first labels
    for (uint32_t i = 0; i < n; ++i) {
        keys[i] = key_of(ev[i]);
        tr.lo = std::min<unsigned long long>(tr.lo, keys[i].hi);
        tr.hi = std::max<unsigned long long>(tr.hi, keys[i].hi);
        lr.lo = std::min<unsigned long long>(lr.lo, ev[i].a);
        lr.hi = std::max<unsigned long long>(lr.hi, ev[i].a);
    }
    std::sort(keys.begin(), keys.end(), key_lt);
    for (int log2_cap : {6, 12, 24}) {
        const int l = bs2_log2_buckets(n, log2_cap);
        CHECK(l >= 6 && l <= log2_cap && (l == log2_cap || (1u << l) * 8u >= n), "log2 %d for n %u cap %d", l, n, log2_cap);
        const uint32_t n_buckets = 2u << l;
        uint32_t prev = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t b = bs2_bucket(keys[i], tr, l, lr);  // (all keys at one time: everything falls into region A)
            CHECK(b < n_buckets, "bucket %u of %u", b, n_buckets);
            CHECK(b >= prev, "two-region cut not monotone at %u: %u after %u (n %u, log2 %d)", i, b, prev, n, l);
            prev = b;
        }
        // the recursive path's single-field cuts: time field over everything, tie field over the keys of one time
        if (tr.lo != tr.hi) {
            prev = 0;
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t b = bs_bucket(keys[i].hi, tr, l, 1);
                CHECK(b < (1u << l) && b >= prev, "time cut at %u: %u after %u", i, b, prev);
                prev = b;
            }
        }
        uint32_t i0 = 0;
        while (i0 < n) {
            uint32_t i1 = i0;
            while (i1 < n && keys[i1].hi == keys[i0].hi) ++i1;
            if (i1 - i0 > 2 && keys[i0].lo != keys[i1 - 1].lo) {
                const BsRange r{keys[i0].lo, keys[i1 - 1].lo};
                prev = 0;
                for (uint32_t i = i0; i < i1; ++i) {
                    const uint32_t b = bs_bucket(keys[i].lo, r, l, 0);
                    CHECK(b < (1u << l) && b >= prev, "tie cut at %u: %u after %u", i, b, prev);
                    prev = b;
                }
            }
            i0 = i1;
        }
    }
    // keys and compact records decode back to the events
    std::sort(ev.begin(), ev.end(), [](const Ev& x, const Ev& y) { return key_lt(key_of(x), key_of(y)); });
    for (uint32_t i = 0; i < n; ++i) {
        const Event16 r = decode_event16(key_of(ev[i]), 0);
        const int kind = (r.a >> 31) ? 0 : ((r.b >> 31) ? 1 : 2);
        CHECK(r.time == ev[i].t + 0.0 && kind == ev[i].kind && (r.a & 0x7fffffffu) == ev[i].a && (r.b & 0x7fffffffu) == ev[i].b, "record %u", i);
    }
}

int main() {
    sm_state = 0xC0111DE5u;
    check_time_key();
    check_cuts(5000, 0.0, 0.1, 0, 0, 0, 99999, 0);            // spread times only
    check_cuts(200000, 0.0, 0.1, 150000, 0, 0, 999999, 0);    // the bench's queue: most events at exactly T
    check_cuts(200000, 12.3, 0.1, 90000, 60000, 0, 999999, 0);  // a second heavy tie at a later time
    check_cuts(120000, 0.7, 0.1, 100000, 0, 6000000, 6999999, 0);  // a sharded rank: labels from 1/8 of the id space
    check_cuts(50000, -3.0, 4.0, 1000, 0, 17, 17 + 40, 1);    // negative times, few distinct labels, pair events only
    check_cuts(3000, 5.0, 0.0, 3000, 0, 0, 4095, 0);          // every key at one time
    check_cuts(1, 0.0, 0.1, 1, 0, 5, 5, 0);
    if (failures) {
        printf("%d failures\n", failures);
        return 1;
    }
    printf("HOST_SORT_OK\n");
    return 0;
}
