// TEST-ONLY pre-GPU check: compiles the product's device arithmetic (collider-rs_b200/csrc/narrowphase.cuh)
// for the host and compares it bit-for-bit with the oracle on seeded random pairs, including
// degenerate ones (touching edges on integer coordinates, still shapes, shrinking shapes, short
// durations).  Built and run by tests/test_narrowphase_host.py; never linked into the product library.
#include <cinttypes>
#include <cstdio>
#include <cstring>

#include "../collider-rs_b200/csrc/narrowphase.cuh"
#include "../oracle/oracle_geom.hpp"

static uint64_t sm_state;
static uint64_t splitmix64() {
    uint64_t z = (sm_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static double uni(double lo, double hi) { return lo + (hi - lo) * ((splitmix64() >> 11) * (1.0 / 9007199254740992.0)); }
static int pick(int n) { return (int)(splitmix64() % (uint64_t)n); }

static cb200::DurHb random_hb(int mode) {
    cb200::DurHb h;
    h.kind = pick(2);
    bool lattice = (mode & 1) != 0;  // integer/half-integer coordinates: exact touching cases
    if (lattice) {
        h.px = pick(13) - 6;
        h.py = pick(13) - 6;
        h.w = 1 + pick(4);
        h.h = h.kind == cb200::kCircle ? h.w : 1 + pick(4);
        h.vx = (pick(9) - 4) * 0.5;
        h.vy = (pick(9) - 4) * 0.5;
    } else {
        h.px = uni(-6, 6);
        h.py = uni(-6, 6);
        h.w = uni(0.05, 4);
        h.h = h.kind == cb200::kCircle ? h.w : uni(0.05, 4);
        h.vx = uni(-3, 3);
        h.vy = uni(-3, 3);
    }
    int r = pick(4);
    if (r == 0) {
        h.rw = uni(-0.2, 0.5);
        h.rh = h.kind == cb200::kCircle ? h.rw : uni(-0.2, 0.5);
    } else {
        h.rw = h.rh = 0.0;
    }
    if (pick(5) == 0) h.vx = h.vy = 0.0;
    if (pick(7) == 0) h.rw = h.rh = 0.0, h.vx = h.vy = 0.0;  // still
    int d = pick(4);
    h.dur = d == 0 ? uni(0.0, 0.3) : (d == 1 ? uni(0.3, 5) : (d == 2 ? 100.0 : (double)(1 + pick(6))));
    // keep dims non-negative over the duration (the reference bounds end_time by time_until_too_small)
    if (h.rw < 0 && h.w + h.rw * h.dur < 0.01) h.rw = h.rh = (h.kind == cb200::kCircle ? 0.0 : h.rh), h.rw = 0.0;
    if (h.rh < 0 && h.h + h.rh * h.dur < 0.01) h.rh = 0.0, h.rw = (h.kind == cb200::kCircle ? 0.0 : h.rw);
    return h;
}

static oracle::DurHitbox to_oracle(const cb200::DurHb& h) {
    oracle::DurHitbox d;
    d.value = oracle::PlacedShape{oracle::v2(h.px, h.py), (oracle::ShapeKind)h.kind, oracle::v2(h.w, h.h)};
    d.vel = oracle::DurHbVel{oracle::v2(h.vx, h.vy), oracle::v2(h.rw, h.rh), h.dur};
    return d;
}

static bool same_bits(double a, double b) {
    if (a != a && b != b) return true;
    uint64_t x, y;
    memcpy(&x, &a, 8);
    memcpy(&y, &b, 8);
    return x == y || (a == 0.0 && b == 0.0);
}

int main(int argc, char** argv) {
    uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 2000000ull;
    sm_state = argc > 2 ? strtoull(argv[2], nullptr, 0) : 0xC0111DE5ull;
    uint64_t bad = 0, finite_c = 0, finite_s = 0, zero_c = 0, filtered = 0, same_dur = 0;
    for (uint64_t i = 0; i < n; ++i) {
        int mode = pick(2);
        cb200::DurHb a = random_hb(mode), b = random_hb(mode);
        double padding = pick(2) ? 0.01 : 0.25;
        double oc, os;
        try {
            oc = oracle::collide_time(to_oracle(a), to_oracle(b));
        } catch (const std::exception&) {
            oc = NAN;
        }
        try {
            os = oracle::separate_time(to_oracle(a), to_oracle(b), padding);
        } catch (const std::exception&) {
            os = NAN;
        }
        double dc = cb200::collide_time(a, b), ds = cb200::separate_time(a, b, padding);
        {
            // the dense kernel's solve (collide_time_merged) and its filter (swept_edges of each record over its own duration):
            // as the solve stage calls it, the partner is advanced by 0.0 and the edges come from the records as stored
            cb200::DurHb b2 = b;
            if (pick(3) != 0) b2.dur = a.dur;  // equal durations: the case the filter decides
            if (pick(50) == 0) b2.vx = INFINITY;
            if (pick(50) == 0) b2.px = -0.0, b2.vx = 1.0;
            const cb200::DurHb B = cb200::advanced(b2, 0.0);
            const double want = cb200::collide_time(a, B), merged = cb200::collide_time_merged(a, B);
            const cb200::SweptEdges ea = cb200::swept_edges(a), eb = cb200::swept_edges(b2);
            const bool pass = cb200::swept_filter_pass(ea, eb.l, eb.r, eb.b, eb.t, eb.dur_key);
            const bool pass_rev = cb200::swept_filter_pass(eb, ea.l, ea.r, ea.b, ea.t, ea.dur_key);
            same_dur += a.dur == b2.dur;
            filtered += !pass;
            const bool is_inf = want == INFINITY;
            if (!same_bits(want, merged) || (!pass && !is_inf) || pass != pass_rev) {
                if (bad < 10) printf("MISMATCH #%" PRIu64 ": collide %.17g merged %.17g filter pass %d / %d | kinds %d %d\n", i, want, merged, (int)pass, (int)pass_rev, a.kind, b2.kind);
                ++bad;
            }
            // with equal durations and plain records the filter IS collide_time's bounds test: what passes and still yields inf
            // must have touching bounds (checked through the oracle's own early exit being the only other source of inf is not
            // possible here, so only the direction that matters for correctness is asserted above)
        }
        finite_c += std::isfinite(oc);
        finite_s += std::isfinite(os) && os > 0;
        zero_c += oc == 0.0;
        if (!same_bits(oc, dc) || !same_bits(os, ds)) {
            if (bad < 10)
                printf("MISMATCH #%" PRIu64 ": collide oracle %.17g dev %.17g | separate oracle %.17g dev %.17g | kinds %d %d\n", i, oc, dc,
                       os, ds, a.kind, b.kind);
            ++bad;
        }
    }
    printf("filter: %" PRIu64 " pairs with equal durations, %" PRIu64 " rejected by the swept-edge filter\n", same_dur, filtered);
    printf("pairs %" PRIu64 " mismatches %" PRIu64 " finite_collide %" PRIu64 " zero_collide %" PRIu64 " positive_separate %" PRIu64 "\n", n, bad,
           finite_c, zero_c, finite_s);
    return bad ? 1 : 0;
}
