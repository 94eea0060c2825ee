// TEST-ONLY: the dense kernel's pair filter and merged solve (collider-rs_b200/csrc/narrowphase.cuh, host instantiation) on
// explicit pairs handed over by tests/test_dense_filter_host.py — the candidate pairs of real seeded scenes, as the
// records of a bulk step hold them (n x 9 doubles: px, py, w, h, vx, vy, rw, rh, duration; kind).  For every pair:
//   passed[i]  = swept_filter_pass(swept_edges(hi), swept_edges(lo))      (what k_solve_dense's filter decides)
//   merged[i]  = collide_time_merged(hi, advanced(lo, 0.0))                (what its solve computes for a survivor)
// The Python side compares both with the oracle's collide_time of the same pairs.  Never linked into the product library.
#include <cstdint>

#include "../collider-rs_b200/csrc/narrowphase.cuh"

static cb200::DurHb load(const double* r, uint8_t kind) {
    cb200::DurHb d;
    d.px = r[0]; d.py = r[1]; d.w = r[2]; d.h = r[3]; d.vx = r[4]; d.vy = r[5]; d.rw = r[6]; d.rh = r[7]; d.dur = r[8];
    d.kind = kind;
    return d;
}

extern "C" void filter_and_solve(uint64_t n, const double* hi, const uint8_t* kind_hi, const double* lo, const uint8_t* kind_lo, uint8_t* passed,
                                 double* merged, double* plain) {
    for (uint64_t i = 0; i < n; ++i) {
        const cb200::DurHb a = load(hi + 9 * i, kind_hi[i]), b = load(lo + 9 * i, kind_lo[i]);
        const cb200::SweptEdges ea = cb200::swept_edges(a), eb = cb200::swept_edges(b);
        passed[i] = cb200::swept_filter_pass(ea, eb.l, eb.r, eb.b, eb.t, eb.dur_key) ? 1 : 0;
        const cb200::DurHb B = cb200::advanced(b, 0.0);
        merged[i] = cb200::collide_time_merged(a, B);
        plain[i] = cb200::collide_time(a, B);
    }
}
