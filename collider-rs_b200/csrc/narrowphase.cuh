// Narrowphase: circle/rect time-of-collision and time-of-separation, f64, FMA-free.
//
// Device arithmetic for SURVEY.md §8 rows a-10 … a-15.  Each function names the reference code whose
// *results* it must reproduce bit-for-bit (paths under the reference's src/).  The structure is not the
// reference's: shapes are flat register structs, the four-cardinal interval sweep is fully unrolled and
// exits through a single predicate, and the rect/circle corner stage re-uses the quadratic solve with
// the corner folded in as a zero-diameter circle.
//
// Bit-exactness rules (DESIGN.md §"f64 contract"): compile with -fmad=false (no contraction), keep
// every product/sum in the reference's association order, use true f64 `/` and `sqrt` (IEEE
// round-to-nearest on sm_100a), never a reciprocal.
//
// CB_HD marks functions that are also compiled for the host *only by tests/host_narrowphase_check.cpp*
// (a pre-GPU logic check of this very source against the oracle); the product library never calls the
// host instantiation.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb200 {

constexpr double kHighTime = 1e50;  // core/mod.rs:28
constexpr int kCircle = 0;
constexpr int kRect = 1;

// A hitbox at the solve time: placed shape + velocity + remaining duration (DurHitbox,
// core/dur_hitbox/mod.rs:27-70), flattened.
struct DurHb {
    double px, py, w, h, vx, vy, rw, rh, dur;
    int kind;
};

CB_HD double cb_inf() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff0000000000000LL);
#else
    return INFINITY;
#endif
}
CB_HD double cb_nan() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff8000000000000LL);
#else
    return NAN;
#endif
}

// PlacedShape::advance (geom/shape/mod.rs:260-265, 99-101): product rounded, then sum rounded.
CB_HD DurHb advanced(const DurHb& s, double t) {
    DurHb r = s;
    r.px = s.px + s.vx * t;
    r.py = s.py + s.vy * t;
    r.w = s.w + s.rw * t;
    r.h = s.h + s.rh * t;
    return r;
}

// Four-cardinal interval sweep of two moving boxes: results of solvers::rect_rect_time
// (core/dur_hitbox/solvers.rs:60-86) with PlacedBounds::card_overlap (geom/shape/mod.rs:285-304) in
// Card::values order MinusX, MinusY, PlusX, PlusY (geom/card.rs:47-49).
//   card_overlap(a, b, c) = b.edge(c) + a.edge(flip c); with edge(-X) = -left the sum (-l_b) + r_a is
//   the same IEEE operation as r_a - l_b, so each card is one subtraction of two edges.
// The reference leaves the loop at the first card that decides the answer; every early exit returns the
// same sentinel (inf for collide, 0 for separate) and the running window only shrinks, so evaluating
// all four cards and testing once at the end selects the same value.
CB_HD double box_sweep_time(const DurHb& a, const DurHb& b, bool for_collide) {
    const double a_l = a.px - a.w * 0.5, a_r = a.px + a.w * 0.5, a_b = a.py - a.h * 0.5, a_t = a.py + a.h * 0.5;
    const double b_l = b.px - b.w * 0.5, b_r = b.px + b.w * 0.5, b_b = b.py - b.h * 0.5, b_t = b.py + b.h * 0.5;
    const double av_l = a.vx - a.rw * 0.5, av_r = a.vx + a.rw * 0.5, av_b = a.vy - a.rh * 0.5, av_t = a.vy + a.rh * 0.5;
    const double bv_l = b.vx - b.rw * 0.5, bv_r = b.vx + b.rw * 0.5, bv_b = b.vy - b.rh * 0.5, bv_t = b.vy + b.rh * 0.5;
    const double o[4] = {a_r - b_l, a_t - b_b, b_r - a_l, b_t - a_b};
    const double ov[4] = {av_r - bv_l, av_t - bv_b, bv_r - av_l, bv_t - av_b};
    double start = 0.0, end = cb_inf();
    bool decided = false;
#if CB200_BOX_SWEEP_TWO_ARMS  /* A/B: the division written once in each arm */
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool apart = o[c] < 0.0;
        if (apart) {
            if (!for_collide || ov[c] <= 0.0) decided = true;
            else start = fmax(start, -o[c] / ov[c]);
        } else if (ov[c] < 0.0) {
            end = fmin(end, -o[c] / ov[c]);
        }
    }
#else
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        // one quotient per card, shared by both arms (an f64 division is ~30 instructions; a diverged warp would otherwise run
        // the sequence once per arm), and none when neither arm needs it
        const bool apart = o[c] < 0.0;
        const bool to_start = apart && for_collide && !(ov[c] <= 0.0);
        const bool to_end = !apart && ov[c] < 0.0;
        decided = decided || (apart && !to_start);
        if (to_start || to_end) {
            const double q = -o[c] / ov[c];
            start = to_start ? fmax(start, q) : start;
            end = to_end ? fmin(end, q) : end;
        }
    }
#endif
    if (decided || start >= end) return for_collide ? cb_inf() : 0.0;
    return for_collide ? start : end;
}

// util::quad_root_ascending (util.rs:23-32) + the `Some(result) if result >= 0.0` filter of
// solvers::circle_circle_time (solvers.rs:105-108).
CB_HD double first_root_or_inf(double a, double b, double c) {
    const double det = b * b - a * c * 4.0;
    if (det <= 0.0) return cb_inf();
    const double r = (b >= 0.0) ? (c * 2.0) / (-b - sqrt(det)) : (-b + sqrt(det)) / (a * 2.0);
    return (r >= 0.0) ? r : cb_inf();
}

// solvers::circle_circle_time (solvers.rs:88-109) on raw scalars so the rect-corner stage can call it
// with a zero-diameter "circle" (solvers.rs:153-165).
CB_HD double disc_sweep_time(double a_px, double a_py, double a_d, double a_vx, double a_vy, double a_rd,  //
                             double b_px, double b_py, double b_d, double b_vx, double b_vy, double b_rd, bool for_collide) {
    const double sign = for_collide ? 1.0 : -1.0;
    const double net_rad = (a_d + b_d) * 0.5;
    const double dx = a_px - b_px, dy = a_py - b_py;
    const double coeff_c = sign * (net_rad * net_rad - (dx * dx + dy * dy));
    if (coeff_c > 0.0) return 0.0;
    const double net_rad_vel = (a_rd + b_rd) * 0.5;
    const double dvx = a_vx - b_vx, dvy = a_vy - b_vy;
    const double coeff_a = sign * (net_rad_vel * net_rad_vel - (dvx * dvx + dvy * dvy));
    const double coeff_b = sign * 2.0 * (net_rad * net_rad_vel - (dx * dvx + dy * dvy));
    return first_root_or_inf(coeff_a, coeff_b, coeff_c);
}

// solvers::rebased_rect_circle_collide_time (solvers.rs:153-165): if the circle centre lies in a corner
// sector of the box (geom/shape/mod.rs:239-243, 330-353), sweep the circle against that moving corner
// (corner of the velocity bounds = corner velocity, geom/shape/mod.rs:306-318).
CB_HD double corner_stage_time(const DurHb& box, const DurHb& disc) {
    const double l = box.px - box.w * 0.5, r = box.px + box.w * 0.5;
    const double b = box.py - box.h * 0.5, t = box.py + box.h * 0.5;
    const int sx = disc.px < l ? -1 : (disc.px > r ? 1 : 0);
    const int sy = disc.py < b ? -1 : (disc.py > t ? 1 : 0);
    if (sx == 0 || sy == 0) return 0.0;
    const double cx = sx < 0 ? l : r, cy = sy < 0 ? b : t;
    const double cvx = sx < 0 ? box.vx - box.rw * 0.5 : box.vx + box.rw * 0.5;
    const double cvy = sy < 0 ? box.vy - box.rh * 0.5 : box.vy + box.rh * 0.5;
    return disc_sweep_time(cx, cy, 0.0, cvx, cvy, 0.0, disc.px, disc.py, disc.w, disc.vx, disc.vy, disc.rw, true);
}

// solvers::rect_circle_collide_time (solvers.rs:119-131)
CB_HD double box_disc_collide(const DurHb& box, const DurHb& disc, double duration) {
    const double base = box_sweep_time(box, disc, true);
    if (base >= duration) return cb_inf();
    return base + corner_stage_time(advanced(box, base), advanced(disc, base));
}

// solvers::rect_circle_separate_time (solvers.rs:133-151): time-reversed corner stage from base_time.
CB_HD double box_disc_separate(const DurHb& box, const DurHb& disc) {
    const double base = box_sweep_time(box, disc, false);
    if (base == 0.0) return 0.0;
    if (base >= kHighTime) return cb_inf();
    DurHb bx = advanced(box, base), dc = advanced(disc, base);
    bx.vx = -bx.vx; bx.vy = -bx.vy; bx.rw = -bx.rw; bx.rh = -bx.rh;  // DurHbVel::negate, dur_hitbox/mod.rs:47-53
    dc.vx = -dc.vx; dc.vy = -dc.vy; dc.rw = -dc.rw; dc.rh = -dc.rh;
    return fmax(base - corner_stage_time(bx, dc), 0.0);
}

// Field-wise select (keeps both operands in registers; a reference select would force them to local memory).
CB_HD DurHb pick(bool first, const DurHb& x, const DurHb& y) {
    DurHb r;
    r.px = first ? x.px : y.px; r.py = first ? x.py : y.py; r.w = first ? x.w : y.w; r.h = first ? x.h : y.h;
    r.vx = first ? x.vx : y.vx; r.vy = first ? x.vy : y.vy; r.rw = first ? x.rw : y.rw; r.rh = first ? x.rh : y.rh;
    r.dur = first ? x.dur : y.dur;
    r.kind = first ? x.kind : y.kind;
    return r;
}

// solvers::time_unpadded (solvers.rs:46-58)
CB_HD double sweep_unpadded(const DurHb& a, const DurHb& b, bool for_collide, double duration) {
    double result;
    if (a.kind == kRect && b.kind == kRect) {
        result = box_sweep_time(a, b, for_collide);
    } else if (a.kind == kCircle && b.kind == kCircle) {
        result = disc_sweep_time(a.px, a.py, a.w, a.vx, a.vy, a.rw, b.px, b.py, b.w, b.vx, b.vy, b.rw, for_collide);
    } else {
        const bool a_is_box = a.kind == kRect;
        const DurHb box = pick(a_is_box, a, b);
        const DurHb disc = pick(a_is_box, b, a);
        result = for_collide ? box_disc_collide(box, disc, duration) : box_disc_separate(box, disc);
    }
    return (result >= duration) ? cb_inf() : result;
}

// Swept bounds of a hitbox over `duration`: DurHitbox::bounding_box_for (dur_hitbox/mod.rs:92-99) =
// as_rect when still (DurHbVel::is_still :43-45), else PlacedShape::bounding_box of start and end shapes
// (geom/shape/mod.rs:249-258) — note the centre is recomputed as left + w*0.5, and the edges used
// afterwards are centre -/+ w*0.5 again, so the round trip is kept.
struct Box4 {
    double cx, cy, w, h;
    bool ok;  // false where the reference would panic (duration >= 1e50 while moving; negative dims)
};
CB_HD Box4 swept_bounds(const DurHb& s, double duration) {
    Box4 o;
    if (s.vx == 0.0 && s.vy == 0.0 && s.rw == 0.0 && s.rh == 0.0) {
        o.cx = s.px; o.cy = s.py; o.w = s.w; o.h = s.h;
        o.ok = (s.w >= 0.0 && s.h >= 0.0);
        return o;
    }
    const DurHb e = advanced(s, duration);
    const double right = fmax(s.px + s.w * 0.5, e.px + e.w * 0.5);
    const double top = fmax(s.py + s.h * 0.5, e.py + e.h * 0.5);
    const double left = fmin(s.px - s.w * 0.5, e.px - e.w * 0.5);
    const double bottom = fmin(s.py - s.h * 0.5, e.py - e.h * 0.5);
    o.w = right - left;
    o.h = top - bottom;
    o.cx = left + o.w * 0.5;
    o.cy = bottom + o.h * 0.5;
    o.ok = (duration < kHighTime) && (o.w >= 0.0) && (o.h >= 0.0);
    return o;
}

// PlacedShape::overlaps for two boxes = min over the four card overlaps >= 0
// (geom/shape/mod.rs:154-156 -> normals.rs:22-30).
CB_HD bool boxes_touch(const Box4& a, const Box4& b) {
    const double a_l = a.cx - a.w * 0.5, a_r = a.cx + a.w * 0.5, a_b = a.cy - a.h * 0.5, a_t = a.cy + a.h * 0.5;
    const double b_l = b.cx - b.w * 0.5, b_r = b.cx + b.w * 0.5, b_b = b.cy - b.h * 0.5, b_t = b.cy + b.h * 0.5;
    return (a_r - b_l >= 0.0) && (a_t - b_b >= 0.0) && (b_r - a_l >= 0.0) && (b_t - a_b >= 0.0);
}

// solvers::collide_time (solvers.rs:25-34).  NaN where the reference would panic.
CB_HD double collide_time(const DurHb& a, const DurHb& b) {
    const double duration = fmin(a.dur, b.dur);
    const Box4 ba = swept_bounds(a, duration), bb = swept_bounds(b, duration);
    if (!(ba.ok && bb.ok)) return cb_nan();
    if (!boxes_touch(ba, bb)) return cb_inf();
    return sweep_unpadded(a, b, true, duration);
}

// collide_time with the three kind pairs folded into ONE box stage and ONE disc stage, for callers whose warps hold mixed
// kind pairs (k_solve_dense): sweep_unpadded above runs box_sweep_time in its rect/rect arm and again in its rect/circle arm,
// and disc_sweep_time in its circle/circle arm and again in the corner stage — a warp with all three pairs pays for all
// four.  Here every lane runs the same operations on the same operands as collide_time would (results are bit-identical,
// tests/host_narrowphase_check.cpp), but a diverged warp executes each stage once.
CB_HD double collide_time_merged(const DurHb& a, const DurHb& b) {
    const double duration = fmin(a.dur, b.dur);
    const Box4 ba = swept_bounds(a, duration), bb = swept_bounds(b, duration);
    if (!(ba.ok && bb.ok)) return cb_nan();
    if (!boxes_touch(ba, bb)) return cb_inf();
    const bool a_rect = a.kind == kRect, b_rect = b.kind == kRect;
    const bool both_disc = !a_rect && !b_rect, mixed = a_rect != b_rect;
    // box stage: rect/rect on (a, b); rect/circle on (box, disc)
    const bool keep = a_rect || !mixed;
    const DurHb x = pick(keep, a, b), y = pick(keep, b, a);
    double result = cb_inf();
    bool disc_stage = both_disc;
    double base = 0.0;
    // disc stage operands: circle/circle = the two shapes; rect/circle = (moving corner as a zero-diameter disc, disc)
    double p_px = a.px, p_py = a.py, p_d = a.w, p_vx = a.vx, p_vy = a.vy, p_rd = a.rw;
    double q_px = b.px, q_py = b.py, q_d = b.w, q_vx = b.vx, q_vy = b.vy, q_rd = b.rw;
    if (!both_disc) {
        base = box_sweep_time(x, y, true);
        result = base;  // rect/rect
        if (mixed) {
            // box_disc_collide + corner_stage_time
            result = cb_inf();
            if (!(base >= duration)) {
                const DurHb bx = advanced(x, base), dc = advanced(y, base);
                const double l = bx.px - bx.w * 0.5, r = bx.px + bx.w * 0.5;
                const double bo = bx.py - bx.h * 0.5, t = bx.py + bx.h * 0.5;
                const int sx = dc.px < l ? -1 : (dc.px > r ? 1 : 0);
                const int sy = dc.py < bo ? -1 : (dc.py > t ? 1 : 0);
                if (sx == 0 || sy == 0) {
                    result = base + 0.0;
                } else {
                    disc_stage = true;
                    p_px = sx < 0 ? l : r; p_py = sy < 0 ? bo : t; p_d = 0.0;
                    p_vx = sx < 0 ? bx.vx - bx.rw * 0.5 : bx.vx + bx.rw * 0.5;
                    p_vy = sy < 0 ? bx.vy - bx.rh * 0.5 : bx.vy + bx.rh * 0.5;
                    p_rd = 0.0;
                    q_px = dc.px; q_py = dc.py; q_d = dc.w; q_vx = dc.vx; q_vy = dc.vy; q_rd = dc.rw;
                }
            }
        }
    }
    if (disc_stage) {
        const double r2 = disc_sweep_time(p_px, p_py, p_d, p_vx, p_vy, p_rd, q_px, q_py, q_d, q_vx, q_vy, q_rd, true);
        result = both_disc ? r2 : base + r2;
    }
    return (result >= duration) ? cb_inf() : result;
}

// The swept bounds of a hitbox over its OWN duration, as the four edges boxes_touch derives, for filtering pairs before
// collide_time (k_solve_dense).  When two hitboxes have the same duration bit for bit these are the bounds collide_time
// computes for the pair (duration = fmin of equals), so `!swept_edges_touch` is exactly its "bounds do not touch -> inf"
// exit.  dur_key = NaN marks edges that must never filter: bounds not `ok` (the reference would panic) or a non-finite
// velocity (the partner is advanced by 0.0 inside collide_time: x + v * 0.0 keeps x for finite v up to the sign of zero,
// which no comparison or later sum observes, but not for an infinite v).
struct SweptEdges {
    double l, r, b, t;
    double dur_key;
};
CB_HD bool cb_finite(double x) { return (x - x) == 0.0; }
CB_HD SweptEdges swept_edges(const DurHb& d) {
    const Box4 bx = swept_bounds(d, d.dur);
    SweptEdges e;
    e.l = bx.cx - bx.w * 0.5; e.r = bx.cx + bx.w * 0.5; e.b = bx.cy - bx.h * 0.5; e.t = bx.cy + bx.h * 0.5;
    const bool plain = bx.ok && cb_finite(d.vx) && cb_finite(d.vy) && cb_finite(d.rw) && cb_finite(d.rh);
    e.dur_key = plain ? d.dur : cb_nan();
    return e;
}
// false: collide_time of the pair is +inf (its swept bounds do not touch); true: it has to be evaluated
CB_HD bool swept_filter_pass(const SweptEdges& a, double b_l, double b_r, double b_b, double b_t, double b_dur_key) {
    const bool touch = (a.r - b_l >= 0.0) && (a.t - b_b >= 0.0) && (b_r - a.l >= 0.0) && (b_t - a.b >= 0.0);
    return touch || !(a.dur_key == b_dur_key);
}

// solvers::separate_time (solvers.rs:36-44): the circle of a (rect, circle) pair goes first, and only the
// first shape is inflated by 2*padding — the result is not symmetric in (a, b) bit-for-bit.
CB_HD double separate_time(const DurHb& a_in, const DurHb& b_in, double padding) {
    const bool swap = (a_in.kind == kRect && b_in.kind == kCircle);
    DurHb a = pick(swap, b_in, a_in);
    const DurHb b = pick(swap, a_in, b_in);
    a.w = a.w + padding * 2.0;
    a.h = a.h + padding * 2.0;
    if (!(a.w >= 0.0 && a.h >= 0.0)) return cb_nan();  // Shape::new, geom/shape/mod.rs:44-47
    return sweep_unpadded(a, b, false, fmin(a.dur, b.dur));
}

}  // namespace cb200
