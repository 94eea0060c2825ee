// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement (C++17) of the geometry / numeric primitives and the narrowphase solvers of
// SergiusIW/collider-rs 0.3.1.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may build, link, load or call anything under oracle/.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Compile with -O2 -ffp-contract=off and without -ffast-math so that every f64 operation is one
// separately rounded IEEE-754 operation, exactly like rustc's x86-64 SSE2 code.
//
// Pinning: tests/test_oracle_kat.py checks this file against every known-answer assertion the
// reference's own tests hold for the path (src/core/dur_hitbox/mod.rs:116-280, src/util.rs:149-156,
// src/index_rect.rs:79-117, src/geom/shape/tests.rs:17-49, src/tests.rs:67-282, src/lib.rs:55-85).
// The one thing those tests cannot pin is the order of *simultaneous* events (hash-set iteration
// order of the un-vendored `fnv` crate): tie order is "parity unpinned" and replaced by the
// canonical ascending-HbId order described in DESIGN.md.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>

namespace oracle {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] inline void panic(const std::string& msg) { throw Panic(msg); }

constexpr double HIGH_TIME = 1e50;  // src/core/mod.rs:28
constexpr double INF = std::numeric_limits<double>::infinity();

// f64::min / f64::max (Rust): if one argument is NaN the other is returned.
inline double fmin_rs(double a, double b) { return std::fmin(a, b); }
inline double fmax_rs(double a, double b) { return std::fmax(a, b); }

// src/float.rs:20-42 — N64 rejects NaN; ordering is the plain partial order.
inline double n64(double v) {
    if (std::isnan(v)) panic("unexpected NaN");
    return v;
}

// src/geom/vec.rs:20-174 (path subset)
struct Vec2 {
    double x = 0.0, y = 0.0;
    double len_sq() const { return x * x + y * y; }  // vec.rs:44-46
};
inline Vec2 v2(double x, double y) { return Vec2{x, y}; }
inline Vec2 operator+(Vec2 a, Vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline Vec2 operator-(Vec2 a, Vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline Vec2 operator-(Vec2 a) { return {-a.x, -a.y}; }
inline Vec2 operator*(Vec2 a, double s) { return {a.x * s, a.y * s}; }
inline double dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }  // vec.rs:117-122
inline bool operator==(Vec2 a, Vec2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(Vec2 a, Vec2 b) { return !(a == b); }

// src/geom/card.rs:20-50
enum class Card { MinusX = 0, MinusY = 1, PlusX = 2, PlusY = 3 };
constexpr Card CARD_VALUES[4] = {Card::MinusX, Card::MinusY, Card::PlusX, Card::PlusY};  // card.rs:47-49
inline Card flip(Card c) {  // card.rs:36-43
    switch (c) {
        case Card::MinusX: return Card::PlusX;
        case Card::PlusX: return Card::MinusX;
        case Card::MinusY: return Card::PlusY;
        default: return Card::MinusY;
    }
}

enum class ShapeKind : uint8_t { Circle = 0, Rect = 1 };  // shape/mod.rs:27-31

// Sector of a point relative to a box — shape/mod.rs:340-353; -1 Less, 0 Equal, +1 Greater.
struct Sector {
    int x, y;
    bool is_corner() const { return x != 0 && y != 0; }
};
inline int interval_sector(double left, double right, double val) {  // shape/mod.rs:330-338
    if (val < left) return -1;
    if (val > right) return 1;
    return 0;
}

// PlacedBounds trait — shape/mod.rs:268-319.  Implemented for PlacedShape (centre = pos,
// dims = shape dims, shape/mod.rs:321-328), for HbVel (core/mod.rs:109-116) and for DurHbVel
// (dur_hitbox/mod.rs:56-63) where centre = velocity and dims = resize velocity.
struct Bounds {
    Vec2 center, dims;
    double bottom() const { return center.y - dims.y * 0.5; }
    double left() const { return center.x - dims.x * 0.5; }
    double top() const { return center.y + dims.y * 0.5; }
    double right() const { return center.x + dims.x * 0.5; }
    double edge(Card c) const {  // shape/mod.rs:285-292
        switch (c) {
            case Card::MinusY: return -bottom();
            case Card::MinusX: return -left();
            case Card::PlusY: return top();
            default: return right();
        }
    }
    double max_edge() const {  // shape/mod.rs:294-300
        double best = 0.0;
        bool first = true;
        for (Card c : CARD_VALUES) {
            double e = n64(std::fabs(edge(c)));
            if (first || e >= best) best = e;  // max_by_key keeps the last maximum
            first = false;
        }
        return best;
    }
    // self.card_overlap(src, card) = src.edge(card) + self.edge(card.flip()) — shape/mod.rs:302-304
    double card_overlap(const Bounds& src, Card c) const { return src.edge(c) + edge(flip(c)); }
    Vec2 corner(Sector s) const {  // shape/mod.rs:306-318
        if (s.x == 0 || s.y == 0) panic("expected corner sector");
        return v2(s.x < 0 ? left() : right(), s.y < 0 ? bottom() : top());
    }
};

// Shape + PlacedShape — shape/mod.rs:37-266 (path subset).
struct PlacedShape {
    Vec2 pos;
    ShapeKind kind = ShapeKind::Rect;
    Vec2 dims;

    Bounds bounds() const { return Bounds{pos, dims}; }
    double min_x() const { return bounds().left(); }
    double min_y() const { return bounds().bottom(); }
    double max_x() const { return bounds().right(); }
    double max_y() const { return bounds().top(); }
};

// Shape::new — shape/mod.rs:44-55
inline void check_shape(ShapeKind kind, Vec2 dims, bool allow_negative) {
    if (!allow_negative && !(dims.x >= 0.0 && dims.y >= 0.0)) panic("dims must be non-negative");
    if (kind == ShapeKind::Circle && !(dims.x == dims.y)) panic("circle width must equal height");
}
inline PlacedShape placed(Vec2 pos, ShapeKind kind, Vec2 dims) {
    check_shape(kind, dims, false);
    return PlacedShape{pos, kind, dims};
}

// PlacedShape::advance — shape/mod.rs:260-265 with Shape::advance 99-101 (negative dims allowed).
inline PlacedShape advance(const PlacedShape& s, Vec2 vel, Vec2 resize_vel, double elapsed) {
    Vec2 dims = s.dims + resize_vel * elapsed;
    check_shape(s.kind, dims, true);
    return PlacedShape{s.pos + vel * elapsed, s.kind, dims};
}

// PlacedShape::as_rect — shape/mod.rs:245-247
inline PlacedShape as_rect(const PlacedShape& s) { return placed(s.pos, ShapeKind::Rect, s.dims); }

// PlacedShape::bounding_box — shape/mod.rs:249-258
inline PlacedShape bounding_box(const PlacedShape& a, const PlacedShape& b) {
    double right = fmax_rs(a.max_x(), b.max_x());
    double top = fmax_rs(a.max_y(), b.max_y());
    double left = fmin_rs(a.min_x(), b.min_x());
    double bottom = fmin_rs(a.min_y(), b.min_y());
    Vec2 dims = v2(right - left, top - bottom);
    check_shape(ShapeKind::Rect, dims, false);
    Vec2 pos = v2(left + dims.x * 0.5, bottom + dims.y * 0.5);
    return PlacedShape{pos, ShapeKind::Rect, dims};
}

// PlacedShape::sector — shape/mod.rs:239-243
inline Sector sector(const PlacedShape& s, Vec2 point) {
    return Sector{interval_sector(s.min_x(), s.max_x(), point.x), interval_sector(s.min_y(), s.max_y(), point.y)};
}

// normals::rect_rect_normal — normals.rs:22-30 — only the length is used on the path
// (PlacedShape::overlaps, shape/mod.rs:154-156).  min_by_key keeps the first minimum.
inline double rect_rect_normal_len(const PlacedShape& dst, const PlacedShape& src) {
    double best = 0.0;
    bool first = true;
    for (Card c : CARD_VALUES) {
        double o = n64(dst.bounds().card_overlap(src.bounds(), c));
        if (first || o < best) best = o;
        first = false;
    }
    return best;
}

// normals::circle_circle_normal length — normals.rs:32-39
inline double circle_circle_normal_len(const PlacedShape& dst, const PlacedShape& src) {
    Vec2 dir = dst.pos - src.pos;
    double dist = std::sqrt(dir.len_sq());
    return (src.dims.x + dst.dims.x) * 0.5 - dist;
}

// normals::rect_circle_normal length — normals.rs:41-51
inline double rect_circle_normal_len(const PlacedShape& dst, const PlacedShape& src) {
    Sector sec = sector(dst, src.pos);
    if (sec.is_corner()) {
        PlacedShape corner = placed(dst.bounds().corner(sec), ShapeKind::Circle, v2(0.0, 0.0));
        return circle_circle_normal_len(corner, src);
    }
    return rect_rect_normal_len(dst, src);
}

// PlacedShape::overlaps — shape/mod.rs:154-165 (normal_from(...).len() >= 0.0)
inline bool overlaps(const PlacedShape& self, const PlacedShape& other) {
    double len;
    if (self.kind == ShapeKind::Rect && other.kind == ShapeKind::Rect) len = rect_rect_normal_len(self, other);
    else if (self.kind == ShapeKind::Rect && other.kind == ShapeKind::Circle) len = rect_circle_normal_len(self, other);
    else if (self.kind == ShapeKind::Circle && other.kind == ShapeKind::Rect) len = rect_circle_normal_len(other, self);
    else len = circle_circle_normal_len(self, other);
    return len >= 0.0;
}

// util::quad_root_ascending — src/util.rs:23-32.  `found` false = None.
inline bool quad_root_ascending(double a, double b, double c, double* root) {
    double determinant = b * b - a * c * 4.0;
    if (determinant <= 0.0) return false;
    if (b >= 0.0) *root = (c * 2.0) / (-b - std::sqrt(determinant));
    else *root = (-b + std::sqrt(determinant)) / (a * 2.0);
    return true;
}

// DurHbVel / DurHitbox — src/core/dur_hitbox/mod.rs:27-108
struct DurHbVel {
    Vec2 value, resize;
    double duration = INF;
    bool is_still() const { return value == Vec2{} && resize == Vec2{}; }  // :43-45
    DurHbVel negate() const { return DurHbVel{-value, -resize, duration}; }  // :47-53
    Bounds bounds() const { return Bounds{value, resize}; }                 // :56-63
};

struct DurHitbox {
    PlacedShape value;
    DurHbVel vel;

    PlacedShape advanced_shape(double time) const {  // :79-86
        if (!(time < HIGH_TIME)) panic("requires time < 1e50");
        return advance(value, vel.value, vel.resize, time);
    }
    PlacedShape bounding_box_for(double duration) const {  // :92-99
        if (vel.is_still()) return as_rect(value);
        PlacedShape end_value = advanced_shape(duration);
        return oracle::bounding_box(value, end_value);
    }
    PlacedShape bounding_box() const { return bounding_box_for(vel.duration); }  // :88-90
};

// ---- src/core/dur_hitbox/solvers.rs ----

// solvers.rs:60-86
inline double rect_rect_time(const DurHitbox& a, const DurHitbox& b, bool for_collide) {
    double overlap_start = 0.0;
    double overlap_end = INF;
    for (Card card : CARD_VALUES) {
        double overlap = a.value.bounds().card_overlap(b.value.bounds(), card);
        double overlap_vel = a.vel.bounds().card_overlap(b.vel.bounds(), card);
        if (overlap < 0.0) {
            if (!for_collide) return 0.0;
            else if (overlap_vel <= 0.0) return INF;
            else overlap_start = fmax_rs(overlap_start, -overlap / overlap_vel);
        } else if (overlap_vel < 0.0) {
            overlap_end = fmin_rs(overlap_end, -overlap / overlap_vel);
        }
        if (overlap_start >= overlap_end) return for_collide ? INF : 0.0;
    }
    return for_collide ? overlap_start : overlap_end;
}

// solvers.rs:88-109
inline double circle_circle_time(const DurHitbox& a, const DurHitbox& b, bool for_collide) {
    double sign = for_collide ? 1.0 : -1.0;
    double net_rad = (a.value.dims.x + b.value.dims.x) * 0.5;
    Vec2 dist = a.value.pos - b.value.pos;
    double coeff_c = sign * (net_rad * net_rad - dist.len_sq());
    if (coeff_c > 0.0) return 0.0;
    double net_rad_vel = (a.vel.resize.x + b.vel.resize.x) * 0.5;
    Vec2 dist_vel = a.vel.value - b.vel.value;
    double coeff_a = sign * (net_rad_vel * net_rad_vel - dist_vel.len_sq());
    double coeff_b = sign * 2.0 * (net_rad * net_rad_vel - dot(dist, dist_vel));
    double result;
    if (quad_root_ascending(coeff_a, coeff_b, coeff_c, &result) && result >= 0.0) return result;
    return INF;
}

// solvers.rs:153-165
inline double rebased_rect_circle_collide_time(const DurHitbox& rect, const DurHitbox& circle) {
    Sector sec = sector(rect.value, circle.value.pos);
    if (sec.is_corner()) {
        DurHitbox corner;
        corner.value = placed(rect.value.bounds().corner(sec), ShapeKind::Circle, v2(0.0, 0.0));
        corner.vel = DurHbVel{};  // DurHbVel::still()
        corner.vel.value = rect.vel.bounds().corner(sec);
        return circle_circle_time(corner, circle, true);
    }
    return 0.0;
}

// solvers.rs:119-131
inline double rect_circle_collide_time(const DurHitbox& rect_in, const DurHitbox& circle_in, double duration) {
    double base_time = rect_rect_time(rect_in, circle_in, true);
    if (base_time >= duration) return INF;
    DurHitbox rect = rect_in;
    rect.value = rect.advanced_shape(base_time);
    DurHitbox circle = circle_in;
    circle.value = circle.advanced_shape(base_time);
    return base_time + rebased_rect_circle_collide_time(rect, circle);
}

// solvers.rs:133-151
inline double rect_circle_separate_time(const DurHitbox& rect_in, const DurHitbox& circle_in) {
    double base_time = rect_rect_time(rect_in, circle_in, false);
    if (base_time == 0.0) return 0.0;
    if (base_time >= HIGH_TIME) return INF;
    DurHitbox rect = rect_in;
    rect.value = rect.advanced_shape(base_time);
    rect.vel = rect.vel.negate();
    DurHitbox circle = circle_in;
    circle.value = circle.advanced_shape(base_time);
    circle.vel = circle.vel.negate();
    return fmax_rs(base_time - rebased_rect_circle_collide_time(rect, circle), 0.0);
}

// solvers.rs:111-117
inline double rect_circle_time(const DurHitbox& rect, const DurHitbox& circle, bool for_collide, double duration) {
    return for_collide ? rect_circle_collide_time(rect, circle, duration) : rect_circle_separate_time(rect, circle);
}

// solvers.rs:46-58
inline double time_unpadded(const DurHitbox& a, const DurHitbox& b, bool for_collide, double duration) {
    double result;
    ShapeKind ka = a.value.kind, kb = b.value.kind;
    if (ka == ShapeKind::Rect && kb == ShapeKind::Rect) result = rect_rect_time(a, b, for_collide);
    else if (ka == ShapeKind::Circle && kb == ShapeKind::Circle) result = circle_circle_time(a, b, for_collide);
    else if (ka == ShapeKind::Rect && kb == ShapeKind::Circle) result = rect_circle_time(a, b, for_collide, duration);
    else result = rect_circle_time(b, a, for_collide, duration);
    return result >= duration ? INF : result;
}

// solvers.rs:25-34
inline double collide_time(const DurHitbox& a, const DurHitbox& b) {
    double duration = fmin_rs(a.vel.duration, b.vel.duration);
    if (overlaps(a.bounding_box_for(duration), b.bounding_box_for(duration)))
        return time_unpadded(a, b, true, duration);
    return INF;
}

// solvers.rs:36-44
inline double separate_time(const DurHitbox& a_in, const DurHitbox& b_in, double padding) {
    const DurHitbox* pa = &a_in;
    const DurHitbox* pb = &b_in;
    if (a_in.value.kind == ShapeKind::Rect && b_in.value.kind == ShapeKind::Circle) {
        pa = &b_in;
        pb = &a_in;
    }
    DurHitbox a = *pa;
    Vec2 dims = a.value.dims + v2(padding, padding) * 2.0;
    check_shape(a.value.kind, dims, false);  // Shape::new
    a.value.dims = dims;
    return time_unpadded(a, *pb, false, fmin_rs(a.vel.duration, pb->vel.duration));
}

}  // namespace oracle
