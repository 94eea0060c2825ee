// collider_b200.hpp — header-only C++ mirror of the reference's `Collider<P: HbProfile>` (src/core/collider.rs:30-318) over
// the C ABI of collider_b200.h.  It is what the Rust wrapper of INTEGRATION.md does, in the host language this image has:
// profiles (arbitrary user data) stay on the host, every method forwards one call to the device library, and a failed
// reference `assert!` comes back as a thrown std::runtime_error carrying the reference's message (Rust: panic!).
//
//   struct MyProfile {                       // HbProfile, core/mod.rs:210-238
//       uint64_t id() const;
//       std::optional<uint32_t> group() const;            // default Some(0)
//       std::vector<uint32_t> interact_groups() const;    // default {0}
//       bool can_interact(const MyProfile&) const;        // default true
//   };
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "collider_b200.h"

namespace collider {

enum class ShapeKind : uint32_t { Circle = CB200_KIND_CIRCLE, Rect = CB200_KIND_RECT };  // geom/shape/mod.rs:27-31
enum class HbEvent : uint32_t { Collide = CB200_EV_COLLIDE, Separate = CB200_EV_SEPARATE };  // collider.rs:500-511

struct Vec2 {
    double x = 0, y = 0;
};
struct HbVel {  // core/mod.rs:34-59
    Vec2 value, resize;
    double end_time = INFINITY;
};
struct PlacedShape {  // geom/shape/mod.rs:105-111
    Vec2 pos;
    ShapeKind kind = ShapeKind::Rect;
    Vec2 dims;
    static PlacedShape circle(double diam, double x, double y) { return PlacedShape{{x, y}, ShapeKind::Circle, {diam, diam}}; }
    static PlacedShape rect(double w, double h, double x, double y) { return PlacedShape{{x, y}, ShapeKind::Rect, {w, h}}; }
    static PlacedShape square(double w, double x, double y) { return rect(w, w, x, y); }
};
struct Hitbox {  // core/mod.rs:119-130
    PlacedShape value;
    HbVel vel;
    static Hitbox still(PlacedShape s) { return Hitbox{s, HbVel{}}; }
    static Hitbox moving(PlacedShape s, double vx, double vy) { return Hitbox{s, HbVel{{vx, vy}, {0, 0}, INFINITY}}; }
};

template <class P>
class Collider {
   public:
    Collider(double cell_width, double padding, int device = 0) : profiles_(std::make_unique<Map>()) {  // Collider::new :59-69
        cb200_collider* c = nullptr;
        if (cb200_collider_new(cell_width, padding, device, &c) != CB200_OK) throw std::runtime_error(cb200_collider_last_error(nullptr));
        dev_.reset(c);
        cb200_collider_set_can_interact(c, &Collider::can_interact_tramp, profiles_.get());
    }

    double time() const { return cb200_collider_time(dev_.get()); }  // :72
    double next_time() {                                            // :85
        double t = 0;
        check(cb200_collider_next_time(dev_.get(), &t));
        return t;
    }
    void set_time(double t) { check(cb200_collider_set_time(dev_.get(), t)); }  // :97

    std::optional<std::tuple<HbEvent, P, P>> next() {  // :113
        int has = 0;
        uint32_t kind = 0;
        uint64_t a = 0, b = 0;
        check(cb200_collider_next(dev_.get(), &has, &kind, &a, &b));
        if (!has) return std::nullopt;
        return std::make_tuple(static_cast<HbEvent>(kind), profiles_->at(a), profiles_->at(b));
    }

    Hitbox get_hitbox(uint64_t id) {  // :200
        cb200_hitbox h;
        check(cb200_collider_get_hitbox(dev_.get(), id, &h));
        return Hitbox{PlacedShape{{h.pos_x, h.pos_y}, static_cast<ShapeKind>(h.kind), {h.dim_x, h.dim_y}},
                      HbVel{{h.vel_x, h.vel_y}, {h.rsz_x, h.rsz_y}, h.end_time}};
    }

    std::vector<P> add_hitbox(const P& profile, const Hitbox& hb) {  // :214
        cb200_profile p{profile.id(), profile.group() ? *profile.group() : CB200_GROUP_NONE, 0};
        for (uint32_t g : profile.interact_groups()) p.interact_mask |= 1u << g;
        cb200_hitbox h{hb.value.pos.x, hb.value.pos.y, hb.value.dims.x, hb.value.dims.y, hb.vel.value.x, hb.vel.value.y,
                       hb.vel.resize.x, hb.vel.resize.y, hb.vel.end_time, static_cast<uint32_t>(hb.value.kind), 0};
        profiles_->emplace(profile.id(), profile);
        std::vector<uint64_t> ids(64);
        uint64_t n = 0;
        const int rc = cb200_collider_add_hitbox(dev_.get(), &p, &h, ids.data(), ids.size(), &n);
        if (rc != CB200_OK) {
            profiles_->erase(profile.id());
            check(rc);
        }
        if (n > ids.size()) return get_overlaps(profile.id());
        ids.resize(n);
        return lookup(ids);
    }

    void set_hitbox_vel(uint64_t id, const HbVel& v) {  // :225
        check(cb200_collider_set_hitbox_vel(dev_.get(), id, v.value.x, v.value.y, v.resize.x, v.resize.y, v.end_time));
    }

    std::vector<P> remove_hitbox(uint64_t id) {  // :256
        std::vector<P> out = get_overlaps(id);
        uint64_t n = 0;
        check(cb200_collider_remove_hitbox(dev_.get(), id, nullptr, 0, &n));
        profiles_->erase(id);
        return out;
    }

    std::vector<P> get_overlaps(uint64_t id) {  // :279
        return ids_call([&](uint64_t* out, uint64_t cap, uint64_t* n) { return cb200_collider_get_overlaps(dev_.get(), id, out, cap, n); });
    }
    const P& get_profile(uint64_t id) const { return profiles_->at(id); }  // :291
    bool is_overlapping(uint64_t a, uint64_t b) {                          // :300
        int r = 0;
        check(cb200_collider_is_overlapping(dev_.get(), a, b, &r));
        return r != 0;
    }
    std::vector<P> query_overlaps(const PlacedShape& s, const P& profile) {  // :309
        cb200_profile p{profile.id(), profile.group() ? *profile.group() : CB200_GROUP_NONE, 0};
        for (uint32_t g : profile.interact_groups()) p.interact_mask |= 1u << g;
        // can_interact is evaluated against a registered profile: make the query profile visible for the duration of the call
        const bool had = profiles_->count(profile.id()) != 0;
        if (!had) profiles_->emplace(profile.id(), profile);
        auto out = ids_call([&](uint64_t* o, uint64_t cap, uint64_t* n) {
            return cb200_collider_query_overlaps(dev_.get(), static_cast<uint32_t>(s.kind), s.pos.x, s.pos.y, s.dims.x, s.dims.y, &p, o, cap, n);
        });
        if (!had) profiles_->erase(profile.id());
        return out;
    }

    // The bulk advance/rebuild entry point: for id ascending, set_hitbox_vel(id, vel[id rank]) at the current time, one device pass.
    cb200_step_result bulk_update(const cb200_vel_view* vel, double end_time) {
        cb200_step_result res;
        check(cb200_collider_bulk_update(dev_.get(), vel, end_time, &res));
        return res;
    }
    uint64_t len() const { return cb200_collider_len(dev_.get()); }

   private:
    using Map = std::unordered_map<uint64_t, P>;
    struct Free {
        void operator()(cb200_collider* c) const { cb200_collider_free(c); }
    };
    std::unique_ptr<cb200_collider, Free> dev_;
    std::unique_ptr<Map> profiles_;  // boxed: the library keeps a pointer to it for can_interact

    static int can_interact_tramp(void* user, uint64_t a, uint64_t b) {
        const Map& m = *static_cast<const Map*>(user);
        auto ia = m.find(a), ib = m.find(b);
        return (ia == m.end() || ib == m.end() || ia->second.can_interact(ib->second)) ? 1 : 0;
    }
    void check(int rc) {
        if (rc != CB200_OK) throw std::runtime_error(cb200_collider_last_error(dev_.get()));
    }
    std::vector<P> lookup(const std::vector<uint64_t>& ids) const {
        std::vector<P> out;
        for (uint64_t id : ids) out.push_back(profiles_->at(id));
        return out;
    }
    template <class F>
    std::vector<P> ids_call(F&& f) {
        std::vector<uint64_t> ids(64);
        for (;;) {
            uint64_t n = 0;
            check(f(ids.data(), ids.size(), &n));
            if (n <= ids.size()) {
                ids.resize(n);
                return lookup(ids);
            }
            ids.resize(n);
        }
    }
};

}  // namespace collider
