// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path (see oracle_geom.hpp header).
//
// CPU restatement (C++17) of the orchestrator, sparse grid and event queue of collider-rs 0.3.1:
//   src/index_rect.rs, src/core/mod.rs, src/core/grid.rs, src/core/events.rs, src/core/collider.rs
// (release-build branch of solitaire_event_check, collider.rs:420-448).
//
// Same data-structure semantics as the reference (hash grid keyed (x, y, group); ordered event map
// keyed (time, insertion index); per-hitbox event-key sets and overlap sets).  The single documented
// deviation: wherever the reference iterates a hash set of HbIds (collider.rs:329, :350 via
// grid.rs:121-134) this oracle iterates in ascending HbId, because the reference's iteration order is a
// function of the un-vendored `fnv` crate + std's table layout and cannot be restated ("parity
// unpinned" for the order of simultaneous events; no reference test asserts it).
#pragma once
#include <algorithm>
#include <functional>
#include <map>
#include <optional>
#include <unordered_map>
#include <vector>

#include "oracle_geom.hpp"

namespace oracle {

using HbId = uint64_t;    // core/mod.rs:31
using HbGroup = uint32_t;  // core/mod.rs:190

// ---- src/index_rect.rs:17-72 ----
struct IndexRect {
    int32_t sx, sy, ex, ey;  // start inclusive, end exclusive
    static IndexRect make(int32_t sx, int32_t sy, int32_t ex, int32_t ey) {  // :25-31
        if (!(sx < ex && sy < ey)) panic("IndexRect contains no elements");
        return IndexRect{sx, sy, ex, ey};
    }
    bool contains(int32_t x, int32_t y) const { return x >= sx && x < ex && y >= sy && y < ey; }  // :37-39
    // Iter::next :53-72 — column-major: x outer, y inner.
    template <class F>
    void for_each(F&& f) const {
        int32_t x = sx, y = sy;
        for (;;) {
            f(x, y);
            if (y == ey - 1) {
                if (x == ex - 1) return;
                x += 1;
                y = sy;
            } else {
                y += 1;
            }
        }
    }
};

// Rust `f64 as i32`: saturating, NaN -> 0.
inline int32_t f64_as_i32(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}

// ---- src/core/mod.rs:34-188 ----
struct HbVel {
    Vec2 value, resize;
    double end_time = INF;
    Bounds bounds() const { return Bounds{value, resize}; }  // :109-116
    bool operator==(const HbVel& o) const { return value == o.value && resize == o.resize && end_time == o.end_time; }
};

struct Hitbox {
    PlacedShape value;
    HbVel vel;

    PlacedShape advanced_shape(double time) const {  // :141-144
        if (!(time < HIGH_TIME)) panic("requires time < 1e50");
        return advance(value, vel.value, vel.resize, time);
    }
    void validate(double min_size, double present_time) const {  // :146-162
        if (!(!std::isnan(vel.end_time) && vel.end_time >= present_time)) panic("end time must exceed present time");
        if (value.kind == ShapeKind::Circle && !(vel.resize.x == vel.resize.y))
            panic("circle resize velocity must maintain aspect ratio");
        if (!(value.dims.x >= min_size && value.dims.y >= min_size)) panic("shape width/height must be at least min_size");
    }
    double time_until_too_small(double min_size) const {  // :164-175
        min_size = min_size * 0.9;
        if (!(value.dims.x > min_size && value.dims.y > min_size)) panic("assertion failed: dims > min_size");
        double time = INF;
        if (vel.resize.x < 0.0) time = fmin_rs(time, (min_size - value.dims.x) / vel.resize.x);
        if (vel.resize.y < 0.0) time = fmin_rs(time, (min_size - value.dims.y) / vel.resize.y);
        return time;
    }
    DurHitbox to_dur_hitbox(double time) const {  // :177-187
        if (!(time <= vel.end_time)) panic("assertion failed: time <= self.vel.end_time");
        return DurHitbox{value, DurHbVel{vel.value, vel.resize, vel.end_time - time}};
    }
};

// HbProfile — core/mod.rs:210-238.  `interact_groups` is kept as a list (iterated in the given order,
// grid.rs:122); can_interact is an optional host callback, default true.
struct Profile {
    HbId id = 0;
    bool has_group = true;
    HbGroup group = 0;
    std::vector<HbGroup> interact_groups{0};
};

// TightSet<T> (util.rs:38-96) — set semantics only; ascending iteration (canonical order).
template <class T, class Less = std::less<T>>
struct SortedSet {
    std::vector<T> v;
    bool insert(const T& x) {
        auto it = std::lower_bound(v.begin(), v.end(), x, Less());
        if (it != v.end() && !Less()(x, *it)) return false;
        v.insert(it, x);
        return true;
    }
    bool remove(const T& x) {
        auto it = std::lower_bound(v.begin(), v.end(), x, Less());
        if (it == v.end() || Less()(x, *it)) return false;
        v.erase(it);
        return true;
    }
    bool contains(const T& x) const { return std::binary_search(v.begin(), v.end(), x, Less()); }
    bool empty() const { return v.empty(); }
    void clear() { v.clear(); }
};

// ---- src/core/events.rs ----
constexpr uint64_t PAIR_BASE = 0x8000000000000000ull;  // :26

struct EventKey {  // :28-68
    double time;
    uint64_t index;
};
struct EventKeyOrd {
    bool operator()(const EventKey& a, const EventKey& b) const {
        if (a.time == b.time) return a.index < b.index;
        return n64(a.time) < n64(b.time);
    }
};
struct EventKeyByIndex {  // PartialEq/Hash use the index only (:40-52)
    bool operator()(const EventKey& a, const EventKey& b) const { return a.index < b.index; }
};

enum class EvKind : uint32_t { Reiterate = 0, Collide = 1, Separate = 2 };  // :74-82 (release build)
struct InternalEvent {
    EvKind kind;
    HbId a, b;  // b unused for Reiterate
    bool other_id(HbId id, HbId* out) const {  // :85-99 + util.rs:107-113
        if (kind == EvKind::Reiterate) {
            if (a != id) panic("other_id: id not involved");
            return false;
        }
        if (a == id) { *out = b; return true; }
        if (b == id) { *out = a; return true; }
        panic("other_id: id not involved");
    }
};

using KeySet = SortedSet<EventKey, EventKeyByIndex>;

struct EventManager {  // :102-194
    std::map<EventKey, InternalEvent, EventKeyOrd> events;
    uint64_t next_event_index = 0;

    std::optional<EventKey> new_event_key(double time, bool for_pair) {  // :156-169
        if (time >= HIGH_TIME) return std::nullopt;
        uint64_t index = next_event_index;
        next_event_index += 1;
        if (!(index < PAIR_BASE)) panic("assertion failed: index < PAIR_BASE");
        if (for_pair) index += PAIR_BASE;
        return EventKey{time, index};
    }
    void add_solitaire_event(double time, InternalEvent ev, KeySet& key_set) {  // :115-125
        if (auto key = new_event_key(time, false)) {
            if (!events.emplace(*key, ev).second) panic("duplicate event key");
            if (!key_set.insert(*key)) panic("duplicate key in set");
        }
    }
    void add_pair_event(double time, InternalEvent ev, KeySet& first, KeySet& second) {  // :127-139
        if (auto key = new_event_key(time, true)) {
            if (!events.emplace(*key, ev).second) panic("duplicate event key");
            if (!first.insert(*key)) panic("duplicate key in set");
            if (!second.insert(*key)) panic("duplicate key in set");
        }
    }
    double peek_time() const { return events.empty() ? INF : events.begin()->first.time; }  // :171-173
};

// ---- src/core/grid.rs ----
struct GridKey {  // :30-34
    int32_t x, y;
    HbGroup group;
    bool operator==(const GridKey& o) const { return x == o.x && y == o.y && group == o.group; }
};
struct GridKeyHash {
    size_t operator()(const GridKey& k) const {  // FNV-1a over the 12 key bytes (fnv crate's published algorithm)
        uint64_t h = 0xcbf29ce484222325ull;
        auto mix = [&h](uint32_t w) {
            for (int i = 0; i < 4; ++i) {
                h ^= (w >> (8 * i)) & 0xff;
                h *= 0x100000001b3ull;
            }
        };
        mix((uint32_t)k.x);
        mix((uint32_t)k.y);
        mix(k.group);
        return (size_t)h;
    }
};
struct GridArea {  // :36-46
    IndexRect rect;
    HbGroup group;
    bool contains(const GridKey& k) const { return group == k.group && rect.contains(k.x, k.y); }
};

struct Grid {  // :48-176
    std::unordered_map<GridKey, SortedSet<HbId>, GridKeyHash> map;
    double cell_width;

    explicit Grid(double cw) : cell_width(cw) {}

    double cell_period(const Hitbox& hitbox, bool has_group) const {  // :61-72
        if (has_group) {
            double speed = hitbox.vel.bounds().max_edge();
            if (speed <= 0.0) return INF;
            return cell_width / speed;
        }
        return INF;
    }
    IndexRect index_bounds(const PlacedShape& bounds) const {  // :101-113
        int32_t start_x = f64_as_i32(std::floor(bounds.min_x() / cell_width));
        int32_t start_y = f64_as_i32(std::floor(bounds.min_y() / cell_width));
        // `start + 1` wraps in a release build (no overflow checks); written out to avoid C++ UB.
        int32_t end_x = std::max(f64_as_i32(std::ceil(bounds.max_x() / cell_width)), (int32_t)((uint32_t)start_x + 1u));
        int32_t end_y = std::max(f64_as_i32(std::ceil(bounds.max_y() / cell_width)), (int32_t)((uint32_t)start_y + 1u));
        return IndexRect::make(start_x, start_y, end_x, end_y);
    }
    GridArea grid_area(const DurHitbox& hitbox, HbGroup group) const {  // :94-99
        return GridArea{index_bounds(hitbox.bounding_box()), group};
    }
    // :115-135 — result ascending + unique (canonical order replaces hash-set iteration order).
    std::vector<HbId> overlapping_ids(std::optional<HbId> hitbox_id, IndexRect rect, const std::vector<HbGroup>& groups) const {
        std::vector<HbId> result;
        for (HbGroup group : groups) {
            rect.for_each([&](int32_t x, int32_t y) {
                auto it = map.find(GridKey{x, y, group});
                if (it != map.end())
                    for (HbId other : it->second.v)
                        if (!(hitbox_id && *hitbox_id == other)) result.push_back(other);
            });
        }
        std::sort(result.begin(), result.end());
        result.erase(std::unique(result.begin(), result.end()), result.end());
        return result;
    }
    std::vector<HbId> shape_cellmates(const PlacedShape& shape, const std::vector<HbGroup>& groups) const {  // :74-77
        return overlapping_ids(std::nullopt, index_bounds(shape), groups);
    }
    void update_area(HbId id, std::optional<GridArea> old_area, std::optional<GridArea> new_area) {  // :137-175
        if (old_area) {
            old_area->rect.for_each([&](int32_t x, int32_t y) {
                GridKey key{x, y, old_area->group};
                if (!new_area || !new_area->contains(key)) {
                    auto it = map.find(key);
                    if (it == map.end()) panic("unreachable: grid cell missing");
                    if (!it->second.remove(id)) panic("assertion failed: grid remove");
                    if (it->second.empty()) map.erase(it);
                }
            });
        }
        if (new_area) {
            new_area->rect.for_each([&](int32_t x, int32_t y) {
                GridKey key{x, y, new_area->group};
                if (!old_area || !old_area->contains(key)) {
                    if (!map[key].insert(id)) panic("assertion failed: grid insert");
                }
            });
        }
    }
    // :79-92
    std::optional<std::vector<HbId>> update_hitbox(HbId id, HbGroup group, const DurHitbox* old_hitbox, const DurHitbox* new_hitbox,
                                                   const std::vector<HbGroup>& groups) {
        if (!(new_hitbox || groups.empty())) panic("assertion failed: new_hitbox.is_some() || groups.is_empty()");
        std::optional<GridArea> old_area, new_area;
        if (old_hitbox) old_area = grid_area(*old_hitbox, group);
        if (new_hitbox) new_area = grid_area(*new_hitbox, group);
        update_area(id, old_area, new_area);
        if (new_area) return overlapping_ids(id, new_area->rect, groups);
        return std::nullopt;
    }
};

// ---- src/core/collider.rs ----
enum class HbEvent : uint32_t { Collide = 1, Separate = 2 };  // :500-511

struct HitboxInfo {  // :457-498
    Profile profile;
    Hitbox hitbox;
    double start_time;
    double pub_end_time;
    KeySet event_keys;
    SortedSet<HbId> overlaps;

    DurHitbox hitbox_at_time(double time) const {  // :478-486
        if (!(time >= start_time && time <= hitbox.vel.end_time)) panic("invalid time");
        Hitbox result = hitbox;
        result.value = result.advanced_shape(time - start_time);
        return result.to_dur_hitbox(time);
    }
    Hitbox pub_hitbox_at_time(double time) const {  // :488-497
        if (!(time >= start_time && time <= pub_end_time)) panic("invalid time");
        Hitbox result = hitbox;
        result.vel.end_time = pub_end_time;
        result.value = result.advanced_shape(time - start_time);
        return result;
    }
};

struct UserEvent {
    HbEvent kind;
    HbId id_1, id_2;  // ascending
};

struct BulkStats {
    uint64_t updates = 0;          // internal_update_hitbox calls
    uint64_t collide_solves = 0;   // collide_time evaluations
    uint64_t separate_solves = 0;  // separate_time evaluations
    uint64_t grid_cells = 0;       // cells in the new index rect, summed
};

class Collider {  // :30-36
   public:
    using CanInteract = std::function<bool(const Profile&, const Profile&)>;

    Collider(double cell_width, double padding) : grid_(cell_width), padding_(padding) {  // :59-69
        if (!(cell_width > padding)) panic("requires cell_width > padding");
        if (!(padding > 0.0)) panic("requires padding > 0.0");
    }
    void set_can_interact(CanInteract f) { can_interact_ = std::move(f); }

    double time() const { return time_; }                       // :72-74
    double next_time() const { return events_.peek_time(); }    // :85-87
    void set_time(double time) {                                // :97-102
        if (!(time >= time_)) panic("cannot rewind time");
        if (!(time <= next_time())) panic("time must not exceed next_time()");
        if (!(time < HIGH_TIME)) panic("time must not exceed 1e50");
        time_ = time;
    }

    std::optional<UserEvent> next() {  // :113-124
        for (;;) {
            std::optional<InternalEvent> event = events_next();
            if (!event) return std::nullopt;
            if (auto ue = process_event(*event)) return ue;
        }
    }

    Hitbox get_hitbox(HbId id) const { return info(id).pub_hitbox_at_time(time_); }  // :200-202

    std::vector<HbId> add_hitbox(const Profile& profile, const Hitbox& hitbox) {  // :214-222
        hitbox.validate(padding_, time_);
        HbId id = profile.id;
        if (hitboxes_.count(id)) panic("assertion failed: self.hitboxes.insert(id, info).is_none()");
        HitboxInfo inf{profile, hitbox, time_, hitbox.vel.end_time, {}, {}};
        solitaire_event_check(id, inf, profile.has_group);
        DurHitbox dur = inf.hitbox.to_dur_hitbox(time_);
        return update_hitbox_tracking(id, std::move(inf), nullptr, dur);
    }

    void set_hitbox_vel(HbId id, const HbVel& vel) {  // :225-229
        if (!(info(id).hitbox.vel == vel)) internal_update_hitbox(id, &vel);
    }

    std::vector<HbId> remove_hitbox(HbId id) {  // :256-275
        auto it = hitboxes_.find(id);
        if (it == hitboxes_.end()) panic("hitbox id " + std::to_string(id) + " not found");
        HitboxInfo inf = std::move(it->second);
        hitboxes_.erase(it);
        clear_related_events(id, inf.event_keys);
        if (inf.profile.has_group) {
            DurHitbox old = inf.hitbox.to_dur_hitbox(inf.start_time);
            grid_.update_hitbox(id, inf.profile.group, &old, nullptr, {});
        }
        return clear_overlaps(id, inf);
    }

    std::vector<HbId> get_overlaps(HbId id) const { return info(id).overlaps.v; }  // :279-288
    const Profile& get_profile(HbId id) const { return info(id).profile; }          // :291-296
    bool is_overlapping(HbId id_1, HbId id_2) const {                               // :300-305
        auto it = hitboxes_.find(id_1);
        return it != hitboxes_.end() && it->second.overlaps.contains(id_2);
    }
    std::vector<HbId> query_overlaps(const PlacedShape& shape, const Profile& profile) const {  // :309-318
        std::vector<HbId> out;
        for (HbId id : grid_.shape_cellmates(shape, profile.interact_groups)) {
            const HitboxInfo& inf = hitboxes_.at(id);
            if (!can_interact(inf.profile, profile)) continue;
            if (!overlaps(inf.pub_hitbox_at_time(time_).value, shape)) continue;
            out.push_back(id);
        }
        return out;
    }

    // Bulk step — reference-equivalent definition: at the current time, for every id in
    // ascending order, run what `set_hitbox_vel(id, vel)` runs when the velocity differs
    // (collider.rs:225-250).  `vel_of(id, current_vel)` returns the velocity to install.
    template <class F>
    void bulk_update(F&& vel_of) {
        std::vector<HbId> ids;
        ids.reserve(hitboxes_.size());
        for (auto& kv : hitboxes_) ids.push_back(kv.first);
        std::sort(ids.begin(), ids.end());
        in_bulk_ = true;
        bulk_candidates_.clear();
        try {
            for (HbId id : ids) {
                HbVel cur = hitboxes_.at(id).hitbox.vel;
                cur.end_time = hitboxes_.at(id).pub_end_time;
                HbVel vel = vel_of(id, cur);
                internal_update_hitbox(id, &vel);
            }
        } catch (...) {
            in_bulk_ = false;
            throw;
        }
        in_bulk_ = false;
    }

    // ---- introspection for parity tests (not in the reference API) ----
    struct QueuedEvent {
        double time;
        uint64_t index;
        EvKind kind;
        HbId a, b;
    };
    std::vector<QueuedEvent> dump_events() const {
        std::vector<QueuedEvent> out;
        for (auto& kv : events_.events) out.push_back({kv.first.time, kv.first.index, kv.second.kind, kv.second.a, kv.second.b});
        return out;
    }
    struct CellEntry {
        int32_t x, y;
        HbGroup group;
        HbId id;
    };
    std::vector<CellEntry> dump_grid() const {
        std::vector<CellEntry> out;
        for (auto& kv : grid_.map)
            for (HbId id : kv.second.v) out.push_back({kv.first.x, kv.first.y, kv.first.group, id});
        return out;
    }
    size_t len() const { return hitboxes_.size(); }
    std::vector<HbId> ids() const {
        std::vector<HbId> v;
        for (auto& kv : hitboxes_) v.push_back(kv.first);
        std::sort(v.begin(), v.end());
        return v;
    }
    // Raw internal state (start_time, internal end_time) for state-parity checks.
    const HitboxInfo& info(HbId id) const {
        auto it = hitboxes_.find(id);
        if (it == hitboxes_.end()) panic("hitbox id " + std::to_string(id) + " not found");
        return it->second;
    }
    bool record_candidates = false;
    // (hi, lo) for every collide_time evaluated during the last bulk_update against a lower id.
    const std::vector<std::pair<HbId, HbId>>& bulk_candidates() const { return bulk_candidates_; }
    BulkStats stats;

   private:
    std::unordered_map<HbId, HitboxInfo> hitboxes_;
    double time_ = 0.0;
    Grid grid_;
    double padding_;
    EventManager events_;
    CanInteract can_interact_;
    bool in_bulk_ = false;
    std::vector<std::pair<HbId, HbId>> bulk_candidates_;

    bool can_interact(const Profile& a, const Profile& b) const { return can_interact_ ? can_interact_(a, b) : true; }

    // EventManager::next — events.rs:175-189
    std::optional<InternalEvent> events_next() {
        if (events_.events.empty()) return std::nullopt;
        auto it = events_.events.begin();
        if (!(it->first.time == time_)) return std::nullopt;
        EventKey key = it->first;
        InternalEvent ev = it->second;
        events_.events.erase(it);
        if (!hitboxes_.at(ev.a).event_keys.remove(key)) panic("assertion failed: event key missing");
        if (ev.kind != EvKind::Reiterate)
            if (!hitboxes_.at(ev.b).event_keys.remove(key)) panic("assertion failed: event key missing");
        return ev;
    }

    // EventManager::clear_related_events — events.rs:141-154 (the hitbox `id` is out of the map here)
    void clear_related_events(HbId id, KeySet& key_set) {
        for (const EventKey& key : key_set.v) {
            auto it = events_.events.find(key);
            if (it == events_.events.end()) panic("called `Option::unwrap()` on a `None` value");
            InternalEvent ev = it->second;
            events_.events.erase(it);
            HbId other;
            if (ev.other_id(id, &other))
                if (!hitboxes_.at(other).event_keys.remove(key)) panic("assertion failed: partner key missing");
        }
        key_set.clear();
    }

    static UserEvent new_event(HbEvent kind, HbId id_1, HbId id_2) {  // :513-519
        if (id_1 == id_2) panic("ids must be different");
        if (id_1 > id_2) std::swap(id_1, id_2);
        return UserEvent{kind, id_1, id_2};
    }

    std::optional<UserEvent> process_event(const InternalEvent& event) {  // :126-175
        switch (event.kind) {
            case EvKind::Collide: {
                HitboxInfo& h1 = hitboxes_.at(event.a);
                HitboxInfo& h2 = hitboxes_.at(event.b);
                process_collision(event.a, h1, event.b, h2);
                return new_event(HbEvent::Collide, event.a, event.b);
            }
            case EvKind::Separate: {
                HitboxInfo& h1 = hitboxes_.at(event.a);
                HitboxInfo& h2 = hitboxes_.at(event.b);
                if (!h1.overlaps.remove(event.b)) panic("assertion failed: overlaps.remove");
                if (!h2.overlaps.remove(event.a)) panic("assertion failed: overlaps.remove");
                double delay = collide_time(h1.hitbox_at_time(time_), h2.hitbox_at_time(time_));
                stats.collide_solves++;
                events_.add_pair_event(time_ + delay, InternalEvent{EvKind::Collide, event.a, event.b}, h1.event_keys, h2.event_keys);
                return new_event(HbEvent::Separate, event.a, event.b);
            }
            default:
                internal_update_hitbox(event.a, nullptr);
                return std::nullopt;
        }
    }

    void process_collision(HbId id_1, HitboxInfo& hb_1, HbId id_2, HitboxInfo& hb_2) {  // :177-197
        if (!hb_1.overlaps.insert(id_2)) panic("assertion failed: overlaps.insert");
        if (!hb_2.overlaps.insert(id_1)) panic("assertion failed: overlaps.insert");
        double delay = separate_time(hb_1.hitbox_at_time(time_), hb_2.hitbox_at_time(time_), padding_);
        stats.separate_solves++;
        events_.add_pair_event(time_ + delay, InternalEvent{EvKind::Separate, id_1, id_2}, hb_1.event_keys, hb_2.event_keys);
    }

    void internal_update_hitbox(HbId id, const HbVel* vel) {  // :231-250
        auto it = hitboxes_.find(id);
        if (it == hitboxes_.end()) panic("hitbox id " + std::to_string(id) + " not found");
        HitboxInfo inf = std::move(it->second);
        hitboxes_.erase(it);
        DurHitbox old_hitbox = inf.hitbox.to_dur_hitbox(inf.start_time);
        inf.hitbox = inf.pub_hitbox_at_time(time_);
        if (vel) {
            inf.hitbox.vel = *vel;
            inf.hitbox.validate(padding_, time_);
        }
        inf.start_time = time_;
        bool has_group = inf.profile.has_group;
        clear_related_events(id, inf.event_keys);
        solitaire_event_check(id, inf, has_group);
        DurHitbox new_hitbox = inf.hitbox.to_dur_hitbox(time_);
        stats.updates++;
        std::vector<HbId> result = update_hitbox_tracking(id, std::move(inf), &old_hitbox, new_hitbox);
        if (!result.empty()) panic("assertion failed: result.is_empty()");
    }

    // :320-381
    std::vector<HbId> update_hitbox_tracking(HbId id, HitboxInfo inf, const DurHitbox* old_hitbox, const DurHitbox& new_hitbox) {
        std::vector<HbId> result;
        if (inf.profile.has_group) {
            HbGroup group = inf.profile.group;
            std::vector<HbId> overlaps_clone = inf.overlaps.v;
            for (HbId other_id : overlaps_clone) {
                HitboxInfo& other = hitboxes_.at(other_id);
                double delay = separate_time(new_hitbox, other.hitbox_at_time(time_), padding_);
                stats.separate_solves++;
                events_.add_pair_event(time_ + delay, InternalEvent{EvKind::Separate, id, other_id}, inf.event_keys, other.event_keys);
            }
            std::vector<HbId> test_ids = *grid_.update_hitbox(id, group, old_hitbox, &new_hitbox, inf.profile.interact_groups);
            {
                IndexRect r = grid_.index_bounds(new_hitbox.bounding_box());
                stats.grid_cells += (uint64_t)(r.ex - r.sx) * (uint64_t)(r.ey - r.sy);
            }
            for (HbId other_id : test_ids) {
                if (old_hitbox == nullptr || !inf.overlaps.contains(other_id)) {
                    HitboxInfo& other = hitboxes_.at(other_id);
                    if (can_interact(inf.profile, other.profile)) {
                        double delay = collide_time(new_hitbox, other.hitbox_at_time(time_));
                        stats.collide_solves++;
                        if (in_bulk_ && record_candidates && other_id < id) bulk_candidates_.emplace_back(id, other_id);
                        if (old_hitbox == nullptr && delay == 0.0) {
                            result.push_back(other_id);
                            process_collision(id, inf, other_id, other);
                        } else {
                            events_.add_pair_event(time_ + delay, InternalEvent{EvKind::Collide, id, other_id}, inf.event_keys,
                                                   other.event_keys);
                        }
                    }
                }
            }
        }
        if (!hitboxes_.emplace(id, std::move(inf)).second) panic("assertion failed: self.hitboxes.insert(id, info).is_none()");
        return result;
    }

    std::vector<HbId> clear_overlaps(HbId id, HitboxInfo& inf) {  // :383-393
        std::vector<HbId> out;
        for (HbId other_id : inf.overlaps.v) {
            HitboxInfo& other = hitboxes_.at(other_id);
            if (!other.overlaps.remove(id)) panic("assertion failed: overlaps.remove");
            out.push_back(other_id);
        }
        inf.overlaps.clear();
        return out;
    }

    // release-build branch — :420-448
    void solitaire_event_check(HbId id, HitboxInfo& inf, bool has_group) {
        inf.pub_end_time = inf.hitbox.vel.end_time;
        double r_time = time_ + grid_.cell_period(inf.hitbox, has_group);
        bool r_flag = true;
        double end_time = inf.hitbox.vel.end_time;
        if (end_time < r_time) {
            r_time = end_time;
            r_flag = false;
        }
        end_time = time_ + inf.hitbox.time_until_too_small(padding_);
        if (end_time < r_time) {
            r_time = end_time;
            r_flag = false;
        }
        inf.hitbox.vel.end_time = r_time;
        if (r_flag) events_.add_solitaire_event(r_time, InternalEvent{EvKind::Reiterate, id, 0}, inf.event_keys);
    }
};

}  // namespace oracle
