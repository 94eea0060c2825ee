// The reference's own tests (src/tests.rs:67-92 test_smoke, :263-282 test_separate_initial_overlap, and the lib.rs:55-85
// doctest) transcribed onto the C++ `collider::Collider<P>` wrapper (include/collider_b200.hpp).  Built with g++ against
// libcollider_b200.so and run on the GPU box by tests/test_collider_gpu.py.
#include <cstdio>
#include <cstdlib>

#include "../include/collider_b200.hpp"

using namespace collider;

struct TestHbProfile {  // tests.rs:25-37
    uint64_t id_;
    uint64_t id() const { return id_; }
    std::optional<uint32_t> group() const { return 0u; }
    std::vector<uint32_t> interact_groups() const { return {0u}; }
    bool can_interact(const TestHbProfile&) const { return true; }
};

#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            std::exit(1);                                                    \
        }                                                                    \
    } while (0)

static void advance(Collider<TestHbProfile>& c, double time) {  // tests.rs:45-52
    while (c.time() < time) {
        REQUIRE(!c.next());
        c.set_time(std::fmin(c.next_time(), time));
    }
    REQUIRE(c.time() == time);
}
static void advance_to_event(Collider<TestHbProfile>& c, double time) {  // tests.rs:39-43
    advance(c, time);
    REQUIRE(c.next_time() == c.time());
}
static bool is_event(const std::optional<std::tuple<HbEvent, TestHbProfile, TestHbProfile>>& e, HbEvent kind, uint64_t a, uint64_t b) {
    return e && std::get<0>(*e) == kind && std::get<1>(*e).id() == a && std::get<2>(*e).id() == b;
}

int main() {
    {  // test_smoke
        Collider<TestHbProfile> c(4.0, 0.25);
        REQUIRE(c.add_hitbox({0}, Hitbox::moving(PlacedShape::square(2.0, -10.0, 0.0), 1.0, 0.0)).empty());
        REQUIRE(c.add_hitbox({1}, Hitbox::moving(PlacedShape::circle(2.0, 10.0, 0.0), -1.0, 0.0)).empty());
        advance_to_event(c, 9.0);
        REQUIRE(is_event(c.next(), HbEvent::Collide, 0, 1));
        advance_to_event(c, 11.125);
        REQUIRE(is_event(c.next(), HbEvent::Separate, 0, 1));
        advance(c, 23.0);
    }
    {  // test_separate_initial_overlap
        Collider<TestHbProfile> c(4.0, 0.25);
        REQUIRE(c.add_hitbox({0}, Hitbox::moving(PlacedShape::square(1.0, 0.0, 0.0), 0.0, 1.0)).empty());
        auto ov = c.add_hitbox({1}, Hitbox::still(PlacedShape::square(1.0, 0.0, 0.0)));
        REQUIRE(ov.size() == 1 && ov[0].id() == 0);
        advance_to_event(c, 1.25);
        REQUIRE(is_event(c.next(), HbEvent::Separate, 0, 1));
        advance(c, 1.5);
    }
    {  // lib.rs doctest: velocity halving on Collide -> Collide at 9, Separate at 13.01
        Collider<TestHbProfile> c(4.0, 0.01);
        c.add_hitbox({0}, Hitbox::moving(PlacedShape::square(2.0, -10.0, 0.0), 1.0, 0.0));
        c.add_hitbox({1}, Hitbox::moving(PlacedShape::square(2.0, 10.0, 0.0), -1.0, 0.0));
        int n = 0;
        while (c.time() < 20.0) {
            c.set_time(std::fmin(c.next_time(), 20.0));
            if (auto ev = c.next()) {
                ++n;
                if (std::get<0>(*ev) == HbEvent::Collide) {
                    REQUIRE(c.time() == 9.0);
                    for (uint64_t id : {std::get<1>(*ev).id(), std::get<2>(*ev).id()}) {
                        HbVel v = c.get_hitbox(id).vel;
                        v.value.x *= 0.5;
                        v.value.y *= 0.5;
                        c.set_hitbox_vel(id, v);
                    }
                } else {
                    REQUIRE(c.time() == 13.01);
                }
            }
        }
        REQUIRE(n == 2);
    }
    {  // contract: Collider::new asserts (collider.rs:60-61), unknown id (collider.rs:235)
        bool threw = false;
        try {
            Collider<TestHbProfile> bad(0.25, 0.25);
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()).find("cell_width > padding") != std::string::npos;
        }
        REQUIRE(threw);
        Collider<TestHbProfile> c(4.0, 0.25);
        threw = false;
        try {
            c.get_hitbox(7);
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()).find("not found") != std::string::npos;
        }
        REQUIRE(threw);
    }
    std::printf("host_collider_check ok\n");
    return 0;
}
