//! Drop-in host side for collider-rs on the B200 engine.  SOURCE ONLY — not compiled in this repository's image
//! (no rustc/cargo); the same C ABI is exercised by the C++ twin `include/collider_b200.hpp`
//! (tests/host_collider_check.cpp) and by the ctypes mirror `collider-rs_b200/collider.py` (tests/test_collider_gpu.py).
//!
//! `Collider<P: HbProfile>` keeps the reference's public API (src/core/collider.rs:59-318): `new(cell_width, padding)`,
//! `time`, `next_time`, `set_time`, `next`, `get_hitbox`, `add_hitbox`, `set_hitbox_vel`, `remove_hitbox`, `get_overlaps`,
//! `is_overlapping`, `query_overlaps` — plus the bulk advance/rebuild entry point `bulk_update`.  Profiles stay on the host
//! (`P` is arbitrary user data); the device sees `(id, group, interact mask)`.  A contract violation the reference would
//! `panic!` on comes back as a negative status + message and is re-raised here as a `panic!` with that message.

use std::collections::HashMap;
use std::ffi::{c_void, CStr};
use std::os::raw::{c_char, c_int};

pub type HbId = u64;
pub type HbGroup = u32;

/// core/mod.rs:210-238
pub trait HbProfile: Copy {
    fn id(&self) -> HbId;
    fn group(&self) -> Option<HbGroup> { Some(0) }
    fn interact_groups(&self) -> &'static [HbGroup] { &[0] }
    fn can_interact(&self, other: &Self) -> bool;
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum HbEvent { Collide, Separate }
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum ShapeKind { Circle, Rect }
/// geom/shape/mod.rs:105-111 + core/mod.rs:34-59,119-130, flattened
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct Hitbox { pub kind: ShapeKind, pub pos: (f64, f64), pub dims: (f64, f64), pub vel: HbVel }
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct HbVel { pub value: (f64, f64), pub resize: (f64, f64), pub end_time: f64 }

// ---- the C ABI (include/collider_b200.h) ----
#[repr(C)] struct Dev { _private: [u8; 0] }
#[repr(C)] #[derive(Default)] struct CHitbox { pos_x: f64, pos_y: f64, dim_x: f64, dim_y: f64, vel_x: f64, vel_y: f64, rsz_x: f64, rsz_y: f64,
                                               end_time: f64, kind: u32, _pad: u32 }
#[repr(C)] struct CProfile { id: u64, group: u32, interact_mask: u32 }
#[repr(C)] pub struct VelView { pub vel_x: *const f64, pub vel_y: *const f64, pub rsz_x: *const f64, pub rsz_y: *const f64, pub end_time: *const f64 }
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct Event { pub time: f64, pub a: u64, pub b: u64, pub kind: u32, pub _pad: u32 }
#[repr(C)] #[derive(Default)] pub struct StepResult { pub next_time: f64, pub next_event: Event, pub n_hitboxes: u64, pub n_entries: u64, pub n_cells: u64,
    pub n_collide_solves: u64, pub n_separate_solves: u64, pub n_events: u64, pub n_solitaire: u64, pub error_flags: u32, pub _pad: u32,
    pub error_slot: u64, pub n_work_items: u64 }

extern "C" {
    fn cb200_collider_new(cell_width: f64, padding: f64, device: c_int, out: *mut *mut Dev) -> c_int;
    fn cb200_collider_free(c: *mut Dev);
    fn cb200_collider_last_error(c: *const Dev) -> *const c_char;
    fn cb200_collider_set_can_interact(c: *mut Dev, f: Option<extern "C" fn(*mut c_void, u64, u64) -> c_int>, user: *mut c_void) -> c_int;
    fn cb200_collider_time(c: *const Dev) -> f64;
    fn cb200_collider_next_time(c: *mut Dev, out: *mut f64) -> c_int;
    fn cb200_collider_set_time(c: *mut Dev, t: f64) -> c_int;
    fn cb200_collider_next(c: *mut Dev, has: *mut c_int, kind: *mut u32, id_1: *mut u64, id_2: *mut u64) -> c_int;
    fn cb200_collider_get_hitbox(c: *mut Dev, id: u64, out: *mut CHitbox) -> c_int;
    fn cb200_collider_add_hitbox(c: *mut Dev, p: *const CProfile, h: *const CHitbox, out: *mut u64, cap: u64, n: *mut u64) -> c_int;
    fn cb200_collider_set_hitbox_vel(c: *mut Dev, id: u64, vx: f64, vy: f64, rw: f64, rh: f64, end_time: f64) -> c_int;
    fn cb200_collider_remove_hitbox(c: *mut Dev, id: u64, out: *mut u64, cap: u64, n: *mut u64) -> c_int;
    fn cb200_collider_get_overlaps(c: *mut Dev, id: u64, out: *mut u64, cap: u64, n: *mut u64) -> c_int;
    fn cb200_collider_is_overlapping(c: *mut Dev, id_1: u64, id_2: u64, out: *mut c_int) -> c_int;
    fn cb200_collider_query_overlaps(c: *mut Dev, kind: u32, px: f64, py: f64, dx: f64, dy: f64, p: *const CProfile, out: *mut u64, cap: u64,
                                     n: *mut u64) -> c_int;
    fn cb200_collider_bulk_update(c: *mut Dev, vel: *const VelView, end_time: f64, out: *mut StepResult) -> c_int;
}

const GROUP_NONE: u32 = 0xFFFF_FFFF;

fn c_profile<P: HbProfile>(p: &P, groups: &mut GroupMap) -> CProfile {
    // the device takes group indices 0..31; arbitrary HbGroup values are mapped to dense indices in order of first use
    CProfile { id: p.id(), group: p.group().map_or(GROUP_NONE, |g| groups.index(g)),
               interact_mask: p.interact_groups().iter().fold(0u32, |m, &g| m | 1u32 << groups.index(g)) }
}
#[derive(Default)]
struct GroupMap { dense: HashMap<HbGroup, u32> }
impl GroupMap {
    fn index(&mut self, g: HbGroup) -> u32 {
        let next = self.dense.len() as u32;
        let i = *self.dense.entry(g).or_insert(next);
        assert!(i < 32, "the device engine supports at most 32 distinct HbGroup values");
        i
    }
}
impl From<Hitbox> for CHitbox {
    fn from(h: Hitbox) -> CHitbox {
        CHitbox { pos_x: h.pos.0, pos_y: h.pos.1, dim_x: h.dims.0, dim_y: h.dims.1, vel_x: h.vel.value.0, vel_y: h.vel.value.1,
                  rsz_x: h.vel.resize.0, rsz_y: h.vel.resize.1, end_time: h.vel.end_time, kind: (h.kind == ShapeKind::Rect) as u32, _pad: 0 }
    }
}

pub struct Collider<P: HbProfile> {
    dev: *mut Dev,
    profiles: Box<HashMap<HbId, P>>, // boxed: the can_interact trampoline holds a pointer to it
    groups: GroupMap,
}

impl<P: HbProfile> Collider<P> {
    /// collider.rs:59-69 (panics unless cell_width > padding > 0.0, same messages)
    pub fn new(cell_width: f64, padding: f64) -> Collider<P> { Self::on_device(cell_width, padding, 0) }
    pub fn on_device(cell_width: f64, padding: f64, device: i32) -> Collider<P> {
        let mut dev = std::ptr::null_mut();
        let rc = unsafe { cb200_collider_new(cell_width, padding, device as c_int, &mut dev) };
        if rc != 0 { panic!("{}", unsafe { CStr::from_ptr(cb200_collider_last_error(std::ptr::null())) }.to_string_lossy()); }
        let mut c = Collider { dev, profiles: Box::new(HashMap::new()), groups: GroupMap::default() };
        let user = &mut *c.profiles as *mut HashMap<HbId, P> as *mut c_void;
        unsafe { cb200_collider_set_can_interact(dev, Some(Self::can_interact_tramp), user) };
        c
    }
    extern "C" fn can_interact_tramp(user: *mut c_void, a: u64, b: u64) -> c_int {
        let profiles = unsafe { &*(user as *const HashMap<HbId, P>) };
        profiles[&a].can_interact(&profiles[&b]) as c_int
    }
    fn check(&self, rc: c_int) {
        if rc != 0 { panic!("{}", unsafe { CStr::from_ptr(cb200_collider_last_error(self.dev)) }.to_string_lossy()); }
    }
    fn ids(&self, mut call: impl FnMut(*mut u64, u64, *mut u64) -> c_int) -> Vec<HbId> {
        let mut buf = vec![0u64; 64];
        loop {
            let mut n = 0u64;
            self.check(call(buf.as_mut_ptr(), buf.len() as u64, &mut n));
            if n as usize <= buf.len() { buf.truncate(n as usize); return buf; }
            buf.resize(n as usize, 0);
        }
    }
    pub fn time(&self) -> f64 { unsafe { cb200_collider_time(self.dev) } }                                   // :72-74
    pub fn next_time(&self) -> f64 { let mut t = 0.0; self.check(unsafe { cb200_collider_next_time(self.dev, &mut t) }); t } // :85-87
    pub fn set_time(&mut self, time: f64) { self.check(unsafe { cb200_collider_set_time(self.dev, time) }) }  // :97-102
    pub fn next(&mut self) -> Option<(HbEvent, P, P)> {                                                       // :113-124
        let (mut has, mut kind, mut a, mut b) = (0, 0u32, 0u64, 0u64);
        self.check(unsafe { cb200_collider_next(self.dev, &mut has, &mut kind, &mut a, &mut b) });
        if has == 0 { return None; }
        Some((if kind == 1 { HbEvent::Collide } else { HbEvent::Separate }, self.profiles[&a], self.profiles[&b]))
    }
    pub fn get_hitbox(&self, id: HbId) -> Hitbox {                                                            // :200-202
        let mut h = CHitbox::default();
        self.check(unsafe { cb200_collider_get_hitbox(self.dev, id, &mut h) });
        Hitbox { kind: if h.kind == 1 { ShapeKind::Rect } else { ShapeKind::Circle }, pos: (h.pos_x, h.pos_y), dims: (h.dim_x, h.dim_y),
                 vel: HbVel { value: (h.vel_x, h.vel_y), resize: (h.rsz_x, h.rsz_y), end_time: h.end_time } }
    }
    pub fn add_hitbox(&mut self, profile: P, hitbox: Hitbox) -> Vec<P> {                                      // :214-222
        let p = c_profile(&profile, &mut self.groups);
        let h: CHitbox = hitbox.into();
        self.profiles.insert(profile.id(), profile);
        let dev = self.dev;
        let mut first = true;
        let id = profile.id();
        let ids = self.ids(|out, cap, n| unsafe {
            if first { first = false; cb200_collider_add_hitbox(dev, &p, &h, out, cap, n) } else { cb200_collider_get_overlaps(dev, id, out, cap, n) }
        });
        ids.iter().map(|i| self.profiles[i]).collect()
    }
    pub fn set_hitbox_vel(&mut self, id: HbId, vel: HbVel) {                                                  // :225-229
        self.check(unsafe { cb200_collider_set_hitbox_vel(self.dev, id, vel.value.0, vel.value.1, vel.resize.0, vel.resize.1, vel.end_time) })
    }
    pub fn remove_hitbox(&mut self, id: HbId) -> Vec<P> {                                                     // :256-275
        let before = self.get_overlaps(id);
        let mut n = 0u64;
        self.check(unsafe { cb200_collider_remove_hitbox(self.dev, id, std::ptr::null_mut(), 0, &mut n) });
        self.profiles.remove(&id);
        before
    }
    pub fn get_overlaps(&self, id: HbId) -> Vec<P> {                                                          // :279-288
        let dev = self.dev;
        self.ids(|out, cap, n| unsafe { cb200_collider_get_overlaps(dev, id, out, cap, n) }).iter().map(|i| self.profiles[i]).collect()
    }
    pub fn get_profile(&self, id: HbId) -> Option<&P> { self.profiles.get(&id) }                              // :291-296
    pub fn is_overlapping(&self, id_1: HbId, id_2: HbId) -> bool {                                            // :300-305
        let mut r = 0;
        self.check(unsafe { cb200_collider_is_overlapping(self.dev, id_1, id_2, &mut r) });
        r != 0
    }
    pub fn query_overlaps(&mut self, kind: ShapeKind, pos: (f64, f64), dims: (f64, f64), profile: &P) -> Vec<P> { // :309-318
        let p = c_profile(profile, &mut self.groups);
        let dev = self.dev;
        self.ids(|out, cap, n| unsafe { cb200_collider_query_overlaps(dev, (kind == ShapeKind::Rect) as u32, pos.0, pos.1, dims.0, dims.1, &p, out, cap, n) })
            .iter().map(|i| self.profiles[i]).collect()
    }
    /// Bulk advance/rebuild: equivalent to `for id in ascending order { self.set_hitbox_vel(id, vel(id)) }` at the current
    /// time, executed in one device pass.  Slices are in ascending-id order; `None` keeps a component.
    pub fn bulk_update(&mut self, vel_x: Option<&[f64]>, vel_y: Option<&[f64]>, end_time: f64) -> StepResult {
        let ptr = |s: Option<&[f64]>| s.map_or(std::ptr::null(), |v| v.as_ptr());
        let view = VelView { vel_x: ptr(vel_x), vel_y: ptr(vel_y), rsz_x: std::ptr::null(), rsz_y: std::ptr::null(), end_time: std::ptr::null() };
        let mut res = StepResult::default();
        self.check(unsafe { cb200_collider_bulk_update(self.dev, &view, end_time, &mut res) });
        res
    }
}

impl<P: HbProfile> Drop for Collider<P> {
    fn drop(&mut self) { unsafe { cb200_collider_free(self.dev) } }
}
