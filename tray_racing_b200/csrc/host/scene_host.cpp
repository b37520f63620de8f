// scene_host.cpp — meshes (OBJ + deterministic synthetic scenes), the cwbvh_gpu_runner marshalling and
// the camera uniform.  CPU only; see include/tray_host.h for the reference file:line each piece mirrors.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tray_host.h"

struct tray_mesh {
    std::vector<float> tris;             // 9 floats per triangle
    std::vector<uint64_t> offsets;       // object k = triangles [offsets[k], offsets[k+1])
    float eye[3] = { 0, 0, 5 }, look_at[3] = { 0, 0, 0 }, fov = 90.f;
    uint64_t count() const { return tris.size() / 9; }
};

namespace {

// binary32 -> binary16, round to nearest even (what the `half` crate's f16::from_f32 does)
uint16_t float_to_half(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));   // inf / nan
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                      // rounds to >= 65520: inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                   // <= 2^-25: zero
    if (x < 0x38800000u) {                                                                        // subnormal half
        const uint32_t e = x >> 23, m = (x & 0x7fffffu) | 0x800000u;
        const uint32_t shift = 126u - e;                        // 14..24: the half subnormal keeps the top 24 - shift bits of m
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((x >> 23) - 112u) << 10 | ((x >> 13) & 0x3ffu);
    const uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;      // may carry into the exponent: still the right value
    return (uint16_t)(sign | h);
}


struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline V3 norm(V3 a) { double l = std::sqrt(dot(a, a)); return l > 0 ? a * (1.0 / l) : V3{ 0, 1, 0 }; }

struct Rng {   // SplitMix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double range(double a, double b) { return a + (b - a) * uni(); }
    double logrange(double a, double b) { return a * std::pow(b / a, uni()); }
    V3 on_sphere() { double z = range(-1, 1), t = range(0, 6.283185307179586), r = std::sqrt(std::max(0.0, 1 - z * z)); return { r * std::cos(t), r * std::sin(t), z }; }
};

struct Gen {
    tray_mesh* m;
    void tri(V3 a, V3 b, V3 c) {
        float v[9] = { (float)a.x, (float)a.y, (float)a.z, (float)b.x, (float)b.y, (float)b.z, (float)c.x, (float)c.y, (float)c.z };
        m->tris.insert(m->tris.end(), v, v + 9);
    }
    void quad(V3 a, V3 b, V3 c, V3 d) { tri(a, b, c); tri(a, c, d); }
    void begin_object() { m->offsets.push_back(m->count()); }
    // parallelogram patch o + u*s + v*t subdivided n x n
    void patch(V3 o, V3 u, V3 v, int n) {
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
            double s0 = (double)i / n, s1 = (double)(i + 1) / n, t0 = (double)j / n, t1 = (double)(j + 1) / n;
            quad(o + u * s0 + v * t0, o + u * s1 + v * t0, o + u * s1 + v * t1, o + u * s0 + v * t1);
        }
    }
    // oriented box: centre c, half-axes ax, ay, az; n x n quads per face  -> 12 n^2 triangles
    void box(V3 c, V3 ax, V3 ay, V3 az, int n) {
        patch(c - ax - ay - az, ax * 2, ay * 2, n); patch(c - ax - ay + az, ax * 2, ay * 2, n);
        patch(c - ax - ay - az, ay * 2, az * 2, n); patch(c + ax - ay - az, ay * 2, az * 2, n);
        patch(c - ax - ay - az, az * 2, ax * 2, n); patch(c - ax + ay - az, az * 2, ax * 2, n);
    }
    // uv sphere with radial displacement; 2 * nlon * (nlat - 1) triangles
    template <class F> void sphere(V3 c, double r, int nlat, int nlon, F disp) {
        auto P = [&](int i, int j) {
            double th = 3.141592653589793 * i / nlat, ph = 6.283185307179586 * (j % nlon) / nlon;
            V3 d{ std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph) };
            return c + d * (r * (1.0 + disp(d)));
        };
        for (int i = 0; i < nlat; i++) for (int j = 0; j < nlon; j++) {
            V3 a = P(i, j), b = P(i + 1, j), cc = P(i + 1, j + 1), d = P(i, j + 1);
            if (i > 0) tri(a, cc, d);
            if (i < nlat - 1) tri(a, b, cc);
        }
    }
    void cylinder(V3 c, double r, double h, int n) {   // 4n triangles
        for (int j = 0; j < n; j++) {
            double a0 = 6.283185307179586 * j / n, a1 = 6.283185307179586 * (j + 1) / n;
            V3 p0{ c.x + r * std::cos(a0), c.y, c.z + r * std::sin(a0) }, p1{ c.x + r * std::cos(a1), c.y, c.z + r * std::sin(a1) };
            V3 q0 = p0 + V3{ 0, h, 0 }, q1 = p1 + V3{ 0, h, 0 };
            quad(p0, p1, q1, q0); tri(c, p1, p0); tri(c + V3{ 0, h, 0 }, q0, q1);
        }
    }
};

inline double lattice(int64_t ix, int64_t iz, uint64_t seed) {
    uint64_t h = (uint64_t)ix * 0x9E3779B97F4A7C15ull ^ ((uint64_t)iz * 0xC2B2AE3D27D4EB4Full) ^ seed;
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull; h = (h ^ (h >> 27)) * 0x94D049BB133111EBull; h ^= h >> 31;
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
inline double vnoise(double x, double z, uint64_t seed) {
    double fx = std::floor(x), fz = std::floor(z); int64_t ix = (int64_t)fx, iz = (int64_t)fz;
    double tx = x - fx, tz = z - fz; tx = tx * tx * (3 - 2 * tx); tz = tz * tz * (3 - 2 * tz);
    double a = lattice(ix, iz, seed), b = lattice(ix + 1, iz, seed), c = lattice(ix, iz + 1, seed), d = lattice(ix + 1, iz + 1, seed);
    return (a * (1 - tx) + b * tx) * (1 - tz) + (c * (1 - tx) + d * tx) * tz;
}
inline double fbm(double x, double z, uint64_t seed, int oct) {
    double s = 0, a = 0.5, f = 1;
    for (int o = 0; o < oct; o++) { s += a * vnoise(x * f, z * f, seed + o); a *= 0.5; f *= 2.03; }
    return s;
}

void heightfield(Gen& g, double x0, double z0, double x1, double z1, int R, double ybase, double yamp, double freq, uint64_t seed) {
    std::vector<double> h((size_t)(R + 1) * (R + 1));
    for (int j = 0; j <= R; j++) for (int i = 0; i <= R; i++) {
        double x = x0 + (x1 - x0) * i / R, z = z0 + (z1 - z0) * j / R;
        h[(size_t)j * (R + 1) + i] = ybase + yamp * fbm(x * freq, z * freq, seed, 6);
    }
    for (int j = 0; j < R; j++) for (int i = 0; i < R; i++) {
        double xa = x0 + (x1 - x0) * i / R, xb = x0 + (x1 - x0) * (i + 1) / R, za = z0 + (z1 - z0) * j / R, zb = z0 + (z1 - z0) * (j + 1) / R;
        V3 a{ xa, h[(size_t)j * (R + 1) + i], za }, b{ xb, h[(size_t)j * (R + 1) + i + 1], za };
        V3 c{ xb, h[(size_t)(j + 1) * (R + 1) + i + 1], zb }, d{ xa, h[(size_t)(j + 1) * (R + 1) + i], zb };
        g.quad(a, d, c, b);
    }
}

void truncate_to(tray_mesh* m, uint64_t target) {
    if (m->count() > target) m->tris.resize(target * 9);
    while (!m->offsets.empty() && m->offsets.back() >= m->count() && m->offsets.size() > 1) m->offsets.pop_back();
}

void set_cam(tray_mesh* m, V3 e, V3 l, double fov) {
    m->eye[0] = (float)e.x; m->eye[1] = (float)e.y; m->eye[2] = (float)e.z;
    m->look_at[0] = (float)l.x; m->look_at[1] = (float)l.y; m->look_at[2] = (float)l.z; m->fov = (float)fov;
}

// C1: kitchen-sized interior, 56,939 triangles (README.md:28), camera assets/scenes/kitchen.ron:3-8
void gen_kitchen(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; Rng r(seed * 7919 + 1);
    uint64_t target = std::max<uint64_t>(64, (uint64_t)(56939 * size));
    g.begin_object();
    const V3 lo{ -5, 0, -3.5 }, hi{ 4, 3, 3 }; V3 c = (lo + hi) * 0.5, h = (hi - lo) * 0.5;
    g.box(c, { h.x, 0, 0 }, { 0, h.y, 0 }, { 0, 0, h.z }, 8);
    while (m->count() < target) {
        int kind = (int)(r.uni() * 3);
        double s = r.logrange(0.05, 0.6);
        V3 p{ r.range(lo.x + 0.7, hi.x - 0.7), r.uni() < 0.7 ? s : r.range(0.3, 2.5), r.range(lo.z + 0.7, hi.z - 0.7) };
        if (kind == 0) {
            V3 ax = norm(V3{ r.range(-1, 1), 0, r.range(-1, 1) }); V3 az = cross(ax, V3{ 0, 1, 0 });
            g.box(p, ax * s, V3{ 0, s * r.range(0.3, 1.5), 0 }, az * (s * r.range(0.3, 1.0)), 1 + (int)(r.uni() * 4));
        } else if (kind == 1) {
            g.cylinder(p, s * 0.5, s * r.range(0.5, 2.0), 12 + (int)(r.uni() * 36));
        } else {
            int nl = 8 + (int)(r.uni() * 16);
            g.sphere(p, s * 0.6, nl, 2 * nl, [](V3) { return 0.0; });
        }
    }
    truncate_to(m, target);
    set_cam(m, { 3.0, 1.5, 1.4 }, { -3.9438584, 1.5, -1.7303504 }, 90);
}

// C2: stand-in for obvhs `demoscene(2048, 0)` (src/main.rs:244-257, external): an fBm height field,
// camera from main.rs:249-255.
void gen_demoscene(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; g.begin_object();
    int R = std::max(8, (int)std::lround(1024 * std::sqrt(size)));
    heightfield(g, -1, -1, 1, 1, R, -0.05, 0.6, 3.0, seed * 31 + 2);
    // the look_at sits above the field: tilt the field towards the camera the way the original frames it
    for (uint64_t i = 0; i < m->count() * 3; i++) { float* v = &m->tris[3 * i]; float y = v[1], z = v[2]; v[1] = -z * 0.35f + y * 0.6f + 0.1f; v[2] = y * 0.8f + z * 0.35f - 0.2f + 0.55f; }
    set_cam(m, { 0, 0, 1.35 }, { 0, 0.16, 0.35 }, 17);
}

// C3: hairball-like triangle soup, 2,880,000 triangles (README.md:30), camera assets/scenes/hairball.ron:3-8
void gen_hairball(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; Rng r(seed * 104729 + 3); g.begin_object();
    const int seg = 48;
    uint64_t strands = std::max<uint64_t>(4, (uint64_t)(30000 * size));
    for (uint64_t s = 0; s < strands; s++) {
        V3 dir = r.on_sphere(); V3 p = dir * 1.0; V3 side = norm(cross(dir, r.on_sphere()));
        double width = r.logrange(0.004, 0.02), step = r.range(0.06, 0.11);
        V3 a = p - side * width, b = p + side * width;
        for (int k = 0; k < seg; k++) {
            dir = norm(dir + r.on_sphere() * 0.45 + norm(p) * 0.12);
            p = p + dir * step;
            double rad = std::sqrt(dot(p, p));
            if (rad > 5.0) { p = p * (5.0 / rad); dir = norm(dir - norm(p) * dot(dir, norm(p))); }
            side = norm(side + r.on_sphere() * 0.2); side = norm(side - dir * dot(side, dir));
            V3 a2 = p - side * width, b2 = p + side * width;
            g.quad(a, b, b2, a2); a = a2; b = b2;
        }
    }
    set_cam(m, { 0, 0, 7 }, { 0, 0, 0 }, 90);
}

void clutter(Gen& g, Rng& r, V3 p, double s, int detail) {
    int kind = (int)(r.uni() * 10);
    uint64_t sd = r.next();
    if (kind == 0) {
        V3 ax = norm(V3{ r.range(-1, 1), r.range(-0.2, 0.2), r.range(-1, 1) }); V3 az = norm(cross(ax, V3{ 0, 1, 0 })); V3 ay = cross(az, ax);
        g.box(p, ax * s, ay * (s * r.range(0.3, 2.0)), az * (s * r.range(0.3, 1.0)), 1 + detail / 8);
    } else {
        int nl = std::max(4, detail);
        g.sphere(p, s, nl, 2 * nl, [sd](V3 d) { return 0.25 * (vnoise(d.x * 3 + 7, d.z * 3 + d.y * 2.7, sd) - 0.5); });
    }
}

// C4: San-Miguel-sized scene, 5,075,977 triangles (README.md:31), camera assets/scenes/san-miguel.ron:3-8
void gen_sanmiguel(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; Rng r(seed * 1299709 + 4);
    uint64_t target = std::max<uint64_t>(256, (uint64_t)(5075977 * size));
    g.begin_object();
    g.box({ 0, 6, 0 }, { 25, 0, 0 }, { 0, 6, 0 }, { 0, 0, 25 }, 16);
    int R = std::max(8, (int)std::lround(1023 * std::sqrt(size)));
    heightfield(g, -25, -25, 25, 25, R, 0.02, 0.8, 0.15, seed + 17);
    while (m->count() < target) {
        double s = r.logrange(0.05, 2.0);
        V3 p{ r.range(-24, 24), r.uni() < 0.75 ? 0.5 + s : r.range(1, 11), r.range(-24, 24) };
        clutter(g, r, p, s, r.uni() < 0.3 ? 50 : 25);
    }
    truncate_to(m, target);
    set_cam(m, { 22.0, 1.5, 13.0 }, { -13.761939, 1.5, -22.647648 }, 90);
}

// C5: Caldera-sized scene, 19,261,109 triangles (README.md:34) in 4,096 objects (one BLAS each under --tlas),
// camera assets/scenes/caldera_hotel_01.ron:3-8
void gen_caldera(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; Rng r(seed * 15485863 + 5);
    uint64_t target = std::max<uint64_t>(4096, (uint64_t)(19261109 * size));
    uint32_t n_obj = (uint32_t)std::max<double>(8, std::min<double>(4096, 4096 * std::sqrt(size)));
    // 64 terrain tiles (or fewer) + blobs; terrain takes ~1/8 of the budget
    uint32_t tiles_side = n_obj >= 1024 ? 8 : (n_obj >= 64 ? 4 : 1);
    uint32_t n_tiles = tiles_side * tiles_side;
    int R = std::max(2, (int)std::sqrt((double)target / 8.0 / 2.0 / n_tiles));
    for (uint32_t tz = 0; tz < tiles_side; tz++) for (uint32_t tx = 0; tx < tiles_side; tx++) {
        g.begin_object();
        double w = 400.0 / tiles_side;
        heightfield(g, -200 + tx * w, -200 + tz * w, -200 + (tx + 1) * w, -200 + (tz + 1) * w, R, -30, 40, 0.012, seed + 99);
    }
    uint32_t blobs = n_obj - n_tiles;
    uint64_t remaining = target > m->count() ? target - m->count() : 0;
    // log-uniform blob sizes, rescaled so they sum to the remaining budget
    std::vector<double> wgt(blobs); double wsum = 0;
    for (auto& w : wgt) { w = r.logrange(1000, 20000); wsum += w; }
    for (uint32_t k = 0; k < blobs; k++) {
        g.begin_object();
        uint64_t want = (uint64_t)(wgt[k] / wsum * (double)remaining);
        int nl = std::max(3, (int)std::lround((1.0 + std::sqrt(1.0 + (double)want)) / 2.0));   // 4 nl (nl - 1) ~ want
        double s = r.logrange(1.0, 10.0);
        double x = r.range(-195, 195), z = r.range(-195, 195);
        double y = -30 + 40 * fbm(x * 0.012, z * 0.012, seed + 99, 6) + (r.uni() < 0.8 ? s * 0.7 : r.range(5, 28));
        uint64_t sd = r.next();
        g.sphere({ x, y, z }, s, nl, 2 * nl, [sd](V3 d) { return 0.3 * (vnoise(d.x * 4 + 3, d.z * 4 + d.y * 3.1, sd) - 0.5); });
    }
    while (m->count() < target) {   // top up the last object, then trim to the exact count
        double s = r.logrange(1.0, 6.0);
        g.sphere({ r.range(-150, 150), r.range(0, 20), r.range(-150, 150) }, s, 40, 80, [](V3) { return 0.0; });
    }
    truncate_to(m, target);
    set_cam(m, { -40.0, 32.0, 84.0 }, { 30.0, -30.0, 0.0 }, 75);
}

// uniform random triangle soup in [-1,1]^3 (tests)
void gen_soup(tray_mesh* m, uint64_t seed, double size) {
    Gen g{ m }; Rng r(seed * 2654435761ull + 6); g.begin_object();
    uint64_t n = std::max<uint64_t>(1, (uint64_t)(1000000 * size));
    for (uint64_t i = 0; i < n; i++) {
        V3 c{ r.range(-1, 1), r.range(-1, 1), r.range(-1, 1) }; double s = r.logrange(0.005, 0.2);
        g.tri(c + r.on_sphere() * s, c + r.on_sphere() * s, c + r.on_sphere() * s);
    }
    set_cam(m, { 0, 0, 3 }, { 0, 0, 0 }, 60);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

int tray_host_mesh_generate(const char* name, uint64_t seed, double size, tray_mesh** out) {
    if (!name || !out || !(size > 0) || size > 1.0) return -1;
    tray_mesh* m = new tray_mesh();
    std::string n(name);
    if (n == "kitchen") gen_kitchen(m, seed, size);
    else if (n == "demoscene") gen_demoscene(m, seed, size);
    else if (n == "hairball") gen_hairball(m, seed, size);
    else if (n == "sanmiguel") gen_sanmiguel(m, seed, size);
    else if (n == "caldera") gen_caldera(m, seed, size);
    else if (n == "soup") gen_soup(m, seed, size);
    else { delete m; return -1; }
    if (m->offsets.empty()) m->offsets.push_back(0);
    m->offsets.push_back(m->count());
    *out = m;
    return 0;
}

int tray_host_mesh_from_tris(const float* tris9, uint64_t n_tris, const uint64_t* object_offsets,
                             uint32_t n_objects, tray_mesh** out) {
    if (!out || (n_tris && !tris9)) return -1;
    tray_mesh* m = new tray_mesh();
    m->tris.assign(tris9, tris9 + 9 * n_tris);
    if (object_offsets && n_objects) m->offsets.assign(object_offsets, object_offsets + n_objects + 1);
    else { m->offsets = { 0, n_tris }; }
    if (m->offsets.back() != n_tris || m->offsets.front() != 0) { delete m; return -1; }
    *out = m;
    return 0;
}

int tray_host_mesh_load_obj(const char* path, tray_mesh** out) {
    if (!path || !out) return -1;
    FILE* f = std::fopen(path, "rb");
    if (!f) return -1;
    tray_mesh* m = new tray_mesh();
    std::vector<float> pos;
    char line[4096];
    bool any_object = false;
    while (std::fgets(line, sizeof line, f)) {
        if (line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
            float x, y, z;
            if (std::sscanf(line + 2, "%f %f %f", &x, &y, &z) == 3) { pos.push_back(x); pos.push_back(y); pos.push_back(z); }
        } else if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\t')) {
            // one Vec<Triangle> per object (main.rs:530-559)
            if (any_object || m->count() > 0) m->offsets.push_back(m->count());
            else m->offsets.push_back(0);
            any_object = true;
        } else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
            long idx[4]; int n = 0; char* p = line + 2;
            while (*p) {
                while (*p == ' ' || *p == '\t') p++;
                if (*p == 0 || *p == '\n' || *p == '\r') break;
                char* end; long v = std::strtol(p, &end, 10);
                if (end == p) break;
                if (n < 4) idx[n] = v < 0 ? (long)(pos.size() / 3) + v : v - 1;
                n++;
                p = end; while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++;
            }
            if (n >= 3) {
                auto P = [&](int k) { return &pos[3 * (size_t)idx[k]]; };
                bool ok = true; for (int k = 0; k < std::min(n, 4); k++) ok = ok && idx[k] >= 0 && (size_t)idx[k] < pos.size() / 3;
                if (!ok) { std::fclose(f); delete m; return -1; }
                if (m->offsets.empty()) m->offsets.push_back(0);
                m->tris.insert(m->tris.end(), P(0), P(0) + 3); m->tris.insert(m->tris.end(), P(1), P(1) + 3); m->tris.insert(m->tris.end(), P(2), P(2) + 3);
                if (n == 4) {   // quad -> second triangle (a, c, d), main.rs:544-551
                    m->tris.insert(m->tris.end(), P(0), P(0) + 3); m->tris.insert(m->tris.end(), P(2), P(2) + 3); m->tris.insert(m->tris.end(), P(3), P(3) + 3);
                }
            }
        }
    }
    std::fclose(f);
    if (m->offsets.empty()) m->offsets.push_back(0);
    m->offsets.push_back(m->count());
    m->offsets.erase(std::unique(m->offsets.begin(), m->offsets.end()), m->offsets.end());   // drop empty objects
    if (m->offsets.size() < 2) m->offsets = { 0, m->count() };
    *out = m;
    return 0;
}

uint64_t tray_host_mesh_tri_count(const tray_mesh* m) { return m ? m->count() : 0; }
uint32_t tray_host_mesh_object_count(const tray_mesh* m) { return m ? (uint32_t)(m->offsets.size() - 1) : 0; }
const float* tray_host_mesh_tris(const tray_mesh* m) { return m ? m->tris.data() : nullptr; }
const uint64_t* tray_host_mesh_object_offsets(const tray_mesh* m) { return m ? m->offsets.data() : nullptr; }
void tray_host_mesh_camera(const tray_mesh* m, float eye[3], float look_at[3], float* fov) {
    for (int a = 0; a < 3; a++) { eye[a] = m->eye[a]; look_at[a] = m->look_at[a]; }
    *fov = m->fov;
}
void tray_host_mesh_free(tray_mesh* m) { delete m; }

}  // extern "C"

// ---- cwbvh_gpu_runner marshalling (src/rt_gpu/mod.rs:16-112) -------------------------------------
struct tray_packed {
    std::vector<tray_cwbvh_node> nodes;      // BLAS0 | BLAS1 | ... | TLAS      (mod.rs:62-69, 88-91)
    std::vector<uint8_t> tri_bytes;          // BVH-ordered records              (mod.rs:34-38, 80-86)
    std::vector<uint32_t> instance;          // blas_offsets in TLAS-leaf order  (mod.rs:72-78); [0;4] when flat (mod.rs:109)
    std::vector<uint32_t> prim_to_mesh_tri;
    std::vector<uint64_t> blas_tri_offsets;
    uint32_t tlas_start = 0, max_depth = 0;
    double build_s = 0, tlas_s = 0;
};

extern "C" {

int tray_host_pack(const tray_mesh* mesh, int use_tlas, uint32_t tri_stride, uint32_t max_prims_per_leaf,
                   int nthreads, tray_packed** out) {
    if (!mesh || !out || (tri_stride != 48 && tri_stride != 64 && tri_stride != 24)) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    tray_packed* p = new tray_packed();
    // flatten unless --tlas (main.rs:300-308)
    std::vector<uint64_t> offs = use_tlas ? mesh->offsets : std::vector<uint64_t>{ 0, mesh->count() };
    const uint32_t n_obj = (uint32_t)offs.size() - 1;
    std::vector<tray_cwbvh*> blas(n_obj, nullptr);
    double t0 = now_s();
    int rc = 0;
    if (n_obj == 1) {
        rc = tray_host_build_cwbvh_from_tris(mesh->tris.data(), mesh->count(), max_prims_per_leaf, nthreads, &blas[0]);
    } else {
        // independent BLAS builds: one thread each
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
        for (int64_t k = 0; k < (int64_t)n_obj; k++) {
            int e = tray_host_build_cwbvh_from_tris(mesh->tris.data() + 9 * offs[k], offs[k + 1] - offs[k], max_prims_per_leaf, 1, &blas[k]);
            if (e) {
#pragma omp atomic write
                rc = e;
            }
        }
    }
    p->build_s = now_s() - t0;
    if (rc) { for (auto b : blas) tray_host_cwbvh_free(b); delete p; return rc; }

    // node offsets (blas_mapping, mod.rs:63-68) and global triangle offsets (tri_offset, mod.rs:28,45-48)
    std::vector<uint64_t> node_off(n_obj + 1, 0);
    p->blas_tri_offsets.assign(n_obj + 1, 0);
    for (uint32_t k = 0; k < n_obj; k++) {
        node_off[k + 1] = node_off[k] + tray_host_cwbvh_node_count(blas[k]);
        p->blas_tri_offsets[k + 1] = p->blas_tri_offsets[k] + tray_host_cwbvh_prim_count(blas[k]);
        p->max_depth = std::max(p->max_depth, tray_host_cwbvh_max_depth(blas[k]));
    }
    const uint64_t n_tris = p->blas_tri_offsets[n_obj];
    p->nodes.resize(node_off[n_obj]);
    p->tri_bytes.assign(n_tris * tri_stride, 0);
    p->prim_to_mesh_tri.resize(n_tris);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int64_t k = 0; k < (int64_t)n_obj; k++) {
        const tray_cwbvh_node* bn = tray_host_cwbvh_nodes(blas[k]);
        const uint64_t nn = tray_host_cwbvh_node_count(blas[k]);
        const uint32_t tri_off = (uint32_t)p->blas_tri_offsets[k];
        for (uint64_t i = 0; i < nn; i++) {
            tray_cwbvh_node n = bn[i];
            n.primitive_base_idx += tri_off;                                  // mod.rs:45-47
            p->nodes[node_off[k] + i] = n;
        }
        const uint32_t* pi = tray_host_cwbvh_prim_indices(blas[k]);
        const uint64_t np = tray_host_cwbvh_prim_count(blas[k]);
        for (uint64_t i = 0; i < np; i++) {                                   // mod.rs:34-38: tris[primitive_indices[i]]
            const uint64_t mesh_tri = offs[k] + pi[i];
            const float* t = mesh->tris.data() + 9 * mesh_tri;
            float* rec = (float*)(p->tri_bytes.data() + (uint64_t)(tri_off + i) * tri_stride);
            // RtTriangle::from(&Triangle): v0, e1 = v0 - v1, e2 = v2 - v0, ng = cross(e1, e2)  
            float v0[3] = { t[0], t[1], t[2] }, e1[3], e2[3];
            for (int a = 0; a < 3; a++) { e1[a] = t[a] - t[3 + a]; e2[a] = t[6 + a] - t[a]; }
            if (tri_stride == 24) {
                // RtCompressedTriangle::from(&Triangle): v0 as f32, e[k] = half(v2 - v0)[k] | half(v1 - v0)[k] << 16
                // (src/rt_gpu/mod.rs:39-43; unpacked at rt_gpu_software_query.hlsl:75-85)
                uint32_t* e = (uint32_t*)(rec + 3);
                for (int a = 0; a < 3; a++) {
                    rec[a] = v0[a];
                    e[a] = (uint32_t)float_to_half(e2[a]) | ((uint32_t)float_to_half(t[3 + a] - t[a]) << 16);
                }
                p->prim_to_mesh_tri[tri_off + i] = (uint32_t)mesh_tri;
                continue;
            }
            for (int a = 0; a < 3; a++) { rec[a] = v0[a]; rec[4 + a] = e1[a]; rec[8 + a] = e2[a]; }
            if (tri_stride == 64) {
                volatile float m0 = e1[1] * e2[2], m1 = e1[2] * e2[1], m2 = e1[2] * e2[0], m3 = e1[0] * e2[2], m4 = e1[0] * e2[1], m5 = e1[1] * e2[0];
                rec[12] = m0 - m1; rec[13] = m2 - m3; rec[14] = m4 - m5;     // separately rounded mul, mul, sub
            }
            p->prim_to_mesh_tri[tri_off + i] = (uint32_t)mesh_tri;
        }
    }
    if (use_tlas) {
        // tlas_from_blas: CWBVH over the BLAS total_aabbs (cwbvh.rs:114,130-132)
        std::vector<float> mn(3 * (size_t)n_obj), mx(3 * (size_t)n_obj);
        for (uint32_t k = 0; k < n_obj; k++) tray_host_cwbvh_aabb(blas[k], &mn[3 * k], &mx[3 * k]);
        double t1 = now_s();
        tray_cwbvh* tlas = nullptr;
        rc = tray_host_build_cwbvh_from_aabbs(mn.data(), mx.data(), n_obj, max_prims_per_leaf, nthreads, &tlas);
        p->tlas_s = now_s() - t1;
        if (!rc) {
            p->tlas_start = (uint32_t)node_off[n_obj];                        // mod.rs:99 (blas_len)
            const uint32_t* tpi = tray_host_cwbvh_prim_indices(tlas);
            const uint64_t tn = tray_host_cwbvh_prim_count(tlas);
            p->instance.resize(tn);
            for (uint64_t i = 0; i < tn; i++) p->instance[i] = (uint32_t)node_off[tpi[i]];   // mod.rs:72-78
            const tray_cwbvh_node* tnodes = tray_host_cwbvh_nodes(tlas);
            p->nodes.insert(p->nodes.end(), tnodes, tnodes + tray_host_cwbvh_node_count(tlas));   // mod.rs:88-91
            p->max_depth += tray_host_cwbvh_max_depth(tlas);
        }
        tray_host_cwbvh_free(tlas);
    } else {
        p->instance.assign(4, 0);                                              // &[0; 16] bytes, mod.rs:109
        p->tlas_start = 0;
    }
    for (auto b : blas) tray_host_cwbvh_free(b);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

const void* tray_host_packed_bvh_bytes(const tray_packed* p, uint64_t* len) { if (len) *len = p->nodes.size() * sizeof(tray_cwbvh_node); return p->nodes.data(); }
const void* tray_host_packed_tri_bytes(const tray_packed* p, uint64_t* len) { if (len) *len = p->tri_bytes.size(); return p->tri_bytes.data(); }
const void* tray_host_packed_instance_bytes(const tray_packed* p, uint64_t* len) { if (len) *len = p->instance.size() * 4; return p->instance.data(); }
const uint32_t* tray_host_packed_prim_to_mesh_tri(const tray_packed* p, uint64_t* count) { if (count) *count = p->prim_to_mesh_tri.size(); return p->prim_to_mesh_tri.data(); }
const uint64_t* tray_host_packed_blas_tri_offsets(const tray_packed* p, uint32_t* n_blas) { if (n_blas) *n_blas = (uint32_t)p->blas_tri_offsets.size() - 1; return p->blas_tri_offsets.data(); }
uint32_t tray_host_packed_tlas_start(const tray_packed* p) { return p->tlas_start; }
uint32_t tray_host_packed_max_depth(const tray_packed* p) { return p->max_depth; }
double tray_host_packed_build_seconds(const tray_packed* p, double* tlas_seconds) { if (tlas_seconds) *tlas_seconds = p->tlas_s; return p->build_s; }
void tray_host_packed_free(tray_packed* p) { delete p; }

// ViewUniform::from_camera (src/main.rs:599-616): proj_inv = perspective_infinite_reverse_rh(fov, aspect, 0.01)^-1,
// view_inv = look_at_rh(eye, look_at, +Y)^-1, column-major.  Both have closed-form inverses; they are evaluated
// in double and rounded once to f32.  The uniform is an INPUT to the oracle and to the kernel alike.
void tray_host_view_from_camera(const float eye[3], const float look_at[3], float fov_deg,
                                float width, float height, float exposure, uint32_t tlas_start, tray_view* out) {
    std::memset(out, 0, sizeof(*out));
    const double aspect = (double)(width / height);
    const double fov = (double)(fov_deg * (3.14159265358979323846f / 180.0f));
    const double f = 1.0 / std::tan(0.5 * fov), zn = 0.01;
    // P = cols (f/aspect,0,0,0) (0,f,0,0) (0,0,0,-1) (0,0,zn,0)  =>  P^-1 = cols (aspect/f,0,0,0) (0,1/f,0,0) (0,0,0,1/zn) (0,0,-1,0)
    out->proj_inv[0] = (float)(aspect / f); out->proj_inv[5] = (float)(1.0 / f);
    out->proj_inv[11] = (float)(1.0 / zn); out->proj_inv[14] = -1.0f;
    V3 e{ eye[0], eye[1], eye[2] }, c{ look_at[0], look_at[1], look_at[2] };
    V3 fw = norm(c - e), s = norm(cross(fw, V3{ 0, 1, 0 })), u = cross(s, fw);
    // V = [R | -R e] with rows s, u, -fw  =>  V^-1 = cols (s,0) (u,0) (-fw,0) (e,1)
    float* m = out->view_inv;
    m[0] = (float)s.x; m[1] = (float)s.y; m[2] = (float)s.z; m[3] = 0;
    m[4] = (float)u.x; m[5] = (float)u.y; m[6] = (float)u.z; m[7] = 0;
    m[8] = (float)-fw.x; m[9] = (float)-fw.y; m[10] = (float)-fw.z; m[11] = 0;
    m[12] = eye[0]; m[13] = eye[1]; m[14] = eye[2]; m[15] = 1;
    out->eye[0] = eye[0]; out->eye[1] = eye[1]; out->eye[2] = eye[2];
    out->exposure = exposure; out->tlas_start = tlas_start;
}

}  // extern "C"
