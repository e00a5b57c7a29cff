// Scene feed: the host-side writer of the scene buffer the CUDA renderer consumes.
//
// Behavioural restatement, in C++ (no Rust toolchain in this image), of the reference's feed:
//   - `Encoder`                       src/lib.rs:79-254
//   - make_cardioid / make_path_test  src/lib.rs:257-284
//   - make_tiger, encode_path[_stroke], parse_color, init_test_scene   src/lib.rs:286-393
//   - flatten_path                    src/flatten.rs:10-47
// plus the pieces of kurbo 0.5.6 (Cargo.lock:8-14; NOT vendored in the reference tree) that the
// feed calls: BezPath::from_svg, Affine::scale * BezPath, CubicBez::to_quads, Arc::from_svg_arc /
// append_iter, Rect::{from_points, union_pt, inflate}.  Those are restated from kurbo's published
// algorithms and are unverifiable here ("parity unpinned" for the feed); the
// hot-path contract is defined on the *encoded* scene, so this cannot affect kernel parity.
//
// All geometry is f64 until it is narrowed to f32 at encode time (point_to_f32s, lib.rs:99-101).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cstdio>
#include <vector>

#include "../../include/piet_metal_b200.h"
#include "pm_scene_format.h"

// The tiger path list (tools/make_tiger_fixture.py) is embedded like the reference embeds the SVG
// with include_bytes! (src/lib.rs:288).
#ifndef PM_ASSET_DIR
#error "PM_ASSET_DIR must point at piet-metal_b200/assets"
#endif
__asm__(".section .rodata\n"
        ".global pm_tiger_pathlist_begin\n"
        "pm_tiger_pathlist_begin:\n"
        ".incbin \"" PM_ASSET_DIR "/tiger.pathlist\"\n"
        ".global pm_tiger_pathlist_end\n"
        "pm_tiger_pathlist_end:\n"
        ".byte 0\n"
        ".previous\n");
extern "C" const char pm_tiger_pathlist_begin[];
extern "C" const char pm_tiger_pathlist_end[];

namespace {

struct Pt { double x, y; };
struct Rect { double x0, y0, x1, y1; };

// ---------------------------------------------------------------------------------------------
// Encoder (src/lib.rs:79-254)
// ---------------------------------------------------------------------------------------------
struct EncoderImpl {
    uint8_t *buf;
    size_t cap;
    size_t free_space = 0;   // lib.rs:81
    size_t group_count = 0;
    size_t group_ix = 0;
    size_t group_start = 0;
    bool overflow = false;   // the reference panics on a slice index instead (lib.rs:127)
    bool in_group = false;

    size_t alloc(size_t size) {  // lib.rs:112-116
        size_t r = free_space;
        free_space += size;
        return r;
    }
    void write(size_t ix, const void *src, size_t len) {  // write_struct, lib.rs:120-130
        if (buf == nullptr || ix + len > cap) { overflow = true; return; }
        memcpy(buf + ix, src, len);
    }
};

uint32_t to_be(uint32_t v) {  // u32::to_be on a little-endian host (lib.rs:181,200,213)
    return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
}

pm_bbox short_bbox(Rect r) {  // ShortBbox::from_rect, lib.rs:88-97
    auto c = [](double v) { return (uint16_t)std::min(std::max(v, 0.0), 65535.0); };
    pm_bbox b;
    b.x0 = c(std::floor(r.x0));
    b.y0 = c(std::floor(r.y0));
    b.x1 = c(std::ceil(r.x1));
    b.y1 = c(std::ceil(r.y1));
    return b;
}

Rect inflate(Rect r, double w, double h) { return Rect{r.x0 - w, r.y0 - h, r.x1 + w, r.y1 + h}; }

int enc_begin_group(EncoderImpl &e, size_t n_items) {  // lib.rs:132-144
    size_t item_start = PM_GROUP_HEADER_SIZE + n_items * PM_BBOX_SIZE;
    size_t total = item_start + n_items * PM_ITEM_SIZE;
    e.group_start = e.alloc(total);
    e.group_count = n_items;
    e.group_ix = 0;
    e.in_group = true;
    pm_group_header g;
    g.n_items = (uint32_t)n_items;
    g.items_ix = (uint32_t)(e.group_start + item_start);
    e.write(e.group_start, &g, sizeof g);
    return PM_OK;
}

int enc_add_item(EncoderImpl &e, const void *item, size_t item_len, pm_bbox bbox) {  // lib.rs:151-163
    if (!e.in_group || e.group_ix >= e.group_count) return PM_ERR_STATE;  // assert! at :152
    size_t bbox_ix = e.group_start + PM_GROUP_HEADER_SIZE + e.group_ix * PM_BBOX_SIZE;
    e.write(bbox_ix, &bbox, sizeof bbox);
    size_t item_ix = e.group_start + PM_GROUP_HEADER_SIZE + e.group_count * PM_BBOX_SIZE + e.group_ix * PM_ITEM_SIZE;
    // The reference copies size_of::<T>() bytes of the variant and leaves the rest of the 32-byte
    // slot as it was; write a zero-padded slot so the encoding is deterministic.
    uint8_t slot[PM_ITEM_SIZE];
    memset(slot, 0, sizeof slot);
    memcpy(slot, item, item_len);
    e.write(item_ix, slot, sizeof slot);
    e.group_ix += 1;
    return PM_OK;
}

// encode_points, lib.rs:224-240
bool enc_points(EncoderImpl &e, const Pt *pts, size_t n, size_t *points_ix, Rect *bbox) {
    if (n == 0) return false;  // .expect("encoded empty points vector")
    size_t ix = e.alloc(n * 8);
    *points_ix = ix;
    Rect bb{pts[0].x, pts[0].y, pts[0].x, pts[0].y};
    for (size_t i = 0; i < n; i++) {
        bb.x0 = std::min(bb.x0, pts[i].x);
        bb.y0 = std::min(bb.y0, pts[i].y);
        bb.x1 = std::max(bb.x1, pts[i].x);
        bb.y1 = std::max(bb.y1, pts[i].y);
        float f[2] = {(float)pts[i].x, (float)pts[i].y};
        e.write(ix + 8 * i, f, 8);
    }
    *bbox = bb;
    return true;
}

int enc_circle(EncoderImpl &e, Pt c, double r) {  // lib.rs:167-174
    uint32_t tag = PM_ITEM_CIRCLE;
    return enc_add_item(e, &tag, 4, short_bbox(Rect{c.x - r, c.y - r, c.x + r, c.y + r}));
}

int enc_stroke_line(EncoderImpl &e, Pt p0, Pt p1, float width, uint32_t rgba) {  // lib.rs:177-192
    pm_item_line it;
    it.tag = PM_ITEM_LINE;
    it.flags = 0;
    it.rgba = to_be(rgba);
    it.width = width;
    it.sx = (float)p0.x; it.sy = (float)p0.y;
    it.ex = (float)p1.x; it.ey = (float)p1.y;
    double hw = (double)(width * 0.5f);
    Rect bb{std::min(p0.x, p1.x), std::min(p0.y, p1.y), std::max(p0.x, p1.x), std::max(p0.y, p1.y)};
    return enc_add_item(e, &it, sizeof it, short_bbox(inflate(bb, hw, hw)));
}

int enc_fill(EncoderImpl &e, const Pt *pts, size_t n, uint32_t rgba, uint32_t flags = 0) {  // lib.rs:195-207 (flags: always 0 there)
    size_t pix; Rect bb;
    if (!enc_points(e, pts, n, &pix, &bb)) return PM_ERR_INVALID_ARG;
    pm_item_fill it;
    memset(&it, 0, sizeof it);
    it.tag = PM_ITEM_FILL;
    it.flags = flags;
    it.rgba = to_be(rgba);
    it.n_points = (uint32_t)n;
    it.points_ix = (uint32_t)pix;
    return enc_add_item(e, &it, 20, short_bbox(bb));
}

// Several closed subpaths as ONE Fill item ("need to deal with subpaths", lib.rs:194): subpath 0, then every further
// subpath closed explicitly and followed by a bridge back to the first point of the path.  The kernels close the
// list from its last point to its first (metal:262); every bridge is walked once in each direction, so winding
// and signed area cancel and what is left is the union of the subpaths' own windings.
int enc_fill_subpaths(EncoderImpl &e, const std::vector<std::vector<Pt>> &sub, uint32_t rgba, uint32_t flags) {
    std::vector<Pt> joined;
    for (size_t i = 0; i < sub.size(); i++) {
        if (sub[i].empty()) continue;
        if (joined.empty()) {
            joined = sub[i];
            joined.push_back(sub[i][0]);
        } else {
            joined.insert(joined.end(), sub[i].begin(), sub[i].end());
            joined.push_back(sub[i][0]);
            joined.push_back(joined[0]);
        }
    }
    if (joined.empty()) return PM_ERR_INVALID_ARG;
    return enc_fill(e, joined.data(), joined.size(), rgba, flags);
}

int enc_polyline(EncoderImpl &e, const Pt *pts, size_t n, uint32_t rgba, float width) {  // lib.rs:209-222
    size_t pix; Rect bb;
    if (!enc_points(e, pts, n, &pix, &bb)) return PM_ERR_INVALID_ARG;
    pm_item_poly it;
    memset(&it, 0, sizeof it);
    it.tag = PM_ITEM_POLY;
    it.rgba = to_be(rgba);
    it.width = width;
    it.n_points = (uint32_t)n;
    it.points_ix = (uint32_t)pix;
    double hw = (double)(width * 0.5f);
    return enc_add_item(e, &it, 20, short_bbox(inflate(bb, hw, hw)));
}

// ---------------------------------------------------------------------------------------------
// Paths: kurbo BezPath::from_svg restated to the SVG path grammar
// ---------------------------------------------------------------------------------------------
enum ElKind { MOVE, LINE, QUAD, CURVE, CLOSE };
struct PathEl { ElKind k; Pt p1, p2, p3; };  // MOVE/LINE use p1; QUAD p1,p2; CURVE p1,p2,p3
typedef std::vector<PathEl> BezPath;

struct Lexer {
    const char *s;
    size_t i = 0, n;
    explicit Lexer(const char *str) : s(str), n(strlen(str)) {}
    void skip_ws() { while (i < n && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r' || s[i] == '\f')) i++; }
    void opt_comma() { skip_ws(); if (i < n && s[i] == ',') { i++; skip_ws(); } }
    bool at_number() { skip_ws(); return i < n && (s[i] == '-' || s[i] == '+' || s[i] == '.' || (s[i] >= '0' && s[i] <= '9')); }
    bool number(double *out) {
        skip_ws();
        size_t st = i;
        if (i < n && (s[i] == '-' || s[i] == '+')) i++;
        size_t digits = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') { i++; digits++; }
        if (i < n && s[i] == '.') {
            i++;
            while (i < n && s[i] >= '0' && s[i] <= '9') { i++; digits++; }
        }
        if (digits == 0) { i = st; return false; }
        if (i < n && (s[i] == 'e' || s[i] == 'E')) {
            size_t save = i;
            i++;
            if (i < n && (s[i] == '-' || s[i] == '+')) i++;
            size_t ed = 0;
            while (i < n && s[i] >= '0' && s[i] <= '9') { i++; ed++; }
            if (ed == 0) i = save;
        }
        *out = strtod(std::string(s + st, i - st).c_str(), nullptr);
        opt_comma();
        return true;
    }
    bool flag(bool *out) {  // arc flags are single characters and need no separator
        skip_ws();
        if (i < n && (s[i] == '0' || s[i] == '1')) { *out = s[i] == '1'; i++; opt_comma(); return true; }
        return false;
    }
    bool pair(Pt *p) { return number(&p->x) && number(&p->y); }
};

void append_arc(BezPath &path, Pt from, Pt to, double rx_in, double ry_in, double x_rot_deg, bool large_arc, bool sweep);

bool parse_svg_path(const char *d, BezPath &path) {
    Lexer lx(d);
    Pt cur{0, 0}, start{0, 0}, last_ctrl{0, 0};
    bool have_ctrl_c = false, have_ctrl_q = false;
    bool subpath_open = false;  // false right after Z: a drawing command then re-opens at `start`
    char cmd = 0;
    lx.skip_ws();
    while (lx.i < lx.n) {
        char c = lx.s[lx.i];
        if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) {
            cmd = c;
            lx.i++;
        } else if (cmd == 0 || !lx.at_number()) {
            return false;
        } else if (cmd == 'M') {
            cmd = 'L';  // implicit lineto after the first moveto pair
        } else if (cmd == 'm') {
            cmd = 'l';
        } else if (cmd == 'z' || cmd == 'Z') {
            return false;
        }
        bool rel = cmd >= 'a' && cmd <= 'z';
        char uc = (char)(rel ? cmd - 32 : cmd);
        auto reopen = [&]() {
            if (!subpath_open) {
                path.push_back(PathEl{MOVE, start, {}, {}});
                cur = start;
                subpath_open = true;
            }
        };
        auto abs_pt = [&](Pt p) { return rel ? Pt{cur.x + p.x, cur.y + p.y} : p; };
        switch (uc) {
            case 'M': {
                Pt p;
                if (!lx.pair(&p)) return false;
                p = abs_pt(p);
                path.push_back(PathEl{MOVE, p, {}, {}});
                cur = start = p;
                subpath_open = true;
                have_ctrl_c = have_ctrl_q = false;
                break;
            }
            case 'L': {
                Pt p;
                if (!lx.pair(&p)) return false;
                reopen();
                p = abs_pt(p);
                path.push_back(PathEl{LINE, p, {}, {}});
                cur = p;
                have_ctrl_c = have_ctrl_q = false;
                break;
            }
            case 'H': case 'V': {
                double v;
                if (!lx.number(&v)) return false;
                reopen();
                Pt p = cur;
                if (uc == 'H') p.x = rel ? cur.x + v : v; else p.y = rel ? cur.y + v : v;
                path.push_back(PathEl{LINE, p, {}, {}});
                cur = p;
                have_ctrl_c = have_ctrl_q = false;
                break;
            }
            case 'C': {
                Pt a, b, p;
                if (!lx.pair(&a) || !lx.pair(&b) || !lx.pair(&p)) return false;
                reopen();
                a = abs_pt(a); b = abs_pt(b); p = abs_pt(p);
                path.push_back(PathEl{CURVE, a, b, p});
                last_ctrl = b; cur = p;
                have_ctrl_c = true; have_ctrl_q = false;
                break;
            }
            case 'S': {
                Pt b, p;
                if (!lx.pair(&b) || !lx.pair(&p)) return false;
                reopen();
                Pt a = have_ctrl_c ? Pt{2 * cur.x - last_ctrl.x, 2 * cur.y - last_ctrl.y} : cur;
                b = abs_pt(b); p = abs_pt(p);
                path.push_back(PathEl{CURVE, a, b, p});
                last_ctrl = b; cur = p;
                have_ctrl_c = true; have_ctrl_q = false;
                break;
            }
            case 'Q': {
                Pt a, p;
                if (!lx.pair(&a) || !lx.pair(&p)) return false;
                reopen();
                a = abs_pt(a); p = abs_pt(p);
                path.push_back(PathEl{QUAD, a, p, {}});
                last_ctrl = a; cur = p;
                have_ctrl_q = true; have_ctrl_c = false;
                break;
            }
            case 'T': {
                Pt p;
                if (!lx.pair(&p)) return false;
                reopen();
                Pt a = have_ctrl_q ? Pt{2 * cur.x - last_ctrl.x, 2 * cur.y - last_ctrl.y} : cur;
                p = abs_pt(p);
                path.push_back(PathEl{QUAD, a, p, {}});
                last_ctrl = a; cur = p;
                have_ctrl_q = true; have_ctrl_c = false;
                break;
            }
            case 'A': {
                double rx, ry, rot;
                bool large, sw;
                Pt p;
                if (!lx.number(&rx) || !lx.number(&ry) || !lx.number(&rot) || !lx.flag(&large) || !lx.flag(&sw) || !lx.pair(&p))
                    return false;
                reopen();
                p = abs_pt(p);
                append_arc(path, cur, p, rx, ry, rot, large, sw);
                cur = p;
                have_ctrl_c = have_ctrl_q = false;
                break;
            }
            case 'Z': {
                path.push_back(PathEl{CLOSE, {}, {}, {}});
                cur = start;
                subpath_open = false;
                have_ctrl_c = have_ctrl_q = false;
                break;
            }
            default:
                return false;
        }
        lx.skip_ws();
    }
    return true;
}

// kurbo Arc::from_svg_arc (SVG implementation notes F.6.5) followed by Arc::append_iter(0.1): the
// elliptical arc becomes n cubic Beziers, n from the tolerance rule below.  Restated from memory.
Pt sample_ellipse(double rx, double ry, double x_rot, double angle) {
    double u = rx * std::cos(angle), v = ry * std::sin(angle);
    return Pt{u * std::cos(x_rot) - v * std::sin(x_rot), u * std::sin(x_rot) + v * std::cos(x_rot)};
}

void append_arc(BezPath &path, Pt from, Pt to, double rx_in, double ry_in, double x_rot_deg, bool large_arc, bool sweep) {
    const double PI = 3.14159265358979323846;
    double rx = std::fabs(rx_in), ry = std::fabs(ry_in);
    // SvgArc::is_straight_line: degenerate radii or coincident end points
    if (rx <= 1e-5 || ry <= 1e-5 || (from.x == to.x && from.y == to.y)) {
        path.push_back(PathEl{LINE, to, {}, {}});
        return;
    }
    double x_rot = x_rot_deg * (PI / 180.0);
    double xr = std::fmod(x_rot, 2.0 * PI);
    double sin_phi = std::sin(xr), cos_phi = std::cos(xr);
    double hd_x = (from.x - to.x) * 0.5, hd_y = (from.y - to.y) * 0.5;
    double hs_x = (from.x + to.x) * 0.5, hs_y = (from.y + to.y) * 0.5;
    Pt p{cos_phi * hd_x + sin_phi * hd_y, -sin_phi * hd_x + cos_phi * hd_y};
    double rf = p.x * p.x / (rx * rx) + p.y * p.y / (ry * ry);
    if (rf > 1.0) {
        double sc = std::sqrt(rf);
        rx *= sc; ry *= sc;
    }
    double rxry = rx * ry, rxpy = rx * p.y, rypx = ry * p.x;
    double sum_of_sq = rxpy * rxpy + rypx * rypx;
    if (sum_of_sq == 0.0) { path.push_back(PathEl{LINE, to, {}, {}}); return; }
    double sign_coe = (large_arc == sweep) ? -1.0 : 1.0;
    double coe = sign_coe * std::sqrt(std::fabs((rxry * rxry - sum_of_sq) / sum_of_sq));
    double tcx = coe * rxpy / ry, tcy = -coe * rypx / rx;
    Pt center{cos_phi * tcx - sin_phi * tcy + hs_x, sin_phi * tcx + cos_phi * tcy + hs_y};
    double start_angle = std::atan2((p.y - tcy) / ry, (p.x - tcx) / rx);
    double end_angle = std::atan2((-p.y - tcy) / ry, (-p.x - tcx) / rx);
    double sweep_angle = std::fmod(end_angle - start_angle, 2.0 * PI);
    if (sweep && sweep_angle < 0.0) sweep_angle += 2.0 * PI;
    else if (!sweep && sweep_angle > 0.0) sweep_angle -= 2.0 * PI;

    const double tolerance = 0.1;
    double sign = sweep_angle > 0 ? 1.0 : (sweep_angle < 0 ? -1.0 : 0.0);
    double scaled_err = std::max(rx, ry) / tolerance;
    double n_err = std::max(std::pow(1.1163 * scaled_err, 1.0 / 6.0), 3.999999);
    double nf = std::ceil(n_err * std::fabs(sweep_angle) * (1.0 / (2.0 * PI)));
    double angle_step = sweep_angle / nf;
    size_t n = (size_t)nf;
    double arm_len = (4.0 / 3.0) * std::tan(std::fabs(0.25 * angle_step)) * sign;
    double angle0 = start_angle;
    Pt p0 = sample_ellipse(rx, ry, x_rot, angle0);
    for (size_t i = 0; i < n; i++) {
        double angle1 = angle0 + angle_step;
        Pt d0 = sample_ellipse(rx, ry, x_rot, angle0 + PI / 2.0);
        Pt c1{p0.x + arm_len * d0.x, p0.y + arm_len * d0.y};
        Pt p3 = sample_ellipse(rx, ry, x_rot, angle1);
        Pt d1 = sample_ellipse(rx, ry, x_rot, angle1 + PI / 2.0);
        Pt c2{p3.x - arm_len * d1.x, p3.y - arm_len * d1.y};
        path.push_back(PathEl{CURVE, Pt{center.x + c1.x, center.y + c1.y}, Pt{center.x + c2.x, center.y + c2.y},
                              Pt{center.x + p3.x, center.y + p3.y}});
        angle0 = angle1;
        p0 = p3;
    }
}

void scale_path(BezPath &path, double s) {  // kurbo::Affine::scale(s) * &BezPath, lib.rs:297,314
    for (auto &el : path) {
        el.p1.x *= s; el.p1.y *= s;
        el.p2.x *= s; el.p2.y *= s;
        el.p3.x *= s; el.p3.y *= s;
    }
}

// kurbo CubicBez::eval
Pt cubic_eval(Pt p0, Pt p1, Pt p2, Pt p3, double t) {
    double mt = 1.0 - t;
    double a = mt * mt * mt, b = mt * mt * 3.0, c = mt * 3.0;
    return Pt{p0.x * a + (p1.x * b + (p2.x * c + p3.x * t) * t) * t,
              p0.y * a + (p1.y * b + (p2.y * c + p3.y * t) * t) * t};
}

// flatten_path (src/flatten.rs:10-47).  Cubics go through kurbo's CubicBez::to_quads(tolerance *
// 1e-2) and only the end point of each quad is kept (:35-37), i.e. uniform subdivision of the
// parameter range into n pieces, n = max(1, ceil((|3p2-p3-3p1+p0|^2 / (432 acc^2))^(1/6))).
// QuadTo and ClosePath elements are ignored (:40).
void flatten_path(const BezPath &path, double tolerance, std::vector<std::vector<Pt>> &result) {
    bool have_cur = false;
    std::vector<Pt> cur;
    Pt last{0, 0};
    for (const PathEl &el : path) {
        switch (el.k) {
            case MOVE:
                if (have_cur) result.push_back(cur);
                cur.clear();
                cur.push_back(el.p1);
                have_cur = true;
                last = el.p1;
                break;
            case LINE:
                if (!have_cur) break;  // the reference would panic on unwrap (:22)
                cur.push_back(el.p1);
                last = el.p1;
                break;
            case CURVE: {
                if (!have_cur) break;
                double acc = tolerance * 1e-2;
                double max_hypot2 = 432.0 * acc * acc;
                double ex = (3.0 * el.p2.x - el.p3.x) - (3.0 * el.p1.x - last.x);
                double ey = (3.0 * el.p2.y - el.p3.y) - (3.0 * el.p1.y - last.y);
                double err = ex * ex + ey * ey;
                double nf = std::max(1.0, std::ceil(std::pow(err / max_hypot2, 1.0 / 6.0)));
                size_t n = (size_t)nf;
                for (size_t i = 0; i < n; i++) {
                    double t1 = (double)(i + 1) / (double)n;
                    cur.push_back(cubic_eval(last, el.p1, el.p2, el.p3, t1));
                }
                last = el.p3;
                break;
            }
            default:
                break;
        }
    }
    if (have_cur) result.push_back(cur);
}

const double TOLERANCE = 0.1;  // lib.rs:330
const float THIN_LINE = 0.7f;  // lib.rs:351

uint32_t parse_color(const char *s) {  // lib.rs:375-385
    if (s[0] == '#') {
        size_t len = strlen(s);
        uint32_t hex = (uint32_t)strtoul(s + 1, nullptr, 16);
        if (len == 4) hex = (hex >> 8) * 0x110000 + ((hex >> 4) & 0xf) * 0x1100 + (hex & 0xf) * 0x11;
        return (hex << 8) + 0xff;
    }
    return 0xff00ff80u;
}

struct PathListEntry { std::string fill, stroke, width, d; };

bool parse_pathlist(const char *text, size_t len, std::vector<PathListEntry> &out) {
    size_t i = 0;
    while (i < len) {
        size_t e = i;
        while (e < len && text[e] != '\n') e++;
        std::string line(text + i, e - i);
        i = e + 1;
        if (line.compare(0, 5, "path ") != 0) continue;
        PathListEntry pe;
        size_t p = 5;
        std::string *fields[3] = {&pe.fill, &pe.stroke, &pe.width};
        for (auto *f : fields) {
            size_t q = line.find(' ', p);
            if (q == std::string::npos) return false;
            *f = line.substr(p, q - p);
            p = q + 1;
        }
        pe.d = line.substr(p);
        out.push_back(pe);
    }
    return true;
}

// make_tiger (lib.rs:286-328) generalised over the path list and the scale.
int encode_pathlist(EncoderImpl &e, const std::vector<PathListEntry> &paths, double scale, uint32_t options = 0) {
    const bool compound = (options & PM_SCENE_OPT_COMPOUND_FILLS) != 0;
    const uint32_t fill_flags = (options & PM_SCENE_OPT_EVEN_ODD) ? PM_FILL_EVEN_ODD : PM_FILL_NONZERO;
    struct Flat { std::vector<std::vector<Pt>> sub; bool ok; };
    std::vector<Flat> flats(paths.size());
    size_t n_items = 0;
    for (size_t i = 0; i < paths.size(); i++) {
        BezPath bp;
        flats[i].ok = parse_svg_path(paths[i].d.c_str(), bp);  // `if let Ok(ref bp)`, :296,:313
        if (!flats[i].ok) continue;
        scale_path(bp, scale);
        flatten_path(bp, TOLERANCE, flats[i].sub);
        if (paths[i].fill != "-") n_items += compound ? (flats[i].sub.empty() ? 0 : 1) : flats[i].sub.size();    // count_fill_items   :332-335
        if (paths[i].stroke != "-") n_items += flats[i].sub.size();  // count_stroke_items :337-340
    }
    enc_begin_group(e, n_items);
    for (size_t i = 0; i < paths.size(); i++) {
        if (!flats[i].ok) continue;
        if (paths[i].fill != "-") {  // encode_path :342-347
            uint32_t rgba = parse_color(paths[i].fill.c_str());
            if (compound) { if (!flats[i].sub.empty()) enc_fill_subpaths(e, flats[i].sub, rgba, fill_flags); }
            else for (auto &sp : flats[i].sub) enc_fill(e, sp.data(), sp.size(), rgba);
        }
        if (paths[i].stroke != "-") {  // :318-323 + encode_path_stroke :353-367
            if (paths[i].width == "-") return PM_ERR_PARSE;  // .unwrap() on stroke-width, :319
            float width = strtof(paths[i].width.c_str(), nullptr) * (float)scale;
            uint32_t rgba = parse_color(paths[i].stroke.c_str());
            if (width < THIN_LINE) {
                float alpha = (float)(rgba & 0xff);
                alpha = alpha * std::sqrt(width / THIN_LINE);
                rgba = (rgba & ~0xffu) | (uint32_t)alpha;
                width = THIN_LINE;
            }
            for (auto &sp : flats[i].sub) enc_polyline(e, sp.data(), sp.size(), rgba, width);
        }
    }
    return e.group_ix == e.group_count ? PM_OK : PM_ERR_STATE;  // end_group assert, :147
}

// ---------------------------------------------------------------------------------------------
// Deterministic synthetic scenes (BASELINE.json configs 4 and 5)
// ---------------------------------------------------------------------------------------------
struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double range(double a, double b) { return a + (b - a) * uni(); }
    uint32_t below(uint32_t n) { return (uint32_t)(uni() * n); }
};

int build_rand_bezier(EncoderImpl &e, uint32_t W, uint32_t H, uint32_t count, uint64_t seed) {
    const double PI = 3.14159265358979323846;
    SplitMix64 rng{seed};
    double rs = (double)W / 8192.0;  // radii are specified at 8192^2 and scale with the surface
    enc_begin_group(e, count);
    std::vector<std::vector<Pt>> sub;
    for (uint32_t i = 0; i < count; i++) {
        Pt c{rng.range(0, W), rng.range(0, H)};
        double r = std::exp(rng.range(std::log(8.0), std::log(512.0))) * rs;
        uint32_t k = 4 + rng.below(5);
        std::vector<Pt> anchors(k), tang(k);
        std::vector<double> arm(k);
        for (uint32_t j = 0; j < k; j++) {
            double th = 2.0 * PI * j / k;
            double rr = r * (1.0 + rng.range(-0.35, 0.35));
            anchors[j] = Pt{c.x + rr * std::cos(th), c.y + rr * std::sin(th)};
            tang[j] = Pt{-std::sin(th), std::cos(th)};
            arm[j] = r * (4.0 / 3.0) * std::tan(PI / (2.0 * k)) * (1.0 + rng.range(-0.25, 0.25));
        }
        BezPath bp;
        bp.push_back(PathEl{MOVE, anchors[0], {}, {}});
        for (uint32_t j = 0; j < k; j++) {
            uint32_t jn = (j + 1) % k;
            Pt c1{anchors[j].x + arm[j] * tang[j].x, anchors[j].y + arm[j] * tang[j].y};
            Pt c2{anchors[jn].x - arm[jn] * tang[jn].x, anchors[jn].y - arm[jn] * tang[jn].y};
            bp.push_back(PathEl{CURVE, c1, c2, anchors[jn]});
        }
        sub.clear();
        flatten_path(bp, TOLERANCE, sub);
        uint32_t rgb = (uint32_t)(rng.next() & 0xffffff);
        uint32_t a = rng.uni() < 0.75 ? 255u : 64u + rng.below(191);
        int st = enc_fill(e, sub[0].data(), sub[0].size(), (rgb << 8) | a);
        if (st != PM_OK) return st;
    }
    return PM_OK;
}

int build_glyphs(EncoderImpl &e, uint32_t W, uint32_t H, uint32_t count, uint64_t seed) {
    const double PI = 3.14159265358979323846;
    SplitMix64 rng{seed};
    double es = (double)W / 4096.0;  // em sizes are specified at 4096^2
    enc_begin_group(e, count);
    double pen_x = 0, line_y = 0, em = rng.range(8.0, 24.0) * es;
    std::vector<Pt> pts;
    for (uint32_t i = 0; i < count; i++) {
        if (pen_x + 0.6 * em > W) {  // wrap
            pen_x = 0;
            line_y += 1.2 * em;
            em = rng.range(8.0, 24.0) * es;
            if (line_y + 1.2 * em > H) line_y = 0;  // page full: start over (glyphs overlap)
        }
        // glyph box: 0.5 em wide, 0.8 em tall, sitting in the line
        double bx = pen_x + 0.05 * em, by = line_y + 0.2 * em, bw = 0.5 * em, bh = 0.8 * em;
        Pt c{bx + 0.5 * bw, by + 0.5 * bh};
        uint32_t corners = 4 + rng.below(6);
        pts.clear();
        for (uint32_t j = 0; j < corners; j++) {
            double th0 = 2.0 * PI * (j + rng.range(-0.3, 0.3)) / corners;
            double rad = rng.range(0.45, 1.0);
            Pt v{c.x + 0.5 * bw * rad * std::cos(th0), c.y + 0.5 * bh * rad * std::sin(th0)};
            if (rng.uni() < 0.5 || pts.empty()) {
                pts.push_back(v);  // sharp corner
            } else {
                // rounded corner: a quadratic from the previous vertex through v, flattened to 4 pieces
                Pt p0 = pts.back();
                double th1 = 2.0 * PI * (j + 0.5) / corners;
                Pt p2{c.x + 0.5 * bw * rad * std::cos(th1), c.y + 0.5 * bh * rad * std::sin(th1)};
                for (int q = 1; q <= 4; q++) {
                    double t = q / 4.0, mt = 1.0 - t;
                    pts.push_back(Pt{mt * mt * p0.x + 2 * mt * t * v.x + t * t * p2.x,
                                     mt * mt * p0.y + 2 * mt * t * v.y + t * t * p2.y});
                }
            }
        }
        int st = enc_fill(e, pts.data(), pts.size(), 0x000000ffu);
        if (st != PM_OK) return st;
        pen_x += 0.6 * em;
    }
    return PM_OK;
}

int build_cardioid(EncoderImpl &e, double scale) {  // make_cardioid, lib.rs:257-270
    const double PI = 3.14159265358979323846;
    if (scale == 0.0) scale = 1.0;
    const int n = 97;
    double dth = PI * 2.0 / (double)n;
    Pt center{1024.0 * scale, 768.0 * scale};
    double r = 750.0 * scale;
    enc_begin_group(e, (size_t)(n - 1) * 2);
    for (int i = 1; i < n; i++) {
        double a0 = (double)i * dth, a1 = (double)((i * 2) % n) * dth;
        Pt p0{center.x + std::cos(a0) * r, center.y + std::sin(a0) * r};
        Pt p1{center.x + std::cos(a1) * r, center.y + std::sin(a1) * r};
        enc_circle(e, p0, 8.0 * scale);
        enc_stroke_line(e, p0, p1, (float)(2.0 * scale), 0x000080e0u);
    }
    return PM_OK;
}

int build_scene(EncoderImpl &e, const pm_scene_desc &d) {
    switch (d.kind) {
        case PM_SCENE_RECT1: {
            Pt p[4] = {{d.rect[0], d.rect[1]}, {d.rect[2], d.rect[1]}, {d.rect[2], d.rect[3]}, {d.rect[0], d.rect[3]}};
            enc_begin_group(e, 1);
            return enc_fill(e, p, 4, d.rgba ? d.rgba : 0x3366ccffu);
        }
        case PM_SCENE_PATH_TEST: {  // make_path_test, lib.rs:273-284
            Pt p[3] = {{10.0, 10.0}, {15.0, 800.0}, {300.0, 500.0}};
            enc_begin_group(e, 1);
            return enc_fill(e, p, 3, 0x80e0u);
        }
        case PM_SCENE_CARDIOID:
            return build_cardioid(e, d.scale);
        case PM_SCENE_TIGER: {
            std::vector<PathListEntry> paths;
            if (!parse_pathlist(pm_tiger_pathlist_begin, (size_t)(pm_tiger_pathlist_end - pm_tiger_pathlist_begin), paths))
                return PM_ERR_PARSE;
            double scale = d.scale != 0.0 ? d.scale : (double)d.width / 200.0;  // viewBox 0 0 200 200
            return encode_pathlist(e, paths, scale, d.options);
        }
        case PM_SCENE_RAND_BEZIER:
            return build_rand_bezier(e, d.width, d.height, d.count ? d.count : 10000u, d.seed ? d.seed : 0x5EED0004ull);
        case PM_SCENE_GLYPHS:
            return build_glyphs(e, d.width, d.height, d.count ? d.count : 100000u, d.seed ? d.seed : 0x5EED0005ull);
        default:
            return PM_ERR_INVALID_ARG;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
struct pm_encoder { EncoderImpl impl; };

extern "C" {

const char *pm_version(void) { return "0.1.0"; }

const char *pm_strerror(int status) {
    switch (status) {
        case PM_OK: return "ok";
        case PM_ERR_INVALID_ARG: return "invalid argument";
        case PM_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case PM_ERR_CUDA: return "CUDA call failed";
        case PM_ERR_SCENE_MALFORMED: return "scene buffer malformed (ref or count out of bounds)";
        case PM_ERR_BUFFER_TOO_SMALL: return "buffer too small";
        case PM_ERR_STATE: return "call out of order";
        case PM_ERR_PARSE: return "parse error";
        case PM_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

int pm_encoder_new(pm_encoder **out, uint8_t *buf, size_t cap) {
    if (!out) return PM_ERR_INVALID_ARG;
    pm_encoder *e = new (std::nothrow) pm_encoder();
    if (!e) return PM_ERR_NOMEM;
    e->impl.buf = buf;
    e->impl.cap = buf ? cap : 0;
    *out = e;
    return PM_OK;
}
static int enc_status(pm_encoder *e, int st) {
    if (st != PM_OK) return st;
    return e->impl.overflow ? PM_ERR_BUFFER_TOO_SMALL : PM_OK;
}
int pm_encoder_begin_group(pm_encoder *e, uint32_t n_items) {
    if (!e) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_begin_group(e->impl, n_items));
}
int pm_encoder_end_group(pm_encoder *e) {
    if (!e) return PM_ERR_INVALID_ARG;
    if (e->impl.group_ix != e->impl.group_count) return PM_ERR_STATE;  // assert_eq!, lib.rs:147
    e->impl.in_group = false;
    return enc_status(e, PM_OK);
}
int pm_encoder_circle(pm_encoder *e, double cx, double cy, double r) {
    if (!e) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_circle(e->impl, Pt{cx, cy}, r));
}
int pm_encoder_stroke_line(pm_encoder *e, double x0, double y0, double x1, double y1, float width, uint32_t rgba) {
    if (!e) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_stroke_line(e->impl, Pt{x0, y0}, Pt{x1, y1}, width, rgba));
}
int pm_encoder_fill(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba) {
    if (!e || !xy || n_points == 0) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_fill(e->impl, reinterpret_cast<const Pt *>(xy), n_points, rgba));
}
int pm_encoder_fill_rule(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba, uint32_t flags) {
    if (!e || !xy || n_points == 0) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_fill(e->impl, reinterpret_cast<const Pt *>(xy), n_points, rgba, flags));
}
int pm_encoder_fill_subpaths(pm_encoder *e, const double *xy, const uint32_t *counts, uint32_t n_subpaths, uint32_t rgba, uint32_t flags) {
    if (!e || !xy || !counts || n_subpaths == 0) return PM_ERR_INVALID_ARG;
    std::vector<std::vector<Pt>> sub(n_subpaths);
    const Pt *p = reinterpret_cast<const Pt *>(xy);
    for (uint32_t i = 0; i < n_subpaths; i++) {
        sub[i].assign(p, p + counts[i]);
        p += counts[i];
    }
    return enc_status(e, enc_fill_subpaths(e->impl, sub, rgba, flags));
}
int pm_encoder_polyline(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba, float width) {
    if (!e || !xy || n_points == 0) return PM_ERR_INVALID_ARG;
    return enc_status(e, enc_polyline(e->impl, reinterpret_cast<const Pt *>(xy), n_points, rgba, width));
}
size_t pm_encoder_bytes(const pm_encoder *e) { return e ? e->impl.free_space : 0; }
void pm_encoder_free(pm_encoder *e) { delete e; }

int64_t pm_flatten_svg_path(const char *d, double scale, double tolerance, double *out_xy, size_t cap_points,
                            uint32_t *out_counts, size_t cap_subpaths, size_t *need_points) {
    if (!d) return PM_ERR_INVALID_ARG;
    BezPath bp;
    if (!parse_svg_path(d, bp)) return PM_ERR_PARSE;
    scale_path(bp, scale);
    std::vector<std::vector<Pt>> sub;
    flatten_path(bp, tolerance, sub);
    size_t total = 0;
    for (auto &s : sub) total += s.size();
    if (need_points) *need_points = total;
    if (total > cap_points || sub.size() > cap_subpaths) return PM_ERR_BUFFER_TOO_SMALL;
    size_t k = 0;
    for (size_t i = 0; i < sub.size(); i++) {
        if (out_counts) out_counts[i] = (uint32_t)sub[i].size();
        for (auto &p : sub[i]) {
            if (out_xy) { out_xy[2 * k] = p.x; out_xy[2 * k + 1] = p.y; }
            k++;
        }
    }
    return (int64_t)sub.size();
}

uint32_t pm_parse_color(const char *s) { return (s && s[0]) ? parse_color(s) : 0xff00ff80u; }

int64_t pm_scene_build(const pm_scene_desc *desc, uint8_t *buf, size_t cap) {
    if (!desc) return PM_ERR_INVALID_ARG;
    EncoderImpl e;
    e.buf = buf;
    e.cap = buf ? cap : 0;
    int st = build_scene(e, *desc);
    if (st != PM_OK) return st;
    if (buf && e.overflow) return PM_ERR_BUFFER_TOO_SMALL;
    return (int64_t)e.free_space;
}

int64_t pm_scene_from_pathlist(const char *text, size_t len, double scale, uint8_t *buf, size_t cap) {
    if (!text) return PM_ERR_INVALID_ARG;
    std::vector<PathListEntry> paths;
    if (!parse_pathlist(text, len, paths)) return PM_ERR_PARSE;
    EncoderImpl e;
    e.buf = buf;
    e.cap = buf ? cap : 0;
    int st = encode_pathlist(e, paths, scale);
    if (st != PM_OK) return st;
    if (buf && e.overflow) return PM_ERR_BUFFER_TOO_SMALL;
    return (int64_t)e.free_space;
}

void init_test_scene(uint8_t *buf, ssize_t buf_size) {  // lib.rs:387-393; make_test_scene -> make_tiger, scale 8
    if (!buf || buf_size < (ssize_t)PM_GROUP_HEADER_SIZE) return;
    pm_scene_desc d;
    memset(&d, 0, sizeof d);
    d.kind = PM_SCENE_TIGER;
    d.scale = 8.0;  // lib.rs:287
    if (pm_scene_build(&d, buf, (size_t)buf_size) < 0) {
        pm_group_header empty = {0, PM_GROUP_HEADER_SIZE};
        memcpy(buf, &empty, sizeof empty);
    }
}

int pm_scene_validate(const uint8_t *scene, size_t len) {
    if (!scene) return PM_ERR_INVALID_ARG;
    if (len < PM_GROUP_HEADER_SIZE || len > 0xffffffffull) return PM_ERR_SCENE_MALFORMED;
    pm_group_header g;
    memcpy(&g, scene, sizeof g);
    uint64_t n = g.n_items;
    if (PM_GROUP_HEADER_SIZE + n * PM_BBOX_SIZE > len) return PM_ERR_SCENE_MALFORMED;
    if ((g.items_ix & 7u) || (uint64_t)g.items_ix + n * PM_ITEM_SIZE > len) return PM_ERR_SCENE_MALFORMED;
    for (uint64_t i = 0; i < n; i++) {
        pm_item_any it;
        memcpy(&it, scene + g.items_ix + i * PM_ITEM_SIZE, sizeof it);
        if (it.tag == PM_ITEM_FILL || it.tag == PM_ITEM_POLY) {
            uint32_t np = it.body[2], pix = it.body[3];  // n_points @12, points_ix @16 in both variants
            if (np == 0 || (pix & 7u) || (uint64_t)pix + (uint64_t)np * 8 > len) return PM_ERR_SCENE_MALFORMED;
        }
    }
    return PM_OK;
}

// ---- row-strip shard: cost estimate per tile row and balanced contiguous partition --------------
// The frame's tile rows are independent given the scene (TestApp/PietRender.metal:167-170, :463-466),
// so the multi-GPU shard is a partition of the rows; equal-height strips leave the GPUs that own
// the middle of a centred drawing with most of the work.  cost[r] models one GPU's time for tile
// row r: a term per tile (the framebuffer store) and a term per (segment, tile) crossing (binning
// and coverage work), in the ratio measured on the 8192^2 and 16384^2 tiger.
#define PM_COST_PER_TILE 0.25f
#define PM_COST_PER_CROSSING 0.32f

int pm_scene_row_costs(const uint8_t *scene, size_t len, uint32_t width, uint32_t height, float *cost, size_t n_rows) {
    if (!scene || !cost || width == 0 || height == 0) return PM_ERR_INVALID_ARG;
    const int st = pm_scene_validate(scene, len);
    if (st != PM_OK) return st;
    const uint32_t n_tx = (width + PM_TILE_W - 1) / PM_TILE_W, n_ty = (height + PM_TILE_H - 1) / PM_TILE_H;
    if (n_rows < n_ty) return PM_ERR_BUFFER_TOO_SMALL;
    std::vector<double> acc(n_ty, 0.0);
    pm_group_header g;
    memcpy(&g, scene, sizeof g);
    auto add_segment = [&](double sx, double sy, double ex, double ey, double hw) {
        if (!(std::isfinite(sx) && std::isfinite(sy) && std::isfinite(ex) && std::isfinite(ey))) return;
        const double y_lo = std::min(sy, ey) - hw, y_hi = std::max(sy, ey) + hw;
        long r0 = (long)std::floor(y_lo / PM_TILE_H), r1 = (long)std::floor(y_hi / PM_TILE_H);
        if (r1 < 0 || r0 >= (long)n_ty) return;
        r0 = std::max(r0, 0L);
        r1 = std::min(r1, (long)n_ty - 1);
        const double dy = ey - sy;
        for (long r = r0; r <= r1; r++) {
            // x extent of the segment inside the band of this tile row
            double xa = sx, xb = ex;
            if (dy != 0.0) {
                const double ya = std::max((double)r * PM_TILE_H - hw, std::min(sy, ey)), yb = std::min((double)(r + 1) * PM_TILE_H + hw, std::max(sy, ey));
                xa = sx + (ya - sy) * (ex - sx) / dy;
                xb = sx + (yb - sy) * (ex - sx) / dy;
            }
            double x_lo = std::max(std::min(xa, xb) - hw, 0.0), x_hi = std::min(std::max(xa, xb) + hw, (double)n_tx * PM_TILE_W);
            if (x_hi < x_lo) continue;
            acc[(size_t)r] += std::floor(x_hi / PM_TILE_W) - std::floor(x_lo / PM_TILE_W) + 1.0;
        }
    };
    for (uint64_t i = 0; i < g.n_items; i++) {
        pm_item_any it;
        memcpy(&it, scene + g.items_ix + i * PM_ITEM_SIZE, sizeof it);
        if (it.tag == PM_ITEM_FILL || it.tag == PM_ITEM_POLY) {
            const uint32_t np = it.body[2], pix = it.body[3];
            const float *pts = reinterpret_cast<const float *>(scene + pix);
            float w = 0.0f;
            if (it.tag == PM_ITEM_POLY) memcpy(&w, &it.body[1], 4);
            const double hw = it.tag == PM_ITEM_POLY ? 0.5 * w + 0.5 : 0.0;
            const uint32_t n_seg = it.tag == PM_ITEM_FILL ? np : np - 1;
            for (uint32_t k = 0; k < n_seg; k++) {
                const uint32_t k1 = k + 1 == np ? 0 : k + 1;
                add_segment(pts[2 * k], pts[2 * k + 1], pts[2 * k1], pts[2 * k1 + 1], hw);
            }
        } else if (it.tag == PM_ITEM_LINE) {
            pm_item_line ln;
            memcpy(&ln, &it, sizeof ln);
            add_segment(ln.sx, ln.sy, ln.ex, ln.ey, 0.5 * ln.width + 0.5);
        } else if (it.tag == PM_ITEM_CIRCLE) {
            pm_bbox bb;
            memcpy(&bb, scene + PM_GROUP_HEADER_SIZE + i * PM_BBOX_SIZE, sizeof bb);
            for (uint32_t r = bb.y0 / PM_TILE_H; r <= bb.y1 / PM_TILE_H && r < n_ty; r++) acc[r] += (double)(bb.x1 / PM_TILE_W - bb.x0 / PM_TILE_W + 1);
        }
    }
    for (uint32_t r = 0; r < n_ty; r++) cost[r] = PM_COST_PER_TILE * (float)n_tx + PM_COST_PER_CROSSING * (float)acc[r];
    return PM_OK;
}

int pm_balance_strips(const float *cost, uint32_t n_rows, uint32_t n_parts, uint32_t *bounds) {
    if (!cost || !bounds || n_parts == 0 || n_rows == 0) return PM_ERR_INVALID_ARG;
    if (n_parts > n_rows) return PM_ERR_INVALID_ARG;  // every strip owns at least one tile row
    std::vector<double> pre(n_rows + 1, 0.0);
    double mx = 0.0;
    for (uint32_t r = 0; r < n_rows; r++) {
        const double c = cost[r] > 0.0f && std::isfinite(cost[r]) ? cost[r] : 0.0;
        pre[r + 1] = pre[r] + c;
        mx = std::max(mx, c);
    }
    // smallest capacity for which a greedy left-to-right packing needs at most n_parts strips
    auto parts_needed = [&](double cap, std::vector<uint32_t> *cuts) {
        uint32_t parts = 0, r = 0;
        while (r < n_rows) {
            // furthest end such that the strip [r, end) fits, leaving at least one row for each remaining strip
            uint32_t end = (uint32_t)(std::upper_bound(pre.begin() + r + 1, pre.end(), pre[r] + cap) - pre.begin()) - 1;
            if (end <= r) end = r + 1;
            if (cuts) {
                const uint32_t remaining = n_parts - 1 - std::min(parts, n_parts - 1);
                if (end > n_rows - remaining) end = n_rows - remaining;
                if (end <= r) end = r + 1;
                cuts->push_back(end);
            }
            r = end;
            parts++;
        }
        return parts;
    };
    double lo = mx, hi = pre[n_rows];
    for (int it = 0; it < 60 && hi - lo > 1e-9 * std::max(1.0, hi); it++) {
        const double mid = 0.5 * (lo + hi);
        if (parts_needed(mid, nullptr) <= n_parts) hi = mid; else lo = mid;
    }
    std::vector<uint32_t> cuts;
    parts_needed(hi * (1.0 + 1e-9), &cuts);
    bounds[0] = 0;
    for (uint32_t g = 1; g <= n_parts; g++) {
        uint32_t b = g - 1 < cuts.size() ? cuts[g - 1] : n_rows;
        if (b < bounds[g - 1] + 1) b = bounds[g - 1] + 1;       // non-empty
        const uint32_t max_b = n_rows - (n_parts - g);              // room for the strips that follow
        if (b > max_b) b = max_b;
        bounds[g] = b;
    }
    bounds[n_parts] = n_rows;
    return PM_OK;
}

// ---- framebuffer egress (host only): the step after the hot path -------------
// The reference hands its texture to MTKView; a headless renderer needs a file.  PPM (P6, alpha dropped) and
// PNG (8-bit RGBA, stored deflate blocks: no compression library needed).
static uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
static void put_be32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void png_chunk(std::vector<uint8_t> &out, const char *type, const std::vector<uint8_t> &data) {
    put_be32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put_be32(out, crc32_update(0xffffffffu, out.data() + start, out.size() - start) ^ 0xffffffffu);
}

int pm_write_ppm(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height, size_t stride) {
    if (!path || !rgba8 || width == 0 || height == 0 || stride < (size_t)width * 4) return PM_ERR_INVALID_ARG;
    FILE *f = fopen(path, "wb");
    if (!f) return PM_ERR_INVALID_ARG;
    fprintf(f, "P6\n%u %u\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (uint32_t y = 0; y < height; y++) {
        const uint8_t *src = rgba8 + (size_t)y * stride;
        for (uint32_t x = 0; x < width; x++) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
        if (fwrite(row.data(), 1, row.size(), f) != row.size()) { fclose(f); return PM_ERR_INVALID_ARG; }
    }
    return fclose(f) == 0 ? PM_OK : PM_ERR_INVALID_ARG;
}

int pm_write_png(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height, size_t stride) {
    if (!path || !rgba8 || width == 0 || height == 0 || stride < (size_t)width * 4) return PM_ERR_INVALID_ARG;
    // raw scanlines: filter byte 0 + RGBA
    const size_t line = (size_t)width * 4 + 1, raw_len = line * height;
    std::vector<uint8_t> idat;
    idat.reserve(raw_len + raw_len / 65535 * 5 + 16);
    idat.push_back(0x78); idat.push_back(0x01);  // zlib header, no compression
    uint32_t a = 1, b = 0;                        // Adler-32 of the raw data
    std::vector<uint8_t> raw(line);
    size_t block_left = 0, done = 0;
    for (uint32_t y = 0; y < height; y++) {
        raw[0] = 0;
        memcpy(raw.data() + 1, rgba8 + (size_t)y * stride, (size_t)width * 4);
        for (size_t i = 0; i < line; i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
        size_t off = 0;
        while (off < line) {
            if (block_left == 0) {  // new stored block of at most 65535 bytes
                const size_t remaining = raw_len - done;
                const size_t len = remaining < 65535 ? remaining : 65535;
                idat.push_back(len == remaining ? 1 : 0);  // BFINAL on the last one, BTYPE = 00
                idat.push_back(len & 0xff); idat.push_back(len >> 8);
                idat.push_back(~len & 0xff); idat.push_back((~len >> 8) & 0xff);
                block_left = len;
            }
            const size_t take = std::min(block_left, line - off);
            idat.insert(idat.end(), raw.begin() + off, raw.begin() + off + take);
            off += take; block_left -= take; done += take;
        }
    }
    put_be32(idat, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // 8-bit RGBA, no interlace
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", idat);
    png_chunk(out, "IEND", std::vector<uint8_t>());
    FILE *f = fopen(path, "wb");
    if (!f) return PM_ERR_INVALID_ARG;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    return (fclose(f) == 0 && ok) ? PM_OK : PM_ERR_INVALID_ARG;
}

}  // extern "C"
