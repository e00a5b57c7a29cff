// pm_context_*: a piet-style RenderContext in front of the scene format.
//
// The reference's README calls itself "an experimental Metal backend for the piet 2D graphics API" (README.md:3)
// but stops at the proto-API of its `Encoder` (src/lib.rs:165-222: circle, stroke_line, fill, polyline) driven by
// make_tiger.  This is the missing front: the calls a piet RenderContext takes -- clear, transform, save / restore,
// fill, fill_even_odd, stroke with a solid brush -- recorded as path control points, and finish(), which hands them
// to the renderer as a pm_path_set: flattening and encoding then run on the device (pm_flatten.cu).  Host only, no CUDA
// in this file.  What make_tiger does per <path> is what fill / stroke do here: Affine * BezPath (lib.rs:297, :314),
// the thin-stroke rule (lib.rs:353-362), one item per stroked subpath; a filled path becomes ONE item whose subpaths
// are joined by zero-area bridges (pm_encoder_fill_subpaths explains why that is exact), so holes are holes.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/piet_metal_b200.h"
#include "pm_scene_format.h"

namespace {

struct Affine {
    double a = 1, b = 0, c = 0, d = 1, e = 0, f = 0;  // x' = a x + c y + e, y' = b x + d y + f (kurbo's [a b c d e f])
    void apply(double x, double y, double *ox, double *oy) const { *ox = a * x + c * y + e; *oy = b * x + d * y + f; }
    Affine then(const Affine &m) const {  // this * m: m is applied first (RenderContext::transform post-multiplies)
        Affine r;
        r.a = a * m.a + c * m.b; r.b = b * m.a + d * m.b;
        r.c = a * m.c + c * m.d; r.d = b * m.c + d * m.d;
        r.e = a * m.e + c * m.f + e; r.f = b * m.e + d * m.f + f;
        return r;
    }
    double scale() const { return std::sqrt(std::fabs(a * d - b * c)); }
};

const float THIN_LINE = 0.7f;  // src/lib.rs:351

}  // namespace

struct pm_context {
    pm_renderer *renderer = nullptr;
    uint32_t width = 0, height = 0;
    Affine ctm;
    std::vector<Affine> stack;
    // the path set being recorded
    std::vector<uint32_t> first{0}, tag, rgba, flags;
    std::vector<double> start, ctrl;
    std::vector<uint8_t> verb;
    std::vector<float> stroke_width;
    int status = PM_OK;

    void begin(double x, double y, uint32_t t, uint32_t color, float w, uint32_t fl) {
        start.push_back(x); start.push_back(y);
        tag.push_back(t); rgba.push_back(color); stroke_width.push_back(w); flags.push_back(fl);
        first.push_back(first.back());
    }
    void line(double x, double y) {
        verb.push_back(PM_VERB_LINE);
        const double z[6] = {0, 0, 0, 0, x, y};
        ctrl.insert(ctrl.end(), z, z + 6);
        first.back()++;
    }
    void curve(double x1, double y1, double x2, double y2, double x3, double y3) {
        verb.push_back(PM_VERB_CURVE);
        const double z[6] = {x1, y1, x2, y2, x3, y3};
        ctrl.insert(ctrl.end(), z, z + 6);
        first.back()++;
    }
};

namespace {

// One path in device space: subpaths of (verb, points), quads raised to cubics, ClosePath dropped (the fill kernel
// closes every point list itself, metal:262; stroked subpaths stay open as in the reference, flatten.rs:40).
struct Sub { double sx, sy; std::vector<uint8_t> verb; std::vector<double> pts; };

bool collect(const pm_context *c, const pm_path_el *els, size_t n, std::vector<Sub> &out) {
    double cx = 0, cy = 0;
    for (size_t i = 0; i < n; i++) {
        const pm_path_el &el = els[i];
        double p[6];
        for (int k = 0; k < 3; k++) c->ctm.apply(el.x[2 * k], el.x[2 * k + 1], &p[2 * k], &p[2 * k + 1]);
        switch (el.verb) {
            case PM_EL_MOVE:
                out.push_back(Sub{p[0], p[1], {}, {}});
                cx = p[0]; cy = p[1];
                break;
            case PM_EL_LINE:
                if (out.empty()) return false;
                out.back().verb.push_back(PM_VERB_LINE);
                { const double z[6] = {0, 0, 0, 0, p[0], p[1]}; out.back().pts.insert(out.back().pts.end(), z, z + 6); }
                cx = p[0]; cy = p[1];
                break;
            case PM_EL_QUAD: {  // exact degree elevation: c1 = p0 + 2/3 (q - p0), c2 = p2 + 2/3 (q - p2)
                if (out.empty()) return false;
                const double z[6] = {cx + (2.0 / 3.0) * (p[0] - cx), cy + (2.0 / 3.0) * (p[1] - cy),
                                     p[2] + (2.0 / 3.0) * (p[0] - p[2]), p[3] + (2.0 / 3.0) * (p[1] - p[3]), p[2], p[3]};
                out.back().verb.push_back(PM_VERB_CURVE);
                out.back().pts.insert(out.back().pts.end(), z, z + 6);
                cx = p[2]; cy = p[3];
                break;
            }
            case PM_EL_CURVE:
                if (out.empty()) return false;
                out.back().verb.push_back(PM_VERB_CURVE);
                out.back().pts.insert(out.back().pts.end(), p, p + 6);
                cx = p[4]; cy = p[5];
                break;
            case PM_EL_CLOSE:
                if (!out.empty()) { cx = out.back().sx; cy = out.back().sy; }
                break;
            default:
                return false;
        }
    }
    return true;
}

int do_fill(pm_context *c, const pm_path_el *els, size_t n, uint32_t color, uint32_t fl) {
    if (!c || (!els && n)) return PM_ERR_INVALID_ARG;
    std::vector<Sub> subs;
    if (!collect(c, els, n, subs)) return PM_ERR_INVALID_ARG;
    if (subs.empty()) return PM_OK;
    // one item: subpath 0, then every further subpath reached by a bridge, closed, and left by the bridge back
    c->begin(subs[0].sx, subs[0].sy, PM_ITEM_FILL, color, 0.0f, fl);
    for (size_t s = 0; s < subs.size(); s++) {
        const Sub &sp = subs[s];
        if (s > 0) c->line(sp.sx, sp.sy);
        for (size_t k = 0; k < sp.verb.size(); k++) {
            const double *p = &sp.pts[6 * k];
            if (sp.verb[k] == PM_VERB_LINE) c->line(p[4], p[5]); else c->curve(p[0], p[1], p[2], p[3], p[4], p[5]);
        }
        if (subs.size() > 1) {
            c->line(sp.sx, sp.sy);
            if (s > 0) c->line(subs[0].sx, subs[0].sy);
        }
    }
    return PM_OK;
}

}  // namespace

extern "C" {

int pm_context_new(pm_context **out, pm_renderer *renderer, uint32_t width, uint32_t height) {
    if (!out || width == 0 || height == 0) return PM_ERR_INVALID_ARG;
    pm_context *c = new (std::nothrow) pm_context();
    if (!c) return PM_ERR_NOMEM;
    c->renderer = renderer;
    c->width = width; c->height = height;
    *out = c;
    return PM_OK;
}

void pm_context_free(pm_context *c) { delete c; }

int pm_context_save(pm_context *c) { if (!c) return PM_ERR_INVALID_ARG; c->stack.push_back(c->ctm); return PM_OK; }
int pm_context_restore(pm_context *c) {
    if (!c) return PM_ERR_INVALID_ARG;
    if (c->stack.empty()) return PM_ERR_STATE;
    c->ctm = c->stack.back();
    c->stack.pop_back();
    return PM_OK;
}
int pm_context_transform(pm_context *c, const double m[6]) {
    if (!c || !m) return PM_ERR_INVALID_ARG;
    Affine t;
    t.a = m[0]; t.b = m[1]; t.c = m[2]; t.d = m[3]; t.e = m[4]; t.f = m[5];
    c->ctm = c->ctm.then(t);
    return PM_OK;
}

int pm_context_clear(pm_context *c, uint32_t color) {  // RenderContext::clear: the whole surface, whatever the transform
    if (!c) return PM_ERR_INVALID_ARG;
    const double w = c->width, h = c->height;
    // (a little beyond the surface and slightly tilted: an exactly horizontal edge that crosses a tile boundary is lost by
    // the reference's left-edge split, metal:336-338)
    c->begin(-8.0, -8.25, PM_ITEM_FILL, color, 0.0f, 0);
    c->line(w + 8.0, -8.0);
    c->line(w + 8.25, h + 8.0);
    c->line(-8.0, h + 8.25);
    return PM_OK;
}

int pm_context_fill(pm_context *c, const pm_path_el *els, size_t n, uint32_t color) { return do_fill(c, els, n, color, PM_FILL_NONZERO); }
int pm_context_fill_even_odd(pm_context *c, const pm_path_el *els, size_t n, uint32_t color) { return do_fill(c, els, n, color, PM_FILL_EVEN_ODD); }

int pm_context_stroke(pm_context *c, const pm_path_el *els, size_t n, uint32_t color, double width) {
    if (!c || (!els && n) || !(width >= 0.0)) return PM_ERR_INVALID_ARG;
    std::vector<Sub> subs;
    if (!collect(c, els, n, subs)) return PM_ERR_INVALID_ARG;
    float w = (float)(width * c->ctm.scale());
    if (w < THIN_LINE) {  // encode_path_stroke, lib.rs:353-362: thinner than 0.7 px -> 0.7 px with the alpha scaled down
        float alpha = (float)(color & 0xff);
        alpha = alpha * std::sqrt(w / THIN_LINE);
        color = (color & ~0xffu) | (uint32_t)alpha;
        w = THIN_LINE;
    }
    for (const Sub &sp : subs) {  // one PietStrokePolyLine per subpath (lib.rs:209-222, :364-366)
        c->begin(sp.sx, sp.sy, PM_ITEM_POLY, color, w, 0);
        for (size_t k = 0; k < sp.verb.size(); k++) {
            const double *p = &sp.pts[6 * k];
            if (sp.verb[k] == PM_VERB_LINE) c->line(p[4], p[5]); else c->curve(p[0], p[1], p[2], p[3], p[4], p[5]);
        }
    }
    return PM_OK;
}

uint32_t pm_context_item_count(const pm_context *c) { return c ? (uint32_t)c->tag.size() : 0; }

// The recorded drawing as a pm_path_set (pointers into the context: valid until the next call on it).
int pm_context_path_set(pm_context *c, pm_path_set *out) {
    if (!c || !out) return PM_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    out->n_subpaths = (uint32_t)c->tag.size();
    out->n_segments = (uint32_t)c->verb.size();
    out->first_segment = c->first.data();
    out->start = c->start.data();
    out->verb = c->verb.data();
    out->ctrl = c->ctrl.data();
    out->tag = c->tag.data();
    out->rgba = c->rgba.data();
    out->width = c->stroke_width.data();
    out->flags = c->flags.data();
    return PM_OK;
}

// RenderContext::finish: hand the drawing to the renderer (flattened and encoded on the device) and start a new one.
// The renderer must have been created with PM_FLAG_FILL_RULES for fill_even_odd to be honoured.
int pm_context_finish(pm_context *c, double tolerance) {
    if (!c) return PM_ERR_INVALID_ARG;
    if (!c->renderer) return PM_ERR_STATE;
    if (c->tag.empty()) pm_context_clear(c, 0xffffffffu);  // an empty drawing is the white background (metal:74, :470)
    pm_path_set ps;
    pm_context_path_set(c, &ps);
    const int st = pm_renderer_set_scene_paths(c->renderer, &ps, 1.0, tolerance > 0.0 ? tolerance : 0.1);
    c->first.assign(1, 0);
    c->tag.clear(); c->rgba.clear(); c->flags.clear(); c->start.clear(); c->ctrl.clear(); c->verb.clear(); c->stroke_width.clear();
    return st;
}

}  // extern "C"
