// Tile-level geometry of the render path: which tiles an item's segments reach, and with what.
//
// These are the predicates of the reference's tileKernel (TestApp/PietRender.metal:160-454),
// factored so that the binning kernel can evaluate them once per (segment, tile row) instead of
// once per (tile, segment) as the reference does.  Every float expression keeps the reference's
// operand order and is compiled without FMA contraction (-fmad=false / -ffp-contract=off): cull
// decisions are signs of `a*x + b*y + c`, and a fused multiply-add would flip knife-edge cases.
//
// The functions are __host__ __device__ so that tests/native/ can drive exactly this code on the
// CPU against the oracle; nothing here is a CPU rendering path of the library.
//
// Exactness of the restructuring (DESIGN.md "why the binned path is exact"):
//  * emission of a Fill/Line command to a tile always carries an x-overlap guard in the reference
//    (metal:334, :350-351, :416-417), so only tiles overlapping the segment's x range need testing;
//  * the backdrop contribution `backdrop -= s00` (metal:331-333) fires for tiles whose top-left
//    corner is on the sign(a) side of the segment; with ymin <= y0 the expression for s00 is
//    bit-identical to the one for sTopLeft, so the contribution is -sign(a); and because every
//    IEEE operation in  fl(fl(fl(a*x0) + fl(y0*b)) + c)  is monotone in x0, the set of such tiles
//    is a suffix of the tile row, found by bisection with the exact expression;
//  * the 256-px strip pre-cull (metal:257-301) is a function of (segment, row, strip) only and is
//    evaluated exactly for every strip that can matter.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pm_pixel_logic.h"
#include "pm_scene_format.h"

struct PmSeg {
    float sx, sy, ex, ey;      // start, end (scene coordinates, pixels)
    float mnx, mny, mxx, mxy;  // component-wise min / max (metal:263-264)
    float a, b, c;             // a*x + b*y + c = 0 (metal:267-269)
};

PM_HD float pm_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }  // MSL sign(): sign(0) = 0 (and sign(NaN) = 0 here)
// sign(u) == sign(v) without materialising the signs (two compares per side; in the bisection loops one side is invariant)
PM_HD bool pm_same_sign(float u, float v) {
    const bool up = u > 0.0f, un = u < 0.0f, vp = v > 0.0f, vn = v < 0.0f;
    return (up && vp) || (un && vn) || (!up && !un && !vp && !vn);
}

PM_HD PmSeg pm_seg(float sx, float sy, float ex, float ey) {
    PmSeg g;
    g.sx = sx; g.sy = sy; g.ex = ex; g.ey = ey;
    g.mnx = fminf(sx, ex); g.mny = fminf(sy, ey);
    g.mxx = fmaxf(sx, ex); g.mxy = fmaxf(sy, ey);
    g.a = ey - sy;
    g.b = sx - ex;
    g.c = -(g.a * sx + g.b * sy);
    return g;
}

// "If all four corners are on same side of line, cull" (metal:237-241 and every copy of it).
PM_HD bool pm_cross4(float s00, float s01, float s10, float s11) { return s00 * s01 + s00 * s10 + s00 * s11 < 3.0f; }
// The same test on the four corner VALUES (v = a*x + b*y + c) instead of their signs: the products of signs sum to 3
// exactly when the four signs are equal and not zero, i.e. when the values are all positive or all negative
// (a NaN compares false both ways, like sign(NaN) = 0 above: "crosses").
PM_HD bool pm_cross4v(float v00, float v01, float v10, float v11) {
    const bool all_pos = v00 > 0.0f && v01 > 0.0f && v10 > 0.0f && v11 > 0.0f;
    const bool all_neg = v00 < 0.0f && v01 < 0.0f && v10 < 0.0f && v11 < 0.0f;
    return !(all_pos || all_neg);
}

// ---------------------------------------------------------------------------------------------
// Fill items (metal:248-365)
// ---------------------------------------------------------------------------------------------

// The y part of the strip pre-cull (metal:265) -- the only y test a fill segment ever gets
// ("no y-based cull here because it's been done in the earlier pass", metal:312-313).
PM_HD bool pm_fill_row_overlap(const PmSeg &g, float y0) { return g.mxy >= y0 && g.mny < y0 + 16.0f; }

// Vote of the strip pre-cull (metal:265-293) for the 256-px strip starting at sx0 in the tile row
// starting at y0.  The caller has already applied pm_fill_row_overlap.
PM_HD bool pm_fill_strip_vote(const PmSeg &g, float y0, float sx0) {
    if (!(g.mnx < sx0 + 256.0f)) return false;
    float left = g.a * sx0;
    float right = g.a * (sx0 + 256.0f);
    float ytop = fmaxf(y0, g.mny);
    float ybot = fminf(y0 + 16.0f, g.mxy);
    float top = g.b * ytop;
    float bot = g.b * ybot;
    float v_top_left = right - g.a * 16.0f + y0 * g.b + g.c;  // top left of rightmost tile in strip
    bool fill_hit = false;
    if (pm_same_sign(v_top_left, g.a) && g.mny <= y0) fill_hit = true;   // left ray intersects, need backdrop
    if (pm_cross4v(top + left + g.c, top + right + g.c, bot + left + g.c, bot + right + g.c) && g.mxx > sx0) fill_hit = true; // intersects strip
    return fill_hit;
}

// sTopLeft == sign(a) for the tile whose left edge is x0 (metal:326, :331).
PM_HD bool pm_fill_backdrop_side(const PmSeg &g, float x0, float y0) {
    return pm_same_sign(g.a * x0 + y0 * g.b + g.c, g.a);
}

// First tile index in [0, n_tiles_x] whose top-left corner is on the sign(a) side; n_tiles_x if
// none.  Valid because the predicate is monotone in x0 (see the header comment).  Needs a != 0.
PM_HD uint32_t pm_fill_backdrop_first_tile(const PmSeg &g, float y0, uint32_t n_tiles_x) {
    uint32_t lo = 0, hi = n_tiles_x;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (pm_fill_backdrop_side(g, (float)(mid * PM_TILE_W), y0)) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// The "left ray intersects" half of the strip vote (metal:277, :282) without the ymin test:
// sTopLeft, taken at the top-left corner of the strip's rightmost tile, is on the sign(a) side.
PM_HD bool pm_fill_strip_side(const PmSeg &g, float y0, float sx0) {
    float right = g.a * (sx0 + 256.0f);
    return pm_same_sign(right - g.a * 16.0f + y0 * g.b + g.c, g.a);
}
// First strip index in [0, n_strips] for which pm_fill_strip_side holds (monotone in the strip
// index for the same reason as the tile test).  Needs a != 0.
PM_HD uint32_t pm_fill_strip_side_first(const PmSeg &g, float y0, uint32_t n_strips) {
    uint32_t lo = 0, hi = n_strips;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (pm_fill_strip_side(g, y0, (float)(mid * PM_STRIP_PX))) hi = mid; else lo = mid + 1;
    }
    return lo;
}

enum { PM_EMIT_NONE = 0, PM_EMIT_WHOLE = 1, PM_EMIT_EDGE = 2 };
struct PmFillEmit {
    int kind;      // PM_EMIT_*
    float s00;     // sign handed to FillEdge (metal:338)
    float y_edge;  // crossing of the tile's left edge (metal:335)
};

// The per-tile exact test (metal:308-353) for the tile at (x0, y0); emission part only.
PM_HD PmFillEmit pm_fill_tile_test(const PmSeg &g, float x0, float y0) {
    PmFillEmit r;
    r.kind = PM_EMIT_NONE; r.y_edge = 0.0f;
    float left = g.a * x0;
    float right = g.a * (x0 + 16.0f);
    float ytop = fmaxf(y0, g.mny);
    float ybot = fminf(y0 + 16.0f, g.mxy);
    float top = g.b * ytop;
    float bot = g.b * ybot;
    const float v00 = top + left + g.c;
    r.s00 = pm_sign(v00);
    bool crosses = pm_cross4v(v00, top + right + g.c, bot + left + g.c, bot + right + g.c);
    if (g.mnx < x0 && g.mxx > x0) {
        float y_edge = g.sy + (g.ey - g.sy) * ((g.sx - x0) / g.b);  // mix(start.y, end.y, (start.x - x0) / b)
        if (y_edge >= y0 && y_edge < y0 + 16.0f) {
            r.kind = PM_EMIT_EDGE;  // line intersects left edge of this tile
            r.y_edge = y_edge;
        } else if (crosses) {
            r.kind = PM_EMIT_WHOLE;
        }
    } else if (crosses && g.mnx < x0 + 16.0f && g.mxx > x0) {
        r.kind = PM_EMIT_WHOLE;
    }
    return r;
}

// Tile range [ta, tb] (clamped to [t_lo, t_hi]) that contains every tile with
// lo_x < x0 + 16 and hi_x > x0.  Returns false if the range is empty (also for NaN input).
PM_HD bool pm_tile_span(float lo_x, float hi_x, uint32_t t_lo, uint32_t t_hi, uint32_t *ta, uint32_t *tb) {
    if (!(lo_x < 65536.0f && hi_x > -16.0f)) return false;
    int ia = (int)floorf(fmaxf(lo_x, 0.0f) * (1.0f / 16.0f)) - 1;
    int ib = (int)floorf(fminf(hi_x, 65535.0f) * (1.0f / 16.0f)) + 1;
    if (ia < (int)t_lo) ia = (int)t_lo;
    if (ib > (int)t_hi) ib = (int)t_hi;
    if (ib < ia) return false;
    *ta = (uint32_t)ia;
    *tb = (uint32_t)ib;
    return true;
}

// x extent of the segment inside the horizontal band [ya, yb], widened by a margin that covers the
// rounding of the corner-sign expressions: their absolute error is below 4 eps (|a| X + |b| Y + |c|)
// with X, Y <= 65536, i.e. a corner farther than 0.033 (1 + |dx/dy|) px from the line has the sign
// exact arithmetic gives it; the margin used is 2 + |dx/dy| / 4 px.  A tile whose x range does not
// touch this extent cannot get a command from the segment in this band: all four corner signs
// agree and the left-edge crossing lies outside the band.
PM_HD void pm_seg_band_x(const PmSeg &g, float ya, float yb, float *lo, float *hi) {
    *lo = g.mnx;
    *hi = g.mxx;
    if (g.a == 0.0f) return;
    float yt = fmaxf(ya, g.mny), ybm = fminf(yb, g.mxy);
    float inv = (g.ex - g.sx) / g.a;  // dx / dy
    float x1 = g.sx + (yt - g.sy) * inv, x2 = g.sx + (ybm - g.sy) * inv;
    float m = 2.0f + 0.25f * fabsf(inv);
    if (!(m < 65536.0f)) return;  // near-horizontal (or NaN): keep the whole extent
    *lo = fmaxf(fminf(x1, x2) - m, g.mnx);
    *hi = fminf(fmaxf(x1, x2) + m, g.mxx);
}

// Tiles [*ta, *tb] of the row at y0 that can receive a command from the segment (a superset; the
// exact tests follow per tile).  False if there is none.
PM_HD bool pm_fill_candidate_span(const PmSeg &g, float y0, uint32_t t_lo, uint32_t t_hi, uint32_t *ta, uint32_t *tb) {
    if (!pm_fill_row_overlap(g, y0)) return false;
    float band_lo, band_hi;
    pm_seg_band_x(g, y0, y0 + 16.0f, &band_lo, &band_hi);
    return pm_tile_span(band_lo, band_hi, t_lo, t_hi, ta, tb);
}

// The exact tests of one candidate tile: x-overlap guard, strip vote, tile test (metal:265-353).
// Sink::fill(t, seg_index, emit, g): one Fill (+FillEdge) command for tile t.
template <class Sink>
PM_HD void pm_fill_candidate_tile(Sink &sink, const PmSeg &g, float y0, uint32_t t, uint32_t seg_index) {
    float x0 = (float)(t * PM_TILE_W);
    if (!(g.mnx < x0 + 16.0f && g.mxx > x0)) return;  // every emitting branch carries this guard
    if (!pm_fill_strip_vote(g, y0, (float)((t / PM_GROUP_TILES_X) * PM_STRIP_PX))) return;
    PmFillEmit e = pm_fill_tile_test(g, x0, y0);
    if (e.kind != PM_EMIT_NONE) sink.fill(t, seg_index, e, g);
}

// Backdrop contribution of the segment to the tiles [t_lo, t_hi] of the row at y0 (metal:331-333).
// Sink::backdrop(ta, tb, delta): backdrop += delta for tiles ta..tb (inclusive).
template <class Sink>
PM_HD void pm_fill_backdrop_row(Sink &sink, const PmSeg &g, float y0, uint32_t t_lo, uint32_t t_hi, uint32_t n_tiles_x) {
    if (!pm_fill_row_overlap(g, y0) || !(g.mny <= y0)) return;  // the segment must reach the row's top line
    float sa = pm_sign(g.a);
    if (sa == 0.0f) return;
    uint32_t t_first = pm_fill_backdrop_first_tile(g, y0, n_tiles_x);
    if (t_first < t_lo) t_first = t_lo;
    if (t_first > t_hi) return;
    const int delta = sa > 0.0f ? -1 : 1;  // backdrop -= s00, s00 == sign(a) here
    // A tile only sees the segment if its 256-px strip voted for it (metal:302):
    //   vote(S) = mnx < sx0+256  &&  (side(S) || (crosses(S) && mxx > sx0)).
    // side(S) is monotone in S like the tile test, mnx < sx0+256 is a suffix too, and
    // crosses(S) && mxx > sx0 can only hold up to the strip containing mxx: so the voting strips
    // are a suffix [s_suf, ..) plus a few strips evaluated one by one.
    const uint32_t n_strips = (n_tiles_x + PM_GROUP_TILES_X - 1) / PM_GROUP_TILES_X;
    const uint32_t s_lo = t_first / PM_GROUP_TILES_X, s_hi = t_hi / PM_GROUP_TILES_X;
    uint32_t s_side = pm_fill_strip_side_first(g, y0, n_strips);
    uint32_t s_a = g.mnx < 256.0f ? 0u : (uint32_t)fminf(floorf(g.mnx * (1.0f / 256.0f)), 4096.0f);
    uint32_t s_suf = s_side > s_a ? s_side : s_a;
    if (s_suf < s_lo) s_suf = s_lo;
    if (g.mxx > 0.0f) {
        uint32_t s_b = (uint32_t)fminf(ceilf(g.mxx * (1.0f / 256.0f)), 4096.0f) - 1u;  // last strip with sx0 < mxx
        uint32_t s_end = s_suf;  // exclusive
        if (s_end > s_hi + 1) s_end = s_hi + 1;
        if (s_end > s_b + 1) s_end = s_b + 1;
        for (uint32_t strip = s_lo > s_a ? s_lo : s_a; strip < s_end; strip++) {
            if (!pm_fill_strip_vote(g, y0, (float)(strip * PM_STRIP_PX))) continue;
            uint32_t a0 = strip * PM_GROUP_TILES_X, a1 = a0 + PM_GROUP_TILES_X - 1;
            if (a0 < t_first) a0 = t_first;
            if (a1 > t_hi) a1 = t_hi;
            sink.backdrop(a0, a1, delta);
        }
    }
    if (s_suf <= s_hi) {
        uint32_t a0 = s_suf * PM_GROUP_TILES_X;
        if (a0 < t_first) a0 = t_first;
        sink.backdrop(a0, t_hi, delta);
    }
}

// All effects of one fill segment on the tiles [t_lo, t_hi] of the tile row starting at y0.
template <class Sink>
PM_HD void pm_fill_segment_row(Sink &sink, const PmSeg &g, float y0, uint32_t t_lo, uint32_t t_hi, uint32_t n_tiles_x,
                               uint32_t seg_index) {
    uint32_t ta = 1, tb = 0;
    if (pm_fill_candidate_span(g, y0, t_lo, t_hi, &ta, &tb))
        for (uint32_t t = ta; t <= tb; t++) pm_fill_candidate_tile(sink, g, y0, t, seg_index);
    pm_fill_backdrop_row(sink, g, y0, t_lo, t_hi, n_tiles_x);
}

// ---------------------------------------------------------------------------------------------
// Stroked polylines (metal:366-445) and single stroked lines (metal:223-247)
// ---------------------------------------------------------------------------------------------

// Tile-level bbox part (metal:416-417): also implies the group-level one (metal:380-381) because
// fl(y0 - hw) >= fl(sy0 - hw) etc. by monotonicity of rounding.
PM_HD bool pm_poly_tile_overlap(const PmSeg &g, float x0, float y0, float hw) {
    return g.mxy > y0 - hw && g.mny < y0 + 16.0f + hw && g.mxx > x0 - hw && g.mnx < x0 + 16.0f + hw;
}

// Corner-sign test against the rectangle [xl - hw, xr + hw] x [yt - hw, yb + hw] (metal:382-393,
// :418-431, :231-241).
PM_HD bool pm_stroke_cross(const PmSeg &g, float xl, float xr, float yt, float yb, float hw) {
    float left = g.a * (xl - hw);
    float right = g.a * (xr + hw);
    float top = g.b * (yt - hw);
    float bot = g.b * (yb + hw);
    return pm_cross4v(top + left + g.c, top + right + g.c, bot + left + g.c, bot + right + g.c);
}

// Group-level pre-cull vote for segment seg_index (metal:375-398).  The corner test uses the y
// range of the *voting lane's* tile row, lane = seg_index & 31, unless
// fix_precull asks for the consuming tile's own row.
PM_HD bool pm_poly_group_vote(const PmSeg &g, float sx0, float sy0, float y0, float hw, uint32_t seg_index, bool fix_precull) {
    if (!(g.mxy > sy0 - hw && g.mny < sy0 + 32.0f + hw && g.mxx > sx0 - hw && g.mnx < sx0 + 256.0f + hw)) return false;
    float lane_y0 = fix_precull ? y0 : sy0 + (float)(PM_TILE_H * ((seg_index & 31u) >> 4));
    return pm_stroke_cross(g, sx0, sx0 + 256.0f, lane_y0, lane_y0 + 16.0f, hw);
}

PM_HD bool pm_poly_candidate_span(const PmSeg &g, float y0, float hw, uint32_t t_lo, uint32_t t_hi, uint32_t *ta, uint32_t *tb) {
    if (!(g.mxy > y0 - hw && g.mny < y0 + 16.0f + hw)) return false;
    float band_lo, band_hi;  // the stroke-inflated tile reaches hw beyond the tile on every side
    pm_seg_band_x(g, y0 - hw - 1.0f, y0 + 16.0f + hw + 1.0f, &band_lo, &band_hi);
    return pm_tile_span(band_lo - hw - 1.0f, band_hi + hw + 1.0f, t_lo, t_hi, ta, tb);
}

// Sink::line(t, seg_index, g): one Line command for tile t
template <class Sink>
PM_HD void pm_poly_candidate_tile(Sink &sink, const PmSeg &g, float y0, float hw, uint32_t t, uint32_t seg_index, bool fix_precull) {
    float x0 = (float)(t * PM_TILE_W);
    if (!pm_poly_tile_overlap(g, x0, y0, hw)) return;
    float sy0 = (float)(((uint32_t)y0) & ~(uint32_t)(PM_GROUP_PX_Y - 1));
    if (!pm_poly_group_vote(g, (float)((t / PM_GROUP_TILES_X) * PM_STRIP_PX), sy0, y0, hw, seg_index, fix_precull)) return;
    if (pm_stroke_cross(g, x0, x0 + 16.0f, y0, y0 + 16.0f, hw)) sink.line(t, seg_index, g);
}

template <class Sink>
PM_HD void pm_poly_segment_row(Sink &sink, const PmSeg &g, float y0, float hw, uint32_t t_lo, uint32_t t_hi,
                               uint32_t seg_index, bool fix_precull) {
    uint32_t ta = 1, tb = 0;
    if (pm_poly_candidate_span(g, y0, hw, t_lo, t_hi, &ta, &tb))
        for (uint32_t t = ta; t <= tb; t++) pm_poly_candidate_tile(sink, g, y0, hw, t, seg_index, fix_precull);
}

// ---------------------------------------------------------------------------------------------
// Record construction: what TileEncoder would have written (metal:85-143), as PmRecords
// ---------------------------------------------------------------------------------------------
PM_HD PmRecord pm_rec_fill(uint32_t item, uint32_t seg, uint32_t t, const PmFillEmit &e, const PmSeg &g) {
    PmRecord r;
    r.item = item;
    r.edge_y = 0.0f;
    r.next = 0;
    if (e.kind == PM_EMIT_EDGE) {  // metal:336-344: FillEdge(s00, yEdge) + the part of the segment right of the edge
        r.key = (seg << PM_REC_KIND_BITS) | (uint32_t)(PM_REC_FILL_EDGE_ZERO + (int)e.s00);
        r.edge_y = e.y_edge;
        float x0 = (float)(t * PM_TILE_W);
        if (g.b > 0.0f) { r.p[0] = g.sx; r.p[1] = g.sy; r.p[2] = x0; r.p[3] = e.y_edge; }
        else            { r.p[0] = x0; r.p[1] = e.y_edge; r.p[2] = g.ex; r.p[3] = g.ey; }
    } else {
        r.key = (seg << PM_REC_KIND_BITS) | PM_REC_FILL;
        r.p[0] = g.sx; r.p[1] = g.sy; r.p[2] = g.ex; r.p[3] = g.ey;
    }
    return r;
}
PM_HD PmRecord pm_rec_line(uint32_t item, uint32_t seg, const PmSeg &g) {
    PmRecord r;
    r.item = item;
    r.key = (seg << PM_REC_KIND_BITS) | PM_REC_LINE;
    r.p[0] = g.sx; r.p[1] = g.sy; r.p[2] = g.ex; r.p[3] = g.ey;
    r.edge_y = 0.0f;
    r.next = 0;
    return r;
}
// DRAWFILL (w0 = backdrop, w1 = rgba), STROKE (w0 = halfWidth bits, w1 = rgba), SOLID (w1 = rgba),
// CIRCLE (w0, w1 = bbox)
PM_HD PmRecord pm_rec_words(uint32_t item, uint32_t kind, uint32_t seg, uint32_t w0, uint32_t w1) {
    PmRecord r;
    r.item = item;
    r.key = (seg << PM_REC_KIND_BITS) | kind;
    r.p[0] = pm_u2f(w0); r.p[1] = pm_u2f(w1); r.p[2] = 0.0f; r.p[3] = 0.0f;
    r.edge_y = 0.0f;
    r.next = 0;
    return r;
}
