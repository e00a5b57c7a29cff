// TEST-ONLY host harness for the product's __host__ __device__ logic headers.
//
// It replays, sequentially on the CPU, exactly what the CUDA kernels do with
// piet-metal_b200/csrc/pm_tile_logic.h and pm_pixel_logic.h: per (item, tile row) binning with the
// exact tile predicates, backdrop prefix, opaque-cover resolution, then per-tile record sort and
// per-pixel interpretation.  tests/test_binned_logic.py compares it with the oracle so that the
// restructured (binned) algorithm is proven equal to the literal per-tile loop before any GPU
// time is spent.  It is NOT part of libpiet_metal_b200.so and is never on a product path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/piet_metal_b200.h"
#include "../../piet-metal_b200/csrc/pm_pixel_logic.h"
#include "../../piet-metal_b200/csrc/pm_scene_format.h"
#include "../../piet-metal_b200/csrc/pm_tile_logic.h"

namespace {

struct TileBin {
    uint64_t occ_color = 0;
    std::vector<PmRecord> recs;
};

struct HostSink {
    std::vector<TileBin> &row_tiles;  // tiles of the current row
    std::vector<int> &delta;          // backdrop difference array, n_tx + 1
    std::vector<uint8_t> &em;
    uint32_t item;
    void fill(uint32_t t, uint32_t seg, const PmFillEmit &e, const PmSeg &g) {
        row_tiles[t].recs.push_back(pm_rec_fill(item, seg, t, e, g));
        em[t] = 1;
    }
    void backdrop(uint32_t ta, uint32_t tb, int d) { delta[ta] += d; delta[tb + 1] -= d; }
    void line(uint32_t t, uint32_t seg, const PmSeg &g) {
        row_tiles[t].recs.push_back(pm_rec_line(item, seg, g));
        em[t] = 1;
    }
};

uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
float rd_f32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }

// Mirrors the fill kernel's per-tile state: fixed-point coverage (near terms + per-row cover
// deltas), the stroke distance field and the linear colour of the 256 pixels.
struct TileAcc {
    int acc[16][16];
    int cov[16][17];
    float dmin[16][16];
    float rgb[16][16][3];
    void near(int row, int j, int fx) { acc[row][j] += fx; }
    void cover(int row, int j, int fx) { cov[row][j] += fx; }
    void dist(int row, int j, float d) { dmin[row][j] = fminf(dmin[row][j], d); }
    void clear_fill() { memset(acc, 0, sizeof acc); memset(cov, 0, sizeof cov); }
    void clear_dist() { for (auto &r : dmin) for (float &v : r) v = 1e9f; }
};

void unpack(const float *lut, uint32_t rgba, float fg[4]) {
    fg[0] = lut[rgba & 0xff]; fg[1] = lut[(rgba >> 8) & 0xff]; fg[2] = lut[(rgba >> 16) & 0xff];
    fg[3] = (float)(rgba >> 24) / 255.0f;
}
void blend_px(float rgb[3], const float fg[4], float a) {
    for (int k = 0; k < 3; k++) rgb[k] = pm_mix(rgb[k], fg[k], a);
}

// One item's records [j0, j1) of a tile, in sorted order; the last one decides what the item is.
void apply_group(TileAcc &t, const std::vector<PmRecord> &recs, size_t j0, size_t j1, const float *lut, float tile_x0, float tile_y0) {
    const PmRecord &last = recs[j1 - 1];
    const uint32_t kind = last.key & 15u;
    float fg[4];
    if (pm_rec_is_drawfill(kind)) {
        t.clear_fill();
        for (size_t j = j0; j + 1 < j1; j++) {
            const PmRecord &q = recs[j];
            int ra, rb;
            pm_fill_rows(q.p[1], q.p[3], tile_y0, &ra, &rb);
            for (int row = ra; row <= rb; row++) pm_fill_pair(t, q.p, row, tile_x0, tile_y0);
            if ((q.key & 15u) != PM_REC_FILL)
                for (int row = 0; row < 16; row++) pm_fill_edge_row(t, q.key & 15u, q.edge_y, row, tile_y0);
        }
        unpack(lut, pm_f2u(last.p[1]), fg);
        const int backdrop = (int)pm_f2u(last.p[0]);
        for (int row = 0; row < 16; row++) {
            int run = 0;
            for (int x = 0; x < 16; x++) {
                run += t.cov[row][x];
                float alpha = pm_resolve_fill(t.acc[row][x] + run, backdrop, kind == PM_REC_DRAWFILL_EO);
                blend_px(t.rgb[row][x], fg, fg[3] * alpha);
            }
        }
    } else if (kind == PM_REC_STROKE) {
        t.clear_dist();
        const float half_width = last.p[0];
        const float reach = half_width + 0.5f;
        for (size_t j = j0; j + 1 < j1; j++) {
            const PmRecord &q = recs[j];
            int ra, rb;
            pm_line_rows(q.p[1], q.p[3], reach, tile_y0, &ra, &rb);
            for (int row = ra; row <= rb; row++) pm_line_pair(t, q.p, reach, row, tile_x0, tile_y0);
        }
        unpack(lut, pm_f2u(last.p[1]), fg);
        for (int row = 0; row < 16; row++)
            for (int x = 0; x < 16; x++) blend_px(t.rgb[row][x], fg, fg[3] * pm_saturate(half_width + 0.5f - t.dmin[row][x]));
    } else if (kind == PM_REC_CIRCLE) {
        const float black[4] = {0, 0, 0, 1};
        for (int row = 0; row < 16; row++)
            for (int x = 0; x < 16; x++)
                blend_px(t.rgb[row][x], black, pm_px_circle_alpha(pm_f2u(last.p[0]), pm_f2u(last.p[1]), tile_x0 + (float)x, tile_y0 + (float)row));
    } else if (kind == PM_REC_SOLID) {
        unpack(lut, pm_f2u(last.p[1]), fg);
        for (int row = 0; row < 16; row++)
            for (int x = 0; x < 16; x++) blend_px(t.rgb[row][x], fg, fg[3]);
    }
}

float encode(float v) { return v < 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f; }

}  // namespace

struct pmh_tile_item { uint32_t item; int32_t backdrop; uint32_t effect; };

namespace {
// Binning of the strip [tile_y0, tile_y1): per-tile record lists and opaque covers.
std::vector<std::vector<TileBin>> bin_scene(const uint8_t *scene, uint32_t n_tx, uint32_t tile_y0, uint32_t tile_y1, bool fix, uint32_t flags = 0) {
    const uint32_t n_rows = tile_y1 - tile_y0;
    const uint32_t n_items = rd_u32(scene), items_ix = rd_u32(scene + 4);
    std::vector<std::vector<TileBin>> tiles(n_rows, std::vector<TileBin>(n_tx));
    std::vector<int> delta(n_tx + 2);
    std::vector<uint8_t> em(n_tx + 1);

    // ---- binning: one (item, row) unit at a time, as k_bin's warps do ----
    for (uint32_t item = 0; item < n_items; item++) {
        pm_bbox bb; memcpy(&bb, scene + 8 + 8 * (size_t)item, 8);
        const uint8_t *it = scene + items_ix + 32 * (size_t)item;
        uint32_t tag = rd_u32(it);
        if (tag < 1 || tag > 4) continue;
        uint32_t t_lo = bb.x0 >> 4, t_hi = bb.x1 >> 4;
        if (t_lo >= n_tx || t_hi < t_lo) continue;
        if (t_hi > n_tx - 1) t_hi = n_tx - 1;
        uint32_t r_lo = std::max<uint32_t>(bb.y0 >> 4, tile_y0), r_hi = std::min<uint32_t>(bb.y1 >> 4, tile_y1 - 1);
        for (uint32_t row = r_lo; row <= r_hi && r_hi != 0xffffffffu && r_lo <= r_hi; row++) {
            std::vector<TileBin> &rt = tiles[row - tile_y0];
            std::fill(delta.begin(), delta.end(), 0);
            std::fill(em.begin(), em.end(), 0);
            HostSink sink{rt, delta, em, item};
            const float y0 = (float)(row * 16);
            if (tag == PM_ITEM_FILL) {
                uint32_t rgba = rd_u32(it + 8), n = rd_u32(it + 12);
                const bool even_odd = (flags & PM_FLAG_FILL_RULES) != 0 && (rd_u32(it + 4) & PM_FILL_EVEN_ODD) != 0;
                const uint8_t *pts = scene + rd_u32(it + 16);
                for (uint32_t k = 0; k < n; k++) {
                    uint32_t k1 = k + 1 == n ? 0 : k + 1;
                    PmSeg g = pm_seg(rd_f32(pts + 8 * k), rd_f32(pts + 8 * k + 4), rd_f32(pts + 8 * k1), rd_f32(pts + 8 * k1 + 4));
                    pm_fill_segment_row(sink, g, y0, t_lo, t_hi, n_tx, k);
                }
                int backdrop = 0;
                for (uint32_t t = t_lo; t <= t_hi; t++) {
                    backdrop += delta[t];
                    if (em[t]) {
                        rt[t].recs.push_back(pm_rec_words(item, even_odd ? PM_REC_DRAWFILL_EO : PM_REC_DRAWFILL, PM_REC_SEG_MAX, (uint32_t)backdrop, rgba));
                    } else if (even_odd ? (backdrop & 1) != 0 : backdrop != 0) {
                        if ((rgba & 0xff000000u) == 0xff000000u)
                            rt[t].occ_color = std::max(rt[t].occ_color, ((uint64_t)(item + 1) << 32) | rgba);
                        else
                            rt[t].recs.push_back(pm_rec_words(item, PM_REC_SOLID, 0, 0, rgba));
                    }
                }
            } else if (tag == PM_ITEM_POLY) {
                uint32_t rgba = rd_u32(it + 4);
                float w = rd_f32(it + 8);
                uint32_t n_seg = rd_u32(it + 12) - 1;
                const uint8_t *pts = scene + rd_u32(it + 16);
                float hw = 0.5f * w + 0.5f;
                for (uint32_t k = 0; k < n_seg; k++) {
                    PmSeg g = pm_seg(rd_f32(pts + 8 * k), rd_f32(pts + 8 * k + 4), rd_f32(pts + 8 * (k + 1)), rd_f32(pts + 8 * (k + 1) + 4));
                    pm_poly_segment_row(sink, g, y0, hw, t_lo, t_hi, k, fix);
                }
                for (uint32_t t = t_lo; t <= t_hi; t++)
                    if (em[t]) rt[t].recs.push_back(pm_rec_words(item, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * w), rgba));
            } else if (tag == PM_ITEM_LINE) {
                uint32_t rgba = rd_u32(it + 8);
                float w = rd_f32(it + 12);
                PmSeg g = pm_seg(rd_f32(it + 16), rd_f32(it + 20), rd_f32(it + 24), rd_f32(it + 28));
                float hw = 0.5f * w + 0.5f;
                for (uint32_t t = t_lo; t <= t_hi; t++) {
                    float x0 = (float)(t * 16);
                    if (pm_stroke_cross(g, x0, x0 + 16.0f, y0, y0 + 16.0f, hw)) {
                        sink.line(t, 0, g);
                        rt[t].recs.push_back(pm_rec_words(item, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * w), rgba));
                    }
                }
            } else {
                uint32_t b_lo = (uint32_t)bb.x0 | ((uint32_t)bb.y0 << 16), b_hi = (uint32_t)bb.x1 | ((uint32_t)bb.y1 << 16);
                for (uint32_t t = t_lo; t <= t_hi; t++) rt[t].recs.push_back(pm_rec_words(item, PM_REC_CIRCLE, 0, b_lo, b_hi));
            }
        }
    }

    return tiles;
}
}  // namespace

extern "C" int pmh_render(const uint8_t *scene, size_t len, uint32_t width, uint32_t height, uint32_t tile_y0, uint32_t tile_y1,
                          uint32_t flags, uint8_t *rgba8, size_t stride8, float *rgba32f, size_t stride32f_bytes,
                          uint32_t *offsets, pmh_tile_item *items, size_t cap_items, size_t *n_items_out, uint32_t *solid) {
    (void)len;
    const uint32_t tile_y0_strip = tile_y0;
    const uint32_t n_tx = (width + 15) / 16, n_ty = (height + 15) / 16;
    if (tile_y1 > n_ty) tile_y1 = n_ty;
    const uint32_t n_rows = tile_y1 - tile_y0;
    const uint32_t n_items = rd_u32(scene), items_ix = rd_u32(scene + 4);
    const bool fix = (flags & 1u) != 0;
    float lut[256];
    for (int i = 0; i < 256; i++) lut[i] = pm_srgb_byte_to_linear((uint32_t)i);

    std::vector<std::vector<TileBin>> tiles = bin_scene(scene, n_tx, tile_y0, tile_y1, fix, flags);

    // ---- fill/blend: per tile, as k_fine does ----
    size_t total = 0;
    for (uint32_t r = 0; r < n_rows; r++) {
        for (uint32_t tx = 0; tx < n_tx; tx++) {
            TileBin &tb = tiles[r][tx];
            const uint32_t occ_item1 = (uint32_t)(tb.occ_color >> 32), occ_rgba = (uint32_t)tb.occ_color;
            std::vector<PmRecord> recs;
            bool has_draw = false;
            for (const PmRecord &q : tb.recs)
                if (q.item >= occ_item1) { recs.push_back(q); if ((q.key & 15u) != PM_REC_SOLID) has_draw = true; }
            std::sort(recs.begin(), recs.end(), [](const PmRecord &a, const PmRecord &b) {
                return (((uint64_t)a.item << 32) | a.key) < (((uint64_t)b.item << 32) | b.key);
            });
            const size_t tile_ix = (size_t)r * n_tx + tx;
            if (offsets) {
                offsets[tile_ix] = (uint32_t)total;
                auto push = [&](uint32_t item, int32_t backdrop, uint32_t effect) {
                    if (items && total < cap_items) { items[total].item = item; items[total].backdrop = backdrop; items[total].effect = effect; }
                    total++;
                };
                if (occ_item1) push(occ_item1 - 1, 0, 1);
                for (size_t k = 0; k < recs.size(); k++) {
                    if (k + 1 != recs.size() && recs[k + 1].item == recs[k].item) continue;
                    uint32_t kind = recs[k].key & 15u;
                    if (pm_rec_is_drawfill(kind)) push(recs[k].item, (int32_t)pm_f2u(recs[k].p[0]), 0);
                    else if (kind == PM_REC_SOLID) push(recs[k].item, 0, 1);
                    else push(recs[k].item, 0, 0);
                }
            }
            const uint32_t solid_color = has_draw ? 0u : (occ_item1 ? occ_rgba : 0xffffffffu);
            if (solid) solid[tile_ix] = solid_color;
            if (!rgba8 && !rgba32f) continue;
            TileAcc tacc;
            if (has_draw) {
                for (int row = 0; row < 16; row++)
                    for (int x = 0; x < 16; x++) tacc.rgb[row][x][0] = tacc.rgb[row][x][1] = tacc.rgb[row][x][2] = 1.0f;
                if (occ_item1) {
                    float fg[4]; unpack(lut, occ_rgba, fg);
                    for (int row = 0; row < 16; row++)
                        for (int x = 0; x < 16; x++) blend_px(tacc.rgb[row][x], fg, fg[3]);
                }
                const float tile_x0 = (float)(tx * 16), tile_y0 = (float)((tile_y0_strip + r) * 16);
                for (size_t j0 = 0; j0 < recs.size();) {
                    size_t j1 = j0 + 1;
                    while (j1 < recs.size() && recs[j1].item == recs[j0].item) j1++;
                    apply_group(tacc, recs, j0, j1, lut, tile_x0, tile_y0);
                    j0 = j1;
                }
            }
            for (uint32_t py = 0; py < 16; py++) {
                uint32_t y = (tile_y0_strip + r) * 16 + py;
                if (y >= height) break;
                for (uint32_t pxi = 0; pxi < 16; pxi++) {
                    uint32_t x = tx * 16 + pxi;
                    if (x >= width) break;
                    float out[4];
                    uint8_t b[4];
                    if (!has_draw) {
                        for (int k = 0; k < 4; k++) { b[k] = (uint8_t)(solid_color >> (8 * k)); out[k] = (float)b[k] / 255.0f; }
                    } else {
                        for (int k = 0; k < 3; k++) { out[k] = encode(tacc.rgb[py][pxi][k]); b[k] = (uint8_t)pm_unorm8(out[k]); }
                        out[3] = 1.0f; b[3] = 255;
                    }
                    size_t ry = (size_t)(y - tile_y0_strip * 16);
                    if (rgba8) memcpy(rgba8 + ry * stride8 + 4 * (size_t)x, b, 4);
                    if (rgba32f) memcpy((uint8_t *)rgba32f + ry * stride32f_bytes + 16 * (size_t)x, out, 16);
                }
            }
        }
    }
    if (offsets) offsets[(size_t)n_rows * n_tx] = (uint32_t)total;
    if (n_items_out) *n_items_out = total;
    return 0;
}

// Workload statistics of the fill kernel for one scene (tools/tile_stats.py): how many records,
// items, (record, pixel row) pairs and near pixels the tiles with records carry.
//   out[0] tiles with records   out[1] records       out[2] items (groups)   out[3] fill pairs
//   out[4] fill near pixels     out[5] line pairs    out[6] line pixels      out[7] fill-edge records
//   out[8] has_draw tiles       out[9] max near px in one pair (sum over pairs of that max is out[10])
//   hist_items[k]: tiles with k items (k capped at 31); hist_recs[k]: tiles with k records (capped 63)
extern "C" int pmh_stats(const uint8_t *scene, uint32_t width, uint32_t height, uint32_t tile_y0, uint32_t tile_y1, uint64_t *out,
                         uint64_t *hist_items, uint64_t *hist_recs) {
    const uint32_t n_tx = (width + 15) / 16, n_ty = (height + 15) / 16;
    if (tile_y1 > n_ty) tile_y1 = n_ty;
    auto tiles = bin_scene(scene, n_tx, tile_y0, tile_y1, false);
    struct Count {
        uint64_t near = 0, pairs = 0, maxnear = 0, cur = 0;
        void near_(int) { near++; cur++; }
    };
    struct FillCounter { uint64_t near = 0, cur = 0; void near_px(int, int, int) {} };
    struct AccCount {
        uint64_t n_near = 0, n_cover = 0, n_dist = 0;
        void near(int, int, int) { n_near++; }
        void cover(int, int, int) { n_cover++; }
        void dist(int, int, float) { n_dist++; }
    };
    for (uint32_t r = 0; r < tile_y1 - tile_y0; r++)
        for (uint32_t tx = 0; tx < n_tx; tx++) {
            TileBin &tb = tiles[r][tx];
            const uint32_t occ_item1 = (uint32_t)(tb.occ_color >> 32);
            if (tb.recs.empty()) continue;
            out[0]++;
            std::vector<PmRecord> recs;
            bool has_draw = false;
            for (const PmRecord &q : tb.recs)
                if (q.item >= occ_item1) { recs.push_back(q); if ((q.key & 15u) != PM_REC_SOLID) has_draw = true; }
            hist_recs[std::min<size_t>(tb.recs.size(), 63)]++;
            if (!has_draw) { hist_items[0]++; continue; }
            out[8]++;
            std::sort(recs.begin(), recs.end(), [](const PmRecord &a, const PmRecord &b) {
                return (((uint64_t)a.item << 32) | a.key) < (((uint64_t)b.item << 32) | b.key);
            });
            out[1] += recs.size();
            const float tile_x0 = (float)(tx * 16), ty0 = (float)((tile_y0 + r) * 16);
            size_t n_groups = 0;
            for (size_t j0 = 0; j0 < recs.size();) {
                size_t j1 = j0 + 1;
                while (j1 < recs.size() && recs[j1].item == recs[j0].item) j1++;
                n_groups++;
                const PmRecord &last = recs[j1 - 1];
                const uint32_t kind = last.key & 15u;
                for (size_t j = j0; j + 1 < j1; j++) {
                    const PmRecord &q = recs[j];
                    int ra, rb;
                    if (pm_rec_is_drawfill(kind)) {
                        pm_fill_rows(q.p[1], q.p[3], ty0, &ra, &rb);
                        for (int row = ra; row <= rb; row++) {
                            AccCount c;
                            pm_fill_pair(c, q.p, row, tile_x0, ty0);
                            out[3]++; out[4] += c.n_near; out[10] += c.n_near; if (c.n_near > out[9]) out[9] = c.n_near;
                        }
                        if ((q.key & 15u) != PM_REC_FILL) out[7]++;
                    } else if (kind == PM_REC_STROKE) {
                        pm_line_rows(q.p[1], q.p[3], last.p[0] + 0.5f, ty0, &ra, &rb);
                        for (int row = ra; row <= rb; row++) {
                            AccCount c;
                            pm_line_pair(c, q.p, last.p[0] + 0.5f, row, tile_x0, ty0);
                            out[5]++; out[6] += c.n_dist;
                        }
                    }
                }
                j0 = j1;
            }
            out[2] += n_groups;
            hist_items[std::min<size_t>(n_groups, 31)]++;
        }
    return 0;
}

// Per-tile workload of one tile row (tools): for each tile of row `ty`: records, items, fill pairs, fill near
// pixels, line pairs, line pixels, LINE records, FILL records.  out has 8 * n_tx entries.
extern "C" int pmh_row_stats(const uint8_t *scene, uint32_t width, uint32_t height, uint32_t ty, uint64_t *out) {
    const uint32_t n_tx = (width + 15) / 16;
    (void)height;
    auto tiles = bin_scene(scene, n_tx, ty, ty + 1, false);
    struct AccCount {
        uint64_t n_near = 0, n_cover = 0, n_dist = 0;
        void near(int, int, int) { n_near++; }
        void cover(int, int, int) { n_cover++; }
        void dist(int, int, float) { n_dist++; }
    };
    for (uint32_t tx = 0; tx < n_tx; tx++) {
        TileBin &tb = tiles[0][tx];
        uint64_t *o = out + 8 * (size_t)tx;
        const uint32_t occ_item1 = (uint32_t)(tb.occ_color >> 32);
        std::vector<PmRecord> recs;
        for (const PmRecord &q : tb.recs) if (q.item >= occ_item1) recs.push_back(q);
        std::sort(recs.begin(), recs.end(), [](const PmRecord &a, const PmRecord &b) {
            return (((uint64_t)a.item << 32) | a.key) < (((uint64_t)b.item << 32) | b.key);
        });
        o[0] = tb.recs.size();
        const float tile_x0 = (float)(tx * 16), ty0 = (float)(ty * 16);
        for (size_t j0 = 0; j0 < recs.size();) {
            size_t j1 = j0 + 1;
            while (j1 < recs.size() && recs[j1].item == recs[j0].item) j1++;
            o[1]++;
            const PmRecord &last = recs[j1 - 1];
            const uint32_t kind = last.key & 15u;
            for (size_t j = j0; j + 1 < j1; j++) {
                const PmRecord &q = recs[j];
                int ra, rb;
                if (pm_rec_is_drawfill(kind)) {
                    o[7]++;
                    pm_fill_rows(q.p[1], q.p[3], ty0, &ra, &rb);
                    for (int row = ra; row <= rb; row++) { AccCount c; pm_fill_pair(c, q.p, row, tile_x0, ty0); o[2]++; o[3] += c.n_near; }
                } else if (kind == PM_REC_STROKE) {
                    o[6]++;
                    pm_line_rows(q.p[1], q.p[3], last.p[0] + 0.5f, ty0, &ra, &rb);
                    for (int row = ra; row <= rb; row++) { AccCount c; pm_line_pair(c, q.p, last.p[0] + 0.5f, row, tile_x0, ty0); o[4]++; o[5] += c.n_dist; }
                }
            }
            j0 = j1;
        }
    }
    return 0;
}
