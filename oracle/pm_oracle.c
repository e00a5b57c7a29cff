/*
 * pm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of the reference's compute render path, used as the parity oracle by
 * tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py.
 * Nothing under piet-metal_b200/ may include, link or call it.
 *
 * PARITY UNPINNED: the reference (linebender/piet-metal @ 71afb99) ships no tests, golden images
 * or known-answer vectors for this path and neither its Metal kernels nor
 * its Rust feed can be built in this image, so this oracle is pinned only by its own
 * hand-derived known-answer tests (tests/test_oracle_kat.py), not by reference output.
 *
 * What it follows, line by line (paths relative to the reference tree):
 *   tile pass     TestApp/PietRender.metal:160-454  tileKernel, one call of pmo_tile() per tile,
 *                 with the threadgroup ballots (:191-208, :257-301, :375-405) evaluated for the
 *                 lanes whose votes this tile consumes
 *   TileEncoder   TestApp/PietRender.metal:69-157   (solidColor tracking, opaque-Solid rewind, Bail)
 *   pixel pass    TestApp/PietRender.metal:457-566  renderKernel, pmo_pixel()
 *   composite     TestApp/PietRender.metal:16-44, :453  solid tile => the solid colour's bytes
 *   formats       TestApp/GenTypes.h (scene readers :49-57,:119-138,:193-209,:257-273; Cmd layouts
 *                 :340-495), TestApp/PietShaderTypes.h:17-22
 * Deliberate differences, all stated in SURVEY.md section 8: fp32 everywhere the reference uses
 * `half` (metal:470-472,502,526,537); no 4096-px / 4096-byte-per-tile caps (command lists grow);
 * output is RGBA8 in memory order R,G,B,A instead of a BGRA8 texture.
 * Metal built-ins are restated per the Metal Shading Language specification: sign(0) = 0,
 * saturate = clamp to [0,1], mix(x,y,a) = x + (y-x)*a, min/max = fmin/fmax, unorm8 write =
 * round-to-nearest-even of clamp(v,0,1)*255.  Build with -ffp-contract=off: no FMA contraction.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE_W 16 /* PietShaderTypes.h:17-18 */
#define TILE_H 16
#define GROUP_W 16 /* tilerGroupWidth  :21 */
#define GROUP_H 2  /* tilerGroupHeight :22 */

enum { ITEM_CIRCLE = 1, ITEM_LINE = 2, ITEM_FILL = 3, ITEM_POLY = 4 };            /* GenTypes.h:325-328 */
enum { CMD_END = 1, CMD_CIRCLE = 2, CMD_LINE = 3, CMD_FILL = 4, CMD_STROKE = 5,   /* GenTypes.h:440-495 */
       CMD_FILLEDGE = 6, CMD_DRAWFILL = 7, CMD_SOLID = 8, CMD_BAIL = 9 };

#define PMO_FLAG_FIX_POLY_PRECULL 1u
/* Extension (off by default: the reference ignores the word): honour bit 0 of PietFill.flags as "even-odd fill
 * rule" -- the rule the reference names but leaves unimplemented (PietRender.metal:538-540 gives the formula;
 * SceneEncoder.h:44 reserves the word "for winding rule").  A DrawFill command of such an item carries 1 in the
 * otherwise unused body word at byte 12. */
#define PMO_FLAG_FILL_RULES 4u
#define FILL_EVEN_ODD 1u

typedef struct { uint32_t tag; uint32_t body[5]; } pmo_cmd; /* GenTypes.h:430-433, 24 bytes */
typedef struct { uint32_t item; int32_t backdrop; uint32_t effect; } pmo_tile_item; /* effect 0 draw, 1 solid */

typedef struct { float x, y; } f2;

static uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static float rd_f32(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }
static f2 rd_f2(const uint8_t *p) { f2 v; memcpy(&v, p, 8); return v; }
static uint32_t f_bits(float f) { uint32_t v; memcpy(&v, &f, 4); return v; }

static float signf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
static float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static float mixf(float x, float y, float a) { return x + (y - x) * a; }

/* ------------------------------------------------------------------------------------------ */
/* TileEncoder (PietRender.metal:69-157) with a growable list and the per-tile item log        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    pmo_cmd *cmds; size_t n, cap;          /* dst - tileBegin, in commands */
    pmo_tile_item *items; size_t ni, icap; /* items that emitted >= 1 command since the last rewind */
    uint32_t solid_color;                  /* :74 */
} encoder;

static void enc_reset(encoder *e) { e->n = 0; e->ni = 0; e->solid_color = 0xffffffffu; }

static pmo_cmd *enc_push(encoder *e) {
    if (e->n == e->cap) {
        e->cap = e->cap ? e->cap * 2 : 64;
        e->cmds = (pmo_cmd *)realloc(e->cmds, e->cap * sizeof(pmo_cmd));
    }
    pmo_cmd *c = &e->cmds[e->n++];
    memset(c, 0, sizeof *c);
    return c;
}
static void enc_log_item(encoder *e, uint32_t item, int32_t backdrop, uint32_t effect) {
    if (e->ni == e->icap) {
        e->icap = e->icap ? e->icap * 2 : 16;
        e->items = (pmo_tile_item *)realloc(e->items, e->icap * sizeof(pmo_tile_item));
    }
    e->items[e->ni].item = item; e->items[e->ni].backdrop = backdrop; e->items[e->ni].effect = effect;
    e->ni++;
}
static void put_f2(pmo_cmd *c, int byte_off, f2 v) { memcpy((uint8_t *)c + byte_off, &v, 8); }

static void encode_circle(encoder *e, const uint16_t bbox[4]) { /* :76-84 */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_CIRCLE;
    memcpy((uint8_t *)c + 8, bbox, 8);
    e->solid_color = 0;
}
static void encode_line(encoder *e, f2 start, f2 end) { /* :85-93 */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_LINE; put_f2(c, 8, start); put_f2(c, 16, end);
    e->solid_color = 0;
}
static void encode_stroke(encoder *e, uint32_t rgba, float width) { /* :94-102 */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_STROKE;
    c->body[0] = f_bits(0.5f * width);
    c->body[1] = rgba;
    e->solid_color = 0;
}
static void encode_fill(encoder *e, f2 start, f2 end) { /* :103-110, does not touch solidColor */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_FILL; put_f2(c, 8, start); put_f2(c, 16, end);
}
static void encode_fill_edge(encoder *e, float sign, float y) { /* :111-118 */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_FILLEDGE;
    c->body[0] = (uint32_t)(int32_t)sign; /* cmd.sign is an int (GenTypes.h:392-396) */
    c->body[1] = f_bits(y);
}
static void encode_draw_fill(encoder *e, uint32_t rgba, int backdrop, uint32_t rule) { /* :119-127; rule: extension, 0 in the reference */
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_DRAWFILL;
    c->body[0] = (uint32_t)backdrop;
    c->body[1] = rgba;
    c->body[2] = rule;
    e->solid_color = 0;
}
static void encode_solid(encoder *e, uint32_t rgba) { /* :128-143 */
    if ((rgba & 0xff000000u) == 0xff000000u) {
        e->solid_color = rgba;
        e->n = 0;  /* dst = tileBegin */
        e->ni = 0; /* the occluded items no longer have commands in the list */
    }
    pmo_cmd *c = enc_push(e);
    c->tag = CMD_SOLID;
    c->body[0] = rgba;
}
static uint32_t enc_end(encoder *e) { /* :145-152 */
    if (e->solid_color) {
        /* Cmd_write_tag(tileBegin, 0, Cmd_Bail) */
        if (e->cap == 0) { enc_push(e); e->n = 0; }
        e->cmds[0].tag = CMD_BAIL;
        if (e->n == 0) e->n = 1;
    } else {
        pmo_cmd *c = enc_push(e);
        c->tag = CMD_END;
    }
    return e->solid_color;
}

/* ------------------------------------------------------------------------------------------ */
/* tileKernel for one tile (PietRender.metal:160-454)                                          */
/* ------------------------------------------------------------------------------------------ */
static uint32_t pmo_tile(const uint8_t *scene, uint32_t gx, uint32_t gy, uint32_t flags, encoder *enc) {
    const uint32_t x0 = gx * TILE_W;  /* ushort in the reference (:168); widths stay below 65536 */
    const uint32_t y0 = gy * TILE_H;
    enc_reset(enc);
    const uint32_t tgs = GROUP_W * GROUP_H; /* 32 */
    const uint32_t stw = GROUP_W * TILE_W;  /* 256, :180 */
    const uint32_t sth = GROUP_H * TILE_H;  /* 32 */
    const uint32_t sx0 = x0 & ~(stw - 1);
    const uint32_t sy0 = y0 & ~(sth - 1);
    const float fx0 = (float)x0, fy0 = (float)y0;

    const uint8_t *bboxes = scene + 8;          /* &group->bbox, :188 */
    const uint32_t n = rd_u32(scene);           /* SimpleGroup_n_items, :189 */
    const uint32_t items_ref = rd_u32(scene + 4);

    for (uint32_t i = 0; i < n; i += tgs) {
        /* first-level ballot over the 32 lanes of the threadgroup (:192-208) */
        uint32_t v = 0;
        for (uint32_t tix = 0; tix < tgs; tix++) {
            if (i + tix < n) {
                uint16_t bb[4]; memcpy(bb, bboxes + 8 * (size_t)(i + tix), 8);
                if (bb[2] >= sx0 && bb[0] < sx0 + stw && bb[3] >= sy0 && bb[1] < sy0 + sth) v |= 1u << (tix & 31);
            }
        }
        while (v) {
            uint32_t ix = i + (uint32_t)__builtin_ctz(v);
            uint16_t bbox[4]; memcpy(bbox, bboxes + 8 * (size_t)ix, 8);
            int hit = bbox[2] >= x0 && bbox[0] < x0 + TILE_W && bbox[3] >= y0 && bbox[1] < y0 + TILE_H; /* :214 */
            const uint8_t *item = scene + items_ref + 32 * (size_t)ix;
            uint32_t item_type = rd_u32(item);
            switch (item_type) {
            case ITEM_CIRCLE: /* :218-222 */
                if (hit) { encode_circle(enc, bbox); enc_log_item(enc, ix, 0, 0); }
                break;
            case ITEM_LINE: { /* :223-247 */
                if (hit) {
                    uint32_t rgba = rd_u32(item + 8);
                    float width = rd_f32(item + 12);
                    f2 start = rd_f2(item + 16), end = rd_f2(item + 24);
                    float a = end.y - start.y;
                    float b = start.x - end.x;
                    float c = -(a * start.x + b * start.y);
                    float hw = 0.5f * width + 0.5f;
                    float left = a * (fx0 - hw);
                    float right = a * ((float)(x0 + TILE_W) + hw);
                    float top = b * (fy0 - hw);
                    float bot = b * ((float)(y0 + TILE_H) + hw);
                    float s00 = signf(top + left + c);
                    float s01 = signf(top + right + c);
                    float s10 = signf(bot + left + c);
                    float s11 = signf(bot + right + c);
                    if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f) {
                        encode_line(enc, start, end);
                        encode_stroke(enc, rgba, width);
                        enc_log_item(enc, ix, 0, 0);
                    }
                }
                break;
            }
            case ITEM_FILL: { /* :248-365 */
                uint32_t rgba = rd_u32(item + 8);
                uint32_t n_points = rd_u32(item + 12);
                const uint8_t *pts = scene + rd_u32(item + 16);
                const uint32_t rule = (flags & PMO_FLAG_FILL_RULES) ? (rd_u32(item + 4) & FILL_EVEN_ODD) : 0u;
                float backdrop = 0;
                int any_fill = 0;
                for (uint32_t j = 0; j < n_points; j += 16) {
                    /* The 16 lanes of this tile's row each pre-test one segment against the
                     * 256-px strip of this row (:258-295); this tile consumes exactly those 16
                     * votes: fillVote = (rd >> (tix & 16)) & 0xffff (:302). */
                    uint32_t fill_vote = 0;
                    for (uint32_t lane = 0; lane < 16; lane++) {
                        uint32_t fill_ix = j + lane;
                        if (fill_ix >= n_points) continue;
                        int fill_hit = 0;
                        f2 start = rd_f2(pts + 8 * (size_t)fill_ix);
                        f2 end = rd_f2(pts + 8 * (size_t)(fill_ix + 1 == n_points ? 0 : fill_ix + 1));
                        f2 xymin = { fminf(start.x, end.x), fminf(start.y, end.y) };
                        f2 xymax = { fmaxf(start.x, end.x), fmaxf(start.y, end.y) };
                        if (xymax.y >= fy0 && xymin.y < (float)(y0 + TILE_H) && xymin.x < (float)(sx0 + stw)) {
                            float a = end.y - start.y;
                            float b = start.x - end.x;
                            float c = -(a * start.x + b * start.y);
                            float left = a * (float)sx0;
                            float right = a * (float)(sx0 + stw);
                            float ytop = fmaxf(fy0, xymin.y);
                            float ybot = fminf((float)(y0 + TILE_H), xymax.y);
                            float top = b * ytop;
                            float bot = b * ybot;
                            float s_top_left = signf(right - a * (float)TILE_W + fy0 * b + c);
                            float s00 = signf(top + left + c);
                            float s01 = signf(top + right + c);
                            float s10 = signf(bot + left + c);
                            float s11 = signf(bot + right + c);
                            if (s_top_left == signf(a) && xymin.y <= fy0) fill_hit = 1;
                            if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f && xymax.x > (float)sx0) fill_hit = 1;
                        }
                        if (fill_hit) fill_vote |= 1u << lane;
                    }
                    while (fill_vote) {
                        uint32_t fill_sub_ix = (uint32_t)__builtin_ctz(fill_vote);
                        uint32_t fill_ix = j + fill_sub_ix;
                        if (hit) { /* :307-355 */
                            f2 start = rd_f2(pts + 8 * (size_t)fill_ix);
                            f2 end = rd_f2(pts + 8 * (size_t)(fill_ix + 1 == n_points ? 0 : fill_ix + 1));
                            f2 xymin = { fminf(start.x, end.x), fminf(start.y, end.y) };
                            f2 xymax = { fmaxf(start.x, end.x), fmaxf(start.y, end.y) };
                            float a = end.y - start.y;
                            float b = start.x - end.x;
                            float c = -(a * start.x + b * start.y);
                            float left = a * fx0;
                            float right = a * (float)(x0 + TILE_W);
                            float ytop = fmaxf(fy0, xymin.y);
                            float ybot = fminf((float)(y0 + TILE_H), xymax.y);
                            float top = b * ytop;
                            float bot = b * ybot;
                            float s_top_left = signf(left + fy0 * b + c);
                            float s00 = signf(top + left + c);
                            float s01 = signf(top + right + c);
                            float s10 = signf(bot + left + c);
                            float s11 = signf(bot + right + c);
                            if (s_top_left == signf(a) && xymin.y <= fy0) backdrop -= s00;
                            if (xymin.x < fx0 && xymax.x > fx0) {
                                float y_edge = mixf(start.y, end.y, (start.x - fx0) / b);
                                if (y_edge >= fy0 && y_edge < (float)(y0 + TILE_H)) {
                                    encode_fill_edge(enc, s00, y_edge);
                                    f2 on_edge = { fx0, y_edge };
                                    if (b > 0.0f) encode_fill(enc, start, on_edge);
                                    else encode_fill(enc, on_edge, end);
                                    any_fill = 1;
                                } else if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f) {
                                    encode_fill(enc, start, end);
                                    any_fill = 1;
                                }
                            } else if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f
                                       && xymin.x < (float)(x0 + TILE_W) && xymax.x > fx0) {
                                encode_fill(enc, start, end);
                                any_fill = 1;
                            }
                        }
                        fill_vote &= ~(1u << fill_sub_ix);
                    }
                }
                if (any_fill) { /* :359-363; float -> int conversion at the call (:119) */
                    encode_draw_fill(enc, rgba, (int)backdrop, rule);
                    enc_log_item(enc, ix, (int)backdrop, 0);
                } else if (rule == FILL_EVEN_ODD ? ((int)backdrop & 1) != 0 : backdrop != 0.0f) { /* even-odd: covered iff the winding number is odd */
                    encode_solid(enc, rgba);
                    enc_log_item(enc, ix, 0, 1);
                }
                break;
            }
            case ITEM_POLY: { /* :366-445 */
                uint32_t rgba = rd_u32(item + 4);
                float width = rd_f32(item + 8);
                const uint8_t *pts = scene + rd_u32(item + 16);
                uint32_t n_seg = rd_u32(item + 12) - 1;
                int any_stroke = 0;
                float hw = 0.5f * width + 0.5f;
                for (uint32_t j = 0; j < n_seg; j += 32) {
                    /* all 32 lanes vote, each with the y-range of ITS OWN tile row (:386-389),
                     * and every tile of the group consumes the whole mask (:406): quirk 10 */
                    uint32_t poly_vote = 0;
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        uint32_t poly_ix = j + lane;
                        if (poly_ix >= n_seg) continue;
                        uint32_t lane_y0 = (flags & PMO_FLAG_FIX_POLY_PRECULL) ? y0 : sy0 + TILE_H * (lane >> 4);
                        f2 start = rd_f2(pts + 8 * (size_t)poly_ix);
                        f2 end = rd_f2(pts + 8 * (size_t)(poly_ix + 1));
                        f2 xymin = { fminf(start.x, end.x), fminf(start.y, end.y) };
                        f2 xymax = { fmaxf(start.x, end.x), fmaxf(start.y, end.y) };
                        if (xymax.y > (float)sy0 - hw && xymin.y < (float)(sy0 + sth) + hw &&
                            xymax.x > (float)sx0 - hw && xymin.x < (float)(sx0 + stw) + hw) {
                            float a = end.y - start.y;
                            float b = start.x - end.x;
                            float c = -(a * start.x + b * start.y);
                            float left = a * ((float)sx0 - hw);
                            float right = a * ((float)(sx0 + stw) + hw);
                            float top = b * ((float)lane_y0 - hw);
                            float bot = b * ((float)(lane_y0 + TILE_H) + hw);
                            float s00 = signf(top + left + c);
                            float s01 = signf(top + right + c);
                            float s10 = signf(bot + left + c);
                            float s11 = signf(bot + right + c);
                            if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f) poly_vote |= 1u << lane;
                        }
                    }
                    while (poly_vote) {
                        uint32_t poly_sub_ix = (uint32_t)__builtin_ctz(poly_vote);
                        uint32_t poly_ix = j + poly_sub_ix;
                        if (hit) { /* :411-436 */
                            f2 start = rd_f2(pts + 8 * (size_t)poly_ix);
                            f2 end = rd_f2(pts + 8 * (size_t)(poly_ix + 1));
                            f2 xymin = { fminf(start.x, end.x), fminf(start.y, end.y) };
                            f2 xymax = { fmaxf(start.x, end.x), fmaxf(start.y, end.y) };
                            if (xymax.y > fy0 - hw && xymin.y < (float)(y0 + TILE_H) + hw &&
                                xymax.x > fx0 - hw && xymin.x < (float)(x0 + TILE_W) + hw) {
                                float a = end.y - start.y;
                                float b = start.x - end.x;
                                float c = -(a * start.x + b * start.y);
                                float left = a * (fx0 - hw);
                                float right = a * ((float)(x0 + TILE_W) + hw);
                                float top = b * (fy0 - hw);
                                float bot = b * ((float)(y0 + TILE_H) + hw);
                                float s00 = signf(top + left + c);
                                float s01 = signf(top + right + c);
                                float s10 = signf(bot + left + c);
                                float s11 = signf(bot + right + c);
                                if (s00 * s01 + s00 * s10 + s00 * s11 < 3.0f) {
                                    encode_line(enc, start, end);
                                    any_stroke = 1;
                                }
                            }
                        }
                        poly_vote &= ~(1u << poly_sub_ix);
                    }
                }
                if (any_stroke) { /* :441-443 */
                    encode_stroke(enc, rgba, width);
                    enc_log_item(enc, ix, 0, 0);
                }
                break;
            }
            default:
                break;
            }
            v &= v - 1;
        }
    }
    return enc_end(enc);
}

/* ------------------------------------------------------------------------------------------ */
/* renderKernel for one pixel (PietRender.metal:457-566), fp32 throughout                      */
/* ------------------------------------------------------------------------------------------ */
static float srgb_to_linear(uint32_t byte) { /* unpack_unorm4x8_srgb_to_half, colour channels */
    float c = (float)byte / 255.0f;
    return c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}
static void unpack_srgb(uint32_t rgba, float fg[4]) {
    fg[0] = srgb_to_linear(rgba & 0xff);
    fg[1] = srgb_to_linear((rgba >> 8) & 0xff);
    fg[2] = srgb_to_linear((rgba >> 16) & 0xff);
    fg[3] = (float)(rgba >> 24) / 255.0f; /* alpha is linear */
}

/* returns 0 and leaves out[] untouched on Bail (:552-553); 2 on an unknown tag (:555-557) */
static int pmo_pixel(const pmo_cmd *src, uint32_t x, uint32_t y, float out[4]) {
    const float px = (float)x, py = (float)y; /* :467: the pixel's integer corner */
    float rgb[3] = { 1.0f, 1.0f, 1.0f };
    float df = 1e9f;
    float signed_area = 0.0f;
    for (;; src++) {
        uint32_t tag = src->tag;
        if (tag == CMD_END) break;
        switch (tag) {
        case CMD_CIRCLE: { /* :481-493 */
            uint16_t bbox[4]; memcpy(bbox, (const uint8_t *)src + 8, 8);
            float x0 = (float)bbox[0], y0 = (float)bbox[1], x1 = (float)bbox[2], y1 = (float)bbox[3];
            float cx = mixf(x0, x1, 0.5f), cy = mixf(y0, y1, 0.5f);
            float dx = px - cx, dy = py - cy;
            float r = sqrtf(dx * dx + dy * dy);
            float circle_r = fminf(cx - x0, cy - y0);
            float alpha = saturatef(circle_r - r);
            for (int k = 0; k < 3; k++) rgb[k] = mixf(rgb[k], 0.0f, alpha);
            break;
        }
        case CMD_LINE: { /* :495-498 and stroke() :49-55 */
            f2 start, end;
            memcpy(&start, (const uint8_t *)src + 8, 8); memcpy(&end, (const uint8_t *)src + 16, 8);
            float lvx = end.x - start.x, lvy = end.y - start.y;
            float dpx = px - start.x, dpy = py - start.y;
            float t = saturatef((lvx * dpx + lvy * dpy) / (lvx * lvx + lvy * lvy));
            float ex = lvx * t - dpx, ey = lvy * t - dpy;
            float field = sqrtf(ex * ex + ey * ey);
            df = fminf(df, field);
            break;
        }
        case CMD_STROKE: { /* :500-507, renderDf :58-60 */
            float half_width; memcpy(&half_width, &src->body[0], 4);
            float alpha = saturatef(half_width + 0.5f - df);
            float fg[4]; unpack_srgb(src->body[1], fg);
            for (int k = 0; k < 3; k++) rgb[k] = mixf(rgb[k], fg[k], fg[3] * alpha);
            df = 1e9f;
            break;
        }
        case CMD_FILL: { /* :508-529 */
            f2 fs, fe;
            memcpy(&fs, (const uint8_t *)src + 8, 8); memcpy(&fe, (const uint8_t *)src + 16, 8);
            float sx = fs.x - px, sy = fs.y - py;
            float ex = fe.x - px, ey = fe.y - py;
            float wx = saturatef(sy), wy = saturatef(ey);
            if (wx != wy) {
                float tx = (wx - sy) / (ey - sy), ty = (wy - sy) / (ey - sy);
                float xsx = mixf(sx, ex, tx), xsy = mixf(sx, ex, ty);
                float xmin = fminf(fminf(xsx, xsy), 1.0f) - 1e-6f;
                float xmax = fmaxf(xsx, xsy);
                float b = fminf(xmax, 1.0f);
                float c = fmaxf(b, 0.0f);
                float d = fmaxf(xmin, 0.0f);
                float area = (b + 0.5f * (d * d - c * c) - xmin) / (xmax - xmin);
                signed_area += area * (wx - wy);
            }
            break;
        }
        case CMD_FILLEDGE: { /* :530-534 */
            int32_t sign = (int32_t)src->body[0];
            float ey; memcpy(&ey, &src->body[1], 4);
            signed_area += (float)sign * saturatef(py - ey + 1.0f);
            break;
        }
        case CMD_DRAWFILL: { /* :535-545 */
            float alpha = signed_area + (float)(int32_t)src->body[0];
            if (src->body[2] == FILL_EVEN_ODD) alpha = fabsf(alpha - 2.0f * roundf(0.5f * alpha)); /* the formula of :539 */
            else alpha = fminf(fabsf(alpha), 1.0f); /* nonzero winding rule */
            float fg[4]; unpack_srgb(src->body[1], fg);
            for (int k = 0; k < 3; k++) rgb[k] = mixf(rgb[k], fg[k], fg[3] * alpha);
            signed_area = 0.0f;
            break;
        }
        case CMD_SOLID: { /* :546-551 */
            float fg[4]; unpack_srgb(src->body[0], fg);
            for (int k = 0; k < 3; k++) rgb[k] = mixf(rgb[k], fg[k], fg[3]);
            break;
        }
        case CMD_BAIL:
            return 0;
        default:
            out[0] = 1.0f; out[1] = 0.0f; out[2] = 1.0f; out[3] = 1.0f;
            return 2;
        }
    }
    for (int k = 0; k < 3; k++) /* :563 */
        out[k] = rgb[k] < 0.0031308f ? 12.92f * rgb[k] : 1.055f * powf(rgb[k], 1.0f / 2.4f) - 0.055f;
    out[3] = 1.0f;
    return 1;
}

static uint8_t unorm8(float v) { return (uint8_t)lrintf(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); }

/* ------------------------------------------------------------------------------------------ */
/* Public entry points (loaded with ctypes by the tests and bench.py)                          */
/* ------------------------------------------------------------------------------------------ */
static int scene_ok(const uint8_t *scene, size_t len) {
    if (!scene || len < 8) return 0;
    uint64_t n = rd_u32(scene), items_ix = rd_u32(scene + 4);
    if (8 + 8 * n > len || items_ix + 32 * n > len) return 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *it = scene + items_ix + 32 * i;
        uint32_t tag = rd_u32(it);
        if (tag == ITEM_FILL || tag == ITEM_POLY) {
            uint64_t np = rd_u32(it + 12), pix = rd_u32(it + 16);
            if (np == 0 || pix + 8 * np > len) return 0;
        }
    }
    return 1;
}

/*
 * Render tile rows [tile_y0, tile_y1) of a width x height surface.
 *   rgba8    (optional) rows 16*tile_y0 .. min(16*tile_y1, height), stride8 bytes apart, R,G,B,A
 *   rgba32f  (optional) same rows, 4 floats per pixel
 *   offsets/items/solid (optional) the per-tile item lists, row-major over the strip's tiles:
 *            offsets has n_tiles+1 entries; *n_items_out is always set; items are written only
 *            while they fit in cap_items.
 * Returns 0, or -1 for a malformed scene / bad arguments.
 */
int pmo_render(const uint8_t *scene, size_t len, uint32_t width, uint32_t height, uint32_t tile_y0, uint32_t tile_y1,
               uint32_t flags, int n_threads, uint8_t *rgba8, size_t stride8, float *rgba32f, size_t stride32f_bytes,
               uint32_t *offsets, pmo_tile_item *items, size_t cap_items, size_t *n_items_out, uint32_t *solid) {
    if (!scene_ok(scene, len) || width == 0 || height == 0 || width > 65535 || height > 65535) return -1;
    const uint32_t ntx = (width + TILE_W - 1) / TILE_W, nty = (height + TILE_H - 1) / TILE_H;
    if (tile_y1 > nty) tile_y1 = nty;
    if (tile_y0 >= tile_y1) return -1;
    const uint32_t rows = tile_y1 - tile_y0;
    const int want_items = offsets != NULL;
    /* per-row item logs are gathered after the parallel loop to keep the order deterministic */
    pmo_tile_item **row_items = NULL; size_t *row_counts = NULL; uint32_t *tile_counts = NULL;
    if (want_items) {
        row_items = (pmo_tile_item **)calloc(rows, sizeof *row_items);
        row_counts = (size_t *)calloc(rows, sizeof *row_counts);
        tile_counts = (uint32_t *)calloc((size_t)rows * ntx, sizeof *tile_counts);
    }
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    #pragma omp parallel num_threads(n_threads)
    {
        encoder enc; memset(&enc, 0, sizeof enc);
        #pragma omp for schedule(dynamic, 1)
        for (uint32_t r = 0; r < rows; r++) {
            const uint32_t gy = tile_y0 + r;
            size_t rcap = 0;
            for (uint32_t gx = 0; gx < ntx; gx++) {
                uint32_t solid_color = pmo_tile(scene, gx, gy, flags, &enc);
                if (solid) solid[(size_t)r * ntx + gx] = solid_color;
                if (want_items) {
                    if (row_counts[r] + enc.ni > rcap) {
                        rcap = (row_counts[r] + enc.ni) * 2 + 64;
                        row_items[r] = (pmo_tile_item *)realloc(row_items[r], rcap * sizeof(pmo_tile_item));
                    }
                    memcpy(row_items[r] + row_counts[r], enc.items, enc.ni * sizeof(pmo_tile_item));
                    row_counts[r] += enc.ni;
                    tile_counts[(size_t)r * ntx + gx] = (uint32_t)enc.ni;
                }
                if (!rgba8 && !rgba32f) continue;
                for (uint32_t py = 0; py < TILE_H; py++) {
                    uint32_t y = gy * TILE_H + py;
                    if (y >= height) break;
                    for (uint32_t pxi = 0; pxi < TILE_W; pxi++) {
                        uint32_t x = gx * TILE_W + pxi;
                        if (x >= width) break;
                        float px[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
                        uint8_t b[4];
                        /* fragmentShader (:34-44): the solid colour wins when its alpha != 0 */
                        if (solid_color >> 24) {
                            for (int k = 0; k < 4; k++) { b[k] = (uint8_t)(solid_color >> (8 * k)); px[k] = (float)b[k] / 255.0f; }
                        } else {
                            pmo_pixel(enc.cmds, x, y, px);
                            for (int k = 0; k < 4; k++) b[k] = unorm8(px[k]);
                        }
                        size_t ry = (size_t)(y - tile_y0 * TILE_H);
                        if (rgba8) memcpy(rgba8 + ry * stride8 + 4 * (size_t)x, b, 4);
                        if (rgba32f) memcpy((uint8_t *)rgba32f + ry * stride32f_bytes + 16 * (size_t)x, px, 16);
                    }
                }
            }
        }
        free(enc.cmds); free(enc.items);
    }
    if (want_items) {
        size_t total = 0;
        for (uint32_t r = 0; r < rows; r++) {
            size_t k = 0;
            for (uint32_t gx = 0; gx < ntx; gx++) {
                offsets[(size_t)r * ntx + gx] = (uint32_t)total;
                uint32_t cnt = tile_counts[(size_t)r * ntx + gx];
                for (uint32_t q = 0; q < cnt; q++, k++, total++)
                    if (items && total < cap_items) items[total] = row_items[r][k];
            }
            free(row_items[r]);
        }
        offsets[(size_t)rows * ntx] = (uint32_t)total;
        if (n_items_out) *n_items_out = total;
        free(row_items); free(row_counts); free(tile_counts);
    }
    return 0;
}

/* The raw 24-byte command stream of one tile (debug / known-answer tests).  *n_cmds is always set. */
int pmo_tile_cmds(const uint8_t *scene, size_t len, uint32_t tx, uint32_t ty, uint32_t flags, uint8_t *cmds,
                  size_t cap_cmds, size_t *n_cmds, uint32_t *solid_color) {
    if (!scene_ok(scene, len)) return -1;
    encoder enc; memset(&enc, 0, sizeof enc);
    uint32_t sc = pmo_tile(scene, tx, ty, flags, &enc);
    if (solid_color) *solid_color = sc;
    if (n_cmds) *n_cmds = enc.n;
    if (cmds) memcpy(cmds, enc.cmds, (enc.n < cap_cmds ? enc.n : cap_cmds) * sizeof(pmo_cmd));
    free(enc.cmds); free(enc.items);
    return 0;
}

int pmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
