// CUDA kernels of the render path (sm_100a).  Compiled with -fmad=false: see pm_tile_logic.h.
//
//   k_validate     bounds/finite check of an uploaded scene                      (the reference has none)
//   k_plan_*       per-item prefixes: segments, (tile row, 32-tile chunk) units of k_row, backdrop-scratch
//                  words; item table; k_row's unit table                          (once per scene/size/strip)
//   k_pieces_*     per segment: the conservative list of (tile row, candidate tile) "pieces" = k_seg threads
//                  (count, prefix, fill: three kernels)
//   k_seg          one thread per piece: the exact tile tests of TestApp/PietRender.metal:248-445 for one
//                  segment and one tile; appends per-tile records, accumulates the row's backdrop deltas
//   k_row          one warp per (item, tile row, 32-tile chunk): prefix-sums the backdrop deltas and closes
//                  the item per tile -- DrawFill / Solid / opaque cover (64-bit atomic max), Stroke; Line and
//                  Circle items are binned here directly                          (k_seg, k_row: per frame)
//   k_list         one thread per tile: the fill kernels' work lists by class from the final record counts (per frame)
//   k_fine         fill/blend: see pm_fine.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/piet_metal_b200.h"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"
#include "pm_tile_logic.h"

#define PM_FULL_MASK 0xffffffffu

namespace cg = cooperative_groups;

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may start while its
// predecessor in the stream is still draining; pm_grid_wait() blocks until the predecessor has completed
// and its writes are visible, pm_grid_launch_dependents() lets the successor's CTAs be scheduled as this
// grid's CTAs retire.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pm_grid_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pm_grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

namespace {

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
__device__ __forceinline__ float ld_f32(const uint8_t *p) { return *reinterpret_cast<const float *>(p); }
__device__ __forceinline__ float2 ld_f2(const uint8_t *p) { return *reinterpret_cast<const float2 *>(p); }

// ---------------------------------------------------------------------------------------------
// k_validate
// ---------------------------------------------------------------------------------------------
__global__ void k_validate(const uint8_t *scene, uint32_t len, uint32_t *err) {
    if (len < PM_GROUP_HEADER_SIZE) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, 1u); return; }
    const uint64_t n = ld_u32(scene);
    const uint64_t items_ix = ld_u32(scene + 4);
    if (PM_GROUP_HEADER_SIZE + n * PM_BBOX_SIZE > len || (items_ix & 7u) || items_ix + n * PM_ITEM_SIZE > len) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, 1u);
        return;
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < n; i += n_warps) {
        const uint8_t *it = scene + items_ix + i * PM_ITEM_SIZE;
        uint32_t tag = ld_u32(it);
        if (tag == PM_ITEM_FILL || tag == PM_ITEM_POLY) {
            uint64_t np = ld_u32(it + 12), pix = ld_u32(it + 16);
            if (np == 0 || np >= PM_REC_SEG_MAX || (pix & 7u) || pix + np * 8 > len) {
                if (lane == 0) atomicOr(err, 2u);
                continue;
            }
            bool bad = false;
            for (uint64_t k = lane; k < np; k += 32) {
                float x = ld_f32(scene + pix + 8 * k), y = ld_f32(scene + pix + 8 * k + 4);
                if (!isfinite(x) || !isfinite(y)) bad = true;
            }
            if (bad) atomicOr(err, 4u);
            if (tag == PM_ITEM_POLY && lane == 0 && !isfinite(ld_f32(it + PM_POLY_WIDTH))) atomicOr(err, 4u);
        } else if (tag == PM_ITEM_LINE && lane == 0) {
            for (int k = 12; k < 32; k += 4)
                if (!isfinite(ld_f32(it + k))) atomicOr(err, 4u);
        }
    }
}

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// k_plan: per-item prefixes for the two binning kernels (once per scene / size / strip)
//   plan_a[i]  low 32 bits: segments of items < i (k_seg: one thread per segment)
//              high 32 bits: (tile row, 32-tile chunk) pairs of items < i (k_row: one warp each)
//   plan_b[i]  words of the backdrop scratch before item i: rows * pm_bd_stride(tile span) per item
// Items that do not touch the strip count nothing.
// ---------------------------------------------------------------------------------------------
struct ItemSpan { uint32_t tag, r_lo, rows, t_lo, t_hi, n_points; };
// words of one (item, tile row) slice of the backdrop scratch: tile span + 1, rounded up to whole 16-byte units so
// that k_row can sum the deltas of the chunks before its own with 128-bit loads
__device__ __forceinline__ uint32_t pm_bd_stride(uint32_t span) { return (span + 1u + 3u) & ~3u; }

__device__ __forceinline__ ItemSpan item_span(const uint8_t *scene, uint32_t items_ix, uint32_t i, uint32_t tile_y0,
                                              uint32_t tile_y1, uint32_t n_tx) {
    ItemSpan sp;
    const pm_bbox bb = *reinterpret_cast<const pm_bbox *>(scene + PM_GROUP_HEADER_SIZE + (size_t)i * PM_BBOX_SIZE);
    const uint8_t *it = scene + items_ix + (size_t)i * PM_ITEM_SIZE;
    sp.tag = ld_u32(it);
    sp.rows = 0;
    sp.n_points = (sp.tag == PM_ITEM_FILL || sp.tag == PM_ITEM_POLY) ? ld_u32(it + PM_FILL_NPOINTS) : 1u;
    // `hit` (metal:214): bbox.z >= x0 && bbox.x < x0 + 16 && bbox.w >= y0 && bbox.y < y0 + 16
    // <=> tile column in [bbox.x >> 4, bbox.z >> 4] and tile row in [bbox.y >> 4, bbox.w >> 4]
    sp.t_lo = bb.x0 >> 4;
    sp.t_hi = bb.x1 >> 4;
    uint32_t r_lo = bb.y0 >> 4, r_hi = bb.y1 >> 4;
    if (r_lo < tile_y0) r_lo = tile_y0;
    if (r_hi >= tile_y1) r_hi = tile_y1 - 1;  // tile_y1 > tile_y0 >= 0
    sp.r_lo = r_lo;
    if (sp.tag < PM_ITEM_CIRCLE || sp.tag > PM_ITEM_POLY) return sp;
    if (sp.t_lo >= n_tx || sp.t_hi < sp.t_lo || r_hi < r_lo) return sp;
    if (sp.t_hi > n_tx - 1) sp.t_hi = n_tx - 1;
    sp.rows = r_hi - r_lo + 1;
    return sp;
}

// block-wide inclusive scan of one u64 per thread (1024 threads); returns the exclusive value and the block total
__device__ __forceinline__ u64 block_scan_excl(u64 v, u64 *warp_excl, u64 *block_total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(PM_FULL_MASK, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_excl[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u64 ws = warp_excl[lane], wi = ws;
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(PM_FULL_MASK, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        warp_excl[lane] = wi - ws;
        if (lane == 31) *block_total = wi;
    }
    __syncthreads();
    u64 r = warp_excl[warp] + incl - v;
    __syncthreads();
    return r;
}

// Three kernels (the plan is redone for every scene that is uploaded, i.e. inside every end-to-end frame; as one
// CTA looping over the items and writing every item's k_row units one after the other it took 0.3 ms for the tiger's
// 304 wide items and 1.9 ms for 100 k glyphs):
//   k_plan_count  one thread per item: its counts (into plan_a / plan_b), its PmItemInfo and its linear colour
//   k_plan_scan   one CTA: the counts become exclusive prefixes in place (every thread a contiguous chunk)
//   k_plan_rows   one thread per k_row unit: its PmRowInfo (the item by binary search in the prefixes)
__global__ void __launch_bounds__(256) k_plan_count(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                                    uint32_t tile_y1, uint32_t n_tx, u64 *plan_a, u64 *plan_b, PmItemInfo *item_info,
                                                    const float *srgb_lut, float4 *item_paint) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const ItemSpan sp = item_span(scene, items_ix, i, tile_y0, tile_y1, n_tx);
    u64 ca = 0, cb = 0;
    if (sp.rows) {
        const uint32_t n_seg = sp.tag == PM_ITEM_FILL ? sp.n_points : (sp.tag == PM_ITEM_POLY ? sp.n_points - 1u : 0u);
        ca = ((u64)(sp.rows * ((sp.t_hi - sp.t_lo + 32u) / 32u)) << 32) | n_seg;  // rows x 32-tile chunks
        if (sp.tag == PM_ITEM_FILL || sp.tag == PM_ITEM_POLY) cb = (u64)sp.rows * pm_bd_stride(sp.t_hi - sp.t_lo + 1u);
    }
    plan_a[i] = ca;
    plan_b[i] = cb;
    PmItemInfo ii;
    ii.t_lo = sp.t_lo; ii.t_hi = sp.t_hi; ii.r_lo = sp.r_lo; ii.rows = sp.rows; ii.bd_base = 0;
    ii.rgba = 0; ii.tag_flags = sp.tag; ii.w0 = 0;
    // the item's colour as the fill kernels blend it: unpack_unorm4x8_srgb_to_half (metal:503, :541, :548)
    // through the look-up table; Cmd_Circle paints opaque black (metal:491)
    float4 paint = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    if (sp.tag == PM_ITEM_LINE || sp.tag == PM_ITEM_FILL || sp.tag == PM_ITEM_POLY) {
        const uint8_t *it = scene + items_ix + (size_t)i * PM_ITEM_SIZE;
        const uint32_t rgba = ld_u32(it + (sp.tag == PM_ITEM_POLY ? PM_POLY_RGBA : PM_FILL_RGBA));
        paint = make_float4(srgb_lut[rgba & 0xffu], srgb_lut[(rgba >> 8) & 0xffu], srgb_lut[(rgba >> 16) & 0xffu], srgb_lut[256u + (rgba >> 24)]);
        ii.rgba = rgba;
        if (sp.tag == PM_ITEM_FILL && (ld_u32(it + PM_FILL_FLAGS) & PM_FILL_EVEN_ODD) != 0) ii.tag_flags |= PM_INFO_EVEN_ODD;
        if (sp.tag == PM_ITEM_POLY) ii.w0 = pm_f2u(0.5f * ld_f32(it + PM_POLY_WIDTH));
        if (sp.tag == PM_ITEM_LINE) ii.w0 = pm_f2u(0.5f * ld_f32(it + PM_LINE_WIDTH));
    }
    item_info[i] = ii;
    item_paint[i] = paint;
}

// plan_a / plan_b [0, n_items): counts -> exclusive prefixes, in place; [n_items] = the totals; result (device)
__global__ void __launch_bounds__(1024) k_plan_scan(uint32_t n_items, u64 *plan_a, u64 *plan_b, PmPlanResult *result) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 total_lo, total_hi, total_b;
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (n_items + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = tid * chunk < n_items ? tid * chunk : n_items;
    const uint32_t hi = lo + chunk < n_items ? lo + chunk : n_items;
    // the two halves of plan_a are summed apart: a carry out of the low half (or 2^31 in either half) would corrupt
    // the packed prefixes and has to be reported, not wrapped
    u64 s_lo = 0, s_hi = 0, s_b = 0;
    for (uint32_t i = lo; i < hi; i++) { const u64 a = plan_a[i]; s_lo += a & 0xffffffffull; s_hi += a >> 32; s_b += plan_b[i]; }
    u64 r_lo = block_scan_excl(s_lo, warp_excl, &total_lo);
    u64 r_hi = block_scan_excl(s_hi, warp_excl, &total_hi);
    u64 r_b = block_scan_excl(s_b, warp_excl, &total_b);
    const bool bad = total_lo >= 0x80000000ull || total_hi >= 0x80000000ull;
    for (uint32_t i = lo; i < hi && !bad; i++) {
        const u64 a = plan_a[i], b = plan_b[i];
        plan_a[i] = (r_hi << 32) | r_lo;
        plan_b[i] = r_b;
        r_lo += a & 0xffffffffull; r_hi += a >> 32; r_b += b;
    }
    if (tid == 0) {
        if (bad) { result->error = 1; return; }
        plan_a[n_items] = (total_hi << 32) | total_lo;
        plan_b[n_items] = total_b;
        result->n_segments = (uint32_t)total_lo;
        result->n_rows = (uint32_t)total_hi;
        result->bd_words = total_b;
    }
}

// largest i with (plan_a[i] >> 32) <= u: the item that owns k_row unit u (items without units share their successor's prefix)
__device__ __forceinline__ uint32_t item_of_unit(const u64 *plan_a, uint32_t n_items, uint32_t u) {
    uint32_t lo = 0, hi = n_items;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)(plan_a[mid] >> 32) <= u) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_plan_rows(uint32_t n_items, uint32_t n_units, const u64 *plan_a, const u64 *plan_b,
                                                   const PmItemInfo *item_info, PmRowInfo *row_info) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const uint32_t i = item_of_unit(plan_a, n_items, u);
    const PmItemInfo ii = item_info[i];
    const uint32_t span = ii.t_hi - ii.t_lo + 1u, chunks = (span + 31u) / 32u;
    const uint32_t k = u - (uint32_t)(plan_a[i] >> 32), r = k / chunks, c = k - r * chunks;
    PmRowInfo ri;
    ri.item = i; ri.row_chunk = ((ii.r_lo + r) << 16) | c;
    ri.bd_row = (uint32_t)plan_b[i] + r * pm_bd_stride(span);
    ri.t_lo_span = ii.t_lo | (span << 16);
    ri.rgba = ii.rgba; ri.tag_flags = ii.tag_flags; ri.w0 = ii.w0; ri.pad = 0;
    row_info[u] = ri;
}

// One segment of a Fill / Poly item and the tile rows of the strip it can reach.  For a Fill the
// range is exact (rows with mxy >= y0 && mny < y0 + 16: division by 16 and floor are exact); for a
// Poly it is conservative and pm_poly_segment_row applies the exact y test.
struct SegCtx { ItemSpan sp; PmSeg sg; float hw; int ra, rb; };

__device__ __forceinline__ SegCtx load_segment(const uint8_t *scene, uint32_t items_ix, uint32_t item, uint32_t k, uint32_t tile_y0,
                                               uint32_t tile_y1, uint32_t n_tx) {
    SegCtx c;
    c.sp = item_span(scene, items_ix, item, tile_y0, tile_y1, n_tx);
    const uint8_t *it = scene + items_ix + (size_t)item * PM_ITEM_SIZE;
    const uint8_t *pts = scene + ld_u32(it + PM_FILL_POINTS_IX);  // same offset in both variants
    const float2 s = ld_f2(pts + 8 * (size_t)k);
    c.hw = 0.0f;
    if (c.sp.tag == PM_ITEM_FILL) {
        const float2 e = ld_f2(pts + 8 * (size_t)(k + 1 == c.sp.n_points ? 0 : k + 1));  // closing segment, metal:262
        c.sg = pm_seg(s.x, s.y, e.x, e.y);
        c.ra = pm_floor_i(c.sg.mny * (1.0f / 16.0f));
        c.rb = pm_floor_i(c.sg.mxy * (1.0f / 16.0f));
    } else {
        const float2 e = ld_f2(pts + 8 * (size_t)(k + 1));
        c.sg = pm_seg(s.x, s.y, e.x, e.y);
        c.hw = 0.5f * ld_f32(it + PM_POLY_WIDTH) + 0.5f;
        c.ra = pm_floor_i((c.sg.mny - c.hw) * (1.0f / 16.0f)) - 1;
        c.rb = pm_floor_i((c.sg.mxy + c.hw) * (1.0f / 16.0f)) + 1;
    }
    const int r_first = (int)c.sp.r_lo, r_last = (int)(c.sp.r_lo + c.sp.rows) - 1;
    if (c.ra < r_first) c.ra = r_first;
    if (c.rb > r_last) c.rb = r_last;
    return c;
}

__device__ __forceinline__ uint32_t item_of_segment(const u64 *plan_a, uint32_t n_items, uint32_t g) {
    uint32_t lo = 0, hi = n_items;
    while (hi - lo > 1) {  // largest i with segment prefix <= g
        uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)plan_a[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// Work decomposition of k_seg: one "piece" per (segment, tile row, candidate tile) -- plus one for a
// (segment, tile row) that has no candidate tile but may still carry backdrop.  The candidate tiles
// are the conservative span of pm_*_candidate_span; the exact tests run in k_seg every frame.
// piece_info[q] = (segment, first << 31 | has_tile << 30 | tile row << 15 | tile column).
// A long flat segment simply becomes many pieces, a tall one too: no thread of k_seg does more than one
// tile's worth of work.
#define PM_PIECE_FIRST 0x80000000u
#define PM_PIECE_TILE 0x40000000u

// Three kernels (the plan is redone for every scene that is uploaded, i.e. inside every end-to-end frame):
//   k_pieces_count  one thread per segment: its PmSegInfo and the number of its pieces
//   k_pieces_scan   one CTA: exclusive prefix of the counts (each thread sums a contiguous chunk)
//   k_pieces_fill   one thread per segment: writes its pieces at its offset
__device__ __forceinline__ SegCtx segment_of(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1,
                                             uint32_t n_tx, const u64 *plan_a, uint32_t g, uint32_t *item_out) {
    const uint32_t item = item_of_segment(plan_a, n_items, g);
    *item_out = item;
    return load_segment(scene, items_ix, item, g - (uint32_t)plan_a[item], tile_y0, tile_y1, n_tx);
}

__global__ void __launch_bounds__(256) k_pieces_count(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                                      uint32_t tile_y1, uint32_t n_tx, const u64 *plan_a, const u64 *plan_b,
                                                      uint32_t n_segments, PmSegInfo *seg_info, uint32_t *piece_cnt) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_segments) return;
    uint32_t item;
    const SegCtx c = segment_of(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, g, &item);
    PmSegInfo si;
    si.sx = c.sg.sx; si.sy = c.sg.sy; si.ex = c.sg.ex; si.ey = c.sg.ey;
    si.item = item; si.k = g - (uint32_t)plan_a[item]; si.hw = c.hw; si.tag = c.sp.tag;
    si.t_lo = c.sp.t_lo; si.t_hi = c.sp.t_hi; si.r_lo = c.sp.r_lo; si.bd_base = (uint32_t)plan_b[item];
    seg_info[g] = si;
    u64 cnt = 0;
    for (int r = c.ra; r <= c.rb; r++) {
        uint32_t ta = 1, tb = 0;
        const float y0 = (float)(r * PM_TILE_H);
        const bool has = c.sp.tag == PM_ITEM_FILL ? pm_fill_candidate_span(c.sg, y0, c.sp.t_lo, c.sp.t_hi, &ta, &tb)
                                                   : pm_poly_candidate_span(c.sg, y0, c.hw, c.sp.t_lo, c.sp.t_hi, &ta, &tb);
        cnt += has ? (u64)(tb - ta + 1) : 1ull;
    }
    piece_cnt[g] = cnt > 0xffffffffull ? 0xffffffffu : (uint32_t)cnt;  // (saturated: the scan reports the overflow)
}

// piece_cnt[g] -> exclusive prefix, in place; result->n_pieces = total; result->error if it does not fit 31 bits
__global__ void __launch_bounds__(1024) k_pieces_scan(uint32_t *piece_cnt, uint32_t n_segments, PmPlanResult *result) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 total;
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (n_segments + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = tid * chunk < n_segments ? tid * chunk : n_segments;
    const uint32_t hi = lo + chunk < n_segments ? lo + chunk : n_segments;
    u64 sum = 0;
    for (uint32_t g = lo; g < hi; g++) sum += piece_cnt[g];
    u64 run = block_scan_excl(sum, warp_excl, &total);
    for (uint32_t g = lo; g < hi; g++) {
        const uint32_t c = piece_cnt[g];
        piece_cnt[g] = (uint32_t)run;
        run += c;
    }
    if (tid == 0) {
        if (total >= 0x80000000ull) result->error = 1;
        result->n_pieces = (uint32_t)total;
    }
}

__global__ void __launch_bounds__(256) k_pieces_fill(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                                     uint32_t tile_y1, uint32_t n_tx, const u64 *plan_a, uint32_t n_segments,
                                                     const uint32_t *piece_off, uint2 *piece_info, uint32_t piece_cap) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_segments) return;
    uint32_t item;
    const SegCtx c = segment_of(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, g, &item);
    u64 q = piece_off[g];
    for (int r = c.ra; r <= c.rb; r++) {
        uint32_t ta = 1, tb = 0;
        const float y0 = (float)(r * PM_TILE_H);
        const bool has = c.sp.tag == PM_ITEM_FILL ? pm_fill_candidate_span(c.sg, y0, c.sp.t_lo, c.sp.t_hi, &ta, &tb)
                                                   : pm_poly_candidate_span(c.sg, y0, c.hw, c.sp.t_lo, c.sp.t_hi, &ta, &tb);
        if (!has) {
            if (q < piece_cap) piece_info[q] = make_uint2(g, PM_PIECE_FIRST | ((uint32_t)r << 15));
            q++;
        } else {
            for (uint32_t t = ta; t <= tb; t++, q++)
                if (q < piece_cap) piece_info[q] = make_uint2(g, (t == ta ? PM_PIECE_FIRST : 0u) | PM_PIECE_TILE | ((uint32_t)r << 15) | t);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_seg + k_row: binning
// ---------------------------------------------------------------------------------------------
// Claims the next record slot of a tile for this frame: returns its position (0 for the first).
// cnt word = stamp << 32 | count.  The atomic max lifts a leftover of an earlier frame (smaller
// stamp) to (stamp, 0) and leaves a current word alone; the add that follows from the same thread
// to the same address is ordered after it, so it always sees this frame's stamp.  No CAS loop:
// a tile that many segments hit at once would make one quadratic in the number of contenders.
__device__ __forceinline__ uint32_t tile_claim_slot(u64 *word, uint32_t stamp) {
    atomicMax(word, (u64)stamp << 32);
    return (uint32_t)atomicAdd(word, 1ull);
}

// Where the exact tile tests of one (item, tile row) send their results.  bd is the row's slice of
// the backdrop scratch: word j <-> tile t_lo + j; bit 0 = "the item has a command in this tile",
// bits 1.. = 2 * (backdrop delta at this tile, difference-encoded along the row).
struct BinSink {
    const PmFrameArgs &A;
    uint32_t *bd;
    uint32_t t_lo;
    uint32_t row_tile0;  // index of the row's first tile
    uint32_t item;

    // bump-allocates overflow block j (header + slots); returns 1 + the header's pool index, or PM_EXT_FAILED
    __device__ __forceinline__ uint32_t alloc_block(uint32_t ovf_region, uint32_t j) const {
        const uint32_t size = pm_blk_size(j) + 1u;
        const uint32_t o = atomicAdd(&A.counters->n_overflow, size);
        if (o + size > A.overflow_cap || o + size < o) return PM_EXT_FAILED;
        const uint32_t mine = ovf_region + o + 1u;
        A.pool[mine - 1u].next = 0;
        __threadfence();
        return mine;
    }
    __device__ __forceinline__ void append(uint32_t t, PmRecord r) {
        const uint32_t tile = row_tile0 + t;
        const uint32_t pos = tile_claim_slot(&A.cnt[tile], A.stamp);
        uint32_t idx;
        r.next = 0;
        if (pos < PM_TILE_SLOTS) {
            idx = tile * PM_TILE_SLOTS + pos;
        } else {
            // Later records go into the tile's chain of overflow blocks (layout: pm_pixel_logic.h).  Lock-free and
            // without waiting: whoever finds a link unpublished allocates the block and tries to install it with a
            // compare-and-swap; a loser adopts the winner's block (its own stays unused -- bump allocation cannot
            // give it back; the count the host sees includes it).
            uint32_t j, off;
            pm_ovf_locate(pos - PM_TILE_SLOTS, &j, &off);
            const uint32_t ovf_region = A.n_rows * A.n_tx * PM_TILE_SLOTS;
            uint32_t base1;  // 1 + pool index of the header of block 0, or PM_EXT_FAILED
            {
                const u64 seen = *reinterpret_cast<volatile u64 *>(&A.ovf[tile]);
                if ((uint32_t)(seen >> 32) == A.stamp) {
                    base1 = (uint32_t)seen;
                } else {
                    const uint32_t mine = alloc_block(ovf_region, 0);
                    const u64 prev = atomicCAS(&A.ovf[tile], seen, ((u64)A.stamp << 32) | (u64)mine);
                    base1 = prev == seen ? mine : (uint32_t)prev;  // (a word that changed was published by somebody else, this frame)
                }
            }
            for (uint32_t k = 0; k < j && base1 != PM_EXT_FAILED; k++) {
                uint32_t *link = &A.pool[base1 - 1u].next;
                uint32_t nxt = *reinterpret_cast<volatile uint32_t *>(link);
                if (nxt == 0) {
                    const uint32_t mine = alloc_block(ovf_region, k + 1u);
                    const uint32_t prev = atomicCAS(link, 0u, mine);
                    nxt = prev == 0 ? mine : prev;
                }
                base1 = nxt;
            }
            if (base1 == PM_EXT_FAILED) return;  // the host sees n_overflow > overflow_cap, grows the pool and renders the frame again
            idx = base1 + off;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(&A.pool[idx]);
        const uint4 *src = reinterpret_cast<const uint4 *>(&r);
        dst[0] = src[0];
        dst[1] = src[1];
    }
    __device__ __forceinline__ void fill(uint32_t t, uint32_t seg, const PmFillEmit &e, const PmSeg &g) {
        append(t, pm_rec_fill(item, seg, t, e, g));
        atomicOr(&bd[t - t_lo], 1u);
    }
    __device__ __forceinline__ void backdrop(uint32_t ta, uint32_t tb, int delta) {
        atomicAdd(&bd[ta - t_lo], (uint32_t)(2 * delta));
        atomicAdd(&bd[tb + 1 - t_lo], (uint32_t)(-2 * delta));
    }
    __device__ __forceinline__ void line(uint32_t t, uint32_t seg, const PmSeg &g) {
        append(t, pm_rec_line(item, seg, g));
        atomicOr(&bd[t - t_lo], 1u);
    }
    __device__ __forceinline__ void trailer(uint32_t t, uint32_t kind, uint32_t seg, uint32_t w0, uint32_t w1) {
        append(t, pm_rec_words(item, kind, seg, w0, w1));
    }
};

// One thread per piece (see k_pieces_count / k_pieces_fill): the exact tile tests of TestApp/PietRender.metal:248-445
// of one segment for one candidate tile; the first piece of a (segment, tile row) also adds the
// row's backdrop intervals.
#ifndef PM_SEG_CTAS
#define PM_SEG_CTAS 5  // resident CTAs per SM k_seg is compiled for (5: 43 registers as the compiler likes it)
#endif
// PM_DEBUG_SEG=1: per-CTA [start, end] in globaltimer ns, two words per CTA (k_seg from word 0, k_row from word 2^18;
// tools/grid_timeline.py)
struct CtaTimer {
    unsigned long long *slot;
    __device__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ CtaTimer(unsigned long long *debug, uint32_t base) : slot(debug && (threadIdx.x & 31) == 0 ? debug + base + 2 * blockIdx.x : nullptr) { if (slot) atomicMin(slot, now()); }  // (one lane per warp)
    __device__ ~CtaTimer() { if (slot) atomicMax(slot + 1, now()); }
};
#ifndef PM_SEG_THREADS
#define PM_SEG_THREADS 256
#endif
__global__ void __launch_bounds__(PM_SEG_THREADS, PM_SEG_CTAS * 256 / PM_SEG_THREADS) k_seg(const PmFrameArgs A) {
    CtaTimer timer(A.debug, 0);
    pm_grid_launch_dependents();
    pm_grid_wait();  // the previous frame's fill kernel is done with the queues, the lists and the scratch
    {   // the backdrop scratch the NEXT frame will use (last touched by the frame before this one): no memset in the frame
        uint4 *z = reinterpret_cast<uint4 *>(A.bd_next);
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.bd_quads; i += (unsigned long long)gridDim.x * blockDim.x)
            z[i] = make_uint4(0, 0, 0, 0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.queue->batch_next = 0;
        A.queue->tile_next = 0;
        A.queue->heavy_next = 0;
        A.queue->heavy_warp_next = 0;
    }
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.n_pieces) return;
    const uint2 pi = A.piece_info[q];
    const uint4 *sp4 = reinterpret_cast<const uint4 *>(&A.seg_info[pi.x]);
    const uint4 s0 = sp4[0], s1 = sp4[1], s2 = sp4[2];  // sx sy ex ey | item k hw tag | t_lo t_hi r_lo bd_base
    const PmSeg sg = pm_seg(pm_u2f(s0.x), pm_u2f(s0.y), pm_u2f(s0.z), pm_u2f(s0.w));
    const uint32_t item = s1.x, k = s1.y, t_lo = s2.x, t_hi = s2.y, r_lo = s2.z;
    const u64 bd_base = s2.w;
    const uint32_t row = (pi.y >> 15) & 0x7fffu, t = pi.y & 0x7fffu;
    const float y0 = (float)(row * PM_TILE_H);
    BinSink sink{A, A.bd + bd_base + (size_t)(row - r_lo) * pm_bd_stride(t_hi - t_lo + 1u), t_lo, (row - A.tile_y0) * A.n_tx, item};
    if (s1.w == PM_ITEM_FILL) {
        if (pi.y & PM_PIECE_FIRST) pm_fill_backdrop_row(sink, sg, y0, t_lo, t_hi, A.n_tx);
        if (pi.y & PM_PIECE_TILE) pm_fill_candidate_tile(sink, sg, y0, t, k);
    } else if (pi.y & PM_PIECE_TILE) {
        pm_poly_candidate_tile(sink, sg, y0, pm_u2f(s1.z), t, k, (A.flags & PM_FLAG_FIX_POLY_PRECULL) != 0);
    }
}

// One warp per (item, tile row, chunk of 32 tiles): closes what k_seg accumulated -- DrawFill /
// Solid / opaque cover per tile of a Fill item (metal:359-363), Stroke per tile of a Poly item
// (metal:441-443).  Line and Circle items have no segments and are binned here directly
// (metal:218-247).  The scratch alternates between two buffers; k_seg clears the one the next frame will use.
#ifndef PM_ROW_WARPS
#define PM_ROW_WARPS 8
#endif
#ifndef PM_ROW_PERSISTENT
#define PM_ROW_PERSISTENT 1  // at most one resident wave of CTAs, every warp walks its units with the next unit's entry in flight
#endif
#ifndef PM_ROW_CTAS
#define PM_ROW_CTAS (2048 / (PM_ROW_WARPS * 32))  // resident CTAs per SM k_row is compiled for
#endif
__global__ void __launch_bounds__(PM_ROW_WARPS * 32, PM_ROW_CTAS) k_row(const PmFrameArgs A) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CtaTimer timer(A.debug, 1u << 18);
    pm_grid_launch_dependents();
    pm_grid_wait();  // k_seg has finished
    const uint32_t stride = gridDim.x * PM_ROW_WARPS;
    uint32_t unit = blockIdx.x * PM_ROW_WARPS + warp;
    if (unit >= A.n_row_units) return;
    // (a unit is a chain of dependent accesses -- its entry, the backdrop words, the slot claim, the record -- and the
    // kernel is as long as the chains it runs one after the other: the next unit's entry is requested before this
    // unit's work begins)
    uint4 n0 = reinterpret_cast<const uint4 *>(&A.row_info[unit])[0], n1 = reinterpret_cast<const uint4 *>(&A.row_info[unit])[1];
  for (;;) {
    const uint4 r0 = n0, r1 = n1;  // item, row << 16 | chunk, bd_row, t_lo | span << 16 || rgba, tag_flags, w0  (k_plan_rows)
    const uint32_t next = unit + stride;
    if (next < A.n_row_units) { n0 = reinterpret_cast<const uint4 *>(&A.row_info[next])[0]; n1 = reinterpret_cast<const uint4 *>(&A.row_info[next])[1]; }
    const uint32_t item = r0.x;
    const uint32_t tag = r1.y & 0xffu, rgba = r1.x;
    const uint8_t *it = A.scene + A.items_ix + (size_t)item * PM_ITEM_SIZE;
    const uint32_t span = r0.w >> 16;
    const uint32_t row = r0.y >> 16, j0 = (r0.y & 0xffffu) * 32u;
    const uint32_t t_lo = r0.w & 0xffffu;
    const float y0 = (float)(row * PM_TILE_H);
    const uint32_t *bd = A.bd + r0.z;
    BinSink sink{A, nullptr, t_lo, (row - A.tile_y0) * A.n_tx, item};
    const uint32_t j = j0 + lane;

    if (tag == PM_ITEM_FILL) {
        // PM_FLAG_FILL_RULES: the item's flags word may ask for the even-odd rule (extension; the reference ignores it)
        const bool even_odd = (A.flags & PM_FLAG_FILL_RULES) != 0 && (r1.y & PM_INFO_EVEN_ODD) != 0;
        // backdrop entering this chunk: sum of the deltas of the tiles before it
        const uint32_t v = j < span ? bd[j] : 0u;
        int carry = 0;
        if (j0) {  // (most items are narrower than 32 tiles: one chunk, nothing before it)
            #pragma unroll 1
            for (uint32_t q = lane * 4u; q < j0; q += 128u) {  // (the slice starts on a 16-byte boundary: pm_bd_stride)
                const uint4 d = *reinterpret_cast<const uint4 *>(&bd[q]);
                carry += ((int)d.x >> 1) + ((int)d.y >> 1) + ((int)d.z >> 1) + ((int)d.w >> 1);
            }
            carry = __reduce_add_sync(PM_FULL_MASK, carry);
        }
        int incl = (int)v >> 1;
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(PM_FULL_MASK, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        const int backdrop = carry + incl;
        if (j < span) {
            const uint32_t t = t_lo + j;
            if (v & 1u) {
                sink.trailer(t, even_odd ? PM_REC_DRAWFILL_EO : PM_REC_DRAWFILL, PM_REC_SEG_MAX, (uint32_t)backdrop, rgba);
            } else if (even_odd ? (backdrop & 1) != 0 : backdrop != 0) {  // covered: nonzero winding number / odd winding number
                if ((rgba & 0xff000000u) == 0xff000000u) {  // opaque full cover: rewinds the tile (metal:132-135)
                    atomicMax(&A.occ[sink.row_tile0 + t], ((u64)A.stamp << 32) | (u64)(item + 1u));
                } else {
                    sink.trailer(t, PM_REC_SOLID, 0, 0, rgba);
                }
            }
        }
    } else if (tag == PM_ITEM_POLY) {
        if (j < span && (bd[j] & 1u)) sink.trailer(t_lo + j, PM_REC_STROKE, PM_REC_SEG_MAX, r1.z, rgba);
    } else if (tag == PM_ITEM_LINE) {  // metal:223-247
        const float width = ld_f32(it + PM_LINE_WIDTH);
        const float2 s = ld_f2(it + PM_LINE_START), e = ld_f2(it + PM_LINE_END);
        const PmSeg g = pm_seg(s.x, s.y, e.x, e.y);
        const float hw = 0.5f * width + 0.5f;
        if (j < span) {
            const uint32_t t = t_lo + j;
            const float x0 = (float)(t * PM_TILE_W);
            if (pm_stroke_cross(g, x0, x0 + 16.0f, y0, y0 + 16.0f, hw)) {
                sink.append(t, pm_rec_line(item, 0, g));
                sink.trailer(t, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * width), rgba);
            }
        }
    } else if (tag == PM_ITEM_CIRCLE) {  // metal:218-222
        const pm_bbox bb = *reinterpret_cast<const pm_bbox *>(A.scene + PM_GROUP_HEADER_SIZE + (size_t)item * PM_BBOX_SIZE);
        const uint32_t b_lo = (uint32_t)bb.x0 | ((uint32_t)bb.y0 << 16), b_hi = (uint32_t)bb.x1 | ((uint32_t)bb.y1 << 16);
        if (j < span) sink.trailer(t_lo + j, PM_REC_CIRCLE, 0, b_lo, b_hi);
    }
    if (next >= A.n_row_units) break;
    unit = next;
  }
}

// One thread per tile, after binning: the fill kernels' work lists from the FINAL record counts -- four disjoint
// classes (heavy for k_heavy; medium, mid, low for k_fine, which walks them in that order: long jobs first, the
// cheapest last), appended warp-wise (one atomic per warp and class) so that the entries stay in tile order.
// (Binning used to append a tile to a list when its count crossed a threshold: three more atomic paths, and a second
// dependent atomic, on k_seg's and k_row's critical path; tiles listed twice and skipped after their prefetch in k_fine.)
#ifndef PM_LIST_SCRAMBLE
#define PM_LIST_SCRAMBLE 0
#endif
#define PM_LIST_THREADS 1024
__global__ void __launch_bounds__(PM_LIST_THREADS) k_list(const PmFrameArgs A) {
    __shared__ uint32_t s_cnt[4][PM_LIST_THREADS / 32];
    pm_grid_launch_dependents();
    pm_grid_wait();  // k_row has finished
    const uint32_t n_tiles = A.n_rows * A.n_tx;
    uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if PM_LIST_SCRAMBLE
    // (experiment) list order decoupled from tile order: thread -> tile through a multiplicative permutation of the
    // power-of-two range that covers the tiles
    {
        uint32_t bits = 1;
        while ((1u << bits) < n_tiles) bits++;
        const uint32_t mask = (1u << bits) - 1u;
        tile = tile <= mask ? (tile * PM_LIST_SCRAMBLE) & mask : tile;
    }
#endif
    uint32_t n = 0;
    if (tile < n_tiles) {
        const u64 cw = A.cnt[tile];
        n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    }
    if (!__syncthreads_or(n != 0)) return;
    const uint32_t cls = n >= PM_HEAVY_MIN ? 1u : (n >= PM_MEDIUM_MIN ? 2u : (n >= PM_MID_MIN ? 3u : 0u));  // quarter of complex_list
    // One atomic per CTA and class (a counter serves an atomic every few cycles: one per warp -- 8,192 of them on
    // the same address at 8192^2 -- made this kernel as long as k_seg).  Per warp: the class masks; warp c then turns
    // the 32 per-warp counts of class c into offsets behind one reservation.
    uint32_t my_mask = 0;
    #pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        const uint32_t m = __ballot_sync(PM_FULL_MASK, n != 0 && cls == c);
        if (cls == c) my_mask = m;
        if (lane == c) s_cnt[c][warp] = (uint32_t)__popc(m);
    }
    __syncthreads();
    if (warp < 4) {
        const uint32_t v = s_cnt[warp][lane];
        uint32_t incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(PM_FULL_MASK, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        uint32_t base = 0;
        if (lane == 31 && incl) {
            uint32_t *counter = warp == 0 ? &A.counters->n_low : (warp == 1 ? &A.counters->n_heavy : (warp == 2 ? &A.counters->n_medium : &A.counters->n_mid));
            base = atomicAdd(counter, incl);
        }
        base = __shfl_sync(PM_FULL_MASK, base, 31);
        s_cnt[warp][lane] = base + incl - v;
    }
    __syncthreads();
    if (n != 0) {
        const uint32_t row = tile / A.n_tx, entry = (row << 16) | (tile - row * A.n_tx);  // (strip-local tile row, tile column)
        A.complex_list[(size_t)cls * n_tiles + s_cnt[cls][warp] + __popc(my_mask & ((1u << lane) - 1u))] = entry;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
void pm_launch_validate(const uint8_t *scene, uint32_t scene_len, uint32_t *err, cudaStream_t s) {
    k_validate<<<296, 256, 0, s>>>(scene, scene_len, err);
}

void pm_launch_plan(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1,
                    uint32_t n_tx, unsigned long long *plan_a, unsigned long long *plan_b, PmItemInfo *item_info,
                    const float *srgb_lut, float4 *item_paint, PmPlanResult *result, cudaStream_t s) {
    if (n_items) k_plan_count<<<(n_items + 255) / 256, 256, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, plan_b, item_info, srgb_lut, item_paint);
    k_plan_scan<<<1, 1024, 0, s>>>(n_items, plan_a, plan_b, result);
}

void pm_launch_plan_rows(uint32_t n_items, uint32_t n_units, const unsigned long long *plan_a, const unsigned long long *plan_b,
                         const PmItemInfo *item_info, PmRowInfo *row_info, cudaStream_t s) {
    if (n_units) k_plan_rows<<<(n_units + 255) / 256, 256, 0, s>>>(n_items, n_units, plan_a, plan_b, item_info, row_info);
}

void pm_launch_pieces_count(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                            const unsigned long long *plan_a, const unsigned long long *plan_b, uint32_t n_segments, PmSegInfo *seg_info,
                            uint32_t *piece_cnt, PmPlanResult *result, cudaStream_t s) {
    if (n_segments) k_pieces_count<<<(n_segments + 255) / 256, 256, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, plan_b, n_segments, seg_info, piece_cnt);
    k_pieces_scan<<<1, 1024, 0, s>>>(piece_cnt, n_segments, result);
}

void pm_launch_pieces_fill(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                           const unsigned long long *plan_a, uint32_t n_segments, const uint32_t *piece_off, uint2 *piece_info,
                           uint32_t piece_cap, cudaStream_t s) {
    if (n_segments) k_pieces_fill<<<(n_segments + 255) / 256, 256, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, n_segments, piece_off, piece_info, piece_cap);
}

template <class K>
static cudaError_t launch_overlapped(K kernel, dim3 grid, dim3 block, bool overlap, cudaStream_t s, const PmFrameArgs &a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

cudaError_t pm_launch_frame(const PmFrameArgs &a, int sm_count, cudaEvent_t mid, cudaEvent_t mid2, bool overlap, cudaStream_t s,
                            uint32_t *n_launched) {
    cudaError_t e;
    uint32_t launched = 0;
    uint32_t grid_seg = (a.n_pieces + PM_SEG_THREADS - 1u) / PM_SEG_THREADS;
    if (grid_seg == 0) grid_seg = 1;  // still clears the fill kernels' queues
    if ((e = launch_overlapped(k_seg, dim3(grid_seg), dim3(PM_SEG_THREADS), overlap, s, a)) != cudaSuccess) return e;
    launched++;
    // (k_row runs even without units when the launches overlap: the chain of grid dependencies must not skip a kernel)
    if (a.n_row_units || overlap) {
        uint32_t grid_row = (a.n_row_units + PM_ROW_WARPS - 1) / PM_ROW_WARPS;
#if PM_ROW_PERSISTENT
        const uint32_t row_wave = (uint32_t)sm_count * PM_ROW_CTAS;  // one resident wave
        if (grid_row > row_wave) grid_row = row_wave;
#endif
        if ((e = launch_overlapped(k_row, dim3(grid_row ? grid_row : 1), dim3(PM_ROW_WARPS * 32), overlap, s, a)) != cudaSuccess) return e;
        launched++;
    }
    {
        const uint32_t n_tiles = a.n_rows * a.n_tx;
        uint32_t n_threads = n_tiles;
#if PM_LIST_SCRAMBLE
        while (n_threads & (n_threads - 1u)) n_threads += n_threads & (0u - n_threads);  // next power of two
#endif
        if ((e = launch_overlapped(k_list, dim3((n_threads + PM_LIST_THREADS - 1u) / PM_LIST_THREADS), dim3(PM_LIST_THREADS), overlap, s, a)) != cudaSuccess) return e;
        launched++;
    }
    if (mid && (e = cudaEventRecord(mid, s)) != cudaSuccess) return e;
    // k_heavy before k_fine: its CTAs (one per heavy tile, the frame's longest jobs) start first; k_fine is released
    // as soon as they have seen binning complete and runs beside them (see the notes on griddepcontrol in both kernels)
    if ((e = pm_launch_heavy(a, sm_count, overlap, s)) != cudaSuccess) return e;
    launched++;
    if (mid2 && (e = cudaEventRecord(mid2, s)) != cudaSuccess) return e;
    if ((e = pm_launch_fine(a, sm_count, overlap, s)) != cudaSuccess) return e;
    launched++;
    if (n_launched) *n_launched = launched;
    return cudaSuccess;
}
