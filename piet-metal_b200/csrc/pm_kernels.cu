// CUDA kernels of the render path (sm_100a).  Compiled with -fmad=false: see pm_tile_logic.h.
//
//   k_validate   bounds/finite check of an uploaded scene            (the reference has none)
//   k_plan       per-item tile-row counts -> work-unit prefix        (once per scene/size/strip)
//   k_bin        one warp per (item, tile row): exact tile tests of TestApp/PietRender.metal:160-454
//                evaluated per segment and row; appends per-tile records, accumulates backdrops,
//                resolves opaque full covers with a 64-bit atomic max   (per frame)
//   k_fine       fill/blend: one warp per tile with records -- renderKernel's arithmetic
//                (metal:457-566) evaluated sparsely: lanes take (record, pixel row) pairs and add
//                fixed-point coverage into shared memory, then 8 pixels per lane are blended in
//                registers and stored -- and 32-tile batches of solid tiles written with full
//                512-byte rows of 128-bit stores (the fused solid-tile composite, metal:16-44)
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
    if (PM_GROUP_HEADER_SIZE + n * PM_BBOX_SIZE > len || (items_ix & 3u) || items_ix + n * PM_ITEM_SIZE > len) {
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
            if (np == 0 || np >= PM_REC_SEG_MAX || (pix & 3u) || pix + np * 8 > len) {
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
//   plan_b[i]  words of the backdrop scratch before item i: rows * (tile span + 1) per item
// Items that do not touch the strip count nothing.
// ---------------------------------------------------------------------------------------------
struct ItemSpan { uint32_t tag, r_lo, rows, t_lo, t_hi, n_points; };

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

__global__ void __launch_bounds__(1024) k_plan(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                               uint32_t tile_y1, uint32_t n_tx, u64 *plan_a, u64 *plan_b, PmItemInfo *item_info,
                                               uint2 *row_info, uint32_t row_info_cap, PmPlanResult *result) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 total_a, total_b, carry_a, carry_b;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n_items; base += blockDim.x) {
        const uint32_t i = base + tid;
        u64 ca = 0, cb = 0;
        if (i < n_items) {
            const ItemSpan sp = item_span(scene, items_ix, i, tile_y0, tile_y1, n_tx);
            if (sp.rows) {
                const uint32_t n_seg = sp.tag == PM_ITEM_FILL ? sp.n_points : (sp.tag == PM_ITEM_POLY ? sp.n_points - 1u : 0u);
                ca = ((u64)(sp.rows * ((sp.t_hi - sp.t_lo + 32u) / 32u)) << 32) | n_seg;  // rows x 32-tile chunks
                if (sp.tag == PM_ITEM_FILL || sp.tag == PM_ITEM_POLY) cb = (u64)sp.rows * (sp.t_hi - sp.t_lo + 2u);
            }
        }
        const u64 ea = carry_a + block_scan_excl(ca, warp_excl, &total_a);
        const u64 eb = carry_b + block_scan_excl(cb, warp_excl, &total_b);
        if (i < n_items) {
            plan_a[i] = ea;
            plan_b[i] = eb;
            {
                const ItemSpan spi = item_span(scene, items_ix, i, tile_y0, tile_y1, n_tx);
                PmItemInfo ii;
                ii.t_lo = spi.t_lo; ii.t_hi = spi.t_hi; ii.r_lo = spi.r_lo; ii.rows = spi.rows; ii.bd_base = eb; ii.pad[0] = ii.pad[1] = 0;
                item_info[i] = ii;
            }
            // second pass (row_info given): tabulate the item's (tile row, 32-tile chunk) units for k_row
            if (row_info && ca != 0) {
                const ItemSpan sp = item_span(scene, items_ix, i, tile_y0, tile_y1, n_tx);
                const uint32_t chunks = (sp.t_hi - sp.t_lo + 32u) / 32u;
                uint32_t u = (uint32_t)(ea >> 32);
                for (uint32_t r = 0; r < sp.rows; r++)
                    for (uint32_t c = 0; c < chunks && u < row_info_cap; c++, u++) row_info[u] = make_uint2(i, ((sp.r_lo + r) << 16) | c);
            }
        }
        // a carry out of the low half (or 2^31 in either half) would corrupt the packed prefixes
        if (((ea + ca) & 0x8000000080000000ull) != 0) result->error = 1;
        __syncthreads();
        if (tid == 0) { carry_a += total_a; carry_b += total_b; }
        __syncthreads();
    }
    if (tid == 0) {
        plan_a[n_items] = carry_a;
        plan_b[n_items] = carry_b;
        result->n_segments = (uint32_t)carry_a;
        result->n_rows = (uint32_t)(carry_a >> 32);
        result->bd_words = carry_b;
    }
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
// Pass 1 (piece_info == nullptr) only counts; pass 2 tabulates.  A long flat segment simply becomes
// many pieces, a tall one too: no thread of k_seg does more than one tile's worth of work.
#define PM_PIECE_FIRST 0x80000000u
#define PM_PIECE_TILE 0x40000000u

__global__ void __launch_bounds__(1024) k_plan_pieces(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                                      uint32_t tile_y1, uint32_t n_tx, const u64 *plan_a, uint32_t n_segments,
                                                      PmSegInfo *seg_info, uint2 *piece_info, uint32_t piece_cap, PmPlanResult *result) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 total, carry;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_segments; base += blockDim.x) {
        const uint32_t g = base + tid;
        u64 cnt = 0;
        SegCtx c;
        c.ra = 1; c.rb = 0;
        if (g < n_segments) {
            const uint32_t item = item_of_segment(plan_a, n_items, g);
            c = load_segment(scene, items_ix, item, g - (uint32_t)plan_a[item], tile_y0, tile_y1, n_tx);
            PmSegInfo si;
            si.sx = c.sg.sx; si.sy = c.sg.sy; si.ex = c.sg.ex; si.ey = c.sg.ey;
            si.item = item; si.k = g - (uint32_t)plan_a[item]; si.hw = c.hw; si.tag = c.sp.tag;
            seg_info[g] = si;
            for (int r = c.ra; r <= c.rb; r++) {
                uint32_t ta = 1, tb = 0;
                const float y0 = (float)(r * PM_TILE_H);
                const bool has = c.sp.tag == PM_ITEM_FILL ? pm_fill_candidate_span(c.sg, y0, c.sp.t_lo, c.sp.t_hi, &ta, &tb)
                                                           : pm_poly_candidate_span(c.sg, y0, c.hw, c.sp.t_lo, c.sp.t_hi, &ta, &tb);
                cnt += has ? (u64)(tb - ta + 1) : 1ull;
            }
        }
        const u64 e = carry + block_scan_excl(cnt, warp_excl, &total);
        if (g < n_segments && piece_info) {
            u64 q = e;
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
        if (e + cnt >= 0x80000000ull) result->error = 1;
        __syncthreads();
        if (tid == 0) carry += total;
        __syncthreads();
    }
    if (tid == 0) result->n_pieces = (uint32_t)carry;
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

    __device__ __forceinline__ void append(uint32_t t, PmRecord r) {
        const uint32_t tile = row_tile0 + t;
        const uint32_t pos = tile_claim_slot(&A.cnt[tile], A.stamp);
        uint32_t idx;
        r.next = 0;
        if (pos < PM_TILE_SLOTS) {
            idx = tile * PM_TILE_SLOTS + pos;
        } else {
            // one counter serves every overflow record of the frame: aggregate the lanes that are here together
            cg::coalesced_group og = cg::coalesced_threads();
            uint32_t o = 0;
            if (og.thread_rank() == 0) o = atomicAdd(&A.counters->n_overflow, og.size());
            o = og.shfl(o, 0) + og.thread_rank();
            if (o >= A.overflow_cap) return;  // the host sees n_overflow > overflow_cap, grows the pool and re-renders
            idx = A.n_rows * A.n_tx * PM_TILE_SLOTS + o;
            u64 prev = atomicExch(&A.ovf[tile], ((u64)A.stamp << 32) | (u64)(idx + 1u));
            if ((uint32_t)(prev >> 32) == A.stamp) r.next = (uint32_t)prev;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(&A.pool[idx]);
        const uint4 *src = reinterpret_cast<const uint4 *>(&r);
        dst[0] = src[0];
        dst[1] = src[1];
        if (pos == PM_TILE_SLOTS) {  // first overflow record: the tile is "heavy", the fill kernel renders those first
            const uint32_t h = atomicAdd(&A.counters->n_heavy, 1u);
            A.complex_list[A.n_rows * A.n_tx + h] = ((tile - t) / A.n_tx << 16) | t;
        }
        if (pos == 0) {  // first record of the tile this frame: queue it for the fill kernel
            cg::coalesced_group g = cg::coalesced_threads();
            uint32_t base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(&A.counters->n_complex, g.size());
            base = g.shfl(base, 0);
            A.complex_list[base + g.thread_rank()] = ((tile - t) / A.n_tx << 16) | t;  // (strip-local tile row, tile column)
        }
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

// One thread per piece (see k_plan_pieces): the exact tile tests of TestApp/PietRender.metal:248-445
// of one segment for one candidate tile; the first piece of a (segment, tile row) also adds the
// row's backdrop intervals.
__global__ void __launch_bounds__(256) k_seg(const PmFrameArgs A) {
    // PM_DEBUG_SEG=1: per-CTA [start, end] in globaltimer ns, two words per CTA
    struct Timer {
        const PmFrameArgs &A;
        __device__ static unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
        __device__ Timer(const PmFrameArgs &a) : A(a) { if (A.debug) atomicMin(&A.debug[2 * blockIdx.x], now()); }
        __device__ ~Timer() { if (A.debug) atomicMax(&A.debug[2 * blockIdx.x + 1], now()); }
    } timer(A);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.queue->complex_next = 0;
        A.queue->batch_next = 0;
        A.queue->heavy_next = 0;
    }
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.n_pieces) return;
    const uint2 pi = A.piece_info[q];
    const uint4 *sp4 = reinterpret_cast<const uint4 *>(&A.seg_info[pi.x]);
    const uint4 s0 = sp4[0], s1 = sp4[1];  // sx sy ex ey | item k hw tag
    const uint4 *ip4 = reinterpret_cast<const uint4 *>(&A.item_info[s1.x]);
    const uint4 i0 = ip4[0], i1 = ip4[1];  // t_lo t_hi r_lo rows | bd_base
    const PmSeg sg = pm_seg(pm_u2f(s0.x), pm_u2f(s0.y), pm_u2f(s0.z), pm_u2f(s0.w));
    const uint32_t item = s1.x, k = s1.y, t_lo = i0.x, t_hi = i0.y, r_lo = i0.z;
    const u64 bd_base = ((u64)i1.y << 32) | i1.x;
    const uint32_t row = (pi.y >> 15) & 0x7fffu, t = pi.y & 0x7fffu;
    const float y0 = (float)(row * PM_TILE_H);
    BinSink sink{A, A.bd + bd_base + (size_t)(row - r_lo) * (t_hi - t_lo + 2u), t_lo, (row - A.tile_y0) * A.n_tx, item};
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
// (metal:218-247).  The scratch is cleared by a memset at the start of the next frame.
#define PM_ROW_WARPS 8
__global__ void __launch_bounds__(PM_ROW_WARPS * 32) k_row(const PmFrameArgs A) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * PM_ROW_WARPS + warp;
    if (unit >= A.n_row_units) return;
    const uint2 ri = A.row_info[unit];  // (item, tile row << 16 | chunk), tabulated by k_plan
    const uint32_t item = ri.x;
    const ItemSpan sp = item_span(A.scene, A.items_ix, item, A.tile_y0, A.tile_y0 + A.n_rows, A.n_tx);
    const uint8_t *it = A.scene + A.items_ix + (size_t)item * PM_ITEM_SIZE;
    const uint32_t span = sp.t_hi - sp.t_lo + 1;
    const uint32_t row = ri.y >> 16, j0 = (ri.y & 0xffffu) * 32u;
    const uint32_t t_lo = sp.t_lo;
    const float y0 = (float)(row * PM_TILE_H);
    const uint32_t *bd = A.bd + A.plan_b[item] + (size_t)(row - sp.r_lo) * (span + 1);
    BinSink sink{A, nullptr, t_lo, (row - A.tile_y0) * A.n_tx, item};
    const uint32_t j = j0 + lane;

    if (sp.tag == PM_ITEM_FILL) {
        const uint32_t rgba = ld_u32(it + PM_FILL_RGBA);
        // backdrop entering this chunk: sum of the deltas of the tiles before it
        int carry = 0;
        for (uint32_t q = lane; q < j0; q += 32) carry += (int)bd[q] >> 1;
        carry = __reduce_add_sync(PM_FULL_MASK, carry);
        const uint32_t v = j < span ? bd[j] : 0u;
        int incl = (int)v >> 1;
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(PM_FULL_MASK, incl, o);
            if (lane >= (uint32_t)o) incl += u;
        }
        const int backdrop = carry + incl;
        if (j < span) {
            const uint32_t t = t_lo + j;
            if (v & 1u) {
                sink.trailer(t, PM_REC_DRAWFILL, PM_REC_SEG_MAX, (uint32_t)backdrop, rgba);
            } else if (backdrop != 0) {
                if ((rgba & 0xff000000u) == 0xff000000u) {  // opaque full cover: rewinds the tile (metal:132-135)
                    atomicMax(&A.occ[sink.row_tile0 + t], ((u64)A.stamp << 32) | (u64)(item + 1u));
                } else {
                    sink.trailer(t, PM_REC_SOLID, 0, 0, rgba);
                }
            }
        }
    } else if (sp.tag == PM_ITEM_POLY) {
        if (j < span && (bd[j] & 1u))
            sink.trailer(t_lo + j, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * ld_f32(it + PM_POLY_WIDTH)), ld_u32(it + PM_POLY_RGBA));
    } else if (sp.tag == PM_ITEM_LINE) {  // metal:223-247
        const uint32_t rgba = ld_u32(it + PM_LINE_RGBA);
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
    } else if (sp.tag == PM_ITEM_CIRCLE) {  // metal:218-222
        const pm_bbox bb = *reinterpret_cast<const pm_bbox *>(A.scene + PM_GROUP_HEADER_SIZE + (size_t)item * PM_BBOX_SIZE);
        const uint32_t b_lo = (uint32_t)bb.x0 | ((uint32_t)bb.y0 << 16), b_hi = (uint32_t)bb.x1 | ((uint32_t)bb.y1 << 16);
        if (j < span) sink.trailer(t_lo + j, PM_REC_CIRCLE, 0, b_lo, b_hi);
    }
}

// ---------------------------------------------------------------------------------------------
// k_fine
// ---------------------------------------------------------------------------------------------
#define PM_FINE_WARPS 8
#define PM_FINE_COMPLEX_WARPS 6    // warps that prefer tiles with records; the rest prefer solid batches
#define PM_FINE_LIST_CAP 256       // records per tile indexed in shared memory; the rest is re-walked
#define PM_ACC_STRIDE 17

// Per-warp shared-memory state of the tile being rendered.
struct FineWarpSmem {
    int acc[16 * PM_ACC_STRIDE];     // near-pixel coverage, 8.24 fixed point
    int cov[16 * PM_ACC_STRIDE];     // per-row cover deltas (pixel x and everything right of it)
    float dmin[16 * PM_ACC_STRIDE];  // stroke distance field
    uint32_t idx[PM_FINE_LIST_CAP];  // pool indices of the tile's records
};

struct FineAcc {
    FineWarpSmem *w;
    __device__ __forceinline__ void near(int row, int j, int fx) { atomicAdd(&w->acc[row * PM_ACC_STRIDE + j], fx); }
    __device__ __forceinline__ void cover(int row, int j, int fx) { atomicAdd(&w->cov[row * PM_ACC_STRIDE + j], fx); }
    __device__ __forceinline__ void dist(int row, int j, float d) {  // d >= 0: unsigned order == float order
        atomicMin(reinterpret_cast<unsigned int *>(&w->dmin[row * PM_ACC_STRIDE + j]), __float_as_uint(d));
    }
};

template <bool EXACT>
__device__ __forceinline__ float linear_to_srgb(float v) {  // metal:563
    if (v < 0.0031308f) return 12.92f * v;
    // default: ex2(lg2(v) / 2.4) on the SFU, a few 1e-7 from powf; PM_FLAG_EXACT_SRGB asks for powf
    float p;
    if (EXACT) {
        p = powf(v, 1.0f / 2.4f);
    } else {  // v in [0.003, ~1]: no denormals, no special cases
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
        l *= 1.0f / 2.4f;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(l));
    }
    return 1.055f * p - 0.055f;
}

// Linear -> sRGB for one pixel, packed RGBA8 (alpha 255).  Out of line: 8 call sites per tile.
template <bool EXACT>
__device__ __noinline__ uint32_t encode_pixel(float r, float g, float b) {
    return pm_unorm8(linear_to_srgb<EXACT>(r)) | (pm_unorm8(linear_to_srgb<EXACT>(g)) << 8) |
           (pm_unorm8(linear_to_srgb<EXACT>(b)) << 16) | 0xff000000u;
}

// lut[0..255]: sRGB byte -> linear; lut[256..511]: alpha byte / 255 (unpack_unorm4x8_srgb_to_half)
__device__ __forceinline__ void unpack_fg(const float *lut, uint32_t rgba, float fg[4]) {
    fg[0] = lut[rgba & 0xffu];
    fg[1] = lut[(rgba >> 8) & 0xffu];
    fg[2] = lut[(rgba >> 16) & 0xffu];
    fg[3] = lut[256u + (rgba >> 24)];
}

__device__ __forceinline__ PmRecord load_record(const PmRecord *pool, uint32_t idx) {
    PmRecord r;
    const uint4 *src = reinterpret_cast<const uint4 *>(&pool[idx]);
    uint4 a = src[0], b = src[1];
    r.item = a.x; r.key = a.y; r.p[0] = pm_u2f(a.z); r.p[1] = pm_u2f(a.w);
    r.p[2] = pm_u2f(b.x); r.p[3] = pm_u2f(b.y); r.edge_y = pm_u2f(b.z); r.next = b.w;
    return r;
}

// Phase A for up to 32 records held one per lane (`mine` = this lane holds a FILL*/LINE record of
// the current item): the (record, pixel row) pairs are enumerated across the lanes and each lane
// adds its pair's coverage / distance into shared memory.
__device__ __noinline__ void fine_pairs(FineAcc &acc, bool mine, const PmRecord &r, bool stroke, float reach,
                                        float tile_x0, float tile_y0, uint32_t lane) {
    const uint32_t kind = r.key & 15u;
    int ra = 1, rb = 0;
    if (mine) {
        if (stroke) pm_line_rows(r.p[1], r.p[3], reach, tile_y0, &ra, &rb);
        else pm_fill_rows(r.p[1], r.p[3], tile_y0, &ra, &rb);
    }
    const int cnt = rb >= ra ? rb - ra + 1 : 0;
    int incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(PM_FULL_MASK, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    const int excl = incl - cnt;
    const int total = __shfl_sync(PM_FULL_MASK, incl, 31);
    for (int q = (int)lane; q - (int)lane < total; q += 32) {
        // owner = last lane whose exclusive prefix is <= q
        int lo = 0;
        #pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            int cand = lo + step;
            int v = __shfl_sync(PM_FULL_MASK, excl, cand & 31);
            if (cand < 32 && v <= q) lo = cand;
        }
        const int o_excl = __shfl_sync(PM_FULL_MASK, excl, lo);
        const int o_ra = __shfl_sync(PM_FULL_MASK, ra, lo);
        float p[4];
        p[0] = __shfl_sync(PM_FULL_MASK, r.p[0], lo);
        p[1] = __shfl_sync(PM_FULL_MASK, r.p[1], lo);
        p[2] = __shfl_sync(PM_FULL_MASK, r.p[2], lo);
        p[3] = __shfl_sync(PM_FULL_MASK, r.p[3], lo);
        if (q < total) {
            const int row = o_ra + (q - o_excl);
            if (stroke) pm_line_pair(acc, p, reach, row, tile_x0, tile_y0);
            else pm_fill_pair(acc, p, row, tile_x0, tile_y0);
        }
    }
    // FillEdge commands: one record at a time, lanes 0..15 take the 16 pixel rows
    if (!stroke) {
        for (uint32_t em = __ballot_sync(PM_FULL_MASK, mine && kind != PM_REC_FILL); em != 0; em &= em - 1) {
            const int src = __ffs(em) - 1;
            const uint32_t e_kind = __shfl_sync(PM_FULL_MASK, kind, src);
            const float e_y = __shfl_sync(PM_FULL_MASK, r.edge_y, src);
            if (lane < 16) pm_fill_edge_row(acc, e_kind, e_y, (int)lane, tile_y0);
        }
    }
}

// One tile that owns records.  All 32 lanes execute this together.  Records are handled in chunks
// of 32, one per lane; the first chunk (all of them, for nearly every tile) stays in registers.
// Blend/store layout: lane l owns pixel row (l >> 1), pixels 8*(l & 1) .. +7.
template <bool F32, bool EXACT, bool GENERAL>
__device__ __forceinline__ void fine_complex_tile_impl(const PmFrameArgs &A, uint32_t packed_tile, u64 cw, u64 ow, FineWarpSmem *w, const float *lut, uint32_t lane) {
    const uint32_t trow = packed_tile >> 16, tx = packed_tile & 0xffffu;
    const uint32_t tile = trow * A.n_tx + tx;
    const uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;

    // index the records: inline slots first, then the overflow chain.  n_cached counts what was
    // actually found (a frame whose overflow pool ran out has fewer links than cnt says; the host
    // re-renders such a frame, it only must not fault).
    const uint32_t n_inline = n < PM_TILE_SLOTS ? n : PM_TILE_SLOTS;
    uint32_t n_cached = n_inline;
    uint32_t tail = 0;  // 1 + pool index of the first record that did not fit the shared-memory index
    if (GENERAL && n > PM_TILE_SLOTS) {
        if (lane < PM_TILE_SLOTS) w->idx[lane] = tile * PM_TILE_SLOTS + lane;
        const u64 vw = A.ovf[tile];
        uint32_t cur = (uint32_t)(vw >> 32) == A.stamp ? (uint32_t)vw : 0u;
        while (cur != 0 && n_cached < PM_FINE_LIST_CAP) {
            if (lane == 0) w->idx[n_cached] = cur - 1u;
            cur = A.pool[cur - 1u].next;
            n_cached++;
        }
        tail = cur;
        __syncwarp();
    }
    const uint32_t n_chunks = GENERAL ? (n_cached + 31u) >> 5 : 1u;
    // chunk 0 lives in registers for the whole tile
    PmRecord r0;
    r0.item = 0xffffffffu; r0.key = 0; r0.p[0] = r0.p[1] = r0.p[2] = r0.p[3] = 0.0f; r0.edge_y = 0.0f; r0.next = 0;
    if (lane < n_cached) r0 = load_record(A.pool, (GENERAL && n > PM_TILE_SLOTS) ? w->idx[lane] : tile * PM_TILE_SLOTS + lane);
    if (r0.item < occ_item1) r0.item = 0xffffffffu;  // below the topmost opaque cover: rewound away (metal:132-135)

    bool has_draw = r0.item != 0xffffffffu && (r0.key & 15u) != PM_REC_SOLID;
    for (uint32_t c = 1; GENERAL && c < n_chunks; c++) {
        const uint32_t i = c * 32u + lane;
        if (i < n_cached) {
            const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[w->idx[i]]);
            if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
        }
    }
    for (uint32_t cur = tail; GENERAL && cur != 0; cur = A.pool[cur - 1u].next) {
        const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[cur - 1u]);
        if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
    }
    has_draw = __any_sync(PM_FULL_MASK, has_draw);

    const uint32_t prow = lane >> 1, half = lane & 1u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + half * 8u) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) +
                     (tx * PM_TILE_W + half * 8u);

    uint32_t occ_rgba = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (occ_item1) occ_rgba = ld_u32(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA);

    if (!has_draw) {
        // Only Solid commands after the last rewind: the tile Bails and shows solidColor (metal:145-147, :34-44)
        const uint32_t c = occ_rgba;
        const uint4 v = make_uint4(c, c, c, c);
        __stcs(reinterpret_cast<uint4 *>(dst), v);
        __stcs(reinterpret_cast<uint4 *>(dst) + 1, v);
        if (F32) {
            const float4 f = make_float4((float)(c & 0xff) / 255.0f, (float)((c >> 8) & 0xff) / 255.0f,
                                         (float)((c >> 16) & 0xff) / 255.0f, (float)(c >> 24) / 255.0f);
            for (int j = 0; j < 8; j++) dst32[j] = f;
        }
        return;
    }

    float rgb[8][3];
    #pragma unroll
    for (int j = 0; j < 8; j++) rgb[j][0] = rgb[j][1] = rgb[j][2] = 1.0f;  // metal:470
    if (occ_item1) {  // the rewound list starts with the cover's Cmd_Solid (metal:136-142, :546-551): same for every pixel
        float fg[4];
        unpack_fg(lut, occ_rgba, fg);
        const float b0 = pm_mix(1.0f, fg[0], fg[3]), b1 = pm_mix(1.0f, fg[1], fg[3]), b2 = pm_mix(1.0f, fg[2], fg[3]);
        #pragma unroll
        for (int j = 0; j < 8; j++) { rgb[j][0] = b0; rgb[j][1] = b1; rgb[j][2] = b2; }
    }
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);  // scene coordinates
    const float px0 = tile_x0 + (float)(half * 8u), py = tile_y0 + (float)prow;
    FineAcc acc{w};
    int *my_acc = &w->acc[prow * PM_ACC_STRIDE + half * 8u];
    int *my_cov = &w->cov[prow * PM_ACC_STRIDE + half * 8u];
    float *my_dmin = &w->dmin[prow * PM_ACC_STRIDE + half * 8u];

    // items in painter's order: repeatedly take the smallest item id above the last one done
    uint32_t last_item = 0;
    bool first = true;
    for (;;) {
        uint32_t cur_item = (first || r0.item > last_item) ? r0.item : 0xffffffffu;
        for (uint32_t c = 1; GENERAL && c < n_chunks; c++) {
            const uint32_t i = c * 32u + lane;
            if (i < n_cached) {
                const uint32_t it = A.pool[w->idx[i]].item;
                if (it >= occ_item1 && (first || it > last_item) && it < cur_item) cur_item = it;
            }
        }
        for (uint32_t cur = tail; GENERAL && cur != 0; cur = A.pool[cur - 1u].next) {
            const uint32_t it = A.pool[cur - 1u].item;
            if (it >= occ_item1 && (first || it > last_item) && it < cur_item) cur_item = it;
        }
        cur_item = __reduce_min_sync(PM_FULL_MASK, cur_item);
        if (cur_item == 0xffffffffu) break;
        first = false;
        last_item = cur_item;

        // the item's closing record says what it is (DrawFill / Stroke / Circle / Solid)
        uint32_t t_kind = 0, t_w0 = 0, t_w1 = 0;
        if (r0.item == cur_item && (r0.key & 15u) >= PM_REC_CIRCLE) { t_kind = r0.key & 15u; t_w0 = pm_f2u(r0.p[0]); t_w1 = pm_f2u(r0.p[1]); }
        for (uint32_t c = 1; GENERAL && c < n_chunks; c++) {
            const uint32_t i = c * 32u + lane;
            if (i < n_cached) {
                const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[w->idx[i]]);
                if (a.x == cur_item && (a.y & 15u) >= PM_REC_CIRCLE) { t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w; }
            }
        }
        for (uint32_t cur = tail; GENERAL && cur != 0; cur = A.pool[cur - 1u].next) {
            const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[cur - 1u]);
            if (a.x == cur_item && (a.y & 15u) >= PM_REC_CIRCLE) { t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w; }
        }
        {
            const uint32_t src = __ffs(__ballot_sync(PM_FULL_MASK, t_kind != 0));
            if (src == 0) continue;  // cannot happen for a well-formed list
            t_kind = __shfl_sync(PM_FULL_MASK, t_kind, src - 1);
            t_w0 = __shfl_sync(PM_FULL_MASK, t_w0, src - 1);
            t_w1 = __shfl_sync(PM_FULL_MASK, t_w1, src - 1);
        }

        // per-pixel blend factor of this item for the lane's 8 pixels, then one shared blend
        float fg[4] = {0.0f, 0.0f, 0.0f, 1.0f};  // Cmd_Circle paints black (metal:491)
        float alpha[8];
        if (t_kind == PM_REC_DRAWFILL || t_kind == PM_REC_STROKE) {
            const bool stroke = t_kind == PM_REC_STROKE;
            const float half_width = pm_u2f(t_w0);
            const float reach = half_width + 0.5f;
            // phase A: coverage of the item's segments, 32 records at a time
            for (uint32_t c = 0; c < n_chunks; c++) {
                PmRecord rc = r0;
                if (GENERAL && c > 0) {
                    const uint32_t i = c * 32u + lane;
                    rc.item = 0xffffffffu;
                    if (i < n_cached) rc = load_record(A.pool, w->idx[i]);
                }
                const bool mine = rc.item == cur_item && (rc.key & 15u) <= PM_REC_LINE;
                if (__any_sync(PM_FULL_MASK, mine)) fine_pairs(acc, mine, rc, stroke, reach, tile_x0, tile_y0, lane);
            }
            for (uint32_t cur = tail; GENERAL && cur != 0;) {  // records beyond the shared-memory index, one at a time
                PmRecord rc = load_record(A.pool, cur - 1u);
                cur = rc.next;
                if (rc.item == cur_item && (rc.key & 15u) <= PM_REC_LINE) fine_pairs(acc, lane == 0, rc, stroke, reach, tile_x0, tile_y0, lane);
            }
            __syncwarp();
            // phase B: resolve this lane's 8 pixels
            unpack_fg(lut, t_w1, fg);
            if (!stroke) {
                int covs[8], accs[8], run = 0;
                #pragma unroll
                for (int j = 0; j < 8; j++) { covs[j] = my_cov[j]; accs[j] = my_acc[j]; my_cov[j] = 0; my_acc[j] = 0; run += covs[j]; }
                const int other = __shfl_xor_sync(PM_FULL_MASK, run, 1);
                run = half ? other : 0;  // covers of the left half carry into the right half
                const int backdrop = (int)t_w0;
                #pragma unroll
                for (int j = 0; j < 8; j++) {
                    run += covs[j];
                    alpha[j] = fg[3] * pm_resolve_fill_alpha(accs[j] + run, backdrop);
                }
            } else {
                #pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float df = my_dmin[j];
                    my_dmin[j] = 1e9f;
                    alpha[j] = fg[3] * pm_saturate(half_width + 0.5f - df);  // renderDf, metal:58-60
                }
            }
            __syncwarp();
        } else if (t_kind == PM_REC_CIRCLE) {
            #pragma unroll 1
            for (int j = 0; j < 8; j++) {
                const float a = pm_px_circle_alpha(t_w0, t_w1, px0 + (float)j, py);
                #pragma unroll
                for (int jj = 0; jj < 8; jj++) if (jj == j) alpha[jj] = a;
            }
        } else {  // PM_REC_SOLID: a translucent full cover
            unpack_fg(lut, t_w1, fg);
            #pragma unroll
            for (int j = 0; j < 8; j++) alpha[j] = fg[3];
        }
        #pragma unroll
        for (int j = 0; j < 8; j++)
            #pragma unroll
            for (int k = 0; k < 3; k++) rgb[j][k] = pm_mix(rgb[j][k], fg[k], alpha[j]);
    }

    uint32_t packed[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        packed[j] = encode_pixel<EXACT>(rgb[j][0], rgb[j][1], rgb[j][2]);
        if (F32)  // debug render: the un-quantised values
            dst32[j] = make_float4(linear_to_srgb<EXACT>(rgb[j][0]), linear_to_srgb<EXACT>(rgb[j][1]), linear_to_srgb<EXACT>(rgb[j][2]), 1.0f);
    }
    __stcs(reinterpret_cast<uint4 *>(dst), make_uint4(packed[0], packed[1], packed[2], packed[3]));
    __stcs(reinterpret_cast<uint4 *>(dst) + 1, make_uint4(packed[4], packed[5], packed[6], packed[7]));
}

// One copy of the tile code for every record count: a second, leaner copy for small tiles was tried
// and lost -- the kernel is instruction-cache bound and two warm copies thrash it.
template <bool F32, bool EXACT>
__device__ __forceinline__ void fine_complex_tile(const PmFrameArgs &A, uint32_t packed_tile, u64 cw, u64 ow, FineWarpSmem *w, const float *lut, uint32_t lane) {
    fine_complex_tile_impl<F32, EXACT, true>(A, packed_tile, cw, ow, w, lut, lane);
}

// 32 consecutive tiles of one tile row; the solid ones are written row-wise: each store
// instruction covers 512 contiguous bytes (128 pixels) of one pixel row.
template <bool F32>
__device__ void fine_solid_batch(const PmFrameArgs &A, uint32_t batch, uint32_t batches_per_row, uint32_t lane) {
    const uint32_t row = batch / batches_per_row;
    const uint32_t t0 = (batch - row * batches_per_row) * 32u;
    const uint32_t t = t0 + lane;
    const bool valid = t < A.n_tx;
    bool solid = false;
    uint32_t colour = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (valid) {
        const size_t tile = (size_t)row * A.n_tx + t;
        const u64 cw = A.cnt[tile], ow = A.occ[tile];
        solid = !((uint32_t)(cw >> 32) == A.stamp && (uint32_t)cw != 0u);
        if (solid && (uint32_t)(ow >> 32) == A.stamp && (uint32_t)ow != 0u)
            colour = ld_u32(A.scene + A.items_ix + (size_t)((uint32_t)ow - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA);
    }
    const uint32_t solid_mask = __ballot_sync(PM_FULL_MASK, solid);
    if (solid_mask == 0) return;
    #pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const uint32_t src = (uint32_t)q * 8u + (lane >> 2);
        const uint32_t c = __shfl_sync(PM_FULL_MASK, colour, src);
        if (!((solid_mask >> src) & 1u)) continue;
        const uint4 v = make_uint4(c, c, c, c);
        uint8_t *dst = A.fb + (size_t)(row * PM_TILE_H) * A.pitch + ((size_t)t0 * PM_TILE_W + (size_t)q * 128u + lane * 4u) * 4u;
        #pragma unroll 4
        for (int y = 0; y < PM_TILE_H; y++) __stcs(reinterpret_cast<uint4 *>(dst + (size_t)y * A.pitch), v);  // streaming: keep L2 for the records
        if (F32) {
            float4 f = make_float4((float)(c & 0xff) / 255.0f, (float)((c >> 8) & 0xff) / 255.0f,
                                   (float)((c >> 16) & 0xff) / 255.0f, (float)(c >> 24) / 255.0f);
            for (int y = 0; y < PM_TILE_H; y++)
                for (int xx = 0; xx < 4; xx++) {
                    float4 *d = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) +
                        (size_t)(row * PM_TILE_H + y) * A.pitch32) + (t0 * PM_TILE_W + q * 128u + lane * 4u + xx);
                    *d = f;
                }
        }
    }
}

template <bool F32, bool EXACT>
__global__ void __launch_bounds__(PM_FINE_WARPS * 32, 3) k_fine(const PmFrameArgs A) {
    __shared__ float s_lut[512];
    __shared__ FineWarpSmem s_warp[PM_FINE_WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s_lut[i] = A.srgb_lut[i];
    FineWarpSmem *w = &s_warp[warp];
    for (uint32_t i = lane; i < 16 * PM_ACC_STRIDE; i += 32) { w->acc[i] = 0; w->cov[i] = 0; w->dmin[i] = 1e9f; }
    const uint32_t n_complex = A.counters->n_complex, n_heavy = A.counters->n_heavy;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.report->n_complex = n_complex;
        A.report->n_overflow = A.counters->n_overflow;
        A.report->frame = A.stamp;
        A.counters_next->n_complex = 0;
        A.counters_next->n_overflow = 0;
        A.counters_next->n_heavy = 0;
    }
    __syncthreads();
    const uint32_t batches_per_row = (A.n_tx + 31u) / 32u;
    const uint32_t n_batches = batches_per_row * A.n_rows;
    bool complex_left = true, batches_left = true, heavy_left = true;
    const bool prefer_complex = warp < PM_FINE_COMPLEX_WARPS;
    while (complex_left || batches_left) {
        const bool take_complex = complex_left && (prefer_complex || !batches_left);
        if (take_complex) {
            // Tiles with records, one per queue access.  Pass 1 renders the heavy ones (more records than
            // inline slots: coincident outlines, deep stacks), which binning listed separately; pass 2
            // walks the full list and skips them.  Heavy tiles first keeps a 20-microsecond tile from
            // starting when everybody else is done; one tile per access spreads list neighbours (which
            // tend to be equally heavy) over as many warps as possible.
            // (one call site for the tile code: the kernel is instruction-cache bound)
            uint32_t pk = 0;
            u64 cw = 0;
            bool have = false;
            {
                const uint32_t n_list = heavy_left ? n_heavy : n_complex;
                uint32_t q = 0;
                if (lane == 0) q = atomicAdd(heavy_left ? &A.queue->heavy_next : &A.queue->complex_next, 1u);
                q = __shfl_sync(PM_FULL_MASK, q, 0);
                if (q >= n_list) {
                    if (heavy_left) heavy_left = false; else complex_left = false;
                    continue;
                }
                pk = A.complex_list[(heavy_left ? A.n_rows * A.n_tx : 0u) + q];
                cw = A.cnt[(pk >> 16) * A.n_tx + (pk & 0xffffu)];
                // pass 2 skips what pass 1 rendered (the stamp is this frame's: the tile is on the list)
                have = heavy_left || (uint32_t)cw <= PM_TILE_SLOTS;
            }
            if (have) fine_complex_tile<F32, EXACT>(A, pk, cw, A.occ[(pk >> 16) * A.n_tx + (pk & 0xffffu)], w, s_lut, lane);
        } else {
            uint32_t q = 0;
            if (lane == 0) q = atomicAdd(&A.queue->batch_next, 1u);
            q = __shfl_sync(PM_FULL_MASK, q, 0);
            if (q >= n_batches) { batches_left = false; continue; }
            fine_solid_batch<F32>(A, q, batches_per_row, lane);
        }
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
                    uint32_t n_tx, unsigned long long *plan_a, unsigned long long *plan_b, PmItemInfo *item_info, uint2 *row_info,
                    uint32_t row_info_cap, PmPlanResult *result, cudaStream_t s) {
    k_plan<<<1, 1024, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, plan_b, item_info, row_info, row_info_cap, result);
}

void pm_launch_plan_pieces(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                           const unsigned long long *plan_a, uint32_t n_segments, PmSegInfo *seg_info, uint2 *piece_info,
                           uint32_t piece_cap, PmPlanResult *result, cudaStream_t s) {
    k_plan_pieces<<<1, 1024, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, plan_a, n_segments, seg_info, piece_info, piece_cap, result);
}

void pm_launch_frame(const PmFrameArgs &a, int sm_count, cudaEvent_t mid, cudaStream_t s) {
    uint32_t grid_seg = (a.n_pieces + 255u) / 256u;
    if (grid_seg == 0) grid_seg = 1;  // still clears the fill kernel's queues
    k_seg<<<grid_seg, 256, 0, s>>>(a);
    if (a.n_row_units) k_row<<<(a.n_row_units + PM_ROW_WARPS - 1) / PM_ROW_WARPS, PM_ROW_WARPS * 32, 0, s>>>(a);
    if (mid) cudaEventRecord(mid, s);
    // persistent fill kernel: enough CTAs to fill every SM, work pulled from two queues
    int grid = sm_count * 4;
    const bool exact = (a.flags & PM_FLAG_EXACT_SRGB) != 0;
    if (a.fb32) {  // debug render with the fp32 parity buffer
        if (exact) k_fine<true, true><<<grid, PM_FINE_WARPS * 32, 0, s>>>(a);
        else       k_fine<true, false><<<grid, PM_FINE_WARPS * 32, 0, s>>>(a);
    } else {
        if (exact) k_fine<false, true><<<grid, PM_FINE_WARPS * 32, 0, s>>>(a);
        else       k_fine<false, false><<<grid, PM_FINE_WARPS * 32, 0, s>>>(a);
    }
}
