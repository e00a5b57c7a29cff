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

// ---------------------------------------------------------------------------------------------
// k_plan: one work unit per (item, tile row of its bbox inside the strip)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t item_row_count(const uint8_t *scene, uint32_t items_ix, uint32_t i, uint32_t tile_y0,
                                                   uint32_t tile_y1, uint32_t n_tx) {
    const pm_bbox bb = *reinterpret_cast<const pm_bbox *>(scene + PM_GROUP_HEADER_SIZE + (size_t)i * PM_BBOX_SIZE);
    uint32_t tag = ld_u32(scene + items_ix + (size_t)i * PM_ITEM_SIZE);
    if (tag < PM_ITEM_CIRCLE || tag > PM_ITEM_POLY) return 0;
    // `hit` (metal:214): bbox.z >= x0 && bbox.x < x0 + 16 && bbox.w >= y0 && bbox.y < y0 + 16
    // <=> tile column in [bbox.x >> 4, bbox.z >> 4] and tile row in [bbox.y >> 4, bbox.w >> 4]
    uint32_t t_lo = bb.x0 >> 4, t_hi = bb.x1 >> 4;
    if (t_lo >= n_tx || t_hi < t_lo) return 0;
    uint32_t r_lo = bb.y0 >> 4, r_hi = bb.y1 >> 4;
    if (r_lo < tile_y0) r_lo = tile_y0;
    if (r_hi >= tile_y1) r_hi = tile_y1 - 1;  // tile_y1 > tile_y0 >= 0
    if (r_hi < r_lo) return 0;
    return r_hi - r_lo + 1;
}

__global__ void __launch_bounds__(1024) k_plan(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0,
                                               uint32_t tile_y1, uint32_t n_tx, uint32_t *unit_base, PmPlanResult *result) {
    __shared__ uint32_t warp_excl[32];
    __shared__ uint32_t block_total;
    __shared__ uint32_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_items; base += blockDim.x) {
        uint32_t i = base + tid;
        uint32_t cnt = i < n_items ? item_row_count(scene, items_ix, i, tile_y0, tile_y1, n_tx) : 0;
        uint32_t incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(PM_FULL_MASK, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) warp_excl[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t ws = warp_excl[lane];
            uint32_t wi = ws;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(PM_FULL_MASK, wi, o);
                if (lane >= (uint32_t)o) wi += v;
            }
            warp_excl[lane] = wi - ws;
            if (lane == 31) block_total = wi;
        }
        __syncthreads();
        uint32_t excl = carry_s + warp_excl[warp] + incl - cnt;
        if (i < n_items) unit_base[i] = excl;
        if (excl + cnt < excl) result->error = 1;  // more than 2^32 work units
        __syncthreads();
        if (tid == 0) carry_s += block_total;
        __syncthreads();
    }
    if (tid == 0) {
        unit_base[n_items] = carry_s;
        result->n_units = carry_s;
    }
}

// ---------------------------------------------------------------------------------------------
// k_bin
// ---------------------------------------------------------------------------------------------
typedef unsigned long long u64;

// Claims the next record slot of a tile for this frame: returns its position (0 for the first).
// cnt word = stamp << 32 | count; a word with another stamp is a leftover of an earlier frame.
__device__ __forceinline__ uint32_t tile_claim_slot(u64 *word, uint32_t stamp) {
    u64 old = atomicAdd(word, 1ull);
    if ((uint32_t)(old >> 32) == stamp) return (uint32_t)old;
    u64 v = old + 1ull;  // we bumped a stale word: race to (re)initialise it
    for (;;) {
        if ((uint32_t)(v >> 32) == stamp) {  // somebody else initialised it (our bump went with the stale word)
            old = atomicAdd(word, 1ull);
            return (uint32_t)old;
        }
        u64 prev = atomicCAS(word, v, ((u64)stamp << 32) | 1ull);
        if (prev == v) return 0;
        v = prev;
    }
}

struct BinSink {
    const PmFrameArgs &A;
    uint32_t *sm;        // per-warp: word j <-> tile t_lo + j; bit 0 = "has a command", bits 1.. = 2 * backdrop delta
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
            uint32_t o = atomicAdd(&A.counters->n_overflow, 1u);
            if (o >= A.overflow_cap) return;  // the host sees n_overflow > overflow_cap, grows the pool and re-renders
            idx = A.n_rows * A.n_tx * PM_TILE_SLOTS + o;
            u64 prev = atomicExch(&A.ovf[tile], ((u64)A.stamp << 32) | (u64)(idx + 1u));
            if ((uint32_t)(prev >> 32) == A.stamp) r.next = (uint32_t)prev;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(&A.pool[idx]);
        const uint4 *src = reinterpret_cast<const uint4 *>(&r);
        dst[0] = src[0];
        dst[1] = src[1];
        if (pos == 0) {  // first record of the tile this frame: queue it for the fill kernel
            cg::coalesced_group g = cg::coalesced_threads();
            uint32_t base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(&A.counters->n_complex, g.size());
            base = g.shfl(base, 0);
            A.complex_list[base + g.thread_rank()] = tile;
        }
    }
    __device__ __forceinline__ void fill(uint32_t t, uint32_t seg, const PmFillEmit &e, const PmSeg &g) {
        append(t, pm_rec_fill(item, seg, t, e, g));
        atomicOr(&sm[t - t_lo], 1u);
    }
    __device__ __forceinline__ void backdrop(uint32_t ta, uint32_t tb, int delta) {
        atomicAdd(&sm[ta - t_lo], (uint32_t)(2 * delta));
        atomicAdd(&sm[tb + 1 - t_lo], (uint32_t)(-2 * delta));
    }
    __device__ __forceinline__ void line(uint32_t t, uint32_t seg, const PmSeg &g) {
        append(t, pm_rec_line(item, seg, g));
        atomicOr(&sm[t - t_lo], 1u);
    }
    __device__ __forceinline__ void trailer(uint32_t t, uint32_t kind, uint32_t seg, uint32_t w0, uint32_t w1) {
        append(t, pm_rec_words(item, kind, seg, w0, w1));
    }
};

#define PM_BIN_UNROLL 4   // point loads in flight per lane while scanning an item's segments
#define PM_BIN_PEND (32 * (PM_BIN_UNROLL + 1))   // pending (row-overlapping) segment indices per warp

// Scans the n_seg segments of an item for the ones whose y range reaches this unit's tile row
// (`overlap(sy, ey)`), compacts their indices and hands them to `process(k)` 32 at a time, so that
// the expensive exact tests run with all lanes busy.  Order does not matter: every effect of a
// segment is an atomic add / or / append.
template <class Overlap, class Process>
__device__ __forceinline__ void scan_segments(const uint8_t *pts, uint32_t n_seg, uint32_t n_points, uint32_t *pend, uint32_t lane,
                                              Overlap overlap, Process process) {
    uint32_t n_pend = 0;
    for (uint32_t k0 = 0; k0 < n_seg; k0 += 32 * PM_BIN_UNROLL) {
        float sy[PM_BIN_UNROLL], ey[PM_BIN_UNROLL];
        #pragma unroll
        for (int u = 0; u < PM_BIN_UNROLL; u++) {
            const uint32_t k = k0 + 32u * u + lane;
            sy[u] = ey[u] = 0.0f;
            if (k < n_seg) {
                sy[u] = ld_f32(pts + 8 * (size_t)k + 4);
                ey[u] = ld_f32(pts + 8 * (size_t)(k + 1 == n_points ? 0 : k + 1) + 4);
            }
        }
        #pragma unroll
        for (int u = 0; u < PM_BIN_UNROLL; u++) {
            const uint32_t k = k0 + 32u * u + lane;
            const bool ov = k < n_seg && overlap(sy[u], ey[u]);
            const uint32_t mask = __ballot_sync(PM_FULL_MASK, ov);
            if (ov) pend[n_pend + __popc(mask & ((1u << lane) - 1u))] = k;
            n_pend += __popc(mask);
        }
        __syncwarp();
        const bool last = k0 + 32 * PM_BIN_UNROLL >= n_seg;
        while (n_pend >= 32 || (last && n_pend > 0)) {  // the one call site of process()
            const uint32_t take = n_pend < 32 ? n_pend : 32;
            if (lane < take) process(pend[n_pend - 1 - lane]);
            n_pend -= take;
        }
        __syncwarp();
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_bin(const PmFrameArgs A) {
    extern __shared__ uint32_t smem_u32[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.queue->complex_next = 0;
        A.queue->batch_next = 0;
    }
    const uint32_t unit = blockIdx.x * WARPS + warp;
    if (unit >= A.n_units) return;

    // item = largest i with unit_base[i] <= unit
    uint32_t lo = 0, hi = A.n_items;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (A.unit_base[mid] <= unit) lo = mid; else hi = mid;
    }
    const uint32_t item = lo;
    const uint8_t *it = A.scene + A.items_ix + (size_t)item * PM_ITEM_SIZE;
    const pm_bbox bb = *reinterpret_cast<const pm_bbox *>(A.scene + PM_GROUP_HEADER_SIZE + (size_t)item * PM_BBOX_SIZE);
    const uint32_t tag = ld_u32(it);
    uint32_t r_lo = bb.y0 >> 4;
    if (r_lo < A.tile_y0) r_lo = A.tile_y0;
    const uint32_t row = r_lo + (unit - A.unit_base[item]);
    const uint32_t t_lo = bb.x0 >> 4;
    uint32_t t_hi = bb.x1 >> 4;
    if (t_hi > A.n_tx - 1) t_hi = A.n_tx - 1;
    const uint32_t span = t_hi - t_lo + 1;
    const float y0 = (float)(row * PM_TILE_H);

    uint32_t *sm = smem_u32 + (size_t)warp * (A.n_tx + 1 + PM_BIN_PEND);
    uint32_t *pend = sm + A.n_tx + 1;
    for (uint32_t j = lane; j <= span; j += 32) sm[j] = 0;
    __syncwarp();

    BinSink sink{A, sm, t_lo, (row - A.tile_y0) * A.n_tx, item};

    if (tag == PM_ITEM_FILL) {
        const uint32_t rgba = ld_u32(it + PM_FILL_RGBA);
        const uint32_t n_points = ld_u32(it + PM_FILL_NPOINTS);
        const uint8_t *pts = A.scene + ld_u32(it + PM_FILL_POINTS_IX);
        const uint32_t n_tx = A.n_tx;
        scan_segments(pts, n_points, n_points, pend, lane,
            [=](float sy, float ey) { return fmaxf(sy, ey) >= y0 && fminf(sy, ey) < y0 + 16.0f; },  // pm_fill_row_overlap
            [&](uint32_t k) {
                float2 s = ld_f2(pts + 8 * (size_t)k);
                float2 e = ld_f2(pts + 8 * (size_t)(k + 1 == n_points ? 0 : k + 1));  // closing segment, metal:262
                PmSeg g = pm_seg(s.x, s.y, e.x, e.y);
                pm_fill_segment_row(sink, g, y0, t_lo, t_hi, n_tx, k);
            });
        // per-tile epilogue (metal:359-363): DrawFill / Solid / nothing
        int carry = 0;
        for (uint32_t base = 0; base < span; base += 32) {
            uint32_t j = base + lane;
            uint32_t v = j < span ? sm[j] : 0u;
            int d = (int)v >> 1;
            int incl = d;
            for (int o = 1; o < 32; o <<= 1) {
                int u = __shfl_up_sync(PM_FULL_MASK, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            int backdrop = carry + incl;
            carry = __shfl_sync(PM_FULL_MASK, backdrop, 31);
            if (j < span) {
                uint32_t t = t_lo + j;
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
        }
    } else if (tag == PM_ITEM_POLY) {
        const uint32_t rgba = ld_u32(it + PM_POLY_RGBA);
        const float width = ld_f32(it + PM_POLY_WIDTH);
        const uint32_t n_points = ld_u32(it + PM_POLY_NPOINTS);
        const uint32_t n_seg = n_points - 1;  // open polyline, metal:369
        const uint8_t *pts = A.scene + ld_u32(it + PM_POLY_POINTS_IX);
        const float hw = 0.5f * width + 0.5f;
        const bool fix = (A.flags & PM_FLAG_FIX_POLY_PRECULL) != 0;
        scan_segments(pts, n_seg, n_points + 1, pend, lane,
            [=](float sy, float ey) { return fmaxf(sy, ey) > y0 - hw && fminf(sy, ey) < y0 + 16.0f + hw; },
            [&](uint32_t k) {
                float2 s = ld_f2(pts + 8 * (size_t)k);
                float2 e = ld_f2(pts + 8 * (size_t)(k + 1));
                PmSeg g = pm_seg(s.x, s.y, e.x, e.y);
                pm_poly_segment_row(sink, g, y0, hw, t_lo, t_hi, k, fix);
            });
        for (uint32_t j = lane; j < span; j += 32)
            if (sm[j] & 1u) sink.trailer(t_lo + j, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * width), rgba);  // metal:441-443
    } else if (tag == PM_ITEM_LINE) {  // metal:223-247
        const uint32_t rgba = ld_u32(it + PM_LINE_RGBA);
        const float width = ld_f32(it + PM_LINE_WIDTH);
        const float2 s = ld_f2(it + PM_LINE_START), e = ld_f2(it + PM_LINE_END);
        const PmSeg g = pm_seg(s.x, s.y, e.x, e.y);
        const float hw = 0.5f * width + 0.5f;
        for (uint32_t j = lane; j < span; j += 32) {
            uint32_t t = t_lo + j;
            float x0 = (float)(t * PM_TILE_W);
            if (pm_stroke_cross(g, x0, x0 + 16.0f, y0, y0 + 16.0f, hw)) {
                sink.line(t, 0, g);
                sink.trailer(t, PM_REC_STROKE, PM_REC_SEG_MAX, pm_f2u(0.5f * width), rgba);
            }
        }
    } else if (tag == PM_ITEM_CIRCLE) {  // metal:218-222
        uint32_t b_lo = (uint32_t)bb.x0 | ((uint32_t)bb.y0 << 16), b_hi = (uint32_t)bb.x1 | ((uint32_t)bb.y1 << 16);
        for (uint32_t j = lane; j < span; j += 32) sink.trailer(t_lo + j, PM_REC_CIRCLE, 0, b_lo, b_hi);
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
    float p = EXACT ? powf(v, 1.0f / 2.4f) : exp2f(__log2f(v) * (1.0f / 2.4f));
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
        else pm_fill_rows(kind, r.p[1], r.p[3], r.edge_y, tile_y0, &ra, &rb);
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
        const uint32_t o_kind = __shfl_sync(PM_FULL_MASK, kind, lo);
        float p[4];
        p[0] = __shfl_sync(PM_FULL_MASK, r.p[0], lo);
        p[1] = __shfl_sync(PM_FULL_MASK, r.p[1], lo);
        p[2] = __shfl_sync(PM_FULL_MASK, r.p[2], lo);
        p[3] = __shfl_sync(PM_FULL_MASK, r.p[3], lo);
        const float o_edge = __shfl_sync(PM_FULL_MASK, r.edge_y, lo);
        if (q < total) {
            const int row = o_ra + (q - o_excl);
            if (stroke) pm_line_pair(acc, p, reach, row, tile_x0, tile_y0);
            else pm_fill_pair(acc, o_kind, p, o_edge, row, tile_x0, tile_y0);
        }
    }
}

// One tile that owns records.  All 32 lanes execute this together.  Records are handled in chunks
// of 32, one per lane; the first chunk (all of them, for nearly every tile) stays in registers.
// Blend/store layout: lane l owns pixel row (l >> 1), pixels 8*(l & 1) .. +7.
template <bool F32, bool EXACT>
__device__ void fine_complex_tile(const PmFrameArgs &A, uint32_t tile, FineWarpSmem *w, const float *lut, uint32_t lane) {
    const u64 cw = A.cnt[tile], ow = A.occ[tile];
    const uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;

    // index the records: inline slots first, then the overflow chain.  n_cached counts what was
    // actually found (a frame whose overflow pool ran out has fewer links than cnt says; the host
    // re-renders such a frame, it only must not fault).
    const uint32_t n_inline = n < PM_TILE_SLOTS ? n : PM_TILE_SLOTS;
    uint32_t n_cached = n_inline;
    uint32_t tail = 0;  // 1 + pool index of the first record that did not fit the shared-memory index
    if (n > PM_TILE_SLOTS) {
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
    const uint32_t n_chunks = (n_cached + 31u) >> 5;
    // chunk 0 lives in registers for the whole tile
    PmRecord r0;
    r0.item = 0xffffffffu; r0.key = 0; r0.p[0] = r0.p[1] = r0.p[2] = r0.p[3] = 0.0f; r0.edge_y = 0.0f; r0.next = 0;
    if (lane < n_cached) r0 = load_record(A.pool, n > PM_TILE_SLOTS ? w->idx[lane] : tile * PM_TILE_SLOTS + lane);
    if (r0.item < occ_item1) r0.item = 0xffffffffu;  // below the topmost opaque cover: rewound away (metal:132-135)

    bool has_draw = r0.item != 0xffffffffu && (r0.key & 15u) != PM_REC_SOLID;
    for (uint32_t c = 1; c < n_chunks; c++) {
        const uint32_t i = c * 32u + lane;
        if (i < n_cached) {
            const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[w->idx[i]]);
            if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
        }
    }
    for (uint32_t cur = tail; cur != 0; cur = A.pool[cur - 1u].next) {
        const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[cur - 1u]);
        if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
    }
    has_draw = __any_sync(PM_FULL_MASK, has_draw);

    const uint32_t trow = tile / A.n_tx, tx = tile - trow * A.n_tx;
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
        reinterpret_cast<uint4 *>(dst)[0] = v;
        reinterpret_cast<uint4 *>(dst)[1] = v;
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
    if (occ_item1) {  // the rewound list starts with the cover's Cmd_Solid (metal:136-142, :546-551)
        float fg[4];
        unpack_fg(lut, occ_rgba, fg);
        #pragma unroll
        for (int j = 0; j < 8; j++)
            #pragma unroll
            for (int k = 0; k < 3; k++) rgb[j][k] = pm_mix(rgb[j][k], fg[k], fg[3]);
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
        for (uint32_t c = 1; c < n_chunks; c++) {
            const uint32_t i = c * 32u + lane;
            if (i < n_cached) {
                const uint32_t it = A.pool[w->idx[i]].item;
                if (it >= occ_item1 && (first || it > last_item) && it < cur_item) cur_item = it;
            }
        }
        for (uint32_t cur = tail; cur != 0; cur = A.pool[cur - 1u].next) {
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
        for (uint32_t c = 1; c < n_chunks; c++) {
            const uint32_t i = c * 32u + lane;
            if (i < n_cached) {
                const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[w->idx[i]]);
                if (a.x == cur_item && (a.y & 15u) >= PM_REC_CIRCLE) { t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w; }
            }
        }
        for (uint32_t cur = tail; cur != 0; cur = A.pool[cur - 1u].next) {
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
                if (c > 0) {
                    const uint32_t i = c * 32u + lane;
                    rc.item = 0xffffffffu;
                    if (i < n_cached) rc = load_record(A.pool, w->idx[i]);
                }
                const bool mine = rc.item == cur_item && (rc.key & 15u) <= PM_REC_LINE;
                if (__any_sync(PM_FULL_MASK, mine)) fine_pairs(acc, mine, rc, stroke, reach, tile_x0, tile_y0, lane);
            }
            for (uint32_t cur = tail; cur != 0;) {  // records beyond the shared-memory index, one at a time
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
    reinterpret_cast<uint4 *>(dst)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    reinterpret_cast<uint4 *>(dst)[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
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
        #pragma unroll
        for (int y = 0; y < PM_TILE_H; y++) *reinterpret_cast<uint4 *>(dst + (size_t)y * A.pitch) = v;
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
    const uint32_t n_complex = A.counters->n_complex;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.report->n_complex = n_complex;
        A.report->n_overflow = A.counters->n_overflow;
        A.report->frame = A.stamp;
        A.counters_next->n_complex = 0;
        A.counters_next->n_overflow = 0;
    }
    __syncthreads();
    const uint32_t batches_per_row = (A.n_tx + 31u) / 32u;
    const uint32_t n_batches = batches_per_row * A.n_rows;
    bool complex_left = true, batches_left = true;
    const bool prefer_complex = warp < PM_FINE_COMPLEX_WARPS;
    while (complex_left || batches_left) {
        const bool take_complex = complex_left && (prefer_complex || !batches_left);
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(take_complex ? &A.queue->complex_next : &A.queue->batch_next, 1u);
        q = __shfl_sync(PM_FULL_MASK, q, 0);
        if (take_complex) {
            if (q >= n_complex) { complex_left = false; continue; }
            fine_complex_tile<F32, EXACT>(A, A.complex_list[q], w, s_lut, lane);
        } else {
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
                    uint32_t n_tx, uint32_t *unit_base, PmPlanResult *result, cudaStream_t s) {
    k_plan<<<1, 1024, 0, s>>>(scene, n_items, items_ix, tile_y0, tile_y1, n_tx, unit_base, result);
}

static size_t bin_smem_bytes(uint32_t n_tx, int *warps_per_cta) {
    size_t per_warp = (size_t)(n_tx + 1 + PM_BIN_PEND) * sizeof(uint32_t);
    int warps = 8;
    while (warps > 1 && per_warp * warps > 160 * 1024) warps >>= 1;
    *warps_per_cta = warps;
    return per_warp * warps;
}

template <int WARPS>
static void launch_bin(const PmFrameArgs &a, size_t smem, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_bin<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = true;
    }
    uint32_t grid = (a.n_units + WARPS - 1) / WARPS;
    if (grid == 0) grid = 1;  // still clears the fill kernel's queues
    k_bin<WARPS><<<grid, WARPS * 32, smem, s>>>(a);
}

void pm_launch_frame(const PmFrameArgs &a, int sm_count, cudaEvent_t mid, cudaStream_t s) {
    int warps = 8;
    size_t smem = bin_smem_bytes(a.n_tx, &warps);
    switch (warps) {
        case 8: launch_bin<8>(a, smem, s); break;
        case 4: launch_bin<4>(a, smem, s); break;
        case 2: launch_bin<2>(a, smem, s); break;
        default: launch_bin<1>(a, smem, s); break;
    }
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
