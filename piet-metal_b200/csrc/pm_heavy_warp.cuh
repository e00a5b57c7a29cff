// Heavy tiles (more records than the 16 inline slots): helpers of k_heavy (pm_heavy.cu; sm_100a, -fmad=false).
//
// pm_heavy_tile_warp: one warp renders one tile of up to PM_HEAVY_WARP_CAP records -- k_fine's algorithm with the records
// streamed from L1/L2 in chunks of 32 instead of held one per lane.  Used when heavy tiles are plentiful and what
// counts is throughput; when they are few, k_heavy spends a whole CTA on each, because then they are the frame's
// critical path.  pm_heavy_warp_mode() is that decision.
// (Measured and rejected: letting k_fine's own warps take the heavy tiles in warp mode -- the out-of-line call costs
// k_fine 400 bytes of stack and 15 % of its speed on every tile, heavy or not.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_cover.cuh"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"

#define PM_HEAVY_WARP_CAP 128u    // warp mode: tiles up to this many records (4 chunks of 32); needs 2 overflow blocks at most
#define PM_HEAVY_WARP_DIR 4u      // overflow blocks a warp indexes

// Warp mode when there are more heavy tiles than a CTA each could take in about the time a warp needs for one -- or when
// the frame has so many tiles with records that the fill/blend kernel runs longer than a warp needs for a heavy tile anyway.
#ifndef PM_HEAVY_WARP_MODE_COMPLEX
#define PM_HEAVY_WARP_MODE_COMPLEX 32768u   // (the 8192^2 tiger, 49 k tiles with records, 619 of them heavy: frame 158 -> 152 us)
#endif
__device__ __forceinline__ bool pm_heavy_warp_mode(uint32_t n_heavy, uint32_t n_ctas, uint32_t n_complex) {
    return n_heavy > 6u * n_ctas || n_complex > PM_HEAVY_WARP_MODE_COMPLEX;
}

__device__ __forceinline__ PmRecord pm_load_record(const PmRecord *pool, uint32_t idx) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&pool[idx]);
    const uint4 a = src[0], b = src[1];
    PmRecord r;
    r.item = a.x; r.key = a.y; r.p[0] = pm_u2f(a.z); r.p[1] = pm_u2f(a.w);
    r.p[2] = pm_u2f(b.x); r.p[3] = pm_u2f(b.y); r.edge_y = pm_u2f(b.z); r.next = b.w;
    return r;
}

// pool index of the tile's record at position pos (pos < the number of indexed records)
__device__ __forceinline__ uint32_t pm_heavy_index(const uint32_t *dir, size_t tile, uint32_t pos) {
    if (pos < PM_TILE_SLOTS) return (uint32_t)tile * PM_TILE_SLOTS + pos;
    uint32_t j, off;
    pm_ovf_locate(pos - PM_TILE_SLOTS, &j, &off);
    return dir[j] + off;
}

// Walks the tile's chain of overflow blocks (one thread): dir[j] = 1 + pool index of block j's header for as many
// blocks as n records need (at most dir_cap).  Returns the number of records that can be reached -- fewer than n if
// the pool ran out (the host renders such a frame again with a larger pool; this pass only must not fault).
__device__ __forceinline__ uint32_t pm_heavy_walk(const PmFrameArgs &A, unsigned long long vw, uint32_t n, uint32_t *dir, uint32_t dir_cap) {
    uint32_t nb = 0, reach = PM_TILE_SLOTS;
    uint32_t link = (uint32_t)(vw >> 32) == A.stamp ? (uint32_t)vw : 0u;
    while (reach < n && link != 0 && link != PM_EXT_FAILED && nb < dir_cap) {
        dir[nb] = link;
        reach += pm_blk_size(nb);
        nb++;
        if (reach < n) link = A.pool[link - 1u].next;
    }
    return reach < n ? reach : n;
}


// Resolves one layer for this lane's 8 pixels (lane l: pixel row l / 2, pixels 8 (l & 1) .. +7) from the coverage
// arrays acc / cov, which it clears: al[0..7] = the layer's alpha (metal:536-537 nonzero rule, :58-60 renderDf,
// :481-490 circle), not yet multiplied by the paint's alpha.
__device__ __forceinline__ void pm_heavy_resolve8(int *acc, int *cov, uint32_t kind, uint32_t w0, uint32_t w1, float tile_x0, float tile_y0, uint32_t lane, float al[8]) {
    const uint32_t prow = lane >> 1, half = lane & 1u;
    const int off0 = pm_cov_swz((int)prow, (int)half * 8), off1 = pm_cov_swz((int)prow, (int)half * 8 + 4);
    if (pm_rec_is_drawfill(kind)) {
        const bool eo = kind == PM_REC_DRAWFILL_EO;
        int4 *pa0 = reinterpret_cast<int4 *>(&acc[off0]), *pa1 = reinterpret_cast<int4 *>(&acc[off1]);
        int4 *pc0 = reinterpret_cast<int4 *>(&cov[off0]), *pc1 = reinterpret_cast<int4 *>(&cov[off1]);
        const int4 a0 = *pa0, a1 = *pa1, c0 = *pc0, c1 = *pc1;
        *pa0 = make_int4(0, 0, 0, 0); *pa1 = make_int4(0, 0, 0, 0);
        *pc0 = make_int4(0, 0, 0, 0); *pc1 = make_int4(0, 0, 0, 0);
        const int sum = ((c0.x + c0.y) + (c0.z + c0.w)) + ((c1.x + c1.y) + (c1.z + c1.w));
        const int other = __shfl_xor_sync(PM_FULL_MASK, sum, 1);  // covers of the left half of the row carry into the right half
        int run = half ? other : 0;
        const int bd = (int)w0;
        run += c0.x; al[0] = pm_resolve_fill(a0.x + run, bd, eo);
        run += c0.y; al[1] = pm_resolve_fill(a0.y + run, bd, eo);
        run += c0.z; al[2] = pm_resolve_fill(a0.z + run, bd, eo);
        run += c0.w; al[3] = pm_resolve_fill(a0.w + run, bd, eo);
        run += c1.x; al[4] = pm_resolve_fill(a1.x + run, bd, eo);
        run += c1.y; al[5] = pm_resolve_fill(a1.y + run, bd, eo);
        run += c1.z; al[6] = pm_resolve_fill(a1.z + run, bd, eo);
        run += c1.w; al[7] = pm_resolve_fill(a1.w + run, bd, eo);
    } else if (kind == PM_REC_STROKE) {
        int4 *pa0 = reinterpret_cast<int4 *>(&acc[off0]), *pa1 = reinterpret_cast<int4 *>(&acc[off1]);
        const int4 a0 = *pa0, a1 = *pa1;
        *pa0 = make_int4(0, 0, 0, 0); *pa1 = make_int4(0, 0, 0, 0);
        const float lim = pm_u2f(w0) + 0.5f;
        const int a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        #pragma unroll
        for (int j = 0; j < 8; j++) al[j] = a[j] ? pm_saturate(lim - __uint_as_float(~(uint32_t)a[j])) : 0.0f;
    } else if (kind == PM_REC_CIRCLE) {
        #pragma unroll 1
        for (int j = 0; j < 8; j++) al[j] = pm_px_circle_alpha(w0, w1, tile_x0 + (float)(half * 8u + (uint32_t)j), tile_y0 + (float)prow);
    } else {  // PM_REC_SOLID: a translucent full cover
        #pragma unroll
        for (int j = 0; j < 8; j++) al[j] = 1.0f;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// warp mode: one warp, one tile of up to PM_HEAVY_WARP_CAP records, streamed 32 at a time
// ---------------------------------------------------------------------------------------------------------------
template <bool F32, bool EXACT>
__device__ __noinline__ void pm_heavy_tile_warp(const PmFrameArgs &A, int *acc, int *cov, float4 (*rgb)[2][32], uint32_t *dir, uint32_t entry, uint32_t lane) {
    const uint32_t trow = entry >> 16, tx = entry & 0xffffu;
    const size_t tile = (size_t)trow * A.n_tx + tx;
    const unsigned long long cw = A.cnt[tile], ow = A.occ[tile];
    uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    if (n > PM_HEAVY_WARP_CAP) return;  // drawn CTA-wise in the second pass
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;
    if (lane == 0 && n > PM_TILE_SLOTS) n = pm_heavy_walk(A, A.ovf[tile], n, dir, PM_HEAVY_WARP_DIR);
    n = __shfl_sync(PM_FULL_MASK, n, 0);
    __syncwarp();
    const uint32_t n_chunks = (n + 31u) / 32u;

    const uint32_t prow = lane >> 1, half = lane & 1u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + half * 8u) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) + (tx * PM_TILE_W + half * 8u);
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);

    // this lane's records: positions lane, lane + 32, ... ; their pool indices and (item, kind) are re-derived on every pass
    bool has_draw = false;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t p = c * 32u + lane;
        if (p < n) {
            const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[pm_heavy_index(dir, tile, p)]);
            if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
        }
    }
    has_draw = __any_sync(PM_FULL_MASK, has_draw);
    if (!has_draw) {
        uint32_t c = 0xffffffffu;
        if (occ_item1) c = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));
        const uint4 v = make_uint4(c, c, c, c);
        reinterpret_cast<uint4 *>(dst)[0] = v;
        reinterpret_cast<uint4 *>(dst)[1] = v;
        if (F32) {
            const float4 f = make_float4((float)(c & 0xff) / 255.0f, (float)((c >> 8) & 0xff) / 255.0f, (float)((c >> 16) & 0xff) / 255.0f, (float)(c >> 24) / 255.0f);
            for (int j = 0; j < 8; j++) dst32[j] = f;
        }
        return;
    }
    {
        float4 base = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        if (occ_item1) base = __ldg(&A.item_paint[occ_item1 - 1u]);
        #pragma unroll
        for (int g = 0; g < 2; g++) {
            rgb[0][g][lane] = make_float4(base.x, base.x, base.x, base.x);
            rgb[1][g][lane] = make_float4(base.y, base.y, base.y, base.y);
            rgb[2][g][lane] = make_float4(base.z, base.z, base.z, base.z);
        }
    }
    PmCoverAcc cacc{acc, cov};
    uint32_t lo_item = occ_item1;
    for (;;) {
        // the next item in painter's order, and its closing record
        uint32_t cand = 0xffffffffu;
        for (uint32_t c = 0; c < n_chunks; c++) {
            const uint32_t p = c * 32u + lane;
            if (p < n) {
                const uint32_t it = A.pool[pm_heavy_index(dir, tile, p)].item;
                if (it >= lo_item && it < cand) cand = it;
            }
        }
        const uint32_t cur = __reduce_min_sync(PM_FULL_MASK, cand);
        if (cur == 0xffffffffu) break;
        lo_item = cur + 1u;
        uint32_t kind = 0, w0 = 0, w1 = 0;
        for (uint32_t c = 0; c < n_chunks; c++) {
            const uint32_t p = c * 32u + lane;
            bool hit = false;
            uint4 a = make_uint4(0, 0, 0, 0);
            if (p < n) {
                a = *reinterpret_cast<const uint4 *>(&A.pool[pm_heavy_index(dir, tile, p)]);
                hit = a.x == cur && (a.y & 15u) >= PM_REC_CIRCLE;
            }
            const uint32_t m = __ballot_sync(PM_FULL_MASK, hit);
            if (m) {
                const int src = __ffs(m) - 1;
                kind = __shfl_sync(PM_FULL_MASK, a.y, src) & 15u;
                w0 = __shfl_sync(PM_FULL_MASK, a.z, src);
                w1 = __shfl_sync(PM_FULL_MASK, a.w, src);
            }
        }
        if (kind == 0) continue;  // cannot happen for a well-formed list
        const float4 paint = kind != PM_REC_CIRCLE ? __ldg(&A.item_paint[cur]) : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
        const bool stroke = kind == PM_REC_STROKE;
        if (stroke || pm_rec_is_drawfill(kind)) {
            const float reach = pm_u2f(w0) + 0.5f;
            for (uint32_t c = 0; c < n_chunks; c++) {
                const uint32_t p = c * 32u + lane;
                bool mine = false;
                PmRecord rc;
                rc.key = 0; rc.p[0] = rc.p[1] = rc.p[2] = rc.p[3] = 0.0f; rc.edge_y = 0.0f;
                if (p < n) {
                    const uint32_t idx = pm_heavy_index(dir, tile, p);
                    const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[idx]);
                    if (ik.x == cur && (ik.y & 15u) <= PM_REC_LINE) { mine = true; rc = pm_load_record(A.pool, idx); }
                }
                if (__any_sync(PM_FULL_MASK, mine))
                    pm_cover_records(cacc, mine, rc.key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, reach, tile_x0, tile_y0, lane);
            }
            __syncwarp();
        }
        float al[8];
        pm_heavy_resolve8(acc, cov, kind, w0, w1, tile_x0, tile_y0, lane, al);
        #pragma unroll
        for (int g = 0; g < 2; g++) {
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                const float fg = k == 0 ? paint.x : (k == 1 ? paint.y : paint.z);
                float4 v = rgb[k][g][lane];
                v.x = pm_mix_fma(v.x, fg, al[4 * g + 0] * paint.w);
                v.y = pm_mix_fma(v.y, fg, al[4 * g + 1] * paint.w);
                v.z = pm_mix_fma(v.z, fg, al[4 * g + 2] * paint.w);
                v.w = pm_mix_fma(v.w, fg, al[4 * g + 3] * paint.w);
                rgb[k][g][lane] = v;
            }
        }
        __syncwarp();
    }
    #pragma unroll 1
    for (int g = 0; g < 2; g++) {
        const float4 r = rgb[0][g][lane], gg = rgb[1][g][lane], bl = rgb[2][g][lane];
        const uint4 px = make_uint4(pm_encode_pixel<EXACT>(r.x, gg.x, bl.x), pm_encode_pixel<EXACT>(r.y, gg.y, bl.y),
                                    pm_encode_pixel<EXACT>(r.z, gg.z, bl.z), pm_encode_pixel<EXACT>(r.w, gg.w, bl.w));
        reinterpret_cast<uint4 *>(dst)[g] = px;
        if (F32) {
            dst32[4 * g + 0] = make_float4(pm_linear_to_srgb<EXACT>(r.x), pm_linear_to_srgb<EXACT>(gg.x), pm_linear_to_srgb<EXACT>(bl.x), 1.0f);
            dst32[4 * g + 1] = make_float4(pm_linear_to_srgb<EXACT>(r.y), pm_linear_to_srgb<EXACT>(gg.y), pm_linear_to_srgb<EXACT>(bl.y), 1.0f);
            dst32[4 * g + 2] = make_float4(pm_linear_to_srgb<EXACT>(r.z), pm_linear_to_srgb<EXACT>(gg.z), pm_linear_to_srgb<EXACT>(bl.z), 1.0f);
            dst32[4 * g + 3] = make_float4(pm_linear_to_srgb<EXACT>(r.w), pm_linear_to_srgb<EXACT>(gg.w), pm_linear_to_srgb<EXACT>(bl.w), 1.0f);
        }
    }
}

