// Coverage accumulation shared by the two fill/blend kernels (k_fine: one warp per tile, pm_fine.cu; k_heavy: one
// CTA per tile, pm_heavy.cu).  Device code for sm_100a, compiled with -fmad=false like the rest of the path.
//
// renderKernel's per-command arithmetic (TestApp/PietRender.metal:495-534) evaluated sparsely by one warp over a
// 16x16-pixel coverage array in shared memory:
//   acc[256]   Fill: sum of the "near" pixels' signed areas (8.24 fixed point); Stroke: max of ~bits(distance)
//   cov[256]   Fill: per-row cover deltas (every pixel to the right gets them: prefix-summed when the item is resolved)
// [pixel row][x] with the 4-pixel groups of a row XOR-swizzled by the row, so that the scattered atomics of the
// accumulation and the row-wise 128-bit accesses of the resolve both spread over the banks.  Integer sums: the result
// does not depend on the order in which lanes (or warps) arrive, which keeps a frame deterministic -- N multi-GPU
// strips reproduce the 1-GPU frame byte for byte.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_pixel_logic.h"

#define PM_FULL_MASK 0xffffffffu

__device__ __forceinline__ int pm_cov_swz(int row, int j) { return row * 16 + (j ^ (((row >> 1) & 3) << 2)); }

struct PmCoverAcc {
    int *acc;
    int *cov;
    __device__ __forceinline__ void near(int row, int j, int fx) { atomicAdd(&acc[pm_cov_swz(row, j)], fx); }
    __device__ __forceinline__ void cover(int row, int j, int fx) { atomicAdd(&cov[pm_cov_swz(row, j)], fx); }
    __device__ __forceinline__ void dist(int row, int j, float d) {  // d >= 0: unsigned order == float order
        atomicMax(reinterpret_cast<unsigned int *>(&acc[pm_cov_swz(row, j)]), ~__float_as_uint(d));
    }
};

// stroke() distance of metal:49-55 with the square root on the SFU (sqrt.approx: 1 ulp; the distances that matter
// are below reach + 1 pixels, so the difference to sqrtf is under 1e-6 of alpha)
__device__ __forceinline__ float pm_px_line_dist_fast(float sx, float sy, float ex, float ey, float px, float py) {
    const float lvx = ex - sx, lvy = ey - sy;
    const float dpx = px - sx, dpy = py - sy;
    const float t = pm_saturate(pm_div(lvx * dpx + lvy * dpy, lvx * lvx + lvy * lvy));
    const float qx = lvx * t - dpx, qy = lvy * t - dpy;
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(qx * qx + qy * qy));
    return r;
}

// Coverage of up to 32 records held one per lane (`mine`: this lane holds a FILL* / LINE record of the item being
// drawn; kind = its PM_REC_* kind; p0..p3 = start.xy, end.xy; edge_y of a FILL_EDGE_* record).  All 32 lanes call.
// Two levels of work distribution, because both the pixel rows a segment crosses and the pixels of a row that need
// arithmetic vary from 0 to 16:
//   level 1: the (record, pixel row) pairs are enumerated across the lanes; a lane computes the row-dependent part of
//            its pair (metal:510-516), adds the row's cover delta and finds the pixel span that needs per-pixel work
//            (fill: the pixels the segment passes through; stroke: the pixels within reach of it);
//   level 2: those (pair, pixel) units are enumerated across the lanes again, one pixel per lane (metal:517-527, :49-55).
// Owner lookup at both levels: exclusive prefix and payload packed into one word that is monotone in the lane,
// binary search with shuffles.
__device__ __forceinline__ void pm_cover_records(PmCoverAcc acc, bool mine, uint32_t kind, float r_p0, float r_p1, float r_p2, float r_p3,
                                                 float r_edge_y, bool stroke, float reach, float tile_x0, float tile_y0, uint32_t lane) {
    int ra = 1, rb = 0;
    if (mine) {
        if (stroke) pm_line_rows(r_p1, r_p3, reach, tile_y0, &ra, &rb);
        else pm_fill_rows(r_p1, r_p3, tile_y0, &ra, &rb);
    }
    const int cnt = rb >= ra ? rb - ra + 1 : 0;
    int incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(PM_FULL_MASK, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    const int key = ((incl - cnt) << 5) | ra;  // (pairs before this lane, first row)
    const int total = __shfl_sync(PM_FULL_MASK, incl, 31);
    #pragma unroll 1
    for (int q = (int)lane; q - (int)lane < total; q += 32) {
        // level 1: owner = last lane whose exclusive prefix is <= q
        const int qk = (q << 5) | 31;
        int lo = 0;
        #pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(PM_FULL_MASK, key, lo + step);
            if (v <= qk) lo += step;
        }
        const int o_key = __shfl_sync(PM_FULL_MASK, key, lo);
        float p[4];
        p[0] = __shfl_sync(PM_FULL_MASK, r_p0, lo);
        p[1] = __shfl_sync(PM_FULL_MASK, r_p1, lo);
        p[2] = __shfl_sync(PM_FULL_MASK, r_p2, lo);
        p[3] = __shfl_sync(PM_FULL_MASK, r_p3, lo);
        // this lane's pair: d0..d5 is what a pixel of it needs (stroke: the segment; fill: sx, ex and the row's window / t)
        int row = 0, j0 = 0, npx = 0;
        float d0 = p[0], d1 = p[1], d2 = p[2], d3 = p[3], d4 = 0.0f, d5 = 0.0f;
        if (q < total) {
            row = (o_key & 31) + (q - (o_key >> 5));
            if (stroke) {
                int ja, jb;
                pm_line_pair_span(p, reach, row, tile_x0, tile_y0, &ja, &jb);
                j0 = ja;
                npx = jb >= ja ? jb - ja + 1 : 0;
            } else {
                PmFillRow fr;
                int j_near, j_cover;
                if (pm_fill_pair_row(p, row, tile_x0, tile_y0, &fr, &j_near, &j_cover)) {
                    if (j_cover < 16) acc.cover(row, j_cover, pm_to_fx(fr.wx - fr.wy));
                    j0 = j_near;
                    npx = j_cover - j_near;
                    d1 = p[2]; d2 = fr.wx; d3 = fr.wy; d4 = fr.tx; d5 = fr.ty;
                }
            }
        }
        // level 2
        int incl2 = npx;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(PM_FULL_MASK, incl2, o);
            if (lane >= (uint32_t)o) incl2 += v;
        }
        const int key2 = ((incl2 - npx) << 9) | (row << 5) | j0;  // (pixels before this lane, row, first pixel)
        const int total2 = __shfl_sync(PM_FULL_MASK, incl2, 31);
        #pragma unroll 1
        for (int u = (int)lane; u - (int)lane < total2; u += 32) {
            const int uk = (u << 9) | 511;
            int lo2 = 0;
            #pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(PM_FULL_MASK, key2, lo2 + step);
                if (v <= uk) lo2 += step;
            }
            const int k2 = __shfl_sync(PM_FULL_MASK, key2, lo2);
            const float e0 = __shfl_sync(PM_FULL_MASK, d0, lo2);
            const float e1 = __shfl_sync(PM_FULL_MASK, d1, lo2);
            const float e2 = __shfl_sync(PM_FULL_MASK, d2, lo2);
            const float e3 = __shfl_sync(PM_FULL_MASK, d3, lo2);
            const int prow = (k2 >> 5) & 15;
            const int j = (k2 & 31) + (u - (k2 >> 9));
            if (stroke) {
                if (u < total2) acc.dist(prow, j, pm_px_line_dist_fast(e0, e1, e2, e3, tile_x0 + (float)j, tile_y0 + (float)prow));
            } else {
                PmFillRow fr;
                fr.wx = e2; fr.wy = e3;
                fr.tx = __shfl_sync(PM_FULL_MASK, d4, lo2);
                fr.ty = __shfl_sync(PM_FULL_MASK, d5, lo2);
                fr.active = true;
                if (u < total2) acc.near(prow, j, pm_fill_pair_px(e0, e1, tile_x0, j, fr));
            }
        }
    }
    // FillEdge commands (metal:530-534): one record at a time, lanes 0..15 take the 16 pixel rows
    if (!stroke) {
        for (uint32_t em = __ballot_sync(PM_FULL_MASK, mine && kind != PM_REC_FILL); em != 0; em &= em - 1) {
            const int src = __ffs(em) - 1;
            const uint32_t e_kind = __shfl_sync(PM_FULL_MASK, kind, src);
            const float e_y = __shfl_sync(PM_FULL_MASK, r_edge_y, src);
            if (lane < 16) pm_fill_edge_row(acc, e_kind, e_y, (int)lane, tile_y0);
        }
    }
}

// ---- pixel helpers ----

// linear -> sRGB (metal:563).  EXACT: powf and the canonical formula (debug renders, PM_FLAG_EXACT_SRGB)
template <bool EXACT>
__device__ __forceinline__ float pm_linear_to_srgb(float v) {
    if (v < 0.0031308f) return 12.92f * v;
    float p;
    if (EXACT) {
        p = powf(v, 1.0f / 2.4f);
    } else {  // ex2(lg2(v) / 2.4) on the SFU, a few 1e-7 from powf
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
        l *= 1.0f / 2.4f;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(l));
    }
    return 1.055f * p - 0.055f;
}
// One channel, linear -> sRGB byte.  Default path: the scale to 0..255 folded into the curve and a saturating convert
// (negative, NaN -> 0; > 1 -> 255), no branch.
template <bool EXACT>
__device__ __forceinline__ uint32_t pm_srgb_byte(float v) {
    if (EXACT) return pm_unorm8(pm_linear_to_srgb<true>(v));
    float l, p;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
    l *= 1.0f / 2.4f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(l));
    const float s = __fmaf_rn(p, 1.055f * 255.0f, -0.055f * 255.0f);
    const float lin = v * (12.92f * 255.0f);
    const float r = v < 0.0031308f ? lin : s;
    uint32_t b;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(b) : "f"(r));
    return b;
}
template <bool EXACT>
__device__ __forceinline__ uint32_t pm_encode_pixel(float r, float g, float b) {
    return pm_srgb_byte<EXACT>(r) | (pm_srgb_byte<EXACT>(g) << 8) | (pm_srgb_byte<EXACT>(b) << 16) | 0xff000000u;
}

// mix(x, y, a) with two FMAs, exact at a == 0 and a == 1 (metal:505, :543, :549: within an ulp or two of x + (y - x) * a)
__device__ __forceinline__ float pm_mix_fma(float x, float y, float a) { return __fmaf_rn(a, y, __fmaf_rn(-a, x, x)); }
