// k_fine: the fill/blend kernel of the render path (sm_100a).  Compiled with -fmad=false like the
// rest of the path (pm_tile_logic.h); the few places that want an FMA ask for one explicitly.
//
// renderKernel's arithmetic (TestApp/PietRender.metal:457-566) evaluated sparsely, one warp per
// tile that owns records, fused with the solid-tile composite (metal:16-44): tiles without records
// are written as 32-tile batches of full 512-byte rows.
//
// Per warp:
//   * the tile's header words and its 16 inline record slots are fetched with cp.async into shared
//     memory while the previous tile is being encoded and stored, so a tile starts without a
//     dependent chain of global loads (queue -> list -> cnt/occ -> records);
//   * coverage of one item is accumulated in shared memory in 8.24 fixed point by lanes that
//     enumerate (record, pixel row) pairs (Cmd_Fill / Cmd_FillEdge, metal:508-534) or as the
//     minimum distance to the item's segments (Cmd_Line, metal:495-498);
//   * the linear colour of the tile's 256 pixels lives in a lane-private slice of shared memory
//     (8 pixels per lane) instead of 24 registers (the kernel runs at 80 registers, 3 CTAs per SM);
//   * the sRGB encode and the 128-bit framebuffer stores happen once per tile.
//
// The kernel is issue-bound and very sensitive to its instruction footprint (L1.5 instruction cache:
// ~20 KB of hot code is the knee) and to registers (80 = 3 CTAs/SM; at 64 the spills cost more than the
// extra warps give).  Measured and rejected on the 8192^2 tiger: 64 registers / 4 CTAs per SM (+15 %
// time), an out-of-line copy of the tile code for heavy tiles (+10 %: they are the long pole, and the
// call boundary spills), prefetching the solid batches' words as well (+3 %: code size), one work
// counter instead of eight (the L2 atomic unit saturates: claims take microseconds), and for strokes a
// first pass that finds, per (segment, pixel row), the pixels certainly inside the stroke analytically
// so that the per-pixel distance work can skip them (+20 %: the sqrt and eight divisions of that test per
// pair cost more than the pixels they save; a single warp runs ~0.1 instructions per cycle, so anything
// that adds serial work to a tile with many segments lengthens the critical path); the pipeline steps
// (fine_step1 / fine_step3) as out-of-line functions (+12 %: the call sites sit in the hot path and the ABI
// saves ~30 live registers around each call), unlike the heavy-tile loops, the circle coverage and the
// dry-sub-queue walk, whose move out of line took the kernel from 137 to 122 us by itself.  Also rejected:
// rendering costly tiles as four independent row-band jobs (the per-item fixed cost is replicated in every band
// and dominates those tiles: no gain on the slowest tiles, +19 % on the frame) and pruning stroke pixels that an
// earlier segment has already saturated (+9 %: the mask bookkeeping costs more than the skipped distances).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/piet_metal_b200.h"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"
#include "pm_tile_logic.h"

#define PM_FULL_MASK 0xffffffffu

namespace {

typedef unsigned long long u64;

#define PM_FINE_WARPS 8
#define PM_FINE_COMPLEX_WARPS 6  // warps that prefer tiles with records; the rest prefer solid batches
#ifndef PM_FINE_TWO_LEVEL
#define PM_FINE_TWO_LEVEL 1      // 1: (pair, pixel) units handed out to the lanes; 0: a lane walks the pixels of its pair
#endif
#ifndef PM_FINE_BATCH_PIPELINE
#define PM_FINE_BATCH_PIPELINE 0 // 1: cnt / occ words of the solid batches prefetched too (measured: the extra code costs more
#endif                           //    in instruction-cache misses than the hidden latency gains; the solid warps are not critical)
#ifndef PM_FINE_TIMELINE
#define PM_FINE_TIMELINE 0       // 1 (debug builds, tools/fine_timeline.py): per-warp timestamps into PmFrameArgs::debug
#endif
#ifndef PM_FINE_BULK
#define PM_FINE_BULK 0           // 1 (NOT yet validated on a GPU; tools/cta_check.sh): the tile prefetch as TMA bulk copies
#endif                           //    (cp.async.bulk + mbarrier, UBLKCP in SASS) instead of 35 cp.async (LDGSTS) per tile
#ifndef PM_FINE_EARLY_CLAIM
#define PM_FINE_EARLY_CLAIM 0    // 1: the next-but-one tile is claimed at the start of a tile; 0: before the encode
#endif
#ifndef PM_FINE_CTAS
#define PM_FINE_CTAS 3           // CTAs per SM the kernel is compiled for (3: 80 registers; 4: 64 registers and PM_FINE_LIST_CAP <= 64)
#endif
#ifndef PM_FINE_LIST_CAP
#define PM_FINE_LIST_CAP 112     // overflow records per tile indexed in shared memory (extension block + 64 of the chain); the rest is re-walked      
#endif

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

// Per-warp shared-memory state.
//   acc / cov: coverage of the item being resolved, [pixel row][x] with the 4-pixel groups of a row
//     XOR-swizzled by the row so that the row-wise 128-bit accesses of the resolve and the scattered
//     atomics of the accumulation both spread over the banks.  For a stroke, acc holds the maximum
//     of ~bits(distance) instead (0 = no segment near), so one zero fill serves both.
//   rgb: lane-private, [channel][4-pixel group][lane].
//   rec / hdr: two prefetch buffers, each the 16 inline record slots of a tile and its cnt / occ / ovf words.
struct FineWarpSmem {
    int acc[256];
    int cov[256];
    float4 rgb[3][2][32];
    uint4 rec[2][2 * PM_TILE_SLOTS];
#if PM_FINE_BULK
    u64 hdr[2][6];   // per buffer: the aligned 16-byte pairs that contain the tile's cnt / occ / ovf words
    u64 mbar[2];     // per buffer: transaction barrier of the bulk copies (phase bits: w->st bits 14, 15)
#else
    u64 hdr[2][4];
#endif
    uint32_t idx[PM_FINE_LIST_CAP];  // heavy tiles: pool indices of the overflow records ...
    uint2 ovk[PM_FINE_LIST_CAP];     // ... and their (item, key), so that only the geometry is read from global memory
    uint32_t pkq[2];  // pipeline state of the walk over the tile list (see fine_entry)
    uint32_t st;
    uint32_t n_over, tail;  // heavy tiles: overflow records indexed in idx / ovk; 1 + pool index of the first one that did not fit
    uint32_t pad[3];
};

#if PM_FINE_TIMELINE
#define TL_MARK(k) do { const unsigned long long tl_t = fine_now(); tl_acc[(heavy ? 8 : 0) + (k)] += tl_t - tl_prev; tl_prev = tl_t; } while (0)
#else
#define TL_MARK(k) do { } while (0)
#endif
#if PM_FINE_TIMELINE
__device__ __forceinline__ unsigned long long fine_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
__device__ __forceinline__ int fine_swz(int row, int j) { return row * 16 + (j ^ (((row >> 1) & 3) << 2)); }

struct FineAcc {
    FineWarpSmem *w;
    __device__ __forceinline__ void near(int row, int j, int fx) { atomicAdd(&w->acc[fine_swz(row, j)], fx); }
    __device__ __forceinline__ void cover(int row, int j, int fx) { atomicAdd(&w->cov[fine_swz(row, j)], fx); }
    __device__ __forceinline__ void dist(int row, int j, float d) {  // d >= 0: unsigned order == float order
        atomicMax(reinterpret_cast<unsigned int *>(&w->acc[fine_swz(row, j)]), ~__float_as_uint(d));
    }
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// the k-th header word (0 cnt, 1 occ, 2 ovf) of the tile prefetched into buffer p
#if PM_FINE_BULK
#define FINE_HDR(w, p, k, tile) ((w)->hdr[p][2 * (k) + ((tile) & 1u)])
// TMA bulk copies (non-tensor): global -> shared, completion signalled on an mbarrier as a byte count.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, uint32_t bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
#define FINE_HDR(w, p, k, tile) ((w)->hdr[p][k])
#endif

// metal:563.  The debug render and PM_FLAG_EXACT_SRGB use this form.
template <bool EXACT>
__device__ __forceinline__ float linear_to_srgb(float v) {
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

// One channel, linear -> sRGB byte.  Default path: the scale to 0..255 folded into the curve and a
// saturating convert (negative, NaN -> 0; > 1 -> 255), no branch.
template <bool EXACT>
__device__ __forceinline__ uint32_t srgb_byte(float v) {
    if (EXACT) return pm_unorm8(linear_to_srgb<true>(v));
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
__device__ __forceinline__ uint32_t encode_pixel(float r, float g, float b) {
    return srgb_byte<EXACT>(r) | (srgb_byte<EXACT>(g) << 8) | (srgb_byte<EXACT>(b) << 16) | 0xff000000u;
}

// mix(x, y, a) with two FMAs, exact at a == 0 and a == 1 (metal:505, :543, :549: within an ulp or
// two of x + (y - x) * a)
__device__ __forceinline__ float mix_fma(float x, float y, float a) { return __fmaf_rn(a, y, __fmaf_rn(-a, x, x)); }

// unpack_unorm4x8_srgb_to_half; lut[0..255]: sRGB byte -> linear, lut[256..511]: alpha byte / 255
// (2 KB of global memory that lives in L1: the index is the same for the whole warp)
__device__ __forceinline__ void unpack_fg(const float *lut, uint32_t rgba, float fg[4]) {
    fg[0] = __ldg(&lut[rgba & 0xffu]);
    fg[1] = __ldg(&lut[(rgba >> 8) & 0xffu]);
    fg[2] = __ldg(&lut[(rgba >> 16) & 0xffu]);
    fg[3] = __ldg(&lut[256u + (rgba >> 24)]);
}

__device__ __forceinline__ PmRecord record_from(const uint4 a, const uint4 b) {
    PmRecord r;
    r.item = a.x; r.key = a.y; r.p[0] = pm_u2f(a.z); r.p[1] = pm_u2f(a.w);
    r.p[2] = pm_u2f(b.x); r.p[3] = pm_u2f(b.y); r.edge_y = pm_u2f(b.z); r.next = b.w;
    return r;
}
__device__ __forceinline__ PmRecord load_record(const PmRecord *pool, uint32_t idx) {
    const uint4 *src = reinterpret_cast<const uint4 *>(&pool[idx]);
    return record_from(src[0], src[1]);
}

// Phase A for up to 32 records held one per lane (`mine` = this lane holds a FILL*/LINE record of
// the current item).  Two levels of work distribution, because both the rows a segment crosses and
// the pixels of a row that need arithmetic vary from 0 to 16:
//   level 1: the (record, pixel row) pairs are enumerated across the lanes; a lane computes the
//            row-dependent part of its pair, adds the row's cover delta, and finds the pixel span
//            that needs per-pixel work (fill: the pixels near the segment; stroke: the pixels
//            within reach of it);
//   level 2: those (pair, pixel) units are enumerated across the lanes again, one pixel per lane.
// Owner lookup at both levels: exclusive prefix and payload packed into one word that is
// monotone in the lane, binary search with shuffles.
__device__ __forceinline__ void fine_pairs(FineWarpSmem *w, bool mine, uint32_t kind, float r_p0, float r_p1, float r_p2, float r_p3,
                                           float r_edge_y, bool stroke, float reach, float tile_x0, float tile_y0, uint32_t lane) {
    FineAcc acc{w};
    int ra = 1, rb = 0;
    if (mine) {
        if (stroke) pm_line_rows(r_p1, r_p3, reach, tile_y0, &ra, &rb);
        else pm_fill_rows(r_p1, r_p3, tile_y0, &ra, &rb);
    }
    const int cnt = rb >= ra ? rb - ra + 1 : 0;
    int incl = cnt;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(PM_FULL_MASK, incl, o);
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
#if PM_FINE_TWO_LEVEL
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
            int v = __shfl_up_sync(PM_FULL_MASK, incl2, o);
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
                if (u < total2) acc.dist(prow, j, pm_px_line_dist(e0, e1, e2, e3, tile_x0 + (float)j, tile_y0 + (float)prow));
            } else {
                PmFillRow fr;
                fr.wx = e2; fr.wy = e3;
                fr.tx = __shfl_sync(PM_FULL_MASK, d4, lo2);
                fr.ty = __shfl_sync(PM_FULL_MASK, d5, lo2);
                fr.active = true;
                if (u < total2) acc.near(prow, j, pm_fill_pair_px(e0, e1, tile_x0, j, fr));
            }
        }
#else
        if (q < total) {
            const int row = (o_key & 31) + (q - (o_key >> 5));
            if (stroke) pm_line_pair(acc, p, reach, row, tile_x0, tile_y0);
            else pm_fill_pair(acc, p, row, tile_x0, tile_y0);
        }
#endif
    }
    // FillEdge commands: one record at a time, lanes 0..15 take the 16 pixel rows
    if (!stroke) {
        for (uint32_t em = __ballot_sync(PM_FULL_MASK, mine && kind != PM_REC_FILL); em != 0; em &= em - 1) {
            const int src = __ffs(em) - 1;
            const uint32_t e_kind = __shfl_sync(PM_FULL_MASK, kind, src);
            const float e_y = __shfl_sync(PM_FULL_MASK, r_edge_y, src);
            if (lane < 16) pm_fill_edge_row(acc, e_kind, e_y, (int)lane, tile_y0);
        }
    }
}

// Pipeline state of a warp that walks the list of tiles with records.  While tile i is rendered, the
// header and inline records of tile i+1 are in flight into the other half of the shared-memory
// buffer; the queue position of tile i+2 is claimed when the coverage of tile i is done and its
// list entry is requested after tile i has been stored.  No global-memory latency of the chain
//   queue counter -> list entry -> cnt / occ / record slots
// is exposed once the pipeline runs.
// List order: the heavy tiles (more records than inline slots: coincident outlines, deep stacks)
// first, then the full list, in which the heavy ones are skipped.  Heavy first keeps a long tile
// from starting when everybody else is done.
//
// The pipeline's state lives in shared memory (w->st, w->pkq), not in registers: the tile code needs every
// register it can get, and state that is spilled to local memory costs an L1 miss each time it is touched.
//   w->pkq[b]  list entry (packed row, column) of the tile that uses prefetch buffer b next
//   w->st      bit b: that tile exists; bit 2+b: it comes from the full list (a heavy tile is skipped
//              there: pass 1 rendered it); bits 4..6: the sub-queue this warp claims from; bits 8..11:
//              sub-queues found empty so far
#define FINE_ST_VALID(b) (1u << (b))
#define FINE_ST_FULL(b) (4u << (b))
#define FINE_ST_HOME(st) (((st) >> 4) & (PM_FINE_SUBQ - 1u))
#define FINE_ST_DRY(st) (((st) >> 8) & 15u)

__device__ __forceinline__ uint32_t fine_claim(const PmFrameArgs &A, const FineWarpSmem *w, uint32_t lane) {
    // (atom.inc with a bound that is never reached, not atom.add: ptxas turns an add -- or an inc bounded by
    // 2^32-1 -- on a warp-uniform address into a warp-aggregated atomic followed by a shuffle of its result,
    // even from inline PTX, and that shuffle waits for the atomic right here)
    uint32_t k = 0;
    if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(k) : "l"(&A.queue->sub[FINE_ST_HOME(w->st)][0]) : "memory");
    return k;
}
// The warp's sub-queue has run dry: move on to the next ones until a position is found or all are dry
// (out of line: only the end of the frame comes here).
__device__ __noinline__ uint32_t fine_next_subqueue(const PmFrameArgs &A, uint32_t n_total, uint32_t *home_io, uint32_t *dry_io) {
    uint32_t home = *home_io, dry = *dry_io, q = 0xffffffffu;
    while (dry < PM_FINE_SUBQ) {
        if (++dry >= PM_FINE_SUBQ) break;
        home = (home + 1u) & (PM_FINE_SUBQ - 1u);
        uint32_t k = 0;
        if ((threadIdx.x & 31u) == 0) k = atomicAdd(&A.queue->sub[home][0], 1u);
        q = home + PM_FINE_SUBQ * __shfl_sync(PM_FULL_MASK, k, 0);
        if (q < n_total) break;
    }
    *home_io = home;
    *dry_io = dry;
    return q;
}
// Turns the claim (lane 0's `claim` = k in the warp's current sub-queue) into a list entry on its way into
// w->pkq[b] (no register waits for it).  The empty asm keeps the compiler from hoisting the shuffle up
// to the atomic.  A sub-queue that has run dry sends the warp on to the next one, until all are dry.
__device__ __forceinline__ void fine_entry(const PmFrameArgs &A, uint32_t claim, FineWarpSmem *w, uint32_t b, uint32_t n_heavy, uint32_t n_total) {
    asm volatile("" : "+r"(claim) : : "memory");
    uint32_t st = w->st & ~(FINE_ST_VALID(b) | FINE_ST_FULL(b));
    uint32_t home = FINE_ST_HOME(st), dry = FINE_ST_DRY(st);
    uint32_t q = home + PM_FINE_SUBQ * __shfl_sync(PM_FULL_MASK, claim, 0);
    if (q >= n_total) q = fine_next_subqueue(A, n_total, &home, &dry);
    st = (st & ~0xff0u) | (home << 4) | (dry << 8);
    if (q < n_total) {
        const bool full = q >= n_heavy;
        st |= FINE_ST_VALID(b) | (full ? FINE_ST_FULL(b) : 0u);
        if ((threadIdx.x & 31u) == 0) cp_async4(&w->pkq[b], full ? &A.complex_list[q - n_heavy] : &A.complex_list[A.n_rows * A.n_tx + q]);
        cp_async_commit();
    }
    __syncwarp();
    w->st = st;
}
// Starts the copy of a tile's header words and inline record slots into buffer `p`.
__device__ __forceinline__ void fine_prefetch(const PmFrameArgs &A, FineWarpSmem *w, uint32_t p, uint32_t pk, uint32_t lane) {
    const size_t tile = (size_t)(pk >> 16) * A.n_tx + (pk & 0xffffu);
#if PM_FINE_BULK
    if (lane == 0) {  // one thread arms the barrier with the byte count and issues the four copies
        const size_t pair = tile & ~(size_t)1;  // the header words are 8 bytes: fetch the aligned 16-byte pair that holds each
        mbar_expect_tx(&w->mbar[p], PM_TILE_SLOTS * (uint32_t)sizeof(PmRecord) + 48u);
        bulk_g2s(&w->rec[p][0], &A.pool[tile * PM_TILE_SLOTS], PM_TILE_SLOTS * (uint32_t)sizeof(PmRecord), &w->mbar[p]);
        bulk_g2s(&w->hdr[p][0], &A.cnt[pair], 16u, &w->mbar[p]);
        bulk_g2s(&w->hdr[p][2], &A.occ[pair], 16u, &w->mbar[p]);
        bulk_g2s(&w->hdr[p][4], &A.ovf[pair], 16u, &w->mbar[p]);
    }
#else
    cp_async16(&w->rec[p][lane], reinterpret_cast<const uint4 *>(&A.pool[tile * PM_TILE_SLOTS]) + lane);
    if (lane < 3) cp_async8(&w->hdr[p][lane], lane == 0 ? &A.cnt[tile] : (lane == 1 ? &A.occ[tile] : &A.ovf[tile]));
    cp_async_commit();
#endif
}
// Pipeline step, part 1 (once the tile's own set-up is done): the next tile's list entry, requested
// when the previous tile was stored, has arrived; start the copy of that tile's data.
__device__ __forceinline__ void fine_step1(const PmFrameArgs &A, FineWarpSmem *w, uint32_t p, uint32_t lane) {
    if (!(w->st & FINE_ST_VALID(p ^ 1u))) return;
    cp_async_wait_all();
    __syncwarp();
    fine_prefetch(A, w, p ^ 1u, w->pkq[p ^ 1u], lane);
}
// Pipeline step, parts 2 and 3: claim the position after the next tile; look its list entry up.
__device__ __forceinline__ uint32_t fine_step2(const PmFrameArgs &A, const FineWarpSmem *w, uint32_t p, uint32_t lane) {
    return (w->st & FINE_ST_VALID(p ^ 1u)) ? fine_claim(A, w, lane) : 0u;
}
__device__ __forceinline__ void fine_step3(const PmFrameArgs &A, uint32_t claim, FineWarpSmem *w, uint32_t p, uint32_t n_heavy, uint32_t n_total) {
    if (w->st & FINE_ST_VALID(p ^ 1u)) {
        fine_entry(A, claim, w, p, n_heavy, n_total);
    } else {
        __syncwarp();
        w->st &= ~(FINE_ST_VALID(p) | FINE_ST_FULL(p));
    }
}

// Cmd_Circle coverage of four consecutive pixels (out of line: rare, and the kernel is sensitive to the size of its hot path)
__device__ __noinline__ float4 fine_circle_alpha4(uint32_t bbox_lo, uint32_t bbox_hi, float px0, float py) {
    return make_float4(pm_px_circle_alpha(bbox_lo, bbox_hi, px0, py), pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 1.0f, py),
                       pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 2.0f, py), pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 3.0f, py));
}

// ---- heavy tiles (more records than inline slots, ~1 % of the tiles): out-of-line helpers, so that their
// loops stay out of the instruction footprint of the common path ----

// Indexes the overflow records: the extension block (positions 16..63, contiguous) and the chain behind it.
// w->n_over counts what was actually found (a frame whose overflow pool ran out has fewer records than cnt
// says; the host re-renders such a frame, it only must not fault).
#if PM_FINE_BULK
__device__ __noinline__ void fine_heavy_index(const PmFrameArgs &A, FineWarpSmem *w, uint32_t p, uint32_t n, uint32_t lane, uint32_t hdr_tile) {
#else
__device__ __noinline__ void fine_heavy_index(const PmFrameArgs &A, FineWarpSmem *w, uint32_t p, uint32_t n, uint32_t lane) {
#endif
    uint32_t n_over = 0, tail = 0;
    const u64 vw = FINE_HDR(w, p, 2, hdr_tile);
    uint32_t base1 = (uint32_t)(vw >> 32) == A.stamp ? (uint32_t)vw : 0u;
    if (base1 == PM_EXT_FAILED) base1 = 0;
    if (base1) {
        n_over = (n < PM_TILE_SLOTS + PM_EXT_SLOTS ? n : PM_TILE_SLOTS + PM_EXT_SLOTS) - PM_TILE_SLOTS;
        for (uint32_t i = lane; i < n_over; i += 32) {
            w->idx[i] = base1 + i;
            w->ovk[i] = *reinterpret_cast<const uint2 *>(&A.pool[base1 + i]);
        }
        if (n > PM_TILE_SLOTS + PM_EXT_SLOTS) {
            const uint32_t n_ext = n_over;
            uint32_t cur = A.pool[base1 - 1u].next;
            while (cur != 0 && n_over < PM_FINE_LIST_CAP) {
                if (lane == 0) w->idx[n_over] = cur - 1u;
                cur = A.pool[cur - 1u].next;
                n_over++;
            }
            tail = cur;
            __syncwarp();
            for (uint32_t i = n_ext + lane; i < n_over; i += 32) w->ovk[i] = *reinterpret_cast<const uint2 *>(&A.pool[w->idx[i]]);
        }
    }
    if (lane == 0) { w->n_over = n_over; w->tail = tail; }
    __syncwarp();
}
__device__ __noinline__ bool fine_heavy_has_draw(const PmFrameArgs &A, const FineWarpSmem *w, uint32_t occ_item1, uint32_t lane) {
    bool has_draw = false;
    for (uint32_t i = lane; i < w->n_over; i += 32) {
        const uint2 ik = w->ovk[i];
        if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
    }
    for (uint32_t cur = w->tail; cur != 0; cur = A.pool[cur - 1u].next) {
        const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[cur - 1u]);
        if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = true;
    }
    return has_draw;
}
// this lane's candidate for the next item in painter's order among the overflow records
__device__ __noinline__ uint32_t fine_heavy_min_item(const PmFrameArgs &A, const FineWarpSmem *w, uint32_t occ_item1, bool first, uint32_t last_item,
                                                     uint32_t cur_item, uint32_t lane) {
    for (uint32_t i = lane; i < w->n_over; i += 32) {
        const uint32_t it = w->ovk[i].x;
        if (it >= occ_item1 && (first || it > last_item) && it < cur_item) cur_item = it;
    }
    for (uint32_t cur = w->tail; cur != 0; cur = A.pool[cur - 1u].next) {
        const uint32_t it = A.pool[cur - 1u].item;
        if (it >= occ_item1 && (first || it > last_item) && it < cur_item) cur_item = it;
    }
    return cur_item;
}
// the item's closing record among the overflow records: (kind, w0, w1), kind 0 if there is none
__device__ __noinline__ uint3 fine_heavy_trailer(const PmFrameArgs &A, const FineWarpSmem *w, uint32_t cur_item, uint32_t lane) {
    uint32_t t_kind = 0, t_w0 = 0, t_w1 = 0;
    for (uint32_t i = lane; i < w->n_over; i += 32) {
        const uint2 ik = w->ovk[i];
        if (ik.x == cur_item && (ik.y & 15u) >= PM_REC_CIRCLE) {
            const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[w->idx[i]]);
            t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w;
        }
    }
    for (uint32_t cur = w->tail; cur != 0; cur = A.pool[cur - 1u].next) {
        const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[cur - 1u]);
        if (a.x == cur_item && (a.y & 15u) >= PM_REC_CIRCLE) { t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w; }
    }
    const uint32_t src = __ffs(__ballot_sync(PM_FULL_MASK, t_kind != 0));
    if (src == 0) return make_uint3(0, 0, 0);
    return make_uint3(__shfl_sync(PM_FULL_MASK, t_kind, src - 1), __shfl_sync(PM_FULL_MASK, t_w0, src - 1), __shfl_sync(PM_FULL_MASK, t_w1, src - 1));
}
// phase A over the overflow records of the item, 32 at a time; those beyond the shared-memory index one at a time
__device__ __noinline__ void fine_heavy_pairs(const PmFrameArgs &A, FineWarpSmem *w, uint32_t cur_item, bool stroke, float reach, float tile_x0, float tile_y0,
                                              uint32_t lane) {
    const uint32_t n_over = w->n_over;
    for (uint32_t i0 = 0; i0 < n_over; i0 += 32) {
        const uint32_t i = i0 + lane;
        PmRecord rc;
        rc.item = 0xffffffffu; rc.key = 0; rc.p[0] = rc.p[1] = rc.p[2] = rc.p[3] = 0.0f; rc.edge_y = 0.0f; rc.next = 0;
        if (i < n_over) {
            const uint2 ik = w->ovk[i];
            if (ik.x == cur_item && (ik.y & 15u) <= PM_REC_LINE) rc = load_record(A.pool, w->idx[i]);
        }
        const bool mine = rc.item == cur_item && (rc.key & 15u) <= PM_REC_LINE;
        if (__any_sync(PM_FULL_MASK, mine))
            fine_pairs(w, mine, rc.key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, reach, tile_x0, tile_y0, lane);
    }
    for (uint32_t cur = w->tail; cur != 0;) {
        PmRecord rc = load_record(A.pool, cur - 1u);
        cur = rc.next;
        if (rc.item == cur_item && (rc.key & 15u) <= PM_REC_LINE)
            fine_pairs(w, lane == 0, rc.key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, reach, tile_x0, tile_y0, lane);
    }
}

// One tile that owns records; its header and inline records are in buffer `p` of w.  All 32 lanes
// execute this together.  Records are handled in chunks of 32, one per lane; chunk 0 is the inline
// slots.  Pixel layout: lane l owns pixel row (l >> 1), pixels 8*(l & 1) .. +7.
template <bool F32, bool EXACT>
__device__ __forceinline__ void fine_complex_tile(const PmFrameArgs &A, FineWarpSmem *w, uint32_t p, uint32_t lane, uint32_t n_heavy, uint32_t n_total
#if PM_FINE_TIMELINE
                                                  , unsigned long long *tl_acc
#endif
                                                  ) {
#if PM_FINE_TIMELINE
    unsigned long long tl_prev = fine_now();
#endif
    cp_async_wait_all();
    __syncwarp();
    const uint32_t packed_tile = w->pkq[p];
    const bool skip_heavy = (w->st & FINE_ST_FULL(p)) != 0;
    const uint32_t trow = packed_tile >> 16, tx = packed_tile & 0xffffu;
#if PM_FINE_BULK
    const uint32_t hdr_tile = trow * A.n_tx + tx;
    {   // the bulk copies of this tile have landed (phase bit of the buffer: w->st bit 14 + p)
        const uint32_t st = w->st;
        mbar_wait(&w->mbar[p], (st >> (14u + p)) & 1u);
        __syncwarp();
        w->st = st ^ (1u << (14u + p));
    }
#endif
#if PM_FINE_EARLY_CLAIM
    const uint32_t claim = fine_step2(A, w, p, lane);
#endif
    const uint4 *rec = w->rec[p];
    const u64 cw = FINE_HDR(w, p, 0, hdr_tile), ow = FINE_HDR(w, p, 1, hdr_tile);
    const uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;
    const bool heavy = n > PM_TILE_SLOTS;
#if PM_CTA_TILES
    const bool by_cta = (w->st & 0x1000u) != 0 && n >= PM_CTA_MIN && n <= PM_CTA_CAP &&
                        (uint32_t)(FINE_HDR(w, p, 2, hdr_tile) >> 32) == A.stamp && (uint32_t)FINE_HDR(w, p, 2, hdr_tile) != PM_EXT_FAILED;  // the test of fine_cta_tile
    if ((heavy && skip_heavy) || by_cta) {
#else
    if (heavy && skip_heavy) {  // pass 1 rendered it
#endif
        TL_MARK(6);
        fine_step1(A, w, p, lane);
#if PM_FINE_EARLY_CLAIM
        fine_step3(A, claim, w, p, n_heavy, n_total);
#else
        fine_step3(A, fine_step2(A, w, p, lane), w, p, n_heavy, n_total);
#endif
        return;
    }
    uint32_t occ_rgba = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (occ_item1) occ_rgba = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));

    const uint32_t n_inline = n < PM_TILE_SLOTS ? n : PM_TILE_SLOTS;
#if PM_FINE_BULK
    if (heavy) fine_heavy_index(A, w, p, n, lane, hdr_tile);
#else
    if (heavy) fine_heavy_index(A, w, p, n, lane);
#endif

    // this lane's inline record: item and key stay in registers, the geometry is re-read when needed
    uint32_t my_item = 0xffffffffu, my_key = 0;
    if (lane < n_inline) {
        const uint2 ik = *reinterpret_cast<const uint2 *>(&rec[2 * lane]);
        my_item = ik.x; my_key = ik.y;
    }
    if (my_item < occ_item1) my_item = 0xffffffffu;  // below the topmost opaque cover: rewound away (metal:132-135)

    bool has_draw = my_item != 0xffffffffu && (my_key & 15u) != PM_REC_SOLID;
    if (heavy && fine_heavy_has_draw(A, w, occ_item1, lane)) has_draw = true;
    has_draw = __any_sync(PM_FULL_MASK, has_draw);

    const uint32_t prow = lane >> 1, half = lane & 1u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + half * 8u) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) +
                     (tx * PM_TILE_W + half * 8u);

    if (!has_draw) {
        TL_MARK(6);
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
        fine_step1(A, w, p, lane);
#if PM_FINE_EARLY_CLAIM
        fine_step3(A, claim, w, p, n_heavy, n_total);
#else
        fine_step3(A, fine_step2(A, w, p, lane), w, p, n_heavy, n_total);
#endif
        return;
    }

    const float *lut = A.srgb_lut;
    {   // the rewound list starts with the cover's Cmd_Solid (metal:136-142, :546-551) over white (metal:470)
        float b0 = 1.0f, b1 = 1.0f, b2 = 1.0f;
        if (occ_item1) {
            float fg[4];
            unpack_fg(lut, occ_rgba, fg);
            b0 = mix_fma(1.0f, fg[0], fg[3]); b1 = mix_fma(1.0f, fg[1], fg[3]); b2 = mix_fma(1.0f, fg[2], fg[3]);
        }
        #pragma unroll
        for (int g = 0; g < 2; g++) {
            w->rgb[0][g][lane] = make_float4(b0, b0, b0, b0);
            w->rgb[1][g][lane] = make_float4(b1, b1, b1, b1);
            w->rgb[2][g][lane] = make_float4(b2, b2, b2, b2);
        }
    }
    fine_step1(A, w, p, lane);
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);  // scene coordinates
    // this lane's two 4-pixel groups of the coverage arrays (word offsets; cov = acc + 256)
    const int my_off0 = fine_swz((int)prow, (int)half * 8), my_off1 = fine_swz((int)prow, (int)half * 8 + 4);

    TL_MARK(0);
    // items in painter's order: repeatedly take the smallest item id above the last one done
    uint32_t last_item = 0;
    bool first = true;
    for (;;) {
        uint32_t cur_item = (first || my_item > last_item) ? my_item : 0xffffffffu;
        if (heavy) cur_item = fine_heavy_min_item(A, w, occ_item1, first, last_item, cur_item, lane);
        cur_item = __reduce_min_sync(PM_FULL_MASK, cur_item);
        if (cur_item == 0xffffffffu) break;
        first = false;
        last_item = cur_item;

        // the item's closing record says what it is (DrawFill / Stroke / Circle / Solid)
        uint32_t t_kind = 0, t_w0 = 0, t_w1 = 0;
        {
            const uint32_t m = __ballot_sync(PM_FULL_MASK, my_item == cur_item && (my_key & 15u) >= PM_REC_CIRCLE);
            if (m) {  // among the inline records: everybody reads it from shared memory
                const uint4 a = rec[2 * (__ffs(m) - 1)];
                t_kind = a.y & 15u; t_w0 = a.z; t_w1 = a.w;
            } else if (heavy) {
                const uint3 tr = fine_heavy_trailer(A, w, cur_item, lane);
                if (tr.x == 0) continue;  // cannot happen for a well-formed list
                t_kind = tr.x; t_w0 = tr.y; t_w1 = tr.z;
            } else {
                continue;  // cannot happen for a well-formed list
            }
        }

        TL_MARK(1);
        float fg[4] = {0.0f, 0.0f, 0.0f, 1.0f};  // Cmd_Circle paints black (metal:491)
        const bool stroke = t_kind == PM_REC_STROKE, fill = t_kind == PM_REC_DRAWFILL;
        const float half_width = pm_u2f(t_w0);
        int run = 0;  // fill: cover entering this lane's pixels from the left
        if (t_kind != PM_REC_CIRCLE) unpack_fg(lut, t_w1, fg);
        if (fill || stroke) {
            const float reach = half_width + 0.5f;
            // phase A: coverage of the item's segments, 32 records at a time
            {
                const bool mine = my_item == cur_item && (my_key & 15u) <= PM_REC_LINE;
                if (__any_sync(PM_FULL_MASK, mine)) {
                    const uint4 a = rec[2 * (lane & (PM_TILE_SLOTS - 1))], b = rec[2 * (lane & (PM_TILE_SLOTS - 1)) + 1];
                    fine_pairs(w, mine, my_key & 15u, pm_u2f(a.z), pm_u2f(a.w), pm_u2f(b.x), pm_u2f(b.y), pm_u2f(b.z), stroke, reach,
                               tile_x0, tile_y0, lane);
                }
            }
            TL_MARK(stroke ? 2 : 7);
            if (heavy) fine_heavy_pairs(A, w, cur_item, stroke, reach, tile_x0, tile_y0, lane);
            __syncwarp();
            TL_MARK(3);
            if (fill) {  // covers of the left half of the pixel row carry into the right half
                const int4 c0 = *reinterpret_cast<const int4 *>(&w->cov[my_off0]), c1 = *reinterpret_cast<const int4 *>(&w->cov[my_off1]);
                const int sum = ((c0.x + c0.y) + (c0.z + c0.w)) + ((c1.x + c1.y) + (c1.z + c1.w));
                const int other = __shfl_xor_sync(PM_FULL_MASK, sum, 1);
                run = half ? other : 0;
            }
        }
        // phase B: resolve this lane's 8 pixels, clear their coverage for the next item, and blend
        // (metal:505, :543, :549).  Two rounds of 4 pixels: the kernel is sensitive to its code size.
        #pragma unroll 1
        for (int g = 0; g < 2; g++) {
            float al[4];
            if (fill || stroke) {
                int4 *pa = reinterpret_cast<int4 *>(&w->acc[g ? my_off1 : my_off0]);
                const int4 a = *pa;
                *pa = make_int4(0, 0, 0, 0);
                if (fill) {
                    int4 *pc = reinterpret_cast<int4 *>(&w->cov[g ? my_off1 : my_off0]);
                    const int4 c = *pc;
                    *pc = make_int4(0, 0, 0, 0);
                    const int backdrop = (int)t_w0;
                    run += c.x; al[0] = pm_resolve_fill_alpha(a.x + run, backdrop);
                    run += c.y; al[1] = pm_resolve_fill_alpha(a.y + run, backdrop);
                    run += c.z; al[2] = pm_resolve_fill_alpha(a.z + run, backdrop);
                    run += c.w; al[3] = pm_resolve_fill_alpha(a.w + run, backdrop);
                } else {  // renderDf, metal:58-60
                    const float lim = half_width + 0.5f;
                    al[0] = a.x ? pm_saturate(lim - __uint_as_float(~(uint32_t)a.x)) : 0.0f;
                    al[1] = a.y ? pm_saturate(lim - __uint_as_float(~(uint32_t)a.y)) : 0.0f;
                    al[2] = a.z ? pm_saturate(lim - __uint_as_float(~(uint32_t)a.z)) : 0.0f;
                    al[3] = a.w ? pm_saturate(lim - __uint_as_float(~(uint32_t)a.w)) : 0.0f;
                }
            } else if (t_kind == PM_REC_CIRCLE) {
                const float4 ca = fine_circle_alpha4(t_w0, t_w1, tile_x0 + (float)(half * 8u + 4u * (uint32_t)g), tile_y0 + (float)prow);
                al[0] = ca.x; al[1] = ca.y; al[2] = ca.z; al[3] = ca.w;
            } else {  // PM_REC_SOLID: a translucent full cover
                al[0] = al[1] = al[2] = al[3] = 1.0f;
            }
            #pragma unroll
            for (int j = 0; j < 4; j++) al[j] *= fg[3];
            #pragma unroll
            for (int k = 0; k < 3; k++) {
                float4 v = w->rgb[k][g][lane];
                v.x = mix_fma(v.x, fg[k], al[0]);
                v.y = mix_fma(v.y, fg[k], al[1]);
                v.z = mix_fma(v.z, fg[k], al[2]);
                v.w = mix_fma(v.w, fg[k], al[3]);
                w->rgb[k][g][lane] = v;
            }
        }
        __syncwarp();
        TL_MARK(4);
    }

    // the position after the next tile is claimed here and looked up after the encode: the claim's result
    // must stay in its register until then (anything that touches it -- a spill included -- waits for the
    // atomic), and this is the stretch of the tile with the fewest live values
#if PM_FINE_EARLY_CLAIM
    fine_step3(A, claim, w, p, n_heavy, n_total);
#else
    const uint32_t claim = fine_step2(A, w, p, lane);
#endif

    #pragma unroll 1
    for (int g = 0; g < 2; g++) {
        const float4 r = w->rgb[0][g][lane], gg = w->rgb[1][g][lane], b = w->rgb[2][g][lane];
        const uint4 px = make_uint4(encode_pixel<EXACT>(r.x, gg.x, b.x), encode_pixel<EXACT>(r.y, gg.y, b.y),
                                    encode_pixel<EXACT>(r.z, gg.z, b.z), encode_pixel<EXACT>(r.w, gg.w, b.w));
        __stcs(reinterpret_cast<uint4 *>(dst) + g, px);
        if (F32) {  // debug render: the un-quantised values
            dst32[4 * g + 0] = make_float4(linear_to_srgb<EXACT>(r.x), linear_to_srgb<EXACT>(gg.x), linear_to_srgb<EXACT>(b.x), 1.0f);
            dst32[4 * g + 1] = make_float4(linear_to_srgb<EXACT>(r.y), linear_to_srgb<EXACT>(gg.y), linear_to_srgb<EXACT>(b.y), 1.0f);
            dst32[4 * g + 2] = make_float4(linear_to_srgb<EXACT>(r.z), linear_to_srgb<EXACT>(gg.z), linear_to_srgb<EXACT>(b.z), 1.0f);
            dst32[4 * g + 3] = make_float4(linear_to_srgb<EXACT>(r.w), linear_to_srgb<EXACT>(gg.w), linear_to_srgb<EXACT>(b.w), 1.0f);
        }
    }
#if !PM_FINE_EARLY_CLAIM
    fine_step3(A, claim, w, p, n_heavy, n_total);
#endif
    TL_MARK(5);
}

#if PM_CTA_TILES
// ---- costly tiles rendered by a whole CTA (see PM_CTA_TILES in pm_kernels.h; not yet validated on a GPU) ----
// Thread t owns pixel (row t >> 4, column t & 15) and holds its linear colour in three registers.  The tile's
// records (at most PM_CTA_CAP) are indexed in shared memory borrowed from the per-warp state of warps 1 and 2
// (their colour planes, which a warp initialises before every use); thread t holds (item, key) of record t, so
// warp w has records 32 w .. 32 w + 31, one per lane, and runs the same fine_pairs() as the per-warp path on them -- all eight warps accumulating into
// warp 0's coverage arrays with shared-memory atomics.  Per item: block-wide minimum of the item ids, coverage,
// barrier, resolve + blend one pixel per thread, barrier.  Same functions, same operand order, integer coverage
// sums: the pixels are bit-identical to the per-warp path.
struct FineCtaShared {      // lives in warp 2's colour planes
    u64 cw, ow, vw;
    uint32_t red[PM_FINE_WARPS];   // block reductions
    uint32_t t_kind, t_w0, t_w1, pad;
};

__device__ __forceinline__ uint32_t fine_block_min(uint32_t v, volatile uint32_t *red, uint32_t lane, uint32_t warp) {
    v = __reduce_min_sync(PM_FULL_MASK, v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    uint32_t m = red[0];
    #pragma unroll
    for (int k = 1; k < PM_FINE_WARPS; k++) m = red[k] < m ? red[k] : m;
    __syncthreads();
    return m;
}

template <bool F32, bool EXACT>
__device__ __noinline__ void fine_cta_tile(const PmFrameArgs &A, FineWarpSmem *ws, uint32_t packed_tile) {
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t trow = packed_tile >> 16, tx = packed_tile & 0xffffu;
    const size_t tile = (size_t)trow * A.n_tx + tx;
    FineWarpSmem *acc_w = &ws[0];                                          // shared coverage accumulators
    uint32_t *idx = reinterpret_cast<uint32_t *>(&ws[1].rgb[0][0][0]);     // [PM_CTA_CAP] pool indices
    FineCtaShared *sh = reinterpret_cast<FineCtaShared *>(&ws[2].rgb[0][0][0]);
    if (t == 0) { sh->cw = A.cnt[tile]; sh->ow = A.occ[tile]; sh->vw = A.ovf[tile]; }
    __syncthreads();
    const u64 cw = sh->cw, ow = sh->ow, vw = sh->vw;
    const uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;
    uint32_t base1 = (uint32_t)(vw >> 32) == A.stamp ? (uint32_t)vw : 0u;
    if (base1 == PM_EXT_FAILED) base1 = 0;
    if (n < PM_CTA_MIN || n > PM_CTA_CAP || base1 == 0) { __syncthreads(); return; }  // (left to the per-warp path, which applies the same test)
    // record index: inline slots, extension block, chain
    if (t < n) {
        if (t < PM_TILE_SLOTS) idx[t] = (uint32_t)tile * PM_TILE_SLOTS + t;
        else if (t < PM_TILE_SLOTS + PM_EXT_SLOTS) idx[t] = base1 + (t - PM_TILE_SLOTS);
    }
    if (t == 0 && n > PM_TILE_SLOTS + PM_EXT_SLOTS) {
        uint32_t cur = A.pool[base1 - 1u].next, k = PM_TILE_SLOTS + PM_EXT_SLOTS;
        for (; cur != 0 && k < n; k++) { idx[k] = cur - 1u; cur = A.pool[cur - 1u].next; }
        for (; k < n; k++) idx[k] = 0xffffffffu;  // (a frame whose overflow pool ran out: fewer links than cnt says)
    }
    __syncthreads();
    uint32_t my_item = 0xffffffffu, my_key = 0;
    if (t < n && idx[t] != 0xffffffffu) {
        const uint2 v = *reinterpret_cast<const uint2 *>(&A.pool[idx[t]]);
        my_item = v.x; my_key = v.y;
    }
    if (my_item < occ_item1) my_item = 0xffffffffu;  // below the topmost opaque cover: rewound away (metal:132-135)
    const int has_draw = __syncthreads_or(my_item != 0xffffffffu && (my_key & 15u) != PM_REC_SOLID);

    const uint32_t prow = t >> 4, px = t & 15u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + px) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) + (tx * PM_TILE_W + px);
    uint32_t occ_rgba = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (occ_item1) occ_rgba = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));
    if (!has_draw) {  // the tile Bails and shows solidColor (metal:145-147, :34-44)
        *reinterpret_cast<uint32_t *>(dst) = occ_rgba;
        if (F32) *dst32 = make_float4((float)(occ_rgba & 0xff) / 255.0f, (float)((occ_rgba >> 8) & 0xff) / 255.0f,
                                      (float)((occ_rgba >> 16) & 0xff) / 255.0f, (float)(occ_rgba >> 24) / 255.0f);
        __syncthreads();
        return;
    }
    const float *lut = A.srgb_lut;
    float c0 = 1.0f, c1 = 1.0f, c2 = 1.0f;  // metal:470; then the cover's Cmd_Solid (metal:136-142, :546-551)
    if (occ_item1) {
        float fg[4];
        unpack_fg(lut, occ_rgba, fg);
        c0 = mix_fma(1.0f, fg[0], fg[3]); c1 = mix_fma(1.0f, fg[1], fg[3]); c2 = mix_fma(1.0f, fg[2], fg[3]);
    }
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);
    const int my_cell = fine_swz((int)prow, (int)px);
    uint32_t last_item = 0;
    bool first = true;
    for (;;) {
        const uint32_t cand = (first || my_item > last_item) ? my_item : 0xffffffffu;
        const uint32_t cur_item = fine_block_min(cand, sh->red, lane, warp);
        if (cur_item == 0xffffffffu) break;
        first = false;
        last_item = cur_item;
        // the item's closing record
        if (t == 0) sh->t_kind = 0;
        __syncthreads();
        if (my_item == cur_item && (my_key & 15u) >= PM_REC_CIRCLE) {
            const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[idx[t]]);
            sh->t_kind = a.y & 15u; sh->t_w0 = a.z; sh->t_w1 = a.w;
        }
        __syncthreads();
        const uint32_t t_kind = sh->t_kind, t_w0 = sh->t_w0, t_w1 = sh->t_w1;
        if (t_kind == 0) continue;  // cannot happen for a well-formed list (uniform: every thread reads the same word)
        float fg[4] = {0.0f, 0.0f, 0.0f, 1.0f};  // Cmd_Circle paints black (metal:491)
        const bool stroke = t_kind == PM_REC_STROKE, fill = t_kind == PM_REC_DRAWFILL;
        const float half_width = pm_u2f(t_w0);
        if (t_kind != PM_REC_CIRCLE) unpack_fg(lut, t_w1, fg);
        float al = 1.0f;  // PM_REC_SOLID: a translucent full cover
        if (fill || stroke) {
            const bool mine = my_item == cur_item && (my_key & 15u) <= PM_REC_LINE;
            if (__any_sync(PM_FULL_MASK, mine)) {
                PmRecord rc;
                rc.p[0] = rc.p[1] = rc.p[2] = rc.p[3] = 0.0f; rc.edge_y = 0.0f;
                if (mine) rc = load_record(A.pool, idx[t]);
                fine_pairs(acc_w, mine, my_key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, half_width + 0.5f, tile_x0, tile_y0, lane);
            }
            __syncthreads();
            const int a = acc_w->acc[my_cell];
            acc_w->acc[my_cell] = 0;
            if (fill) {
                const int c = acc_w->cov[my_cell];
                acc_w->cov[my_cell] = 0;
                int run = c;  // covers of the pixels to the left carry into this one: inclusive scan over the row's 16 lanes
                #pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int v = __shfl_up_sync(PM_FULL_MASK, run, o, 16);
                    if ((int)px >= o) run += v;
                }
                al = pm_resolve_fill_alpha(a + run, (int)t_w0);
            } else {
                al = a ? pm_saturate(half_width + 0.5f - __uint_as_float(~(uint32_t)a)) : 0.0f;  // renderDf, metal:58-60
            }
        } else if (t_kind == PM_REC_CIRCLE) {
            al = pm_px_circle_alpha(t_w0, t_w1, tile_x0 + (float)px, tile_y0 + (float)prow);
        }
        al *= fg[3];
        c0 = mix_fma(c0, fg[0], al); c1 = mix_fma(c1, fg[1], al); c2 = mix_fma(c2, fg[2], al);
        __syncthreads();  // the accumulators are clear again before the next item adds to them
    }
    *reinterpret_cast<uint32_t *>(dst) = encode_pixel<EXACT>(c0, c1, c2);
    if (F32) *dst32 = make_float4(linear_to_srgb<EXACT>(c0), linear_to_srgb<EXACT>(c1), linear_to_srgb<EXACT>(c2), 1.0f);
    __syncthreads();
}
#endif  // PM_CTA_TILES

__device__ __forceinline__ uint32_t fine_batch_claim(const PmFrameArgs &A, uint32_t lane) {
    uint32_t k = 0;  // (atom.inc: see fine_claim)
    if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(k) : "l"(&A.queue->batch_next) : "memory");
    return k;
}
#if PM_FINE_BATCH_PIPELINE
// The solid batches are pipelined like the tiles with records: while batch i is stored, the cnt / occ
// words of batch i+1 are in flight into the other prefetch buffer and the position of batch i+2 is claimed.
__device__ __forceinline__ void fine_batch_prefetch(const PmFrameArgs &A, FineWarpSmem *w, uint32_t b, uint32_t batch, uint32_t batches_per_row,
                                                    uint32_t n_batches, uint32_t lane) {
    if (batch < n_batches) {
        const uint32_t row = batch / batches_per_row;
        const uint32_t t = (batch - row * batches_per_row) * 32u + lane;
        if (t < A.n_tx) {
            const size_t tile = (size_t)row * A.n_tx + t;
            cp_async8(&w->rec[b][lane], &A.cnt[tile]);
            cp_async8(reinterpret_cast<unsigned char *>(&w->rec[b][lane]) + 8, &A.occ[tile]);
        }
    }
    cp_async_commit();  // (an empty group when there is nothing to fetch: the group count stays in step)
}
#endif

// 32 consecutive tiles of one tile row; the solid ones are written row-wise: each store
// instruction covers 512 contiguous bytes (128 pixels) of one pixel row.
template <bool F32>
__device__ __forceinline__ void fine_solid_batch(const PmFrameArgs &A, uint32_t batch, uint32_t batches_per_row, const uint4 *words, uint32_t lane) {
    const uint32_t row = batch / batches_per_row;
    const uint32_t t0 = (batch - row * batches_per_row) * 32u;
    const uint32_t t = t0 + lane;
    const bool valid = t < A.n_tx;
    bool solid = false;
    uint32_t colour = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (valid) {
#if PM_FINE_BATCH_PIPELINE
        const uint4 ww = words[lane];  // this lane's tile: cnt word, occ word (fine_batch_prefetch)
        const u64 cw = ((u64)ww.y << 32) | ww.x, ow = ((u64)ww.w << 32) | ww.z;
#else
        (void)words;
        const size_t tile = (size_t)row * A.n_tx + t;
        const u64 cw = A.cnt[tile], ow = A.occ[tile];
#endif
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
__global__ void __launch_bounds__(PM_FINE_WARPS * 32, PM_FINE_CTAS) k_fine(const PmFrameArgs A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FineWarpSmem *w = reinterpret_cast<FineWarpSmem *>(s_raw) + warp;
    for (uint32_t i = lane; i < 256; i += 32) { w->acc[i] = 0; w->cov[i] = 0; }
#if PM_FINE_BULK
    if (lane == 0) {
        mbar_init(&w->mbar[0], 1);
        mbar_init(&w->mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (programmatic dependent launch, see pm_kernels.cu)
    asm volatile("griddepcontrol.wait;" ::: "memory");               // binning has finished
    const uint32_t n_complex = A.counters->n_complex, n_heavy = A.counters->n_heavy;
    const uint32_t n_total = n_complex + n_heavy;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.report->n_complex = n_complex;
        A.report->n_overflow = A.counters->n_overflow;
        A.report->frame = A.stamp;
        A.counters_next->n_complex = 0;
        A.counters_next->n_overflow = 0;
        A.counters_next->n_heavy = 0;
    }
    __syncwarp();
#if PM_CTA_TILES
    {   // costly tiles first, a whole CTA each, claimed dynamically (the CTAs that are not resident yet must not own any)
        const uint32_t n_costly = A.counters->n_costly;
        const bool cta_mode = n_costly != 0 && n_costly <= PM_CTA_MAX_PER_CTA * gridDim.x;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.counters_next->n_costly = 0;
        __shared__ uint32_t s_costly;
        __syncthreads();  // (every warp has cleared its accumulators)
        while (cta_mode) {
            if (threadIdx.x == 0) s_costly = atomicAdd(&A.queue->costly_next, 1u);
            __syncthreads();
            const uint32_t h = s_costly;
            __syncthreads();
            if (h >= n_costly) break;
            fine_cta_tile<F32, EXACT>(A, reinterpret_cast<FineWarpSmem *>(s_raw), A.complex_list[2u * A.n_rows * A.n_tx + h]);
        }
        if (lane == 0) w->st = cta_mode ? 0x1000u : 0u;  // (the pipeline adds its own bits below)
        __syncwarp();
    }
#endif
    const uint32_t batches_per_row = (A.n_tx + 31u) / 32u;
    const uint32_t n_batches = batches_per_row * A.n_rows;
    bool complex_left = true, batches_left = true;
    // warps 3 and 7 (one of the SM's four schedulers) prefer the solid batches, the rest the tiles with records
    const bool prefer_complex = (warp & 3u) != 3u;
    uint32_t p = 0;
    bool started = false;
#if PM_FINE_TIMELINE
    const unsigned long long tl_begin = fine_now();
    unsigned long long tl_last = tl_begin, tl_long = 0;
    uint32_t tl_tiles = 0, tl_long_pk = 0;
    unsigned long long tl_acc[16];
    for (int k = 0; k < 16; k++) tl_acc[k] = 0;
#endif
#if PM_FINE_BATCH_PIPELINE
    uint32_t b_cur = 0, b_next = 0, bp = 0;
    bool b_started = false;
#endif
#if PM_CTA_TILES
    if (lane == 0) w->st = (w->st & 0x1000u) | ((blockIdx.x & (PM_FINE_SUBQ - 1u)) << 4);
#else
    if (lane == 0) w->st = (blockIdx.x & (PM_FINE_SUBQ - 1u)) << 4;
#endif
    __syncwarp();
    while (complex_left || batches_left) {
        const bool take_complex = complex_left && (prefer_complex || !batches_left);
        if (take_complex) {
            // (one call site for the tile code: the kernel is sensitive to its instruction footprint)
            if (!started) {  // fill the pipeline: this tile's data, the next tile's list entry
                started = true;
                fine_entry(A, fine_claim(A, w, lane), w, p, n_heavy, n_total);
                if (!(w->st & FINE_ST_VALID(p))) { complex_left = false; continue; }
                cp_async_wait_all();
                __syncwarp();
                fine_prefetch(A, w, p, w->pkq[p], lane);
                fine_entry(A, fine_claim(A, w, lane), w, p ^ 1u, n_heavy, n_total);
            }
#if PM_FINE_TIMELINE
            const unsigned long long tl0 = fine_now();
            const uint32_t tl_pk = w->pkq[p];
#endif
#if PM_FINE_TIMELINE
            fine_complex_tile<F32, EXACT>(A, w, p, lane, n_heavy, n_total, tl_acc);
#else
            fine_complex_tile<F32, EXACT>(A, w, p, lane, n_heavy, n_total);
#endif
#if PM_FINE_TIMELINE
            {
                const unsigned long long tl1 = fine_now();
                tl_tiles++;
                tl_last = tl1;
                if (tl1 - tl0 > tl_long) { tl_long = tl1 - tl0; tl_long_pk = tl_pk; }
            }
#endif
            p ^= 1u;
            if (!(w->st & FINE_ST_VALID(p))) complex_left = false;
        } else {
#if PM_FINE_BATCH_PIPELINE
            if (!b_started) {  // fill the pipeline: two batches claimed, their words on the way
                b_started = true;
                cp_async_wait_all();
                __syncwarp();
                b_cur = __shfl_sync(PM_FULL_MASK, fine_batch_claim(A, lane), 0);
                fine_batch_prefetch(A, w, 0, b_cur, batches_per_row, n_batches, lane);
                b_next = __shfl_sync(PM_FULL_MASK, fine_batch_claim(A, lane), 0);
                fine_batch_prefetch(A, w, 1, b_next, batches_per_row, n_batches, lane);
                bp = 0;
            }
            if (b_cur >= n_batches) { batches_left = false; cp_async_wait_all(); __syncwarp(); continue; }
            uint32_t claim = fine_batch_claim(A, lane);  // the batch after the next one
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            fine_solid_batch<F32>(A, b_cur, batches_per_row, w->rec[bp], lane);
            __syncwarp();
            asm volatile("" : "+r"(claim) : : "memory");
            b_cur = b_next;
            b_next = __shfl_sync(PM_FULL_MASK, claim, 0);
            fine_batch_prefetch(A, w, bp, b_next, batches_per_row, n_batches, lane);
            bp ^= 1u;
#else
            uint32_t q = fine_batch_claim(A, lane);
            q = __shfl_sync(PM_FULL_MASK, q, 0);
            if (q >= n_batches) { batches_left = false; continue; }
            fine_solid_batch<F32>(A, q, batches_per_row, nullptr, lane);
#endif
        }
    }
#if PM_FINE_TIMELINE
    if (A.debug && lane == 0) {  // per warp: begin, end of its last tile with records, end, tiles | longest tile (ns << 32 | packed tile)
        unsigned long long *d = A.debug + (size_t)(blockIdx.x * PM_FINE_WARPS + warp) * 24;
        d[0] = tl_begin; d[1] = tl_last; d[2] = fine_now(); d[3] = tl_tiles; d[4] = (tl_long << 32) | tl_long_pk;
        for (int k = 0; k < 16; k++) d[8 + k] = tl_acc[k];
    }
#endif
}

}  // namespace

#define PM_FINE_SMEM (PM_FINE_WARPS * sizeof(FineWarpSmem))

template <bool F32, bool EXACT>
static cudaError_t fine_attr() {
    cudaError_t e = cudaFuncSetAttribute(k_fine<F32, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PM_FINE_SMEM);
    if (e != cudaSuccess) return e;
    // 4 CTAs of ~52 KB per SM: ask for the largest shared-memory carve-out
    return cudaFuncSetAttribute(k_fine<F32, EXACT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

int pm_fine_setup(void) {
    cudaError_t e;
    if ((e = fine_attr<false, false>()) != cudaSuccess) return (int)e;
    if ((e = fine_attr<false, true>()) != cudaSuccess) return (int)e;
    if ((e = fine_attr<true, false>()) != cudaSuccess) return (int)e;
    if ((e = fine_attr<true, true>()) != cudaSuccess) return (int)e;
    return 0;
}

template <bool F32, bool EXACT>
static void fine_launch(const PmFrameArgs &a, int grid, bool overlap, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PM_FINE_WARPS * 32); cfg.dynamicSmemBytes = PM_FINE_SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_fine<F32, EXACT>, a);
}

void pm_launch_fine(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s) {
    // persistent: enough CTAs to fill every SM, work pulled from two queues
    const int grid = sm_count * 4;
    const bool exact = (a.flags & PM_FLAG_EXACT_SRGB) != 0;
    if (a.fb32) {  // debug render with the fp32 parity buffer
        if (exact) fine_launch<true, true>(a, grid, overlap, s);
        else       fine_launch<true, false>(a, grid, overlap, s);
    } else {
        if (exact) fine_launch<false, true>(a, grid, overlap, s);
        else       fine_launch<false, false>(a, grid, overlap, s);
    }
}
