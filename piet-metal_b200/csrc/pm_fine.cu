// k_fine: the fill/blend kernel of the render path (sm_100a) -- the roofline kernel: it stores the framebuffer.
// Compiled with -fmad=false like the rest of the path (pm_tile_logic.h); the places that want an FMA ask for one.
//
// renderKernel's arithmetic (TestApp/PietRender.metal:457-566) evaluated sparsely, one warp per tile, fused with
// the solid-tile composite (metal:16-44): tiles without records are written as 32-tile batches of full 512-byte
// pixel rows.  Tiles with more records than PM_WARP_RECORDS (the 16 inline slots) are left to k_heavy (pm_heavy.cu,
// one CTA per tile), which runs beside this kernel.
//
// Persistent CTAs; a warp pulls work from two queues:
//   * tiles with records: the three lists k_list wrote for this kernel (medium, mid, low: long jobs first, the
//     cheapest last), positions handed out by ticket (see FineList / fine_next).  While tile i is rendered, the
//     header words and the 16 inline record slots (512 B) of tile i+1 are in flight (cp.async into the other half
//     of a double buffer in shared memory), and so is the list entry of tile i+2;
//   * batches of 32 consecutive solid tiles (every store instruction is a full 512-byte run of one pixel row).
// Per tile: the items are taken in painter's order by repeated warp-min over the item ids of the (at most 16)
// records, one record per lane; coverage of one item is accumulated in shared memory in 8.24 fixed point by lanes
// that enumerate (record, pixel row) pairs and then (pair, pixel) units (pm_cover.cuh); lane l owns pixel row
// l / 2, pixels 8 (l & 1) .. +7, resolves their alpha, and blends (packed two-wide FMAs, FFMA2) into the tile's
// linear colour, which goes through a lane-private slice of shared memory only between items; the sRGB encode and
// the two 128-bit framebuffer stores per lane happen once per tile.  An item's linear colour comes from a per-item
// table built at plan time (k_plan), not from the sRGB look-up table.
//
// What the measurements say (profiles/README.md): the kernel is throughput-bound on its instruction mix (issue
// slots ~70 %, L1 data pipe ~58 %; 21 resident warps per SM are only 5 % slower than 28), so what counts is warp
// instructions per tile; its tail is as long as the three tiles a warp owns at a time; claims are atom.inc on
// purpose (an atom.add on a warp-uniform address is compiled into a warp-aggregated atomic plus a shuffle that
// waits for it), and as few as possible (a counter serves one atomic every few cycles).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/piet_metal_b200.h"
#include "pm_cover.cuh"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"

namespace {

typedef unsigned long long u64;

// 4 CTAs of 7 warps per SM: 72 registers per thread, 28 resident warps.  Measured on the 8192^2 tiger (k_fine alone):
// 7 x 4 at 72 registers 103 us, 8 x 3 at 80 registers 106 us, 8 x 4 at 64 registers (a few spills) 111 us.
#ifndef PM_FINE_WARPS
#define PM_FINE_WARPS 7
#endif
#ifndef PM_FINE_CTAS
#define PM_FINE_CTAS 4
#endif
#ifndef PM_FINE_CHUNK
#define PM_FINE_CHUNK 4u         // positions per ticket in the bulk of the list of light tiles (fine_next)
#endif
#ifndef PM_FINE_TAIL_PER_WARP
#define PM_FINE_TAIL_PER_WARP 6u // tiles per warp at the end of the list that are handed out one by one (and as many in twos before them).
                                 // A warp owns up to three tiles at a time (rendering, records in flight, list entry in flight), so the
                                 // kernel's tail is about three tiles long whatever the chunks are; coarse chunks near the end add to it.
                                 // 8192^2 tiger, frame / k_fine alone: 2 -> 139.2 / 98.0 us, 3 -> 137.2 / 96.5, 4 -> 134.5 / 93.6, 6 -> 134.3 / 93.4;
                                 // with the cheapest tiles (one or two records) last, PM_MID_MIN = 3: 6 -> 134.0 / 92.2
#endif
#ifndef PM_FINE_MAGIC_ROUND
#define PM_FINE_MAGIC_ROUND 1    // sRGB bytes rounded with an FADD2 (magic number) instead of cvt.rni.sat.u8 on the XU pipe
#endif
#ifndef PM_FINE_CTA_TICKETS
#define PM_FINE_CTA_TICKETS 1    // the first two tickets of every warp come from one atomic per CTA
#endif
#ifndef PM_FINE_SOLID_EVERY
#define PM_FINE_SOLID_EVERY 4    // one warp in this many prefers the solid batches, the rest the tiles with records
#endif

// Per-warp shared-memory state (6,720 bytes; 7 warps: 46 KB per CTA, four CTAs per SM).
struct FineWarpSmem {
    int acc[256];             // coverage of the item being drawn (pm_cover.cuh)
    int cov[256];
    float4 rgb[3][2][32];     // the tile's linear colour, lane-private: [channel][4-pixel group][lane]
    uint4 rec[3][32];         // [0], [1]: two prefetch buffers, the 16 inline record slots of a tile;
                              // [2]: records 16..31 of a tile that has them (the start of its first overflow block)
    u64 hdr[2][2];            // the prefetched tiles' cnt / occ words
    uint32_t ent[2];          // list entry (packed tile row, column) of the tile that uses buffer b next
    u64 bar[2];               // PM_FINE_BULK: one mbarrier per prefetch buffer (unused otherwise)
    uint32_t pad[2];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- packed fp32 pairs (FFMA2 / FMUL2 on sm_100: two IEEE operations per issue slot, same results as scalar) ----
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// v = mix(v, fg, al) for four pixels of one channel: fma(al, fg, fma(-al, v, v)), exact at al == 0 and al == 1
__device__ __forceinline__ float4 blend4(const float4 v, float fg, u64 al01, u64 al23, u64 nal01, u64 nal23) {
    const u64 v01 = pk2(v.x, v.y), v23 = pk2(v.z, v.w), f = pk2(fg, fg);
    const u64 r01 = fma2(al01, f, fma2(nal01, v01, v01)), r23 = fma2(al23, f, fma2(nal23, v23, v23));
    float4 o;
    upk2(r01, o.x, o.y);
    upk2(r23, o.z, o.w);
    return o;
}

// four values of one channel, linear -> sRGB bytes (default path: the multiplies and the FMA two-wide)
template <bool EXACT>
__device__ __forceinline__ void srgb_bytes4(const float4 v, uint32_t out[4]) {
    if (EXACT) {
        out[0] = pm_srgb_byte<true>(v.x); out[1] = pm_srgb_byte<true>(v.y); out[2] = pm_srgb_byte<true>(v.z); out[3] = pm_srgb_byte<true>(v.w);
        return;
    }
#if PM_FINE_PROBE_NOENC  // (performance probe, wrong pixels: the encode without its six MUFU per pixel)
    out[0] = __float_as_uint(v.x * 255.0f + 12582912.0f); out[1] = __float_as_uint(v.y * 255.0f + 12582912.0f);
    out[2] = __float_as_uint(v.z * 255.0f + 12582912.0f); out[3] = __float_as_uint(v.w * 255.0f + 12582912.0f);
    return;
#endif
    float l0, l1, l2, l3;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(v.x));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(v.y));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(v.z));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l3) : "f"(v.w));
    const u64 g = pk2(1.0f / 2.4f, 1.0f / 2.4f);
    const u64 m01 = mul2(pk2(l0, l1), g), m23 = mul2(pk2(l2, l3), g);
    upk2(m01, l0, l1);
    upk2(m23, l2, l3);
    float p0, p1, p2, p3;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(l0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(l1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p2) : "f"(l2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p3) : "f"(l3));
    const u64 ka = pk2(1.055f * 255.0f, 1.055f * 255.0f), kb = pk2(-0.055f * 255.0f, -0.055f * 255.0f), kl = pk2(12.92f * 255.0f, 12.92f * 255.0f);
    const u64 s01 = fma2(pk2(p0, p1), ka, kb), s23 = fma2(pk2(p2, p3), ka, kb);
    const u64 n01 = mul2(pk2(v.x, v.y), kl), n23 = mul2(pk2(v.z, v.w), kl);
    float s0, s1, s2, s3, n0, n1, n2, n3;
    upk2(s01, s0, s1); upk2(s23, s2, s3);
    upk2(n01, n0, n1); upk2(n23, n2, n3);
    const float r0 = v.x < 0.0031308f ? n0 : s0, r1 = v.y < 0.0031308f ? n1 : s1, r2 = v.z < 0.0031308f ? n2 : s2, r3 = v.w < 0.0031308f ? n3 : s3;
#if PM_FINE_MAGIC_ROUND
    // round to nearest even by adding 1.5 * 2^23: the byte is the low byte of the sum's bit pattern.  The value is
    // within [-0.5, 255.5) (a blend of colours in [0, 1] with weights in [0, 1] stays there up to an ulp), so no clamp;
    // an FADD2 on the FMA pipe instead of four conversions on the XU pipe, which is the busiest unit of this kernel.
    const u64 mg = pk2(12582912.0f, 12582912.0f);
    u64 q01, q23;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q01) : "l"(pk2(r0, r1)), "l"(mg));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q23) : "l"(pk2(r2, r3)), "l"(mg));
    float f0, f1, f2, f3;
    upk2(q01, f0, f1);
    upk2(q23, f2, f3);
    out[0] = __float_as_uint(f0); out[1] = __float_as_uint(f1); out[2] = __float_as_uint(f2); out[3] = __float_as_uint(f3);  // low byte valid
#else
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(out[0]) : "f"(r0));
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(out[1]) : "f"(r1));
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(out[2]) : "f"(r2));
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(out[3]) : "f"(r3));
#endif
}

// pixel = bytes r | g << 8 | b << 16 | 0xff << 24 from the (low bytes of the) three channel words
__device__ __forceinline__ uint32_t pack_rgb(uint32_t r, uint32_t g, uint32_t b) {
#if PM_FINE_MAGIC_ROUND
    return __byte_perm(__byte_perm(r, g, 0x0040), b | 0xff00u, 0x5410);  // (b's word is 0x4b4000bb: byte 1 is zero)
#else
    return r | (g << 8) | (b << 16) | 0xff000000u;
#endif
}

// Cmd_Circle coverage of four consecutive pixels (out of line: rare)
__device__ __noinline__ float4 fine_circle_alpha4(uint32_t bbox_lo, uint32_t bbox_hi, float px0, float py) {
    return make_float4(pm_px_circle_alpha(bbox_lo, bbox_hi, px0, py), pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 1.0f, py),
                       pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 2.0f, py), pm_px_circle_alpha(bbox_lo, bbox_hi, px0 + 3.0f, py));
}

// Starts the copy of a tile's cnt / occ words and inline record slots into buffer b (one commit group).
// PM_FINE_BULK=1 (A/B variant, see profiles/README.md): the 512-byte record block comes as ONE TMA bulk copy
// (cp.async.bulk -> UBLKCP) issued by lane 0 and completes on the buffer's mbarrier instead of 32 lanes' LDGSTS.
#ifndef PM_FINE_BULK
#define PM_FINE_BULK 0
#endif
__device__ __forceinline__ void fine_prefetch(const PmFrameArgs &A, FineWarpSmem *w, uint32_t b, uint32_t entry, uint32_t lane) {
    const size_t tile = (size_t)(entry >> 16) * A.n_tx + (entry & 0xffffu);
#if PM_FINE_BULK
    if (lane == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&w->bar[b]), dst = (uint32_t)__cvta_generic_to_shared(&w->rec[b][0]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last read through the generic proxy
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];"
                     ::"r"(dst), "l"(&A.pool[tile * PM_TILE_SLOTS]), "r"(bar) : "memory");
    }
#else
    cp_async16(&w->rec[b][lane], reinterpret_cast<const uint4 *>(&A.pool[tile * PM_TILE_SLOTS]) + lane);
#endif
    if (lane < 2) cp_async8(&w->hdr[b][lane], lane == 0 ? &A.cnt[tile] : &A.occ[tile]);
    cp_async_commit();
}
// PM_FINE_BULK: waits for the bulk copy into buffer b (phase bit `ph` of its mbarrier)
__device__ __forceinline__ void fine_bulk_wait(FineWarpSmem *w, uint32_t b, uint32_t ph) {
#if PM_FINE_BULK
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&w->bar[b]);
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(ph) : "memory");
    } while (!done);
#else
    (void)w; (void)b; (void)ph;
#endif
}

// Records 16..n-1 of a tile with 17..32 records: the first slots of its first overflow block (pm_pixel_logic.h), copied
// into record buffer 2.  Returns the number of records that can be drawn (16 if the block is missing: the pool ran
// out, the host renders such a frame again).
__device__ __noinline__ uint32_t fine_fetch_ext(const PmFrameArgs &A, FineWarpSmem *w, size_t tile, uint32_t n, uint32_t lane) {
    const u64 vw = A.ovf[tile];
    const uint32_t base1 = (uint32_t)(vw >> 32) == A.stamp ? (uint32_t)vw : 0u;
    if (base1 == 0 || base1 == PM_EXT_FAILED) return PM_TILE_SLOTS;
    if (lane < 2u * (n - PM_TILE_SLOTS)) w->rec[2][lane] = reinterpret_cast<const uint4 *>(&A.pool[base1])[lane];
    __syncwarp();
    return n;
}

// One tile that owns records; its header words and inline records are in buffer b.  All 32 lanes execute this
// together.  Pixel layout: lane l owns pixel row (l >> 1), pixels 8 * (l & 1) .. +7.
template <bool F32, bool EXACT>
__device__ __forceinline__ void fine_tile(const PmFrameArgs &A, FineWarpSmem *w, uint32_t b, uint32_t entry, uint32_t lane) {
    const u64 cw = w->hdr[b][0], ow = w->hdr[b][1];
    uint32_t n = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    if (n > PM_WARP_RECORDS) return;  // (k_heavy's: k_list does not put such a tile on this kernel's lists)
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;
    const uint32_t trow = entry >> 16, tx = entry & 0xffffu;
    if (n > PM_TILE_SLOTS) n = fine_fetch_ext(A, w, (size_t)trow * A.n_tx + tx, n, lane);  // (rare: 1-2 % of the tiles)
    // this lane's record (lanes 16.. : the overflow records): item and kind stay in registers, the geometry is
    // re-read when its item is drawn
    const uint4 *my_rec = &w->rec[lane < PM_TILE_SLOTS ? b : 2u][2 * (lane & (PM_TILE_SLOTS - 1))];
    uint32_t my_item = 0xffffffffu, my_kind = 15u;
    if (lane < n) {
        const uint2 ik = *reinterpret_cast<const uint2 *>(my_rec);
        my_item = ik.x; my_kind = ik.y & 15u;
    }
    if (my_item < occ_item1) my_item = 0xffffffffu;  // below the topmost opaque cover: rewound away (metal:132-135)
    const bool has_draw = __any_sync(PM_FULL_MASK, my_item != 0xffffffffu && my_kind != PM_REC_SOLID);

    const uint32_t prow = lane >> 1, half = lane & 1u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + half * 8u) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) +
                     (tx * PM_TILE_W + half * 8u);

    if (!has_draw) {
        // Only Solid commands after the last rewind: the tile Bails and shows solidColor (metal:145-147, :34-44),
        // which starts as opaque white (metal:74)
        uint32_t c = 0xffffffffu;
        if (occ_item1) c = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));
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

    // the rewound list starts with the cover's Cmd_Solid (metal:136-142, :546-551; the cover is opaque, so the pixel
    // becomes its colour) over white (metal:470).  The load is issued here and consumed after the first item's
    // coverage, where the colour planes are initialised.
    float4 base = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (occ_item1) base = __ldg(&A.item_paint[occ_item1 - 1u]);
    bool fresh = true;
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);  // scene coordinates
    // this lane's two 4-pixel groups of the coverage arrays (word offsets)
    const int my_off0 = pm_cov_swz((int)prow, (int)half * 8), my_off1 = pm_cov_swz((int)prow, (int)half * 8 + 4);
    PmCoverAcc cacc{w->acc, w->cov};

    // items in painter's order: repeatedly take the smallest item id not drawn yet.  The tile's linear colour goes
    // through shared memory only BETWEEN items: the first item blends over the base colour from registers and the
    // last one hands its result straight to the sRGB encode, so a tile with one item never touches the colour planes
    // (the L1 data pipe -- LDS / STS / shuffles -- is the busiest unit of this kernel after the issue slots).
    uint32_t cur_item = __reduce_min_sync(PM_FULL_MASK, my_item);  // (has_draw: there is one)
    for (;;) {
        const bool of_item = my_item == cur_item;
        // the item's closing record says what it is (DrawFill / Stroke / Circle / Solid)
        const uint32_t m_tr = __ballot_sync(PM_FULL_MASK, of_item && my_kind >= PM_REC_CIRCLE);
        const bool mine = of_item && my_kind <= PM_REC_LINE;
        const uint32_t m_geo = __ballot_sync(PM_FULL_MASK, mine);
        if (of_item) my_item = 0xffffffffu;  // done
        const uint32_t next_item = __reduce_min_sync(PM_FULL_MASK, my_item);
        const bool last = next_item == 0xffffffffu;
        uint32_t t_kind = 15u, t_w0 = 0, t_w1 = 0;  // no closing record (cannot happen for a well-formed list): draws nothing
        if (m_tr) {
            const uint32_t tr_lane = __ffs(m_tr) - 1;
            const uint4 tr = w->rec[tr_lane < PM_TILE_SLOTS ? b : 2u][2 * (tr_lane & (PM_TILE_SLOTS - 1))];
            t_kind = tr.y & 15u; t_w0 = tr.z; t_w1 = tr.w;
        }
        float4 paint = make_float4(0.0f, 0.0f, 0.0f, 1.0f);  // Cmd_Circle paints black (metal:491)
        if (t_kind != PM_REC_CIRCLE) paint = __ldg(&A.item_paint[cur_item]);
        const bool stroke = t_kind == PM_REC_STROKE, fill = pm_rec_is_drawfill(t_kind), even_odd = t_kind == PM_REC_DRAWFILL_EO;
        const float half_width = pm_u2f(t_w0);
        int run = 0;  // fill: cover entering this lane's pixels from the left
        const int bd_fx = pm_clamp_i((int)t_w0, -64, 64) << PM_FX_SHIFT;  // the backdrop term of pm_resolve_fill_alpha, once per item
        if (fill || stroke) {
            if (m_geo) {  // coverage of the item's segments
                const uint4 ga = my_rec[0], gb = my_rec[1];
                pm_cover_records(cacc, mine, my_kind, pm_u2f(ga.z), pm_u2f(ga.w), pm_u2f(gb.x), pm_u2f(gb.y), pm_u2f(gb.z), stroke,
                                 half_width + 0.5f, tile_x0, tile_y0, lane);
            }
            __syncwarp();
            if (fill) {  // covers of the left half of the pixel row carry into the right half
                const int4 c0 = *reinterpret_cast<const int4 *>(&w->cov[my_off0]), c1 = *reinterpret_cast<const int4 *>(&w->cov[my_off1]);
                const int sum = ((c0.x + c0.y) + (c0.z + c0.w)) + ((c1.x + c1.y) + (c1.z + c1.w));
                const int other = __shfl_xor_sync(PM_FULL_MASK, sum, 1);
                run = half ? other : 0;
            }
        }
        // resolve this lane's 8 pixels, clear their coverage for the next item, and blend (metal:505, :543, :549).
        // Two rounds of 4 pixels.
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
                    if (!even_odd) {  // (warp-uniform branch: the two rules must not both be evaluated)
                        run += c.x; al[0] = pm_resolve_fill_nz(a.x + run, bd_fx);
                        run += c.y; al[1] = pm_resolve_fill_nz(a.y + run, bd_fx);
                        run += c.z; al[2] = pm_resolve_fill_nz(a.z + run, bd_fx);
                        run += c.w; al[3] = pm_resolve_fill_nz(a.w + run, bd_fx);
                    } else {
                        const int backdrop = (int)t_w0;
                        run += c.x; al[0] = pm_resolve_fill_alpha_eo(a.x + run, backdrop);
                        run += c.y; al[1] = pm_resolve_fill_alpha_eo(a.y + run, backdrop);
                        run += c.z; al[2] = pm_resolve_fill_alpha_eo(a.z + run, backdrop);
                        run += c.w; al[3] = pm_resolve_fill_alpha_eo(a.w + run, backdrop);
                    }
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
            } else {  // PM_REC_SOLID: a translucent full cover (15: nothing)
                al[0] = al[1] = al[2] = al[3] = t_kind == PM_REC_SOLID ? 1.0f : 0.0f;
            }
            const u64 pa2 = pk2(paint.w, paint.w);
            const u64 al01 = mul2(pk2(al[0], al[1]), pa2), al23 = mul2(pk2(al[2], al[3]), pa2);
            float a0, a1, a2, a3;
            upk2(al01, a0, a1);
            upk2(al23, a2, a3);
            const u64 nal01 = pk2(-a0, -a1), nal23 = pk2(-a2, -a3);
            float4 vr, vg, vb;
            if (fresh) {
                vr = make_float4(base.x, base.x, base.x, base.x); vg = make_float4(base.y, base.y, base.y, base.y); vb = make_float4(base.z, base.z, base.z, base.z);
            } else {
                vr = w->rgb[0][g][lane]; vg = w->rgb[1][g][lane]; vb = w->rgb[2][g][lane];
            }
            vr = blend4(vr, paint.x, al01, al23, nal01, nal23);
            vg = blend4(vg, paint.y, al01, al23, nal01, nal23);
            vb = blend4(vb, paint.z, al01, al23, nal01, nal23);
            if (!last) {
                w->rgb[0][g][lane] = vr; w->rgb[1][g][lane] = vg; w->rgb[2][g][lane] = vb;
            } else {
                uint32_t rb[4], gb[4], bb[4];
                srgb_bytes4<EXACT>(vr, rb);
                srgb_bytes4<EXACT>(vg, gb);
                srgb_bytes4<EXACT>(vb, bb);
                const uint4 px = make_uint4(pack_rgb(rb[0], gb[0], bb[0]), pack_rgb(rb[1], gb[1], bb[1]), pack_rgb(rb[2], gb[2], bb[2]), pack_rgb(rb[3], gb[3], bb[3]));
                __stcs(reinterpret_cast<uint4 *>(dst) + g, px);
                if (F32) {  // debug render: the un-quantised values
                    dst32[4 * g + 0] = make_float4(pm_linear_to_srgb<EXACT>(vr.x), pm_linear_to_srgb<EXACT>(vg.x), pm_linear_to_srgb<EXACT>(vb.x), 1.0f);
                    dst32[4 * g + 1] = make_float4(pm_linear_to_srgb<EXACT>(vr.y), pm_linear_to_srgb<EXACT>(vg.y), pm_linear_to_srgb<EXACT>(vb.y), 1.0f);
                    dst32[4 * g + 2] = make_float4(pm_linear_to_srgb<EXACT>(vr.z), pm_linear_to_srgb<EXACT>(vg.z), pm_linear_to_srgb<EXACT>(vb.z), 1.0f);
                    dst32[4 * g + 3] = make_float4(pm_linear_to_srgb<EXACT>(vr.w), pm_linear_to_srgb<EXACT>(vg.w), pm_linear_to_srgb<EXACT>(vb.w), 1.0f);
                }
            }
        }
        __syncwarp();
        if (last) break;
        fresh = false;
        cur_item = next_item;
    }
}

// 32 consecutive tiles of one tile row; the solid ones are written row-wise: each store instruction covers
// 512 contiguous bytes (128 pixels) of one pixel row.
template <bool F32>
__device__ __forceinline__ void fine_solid_batch(const PmFrameArgs &A, uint32_t batch, uint32_t batches_per_row, uint32_t lane) {
    const uint32_t row = batch / batches_per_row;
    const uint32_t t0 = (batch - row * batches_per_row) * 32u;
    const uint32_t t = t0 + lane;
    bool solid = false;
    uint32_t colour = 0xffffffffu;  // solidColor starts as opaque white (metal:74)
    if (t < A.n_tx) {
        const size_t tile = (size_t)row * A.n_tx + t;
        const u64 cw = A.cnt[tile], ow = A.occ[tile];
        solid = !((uint32_t)(cw >> 32) == A.stamp && (uint32_t)cw != 0u);
        if (solid && (uint32_t)(ow >> 32) == A.stamp && (uint32_t)ow != 0u)
            colour = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)((uint32_t)ow - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));
    }
    const uint32_t solid_mask = __ballot_sync(PM_FULL_MASK, solid);
    if (solid_mask == 0) return;
    uint8_t *base = A.fb + (size_t)(row * PM_TILE_H) * A.pitch + ((size_t)t0 * PM_TILE_W + lane * 4u) * 4u;
    #pragma unroll 1
    for (int q = 0; q < 4; q++) {  // a quarter of the batch: 8 tiles = 128 pixels per row, 4 per lane
        const uint32_t src = (uint32_t)q * 8u + (lane >> 2);
        const uint32_t c = __shfl_sync(PM_FULL_MASK, colour, src);
        if ((solid_mask >> src) & 1u) {
            const uint4 v = make_uint4(c, c, c, c);
            uint8_t *dst = base + (size_t)q * 512u;
            #pragma unroll
            for (int y = 0; y < PM_TILE_H; y++) __stcs(reinterpret_cast<uint4 *>(dst + (size_t)y * A.pitch), v);  // streaming: keep L2 for the records
            if (F32) {
                const float4 f = make_float4((float)(c & 0xff) / 255.0f, (float)((c >> 8) & 0xff) / 255.0f,
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
}

// Work list of the tiles with records, as k_fine walks it: k_list's three classes one after the other -- medium
// (PM_MEDIUM_MIN .. PM_WARP_RECORDS records: the long jobs), mid, low (one or two records: the cheapest, last).
// Three tiles are in a warp's pipeline: the tile being rendered (buffer b), the next one (its header words and
// records in flight into buffer b ^ 1) and the one after it (its list entry in flight into ent[b]); only a ticket
// (one L2 atomic per run of positions) is waited for.  Dealing a part of the list statically was measured and lost:
// the CTAs that start late -- k_heavy's CTA holds one of the four slots of an SM -- keep their share waiting.
struct FineList {
    const uint32_t *medium, *mid, *low;  // the three classes k_list wrote for this kernel, walked in this order
    uint32_t n_medium, n_mid, n_total;
    uint32_t n4, n2;  // tickets that stand for 4 / 2 consecutive positions (see fine_next)
};
// (atom.inc with a bound that is never reached, not atom.add: see the header)
__device__ __forceinline__ uint32_t fine_claim(const PmFrameArgs &A, uint32_t lane) {
    uint32_t c = 0;
    if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(c) : "l"(&A.queue->tile_next) : "memory");
    return c;
}
// Positions are handed out by TICKET, and a ticket stands for a run of consecutive positions: one position each
// for the medium tiles at the head of the list (the long jobs), then PM_FINE_CHUNK positions per ticket for most of
// the light tiles, then 2, then 1 again for the tail (guided self-scheduling: the kernel still ends evenly, and the
// counter sees a quarter of the atomics -- with one atomic per tile ~4,000 warps queue on one L2 address and a
// claim took microseconds: a fifth of all warp stall samples in the round-2 profile).
// `pre`, `n_pre`: tickets this warp already owns -- the first two of every warp come out of ONE atomic per CTA at the
// start of the kernel (4,000 warps asking the counter twice each at the same moment kept the last of them waiting
// for 10-20 us: an L2 slice serves one atomic on one address every few cycles).
struct FineRun { uint32_t pos, end, pre, n_pre, pre_stride; };
__device__ __forceinline__ uint32_t fine_next(const PmFrameArgs &A, const FineList &L, FineRun &r, uint32_t lane) {
    if (r.pos < r.end) return r.pos++;
    uint32_t t, start, len = 1;
    if (r.n_pre) { t = r.pre; r.pre += r.pre_stride; r.n_pre--; }
    else t = __shfl_sync(PM_FULL_MASK, fine_claim(A, lane), 0);
    if (t < L.n_medium) start = t;
    else if ((t -= L.n_medium) < L.n4) { start = L.n_medium + PM_FINE_CHUNK * t; len = PM_FINE_CHUNK; }
    else if ((t -= L.n4) < L.n2) { start = L.n_medium + PM_FINE_CHUNK * L.n4 + 2u * t; len = 2; }
    else start = L.n_medium + PM_FINE_CHUNK * L.n4 + 2u * L.n2 + (t - L.n2);
    if (start > L.n_total) start = L.n_total;  // (also keeps the sums below from wrapping after many empty claims)
    r.pos = start + 1u;
    r.end = start + len < L.n_total ? start + len : L.n_total;
    return start;
}
__device__ __forceinline__ void fine_fetch_entry(const FineList &L, FineWarpSmem *w, uint32_t b, uint32_t pos, uint32_t lane) {
    if (lane == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(&w->ent[b]);
        const uint32_t *src = pos < L.n_medium ? &L.medium[pos] : (pos - L.n_medium < L.n_mid ? &L.mid[pos - L.n_medium] : &L.low[pos - L.n_medium - L.n_mid]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(src) : "memory");
    }
    cp_async_commit();
}

#ifndef PM_FINE_TIMELINE
#define PM_FINE_TIMELINE 0
#endif
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool F32, bool EXACT>
__global__ void __launch_bounds__(PM_FINE_WARPS * 32, PM_FINE_CTAS) k_fine(const PmFrameArgs A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FineWarpSmem *w = reinterpret_cast<FineWarpSmem *>(s_raw) + warp;
#if PM_FINE_TIMELINE  // debug build: per-warp [kernel entry, first tile, last tile end, exit], longest tile, tiles, time in tiles
    const unsigned long long tl_start = tl_now();
    unsigned long long tl_first = 0, tl_last = 0, tl_long = 0, tl_tiles = 0, tl_sum = 0, tl_long_start = 1, tl_last_start = 0, tl_last_entry = 0;
#endif
    __shared__ uint32_t s_first_ticket;
    for (uint32_t i = lane; i < 256; i += 32) { w->acc[i] = 0; w->cov[i] = 0; }
#if PM_FINE_BULK
    if (lane < 2) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&w->bar[lane])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t bar_phase = 0;  // bit b: the phase the next wait on buffer b expects
#endif
    // Programmatic dependent launch (pm_kernels.cu): this grid is released by k_heavy, whose CTAs signal only after
    // binning has completed -- so there is no wait here; the wait at the END of the kernel makes this grid's
    // completion imply k_heavy's, which is what the next frame's k_seg depends on.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t n_medium = A.counters->n_medium, n_mid = A.counters->n_mid, n_light = n_mid + A.counters->n_low;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        A.report->n_complex = A.counters->n_heavy + n_medium + n_light;
        A.report->n_overflow = A.counters->n_overflow;
        A.report->n_heavy = A.counters->n_heavy;
        A.report->frame = A.stamp;
    }
    const size_t n_tiles = (size_t)A.n_rows * A.n_tx;
    const bool prefer_complex = (warp % PM_FINE_SOLID_EVERY) != (PM_FINE_SOLID_EVERY - 1);
    constexpr uint32_t kComplexWarps = PM_FINE_WARPS - PM_FINE_WARPS / PM_FINE_SOLID_EVERY;
    // the first two tickets of every warp that starts with the tiles with records: one atomic for the whole CTA
#if PM_FINE_CTA_TICKETS
    // (two tickets per warp when the list is long; one when it holds only a few tiles per warp -- a narrow strip --
    // so that they spread over all the warps)
    const uint32_t n_pre = n_medium + n_light >= 4u * gridDim.x * kComplexWarps ? 2u : 1u;
    if (threadIdx.x == 0) s_first_ticket = atomicAdd(&A.queue->tile_next, n_pre * kComplexWarps);
    __syncthreads();
    FineRun run{0u, 0u, s_first_ticket + (warp - (warp + 1u) / PM_FINE_SOLID_EVERY), prefer_complex ? n_pre : 0u, kComplexWarps};
#else
    FineRun run{0u, 0u, 0u, 0u, 0u};
#endif
    FineList L;
    L.medium = A.complex_list + 2 * n_tiles;
    L.mid = A.complex_list + 3 * n_tiles;
    L.low = A.complex_list;
    L.n_medium = n_medium;
    L.n_mid = n_mid;
    L.n_total = n_medium + n_light;
    {   // guided: the last PM_FINE_TAIL_PER_WARP tiles per warp of the list go out one by one, as many before them in
        // twos, the rest (the bulk of a large frame) PM_FINE_CHUNK at a time; a narrow multi-GPU strip with about one
        // tile per warp is handed out tile by tile
        const uint32_t per = gridDim.x * kComplexWarps * PM_FINE_TAIL_PER_WARP;
        const uint32_t singles = n_light < per ? n_light : per;
        const uint32_t twos = n_light - singles < per ? n_light - singles : per;
        L.n4 = (n_light - singles - twos) / PM_FINE_CHUNK;
        L.n2 = twos / 2u;
    }
    __syncwarp();
    const uint32_t batches_per_row = (A.n_tx + 31u) / 32u;
    const uint32_t n_batches = batches_per_row * A.n_rows;
    bool complex_left = true, batches_left = true;
    while (complex_left || batches_left) {
        if (complex_left && (prefer_complex || !batches_left)) {
            complex_left = false;
            // fill the pipeline
            uint32_t p = fine_next(A, L, run, lane);
            if (p >= L.n_total) continue;
            fine_fetch_entry(L, w, 0, p, lane);
            p = fine_next(A, L, run, lane);
            cp_async_wait<0>();
            __syncwarp();
            fine_prefetch(A, w, 0, w->ent[0], lane);
            bool v_next = p < L.n_total;
            if (v_next) {
                fine_fetch_entry(L, w, 1, p, lane);
            }
            uint32_t b = 0;
            for (;;) {
                cp_async_wait<0>();
#if PM_FINE_BULK
                fine_bulk_wait(w, b, (bar_phase >> b) & 1u);
                bar_phase ^= 1u << b;
#endif
                __syncwarp();
                const uint32_t entry = w->ent[b];
                bool v_nn = false;
                if (v_next) {
                    fine_prefetch(A, w, b ^ 1u, w->ent[b ^ 1u], lane);
                    __syncwarp();  // (every lane has read ent[b] before it is overwritten)
                    p = fine_next(A, L, run, lane);
                    v_nn = p < L.n_total;
                    if (v_nn) {
                        fine_fetch_entry(L, w, b, p, lane);
                    }
                }
#if PM_FINE_TIMELINE
                const unsigned long long tl_a = tl_now();
#endif
                fine_tile<F32, EXACT>(A, w, b, entry, lane);
                __syncwarp();
#if PM_FINE_TIMELINE
                {
                    const unsigned long long tl_b = tl_now(), d = tl_b - tl_a;
                    if (tl_first == 0) tl_first = tl_a;
                    tl_last = tl_b; tl_tiles++; tl_sum += d; tl_last_start = tl_a; tl_last_entry = entry;
                    if (d > (tl_long >> 32)) { tl_long = (d << 32) | entry; tl_long_start = tl_a; }
                }
#endif
                if (!v_next) break;
                b ^= 1u;
                v_next = v_nn;
            }
        } else {
            uint32_t q = 0;
            if (lane == 0) asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(q) : "l"(&A.queue->batch_next) : "memory");
            q = __shfl_sync(PM_FULL_MASK, q, 0);
            if (q >= n_batches) { batches_left = false; continue; }
            fine_solid_batch<F32>(A, q, batches_per_row, lane);
        }
    }
#if PM_FINE_TIMELINE
    if (A.debug && lane == 0) {
        unsigned long long *d = A.debug + (1u << 19) + 8ull * (blockIdx.x * PM_FINE_WARPS + warp);
        d[0] = tl_start; d[1] = tl_first; d[2] = tl_last; d[3] = tl_now(); d[4] = tl_long; d[5] = tl_tiles | (tl_last_entry << 32); d[6] = tl_last_start; d[7] = tl_long_start;
    }
#endif
    asm volatile("griddepcontrol.wait;" ::: "memory");  // k_heavy (and everything before it) has completed
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // the counters the next frame will use (nobody reads this frame's any more)
        A.counters_next->n_complex = 0;
        A.counters_next->n_overflow = 0;
        A.counters_next->n_heavy = 0;
        A.counters_next->n_medium = 0;
        A.counters_next->n_mid = 0;
        A.counters_next->n_low = 0;
    }
}

}  // namespace

#define PM_FINE_SMEM (PM_FINE_WARPS * sizeof(FineWarpSmem))

template <bool F32, bool EXACT>
static cudaError_t fine_attr() {
    cudaError_t e = cudaFuncSetAttribute(k_fine<F32, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PM_FINE_SMEM);
    if (e != cudaSuccess) return e;
    // four CTAs of ~49 KB per SM: ask for the largest shared-memory carve-out
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
static cudaError_t fine_launch(const PmFrameArgs &a, int grid, bool overlap, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PM_FINE_WARPS * 32); cfg.dynamicSmemBytes = PM_FINE_SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_fine<F32, EXACT>, a);
}

cudaError_t pm_launch_fine(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s) {
    // persistent: enough CTAs to fill every SM, work pulled from two queues
    int grid = sm_count * PM_FINE_CTAS;
    {   // experiment switch: fewer resident CTAs per SM with the same code (is the kernel latency- or throughput-bound?)
        static const char *e = getenv("PM_DEBUG_FINE_CTAS");
        if (e && atoi(e) > 0) grid = sm_count * atoi(e);
    }
    const bool exact = (a.flags & PM_FLAG_EXACT_SRGB) != 0;
    if (a.fb32) {  // debug render with the fp32 parity buffer
        return exact ? fine_launch<true, true>(a, grid, overlap, s) : fine_launch<true, false>(a, grid, overlap, s);
    }
    return exact ? fine_launch<false, true>(a, grid, overlap, s) : fine_launch<false, false>(a, grid, overlap, s);
}
