// k_heavy: fill/blend of the tiles with more records than k_fine takes (PM_HEAVY_MIN = 17: more than the inline
// slots; sm_100a).  Same arithmetic as k_fine (renderKernel, TestApp/PietRender.metal:457-566; pm_cover.cuh,
// pm_pixel_logic.h), other decompositions: such a tile has tens to thousands of records -- coincident outlines, deep
// stacks of translucent layers, a whole drawing squeezed into a few tiles.  Rendered the way k_fine renders a light
// tile, one of the tiger's worst tiles at 8192^2 keeps a warp busy for 60-110 us (a warp issues ~0.1 instructions per
// cycle) and the whole 256^2 tiger takes milliseconds: that is the critical path of a narrow multi-GPU strip.
//
// Two modes, chosen per frame from the number of heavy tiles:
//   * few heavy tiles (latency matters): one CTA of 8 warps per tile.
//       1. the tile's records (inline slots + the chain of overflow blocks, pm_pixel_logic.h) are keyed
//          (item, trailer first, position) and sorted in shared memory (bitonic): the records of an item become a
//          contiguous run, the runs are in painter's order;
//       2. the items are taken eight at a time: warp w accumulates the coverage of item 8 g + w into its own
//          coverage arrays and resolves it to a per-pixel alpha (compositing is ordered, coverage is not);
//       3. thread t owns pixel (t / 16, t % 16) with its linear colour in three registers and blends the eight
//          layers in order: one load and two FMAs per channel and layer; encode and store once per tile.
//     A tile with more than PM_HEAVY_SORT_CAP records (a 16x16-pixel tile crossed by > 4096 segments) is drawn
//     without the sort: one pass over all its records per item (correct, slow, only seen in stress tests).
//   * many heavy tiles (throughput matters: the 10 k Bezier and 100 k glyph scenes have 40 k+ of them): one WARP
//     per tile with up to PM_HEAVY_WARP_CAP records (pm_heavy_warp.cuh); the few larger tiles are then drawn
//     CTA-wise as above.
// (Sixteen warps per CTA were measured: a heavy tile is not faster -- its time is sort, loads and the slowest
// item -- and two such CTAs take all of an SM's registers away from k_fine.)
// Integer coverage sums and the same per-pixel functions everywhere: which kernel or mode draws a tile does not
// change a single bit of its pixels.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/piet_metal_b200.h"
#include "pm_cover.cuh"
#include "pm_heavy_warp.cuh"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"

namespace {

typedef unsigned long long u64;

#define PM_HEAVY_WARPS 8
#ifndef PM_HEAVY_GRID_PER_SM
#define PM_HEAVY_GRID_PER_SM 1   // CTAs launched per SM.  One CTA of 8 warps at 64 registers leaves three of k_fine's four CTA slots free
#endif                           // (measured on the 8192^2 tiger, frame: 2 per SM at 80 registers 160.9 us, 1 per SM at 64 registers 152.5 us)
#define PM_HEAVY_THREADS (PM_HEAVY_WARPS * 32)
#define PM_HEAVY_SORT_CAP 4096u
#define PM_HEAVY_DIR_CAP 128u     // overflow blocks indexed per tile (CTA mode): 1488 + 123 * 768 slots ~ 96 k records; the rest is not drawn

struct HeavyWarpState {          // warp mode, per warp (shares its memory with the CTA mode's sort arrays)
    float4 rgb[3][2][32];        // the tile's linear colour, lane-private (as in k_fine)
    uint32_t dir[PM_HEAVY_WARP_DIR];
    uint32_t pad[12];
};

struct HeavyMeta { float4 paint; uint32_t valid, pad[3]; };

struct HeavySmem {
    int acc[PM_HEAVY_WARPS][256];
    int cov[PM_HEAVY_WARPS][256];
    union {
        struct {
            u64 keys[PM_HEAVY_SORT_CAP];           // CTA mode: (item << 32) | (geometry ? 1 << 31 : 0) | position; dropped records: all ones
            uint16_t starts[PM_HEAVY_SORT_CAP + 8];// sorted position of every item's first record, then the number of live records
        };
        HeavyWarpState ws[PM_HEAVY_WARPS];         // warp mode
    };
    uint32_t dir[PM_HEAVY_DIR_CAP];            // 1 + pool index of the header of overflow block j
    HeavyMeta meta[PM_HEAVY_WARPS];
    u64 cw, ow, vw;
    uint32_t red[PM_HEAVY_WARPS];
    uint32_t reach, n_live, n_items, tile_next;
};

__device__ __forceinline__ uint32_t block_min(uint32_t v, HeavySmem *sh, uint32_t lane, uint32_t warp) {
    v = __reduce_min_sync(PM_FULL_MASK, v);
    if (lane == 0) sh->red[warp] = v;
    __syncthreads();
    uint32_t m = sh->red[0];
    #pragma unroll
    for (int k = 1; k < PM_HEAVY_WARPS; k++) m = sh->red[k] < m ? sh->red[k] : m;
    __syncthreads();
    return m;
}

// ---------------------------------------------------------------------------------------------------------------
// CTA mode
// ---------------------------------------------------------------------------------------------------------------

// The layer a warp has prepared: alpha * paint.a of every pixel goes into the warp's acc array as float bits.
__device__ __forceinline__ void heavy_publish_layer(HeavySmem *sh, uint32_t warp, uint32_t kind, uint32_t w0, uint32_t w1, float4 paint,
                                                    float tile_x0, float tile_y0, uint32_t lane) {
    float al[8];
    pm_heavy_resolve8(sh->acc[warp], sh->cov[warp], kind, w0, w1, tile_x0, tile_y0, lane, al);
    const uint32_t prow = lane >> 1, half = lane & 1u;
    const int off0 = pm_cov_swz((int)prow, (int)half * 8), off1 = pm_cov_swz((int)prow, (int)half * 8 + 4);
    *reinterpret_cast<float4 *>(&sh->acc[warp][off0]) = make_float4(al[0] * paint.w, al[1] * paint.w, al[2] * paint.w, al[3] * paint.w);
    *reinterpret_cast<float4 *>(&sh->acc[warp][off1]) = make_float4(al[4] * paint.w, al[5] * paint.w, al[6] * paint.w, al[7] * paint.w);
    if (lane == 0) { sh->meta[warp].paint = paint; sh->meta[warp].valid = 1u; }
}

template <bool F32, bool EXACT>
__device__ void heavy_tile_cta(const PmFrameArgs &A, HeavySmem *sh, uint32_t entry, uint32_t n_min) {
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t trow = entry >> 16, tx = entry & 0xffffu;
    const size_t tile = (size_t)trow * A.n_tx + tx;
    if (t == 0) {
        const u64 cw = A.cnt[tile], vw = A.ovf[tile];
        sh->cw = cw; sh->ow = A.occ[tile]; sh->vw = vw; sh->n_live = 0;
        const uint32_t n0 = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
        sh->reach = (n0 > PM_TILE_SLOTS && n0 >= n_min) ? pm_heavy_walk(A, vw, n0, sh->dir, PM_HEAVY_DIR_CAP) : n0;
    }
    __syncthreads();
    const u64 cw = sh->cw, ow = sh->ow;
    const uint32_t n_cnt = (uint32_t)(cw >> 32) == A.stamp ? (uint32_t)cw : 0u;
    const uint32_t n = sh->reach;
    const uint32_t occ_item1 = (uint32_t)(ow >> 32) == A.stamp ? (uint32_t)ow : 0u;
    __syncthreads();
    if (n_cnt < n_min) return;  // (drawn warp-wise in the pass before this one)

    const bool owner = t < 256;  // owns pixel (t / 16, t % 16)
    const uint32_t prow = (t >> 4) & 15u, px = t & 15u;
    uint8_t *dst = A.fb + (size_t)(trow * PM_TILE_H + prow) * A.pitch + (size_t)(tx * PM_TILE_W + px) * 4u;
    float4 *dst32 = nullptr;
    if (F32) dst32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(A.fb32) + (size_t)(trow * PM_TILE_H + prow) * A.pitch32) + (tx * PM_TILE_W + px);
    const float tile_x0 = (float)(tx * PM_TILE_W), tile_y0 = (float)((A.tile_y0 + trow) * PM_TILE_H);
    const int cell = pm_cov_swz((int)prow, (int)px);
    // metal:470 white; then the cover's Cmd_Solid (metal:136-142, :546-551: opaque, the pixel becomes its colour)
    float c0 = 1.0f, c1 = 1.0f, c2 = 1.0f;
    if (occ_item1) { const float4 b = __ldg(&A.item_paint[occ_item1 - 1u]); c0 = b.x; c1 = b.y; c2 = b.z; }
    int has_draw = 0;
    PmCoverAcc cacc{sh->acc[warp], sh->cov[warp]};

    if (n <= PM_HEAVY_SORT_CAP) {
        // ---- 1. key and sort the records ----
        uint32_t P = 32;
        while (P < n) P <<= 1;
        uint32_t live = 0;
        for (uint32_t p = t; p < P; p += PM_HEAVY_THREADS) {
            u64 key = ~0ull;
            if (p < n) {
                const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[pm_heavy_index(sh->dir, tile, p)]);
                if (ik.x >= occ_item1) {  // (below the topmost opaque cover: rewound away, metal:132-135)
                    const uint32_t kind = ik.y & 15u;
                    key = ((u64)ik.x << 32) | (kind >= PM_REC_CIRCLE ? 0u : 0x80000000u) | p;
                    live++;
                    if (kind != PM_REC_SOLID) has_draw = 1;
                }
            }
            sh->keys[p] = key;
        }
        if (live) atomicAdd(&sh->n_live, live);
        has_draw = __syncthreads_or(has_draw);
        if (has_draw) {
            for (uint32_t k = 2; k <= P; k <<= 1)
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    for (uint32_t i = t; i < P; i += PM_HEAVY_THREADS) {
                        const uint32_t ixj = i ^ j;
                        if (ixj > i) {
                            const u64 a = sh->keys[i], b = sh->keys[ixj];
                            if ((a > b) == ((i & k) == 0)) { sh->keys[i] = b; sh->keys[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            // ---- item starts (compaction of the positions where the item id changes) ----
            const uint32_t nl = sh->n_live;
            const uint32_t per = (nl + PM_HEAVY_THREADS - 1) / PM_HEAVY_THREADS;
            const uint32_t lo = t * per < nl ? t * per : nl, hi = lo + per < nl ? lo + per : nl;
            uint32_t cnt = 0;
            for (uint32_t i = lo; i < hi; i++) cnt += (i == 0 || (uint32_t)(sh->keys[i] >> 32) != (uint32_t)(sh->keys[i - 1] >> 32)) ? 1u : 0u;
            uint32_t incl = cnt;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(PM_FULL_MASK, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            if (lane == 31) sh->red[warp] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (uint32_t k = 0; k < warp; k++) before += sh->red[k];
            uint32_t at = before + incl - cnt;
            for (uint32_t i = lo; i < hi; i++)
                if (i == 0 || (uint32_t)(sh->keys[i] >> 32) != (uint32_t)(sh->keys[i - 1] >> 32)) sh->starts[at++] = (uint16_t)i;
            if (t == PM_HEAVY_THREADS - 1) { sh->n_items = before + incl; sh->starts[before + incl] = (uint16_t)nl; }
            __syncthreads();
            const uint32_t n_items = sh->n_items;

            // ---- 2 + 3. sixteen items at a time: coverage and alpha per warp, then the layers blended in order ----
            for (uint32_t k0 = 0; k0 < n_items; k0 += PM_HEAVY_WARPS) {
                const uint32_t k = k0 + warp;
                if (lane == 0) sh->meta[warp].valid = 0u;
                __syncwarp();
                if (k < n_items) {
                    const uint32_t s = sh->starts[k], e = sh->starts[k + 1];
                    const u64 first = sh->keys[s];
                    if (!(first & 0x80000000ull)) {  // the item's closing record sorts first (an item without one is not drawn)
                        const uint4 tr = *reinterpret_cast<const uint4 *>(&A.pool[pm_heavy_index(sh->dir, tile, (uint32_t)first & 0x7fffffffu)]);
                        const uint32_t kind = tr.y & 15u;
                        float4 paint = make_float4(0.0f, 0.0f, 0.0f, 1.0f);  // Cmd_Circle paints black (metal:491)
                        if (kind != PM_REC_CIRCLE) paint = __ldg(&A.item_paint[(uint32_t)(first >> 32)]);
                        const bool stroke = kind == PM_REC_STROKE;
                        if (stroke || pm_rec_is_drawfill(kind)) {
                            const float reach = pm_u2f(tr.z) + 0.5f;
                            for (uint32_t c = s + 1; c < e; c += 32) {
                                const bool mine = c + lane < e;
                                PmRecord rc;
                                rc.key = 0; rc.p[0] = rc.p[1] = rc.p[2] = rc.p[3] = 0.0f; rc.edge_y = 0.0f;
                                if (mine) rc = pm_load_record(A.pool, pm_heavy_index(sh->dir, tile, (uint32_t)sh->keys[c + lane] & 0x7fffffffu));
                                pm_cover_records(cacc, mine, rc.key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, reach, tile_x0, tile_y0, lane);
                            }
                            __syncwarp();
                        }
                        heavy_publish_layer(sh, warp, kind, tr.z, tr.w, paint, tile_x0, tile_y0, lane);
                    }
                }
                __syncthreads();
                const uint32_t n_here = n_items - k0 < PM_HEAVY_WARPS ? n_items - k0 : PM_HEAVY_WARPS;
                if (owner) {
                    for (uint32_t s = 0; s < n_here; s++) {
                        if (!sh->meta[s].valid) continue;
                        const float4 paint = sh->meta[s].paint;
                        const float al = __int_as_float(sh->acc[s][cell]);
                        sh->acc[s][cell] = 0;
                        c0 = pm_mix_fma(c0, paint.x, al); c1 = pm_mix_fma(c1, paint.y, al); c2 = pm_mix_fma(c2, paint.z, al);
                    }
                }
                __syncthreads();
            }
        }
    } else {
        // ---- more records than the sort holds: one pass over all records per item ----
        for (uint32_t p0 = 0; p0 < n; p0 += PM_HEAVY_THREADS) {
            const uint32_t p = p0 + t;
            if (p < n) {
                const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[pm_heavy_index(sh->dir, tile, p)]);
                if (ik.x >= occ_item1 && (ik.y & 15u) != PM_REC_SOLID) has_draw = 1;
            }
        }
        has_draw = __syncthreads_or(has_draw);
        PmCoverAcc call{sh->acc[0], sh->cov[0]};  // all warps add to one set of arrays
        uint32_t lo_item = occ_item1;
        while (has_draw) {
            uint32_t cand = 0xffffffffu;
            for (uint32_t p = t; p < n; p += PM_HEAVY_THREADS) {
                const uint32_t it = A.pool[pm_heavy_index(sh->dir, tile, p)].item;
                if (it >= lo_item && it < cand) cand = it;
            }
            const uint32_t cur = block_min(cand, sh, lane, warp);
            if (cur == 0xffffffffu) break;
            lo_item = cur + 1u;
            if (t == 0) sh->red[0] = 0u;  // (kind of the closing record)
            __syncthreads();
            for (uint32_t p = t; p < n; p += PM_HEAVY_THREADS) {
                const uint4 a = *reinterpret_cast<const uint4 *>(&A.pool[pm_heavy_index(sh->dir, tile, p)]);
                if (a.x == cur && (a.y & 15u) >= PM_REC_CIRCLE) { sh->red[0] = a.y & 15u; sh->red[1] = a.z; sh->red[2] = a.w; }
            }
            __syncthreads();
            const uint32_t kind = sh->red[0], w0 = sh->red[1], w1 = sh->red[2];
            __syncthreads();
            if (kind == 0) continue;  // (uniform: every thread reads the same words)
            const bool stroke = kind == PM_REC_STROKE;
            if (stroke || pm_rec_is_drawfill(kind)) {
                const float reach = pm_u2f(w0) + 0.5f;
                for (uint32_t p0 = 0; p0 < n; p0 += PM_HEAVY_THREADS) {
                    const uint32_t p = p0 + t;
                    bool mine = false;
                    PmRecord rc;
                    rc.key = 0; rc.p[0] = rc.p[1] = rc.p[2] = rc.p[3] = 0.0f; rc.edge_y = 0.0f;
                    if (p < n) {
                        const uint32_t idx = pm_heavy_index(sh->dir, tile, p);
                        const uint2 ik = *reinterpret_cast<const uint2 *>(&A.pool[idx]);
                        if (ik.x == cur && (ik.y & 15u) <= PM_REC_LINE) { mine = true; rc = pm_load_record(A.pool, idx); }
                    }
                    if (__any_sync(PM_FULL_MASK, mine))
                        pm_cover_records(call, mine, rc.key & 15u, rc.p[0], rc.p[1], rc.p[2], rc.p[3], rc.edge_y, stroke, reach, tile_x0, tile_y0, lane);
                }
            }
            __syncthreads();
            if (warp == 0) {
                const float4 paint = kind != PM_REC_CIRCLE ? __ldg(&A.item_paint[cur]) : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
                heavy_publish_layer(sh, 0, kind, w0, w1, paint, tile_x0, tile_y0, lane);
            }
            __syncthreads();
            if (owner) {
                const float4 paint = sh->meta[0].paint;
                const float al = __int_as_float(sh->acc[0][cell]);
                sh->acc[0][cell] = 0;
                c0 = pm_mix_fma(c0, paint.x, al); c1 = pm_mix_fma(c1, paint.y, al); c2 = pm_mix_fma(c2, paint.z, al);
            }
            __syncthreads();
        }
    }

    if (owner) {
        if (!has_draw) {  // only Solid commands after the last rewind: the tile Bails and shows solidColor (metal:145-147, :34-44)
            uint32_t c = 0xffffffffu;
            if (occ_item1) c = __ldg(reinterpret_cast<const uint32_t *>(A.scene + A.items_ix + (size_t)(occ_item1 - 1u) * PM_ITEM_SIZE + PM_FILL_RGBA));
            *reinterpret_cast<uint32_t *>(dst) = c;
            if (F32) *dst32 = make_float4((float)(c & 0xff) / 255.0f, (float)((c >> 8) & 0xff) / 255.0f, (float)((c >> 16) & 0xff) / 255.0f, (float)(c >> 24) / 255.0f);
        } else {
            *reinterpret_cast<uint32_t *>(dst) = pm_encode_pixel<EXACT>(c0, c1, c2);
            if (F32) *dst32 = make_float4(pm_linear_to_srgb<EXACT>(c0), pm_linear_to_srgb<EXACT>(c1), pm_linear_to_srgb<EXACT>(c2), 1.0f);
        }
    }
    __syncthreads();
}

#ifndef PM_HEAVY_CTAS
#define PM_HEAVY_CTAS 4   // resident CTAs per SM the kernel is compiled for (4: 64 registers, a few spills; 3: 80 registers)
#endif
template <bool F32, bool EXACT>
__global__ void __launch_bounds__(PM_HEAVY_THREADS, PM_HEAVY_CTAS) k_heavy(const PmFrameArgs A) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    HeavySmem *sh = reinterpret_cast<HeavySmem *>(s_raw);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < PM_HEAVY_WARPS * 256; i += PM_HEAVY_THREADS) { (&sh->acc[0][0])[i] = 0; (&sh->cov[0][0])[i] = 0; }
    // Programmatic dependent launch: wait for binning (k_row) to complete, THEN let k_fine launch beside this grid;
    // k_fine itself does not wait at its start (see pm_fine.cu).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t n_heavy = A.counters->n_heavy;
    const uint32_t *list = A.complex_list + (size_t)A.n_rows * A.n_tx;
    __syncthreads();
    uint32_t n_min_cta = 0;
    if (pm_heavy_warp_mode(n_heavy, gridDim.x, n_heavy + A.counters->n_medium + A.counters->n_mid + A.counters->n_low)) {
        // many heavy tiles: a warp each (up to PM_HEAVY_WARP_CAP records); what is left is drawn CTA-wise below
        HeavyWarpState *ws = &sh->ws[warp];
        for (;;) {
            uint32_t h = 0;
            if (lane == 0) h = atomicAdd(&A.queue->heavy_warp_next, 1u);
            h = __shfl_sync(PM_FULL_MASK, h, 0);
            if (h >= n_heavy) break;
            pm_heavy_tile_warp<F32, EXACT>(A, sh->acc[warp], sh->cov[warp], ws->rgb, ws->dir, list[h], lane);
            __syncwarp();
        }
        n_min_cta = PM_HEAVY_WARP_CAP + 1u;
        __syncthreads();
        // (the sort arrays held the warps' colour planes: nothing of them is live any more)
    }
    for (;;) {
        if (threadIdx.x == 0) sh->tile_next = atomicAdd(&A.queue->heavy_next, 1u);
        __syncthreads();
        const uint32_t h = sh->tile_next;
        __syncthreads();
        if (h >= n_heavy) break;
        heavy_tile_cta<F32, EXACT>(A, sh, list[h], n_min_cta);
    }
}

}  // namespace

template <bool F32, bool EXACT>
static cudaError_t heavy_attr() {
    return cudaFuncSetAttribute(k_heavy<F32, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeavySmem));
}

int pm_heavy_setup(void) {
    cudaError_t e;
    if ((e = heavy_attr<false, false>()) != cudaSuccess) return (int)e;
    if ((e = heavy_attr<false, true>()) != cudaSuccess) return (int)e;
    if ((e = heavy_attr<true, false>()) != cudaSuccess) return (int)e;
    if ((e = heavy_attr<true, true>()) != cudaSuccess) return (int)e;
    return 0;
}

template <bool F32, bool EXACT>
static cudaError_t heavy_launch(const PmFrameArgs &a, int grid, bool overlap, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PM_HEAVY_THREADS); cfg.dynamicSmemBytes = sizeof(HeavySmem); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_heavy<F32, EXACT>, a);
}

cudaError_t pm_launch_heavy(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s) {
    // Persistent: the number of heavy tiles is only known on the device; CTAs without work leave at once.  One CTA of
    // eight warps per SM: three quarters of the SM's registers stay free for k_fine's CTAs, which run beside this kernel.
    const int grid = sm_count * (a.heavy_ctas_per_sm ? (int)a.heavy_ctas_per_sm : PM_HEAVY_GRID_PER_SM);
    const bool exact = (a.flags & PM_FLAG_EXACT_SRGB) != 0;
    if (a.fb32) return exact ? heavy_launch<true, true>(a, grid, overlap, s) : heavy_launch<true, false>(a, grid, overlap, s);
    return exact ? heavy_launch<false, true>(a, grid, overlap, s) : heavy_launch<false, false>(a, grid, overlap, s);
}
