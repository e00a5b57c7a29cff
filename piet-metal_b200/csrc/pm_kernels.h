// Host-visible interface of the CUDA kernels (pm_kernels.cu) used by the renderer (pm_renderer.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_pixel_logic.h"

// PM_CTA_TILES (build-time, default 0 = not compiled in; NOT yet validated on a GPU -- written at the end of
// round 1 after the GPU budget was spent, to be measured in round 2): tiles with PM_CTA_MIN..PM_CTA_CAP records
// are listed separately by binning and, when there are few of them, rendered by a whole CTA each (one thread
// per pixel, the eight warps sharing the coverage accumulators) before the per-warp loop starts.  A single warp
// issues ~0.1 instructions per cycle, so such a tile takes 60-110 us on its own and bounds the frame time of a
// narrow multi-GPU strip.  Same per-pixel arithmetic, integer coverage sums: the pixels
// are bit-identical to the per-warp path, so the choice may depend on the strip.
#ifndef PM_CTA_TILES
#define PM_CTA_TILES 0
#endif
#define PM_CTA_MIN 24u        // records from which a tile is "costly"
#define PM_CTA_CAP 256u       // ... and up to which the CTA path takes it (its record index lives in shared memory)
#define PM_CTA_MAX_PER_CTA 6u // the CTA path is used when there are at most this many costly tiles per launched CTA

// Per-frame counters.  Two sets alternate by frame parity so that the fill kernel of frame f can
// clear the set frame f+1 will use (no memset node in the frame).
struct PmBinCounters {
    uint32_t n_complex;   // tiles that own at least one record
    uint32_t n_overflow;  // records that did not fit the inline slots of their tile
    uint32_t n_heavy;     // tiles with more records than inline slots
    uint32_t n_costly;    // tiles with at least PM_CTA_MIN records (only counted when PM_CTA_TILES is compiled in)
};
// Work queues of the fill kernel; cleared by the binning kernel of the same frame.
// The list of tiles with records is handed out through PM_FINE_SUBQ counters instead of one: position
// s + PM_FINE_SUBQ * k belongs to counter s.  A single counter would see one atomic every few cycles,
// which is what an L2 slice can do on one address: the claims would queue up for microseconds.
#define PM_FINE_SUBQ 8
struct PmFineQueue {
    uint32_t batch_next;    // 32-tile batches of solid tiles
    uint32_t costly_next;   // costly tiles, claimed by whole CTAs (PM_CTA_TILES)
    uint32_t pad[62];
    uint32_t sub[PM_FINE_SUBQ][64];  // [s][0]: next k of sub-queue s (each on a cache line of its own)
};
// Written by the device into mapped host memory at the end of every frame.
struct PmFrameReport {
    uint32_t n_complex;
    uint32_t n_overflow;
    uint32_t frame;
    uint32_t pad;
};

// Plan-time copies of what k_seg needs about a segment / an item, laid out for two 16-byte loads each
// (the scene's own layout would cost five dependent loads per thread, and k_seg is latency bound).
struct alignas(16) PmSegInfo { float sx, sy, ex, ey; uint32_t item, k; float hw; uint32_t tag; };
struct alignas(16) PmItemInfo { uint32_t t_lo, t_hi, r_lo, rows; unsigned long long bd_base; uint32_t pad[2]; };

struct PmFrameArgs {
    const uint8_t *scene;       // encoded scene in device memory
    uint32_t scene_len;
    uint32_t n_items;
    uint32_t items_ix;
    const unsigned long long *plan_a;  // per item: rows-before << 32 | segments-before (n_items + 1 entries)
    const unsigned long long *plan_b;  // per item: backdrop-scratch words before it
    const uint2 *piece_info;           // per k_seg thread: segment, flags | tile row << 15 | tile column (k_pieces_*)
    const PmSegInfo *seg_info;         // per segment (k_pieces_*)
    const PmItemInfo *item_info;       // per item (k_plan)
    const uint2 *row_info;             // per k_row unit: item, tile row << 16 | 32-tile chunk
    uint32_t n_segments;        // segments of the Fill / Poly items that touch the strip
    uint32_t n_pieces;          // k_seg threads
    uint32_t n_row_units;       // k_row warps: (item, tile row) pairs inside the strip
    uint32_t *bd;               // backdrop scratch of this frame, all zero when the frame starts
    uint32_t *bd_next;          // the next frame's (the two alternate); k_seg clears it
    unsigned long long bd_quads; // size of each in 16-byte units
    uint32_t tile_y0;           // first tile row of the strip
    uint32_t n_rows;            // tile rows in the strip
    uint32_t n_tx;              // tiles per row
    unsigned long long *occ;    // n_rows * n_tx stamped words (see pm_pixel_logic.h)
    unsigned long long *cnt;
    unsigned long long *ovf;
    PmRecord *pool;             // [n_tiles * PM_TILE_SLOTS inline slots][overflow_cap records]
    uint32_t overflow_cap;
    uint32_t *complex_list;     // 2 * n_rows * n_tx: tiles with records, then (second half) the heavy ones among them
    PmBinCounters *counters;    // this frame's set
    PmBinCounters *counters_next;
    PmFineQueue *queue;
    PmFrameReport *report;      // mapped host memory
    uint32_t stamp;             // frame number + 1, never 0
    uint32_t flags;             // PM_FLAG_*
    uint8_t *fb;                // RGBA8 strip, n_rows*16 rows, pitch bytes apart
    size_t pitch;
    float *fb32;                // optional fp32 RGBA strip (debug renders)
    size_t pitch32;
    unsigned long long *debug;  // optional per-CTA cycle counts of k_seg (PM_DEBUG_SEG=1)
    const float *srgb_lut;      // 512 floats: [0,256) sRGB byte -> linear, [256,512) alpha byte / 255
};

struct PmPlanResult { uint32_t n_segments; uint32_t n_rows; unsigned long long bd_words; uint32_t error; uint32_t n_pieces; };

// Validates an encoded scene on the device.  *err (device) becomes non-zero if a ref or count is
// out of bounds or a coordinate is not finite.
void pm_launch_validate(const uint8_t *scene, uint32_t scene_len, uint32_t *err, cudaStream_t s);
// Fills plan_a / plan_b [0..n_items] and result (device) for the given strip.
void pm_launch_plan(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1,
                    uint32_t n_tx, unsigned long long *plan_a, unsigned long long *plan_b, PmItemInfo *item_info, uint2 *row_info,
                    uint32_t row_info_cap, PmPlanResult *result, cudaStream_t s);
// The k_seg work list: count + prefix (result->n_pieces, piece_cnt becomes the per-segment offset), then fill.
void pm_launch_pieces_count(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                            const unsigned long long *plan_a, uint32_t n_segments, PmSegInfo *seg_info, uint32_t *piece_cnt,
                            PmPlanResult *result, cudaStream_t s);
void pm_launch_pieces_fill(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                           const unsigned long long *plan_a, uint32_t n_segments, const uint32_t *piece_off, uint2 *piece_info,
                           uint32_t piece_cap, cudaStream_t s);
// One frame: binning (k_seg, k_row) then fill/blend (k_fine).  `mid` (optional) is recorded before k_fine.
// `overlap`: programmatic dependent launch between the frame's kernels and from one frame to the next (no event
// may sit between them, so `mid` must be null).
void pm_launch_frame(const PmFrameArgs &a, int sm_count, cudaEvent_t mid, bool overlap, cudaStream_t s);
// The fill/blend kernel alone (pm_fine.cu), and its one-time set-up on the current device
// (shared-memory attributes of the kernel).
void pm_launch_fine(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s);
int pm_fine_setup(void);
