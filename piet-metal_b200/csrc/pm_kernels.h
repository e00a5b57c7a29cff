// Host-visible interface of the CUDA kernels (pm_kernels.cu) used by the renderer (pm_renderer.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm_pixel_logic.h"

// Per-frame counters.  Two sets alternate by frame parity so that the fill kernel of frame f can
// clear the set frame f+1 will use (no memset node in the frame).  Each counter sits on a cache line of its own:
// binning adds to all of them concurrently, and atomics on one line are served one at a time.
struct PmBinCounters {
    uint32_t n_complex;   // (unused: the number of tiles with records is the sum of the four classes below)
    uint32_t pad0[31];
    uint32_t n_overflow;  // pool records allocated behind the inline slots (overflow blocks, headers included)
    uint32_t pad1[31];
    uint32_t n_heavy;     // tiles with at least PM_HEAVY_MIN records: rendered by k_heavy
    uint32_t pad2[31];
    uint32_t n_medium;    // tiles with PM_MEDIUM_MIN .. PM_WARP_RECORDS records: k_fine starts with these (the long jobs)
    uint32_t pad3[31];
    uint32_t n_mid;       // tiles with PM_MID_MIN .. PM_MEDIUM_MIN - 1 records: next
    uint32_t pad4[31];
    uint32_t n_low;       // tiles with fewer: last, so that the kernel ends on its cheapest jobs
    uint32_t pad5[31];
};
// The work lists of the fill kernels, written by k_list after binning from the final record counts: four disjoint
// classes, each in its own quarter of complex_list (n_tiles entries: low | heavy | medium | mid), entries in tile order.
#ifndef PM_MEDIUM_MIN
#define PM_MEDIUM_MIN 6u
#endif
#ifndef PM_MID_MIN
#define PM_MID_MIN 3u     // (8192^2 tiger, with six single-tile tickets per warp at the end of the list: 3 -> frame 134.0 us, 2 -> 135.1, 4 -> 137.0)
#endif
// Records up to which a tile is k_fine's (one warp per tile); tiles with more go to k_heavy (one CTA per tile).
// 16 = the inline slots.  k_fine can take up to 32 (one record per lane; records 16..31 are the start of the
// tile's first overflow block), measured on the 8192^2 tiger: k_fine +6 us, the frame +7 us -- a 30-record tile
// keeps one warp busy for tens of microseconds, which is exactly what k_heavy is for.
#ifndef PM_WARP_RECORDS
#define PM_WARP_RECORDS 16u
#endif
#define PM_HEAVY_MIN (PM_WARP_RECORDS + 1u)
// Work queues of the fill kernels; cleared by the binning kernel of the same frame.  Each counter sits on a
// cache line of its own.
struct PmFineQueue {
    uint32_t batch_next;    // k_fine: 32-tile batches of solid tiles
    uint32_t pad0[31];
    uint32_t tile_next;     // k_fine: position in the dynamically claimed part of its work list
    uint32_t pad1[31];
    uint32_t heavy_next;    // k_heavy: position in the list of heavy tiles, one CTA each
    uint32_t pad2[31];
    uint32_t heavy_warp_next; // k_heavy, warp mode: position in the same list, one warp each
    uint32_t pad3[31];
};
// Written by the device into mapped host memory at the end of every frame.
struct PmFrameReport {
    uint32_t n_complex;
    uint32_t n_overflow;
    uint32_t frame;
    uint32_t n_heavy;
};

// Plan-time copies of what k_seg needs about a segment / an item, laid out for two 16-byte loads each
// (the scene's own layout would cost five dependent loads per thread, and k_seg is latency bound).
struct alignas(16) PmSegInfo {
    float sx, sy, ex, ey; uint32_t item, k; float hw; uint32_t tag;
    uint32_t t_lo, t_hi, r_lo, bd_base;  // copied from the segment's item (PmItemInfo): one dependent load less in k_seg
};
// PmItemInfo (plan time only: copied into PmSegInfo / PmRowInfo, which is what the frame's kernels read): tile span
// of the item's bbox inside the strip, its colour (rgba8 as encoded), its tag with the even-odd bit (bit 8), and
// w0 = the bits of half its stroke width (Poly / Line).
struct alignas(16) PmItemInfo { uint32_t t_lo, t_hi, r_lo, rows; uint32_t bd_base /* unused */, rgba, tag_flags, w0; };
#define PM_INFO_EVEN_ODD 0x100u

// One k_row unit = (item, tile row, chunk of 32 tiles), with what k_row needs of the item copied in (one dependent load
// less: k_row is a chain of dependent loads and atomics).
struct alignas(16) PmRowInfo {
    uint32_t item, row_chunk;   // tile row << 16 | chunk
    uint32_t bd_row;            // first word of the (item, row)'s slice of the backdrop scratch
    uint32_t t_lo_span;         // first tile column of the item | (tile span << 16)
    uint32_t rgba, tag_flags, w0, pad;
};

struct PmFrameArgs {
    const uint8_t *scene;       // encoded scene in device memory
    uint32_t scene_len;
    uint32_t n_items;
    uint32_t items_ix;
    const unsigned long long *plan_a;  // per item: rows-before << 32 | segments-before (n_items + 1 entries)
    const unsigned long long *plan_b;  // per item: backdrop-scratch words before it
    const uint2 *piece_info;           // per k_seg thread: segment, flags | tile row << 15 | tile column (k_pieces_*)
    const PmSegInfo *seg_info;         // per segment (k_pieces_*)
    const PmRowInfo *row_info;         // per k_row unit (k_plan): everything the warp needs in one 32-byte load
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
    uint32_t *complex_list;     // 4 * n_rows * n_tx: the tiles with records by class (k_list): low | heavy | medium | mid
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
    const float4 *item_paint;   // per item: linear r, g, b and alpha of its colour (k_plan; Circle: opaque black)
    uint32_t heavy_ctas_per_sm; // k_heavy's grid: 1 beside a k_fine that dominates the frame, 3 when the heavy tiles do (host heuristic)
};

// A set of paths in device memory (pm_flatten.cu; the host-side description is pm_path_set in the public header).
struct PmPathSetDev {
    uint32_t n_subpaths, n_segments;
    const uint32_t *first;   // n_subpaths + 1
    const double *start;     // n_subpaths x 2
    const uint8_t *verb;     // n_segments
    const double *ctrl;      // n_segments x 6
    const uint32_t *tag, *rgba, *flags;  // n_subpaths (flags may be null)
    const float *width;      // n_subpaths
};
// Flattening + scene encoding on the device: counts and their prefix (cnt becomes the offsets, *total the number of
// emitted points without the MoveTo points), then points, bounding boxes, items and header into `scene`.
void pm_launch_flat_count(const PmPathSetDev &P, double scale, double tolerance, uint32_t *cnt, long long *bbox, unsigned long long *total, cudaStream_t s);
void pm_launch_flat_emit(const PmPathSetDev &P, double scale, double tolerance, const uint32_t *off, unsigned long long total_points, uint8_t *scene,
                         uint32_t items_ix, uint32_t pts_base, long long *bbox, cudaStream_t s);

struct PmPlanResult { uint32_t n_segments; uint32_t n_rows; unsigned long long bd_words; uint32_t error; uint32_t n_pieces; };

// Validates an encoded scene on the device.  *err (device) becomes non-zero if a ref or count is
// out of bounds or a coordinate is not finite.
void pm_launch_validate(const uint8_t *scene, uint32_t scene_len, uint32_t *err, cudaStream_t s);
// Fills plan_a / plan_b [0..n_items], the item tables and result (device) for the given strip; then, once the host
// knows result->n_rows (and has made room), the k_row unit table.
void pm_launch_plan(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1,
                    uint32_t n_tx, unsigned long long *plan_a, unsigned long long *plan_b, PmItemInfo *item_info,
                    const float *srgb_lut, float4 *item_paint, PmPlanResult *result, cudaStream_t s);
void pm_launch_plan_rows(uint32_t n_items, uint32_t n_units, const unsigned long long *plan_a, const unsigned long long *plan_b,
                         const PmItemInfo *item_info, PmRowInfo *row_info, cudaStream_t s);
// The k_seg work list: count + prefix (result->n_pieces, piece_cnt becomes the per-segment offset), then fill.
void pm_launch_pieces_count(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                            const unsigned long long *plan_a, const unsigned long long *plan_b, uint32_t n_segments, PmSegInfo *seg_info,
                            uint32_t *piece_cnt, PmPlanResult *result, cudaStream_t s);
void pm_launch_pieces_fill(const uint8_t *scene, uint32_t n_items, uint32_t items_ix, uint32_t tile_y0, uint32_t tile_y1, uint32_t n_tx,
                           const unsigned long long *plan_a, uint32_t n_segments, const uint32_t *piece_off, uint2 *piece_info,
                           uint32_t piece_cap, cudaStream_t s);
// One frame: binning (k_seg, k_row), then fill/blend: k_fine (solid tiles and tiles whose records fit the inline
// slots, one warp per tile) and k_heavy (the other tiles, one CTA per tile).  `mid` / `mid2` (optional) are recorded
// between binning and k_heavy / between k_heavy and k_fine.  `overlap`: programmatic dependent launch between the
// frame's kernels and from one frame to the next (no event may sit between them, so the events must be null); k_heavy
// and k_fine then run side by side.  Returns the first launch error.
cudaError_t pm_launch_frame(const PmFrameArgs &a, int sm_count, cudaEvent_t mid, cudaEvent_t mid2, bool overlap, cudaStream_t s,
                            uint32_t *n_launched);
// The fill/blend kernels (pm_fine.cu, pm_heavy.cu) and their one-time set-up on the current device
// (shared-memory attributes).
cudaError_t pm_launch_fine(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s);
cudaError_t pm_launch_heavy(const PmFrameArgs &a, int sm_count, bool overlap, cudaStream_t s);
int pm_fine_setup(void);
int pm_heavy_setup(void);
