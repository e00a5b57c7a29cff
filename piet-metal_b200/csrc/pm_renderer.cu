// Host side of the renderer behind the C ABI of include/piet_metal_b200.h.
//
// Replaces the reference's Objective-C/Metal dispatch layer (TestApp/PietRenderer.m): pipeline and
// buffer creation (:23-57), per-resize surface allocation (:105-146), scene hand-over (:203-205)
// and the per-frame encode of the two compute passes (:59-88).  The render-pass composite
// (:90-101) has no counterpart: solid tiles are written straight into the framebuffer.
//
// There is no CPU fallback: without a CUDA device every renderer entry point fails with
// PM_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "../../include/piet_metal_b200.h"
#include "pm_kernels.h"
#include "pm_pixel_logic.h"
#include "pm_scene_format.h"

namespace {

thread_local std::string g_last_error;

int cuda_fail(cudaError_t e, const char *what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed at pm_renderer.cu:%d: %s", what, line, cudaGetErrorString(e));
    g_last_error = buf;
    return PM_ERR_CUDA;
}
#define PM_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__);  \
    } while (0)

const int EVENT_RING = 512;

}  // namespace

void pm_set_last_error(const char *text) { g_last_error = text ? text : ""; }  // (pm_group.cu reports through the same channel)

struct pm_renderer {
    int device = 0;
    int sm_count = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;

    // scene
    uint8_t *scene = nullptr;
    size_t scene_cap = 0;
    uint32_t scene_len = 0, n_items = 0, items_ix = 0;
    unsigned long long *plan_a = nullptr, *plan_b = nullptr;  // per-item prefixes (pm_kernels.cu, k_plan)
    size_t plan_cap = 0;
    uint32_t n_segments = 0, n_row_units = 0, n_pieces = 0;
    PmSegInfo *seg_info = nullptr;
    uint32_t *piece_off = nullptr;  // per segment: offset of its pieces in piece_info
    size_t seg_cap = 0;
    PmItemInfo *item_info = nullptr;
    float4 *item_paint = nullptr;   // per item: linear colour + alpha (k_plan)
    uint2 *piece_info = nullptr;
    PmRowInfo *row_info = nullptr;
    size_t piece_cap = 0, row_info_cap = 0;
    uint32_t *bd = nullptr;  // backdrop scratch, zero between frames
    unsigned long long *debug = nullptr;
    size_t bd_cap = 0, bd_words = 0;
    bool have_scene = false, plan_dirty = true;
    int frame_events = 1;       // 0: no events, launches overlap across frames; 1: events around every kernel group (serial);
                                // 2: events around the frame only, its kernels overlap (pm_renderer_set_frame_events)
    double plan_ms = 0.0;       // host wall time of the last plan (validate is not included), milliseconds
    uint32_t *dev_err = nullptr;
    PmPlanResult *dev_plan = nullptr;

    // surface
    uint32_t width = 0, height = 0, n_tx = 0, n_ty = 0, tile_y0 = 0, tile_y1 = 0;
    bool have_surface = false, strip_explicit = false;
    uint8_t *fb = nullptr;
    size_t pitch = 0;
    float *fb32 = nullptr;
    size_t pitch32 = 0;
    unsigned long long *occ = nullptr, *cnt = nullptr, *ovf = nullptr;  // stamped per-tile words
    uint32_t *complex_list = nullptr;
    size_t tiles_cap = 0, fb_cap = 0;

    // per-frame scratch: [n_tiles * PM_TILE_SLOTS inline record slots][overflow_cap records]
    PmRecord *pool = nullptr;
    size_t pool_records = 0;
    uint32_t overflow_cap = 0;
    uint64_t pool_bytes_cfg = 0;
    PmBinCounters *counters = nullptr;  // [2]
    PmFineQueue *queue = nullptr;
    PmFrameReport *report = nullptr;      // mapped host memory
    PmFrameReport *report_dev = nullptr;
    float *lut = nullptr;

    // timing
    cudaEvent_t ev_start[EVENT_RING], ev_mid[EVENT_RING], ev_mid2[EVENT_RING], ev_end[EVENT_RING];
    uint32_t n_launches = 0;    // kernels of the last frame
    uint32_t frame = 0, frames_unsynced = 0, stamp = 0;
    uint32_t retries = 0;
};

namespace {

int use_device(pm_renderer *r) {
    PM_CUDA(cudaSetDevice(r->device));
    return PM_OK;
}

size_t strip_tiles(const pm_renderer *r) { return (size_t)(r->tile_y1 - r->tile_y0) * r->n_tx; }

// The record pool holds PM_TILE_SLOTS inline slots per tile plus `want_overflow` overflow records.
int ensure_pool(pm_renderer *r, uint32_t want_overflow) {
    const size_t want = strip_tiles(r) * PM_TILE_SLOTS + want_overflow;
    if (r->pool && r->pool_records >= want) { r->overflow_cap = (uint32_t)std::min<size_t>(r->pool_records - strip_tiles(r) * PM_TILE_SLOTS, 0xfffffff0u); return PM_OK; }
    if (r->pool) { PM_CUDA(cudaStreamSynchronize(r->stream)); PM_CUDA(cudaFree(r->pool)); r->pool = nullptr; r->pool_records = 0; }
    PM_CUDA(cudaMalloc(&r->pool, want * sizeof(PmRecord)));
    r->pool_records = want;
    r->overflow_cap = want_overflow;
    return PM_OK;
}

int alloc_surface(pm_renderer *r) {
    const uint32_t n_rows = r->tile_y1 - r->tile_y0;
    const size_t n_tiles = (size_t)n_rows * r->n_tx;
    const size_t pitch = (size_t)r->n_tx * PM_TILE_W * 4;
    const size_t fb_bytes = pitch * n_rows * PM_TILE_H;
    PM_CUDA(cudaStreamSynchronize(r->stream));
    r->have_surface = false;  // until every buffer below exists
    if (fb_bytes > r->fb_cap) {
        if (r->fb) PM_CUDA(cudaFree(r->fb));
        r->fb = nullptr;
        r->fb_cap = 0;
        PM_CUDA(cudaMalloc(&r->fb, fb_bytes));
        r->fb_cap = fb_bytes;
    }
    if (n_tiles > r->tiles_cap) {
        if (r->occ) PM_CUDA(cudaFree(r->occ));
        if (r->cnt) PM_CUDA(cudaFree(r->cnt));
        if (r->ovf) PM_CUDA(cudaFree(r->ovf));
        if (r->complex_list) PM_CUDA(cudaFree(r->complex_list));
        r->occ = r->cnt = r->ovf = nullptr; r->complex_list = nullptr;
        r->tiles_cap = 0;
        PM_CUDA(cudaMalloc(&r->occ, n_tiles * sizeof(unsigned long long)));
        PM_CUDA(cudaMalloc(&r->cnt, n_tiles * sizeof(unsigned long long)));
        PM_CUDA(cudaMalloc(&r->ovf, n_tiles * sizeof(unsigned long long)));
        PM_CUDA(cudaMalloc(&r->complex_list, 4 * n_tiles * sizeof(uint32_t)));  // the tiles with records by class: low | heavy | medium | mid (k_list)
        r->tiles_cap = n_tiles;
    }
    if (r->fb32) { PM_CUDA(cudaFree(r->fb32)); r->fb32 = nullptr; }
    r->pitch = pitch;
    // stamp 0 is never used by a frame, so zeroed words read as "empty"
    PM_CUDA(cudaMemsetAsync(r->occ, 0, n_tiles * sizeof(unsigned long long), r->stream));
    PM_CUDA(cudaMemsetAsync(r->cnt, 0, n_tiles * sizeof(unsigned long long), r->stream));
    PM_CUDA(cudaMemsetAsync(r->ovf, 0, n_tiles * sizeof(unsigned long long), r->stream));
    PM_CUDA(cudaMemsetAsync(r->counters, 0, 2 * sizeof(PmBinCounters), r->stream));
    memset(r->report, 0, sizeof(PmFrameReport));  // (a report of the previous surface must not trigger a pool growth for this one)
    {
        uint64_t want = r->pool_bytes_cfg ? r->pool_bytes_cfg / sizeof(PmRecord) : std::max<uint64_t>(1u << 16, n_tiles / 4);
        int st = ensure_pool(r, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 64), 1u << 28));
        if (st != PM_OK) return st;
    }
    r->have_surface = true;
    r->plan_dirty = true;
    return PM_OK;
}

int run_plan_timed(pm_renderer *r);
int run_plan(pm_renderer *r) {
    PmPlanResult res;
    PM_CUDA(cudaMemsetAsync(r->dev_plan, 0, sizeof(PmPlanResult), r->stream));
    pm_launch_plan(r->scene, r->n_items, r->items_ix, r->tile_y0, r->tile_y1, r->n_tx, r->plan_a, r->plan_b, r->item_info,
                   r->lut, r->item_paint, r->dev_plan, r->stream);
    PM_CUDA(cudaGetLastError());
    PM_CUDA(cudaMemcpyAsync(&res, r->dev_plan, sizeof res, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    if (!res.error) {  // the k_row unit table, now that its size is known
        if ((size_t)res.n_rows > r->row_info_cap) {
            if (r->row_info) PM_CUDA(cudaFree(r->row_info));
            r->row_info = nullptr;
            r->row_info_cap = 0;
            PM_CUDA(cudaMalloc(&r->row_info, ((size_t)res.n_rows + res.n_rows / 8 + 1) * sizeof(PmRowInfo)));
            r->row_info_cap = (size_t)res.n_rows + res.n_rows / 8 + 1;
        }
        pm_launch_plan_rows(r->n_items, res.n_rows, r->plan_a, r->plan_b, r->item_info, r->row_info, r->stream);
        PM_CUDA(cudaGetLastError());
    }
    if (res.error) { g_last_error = "scene has more than 2^31 segments or (item, tile row) pairs"; return PM_ERR_INVALID_ARG; }
    if (res.bd_words > (1ull << 31)) { g_last_error = "item bounding boxes cover more than 2^31 tiles in total"; return PM_ERR_NOMEM; }
    r->n_segments = res.n_segments;
    if ((size_t)res.n_segments + 1 > r->seg_cap) {
        if (r->seg_info) PM_CUDA(cudaFree(r->seg_info));
        if (r->piece_off) PM_CUDA(cudaFree(r->piece_off));
        r->seg_info = nullptr; r->piece_off = nullptr;
        PM_CUDA(cudaMalloc(&r->seg_info, ((size_t)res.n_segments + 1) * sizeof(PmSegInfo)));
        PM_CUDA(cudaMalloc(&r->piece_off, ((size_t)res.n_segments + 1) * sizeof(uint32_t)));
        r->seg_cap = (size_t)res.n_segments + 1;
    }
    // the k_seg pieces: count and prefix, then (once the table is large enough) fill
    pm_launch_pieces_count(r->scene, r->n_items, r->items_ix, r->tile_y0, r->tile_y1, r->n_tx, r->plan_a, r->plan_b, res.n_segments,
                           r->seg_info, r->piece_off, r->dev_plan, r->stream);
    PM_CUDA(cudaGetLastError());
    PM_CUDA(cudaMemcpyAsync(&res, r->dev_plan, sizeof res, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    if (res.error) { g_last_error = "scene has more than 2^31 (segment, tile row, tile) pieces"; return PM_ERR_INVALID_ARG; }
    if ((size_t)res.n_pieces + 1 > r->piece_cap) {
        if (r->piece_info) PM_CUDA(cudaFree(r->piece_info));
        r->piece_info = nullptr;
        PM_CUDA(cudaMalloc(&r->piece_info, ((size_t)res.n_pieces + 1) * sizeof(uint2)));
        r->piece_cap = (size_t)res.n_pieces + 1;
    }
    pm_launch_pieces_fill(r->scene, r->n_items, r->items_ix, r->tile_y0, r->tile_y1, r->n_tx, r->plan_a, res.n_segments, r->piece_off,
                          r->piece_info, (uint32_t)r->piece_cap, r->stream);
    PM_CUDA(cudaGetLastError());
    r->n_pieces = res.n_pieces;
    if (getenv("PM_L2_PERSIST")) {
        // Experiment switch (off by default, to be measured): keep the k_seg work list resident in L2 across
        // frames.  Every frame streams its whole framebuffer through L2, so the plan-time tables are read from
        // DRAM again each time (k_seg: 30 % L2 hit rate, latency bound).
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, r->device);
        const size_t bytes = std::min<size_t>((size_t)r->n_pieces * sizeof(uint2), (size_t)std::max(max_persist, 0));
        if (bytes) {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
            cudaStreamAttrValue v;
            memset(&v, 0, sizeof v);
            v.accessPolicyWindow.base_ptr = r->piece_info;
            v.accessPolicyWindow.num_bytes = bytes;
            v.accessPolicyWindow.hitRatio = 1.0f;
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(r->stream, cudaStreamAttributeAccessPolicyWindow, &v);
            cudaGetLastError();
        }
    }
    if (getenv("PM_DEBUG_SEG") || getenv("PM_DEBUG_FINE")) {
        if (r->debug) cudaFree(r->debug);
        PM_CUDA(cudaMalloc(&r->debug, (size_t)(1u << 20) * 8));
        {
            std::vector<unsigned long long> init(1u << 20);
            for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ull; init[i + 1] = 0; }
            PM_CUDA(cudaMemcpy(r->debug, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
        }
    }
    r->n_row_units = res.n_rows;
    const size_t want = (((size_t)res.bd_words + 1) + 3) & ~(size_t)3;  // whole 16-byte units (k_seg clears with 128-bit stores)
    if (want > r->bd_cap) {
        if (r->bd) PM_CUDA(cudaFree(r->bd));
        r->bd = nullptr;
        PM_CUDA(cudaMalloc(&r->bd, 2 * want * sizeof(uint32_t)));  // two buffers: frames alternate, each clears the other's
        r->bd_cap = want;
    }
    r->bd_words = want;
    PM_CUDA(cudaMemsetAsync(r->bd, 0, 2 * want * sizeof(uint32_t), r->stream));
    r->plan_dirty = false;
    return PM_OK;
}

int run_plan_timed(pm_renderer *r) {
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const int st = run_plan(r);
    if (st == PM_OK) cudaStreamSynchronize(r->stream);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    r->plan_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
    return st;
}

// Enqueue one frame.  debug_f32 also writes the fp32 parity buffer.
int enqueue_frame(pm_renderer *r, bool debug_f32) {
    if (!r->have_scene || !r->have_surface) return PM_ERR_STATE;
    if (!r->pool || !r->fb) { g_last_error = "surface allocation failed earlier"; return PM_ERR_NOMEM; }
    if (r->plan_dirty) { int st = run_plan_timed(r); if (st != PM_OK) return st; }
    const size_t n_tiles = strip_tiles(r);
    r->stamp++;
    if (r->stamp == 0) {  // 2^32 frames: restart the stamps from a clean slate
        PM_CUDA(cudaMemsetAsync(r->occ, 0, n_tiles * sizeof(unsigned long long), r->stream));
        PM_CUDA(cudaMemsetAsync(r->cnt, 0, n_tiles * sizeof(unsigned long long), r->stream));
        PM_CUDA(cudaMemsetAsync(r->ovf, 0, n_tiles * sizeof(unsigned long long), r->stream));
        r->stamp = 1;
    }
    PmFrameArgs a;
    memset(&a, 0, sizeof a);
    a.scene = r->scene; a.scene_len = r->scene_len; a.n_items = r->n_items; a.items_ix = r->items_ix;
    a.plan_a = r->plan_a; a.plan_b = r->plan_b; a.n_segments = r->n_segments; a.n_row_units = r->n_row_units; a.bd = r->bd + (size_t)(r->frame & 1) * r->bd_words; a.bd_next = r->bd + (size_t)((r->frame + 1) & 1) * r->bd_words; a.bd_quads = r->bd_words / 4;
    a.piece_info = r->piece_info; a.seg_info = r->seg_info; a.n_pieces = r->n_pieces; a.row_info = r->row_info;
    a.tile_y0 = r->tile_y0; a.n_rows = r->tile_y1 - r->tile_y0; a.n_tx = r->n_tx;
    a.occ = r->occ; a.cnt = r->cnt; a.ovf = r->ovf;
    a.pool = r->pool; a.overflow_cap = r->overflow_cap; a.complex_list = r->complex_list;
    a.counters = &r->counters[r->frame & 1]; a.counters_next = &r->counters[(r->frame + 1) & 1];
    a.queue = r->queue; a.report = r->report_dev; a.stamp = r->stamp; a.flags = r->flags;
    a.fb = r->fb; a.pitch = r->pitch;
    a.fb32 = debug_f32 ? r->fb32 : nullptr; a.pitch32 = r->pitch32;
    a.srgb_lut = r->lut;
    a.item_paint = r->item_paint;
    a.debug = r->debug;
    {   // k_heavy's share of the SMs, from what the last reported frame held (mapped host memory, read without a sync:
        // a heuristic, any value renders the same pixels): when the records behind the inline slots outweigh the tiles
        // k_fine draws -- small surfaces, the dense synthetic scenes -- k_heavy is the frame and gets three CTAs per SM
        const volatile PmFrameReport *rep = r->report;
        const uint32_t n_ovf = rep->n_overflow, n_cplx = rep->n_complex;
        a.heavy_ctas_per_sm = (rep->frame != 0 && (uint64_t)n_ovf > 4ull * n_cplx) ? 3u : 1u;
        if (const char *e = getenv("PM_DEBUG_HEAVY_CTAS")) { if (atoi(e) > 0) a.heavy_ctas_per_sm = (uint32_t)atoi(e); }
    }
    const uint32_t slot = r->frame % EVENT_RING;
    const int mode = debug_f32 ? 1 : r->frame_events;
    const bool events = mode != 0, split = mode == 1;
    if (events) PM_CUDA(cudaEventRecord(r->ev_start[slot], r->stream));
    PM_CUDA(pm_launch_frame(a, r->sm_count, split ? r->ev_mid[slot] : nullptr, split ? r->ev_mid2[slot] : nullptr, !split, r->stream, &r->n_launches));
    if (events) PM_CUDA(cudaEventRecord(r->ev_end[slot], r->stream));
    PM_CUDA(cudaGetLastError());
    r->frame++;
    if (events) r->frames_unsynced++;
    return PM_OK;
}

// Wait for the stream; if the overflow part of the record pool was too small, grow it and render
// the frame again.
int finish_frames(pm_renderer *r, bool debug_f32) {
    for (int attempt = 0; attempt < 8; attempt++) {
        PM_CUDA(cudaStreamSynchronize(r->stream));
        // (the report is the last frame's only if its stamp says so: a stale one must not grow the pool)
        if (r->frame == 0 || r->report->frame != r->stamp || r->report->n_overflow <= r->overflow_cap) return PM_OK;
        uint64_t want = std::max<uint64_t>((uint64_t)r->report->n_overflow + r->report->n_overflow / 4, (uint64_t)r->overflow_cap * 2);
        if (want > 0xf0000000ull) { g_last_error = "record pool would exceed 2^32 records"; return PM_ERR_NOMEM; }
        int st = ensure_pool(r, (uint32_t)want);
        if (st != PM_OK) return st;
        r->retries++;
        st = enqueue_frame(r, debug_f32);
        if (st != PM_OK) return st;
    }
    g_last_error = "record pool still too small after 8 retries";
    return PM_ERR_NOMEM;
}

int ensure_scene_cap(pm_renderer *r, size_t len) {
    if (len > r->scene_cap) {
        if (r->scene) PM_CUDA(cudaFree(r->scene));
        r->scene = nullptr;
        r->scene_cap = 0;
        size_t cap = (len + 255) & ~(size_t)255;
        PM_CUDA(cudaMalloc(&r->scene, cap));
        r->scene_cap = cap;
    }
    return PM_OK;
}

// The scene bytes are in r->scene (enqueued on the stream): validate them on the device and size the per-item tables.
int adopt_scene(pm_renderer *r, size_t len) {
    PM_CUDA(cudaMemsetAsync(r->dev_err, 0, sizeof(uint32_t), r->stream));
    pm_launch_validate(r->scene, (uint32_t)len, r->dev_err, r->stream);
    PM_CUDA(cudaGetLastError());
    uint32_t err = 0;
    pm_group_header hdr;
    PM_CUDA(cudaMemcpyAsync(&err, r->dev_err, sizeof err, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaMemcpyAsync(&hdr, r->scene, sizeof hdr, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    if (err) return PM_ERR_SCENE_MALFORMED;
    r->scene_len = (uint32_t)len;
    r->n_items = hdr.n_items;
    r->items_ix = hdr.items_ix;
    if ((size_t)r->n_items + 1 > r->plan_cap) {
        if (r->plan_a) PM_CUDA(cudaFree(r->plan_a));
        if (r->plan_b) PM_CUDA(cudaFree(r->plan_b));
        r->plan_a = r->plan_b = nullptr;
        PM_CUDA(cudaMalloc(&r->plan_a, ((size_t)r->n_items + 1) * sizeof(unsigned long long)));
        PM_CUDA(cudaMalloc(&r->plan_b, ((size_t)r->n_items + 1) * sizeof(unsigned long long)));
        if (r->item_info) PM_CUDA(cudaFree(r->item_info));
        if (r->item_paint) PM_CUDA(cudaFree(r->item_paint));
        r->item_info = nullptr; r->item_paint = nullptr;
        PM_CUDA(cudaMalloc(&r->item_info, ((size_t)r->n_items + 1) * sizeof(PmItemInfo)));
        PM_CUDA(cudaMalloc(&r->item_paint, ((size_t)r->n_items + 1) * sizeof(float4)));
        r->plan_cap = (size_t)r->n_items + 1;
    }
    r->have_scene = true;
    r->plan_dirty = true;
    return PM_OK;
}

int install_scene(pm_renderer *r, const void *src, size_t len, cudaMemcpyKind kind) {
    if (!src || len < PM_GROUP_HEADER_SIZE || len > 0xfffffff0ull) return len < PM_GROUP_HEADER_SIZE ? PM_ERR_SCENE_MALFORMED : PM_ERR_INVALID_ARG;
    PM_CUDA(cudaStreamSynchronize(r->stream));
    int st = ensure_scene_cap(r, len);
    if (st != PM_OK) return st;
    r->have_scene = false;
    PM_CUDA(cudaMemcpyAsync(r->scene, src, len, kind, r->stream));
    return adopt_scene(r, len);
}

// Flattening + encoding on the device (pm_flatten.cu).  One temporary allocation holds the uploaded path set and the
// scratch of the kernels; it is freed before returning (this runs once per scene, not per frame).
int install_paths(pm_renderer *r, const pm_path_set *ps, double scale, double tolerance) {
    const size_t ns = ps->n_subpaths, ng = ps->n_segments;
    if (ns == 0 || ns > 0x3fffffffull || ng > 0x7fffffffull) return PM_ERR_INVALID_ARG;
    if (!ps->first_segment || !ps->start || !ps->tag || !ps->rgba || !ps->width || (ng && (!ps->verb || !ps->ctrl))) return PM_ERR_INVALID_ARG;
    if (ps->first_segment[0] != 0 || ps->first_segment[ns] != ng) return PM_ERR_INVALID_ARG;
    for (size_t i = 0; i < ns; i++) {
        if (ps->first_segment[i] > ps->first_segment[i + 1]) return PM_ERR_INVALID_ARG;
        if (ps->tag[i] != PM_ITEM_FILL && ps->tag[i] != PM_ITEM_POLY) return PM_ERR_INVALID_ARG;
    }
    if (!(scale == scale) || !(tolerance > 0.0)) return PM_ERR_INVALID_ARG;
    PM_CUDA(cudaStreamSynchronize(r->stream));
    r->have_scene = false;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_first = 0, o_start = o_first + al((ns + 1) * 4), o_ctrl = o_start + al(ns * 16), o_verb = o_ctrl + al(ng * 48 + 16),
                 o_tag = o_verb + al(ng + 1), o_rgba = o_tag + al(ns * 4), o_flags = o_rgba + al(ns * 4), o_width = o_flags + al(ns * 4),
                 o_cnt = o_width + al(ns * 4), o_bbox = o_cnt + al((ng + 1) * 4), o_total = o_bbox + al(ns * 32), total_bytes = o_total + 256;
    uint8_t *tmp = nullptr;
    PM_CUDA(cudaMalloc(&tmp, total_bytes));
    auto fail = [&](int st) { cudaFree(tmp); return st; };
#define PM_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(cuda_fail(e_, #call, __LINE__)); } while (0)
    PM_TRY(cudaMemcpyAsync(tmp + o_first, ps->first_segment, (ns + 1) * 4, cudaMemcpyHostToDevice, r->stream));
    PM_TRY(cudaMemcpyAsync(tmp + o_start, ps->start, ns * 16, cudaMemcpyHostToDevice, r->stream));
    if (ng) {
        PM_TRY(cudaMemcpyAsync(tmp + o_ctrl, ps->ctrl, ng * 48, cudaMemcpyHostToDevice, r->stream));
        PM_TRY(cudaMemcpyAsync(tmp + o_verb, ps->verb, ng, cudaMemcpyHostToDevice, r->stream));
    }
    PM_TRY(cudaMemcpyAsync(tmp + o_tag, ps->tag, ns * 4, cudaMemcpyHostToDevice, r->stream));
    PM_TRY(cudaMemcpyAsync(tmp + o_rgba, ps->rgba, ns * 4, cudaMemcpyHostToDevice, r->stream));
    PM_TRY(cudaMemcpyAsync(tmp + o_width, ps->width, ns * 4, cudaMemcpyHostToDevice, r->stream));
    if (ps->flags) PM_TRY(cudaMemcpyAsync(tmp + o_flags, ps->flags, ns * 4, cudaMemcpyHostToDevice, r->stream));
    PmPathSetDev P;
    P.n_subpaths = (uint32_t)ns; P.n_segments = (uint32_t)ng;
    P.first = reinterpret_cast<const uint32_t *>(tmp + o_first);
    P.start = reinterpret_cast<const double *>(tmp + o_start);
    P.ctrl = reinterpret_cast<const double *>(tmp + o_ctrl);
    P.verb = tmp + o_verb;
    P.tag = reinterpret_cast<const uint32_t *>(tmp + o_tag);
    P.rgba = reinterpret_cast<const uint32_t *>(tmp + o_rgba);
    P.flags = ps->flags ? reinterpret_cast<const uint32_t *>(tmp + o_flags) : nullptr;
    P.width = reinterpret_cast<const float *>(tmp + o_width);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(tmp + o_cnt);
    long long *bbox = reinterpret_cast<long long *>(tmp + o_bbox);
    unsigned long long *total_dev = reinterpret_cast<unsigned long long *>(tmp + o_total);
    pm_launch_flat_count(P, scale, tolerance, cnt, bbox, total_dev, r->stream);
    PM_TRY(cudaGetLastError());
    unsigned long long total = 0;
    PM_TRY(cudaMemcpyAsync(&total, total_dev, 8, cudaMemcpyDeviceToHost, r->stream));
    PM_TRY(cudaStreamSynchronize(r->stream));
    const unsigned long long items_ix = PM_GROUP_HEADER_SIZE + ns * PM_BBOX_SIZE, pts_base = items_ix + ns * PM_ITEM_SIZE;
    const unsigned long long len = pts_base + 8ull * (total + ns);
    if (len > 0xfffffff0ull) { g_last_error = "the flattened scene would exceed 4 GiB"; return fail(PM_ERR_NOMEM); }
    int st = ensure_scene_cap(r, (size_t)len);
    if (st != PM_OK) return fail(st);
    pm_launch_flat_emit(P, scale, tolerance, cnt, total, r->scene, (uint32_t)items_ix, (uint32_t)pts_base, bbox, r->stream);
    PM_TRY(cudaGetLastError());
#undef PM_TRY
    st = adopt_scene(r, (size_t)len);  // (synchronises the stream: the temporary can go)
    cudaFree(tmp);
    return st;
}

}  // namespace

extern "C" {

const char *pm_last_error(void) { return g_last_error.c_str(); }

int pm_renderer_create(pm_renderer **out, const pm_config *cfg) {
    if (!out) return PM_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_last_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return PM_ERR_NO_DEVICE;
    }
    int device = cfg ? cfg->device : 0;
    if (device < 0 || device >= n_dev) return PM_ERR_INVALID_ARG;
    cudaDeviceProp prop;
    PM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        g_last_error = std::string("device ") + prop.name + " is not sm_100: the kernels are built for sm_100a only";
        return PM_ERR_NO_DEVICE;
    }
    pm_renderer *r = new (std::nothrow) pm_renderer();
    if (!r) return PM_ERR_NOMEM;
    r->device = device;
    r->sm_count = prop.multiProcessorCount;
    r->flags = cfg ? cfg->flags : 0;
    r->pool_bytes_cfg = cfg ? cfg->scratch_bytes : 0;
    int st = PM_OK;
    auto fail = [&](int s) { pm_renderer_destroy(r); return s; };
    if ((st = use_device(r)) != PM_OK) return fail(st);
#define PM_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(cuda_fail(e_, #call, __LINE__)); } while (0)
    PM_TRY(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    for (int i = 0; i < EVENT_RING; i++) {
        PM_TRY(cudaEventCreate(&r->ev_start[i]));
        PM_TRY(cudaEventCreate(&r->ev_mid[i]));
        PM_TRY(cudaEventCreate(&r->ev_mid2[i]));
        PM_TRY(cudaEventCreate(&r->ev_end[i]));
    }
    PM_TRY(cudaMalloc(&r->counters, 2 * sizeof(PmBinCounters)));
    PM_TRY(cudaMalloc(&r->queue, sizeof(PmFineQueue)));
    PM_TRY(cudaMalloc(&r->dev_err, sizeof(uint32_t)));
    PM_TRY(cudaMalloc(&r->dev_plan, sizeof(PmPlanResult)));
    PM_TRY(cudaMalloc(&r->lut, 512 * sizeof(float)));
    PM_TRY(cudaHostAlloc(&r->report, sizeof(PmFrameReport), cudaHostAllocMapped));
    memset(r->report, 0, sizeof(PmFrameReport));
    PM_TRY(cudaHostGetDevicePointer(&r->report_dev, r->report, 0));
    PM_TRY(cudaMemsetAsync(r->counters, 0, 2 * sizeof(PmBinCounters), r->stream));
    PM_TRY(cudaMemsetAsync(r->queue, 0, sizeof(PmFineQueue), r->stream));
    float lut[512];  // [0,256): sRGB byte -> linear; [256,512): alpha byte / 255
    for (int i = 0; i < 256; i++) { lut[i] = pm_srgb_byte_to_linear((uint32_t)i); lut[256 + i] = (float)i / 255.0f; }
    PM_TRY(cudaMemcpyAsync(r->lut, lut, sizeof lut, cudaMemcpyHostToDevice, r->stream));
    PM_TRY((cudaError_t)pm_fine_setup());
    PM_TRY((cudaError_t)pm_heavy_setup());
    PM_TRY(cudaStreamSynchronize(r->stream));
#undef PM_TRY
    *out = r;
    return PM_OK;
}

void pm_renderer_destroy(pm_renderer *r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) cudaStreamSynchronize(r->stream);
    cudaFree(r->scene); cudaFree(r->plan_a); cudaFree(r->plan_b); cudaFree(r->bd); cudaFree(r->seg_info); cudaFree(r->piece_off); cudaFree(r->item_info); cudaFree(r->item_paint); cudaFree(r->piece_info); cudaFree(r->row_info); cudaFree(r->debug); cudaFree(r->dev_err); cudaFree(r->dev_plan);
    cudaFree(r->fb); cudaFree(r->fb32); cudaFree(r->occ); cudaFree(r->cnt); cudaFree(r->ovf); cudaFree(r->complex_list);
    cudaFree(r->pool); cudaFree(r->counters); cudaFree(r->queue); cudaFree(r->lut);
    if (r->report) cudaFreeHost(r->report);
    for (int i = 0; i < EVENT_RING; i++) {
        if (r->ev_start[i]) cudaEventDestroy(r->ev_start[i]);
        if (r->ev_mid[i]) cudaEventDestroy(r->ev_mid[i]);
        if (r->ev_mid2[i]) cudaEventDestroy(r->ev_mid2[i]);
        if (r->ev_end[i]) cudaEventDestroy(r->ev_end[i]);
    }
    if (r->stream) cudaStreamDestroy(r->stream);
    cudaGetLastError();
    delete r;
}

int pm_renderer_resize(pm_renderer *r, uint32_t width, uint32_t height) {
    if (!r || width == 0 || height == 0 || width > 65535 || height > 65535) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    r->width = width; r->height = height;
    r->n_tx = (width + PM_TILE_W - 1) / PM_TILE_W;
    r->n_ty = (height + PM_TILE_H - 1) / PM_TILE_H;
    r->tile_y0 = 0; r->tile_y1 = r->n_ty;
    r->strip_explicit = false;
    return alloc_surface(r);
}

int pm_renderer_set_strip(pm_renderer *r, uint32_t tile_y0, uint32_t tile_y1) {
    if (!r) return PM_ERR_INVALID_ARG;
    if (r->n_ty == 0) return PM_ERR_STATE;
    if (tile_y0 >= tile_y1 || tile_y1 > r->n_ty) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    r->tile_y0 = tile_y0; r->tile_y1 = tile_y1;
    r->strip_explicit = true;
    return alloc_surface(r);
}

int pm_renderer_set_scene(pm_renderer *r, const uint8_t *scene, size_t len) {
    if (!r || !scene) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    return install_scene(r, scene, len, cudaMemcpyHostToDevice);
}

int pm_renderer_set_scene_device(pm_renderer *r, const void *scene_dev, size_t len) {
    if (!r || !scene_dev) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    return install_scene(r, scene_dev, len, cudaMemcpyDeviceToDevice);
}

int pm_renderer_set_scene_paths(pm_renderer *r, const pm_path_set *paths, double scale, double tolerance) {
    if (!r || !paths) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    return install_paths(r, paths, scale, tolerance);
}

int pm_renderer_read_scene(pm_renderer *r, uint8_t *dst, size_t cap, size_t *len) {
    if (!r) return PM_ERR_INVALID_ARG;
    if (!r->have_scene) return PM_ERR_STATE;
    if (len) *len = r->scene_len;
    if (!dst || cap < r->scene_len) return PM_ERR_BUFFER_TOO_SMALL;
    int st = use_device(r);
    if (st != PM_OK) return st;
    PM_CUDA(cudaMemcpyAsync(dst, r->scene, r->scene_len, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    return PM_OK;
}

int pm_renderer_render(pm_renderer *r) {
    if (!r) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    return enqueue_frame(r, false);
}

int pm_renderer_set_frame_events(pm_renderer *r, int enabled) {
    if (!r) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    PM_CUDA(cudaStreamSynchronize(r->stream));
    r->frame_events = enabled < 0 || enabled > 2 ? 1 : enabled;
    r->frames_unsynced = 0;
    return PM_OK;
}

int pm_renderer_sync(pm_renderer *r, pm_frame_stats *stats) {
    if (!r) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    const uint32_t retries_before = r->retries;
    st = finish_frames(r, false);
    if (st != PM_OK) return st;
    if (r->debug && getenv("PM_DEBUG_SEG")) {
        size_t nb = std::min<size_t>((size_t)r->n_pieces / 256 + 1, 1u << 19);
        std::vector<unsigned long long> d(2 * nb);
        cudaMemcpy(d.data(), r->debug, 2 * nb * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull, t1 = 0;
        for (size_t i = 0; i < nb; i++) { t0 = std::min(t0, d[2 * i]); t1 = std::max(t1, d[2 * i + 1]); }
        std::vector<size_t> order(nb);
        for (size_t i = 0; i < nb; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return d[2 * a + 1] > d[2 * b + 1]; });
        fprintf(stderr, "[PM_DEBUG_SEG] k_seg span %.1f us over %zu CTAs\n", (t1 - t0) * 1e-3, nb);
        for (size_t i = 0; i < std::min<size_t>(nb, 6); i++)
            fprintf(stderr, "[PM_DEBUG_SEG]   block %zu: start +%.1f us, end +%.1f us\n", order[i], (d[2 * order[i]] - t0) * 1e-3, (d[2 * order[i] + 1] - t0) * 1e-3);
        std::vector<unsigned long long> init(2 * nb);
        for (size_t i = 0; i < nb; i++) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
        cudaMemcpy(r->debug, init.data(), 2 * nb * 8, cudaMemcpyHostToDevice);
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        uint32_t n = r->frame_events ? std::min<uint32_t>(r->frames_unsynced, EVENT_RING) : 0;
        const bool split = r->frame_events == 1;
        double sum_total = 0, sum_bin = 0, sum_fine = 0, sum_heavy = 0;
        for (uint32_t k = 0; k < n; k++) {
            uint32_t slot = (r->frame - 1 - k) % EVENT_RING;
            float t = 0, b = 0, f = 0, h = 0;
            PM_CUDA(cudaEventElapsedTime(&t, r->ev_start[slot], r->ev_end[slot]));
            if (split) {
                PM_CUDA(cudaEventElapsedTime(&b, r->ev_start[slot], r->ev_mid[slot]));
                PM_CUDA(cudaEventElapsedTime(&h, r->ev_mid[slot], r->ev_mid2[slot]));
                PM_CUDA(cudaEventElapsedTime(&f, r->ev_mid2[slot], r->ev_end[slot]));
            }
            if (k == 0) { stats->ms_total = t; stats->ms_bin = b; stats->ms_fine = f; stats->ms_heavy = h; }
            sum_total += t; sum_bin += b; sum_fine += f; sum_heavy += h;
        }
        stats->ms_heavy_sum = (float)sum_heavy;
        stats->n_heavy_tiles = r->report->n_heavy;
        stats->ms_plan = (float)r->plan_ms;
        stats->frames = n;
        stats->ms_total_sum = (float)sum_total;
        stats->ms_bin_sum = (float)sum_bin;
        stats->ms_fine_sum = (float)sum_fine;
        stats->n_tiles = (r->tile_y1 - r->tile_y0) * r->n_tx;
        stats->n_overflow_records = r->report->n_overflow;
        stats->n_complex_tiles = r->report->n_complex;
        stats->n_launches = r->n_launches;  /* k_seg, k_row, k_list, k_heavy, k_fine */
        stats->retries = r->retries - retries_before;
    }
    r->frames_unsynced = 0;
    return PM_OK;
}

int pm_renderer_read_rgba8(pm_renderer *r, uint8_t *dst, size_t stride) {
    if (!r || !dst) return PM_ERR_INVALID_ARG;
    if (!r->have_surface) return PM_ERR_STATE;
    if (stride < (size_t)r->width * 4) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    st = finish_frames(r, false);
    if (st != PM_OK) return st;
    uint32_t y_begin = r->tile_y0 * PM_TILE_H, y_end = std::min(r->tile_y1 * PM_TILE_H, r->height);
    if (stride == r->pitch && (size_t)r->width * 4 == r->pitch)  // contiguous on both sides: one linear copy
        PM_CUDA(cudaMemcpyAsync(dst, r->fb, r->pitch * (y_end - y_begin), cudaMemcpyDeviceToHost, r->stream));
    else
        PM_CUDA(cudaMemcpy2DAsync(dst, stride, r->fb, r->pitch, (size_t)r->width * 4, y_end - y_begin, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    return PM_OK;
}

int pm_renderer_render_host(pm_renderer *r, const uint8_t *scene, size_t len, uint8_t *dst, size_t stride, pm_frame_stats *stats) {
    int st = pm_renderer_set_scene(r, scene, len);
    if (st != PM_OK) return st;
    st = pm_renderer_render(r);
    if (st != PM_OK) return st;
    st = pm_renderer_read_rgba8(r, dst, stride);
    if (st != PM_OK) return st;
    return stats ? pm_renderer_sync(r, stats) : PM_OK;
}

int pm_renderer_framebuffer(pm_renderer *r, void **dev_ptr, size_t *pitch_bytes, uint32_t *rows) {
    if (!r) return PM_ERR_INVALID_ARG;
    if (!r->have_surface) return PM_ERR_STATE;
    if (dev_ptr) *dev_ptr = r->fb;
    if (pitch_bytes) *pitch_bytes = r->pitch;
    if (rows) *rows = (r->tile_y1 - r->tile_y0) * PM_TILE_H;
    return PM_OK;
}

int pm_renderer_stream(pm_renderer *r, void **cuda_stream) {
    if (!r || !cuda_stream) return PM_ERR_INVALID_ARG;
    *cuda_stream = r->stream;
    return PM_OK;
}

int pm_renderer_read_rgba32f(pm_renderer *r, float *dst, size_t stride_bytes) {
    if (!r || !dst) return PM_ERR_INVALID_ARG;
    if (!r->have_surface || !r->have_scene) return PM_ERR_STATE;
    if (stride_bytes < (size_t)r->width * 16) return PM_ERR_INVALID_ARG;
    int st = use_device(r);
    if (st != PM_OK) return st;
    st = finish_frames(r, false);
    if (st != PM_OK) return st;
    const uint32_t n_rows = r->tile_y1 - r->tile_y0;
    r->pitch32 = (size_t)r->n_tx * PM_TILE_W * 16;
    if (!r->fb32) PM_CUDA(cudaMalloc(&r->fb32, r->pitch32 * n_rows * PM_TILE_H));
    st = enqueue_frame(r, true);
    if (st != PM_OK) return st;
    st = finish_frames(r, true);
    if (st != PM_OK) return st;
    uint32_t y_begin = r->tile_y0 * PM_TILE_H, y_end = std::min(r->tile_y1 * PM_TILE_H, r->height);
    PM_CUDA(cudaMemcpy2DAsync(dst, stride_bytes, r->fb32, r->pitch32, (size_t)r->width * 16, y_end - y_begin, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    PM_CUDA(cudaFree(r->fb32));
    r->fb32 = nullptr;
    return PM_OK;
}

int pm_renderer_read_tile_items(pm_renderer *r, uint32_t *offsets, pm_tile_item *items, size_t cap_items, size_t *n_items_out,
                                uint32_t *solid_colors) {
    if (!r || !offsets) return PM_ERR_INVALID_ARG;
    if (!r->have_surface || !r->have_scene) return PM_ERR_STATE;
    int st = use_device(r);
    if (st != PM_OK) return st;
    st = finish_frames(r, false);
    if (st != PM_OK) return st;
    // the per-tile words and the record pool of the last frame are still in place (nothing is ever cleared)
    if (r->frame == 0) {
        st = enqueue_frame(r, false);
        if (st != PM_OK) return st;
        st = finish_frames(r, false);
        if (st != PM_OK) return st;
    }
    const size_t n_tiles = strip_tiles(r);
    const size_t n_rec = n_tiles * PM_TILE_SLOTS + std::min<uint32_t>(r->report->n_overflow, r->overflow_cap);
    const uint32_t stamp = r->stamp;
    std::vector<unsigned long long> occ(n_tiles), cnt(n_tiles), ovf(n_tiles);
    std::vector<PmRecord> rec(n_rec);
    PM_CUDA(cudaMemcpyAsync(occ.data(), r->occ, n_tiles * 8, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaMemcpyAsync(cnt.data(), r->cnt, n_tiles * 8, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaMemcpyAsync(ovf.data(), r->ovf, n_tiles * 8, cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaMemcpyAsync(rec.data(), r->pool, n_rec * sizeof(PmRecord), cudaMemcpyDeviceToHost, r->stream));
    PM_CUDA(cudaStreamSynchronize(r->stream));
    std::vector<uint32_t> rgba_of(r->n_items);
    {
        std::vector<uint8_t> items((size_t)r->n_items * PM_ITEM_SIZE);
        if (r->n_items) PM_CUDA(cudaMemcpy(items.data(), r->scene + r->items_ix, items.size(), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < r->n_items; i++) memcpy(&rgba_of[i], items.data() + (size_t)i * PM_ITEM_SIZE + PM_FILL_RGBA, 4);
    }

    size_t total = 0;
    std::vector<std::pair<uint64_t, uint32_t>> keyed;
    for (size_t t = 0; t < n_tiles; t++) {
        offsets[t] = (uint32_t)total;
        const uint32_t occ_item1 = (uint32_t)(occ[t] >> 32) == stamp ? (uint32_t)occ[t] : 0u;
        const uint32_t occ_rgba = occ_item1 ? rgba_of[occ_item1 - 1] : 0xffffffffu;
        const uint32_t n = (uint32_t)(cnt[t] >> 32) == stamp ? (uint32_t)cnt[t] : 0u;
        keyed.clear();
        bool has_draw = false;
        auto visit = [&](uint32_t idx) {
            const PmRecord &q = rec[idx];
            if (q.item < occ_item1) return;  // below the topmost opaque cover: rewound away
            if ((q.key & 15u) != PM_REC_SOLID) has_draw = true;
            keyed.push_back({((uint64_t)q.item << 32) | q.key, idx});
        };
        for (uint32_t k = 0; k < std::min<uint32_t>(n, PM_TILE_SLOTS); k++) visit((uint32_t)(t * PM_TILE_SLOTS + k));
        if (n > PM_TILE_SLOTS && (uint32_t)(ovf[t] >> 32) == stamp) {
            // the chain of overflow blocks (pm_pixel_logic.h): block j holds positions pm_blk_first(j) .. + pm_blk_size(j)
            uint32_t link = (uint32_t)ovf[t], pos = PM_TILE_SLOTS;
            for (uint32_t j = 0; link != 0 && link != PM_EXT_FAILED && link - 1 < n_rec && pos < n; j++) {
                const uint32_t size = pm_blk_size(j);
                for (uint32_t k = 0; k < size && pos < n && link + k < n_rec; k++, pos++) visit(link + k);
                link = rec[link - 1].next;
            }
        }
        std::sort(keyed.begin(), keyed.end());
        auto push = [&](uint32_t item, int32_t backdrop, uint32_t effect) {
            if (items && total < cap_items) { items[total].item = item; items[total].backdrop = backdrop; items[total].effect = effect; }
            total++;
        };
        if (occ_item1) push(occ_item1 - 1, 0, 1);
        for (size_t k = 0; k < keyed.size(); k++) {
            const PmRecord &q = rec[keyed[k].second];
            const bool last_of_item = k + 1 == keyed.size() || (keyed[k + 1].first >> 32) != q.item;
            if (!last_of_item) continue;  // the item's trailer (DrawFill / Stroke) or its only record sorts last
            const uint32_t kind = q.key & 15u;
            if (pm_rec_is_drawfill(kind)) push(q.item, (int32_t)pm_f2u(q.p[0]), 0);
            else if (kind == PM_REC_SOLID) push(q.item, 0, 1);
            else push(q.item, 0, 0);
        }
        if (solid_colors) solid_colors[t] = has_draw ? 0u : occ_rgba;
    }
    offsets[n_tiles] = (uint32_t)total;
    if (n_items_out) *n_items_out = total;
    return (items && total > cap_items) ? PM_ERR_BUFFER_TOO_SMALL : PM_OK;
}

/* Debug builds only (PM_DEBUG_FINE=1 and a library built with -DPM_FINE_TIMELINE=1): copies the first
 * n_words 64-bit words of the kernels' debug buffer.  Not declared in the public header. */
int pm_debug_read(pm_renderer *r, unsigned long long *dst, size_t n_words) {
    if (!r || !dst || !r->debug || n_words > (1u << 20)) return PM_ERR_STATE;
    PM_CUDA(cudaStreamSynchronize(r->stream));
    PM_CUDA(cudaMemcpy(dst, r->debug, n_words * 8, cudaMemcpyDeviceToHost));
    {   // start over: the kernels keep the minimum start / maximum end per slot
        std::vector<unsigned long long> init(1u << 20);
        for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ull; init[i + 1] = 0; }
        PM_CUDA(cudaMemcpy(r->debug, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    }
    return PM_OK;
}

int pm_host_alloc(void **out, size_t bytes) {
    if (!out) return PM_ERR_INVALID_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? PM_ERR_NO_DEVICE : cuda_fail(e, "cudaHostAlloc", __LINE__); }
    return PM_OK;
}

void pm_host_free(void *p) {
    if (p) cudaFreeHost(p);
    cudaGetLastError();
}

}  // extern "C"
