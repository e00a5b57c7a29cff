/*
 * piet_metal_b200.h -- C ABI of the B200-native drop-in for piet-metal's compute render path.
 *
 * Plain C types only (pointers, sizes, fixed-width integers) so that the reference's Rust host can
 * bind it with an `extern "C"` block and its Objective-C host can include it directly; see
 * INTEGRATION.md for the binding a maintainer would add.  Every function returns 0 (PM_OK) or a
 * negative pm_status and never throws or aborts across the boundary.
 *
 * What each group replaces in the reference (paths relative to linebender/piet-metal):
 *
 *   feed      init_test_scene                 include/piet_metal.h:3, src/lib.rs:387-393
 *             pm_encoder_*                    `Encoder` src/lib.rs:79-254
 *             pm_scene_build                  make_tiger / make_cardioid / make_path_test
 *                                             src/lib.rs:257-328 (+ the synthetic stress scenes)
 *   renderer  pm_renderer_create/destroy      -[PietRenderer initWithMetalKitView:]
 *                                             TestApp/PietRenderer.m:23-57
 *             pm_renderer_resize              -[PietRenderer mtkView:drawableSizeWillChange:]
 *                                             TestApp/PietRenderer.m:105-146
 *             pm_renderer_set_scene           -[PietRenderer initScene] TestApp/PietRenderer.m:203-205
 *                                             (the 16 MiB shared MTLBuffer, :52-53, is gone: the
 *                                             renderer owns device memory sized to the scene)
 *             pm_renderer_render              -[PietRenderer drawInMTKView:] compute passes,
 *                                             TestApp/PietRenderer.m:59-88, i.e. tileKernel +
 *                                             renderKernel of TestApp/PietRender.metal:160-566
 *                                             and the solid-tile composite (:16-44)
 *
 * Threading: one renderer is driven by one host thread at a time (like an MTKView delegate).
 * Ownership: the renderer owns all device memory; the caller owns every host buffer it passes in.
 */
#ifndef PIET_METAL_B200_H
#define PIET_METAL_B200_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h> /* ssize_t, as in the reference header */

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pm_status {
    PM_OK = 0,
    PM_ERR_INVALID_ARG = -1,
    PM_ERR_NO_DEVICE = -2,      /* no CUDA device / driver: there is no CPU fallback */
    PM_ERR_CUDA = -3,           /* a CUDA call failed; pm_last_error() has the text */
    PM_ERR_SCENE_MALFORMED = -4,/* a ref or count in the scene points outside the buffer */
    PM_ERR_BUFFER_TOO_SMALL = -5,
    PM_ERR_STATE = -6,          /* render before set_scene/resize, encoder misuse, ... */
    PM_ERR_PARSE = -7,          /* SVG path data / path list could not be parsed */
    PM_ERR_NOMEM = -8
} pm_status;

const char *pm_strerror(int status);
/* Text of the last failing CUDA call on this thread ("" if none). */
const char *pm_last_error(void);
/* Library version, "major.minor.patch". */
const char *pm_version(void);

/* ------------------------------------------------------------------------------------------- */
/* Feed: scene encoding (host only, no GPU needed)                                              */
/* ------------------------------------------------------------------------------------------- */

/* Same symbol, signature and behaviour as the reference (include/piet_metal.h:3): writes the
 * default test scene (the tiger at scale 8, src/lib.rs:286-328,369-373) from offset 0 of `buf`.
 * The reference panics on overflow; this one writes nothing past buf_size and leaves n_items = 0. */
void init_test_scene(uint8_t *buf, ssize_t buf_size);

/* Bump-allocating scene writer, mirroring `Encoder` (src/lib.rs:79-254).  Colours are
 * 0xRRGGBBAA as in the reference API and are stored byte-swapped (rgba.to_be(), lib.rs:181). */
typedef struct pm_encoder pm_encoder;
int pm_encoder_new(pm_encoder **out, uint8_t *buf, size_t cap);          /* Encoder::new   :104 */
int pm_encoder_begin_group(pm_encoder *e, uint32_t n_items);             /* begin_group    :132 */
int pm_encoder_end_group(pm_encoder *e);                                 /* end_group      :146 */
int pm_encoder_circle(pm_encoder *e, double cx, double cy, double r);    /* circle         :167 */
int pm_encoder_stroke_line(pm_encoder *e, double x0, double y0, double x1, double y1,
                           float width, uint32_t rgba);                  /* stroke_line    :177 */
int pm_encoder_fill(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba);   /* :195 */
int pm_encoder_polyline(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba,
                        float width);                                    /* polyline       :209 */
/* Extensions for what the reference leaves open ("need to deal with subpaths", src/lib.rs:194; flags "will be
 * used for winding rule", TestApp/SceneEncoder.h:44).  pm_encoder_fill_rule is pm_encoder_fill with the item's flags word
 * (PM_FILL_NONZERO / PM_FILL_EVEN_ODD; read by a renderer created with PM_FLAG_FILL_RULES).  pm_encoder_fill_subpaths
 * writes ONE Fill item for a path of several closed subpaths -- holes are cut out instead of painted over: the
 * subpaths are joined into a single closed point list by zero-area bridges (each subpath is closed explicitly and
 * followed by a segment back to the first point of the path; every bridge is walked once in each direction, so its
 * winding and area cancel), which the reference's own kernels can render as it is.  counts[i] = points of subpath i. */
int pm_encoder_fill_rule(pm_encoder *e, const double *xy, uint32_t n_points, uint32_t rgba, uint32_t flags);
int pm_encoder_fill_subpaths(pm_encoder *e, const double *xy, const uint32_t *counts, uint32_t n_subpaths,
                             uint32_t rgba, uint32_t flags);
/* Bytes used so far (free_space, lib.rs:112-116). */
size_t pm_encoder_bytes(const pm_encoder *e);
void pm_encoder_free(pm_encoder *e);

/* Path-level helpers: parse SVG path data (kurbo BezPath::from_svg as used at lib.rs:296,312),
 * scale it, flatten it with flatten_path's rule (src/flatten.rs:10-47) and report the subpaths.
 * out_xy receives x,y pairs of all subpaths back to back; out_counts the points per subpath.
 * Returns the number of subpaths, or a negative pm_status (PM_ERR_BUFFER_TOO_SMALL if either
 * capacity is exceeded; *need_points is then the total required). */
int64_t pm_flatten_svg_path(const char *d, double scale, double tolerance, double *out_xy,
                            size_t cap_points, uint32_t *out_counts, size_t cap_subpaths,
                            size_t *need_points);
/* parse_color (src/lib.rs:375-385): "#RGB" / "#RRGGBB" -> 0xRRGGBBff, anything else 0xff00ff80. */
uint32_t pm_parse_color(const char *s);

typedef enum pm_scene_kind {
    PM_SCENE_RECT1 = 0,      /* BASELINE config 1: one solid-fill rectangle (x0,y0,x1,y1 in rect[]) */
    PM_SCENE_PATH_TEST = 1,  /* make_path_test, src/lib.rs:273-284 */
    PM_SCENE_CARDIOID = 2,   /* make_cardioid,  src/lib.rs:257-270 */
    PM_SCENE_TIGER = 3,      /* make_tiger,     src/lib.rs:286-328 with scale = width / 200 */
    PM_SCENE_RAND_BEZIER = 4,/* BASELINE config 4: `count` random filled Bezier paths */
    PM_SCENE_GLYPHS = 5      /* BASELINE config 5: `count` small glyph-like outlines */
} pm_scene_kind;

typedef struct pm_scene_desc {
    uint32_t kind;     /* pm_scene_kind */
    uint32_t width;    /* target surface in pixels (scenes are encoded in pixel coordinates) */
    uint32_t height;
    uint32_t count;    /* RAND_BEZIER / GLYPHS: number of paths (0 = the config's default) */
    uint64_t seed;     /* SplitMix64 seed (0 = the config's default) */
    double scale;      /* TIGER: 0 = width/200; CARDIOID: 0 = 1 (coordinates as in the reference) */
    double rect[4];    /* RECT1 */
    uint32_t rgba;     /* RECT1 colour, 0xRRGGBBAA (0 = 0x3366ccff) */
    uint32_t options;  /* PM_SCENE_OPT_* */
} pm_scene_desc;
enum {
    PM_SCENE_OPT_COMPOUND_FILLS = 1u << 0, /* TIGER: one Fill item per <path> (pm_encoder_fill_subpaths, nonzero rule)
                                              instead of one per subpath (src/lib.rs:342-347): holes are cut out */
    PM_SCENE_OPT_EVEN_ODD = 1u << 1        /* ... with PM_FILL_EVEN_ODD in the items' flags */
};

/* Writes the scene into buf; returns bytes written, or a negative pm_status.  With buf == NULL it
 * returns the size needed. */
int64_t pm_scene_build(const pm_scene_desc *desc, uint8_t *buf, size_t cap);
/* Same as make_tiger but for an arbitrary "path list" text (see tools/make_tiger_fixture.py). */
int64_t pm_scene_from_pathlist(const char *text, size_t len, double scale, uint8_t *buf, size_t cap);
/* Bounds-checks every ref/count of an encoded scene (the reference never does).  items_ix and every points_ix
 * must be multiples of 8 -- what the reference's encoder produces (src/lib.rs:132-163, :224-240); the kernels read
 * points with 64-bit loads. */
int pm_scene_validate(const uint8_t *scene, size_t len);

/* Multi-GPU row-strip shard (host only).  The reference is single-device; tiles are independent
 * given the scene (TestApp/PietRender.metal:167-170, :463-466), so N GPUs take N contiguous strips
 * of tile rows.  pm_scene_row_costs estimates the relative cost of every tile row of the frame
 * (n_rows >= ceil(height / 16) entries); pm_balance_strips cuts the rows into n_parts contiguous,
 * non-empty strips whose largest cost is minimal: strip g = rows [bounds[g], bounds[g+1]),
 * bounds has n_parts + 1 entries.  Both are deterministic, so every rank can compute its own strip
 * from the broadcast scene without a further collective. */
int pm_scene_row_costs(const uint8_t *scene, size_t len, uint32_t width, uint32_t height,
                       float *cost, size_t n_rows);
int pm_balance_strips(const float *cost, uint32_t n_rows, uint32_t n_parts, uint32_t *bounds);

/* Framebuffer egress (host only; SURVEY.md 8(f) rank 4: the step after the hot path).  The reference hands
 * its texture to MTKView (TestApp/PietRenderer.m:90-101); a headless renderer writes a file: binary PPM
 * (alpha dropped) or PNG (8-bit RGBA, stored deflate blocks).  rgba8 as returned by pm_renderer_read_rgba8. */
int pm_write_ppm(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height, size_t stride);
int pm_write_png(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height, size_t stride);

/* ------------------------------------------------------------------------------------------- */
/* Renderer (CUDA, sm_100a)                                                                     */
/* ------------------------------------------------------------------------------------------- */

typedef struct pm_renderer pm_renderer;

enum {
    PM_FLAG_FIX_POLY_PRECULL = 1u << 0, /* use the per-row polyline pre-cull instead of the
                                           reference's lane-row one */
    PM_FLAG_EXACT_SRGB = 1u << 1,       /* powf() for the linear->sRGB encode instead of ex2/lg2 */
    PM_FLAG_FILL_RULES = 1u << 2        /* honour PietFill.flags (the word the reference reserves "for winding rule",
                                           TestApp/SceneEncoder.h:44, and never reads): bit 0 = PM_FILL_EVEN_ODD fills the
                                           item by the even-odd rule, with the formula the reference names but leaves
                                           out (TestApp/PietRender.metal:538-540).  Off: the word is ignored, as upstream */
};
enum { PM_FILL_NONZERO = 0u, PM_FILL_EVEN_ODD = 1u };  /* PietFill.flags */

typedef struct pm_config {
    int32_t device;          /* CUDA device ordinal */
    uint32_t flags;
    uint64_t scratch_bytes;  /* overflow part of the record pool (beyond the 16 inline slots per
                                tile); 0 = sized automatically; it grows on demand either way */
} pm_config;

typedef struct pm_frame_stats {
    float ms_total;          /* device time of the last frame (CUDA events on the render stream) */
    float ms_bin;            /*   its binning kernel */
    float ms_fine;           /*   its fill/blend kernel (k_fine: the kernel that stores the framebuffer) */
    uint32_t frames;         /* frames enqueued since the previous sync that the sums below cover */
    float ms_total_sum;      /* the same three times summed over those frames */
    float ms_bin_sum;
    float ms_fine_sum;
    uint32_t n_tiles;        /* tiles in the strip */
    uint32_t n_overflow_records; /* pool records used beyond the 16 inline slots per tile (last frame) */
    uint32_t n_complex_tiles;/* tiles that own at least one record (last frame) */
    uint32_t n_launches;     /* kernels launched per frame */
    uint32_t retries;        /* re-renders after growing the record pool */
    float ms_heavy;          /* device time of the heavy-tile kernel of the last frame (it is part of neither ms_bin nor
                                ms_fine; with frame events off it runs beside the fill/blend kernel) */
    float ms_heavy_sum;
    uint32_t n_heavy_tiles;  /* tiles with more records than inline slots (last frame) */
    float ms_plan;           /* host wall time of the last per-scene plan (once per scene / size / strip; it is part of
                                the first frame after set_scene and of every pm_renderer_render_host call) */
} pm_frame_stats;

int pm_renderer_create(pm_renderer **out, const pm_config *cfg);
void pm_renderer_destroy(pm_renderer *r);

/* Surface size in pixels.  The framebuffer is RGBA8 (bytes R,G,B,A), row-major, top-left origin.
 * By default the renderer owns the whole frame; pm_renderer_set_strip restricts it to the
 * contiguous tile rows [tile_y0, tile_y1) for the multi-GPU row-strip shard. */
int pm_renderer_resize(pm_renderer *r, uint32_t width, uint32_t height);
int pm_renderer_set_strip(pm_renderer *r, uint32_t tile_y0, uint32_t tile_y1);

/* Upload (and validate) an encoded scene from host memory. */
int pm_renderer_set_scene(pm_renderer *r, const uint8_t *scene, size_t len);
/* Adopt a scene that already is in this device's memory (e.g. the output of an NCCL broadcast);
 * it is copied device-to-device on the render stream and validated on the device. */
int pm_renderer_set_scene_device(pm_renderer *r, const void *scene_dev, size_t len);

/* Flattening and scene encoding ON THE DEVICE (the step in front of the hot path: src/flatten.rs:10-47 + Encoder::fill /
 * polyline, src/lib.rs:195-240, which the reference runs on the CPU).  The caller hands over path control points as they
 * are -- one item per subpath, in painter's order, exactly what make_tiger produces after BezPath::from_svg (lib.rs:296-323)
 * -- and the renderer builds the encoded scene in its own scene buffer: every coordinate times `scale` (Affine::scale,
 * lib.rs:297), cubics cut into n = max(1, ceil((|3 p2 - p3 - 3 p1 + p0|^2 / (432 (tolerance / 100)^2))^(1/6))) uniform
 * steps in f64, points narrowed to f32, u16 bounding boxes, PietFill / PietStrokePolyLine items.  Host arrays; the
 * renderer copies them.  pm_renderer_read_scene copies the scene the renderer holds (however it was set) back out. */
enum { PM_VERB_LINE = 0u, PM_VERB_CURVE = 1u };
typedef struct pm_path_set {
    uint32_t n_subpaths;            /* one item per subpath */
    uint32_t n_segments;            /* LineTo / CurveTo elements of all subpaths */
    const uint32_t *first_segment;  /* n_subpaths + 1 entries: subpath i owns segments [first_segment[i], first_segment[i+1]) */
    const double *start;            /* n_subpaths x 2: the subpath's MoveTo point */
    const uint8_t *verb;            /* n_segments: PM_VERB_LINE / PM_VERB_CURVE */
    const double *ctrl;             /* n_segments x 6: CurveTo c1.x c1.y c2.x c2.y end.x end.y; LineTo: its end point in the last two */
    const uint32_t *tag;            /* n_subpaths: 3 (PietFill) or 4 (PietStrokePolyLine) */
    const uint32_t *rgba;           /* n_subpaths: 0xRRGGBBAA */
    const float *width;             /* n_subpaths: stroke width (read for tag 4) */
    const uint32_t *flags;          /* n_subpaths or NULL: PietFill.flags */
} pm_path_set;
int pm_renderer_set_scene_paths(pm_renderer *r, const pm_path_set *paths, double scale, double tolerance);
int pm_renderer_read_scene(pm_renderer *r, uint8_t *dst, size_t cap, size_t *len);

/* Enqueue one frame on the renderer's stream (asynchronous, like drawInMTKView's commit). */
int pm_renderer_render(pm_renderer *r);
/* Per-frame CUDA events (the ms_* fields of pm_frame_stats):
 *   1 (default)  events around binning, the heavy-tile kernel and the fill/blend kernel: every kernel is timed, the
 *                kernels of a frame run one after the other;
 *   2            events around the frame only (ms_total): inside the frame the kernels are chained by programmatic
 *                dependent launch -- each kernel's launch and prologue overlap the tail of the one before it, and the
 *                heavy-tile kernel runs beside the fill/blend kernel;
 *   0            no events: consecutive frames are chained the same way, like the reference's back-to-back command
 *                buffers (TestApp/PietRenderer.m:59-103 commits and never waits); pm_frame_stats reports frames = 0. */
int pm_renderer_set_frame_events(pm_renderer *r, int enabled);
/* Wait for the stream; report the last frame's statistics (stats may be NULL). */
int pm_renderer_sync(pm_renderer *r, pm_frame_stats *stats);
/* Copy the strip's pixels to host memory: rows [16*tile_y0, min(16*tile_y1, height)),
 * `stride` bytes between rows (>= 4*width). */
int pm_renderer_read_rgba8(pm_renderer *r, uint8_t *dst, size_t stride);
/* One call from host scene bytes to host pixels: upload, render, read back, synchronise. */
int pm_renderer_render_host(pm_renderer *r, const uint8_t *scene, size_t len, uint8_t *dst,
                            size_t stride, pm_frame_stats *stats);

/* Device pointer / pitch of the strip's framebuffer and the stream the frame is enqueued on
 * (cudaStream_t), for zero-copy consumers such as a torch tensor view.  A zero-copy consumer must call
 * pm_renderer_sync() before it uses the pixels of a frame: that is where a frame whose records did not fit the
 * record pool is detected (the pool grows and the frame is rendered again; pm_frame_stats.retries counts it) --
 * ordering work on the stream alone would see such a frame with records missing. */
int pm_renderer_framebuffer(pm_renderer *r, void **dev_ptr, size_t *pitch_bytes, uint32_t *rows);
int pm_renderer_stream(pm_renderer *r, void **cuda_stream);

/* Debug / parity read-backs (never on the timed path). */
/* Re-render the strip with fp32 RGBA output (4 floats per pixel, alpha = 1) and copy it out. */
int pm_renderer_read_rgba32f(pm_renderer *r, float *dst, size_t stride_bytes);

typedef struct pm_tile_item {
    uint32_t item;      /* item index */
    int32_t backdrop;   /* Fill items: integer backdrop handed to DrawFill; 0 otherwise */
    uint32_t effect;    /* 0 = draw, 1 = solid */
} pm_tile_item;
/* Per-tile item lists of the last frame, in the definition of SURVEY.md 8(a): for tile t (row-major
 * within the strip) items[offsets[t] .. offsets[t+1]) ascending, truncated at the last opaque solid
 * cover, plus the tile's final solid colour (0 if the tile is not solid).  offsets has n_tiles+1
 * entries.  Returns PM_ERR_BUFFER_TOO_SMALL (with *n_items_out set) if cap_items is too small. */
int pm_renderer_read_tile_items(pm_renderer *r, uint32_t *offsets, pm_tile_item *items,
                                size_t cap_items, size_t *n_items_out, uint32_t *solid_colors);

/* ------------------------------------------------------------------------------------------- */
/* Multi-GPU group (one host thread, the GPUs of one box)                                       */
/* ------------------------------------------------------------------------------------------- */
/* The reference renders on one device (TestApp/ViewController.m:16, TestApp/PietRenderer.m:48) and its
 * -[PietRenderer initScene] (PietRenderer.m:203-205) fills that device's scene buffer.  A group is the same
 * four-call life cycle over N GPUs: the frame's tile rows are cut into N contiguous, cost-balanced row-strips;
 * pm_group_set_scene uploads the scene once and broadcasts it with ONE ncclBroadcast over NVLink (communicators from
 * ncclCommInitAll: no launcher, no rendezvous; libnccl.so.2 is loaded with dlopen on first use); pm_group_render
 * enqueues the frame on every GPU with no collective and no host synchronisation.  devices == NULL means 0..n-1. */
typedef struct pm_group pm_group;
int pm_group_create(pm_group **out, const int32_t *devices, uint32_t n_devices, uint32_t flags);
void pm_group_destroy(pm_group *g);
uint32_t pm_group_size(const pm_group *g);
int pm_group_member(pm_group *g, uint32_t index, pm_renderer **out);   /* the renderer that owns strip `index` */
int pm_group_resize(pm_group *g, uint32_t width, uint32_t height);
int pm_group_set_scene(pm_group *g, const uint8_t *scene, size_t len); /* host scene -> device 0 -> ncclBroadcast -> validate + plan */
int pm_group_strip_bounds(const pm_group *g, uint32_t *bounds, size_t cap); /* n + 1 tile-row bounds */
int pm_group_set_frame_events(pm_group *g, int mode);                  /* pm_renderer_set_frame_events on every member */
int pm_group_render(pm_group *g);
/* Wait for every member; stats (optional, n_stats entries) per member; *ms_frame_max = the slowest member's last frame. */
int pm_group_sync(pm_group *g, pm_frame_stats *stats, size_t n_stats, float *ms_frame_max);
/* Off the hot path: the whole frame to host memory (row `stride` >= 4 * width), or gathered into one buffer on member
 * `root`'s device (strips are of unequal height: ncclSend / ncclRecv, not an all-gather); *pitch_bytes = 64 * ceil(width / 16). */
int pm_group_read_rgba8(pm_group *g, uint8_t *dst, size_t stride);
int pm_group_gather_device(pm_group *g, uint32_t root, void **dev_ptr, size_t *pitch_bytes);
/* NCCL_VERSION_CODE of the library that was loaded, 0 if none could be. */
int pm_group_nccl_version(void);

/* ------------------------------------------------------------------------------------------- */
/* RenderContext facade (host side; the drawing is flattened and encoded on the device)          */
/* ------------------------------------------------------------------------------------------- */
/* The piet calls the reference's README aspires to (README.md:3) and its `Encoder` stops short of (src/lib.rs:165-222):
 * clear, transform / save / restore, fill, fill_even_odd, stroke with a solid colour (0xRRGGBBAA), finish.  Paths are
 * BezPath-like element lists; quadratics are raised to cubics; a filled path with several subpaths becomes one item
 * (holes are holes); strokes follow make_tiger: one polyline item per subpath, thin-stroke rule (lib.rs:353-362).
 * finish() installs the drawing as the renderer's scene (pm_renderer_set_scene_paths) and empties the context;
 * render and read back with pm_renderer_render / pm_renderer_read_rgba8 as usual. */
enum { PM_EL_MOVE = 0u, PM_EL_LINE = 1u, PM_EL_QUAD = 2u, PM_EL_CURVE = 3u, PM_EL_CLOSE = 4u };
typedef struct pm_path_el {
    uint32_t verb;   /* PM_EL_* */
    uint32_t pad;
    double x[6];     /* MOVE / LINE: x y; QUAD: cx cy x y; CURVE: c1x c1y c2x c2y x y */
} pm_path_el;
typedef struct pm_context pm_context;
int pm_context_new(pm_context **out, pm_renderer *renderer, uint32_t width, uint32_t height);
void pm_context_free(pm_context *c);
int pm_context_save(pm_context *c);
int pm_context_restore(pm_context *c);
int pm_context_transform(pm_context *c, const double affine[6]);   /* kurbo order [a b c d e f]: x' = a x + c y + e */
int pm_context_clear(pm_context *c, uint32_t rgba);
int pm_context_fill(pm_context *c, const pm_path_el *els, size_t n, uint32_t rgba);
int pm_context_fill_even_odd(pm_context *c, const pm_path_el *els, size_t n, uint32_t rgba);
int pm_context_stroke(pm_context *c, const pm_path_el *els, size_t n, uint32_t rgba, double width);
uint32_t pm_context_item_count(const pm_context *c);
int pm_context_path_set(pm_context *c, pm_path_set *out);          /* the recorded drawing (pointers into the context) */
int pm_context_finish(pm_context *c, double tolerance);            /* tolerance <= 0: the reference's 0.1 (lib.rs:330) */

/* Pinned host memory for fast host<->device copies of scenes and frames. */
int pm_host_alloc(void **out, size_t bytes);
void pm_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* PIET_METAL_B200_H */
