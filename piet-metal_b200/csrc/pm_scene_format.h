// Scene-buffer wire format shared by the feed (host C++) and the CUDA kernels.
//
// This is the encoding the reference's Rust `Encoder` writes (src/lib.rs:15-77, 132-163, 224-240)
// and its Metal kernels read through TestApp/GenTypes.h (readers :49-57, :119-138, :193-209,
// :257-273, :317-328).  Everything is little-endian; a "ref" is a u32 byte offset from the start of
// the scene buffer.  Layout:
//
//   0            u32  n_items
//   4            u32  items_ix            (= 8 + 8*n_items for a group at offset 0)
//   8            u16[4] x n_items         bbox (x0,y0,x1,y1): floor/floor/ceil/ceil, clamped 0..65535
//   items_ix     32 B x n_items           item union, tag in the first word
//   ...          (f32,f32) arrays         point lists referenced by Fill / Poly items
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif

// Tile geometry: TestApp/PietShaderTypes.h:17-22.  The reference's 4096-px / 170-command caps
// (PietShaderTypes.h:24-32) are deliberately not carried over.
#define PM_TILE_W 16
#define PM_TILE_H 16
#define PM_GROUP_TILES_X 16   /* tilerGroupWidth  */
#define PM_GROUP_TILES_Y 2    /* tilerGroupHeight */
#define PM_STRIP_PX (PM_GROUP_TILES_X * PM_TILE_W)   /* 256: x extent of a tiler group */
#define PM_GROUP_PX_Y (PM_GROUP_TILES_Y * PM_TILE_H) /* 32 */

// Item tags: TestApp/GenTypes.h:325-328 ("manually fixed up"), src/lib.rs:70-77.
enum { PM_ITEM_CIRCLE = 1, PM_ITEM_LINE = 2, PM_ITEM_FILL = 3, PM_ITEM_POLY = 4 };

#define PM_GROUP_HEADER_SIZE 8
#define PM_BBOX_SIZE 8
#define PM_ITEM_SIZE 32

// Byte offsets inside a 32-byte item (GenTypes.h:110-138, 185-209, 249-273).
#define PM_LINE_FLAGS 4
#define PM_LINE_RGBA 8
#define PM_LINE_WIDTH 12
#define PM_LINE_START 16
#define PM_LINE_END 24
#define PM_FILL_FLAGS 4
#define PM_FILL_RGBA 8
#define PM_FILL_NPOINTS 12
#define PM_FILL_POINTS_IX 16
#define PM_POLY_RGBA 4
#define PM_POLY_WIDTH 8
#define PM_POLY_NPOINTS 12
#define PM_POLY_POINTS_IX 16

typedef struct { uint32_t n_items; uint32_t items_ix; } pm_group_header;
typedef struct { uint16_t x0, y0, x1, y1; } pm_bbox;
typedef struct { uint32_t tag, flags, rgba; float width; float sx, sy, ex, ey; } pm_item_line;
typedef struct { uint32_t tag, flags, rgba, n_points, points_ix; uint32_t pad[3]; } pm_item_fill;
typedef struct { uint32_t tag, rgba; float width; uint32_t n_points, points_ix; uint32_t pad[3]; } pm_item_poly;
typedef struct { uint32_t tag; uint32_t body[7]; } pm_item_any;

#if defined(__cplusplus)
static_assert(sizeof(pm_group_header) == PM_GROUP_HEADER_SIZE, "SimpleGroup header is 8 bytes (lib.rs:15-20)");
static_assert(sizeof(pm_bbox) == PM_BBOX_SIZE, "ShortBbox is 4 x u16 (lib.rs:22-24)");
static_assert(sizeof(pm_item_line) == PM_ITEM_SIZE, "PietStrokeLine fills the 32-byte union");
static_assert(sizeof(pm_item_fill) == PM_ITEM_SIZE, "item slot is 32 bytes (GenTypes.h:323)");
static_assert(sizeof(pm_item_poly) == PM_ITEM_SIZE, "item slot is 32 bytes (GenTypes.h:323)");
static_assert(sizeof(pm_item_any) == PM_ITEM_SIZE, "PietItem is tag + 7 words (GenTypes.h:313-316)");
static_assert(offsetof(pm_item_line, rgba) == PM_LINE_RGBA && offsetof(pm_item_line, width) == PM_LINE_WIDTH &&
              offsetof(pm_item_line, sx) == PM_LINE_START && offsetof(pm_item_line, ex) == PM_LINE_END, "PietStrokeLine offsets");
static_assert(offsetof(pm_item_fill, rgba) == PM_FILL_RGBA && offsetof(pm_item_fill, n_points) == PM_FILL_NPOINTS &&
              offsetof(pm_item_fill, points_ix) == PM_FILL_POINTS_IX, "PietFill offsets");
static_assert(offsetof(pm_item_poly, rgba) == PM_POLY_RGBA && offsetof(pm_item_poly, width) == PM_POLY_WIDTH &&
              offsetof(pm_item_poly, n_points) == PM_POLY_NPOINTS && offsetof(pm_item_poly, points_ix) == PM_POLY_POINTS_IX, "PietStrokePolyLine offsets");
#endif
