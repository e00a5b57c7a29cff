// Flattening and scene encoding on the device (sm_100a): the step in front of the hot path.
//
// The reference flattens and encodes on the CPU: flatten_path (src/flatten.rs:10-47) walks a path's elements --
// MoveTo starts a subpath, LineTo adds a point, CurveTo goes through kurbo's CubicBez::to_quads(tolerance * 1e-2) of
// which only the end points are kept, i.e. n = max(1, ceil((|3 p2 - p3 - 3 p1 + p0|^2 / (432 acc^2))^(1/6))) uniform
// parameter steps -- and Encoder::fill / polyline (src/lib.rs:195-240) write one item per subpath: points narrowed
// to f32, bounding box floor/floor/ceil/ceil clamped to u16 (lib.rs:88-97).  At 100 k paths that is what an
// end-to-end frame spends its time on.  Here the control points are uploaded as they are and four small kernels
// produce the encoded scene -- byte for byte the layout of SURVEY.md 2.2 -- directly in the renderer's scene buffer:
//   k_flat_count   one thread per path segment: points it will emit (f64, the reference's formula and operand order)
//   k_flat_scan    exclusive prefix of the counts (one CTA, every thread a contiguous chunk)
//   k_flat_emit    one thread per segment: its points, narrowed to f32, at their final place; per-subpath bounding box
//                  in f64 with ordered-integer atomics
//   k_flat_items   one thread per subpath: u16 bbox and the PietFill / PietStrokePolyLine item; thread 0: the header
// All arithmetic is f64 like kurbo's; the only difference to the CPU feed (pm_feed.cpp) a test can see is pow():
// CUDA's differs from glibc's in the last bit at most, which moves a point count only if the sixth root lands within
// an ulp of an integer.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "pm_kernels.h"
#include "pm_scene_format.h"

namespace {

typedef unsigned long long u64;
typedef long long i64;

// total order of doubles as signed 64-bit integers (for atomicMin / atomicMax)
__device__ __forceinline__ i64 dkey(double d) {
    i64 b = __double_as_longlong(d);
    return b < 0 ? b ^ 0x7fffffffffffffffll : b;
}
__device__ __forceinline__ double dunkey(i64 k) { return __longlong_as_double(k < 0 ? k ^ 0x7fffffffffffffffll : k); }

__device__ __forceinline__ uint32_t subpath_of(const uint32_t *first, uint32_t n_sub, uint32_t s) {
    uint32_t lo = 0, hi = n_sub;  // largest i with first[i] <= s  (first[n_sub] = n_segments > s)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (first[mid] <= s) lo = mid; else hi = mid;
    }
    // empty subpaths share their first[] with the next one: the segment belongs to the last of them
    while (lo + 1 < n_sub && first[lo + 1] <= s) lo++;
    return lo;
}

struct Seg { double p0x, p0y, p1x, p1y, p2x, p2y, p3x, p3y; bool curve; };

// the segment's control points after Affine::scale (lib.rs:297, :314: every coordinate times s)
__device__ __forceinline__ Seg load_seg(const PmPathSetDev &P, uint32_t s, uint32_t sub, double scale) {
    Seg g;
    const double *c = P.ctrl + 6 * (size_t)s;
    if (s == P.first[sub]) { g.p0x = P.start[2 * (size_t)sub] * scale; g.p0y = P.start[2 * (size_t)sub + 1] * scale; }
    else { g.p0x = c[-2] * scale; g.p0y = c[-1] * scale; }  // the end point of the segment before
    g.p1x = c[0] * scale; g.p1y = c[1] * scale; g.p2x = c[2] * scale; g.p2y = c[3] * scale; g.p3x = c[4] * scale; g.p3y = c[5] * scale;
    g.curve = P.verb[s] != 0;
    return g;
}

__device__ __forceinline__ uint32_t seg_points(const Seg &g, double tolerance) {
    if (!g.curve) return 1u;
    const double acc = tolerance * 1e-2;  // flatten.rs:35
    const double max_hypot2 = 432.0 * acc * acc;
    const double ex = (3.0 * g.p2x - g.p3x) - (3.0 * g.p1x - g.p0x);
    const double ey = (3.0 * g.p2y - g.p3y) - (3.0 * g.p1y - g.p0y);
    const double err = ex * ex + ey * ey;
    const double nf = fmax(1.0, ceil(pow(err / max_hypot2, 1.0 / 6.0)));
    return nf < 1.0e6 ? (uint32_t)nf : 1000000u;
}

__global__ void __launch_bounds__(256) k_flat_count(const PmPathSetDev P, double scale, double tolerance, uint32_t *cnt, i64 *bbox) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    // subpaths: the bounding box starts at the MoveTo point (encode_points, lib.rs:230-231)
    for (uint32_t i = s; i < P.n_subpaths; i += gridDim.x * blockDim.x) {
        const double x = P.start[2 * (size_t)i] * scale, y = P.start[2 * (size_t)i + 1] * scale;
        bbox[4 * (size_t)i + 0] = dkey(x); bbox[4 * (size_t)i + 1] = dkey(y);
        bbox[4 * (size_t)i + 2] = dkey(x); bbox[4 * (size_t)i + 3] = dkey(y);
    }
    if (s >= P.n_segments) return;
    const uint32_t sub = subpath_of(P.first, P.n_subpaths, s);
    cnt[s] = seg_points(load_seg(P, s, sub, scale), tolerance);
}

// cnt[s] -> exclusive prefix in place; *total = sum (saturating at 2^32 - 1)
__global__ void __launch_bounds__(1024) k_flat_scan(uint32_t *cnt, uint32_t n, u64 *total) {
    __shared__ u64 part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = tid * chunk < n ? tid * chunk : n, hi = lo + chunk < n ? lo + chunk : n;
    u64 sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += cnt[i];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) {
        u64 run = 0;
        for (uint32_t k = 0; k < blockDim.x; k++) { const u64 v = part[k]; part[k] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    u64 run = part[tid];
    for (uint32_t i = lo; i < hi; i++) { const uint32_t c = cnt[i]; cnt[i] = run > 0xffffffffull ? 0xffffffffu : (uint32_t)run; run += c; }
}

// kurbo CubicBez::eval, operand order of pm_feed.cpp's cubic_eval
__device__ __forceinline__ void cubic_eval(const Seg &g, double t, double *x, double *y) {
    const double mt = 1.0 - t;
    const double a = mt * mt * mt, b = mt * mt * 3.0, c = mt * 3.0;
    *x = g.p0x * a + (g.p1x * b + (g.p2x * c + g.p3x * t) * t) * t;
    *y = g.p0y * a + (g.p1y * b + (g.p2y * c + g.p3y * t) * t) * t;
}

__global__ void __launch_bounds__(256) k_flat_emit(const PmPathSetDev P, double scale, double tolerance, const uint32_t *off, uint8_t *scene,
                                                   uint32_t pts_base, i64 *bbox) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.n_segments) return;
    const uint32_t sub = subpath_of(P.first, P.n_subpaths, s);
    const Seg g = load_seg(P, s, sub, scale);
    // points before this segment's: those of earlier segments plus one MoveTo point per subpath up to and including this one
    float2 *dst = reinterpret_cast<float2 *>(scene + pts_base) + ((size_t)off[s] + sub + 1u);
    if (s == P.first[sub]) dst[-1] = make_float2((float)g.p0x, (float)g.p0y);  // the subpath's MoveTo point
    double x0 = g.p3x, y0 = g.p3y, x1 = g.p3x, y1 = g.p3y;
    if (!g.curve) {
        dst[0] = make_float2((float)g.p3x, (float)g.p3y);
    } else {
        const uint32_t n = seg_points(g, tolerance);
        for (uint32_t i = 0; i < n; i++) {
            const double t1 = (double)(i + 1) / (double)n;
            double x, y;
            cubic_eval(g, t1, &x, &y);
            dst[i] = make_float2((float)x, (float)y);
            x0 = fmin(x0, x); y0 = fmin(y0, y); x1 = fmax(x1, x); y1 = fmax(y1, y);
        }
    }
    atomicMin(&bbox[4 * (size_t)sub + 0], dkey(x0)); atomicMin(&bbox[4 * (size_t)sub + 1], dkey(y0));
    atomicMax(&bbox[4 * (size_t)sub + 2], dkey(x1)); atomicMax(&bbox[4 * (size_t)sub + 3], dkey(y1));
}

__device__ __forceinline__ uint16_t clamp_u16(double v) { return (uint16_t)fmin(fmax(v, 0.0), 65535.0); }  // lib.rs:91-94

__global__ void __launch_bounds__(256) k_flat_items(const PmPathSetDev P, const uint32_t *off, u64 total_points, uint8_t *scene,
                                                    uint32_t items_ix, uint32_t pts_base, const i64 *bbox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {  // SimpleGroup header (lib.rs:137-140)
        reinterpret_cast<uint32_t *>(scene)[0] = P.n_subpaths;
        reinterpret_cast<uint32_t *>(scene)[1] = items_ix;
    }
    if (i >= P.n_subpaths) return;
    const uint32_t s0 = P.first[i], s1 = P.first[i + 1];
    const u64 before = (s0 < P.n_segments ? (u64)off[s0] : total_points) + i;
    const u64 after = (s1 < P.n_segments ? (u64)off[s1] : total_points) + i + 1u;
    const uint32_t n_points = (uint32_t)(after - before);
    const uint32_t points_ix = pts_base + 8u * (uint32_t)before;
    const uint32_t tag = P.tag[i];
    double x0 = dunkey(bbox[4 * (size_t)i]), y0 = dunkey(bbox[4 * (size_t)i + 1]), x1 = dunkey(bbox[4 * (size_t)i + 2]), y1 = dunkey(bbox[4 * (size_t)i + 3]);
    const double sx = x0, sy = y0;  // (for a subpath without segments: its MoveTo point)
    const uint32_t rgba = __byte_perm(P.rgba[i], 0, 0x0123);  // rgba.to_be(): bytes R, G, B, A in memory (lib.rs:200, :213)
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tag == PM_ITEM_POLY) {  // PietStrokePolyLine (GenTypes.h:249-273); bbox inflated by half the width (lib.rs:220)
        const float width = P.width[i];
        const double hw = (double)(width * 0.5f);
        x0 -= hw; y0 -= hw; x1 += hw; y1 += hw;
        w[0] = PM_ITEM_POLY; w[1] = rgba; w[2] = __float_as_uint(width); w[3] = n_points; w[4] = points_ix;
    } else {                    // PietFill (GenTypes.h:185-209)
        w[0] = PM_ITEM_FILL; w[1] = P.flags ? P.flags[i] : 0u; w[2] = rgba; w[3] = n_points; w[4] = points_ix;
    }
    if (s0 == s1) {  // a subpath without segments: its MoveTo point is all there is (k_flat_emit had no thread for it)
        reinterpret_cast<float2 *>(scene + pts_base)[before] = make_float2((float)sx, (float)sy);
    }
    uint16_t *bb = reinterpret_cast<uint16_t *>(scene + PM_GROUP_HEADER_SIZE + (size_t)i * PM_BBOX_SIZE);
    bb[0] = clamp_u16(floor(x0)); bb[1] = clamp_u16(floor(y0)); bb[2] = clamp_u16(ceil(x1)); bb[3] = clamp_u16(ceil(y1));
    uint2 *it = reinterpret_cast<uint2 *>(scene + items_ix + (size_t)i * PM_ITEM_SIZE);  // (items_ix = 8 + 8 n: 8-byte aligned only)
    it[0] = make_uint2(w[0], w[1]);
    it[1] = make_uint2(w[2], w[3]);
    it[2] = make_uint2(w[4], 0);
    it[3] = make_uint2(0, 0);
}

}  // namespace

void pm_launch_flat_count(const PmPathSetDev &P, double scale, double tolerance, uint32_t *cnt, long long *bbox, unsigned long long *total, cudaStream_t s) {
    const uint32_t n = P.n_segments > P.n_subpaths ? P.n_segments : P.n_subpaths;
    k_flat_count<<<(n + 255) / 256 ? (n + 255) / 256 : 1, 256, 0, s>>>(P, scale, tolerance, cnt, bbox);
    k_flat_scan<<<1, 1024, 0, s>>>(cnt, P.n_segments, total);
}

void pm_launch_flat_emit(const PmPathSetDev &P, double scale, double tolerance, const uint32_t *off, unsigned long long total_points, uint8_t *scene,
                         uint32_t items_ix, uint32_t pts_base, long long *bbox, cudaStream_t s) {
    if (P.n_segments) k_flat_emit<<<(P.n_segments + 255) / 256, 256, 0, s>>>(P, scale, tolerance, off, scene, pts_base, bbox);
    k_flat_items<<<(P.n_subpaths + 255) / 256 ? (P.n_subpaths + 255) / 256 : 1, 256, 0, s>>>(P, off, total_points, scene, items_ix, pts_base, bbox);
}
