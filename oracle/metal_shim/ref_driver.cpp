// ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs the reference's own shaders on the CPU: this translation unit #includes
// /root/reference/TestApp/PietRender.metal UNMODIFIED (through the metal_stdlib stand-in next to this
// file) and plays the part of the Metal runtime around it, following the reference's host code
// TestApp/PietRenderer.m:
//   :59-78    tileKernel: threadgroups of tilerGroupWidth x tilerGroupHeight = 16 x 2 threads over
//             ceil(nTilesX / 16) x ceil(nTilesY / 2) groups; buffer(0) = scene, buffer(1) = tiles
//             (maxTilesWidth * maxTilesHeight * tileBufSize bytes, :50-54), texture(0) = loTexture
//   :80-88    renderKernel: threadgroups of 16 x 16 threads, one per tile; texture(0) = the W x H
//             pixel texture, buffer(0) = tiles
//   :90-99    one point sprite per tile (vertex buffer of :125-143): vertexShader reads the tile's
//             solid colour, fragmentShader picks it or the pixel texture
// The 32 threads of a tiler threadgroup are cooperative fibers (ucontext) on one OS thread;
// threadgroup_barrier() switches to the next fiber, so every thread reaches barrier k before any
// thread leaves it, and `threadgroup` variables are thread_local statics shared by those fibers.
// Threadgroups are independent and are spread over OpenMP threads.
//
// Limits inherited from the unmodified source: surfaces up to 4096 x 4096 (maxTilesWidth/Height 256,
// PietShaderTypes.h:31-32) and 4096 bytes = 170 commands per tile with NO overflow check
// (PietShaderTypes.h:27, PietRender.metal:82-83): a tile that needs more writes into its right-hand
// neighbour's buffer.  pmref_render reports the command count of every tile (-1: no terminator
// inside the tile's 4096 bytes) so that the tests can keep such tiles out of a comparison.
//
// Textures are BGRA8Unorm in the reference (PietRenderer.m:29, :112): the stand-in keeps the fp32
// values the shaders wrote and the driver applies the unorm8 write rule (round to nearest of
// clamp(v, 0, 1) * 255) when it produces the RGBA8 frame, which it stores R, G, B, A.
#include <metal_stdlib>
using namespace metal;

#include <ucontext.h>

#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// ---- the reference's shader source, as it is ----
#include "PietRender.metal"

namespace {

const int kGroupThreads = tilerGroupWidth * tilerGroupHeight;  // 32
const size_t kFiberStack = 256 * 1024;

struct TileJob {
    const char *scene;
    char *tiles;
    texture2d<half, access::write> lo;
    uint group_x, group_y;
};

struct FiberSet {
    ucontext_t main_ctx;
    ucontext_t ctx[kGroupThreads];
    std::vector<char> stacks;
    bool done[kGroupThreads];
    int current;
    const TileJob *job;
};

thread_local FiberSet *t_fibers = nullptr;

void fiber_entry() {
    FiberSet *fs = t_fibers;
    const int t = fs->current;
    const TileJob *j = fs->job;
    const uint lx = (uint)t % tilerGroupWidth, ly = (uint)t / tilerGroupWidth;
    // thread_position_in_grid / thread_index_in_threadgroup of a 16 x 2 threadgroup
    tileKernel(j->scene, j->tiles, j->lo, uint2(j->group_x * tilerGroupWidth + lx, j->group_y * tilerGroupHeight + ly), (uint)t);
    fs->done[t] = true;
    swapcontext(&fs->ctx[t], &fs->main_ctx);
}

void run_tile_group(FiberSet *fs, const TileJob *job) {
    fs->job = job;
    for (int t = 0; t < kGroupThreads; t++) {
        getcontext(&fs->ctx[t]);
        fs->ctx[t].uc_stack.ss_sp = fs->stacks.data() + (size_t)t * kFiberStack;
        fs->ctx[t].uc_stack.ss_size = kFiberStack;
        fs->ctx[t].uc_link = &fs->main_ctx;
        makecontext(&fs->ctx[t], fiber_entry, 0);
        fs->done[t] = false;
    }
    // round-robin: every live fiber runs up to its next barrier (or its end), then the next one
    for (;;) {
        bool any = false;
        for (int t = 0; t < kGroupThreads; t++) {
            if (fs->done[t]) continue;
            any = true;
            fs->current = t;
            swapcontext(&fs->main_ctx, &fs->ctx[t]);
        }
        if (!any) break;
    }
}

uint8_t unorm8(float v) { return (uint8_t)lrintf(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); }

}  // namespace

extern "C" void pmref_threadgroup_barrier(void) {
    FiberSet *fs = t_fibers;
    swapcontext(&fs->ctx[fs->current], &fs->main_ctx);
}

extern "C" {

// 1 if `half` is fp32 in this build, 0 if it is a 16-bit float
int pmref_half_is_float(void) { return sizeof(half) == 4; }
int pmref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
uint32_t pmref_tile_buf_size(void) { return tileBufSize; }

// Renders the scene the way PietRenderer.m drives the three passes.
//   rgba8    H x W x 4 bytes (R, G, B, A), may be NULL: the composited frame
//   rgba32f  H x W x 4 floats, may be NULL: the same before the unorm8 write rule
//   solid    nTilesY x nTilesX words, may be NULL: tileKernel's per-tile solidColor (loTexture, as bytes)
//   n_cmds   nTilesY x nTilesX, may be NULL: commands in the tile's list including End / Bail; -1 if none was found
//   cmds     nTilesY x nTilesX x tileBufSize bytes, may be NULL: the raw per-tile command buffers
// returns 0, or -1 for a surface beyond maxTilesWidth x maxTilesHeight tiles / bad arguments
int pmref_render(const uint8_t *scene, size_t scene_len, uint32_t width, uint32_t height, int threads, uint8_t *rgba8, float *rgba32f,
                 uint32_t *solid, int32_t *n_cmds, uint8_t *cmds) {
    if (!scene || scene_len < 8 || width == 0 || height == 0) return -1;
    const uint nTilesX = (width + tileWidth - 1) / tileWidth, nTilesY = (height + tileHeight - 1) / tileHeight;  // PietRenderer.m:63-64
    if (nTilesX > maxTilesWidth || nTilesY > maxTilesHeight) return -1;
    const uint nTilerGroupsX = (nTilesX + tilerGroupWidth - 1) / tilerGroupWidth;                                  // :66-67
    const uint nTilerGroupsY = (nTilesY + tilerGroupHeight - 1) / tilerGroupHeight;
    // _tileBuf (:50,:54) -- only the rows the dispatch can touch, zero-filled
    const size_t tile_rows = (size_t)nTilerGroupsY * tilerGroupHeight;
    std::vector<char> tiles(tile_rows * maxTilesWidth * tileBufSize + tileBufSize, 0);
    std::vector<float> lo_texels((size_t)nTilesX * nTilesY * 4, 0.0f), px_texels((size_t)width * height * 4, 0.0f);
    pm_texture_store lo_store{lo_texels.data(), nTilesX, nTilesY}, px_store{px_texels.data(), width, height};
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif

    // pass 1: tileKernel
    #pragma omp parallel
    {
        FiberSet fs;
        fs.stacks.resize((size_t)kGroupThreads * kFiberStack);
        t_fibers = &fs;
        #pragma omp for schedule(dynamic, 1) collapse(2)
        for (uint gy = 0; gy < nTilerGroupsY; gy++)
            for (uint gx = 0; gx < nTilerGroupsX; gx++) {
                TileJob job{(const char *)scene, tiles.data(), texture2d<half, access::write>{&lo_store}, gx, gy};
                run_tile_group(&fs, &job);
            }
        t_fibers = nullptr;
    }

    // pass 2: renderKernel, one 16 x 16 threadgroup per tile (no barriers in it: a plain loop)
    #pragma omp parallel for schedule(dynamic, 1)
    for (uint ty = 0; ty < nTilesY; ty++)
        for (uint tx = 0; tx < nTilesX; tx++)
            for (uint y = 0; y < tileHeight; y++)
                for (uint x = 0; x < tileWidth; x++)
                    renderKernel(texture2d<half, access::write>{&px_store}, tiles.data(), uint2(tx * tileWidth + x, ty * tileHeight + y), uint2(tx, ty));

    // pass 3: point sprites (vertex buffer as PietRenderer.m:125-143 builds it)
    const float scaleX = 2.0f / (float)width, scaleY = 2.0f / (float)height;
    std::vector<RenderVertex> vertices((size_t)nTilesX * nTilesY);
    {
        size_t ix = 0;
        for (uint y = 0; y < nTilesY; y++)
            for (uint x = 0; x < nTilesX; x++) {
                RenderVertex rv;
                const uint x0 = x * tileWidth + (tileWidth / 2), y0 = y * tileHeight + (tileHeight / 2);
                rv.position.x = x0 * scaleX - 1.0f;
                rv.position.y = y0 * -scaleY + 1.0f;
                rv.textureCoordinate.x = x0;
                rv.textureCoordinate.y = y0;
                vertices[ix++] = rv;
            }
    }
    #pragma omp parallel for schedule(dynamic, 1)
    for (uint ty = 0; ty < nTilesY; ty++)
        for (uint tx = 0; tx < nTilesX; tx++) {
            const uint vid = ty * nTilesX + tx;
            const RenderData rd = vertexShader(vid, vertices.data(), texture2d<half>{&lo_store});
            // the rasteriser: a pointSize-wide square around the vertex; fragment position = pixel centre
            const int psz = (int)rd.pointSize;
            const int cx = (int)vertices[vid].textureCoordinate.x, cy = (int)vertices[vid].textureCoordinate.y;
            for (int py = cy - psz / 2; py < cy + psz / 2; py++)
                for (int px = cx - psz / 2; px < cx + psz / 2; px++) {
                    if (px < 0 || py < 0 || px >= (int)width || py >= (int)height) continue;
                    RenderData in = rd;
                    in.clipSpacePosition = float4((float)px + 0.5f, (float)py + 0.5f, 0.0f, 1.0f);
                    const half4 c = fragmentShader(in, texture2d<half>{&px_store});
                    const size_t o = ((size_t)py * width + (size_t)px) * 4;
                    const float v[4] = {(float)c.x, (float)c.y, (float)c.z, (float)c.w};
                    for (int k = 0; k < 4; k++) {
                        if (rgba32f) rgba32f[o + k] = v[k];
                        if (rgba8) rgba8[o + k] = unorm8(v[k]);
                    }
                }
        }

    for (uint ty = 0; ty < nTilesY; ty++)
        for (uint tx = 0; tx < nTilesX; tx++) {
            const size_t t = (size_t)ty * nTilesX + tx;
            const char *src = tiles.data() + ((size_t)ty * maxTilesWidth + tx) * tileBufSize;
            if (solid) {
                const float *p = &lo_texels[t * 4];
                solid[t] = (uint32_t)unorm8(p[0]) | ((uint32_t)unorm8(p[1]) << 8) | ((uint32_t)unorm8(p[2]) << 16) | ((uint32_t)unorm8(p[3]) << 24);
            }
            if (n_cmds) {
                int32_t n = -1;
                for (uint k = 0; k < tileBufSize / sizeof(Cmd); k++) {
                    uint tag;
                    memcpy(&tag, src + k * sizeof(Cmd), 4);
                    if (tag == Cmd_End || tag == Cmd_Bail) { n = (int32_t)k + 1; break; }
                    if (tag < Cmd_End || tag > Cmd_Bail) break;
                }
                n_cmds[t] = n;
            }
            if (cmds) memcpy(cmds + t * tileBufSize, src, tileBufSize);
        }
    return 0;
}

}  // extern "C"
