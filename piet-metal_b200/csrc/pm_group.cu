// pm_group_*: one host thread driving the GPUs of one box -- the multi-GPU face of the C ABI.
//
// The reference is single-device (one MTLDevice, one command queue: TestApp/ViewController.m:16,
// TestApp/PietRenderer.m:48); -[PietRenderer initScene] (PietRenderer.m:203-205) hands the scene to that one
// device.  Tiles are independent given the scene (TestApp/PietRender.metal:167-170, :463-466), so a group shards the
// frame's tile rows into contiguous row-strips, one per GPU:
//   * pm_group_set_scene uploads the scene ONCE, to the first device, and broadcasts it to the others over
//     NVLink/NVSwitch with one ncclBroadcast (ncclUint8, root 0); every member validates its copy on the device and
//     plans its own strip (cost-balanced bounds, computed once on the host from the scene);
//   * pm_group_render enqueues the frame on every member's stream: no collective, no host synchronisation;
//   * pm_group_read_rgba8 / pm_group_gather_device collect the strips -- off the hot path: D2H copies, or
//     ncclSend/ncclRecv of the (unequal) strips to one device.
// NCCL is loaded with dlopen when the first group is created, so that a process that only renders on one GPU -- or
// one that already carries a libnccl of its own, as PyTorch does -- is not forced to link another.  Communicators
// come from ncclCommInitAll: one process, no rendezvous, no environment variables.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/piet_metal_b200.h"
#include "pm_scene_format.h"

void pm_set_last_error(const char *text);  // pm_renderer.cu

namespace {

struct NcclApi {
    void *handle = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};
NcclApi g_nccl;

int fail(int status, const std::string &text) {
    pm_set_last_error(text.c_str());
    return status;
}

int load_nccl() {
    if (g_nccl.handle) return PM_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail(PM_ERR_NO_DEVICE, std::string("NCCL is not available: ") + dlerror());
    NcclApi a;
    a.handle = h;
#define PM_SYM(field, name) *(void **)(&a.field) = dlsym(h, name); if (!a.field) return fail(PM_ERR_NO_DEVICE, std::string("libnccl lacks ") + name)
    PM_SYM(GetErrorString, "ncclGetErrorString");
    PM_SYM(CommInitAll, "ncclCommInitAll");
    PM_SYM(CommDestroy, "ncclCommDestroy");
    PM_SYM(Broadcast, "ncclBroadcast");
    PM_SYM(GroupStart, "ncclGroupStart");
    PM_SYM(GroupEnd, "ncclGroupEnd");
    PM_SYM(Send, "ncclSend");
    PM_SYM(Recv, "ncclRecv");
    PM_SYM(GetVersion, "ncclGetVersion");
#undef PM_SYM
    g_nccl = a;
    return PM_OK;
}

}  // namespace

struct pm_group {
    std::vector<int> devices;
    std::vector<pm_renderer *> members;
    std::vector<ncclComm_t> comms;
    std::vector<uint8_t *> scene_dev;   // per device: the broadcast's landing buffer
    std::vector<size_t> scene_cap;
    std::vector<uint32_t> bounds;       // n + 1 tile-row bounds of the strips
    uint32_t width = 0, height = 0, n_ty = 0;
    bool have_scene = false;
    uint8_t *gather = nullptr;          // full frame on the gather root (pm_group_gather_device)
    size_t gather_cap = 0;
    int gather_root = -1;
};

namespace {

#define PM_G_CUDA(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(PM_ERR_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e_)); \
    } while (0)
#define PM_G_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(PM_ERR_CUDA, std::string(#call " failed: ") + g_nccl.GetErrorString(r_)); \
    } while (0)

cudaStream_t member_stream(pm_group *g, size_t i) {
    void *s = nullptr;
    pm_renderer_stream(g->members[i], &s);
    return (cudaStream_t)s;
}

}  // namespace

extern "C" {

int pm_group_create(pm_group **out, const int32_t *devices, uint32_t n_devices, uint32_t flags) {
    if (!out || n_devices == 0 || n_devices > 64) return PM_ERR_INVALID_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(PM_ERR_NO_DEVICE, "no CUDA device"); }
    pm_group *g = new (std::nothrow) pm_group();
    if (!g) return PM_ERR_NOMEM;
    for (uint32_t i = 0; i < n_devices; i++) {
        const int d = devices ? devices[i] : (int)i;
        if (d < 0 || d >= n_dev) { delete g; return PM_ERR_INVALID_ARG; }
        for (int e : g->devices) if (e == d) { delete g; return fail(PM_ERR_INVALID_ARG, "a device may be a member of a group only once"); }
        g->devices.push_back(d);
    }
    int st = load_nccl();
    if (st != PM_OK) { delete g; return st; }
    for (uint32_t i = 0; i < n_devices; i++) {
        pm_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.device = g->devices[i];
        cfg.flags = flags;
        pm_renderer *r = nullptr;
        st = pm_renderer_create(&r, &cfg);
        if (st != PM_OK) { pm_group_destroy(g); return st; }
        g->members.push_back(r);
    }
    g->comms.assign(n_devices, nullptr);
    ncclResult_t nr = g_nccl.CommInitAll(g->comms.data(), (int)n_devices, g->devices.data());
    if (nr != ncclSuccess) {
        g->comms.clear();
        pm_group_destroy(g);
        return fail(PM_ERR_CUDA, std::string("ncclCommInitAll failed: ") + g_nccl.GetErrorString(nr));
    }
    g->scene_dev.assign(n_devices, nullptr);
    g->scene_cap.assign(n_devices, 0);
    *out = g;
    return PM_OK;
}

void pm_group_destroy(pm_group *g) {
    if (!g) return;
    for (size_t i = 0; i < g->members.size(); i++) {
        cudaSetDevice(g->devices[i]);
        pm_renderer_sync(g->members[i], nullptr);
    }
    for (size_t i = 0; i < g->comms.size(); i++)
        if (g->comms[i]) g_nccl.CommDestroy(g->comms[i]);
    for (size_t i = 0; i < g->scene_dev.size(); i++) {
        cudaSetDevice(g->devices[i]);
        cudaFree(g->scene_dev[i]);
    }
    if (g->gather && g->gather_root >= 0) { cudaSetDevice(g->devices[g->gather_root]); cudaFree(g->gather); }
    for (pm_renderer *r : g->members) pm_renderer_destroy(r);
    cudaGetLastError();
    delete g;
}

uint32_t pm_group_size(const pm_group *g) { return g ? (uint32_t)g->members.size() : 0; }

int pm_group_member(pm_group *g, uint32_t index, pm_renderer **out) {
    if (!g || !out || index >= g->members.size()) return PM_ERR_INVALID_ARG;
    *out = g->members[index];
    return PM_OK;
}

int pm_group_resize(pm_group *g, uint32_t width, uint32_t height) {
    if (!g) return PM_ERR_INVALID_ARG;
    for (size_t i = 0; i < g->members.size(); i++) {
        int st = pm_renderer_resize(g->members[i], width, height);
        if (st != PM_OK) return st;
    }
    g->width = width; g->height = height;
    g->n_ty = (height + PM_TILE_H - 1) / PM_TILE_H;
    g->have_scene = false;
    g->bounds.clear();
    return PM_OK;
}

int pm_group_set_scene(pm_group *g, const uint8_t *scene, size_t len) {
    if (!g || !scene) return PM_ERR_INVALID_ARG;
    if (g->width == 0) return PM_ERR_STATE;
    const size_t n = g->members.size();
    if (n > g->n_ty) return fail(PM_ERR_INVALID_ARG, "more GPUs than tile rows");
    int st = pm_scene_validate(scene, len);
    if (st != PM_OK) return st;
    // the strips: contiguous tile rows, cut where the estimated cost is even (every GPU would derive the same bounds)
    std::vector<float> cost(g->n_ty);
    st = pm_scene_row_costs(scene, len, g->width, g->height, cost.data(), cost.size());
    if (st != PM_OK) return st;
    g->bounds.assign(n + 1, 0);
    st = pm_balance_strips(cost.data(), g->n_ty, (uint32_t)n, g->bounds.data());
    if (st != PM_OK) return st;
    for (size_t i = 0; i < n; i++) {
        PM_G_CUDA(cudaSetDevice(g->devices[i]));
        if (len > g->scene_cap[i]) {
            PM_G_CUDA(cudaStreamSynchronize(member_stream(g, i)));
            cudaFree(g->scene_dev[i]);
            g->scene_dev[i] = nullptr;
            g->scene_cap[i] = 0;
            PM_G_CUDA(cudaMalloc(&g->scene_dev[i], (len + 255) & ~(size_t)255));
            g->scene_cap[i] = (len + 255) & ~(size_t)255;
        }
    }
    // one upload, one broadcast
    PM_G_CUDA(cudaSetDevice(g->devices[0]));
    PM_G_CUDA(cudaMemcpyAsync(g->scene_dev[0], scene, len, cudaMemcpyHostToDevice, member_stream(g, 0)));
    if (n > 1) {
        PM_G_NCCL(g_nccl.GroupStart());
        for (size_t i = 0; i < n; i++) {
            ncclResult_t r = g_nccl.Broadcast(g->scene_dev[0], g->scene_dev[i], len, ncclUint8, 0, g->comms[i], member_stream(g, i));
            if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(PM_ERR_CUDA, std::string("ncclBroadcast failed: ") + g_nccl.GetErrorString(r)); }
        }
        PM_G_NCCL(g_nccl.GroupEnd());
    }
    for (size_t i = 0; i < n; i++) {
        st = pm_renderer_set_strip(g->members[i], g->bounds[i], g->bounds[i + 1]);
        if (st != PM_OK) return st;
        st = pm_renderer_set_scene_device(g->members[i], g->scene_dev[i], len);  // (same stream as the broadcast: ordered behind it)
        if (st != PM_OK) return st;
    }
    g->have_scene = true;
    return PM_OK;
}

int pm_group_strip_bounds(const pm_group *g, uint32_t *bounds, size_t cap) {
    if (!g || !bounds) return PM_ERR_INVALID_ARG;
    if (g->bounds.empty()) return PM_ERR_STATE;
    if (cap < g->bounds.size()) return PM_ERR_BUFFER_TOO_SMALL;
    memcpy(bounds, g->bounds.data(), g->bounds.size() * sizeof(uint32_t));
    return PM_OK;
}

int pm_group_set_frame_events(pm_group *g, int mode) {
    if (!g) return PM_ERR_INVALID_ARG;
    for (pm_renderer *r : g->members) {
        int st = pm_renderer_set_frame_events(r, mode);
        if (st != PM_OK) return st;
    }
    return PM_OK;
}

int pm_group_render(pm_group *g) {
    if (!g) return PM_ERR_INVALID_ARG;
    if (!g->have_scene) return PM_ERR_STATE;
    for (pm_renderer *r : g->members) {  // asynchronous: the GPUs render their strips side by side
        int st = pm_renderer_render(r);
        if (st != PM_OK) return st;
    }
    return PM_OK;
}

int pm_group_sync(pm_group *g, pm_frame_stats *stats, size_t n_stats, float *ms_frame_max) {
    if (!g) return PM_ERR_INVALID_ARG;
    float worst = 0.0f;
    for (size_t i = 0; i < g->members.size(); i++) {
        pm_frame_stats st;
        int rc = pm_renderer_sync(g->members[i], &st);
        if (rc != PM_OK) return rc;
        if (stats && i < n_stats) stats[i] = st;
        if (st.ms_total > worst) worst = st.ms_total;
    }
    if (ms_frame_max) *ms_frame_max = worst;
    return PM_OK;
}

int pm_group_read_rgba8(pm_group *g, uint8_t *dst, size_t stride) {
    if (!g || !dst) return PM_ERR_INVALID_ARG;
    if (!g->have_scene) return PM_ERR_STATE;
    if (stride < (size_t)g->width * 4) return PM_ERR_INVALID_ARG;
    for (size_t i = 0; i < g->members.size(); i++) {
        int st = pm_renderer_read_rgba8(g->members[i], dst + (size_t)g->bounds[i] * PM_TILE_H * stride, stride);
        if (st != PM_OK) return st;
    }
    return PM_OK;
}

int pm_group_gather_device(pm_group *g, uint32_t root, void **dev_ptr, size_t *pitch_bytes) {
    if (!g || root >= g->members.size() || !dev_ptr) return PM_ERR_INVALID_ARG;
    if (!g->have_scene) return PM_ERR_STATE;
    const size_t n = g->members.size();
    std::vector<void *> fb(n);
    std::vector<size_t> pitch(n);
    std::vector<uint32_t> rows(n);
    for (size_t i = 0; i < n; i++) {
        int st = pm_renderer_sync(g->members[i], nullptr);  // (also re-renders a frame whose record pool had to grow)
        if (st != PM_OK) return st;
        st = pm_renderer_framebuffer(g->members[i], &fb[i], &pitch[i], &rows[i]);
        if (st != PM_OK) return st;
    }
    const size_t p = pitch[0], total = p * (size_t)g->n_ty * PM_TILE_H;
    PM_G_CUDA(cudaSetDevice(g->devices[root]));
    if (g->gather_root != (int)root || total > g->gather_cap) {
        if (g->gather && g->gather_root >= 0) { cudaSetDevice(g->devices[g->gather_root]); cudaFree(g->gather); cudaSetDevice(g->devices[root]); }
        g->gather = nullptr; g->gather_cap = 0;
        PM_G_CUDA(cudaMalloc(&g->gather, total));
        g->gather_cap = total;
        g->gather_root = (int)root;
    }
    // the strips are of unequal height: point-to-point sends to the root instead of an all-gather
    PM_G_CUDA(cudaMemcpyAsync(g->gather + (size_t)g->bounds[root] * PM_TILE_H * p, fb[root], p * rows[root], cudaMemcpyDeviceToDevice, member_stream(g, root)));
    if (n > 1) {
        PM_G_NCCL(g_nccl.GroupStart());
        for (size_t i = 0; i < n; i++) {
            if (i == root) continue;
            ncclResult_t r = g_nccl.Send(fb[i], p * rows[i], ncclUint8, (int)root, g->comms[i], member_stream(g, i));
            if (r == ncclSuccess) r = g_nccl.Recv(g->gather + (size_t)g->bounds[i] * PM_TILE_H * p, p * rows[i], ncclUint8, (int)i, g->comms[root], member_stream(g, root));
            if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(PM_ERR_CUDA, std::string("ncclSend/ncclRecv failed: ") + g_nccl.GetErrorString(r)); }
        }
        PM_G_NCCL(g_nccl.GroupEnd());
    }
    for (size_t i = 0; i < n; i++) {
        PM_G_CUDA(cudaSetDevice(g->devices[i]));
        PM_G_CUDA(cudaStreamSynchronize(member_stream(g, i)));
    }
    *dev_ptr = g->gather;
    if (pitch_bytes) *pitch_bytes = p;
    return PM_OK;
}

int pm_group_nccl_version(void) {
    if (load_nccl() != PM_OK) return 0;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

}  // extern "C"
