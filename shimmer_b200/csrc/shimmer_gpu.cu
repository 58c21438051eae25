// libshimmer_gpu.so -- C ABI implementation (include/shimmer_gpu.h).
// Host side: scene upload (one-time staging into HBM), wavefront scheduling on one CUDA
// stream with device-side queue counters (no host sync inside a batch), film readback.
// There is NO CPU fallback: every entry point either runs the sm_100a kernels or fails.
#include "sg_kernels.h"
#include "sg_image.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <functional>
#include <vector>
#include <new>
#include <thread>
#include <algorithm>
#include <mutex>
#include <dlfcn.h>
#include <nccl.h>          // types and prototypes only: the library is dlopen'ed on first use (see Nccl below)

using namespace sg;

namespace {

thread_local std::string g_err;
// One entry per GPU this process drives.  g_dev[0] is the primary device: every entry point that is not a multi-GPU
// render runs there.  A scene replica remembers the index of its device.
constexpr int kMaxWavefronts = 2;   // 3 and 4 in flight measured no better than 2 (gpurun_out/r02_overlap2.log)
// aux[0]: the second wavefront of render_on; aux[1], aux[2]: the shadow-ray side streams of wavefront 0 and 1
struct Device { int id = -1; cudaStream_t stream = nullptr; cudaStream_t aux[3] = {}; int num_sms = 0; ncclComm_t comm = nullptr; };
std::vector<Device> g_dev;
#define g_device (g_dev.empty() ? -1 : g_dev[0].id)
#define g_stream (g_dev[0].stream)
// process communicator (one process per GPU): sg_comm_init_rank
ncclComm_t g_proc_comm = nullptr; int g_proc_rank = 0, g_proc_nranks = 1;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

// NCCL entry points resolved at run time: a single-GPU host does not need libnccl, and inside a process that already
// carries one (e.g. torch's bundled copy) the same library instance is used.
struct Nccl {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
} g_nccl;
int nccl_load() {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (g_nccl.h) return SG_OK;
    // SG_NCCL_LIB overrides the library; otherwise a copy the process already carries (e.g. torch's bundled one) is reused --
    // the loader keys on the SONAME, so loading a second, older libnccl.so.2 first would break whoever needs the newer one.
    void* h = nullptr;
    if (const char* path = std::getenv("SG_NCCL_LIB")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { if (h) break; h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); }
    if (!h) return fail(SG_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define NCCL_SYM(field, name) g_nccl.field = (decltype(g_nccl.field))dlsym(h, name); if (!g_nccl.field) return fail(SG_ERR_NCCL, std::string("libnccl lacks ") + name)
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank"); NCCL_SYM(CommInitAll, "ncclCommInitAll");
    NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(Reduce, "ncclReduce"); NCCL_SYM(GroupStart, "ncclGroupStart");
    NCCL_SYM(GroupEnd, "ncclGroupEnd"); NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
    g_nccl.h = h;
    return SG_OK;
}
#define NC(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) { \
    return fail(SG_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(r__)); } } while (0)

// Entry points run on the device their scene lives on (default: the primary) and leave the caller's current device alone.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev_index) { cudaGetDevice(&prev); if (dev_index >= 0 && dev_index < (int)g_dev.size() && prev != g_dev[dev_index].id) cudaSetDevice(g_dev[dev_index].id); }
    ~DeviceGuard() { int cur = -1; cudaGetDevice(&cur); if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};
#define ENTER(dev_index) if (g_dev.empty()) return fail(SG_ERR_NOT_INITIALIZED, "sg_init has not been called"); DeviceGuard guard__(dev_index)

// destroys CUDA events on every exit path of a function
struct EventBag {
    std::vector<cudaEvent_t> ev;
    cudaEvent_t make() { cudaEvent_t e = nullptr; if (cudaEventCreate(&e) != cudaSuccess) return nullptr; ev.push_back(e); return e; }
    ~EventBag() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};
#define CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { \
    return fail(e__ == cudaErrorMemoryAllocation ? SG_ERR_OUT_OF_MEMORY : SG_ERR_CUDA, \
                std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)

template <class T> int upload(const T* src, size_t n, T** dst, std::vector<void*>& owned) {
    *dst = nullptr;
    if (n == 0 || src == nullptr) return SG_OK;
    void* p = nullptr;
    CU(cudaMalloc(&p, n * sizeof(T)));
    owned.push_back(p);
    CU(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (T*)p;
    return SG_OK;
}

struct Workspace {
    uint32_t capacity = 0;
    int max_depth = -1;
    PathState st{};
    Queues q{};
    std::vector<void*> owned;
    void release() { for (void* p : owned) cudaFree(p); owned.clear(); capacity = 0; max_depth = -1; }
};

}  // namespace

struct SgScene {
    int dev = 0;                    // index into g_dev of the device that holds this replica
    std::vector<SgScene*> peers;    // replicas on g_dev[1..] (sg_init_multi); owned by the primary
    DScene d{};
    TraceScene ts{};
    size_t smem_closest = 0, smem_shadow = 0, smem_closest_dual = 0, smem_shadow_dual = 0;
    std::vector<void*> owned;
    Workspace ws[2];                 // wavefronts in flight (render_on deals batches round-robin to this many streams); kMaxWavefronts
    DevStats* d_stats = nullptr;
    unsigned long long* d_cursor = nullptr;
    bool kinds_present[8] = {false, false, false, false, false, false, false, false};
    bool instanced = false;         // object instances: k_trace<.., INST = true> and the hit_inst path-state array
    bool general_lights = false;    // sphere / patch / point / image-infinite lights: k_shade<KIND, true, true, true>
    bool has_mix = false;           // Mix materials: k_resolve_mix + the mat_override path-state array
    bool tex_path = false;          // image textures, a non-zero constant displacement or non-triangle emitters: k_shade<KIND, true>
    bool sort_by_material = false;  // tex_path and some kind has more than one material: shade queues are re-ordered by material id
    bool staged_shading = false;    // tex_path with Diffuse materials: k_shade STAGE 1 (get_bsdf -> record) + STAGE 2 (the rest) for the path integrator
    int materials_of_kind[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double* d_film = nullptr; size_t film_pixels = 0;
    SgFilmPixel* h_film = nullptr; size_t h_film_pixels = 0;      // pinned staging for sg_render
    uint64_t n_pixels() const { return (uint64_t)(d.film.pixel_bounds[2] - d.film.pixel_bounds[0]) * (uint64_t)(d.film.pixel_bounds[3] - d.film.pixel_bounds[1]); }
};

namespace {

template <class T> int ws_alloc(Workspace& w, T** p, size_t n) {
    void* q = nullptr;
    CU(cudaMalloc(&q, n * sizeof(T)));
    w.owned.push_back(q);
    *p = (T*)q;
    return SG_OK;
}
int ensure_workspace(SgScene* s, int which, uint32_t capacity, int max_depth) {
    Workspace& w = s->ws[which];
    if (w.capacity >= capacity && w.max_depth >= max_depth) return SG_OK;
    w.release();
    w.st = PathState{};
    int rc;
    const size_t n = capacity;
#define WS(field) if ((rc = ws_alloc(w, &w.st.field, n)) != SG_OK) return rc
    WS(ray_o); WS(ray_d); WS(hit_b); WS(hit_prim); WS(L); WS(beta); WS(lambda); WS(lpdf); WS(rng_a); WS(rng_b);
    WS(pixel); WS(flags); WS(pb_eta); WS(ctx0); WS(ctx1); WS(ctx2); WS(sh_o); WS(sh_d); WS(sh_L);
    if (s->d.n_textures > 0) { WS(aux0); WS(aux1); WS(aux2); }
    if (s->staged_shading) for (int r = 0; r < 6; ++r) if ((rc = ws_alloc(w, &w.st.rec[r], n)) != SG_OK) return rc;
    if (s->instanced) { WS(hit_inst); }
    if (s->has_mix) { WS(mat_override); }      // ray differentials only feed image-texture filtering
#undef WS
    if ((rc = ws_alloc(w, &w.q.ray[0], n)) != SG_OK) return rc;
    if ((rc = ws_alloc(w, &w.q.ray[1], n)) != SG_OK) return rc;
    for (int k = 0; k < Q_NKINDS; ++k) if ((rc = ws_alloc(w, &w.q.shade[k], n)) != SG_OK) return rc;
    if ((rc = ws_alloc(w, &w.q.shadow, n)) != SG_OK) return rc;
    if ((rc = ws_alloc(w, &w.q.counters, (size_t)(max_depth + 3) * C_STRIDE)) != SG_OK) return rc;
    w.q.sorted = nullptr; w.q.sort_hist = nullptr; w.q.ray_kind = nullptr; w.q.block_counts = nullptr;
    {   // order-preserving shade queues (sg_wavefront.cuh k_queue_*); the 8 extra bytes let the last thread of a block load whole words
        static const int ordered_env = [] { const char* v = std::getenv("SG_ORDERED_QUEUES"); return v ? std::atoi(v) : 1; }();
        if (ordered_env) {
            if ((rc = ws_alloc(w, &w.q.ray_kind, n + 8)) != SG_OK) return rc;
            if ((rc = ws_alloc(w, &w.q.block_counts, ((n + kQueueBlock - 1) / kQueueBlock) * (size_t)Q_NKINDS)) != SG_OK) return rc;
        }
    }
    if (s->sort_by_material) {                  // textured scenes with several materials of a kind: see k_sort_queue_* (sg_wavefront.cuh)
        if ((rc = ws_alloc(w, &w.q.sorted, n)) != SG_OK) return rc;
        if ((rc = ws_alloc(w, &w.q.sort_hist, (size_t)128)) != SG_OK) return rc;
    }
    w.capacity = capacity; w.max_depth = max_depth;
    return SG_OK;
}

typedef void (*TraceKernel)(const DScene, const TraceScene, PathState, Queues, int, DevStats*);
TraceKernel trace_kernel(bool any, bool count, bool inst) {
    if (any) return count ? (inst ? k_trace<true, true, true> : k_trace<true, true, false>) : (inst ? k_trace<true, false, true> : k_trace<true, false, false>);
    return count ? (inst ? k_trace<false, true, true> : k_trace<false, true, false>) : (inst ? k_trace<false, false, true> : k_trace<false, false, false>);
}
TraceKernel trace_kernel_dual(bool any) { return any ? k_trace_dual<true> : k_trace_dual<false>; }
typedef void (*TraceRaysKernel)(const DScene, const TraceScene, long long, const float*, const float*, const float*, SgHit*, unsigned long long*, DevStats*);
TraceRaysKernel trace_rays_kernel(bool any, bool count, bool inst) {
    if (any) return count ? (inst ? k_trace_rays<true, true, true> : k_trace_rays<true, true, false>) : (inst ? k_trace_rays<true, false, true> : k_trace_rays<true, false, false>);
    return count ? (inst ? k_trace_rays<false, true, true> : k_trace_rays<false, true, false>) : (inst ? k_trace_rays<false, false, true> : k_trace_rays<false, false, false>);
}

int persistent_grid(int num_sms, const void* kernel, int threads, size_t smem) {
    int per_sm = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (const char* v = std::getenv("SG_TRACE_BLOCKS_PER_SM")) { const int cap = std::atoi(v); if (cap >= 1 && cap < per_sm) per_sm = cap; }   // A/B knob: leave room for other streams' CTAs
    return num_sms * per_sm;        // grid = SM count x resident CTAs: one full wave, persistent
}

}  // namespace

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }
const char* sg_last_error(void) { return g_err.c_str(); }

static int device_open(int device, Device& D) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(SG_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SG_ERR_INVALID_ARGUMENT, "device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SG_ERR_UNSUPPORTED, std::string("kernels are built for sm_100a only; found ") + prop.name);
    D.id = device; D.num_sms = prop.multiProcessorCount; D.comm = nullptr;
    CU(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
    for (cudaStream_t& a : D.aux) CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    return SG_OK;
}
static void devices_close() {
    for (Device& D : g_dev) {
        if (D.id >= 0) cudaSetDevice(D.id);
        if (D.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(D.comm);
        if (D.stream) cudaStreamDestroy(D.stream);
        for (cudaStream_t a : D.aux) if (a) cudaStreamDestroy(a);
    }
    g_dev.clear();
}

int sg_init_multi(const int* devices, int n) {
    if (!devices || n < 1) return fail(SG_ERR_INVALID_ARGUMENT, "sg_init_multi: need at least one device");
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) if (devices[i] == devices[j]) return fail(SG_ERR_INVALID_ARGUMENT, "sg_init_multi: duplicate device");
    bool same = (int)g_dev.size() == n;
    for (int i = 0; same && i < n; ++i) same = g_dev[i].id == devices[i];
    if (same) { CU(cudaSetDevice(g_dev[0].id)); return SG_OK; }
    devices_close();                        // a different device set: streams (and the communicator) belong to the old one
    std::vector<Device> devs(n);
    for (int i = 0; i < n; ++i) {
        const int rc = device_open(devices[i], devs[i]);
        if (rc != SG_OK) { g_dev.assign(devs.begin(), devs.begin() + i); devices_close(); return rc; }
    }
    g_dev = devs;
    if (n > 1) {
        int rc = nccl_load();
        if (rc != SG_OK) { devices_close(); return rc; }
        std::vector<ncclComm_t> comms(n, nullptr);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devices);
        if (r != ncclSuccess) { devices_close(); return fail(SG_ERR_NCCL, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r)); }
        for (int i = 0; i < n; ++i) g_dev[i].comm = comms[i];
    }
    CU(cudaSetDevice(g_dev[0].id));
    return SG_OK;
}
int sg_init(int device) { return sg_init_multi(&device, 1); }
int sg_device_count(void) { return (int)g_dev.size(); }

int sg_shutdown(void) {
    if (g_proc_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(g_proc_comm); }
    g_proc_comm = nullptr; g_proc_rank = 0; g_proc_nranks = 1;
    devices_close();
    return SG_OK;
}

int sg_comm_get_unique_id(void* id_out) {
    if (!id_out) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    static_assert(sizeof(ncclUniqueId) == SG_COMM_ID_BYTES, "SG_COMM_ID_BYTES must match ncclUniqueId");
    int rc = nccl_load();
    if (rc != SG_OK) return rc;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return SG_OK;
}
int sg_comm_init_rank(const void* id_in, int rank, int n_ranks) {
    ENTER(0);
    if (!id_in || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SG_ERR_INVALID_ARGUMENT, "sg_comm_init_rank: bad rank / n_ranks");
    if (g_dev.size() != 1) return fail(SG_ERR_INVALID_ARGUMENT, "a process communicator needs a single-device process (sg_init, not sg_init_multi)");
    if (g_proc_comm) return fail(SG_ERR_INVALID_ARGUMENT, "a process communicator already exists (sg_comm_destroy first)");
    int rc = nccl_load();
    if (rc != SG_OK) return rc;
    CU(cudaSetDevice(g_dev[0].id));
    ncclUniqueId id; std::memcpy(&id, id_in, sizeof id);
    NC(g_nccl.CommInitRank(&g_proc_comm, n_ranks, id, rank));
    g_proc_rank = rank; g_proc_nranks = n_ranks;
    return SG_OK;
}
int sg_comm_destroy(void) {
    if (g_proc_comm) { cudaSetDevice(g_dev.empty() ? 0 : g_dev[0].id); cudaDeviceSynchronize(); NC(g_nccl.CommDestroy(g_proc_comm)); }
    g_proc_comm = nullptr; g_proc_rank = 0; g_proc_nranks = 1;
    return SG_OK;
}
int sg_comm_rank(int* rank, int* n_ranks) {
    if (rank) *rank = g_proc_rank;
    if (n_ranks) *n_ranks = g_proc_nranks;
    return SG_OK;
}
int sg_sample_range_for_rank(int32_t begin, int32_t end, int rank, int n_ranks, int32_t* out_begin, int32_t* out_end) {
    if (!out_begin || !out_end || end < begin || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SG_ERR_INVALID_ARGUMENT, "sg_sample_range_for_rank: bad arguments");
    const int32_t n = end - begin, base = n / n_ranks, rem = n % n_ranks;
    *out_begin = begin + rank * base + std::min(rank, rem);
    *out_end = *out_begin + base + (rank < rem ? 1 : 0);
    return SG_OK;
}
int sg_film_reduce_device(void* d_film, int64_t n_pixels, void* stream_v) {
    ENTER(0);
    if (!d_film || n_pixels < 0) return fail(SG_ERR_INVALID_ARGUMENT, "sg_film_reduce_device: bad arguments");
    if (!g_proc_comm || g_proc_nranks == 1 || n_pixels == 0) return SG_OK;
    NC(g_nccl.Reduce(d_film, d_film, (size_t)n_pixels * 4, ncclDouble, ncclSum, 0, g_proc_comm, (cudaStream_t)stream_v));
    return SG_OK;
}

static int scene_create_on(const SgSceneDesc* desc, int dev_index, SgScene** out);
int sg_scene_create(const SgSceneDesc* desc, SgScene** out) {
    if (g_dev.empty()) return fail(SG_ERR_NOT_INITIALIZED, "sg_init has not been called");
    if (!desc || !out) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    SgScene* primary = nullptr;
    int rc = scene_create_on(desc, 0, &primary);
    if (rc != SG_OK) return rc;
    for (int i = 1; i < (int)g_dev.size(); ++i) {           // sg_init_multi: one replica per device, staged once
        SgScene* peer = nullptr;
        rc = scene_create_on(desc, i, &peer);
        if (rc != SG_OK) { const std::string msg = g_err; sg_scene_destroy(primary); g_err = msg; return rc; }
        primary->peers.push_back(peer);
    }
    *out = primary;
    return SG_OK;
}
static int scene_create_on(const SgSceneDesc* desc, int dev_index, SgScene** out) {
    DeviceGuard guard__(dev_index);
    if (desc->abi_version != SG_ABI_VERSION) return fail(SG_ERR_INVALID_ARGUMENT, "SgSceneDesc.abi_version mismatch");
    if (desc->n_primitives > 0 && (!desc->nodes || !desc->primitives)) return fail(SG_ERR_INVALID_ARGUMENT, "geometry arrays missing");
    if (desc->n_spheres && !desc->spheres) return fail(SG_ERR_INVALID_ARGUMENT, "sphere array missing");
    if ((desc->n_textures && !desc->textures) || (desc->n_materials && !desc->materials) || (desc->n_lights && !desc->lights) ||
        (desc->n_spectra && !desc->spectra) || (desc->n_meshes && !desc->meshes))
        return fail(SG_ERR_INVALID_ARGUMENT, "texture / material / light / spectrum / mesh arrays missing");
    if (desc->camera.kind != SG_CAMERA_PERSPECTIVE && desc->camera.kind != SG_CAMERA_ORTHOGRAPHIC) return fail(SG_ERR_UNSUPPORTED, "camera kind is not on the GPU path");
    if (desc->n_primitives >= (1u << 31)) return fail(SG_ERR_UNSUPPORTED, "too many primitives");
    // validate references so device code never reads out of bounds
    const uint32_t n_top_nodes = desc->n_top_nodes ? desc->n_top_nodes : desc->n_nodes;
    const uint32_t n_top_prims = desc->n_top_primitives ? desc->n_top_primitives : desc->n_primitives;
    if (n_top_nodes > desc->n_nodes || n_top_prims > desc->n_primitives) return fail(SG_ERR_INVALID_ARGUMENT, "top-level ranges exceed the node / primitive arrays");
    if ((desc->n_objects && !desc->objects) || (desc->n_instances && !desc->instances)) return fail(SG_ERR_INVALID_ARGUMENT, "object / instance arrays missing");
    for (uint32_t i = 0; i < desc->n_primitives; ++i) {
        const SgPrimitive& p = desc->primitives[i];
        if (p.mesh == SG_PRIM_INSTANCE) {
            if (i >= n_top_prims) return fail(SG_ERR_UNSUPPORTED, "nested object instances are not supported (pbrt-v4 scene format)");
            if (p.tri >= desc->n_instances) return fail(SG_ERR_INVALID_ARGUMENT, "primitive " + std::to_string(i) + " references an out-of-range instance");
            continue;
        }
        if (p.mesh == SG_PRIM_SPHERE) {
            if (i >= n_top_prims && p.light >= 0) return fail(SG_ERR_UNSUPPORTED, "area lights inside object definitions are not supported (pbrt-v4 scene format)");
            if (p.tri >= desc->n_spheres || p.material >= desc->n_materials) return fail(SG_ERR_INVALID_ARGUMENT, "primitive " + std::to_string(i) + " references an out-of-range sphere/material");
            if (p.light >= (int32_t)desc->n_lights || (p.light >= 0 && (desc->lights[p.light].kind != SG_LIGHT_DIFFUSE_AREA_SPHERE || desc->lights[p.light].tri != p.tri)))
                return fail(SG_ERR_INVALID_ARGUMENT, "an emissive sphere must point at an SG_LIGHT_DIFFUSE_AREA_SPHERE light over that sphere");
            const SgSphere& sp = desc->spheres[p.tri];
            const float* r3 = sp.render_from_object + 12; const float* q3 = sp.object_from_render + 12;
            if (r3[0] != 0.0f || r3[1] != 0.0f || r3[2] != 0.0f || r3[3] != 1.0f || q3[0] != 0.0f || q3[1] != 0.0f || q3[2] != 0.0f || q3[3] != 1.0f)
                return fail(SG_ERR_UNSUPPORTED, "sphere transforms must be affine");
            if (!(sp.radius > 0.0f)) return fail(SG_ERR_INVALID_ARGUMENT, "sphere radius must be positive");
            continue;
        }
        if (!desc->meshes || !desc->indices || !desc->p) return fail(SG_ERR_INVALID_ARGUMENT, "geometry arrays missing");
        if (p.mesh < desc->n_meshes && (desc->meshes[p.mesh].flags & SG_MESH_BILINEAR)) {
            if (i >= n_top_prims && p.light >= 0) return fail(SG_ERR_UNSUPPORTED, "area lights inside object definitions are not supported (pbrt-v4 scene format)");
            if (p.light >= (int32_t)desc->n_lights || (p.light >= 0 && (desc->lights[p.light].kind != SG_LIGHT_DIFFUSE_AREA_PATCH || desc->lights[p.light].mesh != p.mesh ||
                                                                     desc->lights[p.light].tri != p.tri)))
                return fail(SG_ERR_INVALID_ARGUMENT, "an emissive bilinear patch must point at an SG_LIGHT_DIFFUSE_AREA_PATCH light over that patch");
        }
        if (p.mesh >= desc->n_meshes || p.tri >= desc->meshes[p.mesh].n_triangles || p.material >= desc->n_materials ||
            p.light >= (int32_t)desc->n_lights)
            return fail(SG_ERR_INVALID_ARGUMENT, "primitive " + std::to_string(i) + " references out-of-range mesh/triangle/material/light");
        if (!(desc->meshes[p.mesh].flags & SG_MESH_BILINEAR) && p.light >= 0) {
            // light sampling reads the emitter's render-space vertices (light_verts): an emitter inside an object definition would be
            // sampled untransformed, and a light row of another kind / over another triangle would be sampled instead of this one
            if (i >= n_top_prims) return fail(SG_ERR_UNSUPPORTED, "area lights inside object definitions are not supported (pbrt-v4 scene format)");
            const SgLight& L = desc->lights[p.light];
            if (L.kind != SG_LIGHT_DIFFUSE_AREA || L.mesh != p.mesh || L.tri != p.tri)
                return fail(SG_ERR_INVALID_ARGUMENT, "an emissive triangle must point at an SG_LIGHT_DIFFUSE_AREA light over that triangle");
        }
    }
    for (uint32_t i = 0; i < desc->n_objects; ++i) {
        const SgObject& o = desc->objects[i];
        if ((uint64_t)o.first_node + o.n_nodes > desc->n_nodes || (uint64_t)o.first_prim + o.n_prims > desc->n_primitives || o.n_prims == 0 ||
            o.first_prim < n_top_prims || (o.n_nodes && o.first_node < n_top_nodes))
            return fail(SG_ERR_INVALID_ARGUMENT, "object " + std::to_string(i) + ": node / primitive range is malformed");
    }
    for (uint32_t i = 0; i < desc->n_instances; ++i)
        if (desc->instances[i].object >= desc->n_objects) return fail(SG_ERR_INVALID_ARGUMENT, "instance " + std::to_string(i) + " references an out-of-range object");
    for (uint32_t i = 0; i < desc->n_lights; ++i) {
        const SgLight& L = desc->lights[i];
        if (L.kind < SG_LIGHT_DIFFUSE_AREA || L.kind > SG_LIGHT_DIFFUSE_AREA_PATCH) return fail(SG_ERR_UNSUPPORTED, "light kind " + std::to_string(L.kind) + " is not on the GPU path");
        if (L.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE && L.tri >= desc->n_spheres) return fail(SG_ERR_INVALID_ARGUMENT, "light " + std::to_string(i) + " references an out-of-range sphere");
        if (L.kind == SG_LIGHT_IMAGE_INFINITE && (L.tri >= desc->n_env_maps || !desc->env_maps)) return fail(SG_ERR_INVALID_ARGUMENT, "light " + std::to_string(i) + " references an out-of-range environment map");
        if (L.spectrum < 0 || L.spectrum >= (int32_t)desc->n_spectra) return fail(SG_ERR_INVALID_ARGUMENT, "light " + std::to_string(i) + ": spectrum id out of range");
    }
    for (uint32_t i = 0; i < desc->n_env_maps; ++i) {
        const SgEnvMap& E = desc->env_maps[i];
        if (E.res < 1 || !desc->texels || E.texel_offset + (uint64_t)E.res * (uint64_t)E.res * 3 > desc->n_texels)
            return fail(SG_ERR_INVALID_ARGUMENT, "environment map " + std::to_string(i) + ": image outside the texel pool");
        for (const SgDistribution2D* D2 : {&E.distribution, &E.compensated}) {
            const uint64_t nu = (uint64_t)D2->nu, nv = (uint64_t)D2->nv;
            if (D2->nu < 1 || D2->nv < 1 || D2->func_off + nu * nv > desc->n_pool || D2->cdf_off + (nu + 1) * nv > desc->n_pool ||
                D2->marg_func_off + nv > desc->n_pool || D2->marg_cdf_off + nv + 1 > desc->n_pool)
                return fail(SG_ERR_INVALID_ARGUMENT, "environment map " + std::to_string(i) + ": sampling distribution outside spectrum_pool");
        }
    }
    for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const SgMaterial& m = desc->materials[i];
        if (m.kind < 0 || m.kind > SG_MATERIAL_MIX) return fail(SG_ERR_UNSUPPORTED, "material kind " + std::to_string(m.kind) + " is not on the GPU path");
        if (m.kind == SG_MATERIAL_MIX) continue;                 // validated below
        if (m.kind == SG_MATERIAL_COATED_CONDUCTOR && (m.spec_b < 0 || m.spec_b >= (int32_t)desc->n_spectra || m.spec_c < 0 || m.spec_c >= (int32_t)desc->n_spectra ||
                                                       m.max_depth < 0 || m.n_samples < 1 ||
                                                       (!(m.flags & SG_MAT_CONDUCTOR_REFLECTANCE) && (m.spec_d < 0 || m.spec_d >= (int32_t)desc->n_spectra))))
            return fail(SG_ERR_INVALID_ARGUMENT, "coated conductor material parameters out of range");
        if (m.kind == SG_MATERIAL_COATED_DIFFUSE && (m.spec_b < 0 || m.spec_b >= (int32_t)desc->n_spectra || m.spec_c < 0 || m.spec_c >= (int32_t)desc->n_spectra ||
                                                     m.max_depth < 0 || m.n_samples < 1))
            return fail(SG_ERR_INVALID_ARGUMENT, "coated diffuse material parameters out of range");
        if (m.spec_a < 0 || m.spec_a >= (int32_t)desc->n_spectra || (m.kind == SG_MATERIAL_CONDUCTOR && (m.spec_b < 0 || m.spec_b >= (int32_t)desc->n_spectra)))
            return fail(SG_ERR_INVALID_ARGUMENT, "material spectrum id out of range");
    }
    for (uint32_t i = 0; i < desc->n_materials; ++i) {                 // Mix: valid children, no cycles, resolves within 64 steps
        const SgMaterial& m = desc->materials[i];
        if (m.kind != SG_MATERIAL_MIX) continue;
        for (int c = 0; c < 2; ++c) {
            int32_t cur = m.mix_materials[c]; int steps = 0;
            for (;;) {
                if (cur < 0 || cur >= (int32_t)desc->n_materials) return fail(SG_ERR_INVALID_ARGUMENT, "mix material " + std::to_string(i) + " references an out-of-range material");
                if (desc->materials[cur].kind != SG_MATERIAL_MIX) break;
                if (++steps > 32) return fail(SG_ERR_INVALID_ARGUMENT, "mix material " + std::to_string(i) + ": nesting deeper than 32 (or cyclic)");
                cur = desc->materials[cur].mix_materials[0];     // left spine; the right spines are checked from their own rows
            }
        }
        if (m.tex_mix_amount >= (int32_t)desc->n_textures || (m.tex_mix_amount >= 0 && desc->textures[m.tex_mix_amount].n_channels != 1))
            return fail(SG_ERR_INVALID_ARGUMENT, "mix material " + std::to_string(i) + ": `amount` texture must be a one-channel image");
    }
    for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const SgMaterial& m = desc->materials[i];
        if (m.kind == SG_MATERIAL_MIX) continue;
        if (m.normal_map >= (int32_t)desc->n_textures || (m.normal_map >= 0 && (desc->textures[m.normal_map].n_channels != 3 || desc->textures[m.normal_map].kind != SG_TEXTURE_IMAGE)))
            return fail(SG_ERR_INVALID_ARGUMENT, "material " + std::to_string(i) + ": normal maps must be three-channel images");
        for (int32_t t : {m.tex_reflectance, m.tex_displacement})
            if (t >= (int32_t)desc->n_textures) return fail(SG_ERR_INVALID_ARGUMENT, "material " + std::to_string(i) + " references an out-of-range texture");
        if (m.tex_reflectance >= 0 && m.kind != SG_MATERIAL_DIFFUSE && m.kind != SG_MATERIAL_COATED_DIFFUSE)
            return fail(SG_ERR_UNSUPPORTED, "reflectance textures are on the GPU path for diffuse and coated-diffuse materials only");
        if (m.tex_displacement >= 0 && desc->textures[m.tex_displacement].n_channels != 1)
            return fail(SG_ERR_UNSUPPORTED, "displacement textures must be one-channel images");
    }
    bool need_rgb2spec = false;
    // depth of a texture tree below row i (image / constant rows: 0); -1 = malformed (bad ids, wrong operand type, cycle, too deep)
    std::function<int(int32_t, bool, int)> tex_depth = [&](int32_t id, bool want_float, int budget) -> int {
        if (id < 0 || (uint32_t)id >= desc->n_textures || budget < 0) return -1;
        const SgTexture& t = desc->textures[id];
        if (want_float && t.n_channels != 1) return -1;
        if (t.kind == SG_TEXTURE_IMAGE) return 0;
        if (t.kind < SG_TEXTURE_CONSTANT || t.kind > SG_TEXTURE_DIRECTION_MIX || !desc->texture_nodes || t.node < 0 || (uint32_t)t.node >= desc->n_texture_nodes) return -1;
        const SgTextureNode& nd = desc->texture_nodes[t.node];
        if (t.kind == SG_TEXTURE_CONSTANT) return (t.n_channels == 1 || nd.spectrum < 0 || (uint32_t)nd.spectrum < desc->n_spectra) ? 0 : -1;
        const bool is_float = t.n_channels == 1;
        const int a = tex_depth(nd.tex1, is_float, budget - 1);
        const int b = tex_depth(nd.tex2, t.kind == SG_TEXTURE_SCALED ? true : is_float, budget - 1);
        const int c = t.kind == SG_TEXTURE_MIX ? tex_depth(nd.amount, true, budget - 1) : 0;
        if (a < 0 || b < 0 || c < 0) return -1;
        return 1 + std::max(a, std::max(b, c));
    };
    for (uint32_t i = 0; i < desc->n_textures; ++i) {
        const SgTexture& t = desc->textures[i];
        if (t.kind != SG_TEXTURE_IMAGE) {
            if (t.n_channels < 1 || tex_depth((int32_t)i, false, SG_MAX_TEXTURE_DEPTH) < 0)
                return fail(SG_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": bad kind / node / operand ids or types, a cycle, or operands nested deeper than SG_MAX_TEXTURE_DEPTH");
            continue;
        }
        if (!desc->image_levels || !desc->texels) return fail(SG_ERR_INVALID_ARGUMENT, "texture arrays missing");
        if ((t.n_channels != 1 && t.n_channels != 3) || t.n_levels < 1 || (uint64_t)t.first_level + (uint64_t)t.n_levels > desc->n_image_levels)
            return fail(SG_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": bad channel count or level range");
        if (t.wrap < SG_WRAP_REPEAT || t.wrap > SG_WRAP_CLAMP || t.filter < SG_FILTER_POINT || t.filter > SG_FILTER_EWA ||
            t.spectrum_type < SG_SPECTRUM_TYPE_ALBEDO || t.spectrum_type > SG_SPECTRUM_TYPE_UNBOUNDED)
            return fail(SG_ERR_UNSUPPORTED, "texture " + std::to_string(i) + ": wrap / filter / spectrum type not on the GPU path");
        if (t.mapping >= (int32_t)desc->n_texture_mappings || (t.mapping >= 0 && (!desc->texture_mappings || desc->texture_mappings[t.mapping].kind < SG_MAPPING_SPHERICAL ||
                                                                            desc->texture_mappings[t.mapping].kind > SG_MAPPING_PLANAR)))
            return fail(SG_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": texture mapping out of range");
        if (t.filter == SG_FILTER_EWA && !desc->mip_filter_lut) return fail(SG_ERR_INVALID_ARGUMENT, "EWA filtering needs mip_filter_lut");
        for (int32_t l = 0; l < t.n_levels; ++l) {
            const SgImageLevel& L = desc->image_levels[t.first_level + l];
            if (L.res[0] < 1 || L.res[1] < 1 || (uint64_t)L.offset + (uint64_t)L.res[0] * (uint64_t)L.res[1] * (uint64_t)t.n_channels > desc->n_texels)
                return fail(SG_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": MIP level outside the texel pool");
        }
        const SgImageLevel& last = desc->image_levels[t.first_level + t.n_levels - 1];
        if (last.res[0] != 1 || last.res[1] != 1) return fail(SG_ERR_INVALID_ARGUMENT, "texture " + std::to_string(i) + ": the last MIP level must be 1x1 (image.rs:783)");
        need_rgb2spec |= t.n_channels == 3;
    }
    if (desc->material_textures) for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const SgMaterialTextures& mt = desc->material_textures[i];
        for (int32_t t : {mt.u_roughness, mt.v_roughness, mt.thickness, mt.g, mt.u_roughness2, mt.v_roughness2})
            if (t >= (int32_t)desc->n_textures || (t >= 0 && desc->textures[t].n_channels != 1))
                return fail(SG_ERR_INVALID_ARGUMENT, "material " + std::to_string(i) + ": a float parameter references a missing or non-float texture");
        for (int32_t t : {mt.spec_a, mt.spec_b, mt.spec_d})
            if (t >= (int32_t)desc->n_textures) return fail(SG_ERR_INVALID_ARGUMENT, "material " + std::to_string(i) + ": a spectrum parameter references an out-of-range texture");
    }
    need_rgb2spec |= desc->n_env_maps > 0;
    if (need_rgb2spec && (desc->rgb2spec_res < 2 || !desc->rgb2spec_scale || !desc->rgb2spec_data))
        return fail(SG_ERR_INVALID_ARGUMENT, "three-channel textures need the rgb2spec table of the scene colour space");
    auto check_bvh = [&](uint32_t node_base, uint32_t nn, uint32_t np) -> bool {          // offsets are relative to the BVH's own ranges
        for (uint32_t i = 0; i < nn; ++i) {
            const SgBvhNode& nd = desc->nodes[node_base + i];
            if (nd.n_prims > 0 ? ((uint64_t)nd.offset + nd.n_prims > np) : (nd.offset >= nn || nd.offset <= i || i + 1 >= nn)) return false;
        }
        return true;
    };
    if (!check_bvh(0, n_top_nodes, n_top_prims)) return fail(SG_ERR_INVALID_ARGUMENT, "top-level BVH is malformed");
    for (uint32_t i = 0; i < desc->n_objects; ++i)
        if (!check_bvh(desc->objects[i].first_node, desc->objects[i].n_nodes, desc->objects[i].n_prims))
            return fail(SG_ERR_INVALID_ARGUMENT, "BVH of object " + std::to_string(i) + " is malformed");
    SgScene* s = new (std::nothrow) SgScene();
    if (!s) return fail(SG_ERR_OUT_OF_MEMORY, "host allocation failed");
    s->dev = dev_index;
    int rc = SG_OK;
    auto bail = [&](int code) { sg_scene_destroy(s); return code; };
    DScene& d = s->d;
    // pre-gathered triangle vertices in BVH-leaf order (DESIGN.md "Data layout")
    std::vector<float4> tv((size_t)desc->n_primitives * 3);
    std::vector<float4> pv;                                 // bilinear patch vertices, 4 float4 per patch primitive
    for (uint32_t i = 0; i < desc->n_primitives; ++i) {
        const SgPrimitive& p = desc->primitives[i];
        if (p.mesh == SG_PRIM_INSTANCE) {                   // TransformedPrimitive: no vertices; kind 7, instance id in word 1
            const uint32_t w[3] = {kKindInstance << 28, p.tri, 0u};
            for (int k = 0; k < 3; ++k) { float wf; std::memcpy(&wf, &w[k], 4); tv[3 * (size_t)i + k] = make_float4(0.0f, 0.0f, 0.0f, wf); }
            continue;
        }
        if (p.mesh == SG_PRIM_SPHERE) {                     // Shape::Sphere: material + kind like a triangle, sphere index | kSphereBit in word 2
            if (p.material >= (1u << 23)) { g_err = "more than 2^23 materials"; return bail(SG_ERR_UNSUPPORTED); }
            const uint32_t w[3] = {p.material | ((uint32_t)desc->materials[p.material].kind << 28), (uint32_t)p.light, p.tri | kSphereBit};
            const float* M = desc->spheres[p.tri].render_from_object;
            for (int k = 0; k < 3; ++k) { float wf; std::memcpy(&wf, &w[k], 4); tv[3 * (size_t)i + k] = make_float4(M[3], M[7], M[11], wf); }
            s->kinds_present[desc->materials[p.material].kind] = true;
            continue;
        }
        const SgMesh& m = desc->meshes[p.mesh];
        if (m.flags & SG_MESH_BILINEAR) {                   // Shape::BilinearPatch: 4 vertices in patch_verts, record index | kPatchBit in word 2
            if (p.material >= (1u << 23)) { g_err = "more than 2^23 materials"; return bail(SG_ERR_UNSUPPORTED); }
            const uint32_t rec = (uint32_t)(pv.size() / 4);
            const uint32_t* ix4 = desc->indices + m.first_index + 4 * (size_t)p.tri;
            const uint32_t meta[4] = {m.flags, p.mesh, p.tri, 0u};
            for (int k = 0; k < 4; ++k) {
                if (ix4[k] >= m.n_vertices) { g_err = "vertex index out of range"; return bail(SG_ERR_INVALID_ARGUMENT); }
                const float* q = desc->p + 3 * (size_t)(m.first_vertex + ix4[k]);
                float wf; std::memcpy(&wf, &meta[k], 4);
                pv.push_back(make_float4(q[0], q[1], q[2], wf));
            }
            const uint32_t w[3] = {p.material | ((uint32_t)desc->materials[p.material].kind << 28), (uint32_t)p.light, rec | kPatchBit};
            for (int k = 0; k < 3; ++k) { float wf; std::memcpy(&wf, &w[k], 4); tv[3 * (size_t)i + k] = make_float4(0.0f, 0.0f, 0.0f, wf); }
            s->kinds_present[desc->materials[p.material].kind] = true;
            continue;
        }
        const uint32_t* ix = desc->indices + m.first_index + 3 * (size_t)p.tri;
        if (p.material >= (1u << 23)) { g_err = "more than 2^23 materials"; return bail(SG_ERR_UNSUPPORTED); }
        const uint32_t w[3] = {p.material | ((m.flags & 31u) << 23) | ((uint32_t)desc->materials[p.material].kind << 28), (uint32_t)p.light, p.mesh};
        for (int k = 0; k < 3; ++k) {
            if (ix[k] >= m.n_vertices) { g_err = "vertex index out of range"; return bail(SG_ERR_INVALID_ARGUMENT); }
            const float* q = desc->p + 3 * (size_t)(m.first_vertex + ix[k]);
            float wf; std::memcpy(&wf, &w[k], 4);
            tv[3 * (size_t)i + k] = make_float4(q[0], q[1], q[2], wf);
        }
        s->kinds_present[desc->materials[p.material].kind] = true;
    }
    // Node64 (sg_trace2.cuh): interior nodes with both children's bounds; leaves are folded into child refs.  One
    // conversion per BvhAggregate (the top level and every object definition); refs index the shared node64 / tri_verts arrays.
    std::vector<float4> n64;
    std::vector<DInstance> dinst(desc->n_instances);
    {
        struct Root { uint32_t ref; uint32_t depth; };
        bool too_deep = false;
        auto convert = [&](uint32_t node_base, uint32_t N, uint32_t prim_base, uint32_t n_direct_prims) -> Root {
            if (N == 0) {                                    // bare primitives without an aggregate: one leaf
                if (n_direct_prims == 0) return Root{kEmptyRef, 0};
                uint32_t last = prim_base + n_direct_prims - 1, wbits;
                std::memcpy(&wbits, &tv[3 * (size_t)last + 2].w, 4); wbits |= kLastInLeaf; std::memcpy(&tv[3 * (size_t)last + 2].w, &wbits, 4);
                return Root{kLeafBit | prim_base, 1};
            }
            const SgBvhNode* nodes = desc->nodes + node_base;
            std::vector<uint32_t> idx64(N, 0), depth(N, 0);
            const uint32_t first64 = (uint32_t)(n64.size() / 4);
            uint32_t n_interior = 0, max_depth = 1;
            for (uint32_t i = 0; i < N; ++i) if (nodes[i].n_prims == 0) idx64[i] = first64 + n_interior++;
            depth[0] = 1;
            n64.resize(n64.size() + (size_t)n_interior * 4);
            for (uint32_t i = 0; i < N; ++i) {
                const SgBvhNode& nd = nodes[i];
                if (depth[i] > max_depth) max_depth = depth[i];
                if (nd.n_prims > 0) {
                    uint32_t last = prim_base + nd.offset + nd.n_prims - 1, wbits;
                    std::memcpy(&wbits, &tv[3 * (size_t)last + 2].w, 4); wbits |= kLastInLeaf; std::memcpy(&tv[3 * (size_t)last + 2].w, &wbits, 4);
                    continue;
                }
                const uint32_t c[2] = {i + 1, nd.offset};
                uint32_t ref[2];
                for (int k = 0; k < 2; ++k) {
                    const SgBvhNode& ch = nodes[c[k]];
                    depth[c[k]] = depth[i] + 1;
                    ref[k] = ch.n_prims > 0 ? (kLeafBit | (prim_base + ch.offset)) : idx64[c[k]];
                }
                const SgBvhNode& a = nodes[c[0]]; const SgBvhNode& b = nodes[c[1]];
                float4* o = &n64[(size_t)idx64[i] * 4];
                float r0, r1, mt; uint32_t meta = nd.axis;
                std::memcpy(&r0, &ref[0], 4); std::memcpy(&r1, &ref[1], 4); std::memcpy(&mt, &meta, 4);
                o[0] = make_float4(a.bmin[0], a.bmin[1], a.bmin[2], a.bmax[0]);
                o[1] = make_float4(a.bmax[1], a.bmax[2], b.bmin[0], b.bmin[1]);
                o[2] = make_float4(b.bmin[2], b.bmax[0], b.bmax[1], b.bmax[2]);
                o[3] = make_float4(r0, r1, mt, 0.0f);
            }
            if (max_depth > 64) too_deep = true;
            return Root{nodes[0].n_prims > 0 ? (kLeafBit | (prim_base + nodes[0].offset)) : idx64[0], max_depth};
        };
        if (desc->n_primitives >= (1u << 28)) { g_err = "too many primitives"; return bail(SG_ERR_UNSUPPORTED); }
        const Root top = convert(0, n_top_nodes, 0, 0);
        uint32_t obj_depth = 0;
        std::vector<Root> oroot(desc->n_objects);
        for (uint32_t i = 0; i < desc->n_objects; ++i) {
            const SgObject& o = desc->objects[i];
            oroot[i] = convert(o.first_node, o.n_nodes, o.first_prim, o.n_prims);
            obj_depth = std::max(obj_depth, oroot[i].depth);
        }
        if (too_deep) { g_err = "BVH deeper than 64 levels: the reference's fixed traversal stack (aggregate.rs:90) would overflow"; return bail(SG_ERR_UNSUPPORTED); }
        for (uint32_t i = 0; i < desc->n_instances; ++i) {
            const SgInstance& I = desc->instances[i]; const SgObject& o = desc->objects[I.object];
            DInstance& D = dinst[i];
            std::memcpy(D.m, I.render_from_primitive, 12 * sizeof(float)); std::memcpy(D.mi, I.primitive_from_render, 12 * sizeof(float));
            const float* r3 = I.render_from_primitive + 12;
            if (r3[0] != 0.0f || r3[1] != 0.0f || r3[2] != 0.0f || r3[3] != 1.0f) { g_err = "instance transforms must be affine"; return bail(SG_ERR_UNSUPPORTED); }
            D.root_ref = oroot[I.object].ref; D.root_has_bounds = o.n_nodes > 0 ? 1u : 0u;
            for (int k = 0; k < 3; ++k) { D.bmin[k] = o.n_nodes ? desc->nodes[o.first_node].bmin[k] : 0.0f; D.bmax[k] = o.n_nodes ? desc->nodes[o.first_node].bmax[k] : 0.0f; }
        }
        const uint32_t N = n_top_nodes;
        // stack levels: the top-level tree, the deepest object tree, and one resume entry per instance entered from a multi-primitive leaf
        const uint32_t max_depth = top.depth + (desc->n_instances ? obj_depth + 1 : 0);
        s->ts.stack_depth = max_depth < 1 ? 1 : (int)max_depth;
        s->ts.root_ref = N == 0 ? kEmptyRef : top.ref;
        for (int k = 0; k < 3 && N; ++k) { s->ts.root_bmin[k] = desc->nodes[0].bmin[k]; s->ts.root_bmax[k] = desc->nodes[0].bmax[k]; }
        s->ts.scene_flags = desc->scene_flags;
        auto env_int = [](const char* name, int dflt) { const char* v = std::getenv(name); return v ? std::atoi(v) : dflt; };
        // Scheduling knobs of the vote loop (sg_wavefront.cuh trace_persistent), tuned per kernel family on the B200
        // (gpurun_out/r02_sweep*.log): triangle-only kernels refill at 14 waiting lanes and run interior steps in bursts of 4;
        // the instanced kernels meet a leaf event (an instance to enter) every few steps, so bursts only make lanes wait --
        // burst 1 and refill at 20: C4 276 -> 314 Mpaths/s (burst 2 / 3 / 8: 301 / 289 / 231).
        const bool general_kernels = desc->n_instances > 0 || desc->n_spheres > 0 || !pv.empty();
        s->ts.leaf_threshold = env_int("SG_LEAF_THRESHOLD", 6);
        s->ts.refill_threshold = env_int("SG_REFILL_THRESHOLD", general_kernels ? 20 : 14);
        s->ts.interior_burst = env_int("SG_INTERIOR_BURST", general_kernels ? 1 : 4);
        // Depth-0 closest-hit launch: refill only when ALL 32 lanes are done.  Camera rays of one warp are near-identical (pixel-major
        // wavefront), so nobody waits long -- and the warp's hits are then appended to the shade queue as one group of 32 consecutive
        // path slots, which keeps the depth-0 shade kernel's state accesses coalesced: C2 600 -> 629, C4 324 -> 350, C5 1709 -> 1741
        // Mpaths/s (gpurun_out/r02_sweep5_*.log; both the traversal and the shading time drop).
        s->ts.refill_threshold_d0 = env_int("SG_REFILL_THRESHOLD_D0", 32);
        s->ts.interior_burst_d0 = env_int("SG_INTERIOR_BURST_D0", s->ts.interior_burst);
        s->ts.prefetch = env_int("SG_PREFETCH", 0);
        // shared-memory part of the per-thread stack: 20 levels x 8 B x 128 threads = 20.5 KB -> 9 CTAs (36 warps, the register limit at 56 regs) per SM;
        // deeper levels (if the tree has them) spill to local memory (sg_trace2.cuh Stack)
        s->ts.smem_levels = std::min(s->ts.stack_depth, std::max(1, env_int("SG_SMEM_LEVELS", 20)));
        if (s->ts.stack_depth - s->ts.smem_levels > kSpillLevels) s->ts.smem_levels = s->ts.stack_depth - kSpillLevels;
        s->smem_closest = (size_t)s->ts.smem_levels * kTraceThreads * 8;
        s->smem_shadow = (size_t)s->ts.smem_levels * kTraceThreads * 4;
        // two-rays-per-lane kernels (SG_TRACE_DUAL=1, triangle-only scenes): two stacks + one parked ray per thread
        s->ts.dual_levels_closest = std::max(1, std::min(s->ts.stack_depth, env_int("SG_DUAL_LEVELS", 12)));
        s->ts.dual_levels_shadow = std::max(1, std::min(s->ts.stack_depth, env_int("SG_DUAL_LEVELS_SHADOW", 20)));
        s->ts.refill_threshold_dual = env_int("SG_DUAL_REFILL", 24);
        s->ts.leaf_threshold_dual = env_int("SG_DUAL_LEAF", 12);
        s->smem_closest_dual = (size_t)kTraceThreads * (2 * s->ts.dual_levels_closest * 8 + kParkWords * 4);
        s->smem_shadow_dual = (size_t)kTraceThreads * (2 * s->ts.dual_levels_shadow * 4 + kParkWords * 4);
    }
    s->instanced = desc->n_instances > 0 || desc->n_spheres > 0 || !pv.empty();      // anything that is not a triangle -> the general kernels
    if (s->instanced) {                          // + the parked render-space ray of a lane inside an instance (sg_trace2.cuh lane_save_ray)
        s->smem_closest += (size_t)kTraceThreads * 10 * sizeof(float);
        s->smem_shadow += (size_t)kTraceThreads * 10 * sizeof(float);
    }
    {
        float4* d_pv = nullptr;
        if ((rc = upload(pv.data(), pv.size(), &d_pv, s->owned)) != SG_OK) return bail(rc);
        s->ts.patch_verts = d_pv; d.patch_verts = d_pv;
    }
    {
        std::vector<DSphere> dsph(desc->n_spheres);
        for (uint32_t i = 0; i < desc->n_spheres; ++i) {
            const SgSphere& sp = desc->spheres[i]; DSphere& D = dsph[i];
            std::memcpy(D.m, sp.render_from_object, 12 * sizeof(float)); std::memcpy(D.mi, sp.object_from_render, 12 * sizeof(float));
            D.radius = sp.radius; D.z_min = sp.z_min; D.z_max = sp.z_max; D.theta_z_min = sp.theta_z_min; D.theta_z_max = sp.theta_z_max;
            D.phi_max = sp.phi_max; D.flags = sp.flags; D.pad = 0;
        }
        DSphere* d_sph = nullptr;
        if ((rc = upload(dsph.data(), dsph.size(), &d_sph, s->owned)) != SG_OK) return bail(rc);
        s->ts.spheres = d_sph; d.spheres = d_sph;
    }
    {
        DInstance* d_inst = nullptr;
        if ((rc = upload(dinst.data(), dinst.size(), &d_inst, s->owned)) != SG_OK) return bail(rc);
        s->ts.instances = d_inst; d.instances = d_inst; d.n_instances = desc->n_instances; d.scene_flags = desc->scene_flags;
    }
    float4* d_nodes = nullptr; float4* d_tv = nullptr; float4* d_n64 = nullptr;
    if ((rc = upload(n64.data(), n64.size(), &d_n64, s->owned)) != SG_OK) return bail(rc);
    s->ts.node64 = d_n64;
    if ((rc = upload((const float4*)desc->nodes, (size_t)desc->n_nodes * 2, &d_nodes, s->owned)) != SG_OK) return bail(rc);
    if ((rc = upload(tv.data(), tv.size(), &d_tv, s->owned)) != SG_OK) return bail(rc);
    d.nodes = d_nodes; d.tri_verts = d_tv; s->ts.tri_verts = d_tv;
    if (!tv.empty()) {                                      // degenerate-triangle flags (sg_scene.cuh kDegenerateBit)
        const uint32_t n_rec = (uint32_t)(tv.size() / 3);
        k_mark_degenerate<<<(n_rec + 255) / 256, 256, 0, g_dev[s->dev].stream>>>(d_tv, n_rec);
        cudaError_t e = cudaStreamSynchronize(g_dev[s->dev].stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return bail(fail(SG_ERR_CUDA, std::string("k_mark_degenerate: ") + cudaGetErrorString(e)));
    }
#define UP(field, src, n, T) { T* p__ = nullptr; if ((rc = upload((const T*)(src), (size_t)(n), &p__, s->owned)) != SG_OK) return bail(rc); d.field = p__; }
    UP(prims, desc->primitives, desc->n_primitives, SgPrimitive);
    UP(meshes, desc->meshes, desc->n_meshes, SgMesh);
    UP(indices, desc->indices, desc->n_indices, uint32_t);
    UP(p, desc->p, (size_t)desc->n_vertices * 3, float);
    UP(n, desc->n, desc->n ? (size_t)desc->n_vertices * 3 : 0, float);
    UP(uv, desc->uv, desc->uv ? (size_t)desc->n_vertices * 2 : 0, float);
    UP(s, desc->s, desc->s ? (size_t)desc->n_vertices * 3 : 0, float);
    {   // piecewise-linear spectra: tabulate find_interval at every integer wavelength (spectrum_get steps forward from there)
        std::vector<SgSpectrum> spectra_dev(desc->spectra, desc->spectra + desc->n_spectra);
        std::vector<uint16_t> lut;
        for (uint32_t i = 0; i < desc->n_spectra; ++i) {
            SgSpectrum& sp = spectra_dev[i];
            sp.pad = 0u;
            if (sp.kind != SG_SPECTRUM_PIECEWISE_LINEAR || sp.n < 2 || sp.n > 65535) continue;
            if ((uint64_t)sp.off_a + (uint64_t)sp.n > desc->n_pool || (uint64_t)sp.off_b + (uint64_t)sp.n > desc->n_pool)
                return bail(fail(SG_ERR_INVALID_ARGUMENT, "spectrum " + std::to_string(i) + ": samples outside spectrum_pool"));
            const size_t at = lut.size();
            lut.resize(at + kSpecLutBins);
            if (build_spectrum_lut(desc->spectrum_pool + sp.off_a, sp.n, lut.data() + at)) sp.pad = (uint32_t)at + 1u;
            else lut.resize(at);                            // unsorted knots: keep the binary search
        }
        SgSpectrum* p_sp = nullptr; uint16_t* p_lut = nullptr;
        if ((rc = upload(spectra_dev.data(), spectra_dev.size(), &p_sp, s->owned)) != SG_OK) return bail(rc);
        if ((rc = upload(lut.data(), lut.size(), &p_lut, s->owned)) != SG_OK) return bail(rc);
        d.spectra = p_sp; d.spec_lut = p_lut;
    }
    UP(pool, desc->spectrum_pool, desc->n_pool, float);
    UP(materials, desc->materials, desc->n_materials, SgMaterial);
    UP(lights, desc->lights, desc->n_lights, SgLight);
    UP(textures, desc->textures, desc->n_textures, SgTexture);
    UP(image_levels, desc->image_levels, desc->n_textures ? desc->n_image_levels : 0, SgImageLevel);
    UP(texels, desc->texels, (desc->n_textures || desc->n_env_maps) ? desc->n_texels : 0, float);
    UP(texture_mappings, desc->texture_mappings, desc->n_texture_mappings, SgTextureMapping);
    UP(texture_nodes, desc->texture_nodes, desc->n_texture_nodes, SgTextureNode);
    UP(material_textures, desc->material_textures, desc->material_textures ? desc->n_materials : 0, SgMaterialTextures);
    UP(env_maps, desc->env_maps, desc->n_env_maps, SgEnvMap);
    UP(mip_lut, desc->mip_filter_lut, desc->mip_filter_lut ? 128 : 0, float);
    UP(rgb2spec_scale, desc->rgb2spec_scale, need_rgb2spec ? desc->rgb2spec_res : 0, float);
    UP(rgb2spec_data, desc->rgb2spec_data, need_rgb2spec ? (size_t)9 * desc->rgb2spec_res * desc->rgb2spec_res * desc->rgb2spec_res : 0, float);
#undef UP
    for (uint32_t i = 0; i < desc->n_materials; ++i) if (desc->materials[i].kind == SG_MATERIAL_MIX) s->has_mix = true;
    if (s->has_mix)                                         // any material may come out of a mix: launch every kind that exists in the table
        for (uint32_t i = 0; i < desc->n_materials; ++i) if (desc->materials[i].kind != SG_MATERIAL_MIX) s->kinds_present[desc->materials[i].kind] = true;
    s->ts.queue_mask = 1u;                                 // Q_MISS
    for (int k = 0; k < 8; ++k) if (s->kinds_present[k]) s->ts.queue_mask |= 1u << (1 + k);
    if (s->has_mix) s->ts.queue_mask |= 1u << Q_MIX;
    d.n_textures = desc->n_textures; d.rgb2spec_res = need_rgb2spec ? desc->rgb2spec_res : 0;
    s->tex_path = desc->n_textures > 0;
    for (uint32_t i = 0; i < desc->n_materials; ++i)
        if ((desc->materials[i].flags & SG_MAT_HAS_DISPLACEMENT) && desc->materials[i].displacement != 0.0f) s->tex_path = true;
    for (uint32_t i = 0; i < desc->n_materials; ++i) if (desc->materials[i].kind >= 0 && desc->materials[i].kind < 8) s->materials_of_kind[desc->materials[i].kind]++;
    {
        static const int sort_env = [] { const char* v = std::getenv("SG_SORT_MATERIALS"); return v ? std::atoi(v) : 1; }();
        for (int kd = 0; kd < 8; ++kd) if (s->tex_path && sort_env && s->materials_of_kind[kd] > 1) s->sort_by_material = true;
        static const int stage_env = [] { const char* v = std::getenv("SG_STAGED_SHADING"); return v ? std::atoi(v) : 1; }();
        s->staged_shading = stage_env != 0 && s->tex_path && s->kinds_present[SG_MATERIAL_DIFFUSE];
    }
    for (uint32_t i = 0; i < desc->n_lights; ++i)               // lights only the general shade kernels handle (k_shade<.., LG = true>)
        if (desc->lights[i].kind != SG_LIGHT_DIFFUSE_AREA && desc->lights[i].kind != SG_LIGHT_UNIFORM_INFINITE) s->general_lights = true;
    d.n_nodes = desc->n_nodes; d.n_prims = desc->n_primitives; d.n_lights = desc->n_lights; d.n_materials = desc->n_materials;
    d.n_infinite = 0;
    for (uint32_t i = 0; i < desc->n_lights; ++i) if (desc->lights[i].kind == SG_LIGHT_UNIFORM_INFINITE || desc->lights[i].kind == SG_LIGHT_IMAGE_INFINITE) {
        if (d.n_infinite >= 4) { g_err = "more than 4 infinite lights"; return bail(SG_ERR_UNSUPPORTED); }
        d.infinite_ids[d.n_infinite++] = (int32_t)i;
    }
    {   // emitter triangles pre-gathered per light (light.rs:524-535: each area light owns its Shape)
        std::vector<float4> lv((size_t)desc->n_lights * 3, make_float4(0, 0, 0, 0));
        for (uint32_t i = 0; i < desc->n_lights; ++i) {
            const SgLight& L = desc->lights[i];
            if (L.kind != SG_LIGHT_DIFFUSE_AREA) continue;
            if (L.mesh >= desc->n_meshes || L.tri >= desc->meshes[L.mesh].n_triangles) { g_err = "light references out-of-range triangle"; return bail(SG_ERR_INVALID_ARGUMENT); }
            const SgMesh& m = desc->meshes[L.mesh];
            const uint32_t* ix = desc->indices + m.first_index + 3 * (size_t)L.tri;
            for (int k = 0; k < 3; ++k) {
                const float* q = desc->p + 3 * (size_t)(m.first_vertex + ix[k]);
                float wf = 0.0f; if (k == 0) { uint32_t fl = m.flags; std::memcpy(&wf, &fl, 4); }
                lv[3 * (size_t)i + k] = make_float4(q[0], q[1], q[2], wf);
            }
        }
        // patch emitters: the record of the light's patch in patch_verts (records were assigned in primitive order above)
        {
            uint32_t rec = 0; uint32_t found = 0, wanted = 0;
            for (uint32_t i = 0; i < desc->n_lights; ++i) if (desc->lights[i].kind == SG_LIGHT_DIFFUSE_AREA_PATCH) ++wanted;
            for (uint32_t i = 0; i < desc->n_primitives && wanted; ++i) {
                const SgPrimitive& p = desc->primitives[i];
                if (p.mesh == SG_PRIM_INSTANCE || p.mesh == SG_PRIM_SPHERE || !(desc->meshes[p.mesh].flags & SG_MESH_BILINEAR)) continue;
                if (p.light >= 0) { float wf; std::memcpy(&wf, &rec, 4); lv[3 * (size_t)p.light].w = wf; ++found; }
                ++rec;
            }
            if (found != wanted) { g_err = "every SG_LIGHT_DIFFUSE_AREA_PATCH light needs exactly one patch primitive that points at it"; return bail(SG_ERR_INVALID_ARGUMENT); }
        }
        float4* d_lv = nullptr;
        if ((rc = upload(lv.data(), lv.size(), &d_lv, s->owned)) != SG_OK) return bail(rc);
        d.light_verts = d_lv;
    }
    d.camera = desc->camera; d.film = desc->film;
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(DevStats)) != cudaSuccess) { g_err = "cudaMalloc stats"; return bail(SG_ERR_OUT_OF_MEMORY); }
    s->owned.push_back(p); s->d_stats = (DevStats*)p;
    if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) { g_err = "cudaMalloc cursor"; return bail(SG_ERR_OUT_OF_MEMORY); }
    s->owned.push_back(p); s->d_cursor = (unsigned long long*)p;
    *out = s;
    return SG_OK;
}

int sg_scene_destroy(SgScene* s) {
    if (!s) return SG_OK;
    for (SgScene* peer : s->peers) sg_scene_destroy(peer);
    s->peers.clear();
    DeviceGuard guard__(s->dev);
    cudaDeviceSynchronize();
    for (Workspace& w : s->ws) w.release();
    for (void* p : s->owned) cudaFree(p);
    if (s->d_film) cudaFree(s->d_film);
    if (s->h_film) cudaFreeHost(s->h_film);
    delete s;
    return SG_OK;
}

// The wavefront loop of one scene replica on its own device: samples [rp->sample_begin, rp->sample_end) of every pixel are
// ADDED into d_film.  With `reduce` the film is then summed onto rank 0 of the process communicator in stream order.
// Blocks until the stream has drained (the stats are read back).
static int render_on(SgScene* s, const SgRenderParams* rp, void* d_film, SgStats* stats, cudaStream_t stream, bool reduce) {
    if (!s || !rp || !d_film) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (rp->sample_end < rp->sample_begin || rp->sample_begin < 0) return fail(SG_ERR_INVALID_ARGUMENT, "bad sample range");
    if (rp->max_depth < 0 || rp->max_depth > 254) return fail(SG_ERR_INVALID_ARGUMENT, "max_depth out of range");
    const bool force_diffuse = (rp->option_flags & SG_OPT_FORCE_DIFFUSE) != 0;
    if (force_diffuse && rp->integrator != SG_INTEGRATOR_PATH) return fail(SG_ERR_UNSUPPORTED, "force_diffuse is on the GPU path for the path integrator only");
    if (rp->integrator < SG_INTEGRATOR_PATH || rp->integrator > SG_INTEGRATOR_RANDOM_WALK) return fail(SG_ERR_INVALID_ARGUMENT, "unknown integrator kind");
    DeviceGuard guard__(s->dev);
    const int num_sms = g_dev[s->dev].num_sms;
    const bool count = (rp->flags & SG_RENDER_COUNT_VISITS) != 0;
    const uint64_t npix = s->n_pixels();
    const uint64_t total = npix * (uint64_t)(rp->sample_end - rp->sample_begin);
    uint64_t cap64 = rp->max_paths_in_flight > 0 ? (uint64_t)rp->max_paths_in_flight : (1ull << 26);   // 64 Mi paths = 18.5 GB of 180 GB
    if (cap64 > total) cap64 = total;
    if (cap64 == 0) cap64 = 1;
    // Several wavefronts in flight (default 2): batches are dealt round-robin to the caller's stream and auxiliary ones, each with
    // its own path state and queues.  The persistent traversal kernels and the grid-stride shade kernels of one batch leave SMs
    // idle in their tails and in the small late-depth launches; the other batches' kernels fill them (the hardware schedules CTAs
    // of all streams).  Film and statistics updates are atomic, so the batches commute.  A one-batch job below 32 Mi paths is cut
    // into equal parts (C1: +5 %); a bigger single batch stays whole (C2 at 64 Mi paths: halves lose 2 % -- large wavefronts
    // amortise their own tails).  Per-kernel timing / visit counting keep one stream (their events bracket one kernel at a time).
    const int overlap_env = [] { const char* v = std::getenv("SG_OVERLAP"); return v ? std::atoi(v) : 2; }();      // read per render (tests flip it)
    const bool time_or_count = (rp->flags & (SG_RENDER_TIME_KERNELS | SG_RENDER_COUNT_VISITS)) != 0;
    int n_inflight = (overlap_env >= 2 && !time_or_count && total >= (1ull << 21)) ? std::min(overlap_env, kMaxWavefronts) : 1;
    if (n_inflight > 1 && cap64 >= total) {
        if (total < (1ull << 25)) cap64 = (total + n_inflight - 1) / n_inflight; else n_inflight = 1;
    }
    const uint32_t capacity = (uint32_t)cap64;
    int rc = SG_OK;
    for (int i = 0; i < n_inflight; ++i) if ((rc = ensure_workspace(s, i, capacity, rp->max_depth)) != SG_OK) return rc;
    RenderConst k;
    k.seed = rp->seed; k.sample_begin = rp->sample_begin; k.spp = rp->samples_per_pixel; k.max_depth = rp->max_depth;
    k.regularize = rp->regularize; k.option_flags = rp->option_flags;
    k.win_x0 = s->d.film.pixel_bounds[0]; k.win_y0 = s->d.film.pixel_bounds[1];
    k.win_w = s->d.film.pixel_bounds[2] - k.win_x0; k.win_h = s->d.film.pixel_bounds[3] - k.win_y0;
    k.full_res_x = s->d.film.full_resolution[0];
    k.n_samples = rp->sample_end - rp->sample_begin;
    { const char* v = std::getenv("SG_PATH_ORDER"); k.path_order = v ? std::atoi(v) : 1; }
    k.integrator = rp->integrator; k.integrator_flags = rp->integrator_flags;
    { const char* v = std::getenv("SG_SHADE_SYNC"); k.shade_sync = v ? std::atoi(v) : 12; }           // barriers around sample_ld (see k_shade)
    { const char* v = std::getenv("SG_SHADE_SYNC_TEX"); k.shade_sync_tex = v ? std::atoi(v) : 0; }
    const bool path_integrator = rp->integrator == SG_INTEGRATOR_PATH;
    if (k.n_samples == 0) k.n_samples = 1;
    const bool time_trace = (rp->flags & SG_RENDER_TIME_KERNELS) != 0;
    const int n_depths = rp->max_depth + 1;
    // A/B switch: two rays per lane (triangle-only scenes, no visit counting); bit 0 = closest-hit launches, bit 1 = any-hit launches
    const int dual_env = [] { const char* v = std::getenv("SG_TRACE_DUAL"); return v ? std::atoi(v) : 0; }();
    const bool dual_ok = !count && !s->instanced && s->ts.stack_depth - std::min(s->ts.dual_levels_closest, s->ts.dual_levels_shadow) <= kSpillDual;
    const bool dual_c = dual_ok && (dual_env & 1), dual_s = dual_ok && (dual_env & 2);
    const size_t smc = dual_c ? s->smem_closest_dual : s->smem_closest, sms = dual_s ? s->smem_shadow_dual : s->smem_shadow;
    const TraceKernel kern_closest = dual_c ? trace_kernel_dual(false) : trace_kernel(false, count, s->instanced);
    const TraceKernel kern_shadow = dual_s ? trace_kernel_dual(true) : trace_kernel(true, count, s->instanced);
    const int grid_closest = persistent_grid(num_sms, (const void*)kern_closest, kTraceThreads, smc);
    const int grid_shadow = persistent_grid(num_sms, (const void*)kern_shadow, kTraceThreads, sms);
    const int shade_grid = num_sms * [] { const char* v = std::getenv("SG_SHADE_GRID"); return v && std::atoi(v) > 0 ? std::atoi(v) : 8; }();   // grid-stride shade kernels: CTAs of 128 threads per SM
    const int shade_grid_lean = shade_grid * 128 / SG_SHADE_THREADS, shade_grid_tex = shade_grid * 128 / SG_SHADE_THREADS_TEX;
    CU(cudaMemsetAsync(s->d_stats, 0, sizeof(DevStats), stream));
    EventBag bag;                           // events die on every return path
    cudaEvent_t ev0 = bag.make(), ev1 = bag.make(), ev2 = bag.make();
    if (!ev0 || !ev1 || !ev2) return fail(SG_ERR_CUDA, "cudaEventCreate failed");
    std::vector<cudaEvent_t> tev, sev;      // event pairs around closest-hit / any-hit launches
    auto mark = [&](std::vector<cudaEvent_t>& v) -> int { cudaEvent_t a = bag.make(); if (!a) return fail(SG_ERR_CUDA, "cudaEventCreate failed"); CU(cudaEventRecord(a, stream)); v.push_back(a); return SG_OK; };
    uint64_t launches = 0, closest_launches = 0, shadow_launches = 0;
    // Inside a batch the any-hit (shadow) traversal of depth d runs on a side stream, concurrently with the closest-hit
    // traversal of depth d + 1: they touch disjoint state (shadow: sh_*, L; closest: ray_*, hit_*, the queues of d + 1) and
    // each is a persistent kernel whose tail leaves SMs idle.  The shade kernels of d + 1 (which read L) wait for shadow(d).
    static const int side_env = [] { const char* v = std::getenv("SG_SHADOW_SIDE_STREAM"); return v ? std::atoi(v) : 1; }();
    // Only for one-wavefront jobs (C2: +1.4 %); with two wavefronts in flight the other batch already fills those tails and the
    // side stream measured neutral to -1 % (gpurun_out/r02_side.log).
    const bool side_stream = side_env != 0 && !time_or_count && n_inflight == 1;
    cudaStream_t lanes[kMaxWavefronts] = {stream, g_dev[s->dev].aux[0]};
    cudaStream_t sides[kMaxWavefronts] = {g_dev[s->dev].aux[1], g_dev[s->dev].aux[2]};
    cudaEvent_t ev_shaded[kMaxWavefronts], ev_shadowed[kMaxWavefronts];
    for (int i = 0; i < kMaxWavefronts; ++i) { ev_shaded[i] = bag.make(); ev_shadowed[i] = bag.make(); if (!ev_shaded[i] || !ev_shadowed[i]) return fail(SG_ERR_CUDA, "cudaEventCreate failed"); }
    CU(cudaEventRecord(ev0, stream));
    for (int i = 1; i < n_inflight; ++i) CU(cudaStreamWaitEvent(lanes[i], ev0, 0));   // the auxiliary wavefront starts after the caller's prior work too
    uint64_t batch = 0;
    for (uint64_t first = 0; first < total; first += capacity, ++batch) {
        const uint32_t cnt = (uint32_t)((total - first) < capacity ? (total - first) : capacity);
        const int lane = (int)(batch % (uint64_t)n_inflight);
        cudaStream_t stream = lanes[lane];                                     // shadows the caller's stream inside the batch
        cudaStream_t side = side_stream ? sides[lane] : stream;
        Workspace& w = s->ws[lane];
        bool shadow_pending = false;
        CU(cudaMemsetAsync(w.q.counters, 0, (size_t)(rp->max_depth + 3) * C_STRIDE * sizeof(uint32_t), stream));
        // first = first_ord * n_samples + first_rem (pixel-order index and sample offset of the batch's first path)
        k_generate<<<(cnt + 255) / 256, 256, 0, stream>>>(s->d, w.st, w.q, k, first, (uint32_t)(first / (uint64_t)k.n_samples), (uint32_t)(first % (uint64_t)k.n_samples), cnt); ++launches;
        for (int depth = 0; depth < n_depths; ++depth) {
            if (time_trace && (rc = mark(tev)) != SG_OK) return rc;
            kern_closest<<<grid_closest, kTraceThreads, smc, stream>>>(s->d, s->ts, w.st, w.q, depth, s->d_stats);
            ++launches; ++closest_launches;
            if (time_trace && (rc = mark(tev)) != SG_OK) return rc;
            if (w.q.ray_kind) {                                    // build the shade queues of this depth in ray-queue order
                const unsigned nblk = std::min((unsigned)((cnt + kQueueBlock - 1) / kQueueBlock), (unsigned)num_sms * 8u);
                k_queue_count<<<nblk, 256, 0, stream>>>(w.q, depth);
                k_queue_scan<<<1, 1024, 0, stream>>>(w.q, depth, s->ts.queue_mask);
                k_queue_scatter<<<nblk, 256, 0, stream>>>(w.q, depth);
                launches += 3;
            }
            if (shadow_pending) { CU(cudaStreamWaitEvent(stream, ev_shadowed[lane], 0)); shadow_pending = false; }   // L of depth - 1 is final
            if (s->d.n_infinite > 0) { k_shade_miss<<<shade_grid, 128, 0, stream>>>(s->d, w.st, w.q, k, depth); ++launches; }
            if (s->has_mix) { resolve_mix_kernel(s->tex_path)<<<shade_grid, 128, 0, stream>>>(s->d, w.st, w.q, k, depth); ++launches; }
            for (int kind = 0; kind <= SG_MATERIAL_COATED_CONDUCTOR; ++kind) {
                if (!s->kinds_present[kind]) continue;
                Queues qs = w.q;
                if (w.q.sorted && (s->materials_of_kind[kind] > 1 || s->has_mix)) {     // re-order this kind's queue by material id
                    CU(cudaMemsetAsync(w.q.sort_hist, 0, 128 * sizeof(uint32_t), stream));
                    k_sort_queue_count<<<num_sms * 4, 256, 0, stream>>>(s->d, w.st, w.q, depth, 1 + kind);
                    k_sort_queue_scatter<<<num_sms * 4, 256, 0, stream>>>(s->d, w.st, w.q, depth, 1 + kind);
                    launches += 2;
                    qs.shade[1 + kind] = w.q.sorted;
                }
                if (s->staged_shading && path_integrator && !force_diffuse && shade_kernel_stage1(kind)) {
                    shade_kernel_stage1(kind)<<<shade_grid_tex, SG_SHADE_THREADS_TEX, 0, stream>>>(s->d, w.st, qs, k, depth);
                    shade_kernel_stage2(kind, s->general_lights)<<<shade_grid_lean, SG_SHADE_THREADS, 0, stream>>>(s->d, w.st, qs, k, depth);
                    launches += 2;
                    continue;
                }
                const bool lean = !force_diffuse && path_integrator && !s->general_lights && !s->tex_path;     // shade_kernel_lean: TEX = false
                shade_kernel(kind, s->tex_path, s->general_lights, path_integrator, force_diffuse)<<<lean ? shade_grid_lean : shade_grid_tex, lean ? SG_SHADE_THREADS : SG_SHADE_THREADS_TEX, 0, stream>>>(s->d, w.st, qs, k, depth);
                ++launches;
            }
            if (depth < rp->max_depth && s->d.n_lights > 0) {
                if (time_trace && (rc = mark(sev)) != SG_OK) return rc;
                if (side != stream) { CU(cudaEventRecord(ev_shaded[lane], stream)); CU(cudaStreamWaitEvent(side, ev_shaded[lane], 0)); }
                kern_shadow<<<grid_shadow, kTraceThreads, sms, side>>>(s->d, s->ts, w.st, w.q, depth, s->d_stats);
                ++launches; ++shadow_launches;
                if (side != stream) { CU(cudaEventRecord(ev_shadowed[lane], side)); shadow_pending = true; }
                if (time_trace && (rc = mark(sev)) != SG_OK) return rc;
            }
        }
        if (shadow_pending) CU(cudaStreamWaitEvent(stream, ev_shadowed[lane], 0));
        k_film<<<(cnt + 255) / 256, 256, 0, stream>>>(s->d, w.st, cnt, (double*)d_film); ++launches;
        k_accum_stats<<<1, 1, 0, stream>>>(w.q.counters, n_depths, s->d_stats); ++launches;
    }
    for (int i = 1; i < n_inflight; ++i) {                                     // join: the caller's stream continues after every wavefront
        cudaEvent_t e = bag.make();
        if (!e) return fail(SG_ERR_CUDA, "cudaEventCreate failed");
        CU(cudaEventRecord(e, lanes[i]));
        CU(cudaStreamWaitEvent(stream, e, 0));
    }
    CU(cudaEventRecord(ev1, stream));
    const bool do_reduce = reduce && g_proc_comm && g_proc_nranks > 1;
    if (do_reduce) NC(g_nccl.Reduce(d_film, d_film, (size_t)npix * 4, ncclDouble, ncclSum, 0, g_proc_comm, stream));
    CU(cudaEventRecord(ev2, stream));
    CU(cudaEventSynchronize(ev2));
    CU(cudaGetLastError());
    float ms = 0.0f, rms = 0.0f;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    if (do_reduce) CU(cudaEventElapsedTime(&rms, ev1, ev2));
    double closest_ms = 0.0, shadow_ms = 0.0;
    for (size_t i = 0; i + 1 < tev.size(); i += 2) { float t = 0.0f; cudaEventElapsedTime(&t, tev[i], tev[i + 1]); closest_ms += t; }
    for (size_t i = 0; i + 1 < sev.size(); i += 2) { float t = 0.0f; cudaEventElapsedTime(&t, sev[i], sev[i + 1]); shadow_ms += t; }
    if (time_trace && std::getenv("SG_DEBUG_TIMING")) {      // per-launch device times, one line per render call
        std::fprintf(stderr, "[sg] closest ms:");
        for (size_t i = 0; i + 1 < tev.size(); i += 2) { float t = 0.0f; cudaEventElapsedTime(&t, tev[i], tev[i + 1]); std::fprintf(stderr, " %.3f", t); }
        std::fprintf(stderr, " | shadow ms:");
        for (size_t i = 0; i + 1 < sev.size(); i += 2) { float t = 0.0f; cudaEventElapsedTime(&t, sev[i], sev[i + 1]); std::fprintf(stderr, " %.3f", t); }
        std::fprintf(stderr, " | total %.3f\n", ms);
    }
    if (stats) {
        DevStats h;
        CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        std::memset(stats, 0, sizeof *stats);
        stats->camera_paths = total; stats->closest_hit_rays = h.closest; stats->shadow_rays = h.shadow;
        stats->nodes_visited = h.nodes; stats->tris_tested = h.tris; stats->kernel_launches = launches;
        stats->render_ms = ms; stats->trace_ms = closest_ms + shadow_ms;
        stats->closest_nodes = h.nodes_closest; stats->closest_tris = h.tris_closest;
        stats->closest_launches = closest_launches; stats->shadow_launches = shadow_launches;
        stats->closest_ms = closest_ms; stats->shadow_ms = shadow_ms;
        stats->reduce_ms = rms; stats->n_devices = (uint32_t)g_proc_nranks; stats->rank = (uint32_t)g_proc_rank;
    }
    return SG_OK;
}

// this rank's share of the call's sample range (SG_RENDER_SPLIT_SAMPLES with a process communicator)
static SgRenderParams split_for_rank(const SgRenderParams& rp) {
    SgRenderParams p = rp;
    if ((rp.flags & SG_RENDER_SPLIT_SAMPLES) && g_proc_comm && g_proc_nranks > 1)
        sg_sample_range_for_rank(rp.sample_begin, rp.sample_end, g_proc_rank, g_proc_nranks, &p.sample_begin, &p.sample_end);
    return p;
}

int sg_render_device(SgScene* s, const SgRenderParams* rp, void* d_film, SgStats* stats, void* stream_v) {
    if (g_dev.empty()) return fail(SG_ERR_NOT_INITIALIZED, "sg_init has not been called");
    if (!s || !rp || !d_film) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    // NULL = the legacy default stream: the caller's zeroing of d_film (and whatever reads it next) is ordered there
    const SgRenderParams p = split_for_rank(*rp);
    return render_on(s, &p, d_film, stats, (cudaStream_t)stream_v, (rp->flags & SG_RENDER_REDUCE_FILM) != 0);
}

static int ensure_film(SgScene* s, size_t npix, bool host) {
    DeviceGuard guard__(s->dev);
    if (s->film_pixels < npix) {
        if (s->d_film) cudaFree(s->d_film);
        s->d_film = nullptr; s->film_pixels = 0;
        CU(cudaMalloc((void**)&s->d_film, npix * sizeof(SgFilmPixel)));
        s->film_pixels = npix;
    }
    if (host && s->h_film_pixels < npix) {                  // pinned staging, allocated once per scene
        if (s->h_film) cudaFreeHost(s->h_film);
        s->h_film = nullptr; s->h_film_pixels = 0;
        CU(cudaHostAlloc((void**)&s->h_film, npix * sizeof(SgFilmPixel), cudaHostAllocDefault));
        s->h_film_pixels = npix;
    }
    return SG_OK;
}

int sg_render(SgScene* s, const SgRenderParams* rp, SgFilmPixel* film, SgStats* stats) {
    if (g_dev.empty()) return fail(SG_ERR_NOT_INITIALIZED, "sg_init has not been called");
    if (!s || !rp) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    const bool proc_reduce = (rp->flags & SG_RENDER_REDUCE_FILM) && g_proc_comm && g_proc_nranks > 1;
    const bool root = !proc_reduce || g_proc_rank == 0;     // only the root's host film is written
    if (root && !film) return fail(SG_ERR_INVALID_ARGUMENT, "null film");
    const size_t npix = (size_t)s->n_pixels();
    const int n_dev = 1 + (int)s->peers.size();
    DeviceGuard guard__(s->dev);
    cudaStream_t stream0 = g_dev[s->dev].stream;
    int rc = ensure_film(s, npix, root);
    if (rc != SG_OK) return rc;
    SgStats st0; std::memset(&st0, 0, sizeof st0);
    double reduce_ms = 0.0;
    if (n_dev == 1) {
        CU(cudaMemsetAsync(s->d_film, 0, npix * sizeof(SgFilmPixel), stream0));
        const SgRenderParams p = split_for_rank(*rp);
        rc = render_on(s, &p, s->d_film, &st0, stream0, proc_reduce);
        if (rc != SG_OK) return rc;
        reduce_ms = st0.reduce_ms;
    } else {
        // single process, n GPUs: one host thread per device renders that device's share of the sample range into its own
        // film (integrator.rs:235-245: the reference fans tiles out over rayon's pool); then ONE ncclReduce onto devices[0]
        std::vector<SgScene*> rep(n_dev); rep[0] = s;
        for (int i = 1; i < n_dev; ++i) rep[i] = s->peers[i - 1];
        std::vector<SgStats> st(n_dev); std::vector<int> rcs(n_dev, SG_OK); std::vector<std::string> errs(n_dev);
        auto work = [&](int i) {
            SgScene* r = rep[i];
            DeviceGuard g(r->dev);
            cudaStream_t stream = g_dev[r->dev].stream;
            int c = ensure_film(r, npix, false);
            if (c == SG_OK && cudaMemsetAsync(r->d_film, 0, npix * sizeof(SgFilmPixel), stream) != cudaSuccess) c = fail(SG_ERR_CUDA, "cudaMemsetAsync(film)");
            if (c == SG_OK) {
                SgRenderParams p = *rp;
                sg_sample_range_for_rank(rp->sample_begin, rp->sample_end, i, n_dev, &p.sample_begin, &p.sample_end);
                c = render_on(r, &p, r->d_film, &st[i], stream, false);
            }
            rcs[i] = c; if (c != SG_OK) errs[i] = g_err;
        };
        std::vector<std::thread> th;
        for (int i = 1; i < n_dev; ++i) th.emplace_back(work, i);
        work(0);
        for (auto& t : th) t.join();
        for (int i = 0; i < n_dev; ++i) if (rcs[i] != SG_OK) return fail(rcs[i], "device " + std::to_string(g_dev[rep[i]->dev].id) + ": " + errs[i]);
        st0 = st[0];
        for (int i = 1; i < n_dev; ++i) {
            st0.camera_paths += st[i].camera_paths; st0.closest_hit_rays += st[i].closest_hit_rays; st0.shadow_rays += st[i].shadow_rays;
            st0.nodes_visited += st[i].nodes_visited; st0.tris_tested += st[i].tris_tested; st0.kernel_launches += st[i].kernel_launches;
            st0.closest_nodes += st[i].closest_nodes; st0.closest_tris += st[i].closest_tris;
            st0.closest_launches += st[i].closest_launches; st0.shadow_launches += st[i].shadow_launches;
            st0.render_ms = std::max(st0.render_ms, st[i].render_ms); st0.trace_ms = std::max(st0.trace_ms, st[i].trace_ms);
            st0.closest_ms = std::max(st0.closest_ms, st[i].closest_ms); st0.shadow_ms = std::max(st0.shadow_ms, st[i].shadow_ms);
        }
        EventBag bag; cudaEvent_t e0 = bag.make(), e1 = bag.make();
        if (!e0 || !e1) return fail(SG_ERR_CUDA, "cudaEventCreate failed");
        CU(cudaEventRecord(e0, stream0));
        NC(g_nccl.GroupStart());
        for (int i = 0; i < n_dev; ++i) {
            const ncclResult_t r = g_nccl.Reduce(rep[i]->d_film, rep[i]->d_film, npix * 4, ncclDouble, ncclSum, 0, g_dev[rep[i]->dev].comm, g_dev[rep[i]->dev].stream);
            if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(SG_ERR_NCCL, std::string("ncclReduce: ") + g_nccl.GetErrorString(r)); }
        }
        NC(g_nccl.GroupEnd());
        CU(cudaEventRecord(e1, stream0));
        for (int i = 1; i < n_dev; ++i) { DeviceGuard g(rep[i]->dev); CU(cudaStreamSynchronize(g_dev[rep[i]->dev].stream)); }
        CU(cudaEventSynchronize(e1));
        float t = 0.0f; CU(cudaEventElapsedTime(&t, e0, e1)); reduce_ms = t;
    }
    double d2h_ms = 0.0;
    if (root) {
        // D2H through the pinned staging buffer, then accumulate into the caller's film
        EventBag bag; cudaEvent_t e0 = bag.make(), e1 = bag.make();
        if (!e0 || !e1) return fail(SG_ERR_CUDA, "cudaEventCreate failed");
        CU(cudaEventRecord(e0, stream0));
        CU(cudaMemcpyAsync(s->h_film, s->d_film, npix * sizeof(SgFilmPixel), cudaMemcpyDeviceToHost, stream0));
        CU(cudaEventRecord(e1, stream0));
        CU(cudaStreamSynchronize(stream0));
        float t = 0.0f; CU(cudaEventElapsedTime(&t, e0, e1)); d2h_ms = t;
        const double* src = reinterpret_cast<const double*>(s->h_film);
        double* dst = reinterpret_cast<double*>(film);
        const bool overwrite = (rp->flags & SG_RENDER_OVERWRITE_FILM) != 0;
        const size_t n = 4 * npix;
        const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        auto work = [&](unsigned t) {
            const size_t b = n * t / nt, e = n * (t + 1) / nt;
            if (overwrite) std::memcpy(dst + b, src + b, (e - b) * sizeof(double));
            else for (size_t i = b; i < e; ++i) dst[i] += src[i];
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
    }
    if (stats) {
        *stats = st0;
        stats->reduce_ms = reduce_ms; stats->d2h_ms = d2h_ms;
        stats->n_devices = n_dev > 1 ? (uint32_t)n_dev : (uint32_t)g_proc_nranks;
        stats->rank = (uint32_t)g_proc_rank;
    }
    return SG_OK;
}

int sg_trace_device(SgScene* s, int64_t n, const void* d_o, const void* d_d, const void* d_t_max, int any_hit, void* d_out,
                    SgStats* stats, void* stream_v) {
    if (g_dev.empty()) return fail(SG_ERR_NOT_INITIALIZED, "sg_init has not been called");
    if (!s || n < 0 || (n > 0 && (!d_o || !d_d || !d_t_max || !d_out))) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard__(s->dev);
    cudaStream_t stream = (cudaStream_t)stream_v;            // NULL = the legacy default stream (see sg_render_device)
    const bool count = stats != nullptr;
    CU(cudaMemsetAsync(s->d_stats, 0, sizeof(DevStats), stream));
    CU(cudaMemsetAsync(s->d_cursor, 0, sizeof(unsigned long long), stream));
    EventBag bag; cudaEvent_t ev0 = bag.make(), ev1 = bag.make();
    if (!ev0 || !ev1) return fail(SG_ERR_CUDA, "cudaEventCreate failed");
    CU(cudaEventRecord(ev0, stream));
    if (n > 0) {
        const size_t sm = any_hit ? s->smem_shadow : s->smem_closest;
        const TraceRaysKernel kern = trace_rays_kernel(any_hit != 0, count, s->instanced);
        const int g = persistent_grid(g_dev[s->dev].num_sms, (const void*)kern, kTraceThreads, sm);
        kern<<<g, kTraceThreads, sm, stream>>>(s->d, s->ts, (long long)n, (const float*)d_o, (const float*)d_d, (const float*)d_t_max, (SgHit*)d_out, s->d_cursor, s->d_stats);
    }
    CU(cudaEventRecord(ev1, stream));
    CU(cudaEventSynchronize(ev1));
    CU(cudaGetLastError());
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    if (stats) {
        DevStats h;
        CU(cudaMemcpy(&h, s->d_stats, sizeof h, cudaMemcpyDeviceToHost));
        std::memset(stats, 0, sizeof *stats);
        if (any_hit) stats->shadow_rays = (uint64_t)n; else stats->closest_hit_rays = (uint64_t)n;
        stats->nodes_visited = h.nodes; stats->tris_tested = h.tris; stats->kernel_launches = n > 0 ? 1 : 0;
        stats->render_ms = ms; stats->trace_ms = ms;
        if (any_hit) { stats->shadow_ms = ms; stats->shadow_launches = n > 0 ? 1 : 0; }
        else { stats->closest_ms = ms; stats->closest_launches = n > 0 ? 1 : 0; stats->closest_nodes = h.nodes; stats->closest_tris = h.tris; }
    }
    return SG_OK;
}

int sg_trace(SgScene* s, int64_t n, const float* o, const float* d, const float* t_max, int any_hit, SgHit* out, SgStats* stats) {
    ENTER(0);
    if (!s || n < 0 || (n > 0 && (!o || !d || !t_max || !out))) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) { if (stats) std::memset(stats, 0, sizeof *stats); return SG_OK; }
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr; SgHit* d_out = nullptr;
    int rc = SG_OK;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_out); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CUX(cudaMalloc((void**)&d_o, (size_t)n * 12)); CUX(cudaMalloc((void**)&d_d, (size_t)n * 12));
    CUX(cudaMalloc((void**)&d_t, (size_t)n * 4)); CUX(cudaMalloc((void**)&d_out, (size_t)n * sizeof(SgHit)));
    CUX(cudaMemcpyAsync(d_o, o, (size_t)n * 12, cudaMemcpyHostToDevice, g_stream));
    CUX(cudaMemcpyAsync(d_d, d, (size_t)n * 12, cudaMemcpyHostToDevice, g_stream));
    CUX(cudaMemcpyAsync(d_t, t_max, (size_t)n * 4, cudaMemcpyHostToDevice, g_stream));
    rc = sg_trace_device(s, n, d_o, d_d, d_t, any_hit, d_out, stats, g_stream);
    if (rc == SG_OK) CUX(cudaMemcpy(out, d_out, (size_t)n * sizeof(SgHit), cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return rc;
}

int sg_sampler_fill(uint64_t seed, int raw, uint32_t pixel_index, uint32_t sample_index, int64_t n, float* out) {
    ENTER(0);
    if (n < 0 || (n > 0 && !out)) return fail(SG_ERR_INVALID_ARGUMENT, "bad arguments");
    if (n == 0) return SG_OK;
    float* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)n * 4));
    k_sampler_fill<<<1, 1, 0, g_stream>>>(seed, raw, pixel_index, sample_index, (long long)n, d);
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(SG_ERR_CUDA, cudaGetErrorString(e));
    return SG_OK;
}

int sg_camera_rays(SgScene* s, const SgRenderParams* rp, int64_t n, const int32_t* pixel_xy, const int32_t* sample_index,
                   float* out_rays, float* out_lambda) {
    ENTER(0);
    if (!s || !rp || n < 0 || (n > 0 && (!pixel_xy || !sample_index || !out_rays || !out_lambda))) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return SG_OK;
    int *d_xy = nullptr, *d_si = nullptr; float *d_r = nullptr, *d_l = nullptr;
    auto cleanup = [&]() { cudaFree(d_xy); cudaFree(d_si); cudaFree(d_r); cudaFree(d_l); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CUX(cudaMalloc((void**)&d_xy, (size_t)n * 8)); CUX(cudaMalloc((void**)&d_si, (size_t)n * 4));
    CUX(cudaMalloc((void**)&d_r, (size_t)n * 24)); CUX(cudaMalloc((void**)&d_l, (size_t)n * 32));
    CUX(cudaMemcpy(d_xy, pixel_xy, (size_t)n * 8, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_si, sample_index, (size_t)n * 4, cudaMemcpyHostToDevice));
    RenderConst k{}; k.seed = rp->seed; k.option_flags = rp->option_flags; k.full_res_x = s->d.film.full_resolution[0];
    k_camera_rays<<<(unsigned)((n + 127) / 128), 128, 0, g_stream>>>(s->d, k, (long long)n, d_xy, d_si, d_r, d_l);
    CUX(cudaStreamSynchronize(g_stream));
    CUX(cudaMemcpy(out_rays, d_r, (size_t)n * 24, cudaMemcpyDeviceToHost));
    CUX(cudaMemcpy(out_lambda, d_l, (size_t)n * 32, cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return SG_OK;
}

int sg_texture_eval_ctx(SgScene* s, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* nrm, const float* lambda, float* out) {
    ENTER(0);
    if (!s || n < 0 || (n > 0 && (!q || !lambda || !out))) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (tex < 0 || (uint32_t)tex >= s->d.n_textures) return fail(SG_ERR_INVALID_ARGUMENT, "texture id out of range");
    if (n == 0) return SG_OK;
    float *d_q = nullptr, *d_l = nullptr, *d_o = nullptr, *d_p = nullptr, *d_n = nullptr;
    auto cleanup = [&]() { cudaFree(d_q); cudaFree(d_l); cudaFree(d_o); cudaFree(d_p); cudaFree(d_n); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CUX(cudaMalloc((void**)&d_q, (size_t)n * 24)); CUX(cudaMalloc((void**)&d_l, (size_t)n * 16)); CUX(cudaMalloc((void**)&d_o, (size_t)n * 16));
    CUX(cudaMemcpy(d_q, q, (size_t)n * 24, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_l, lambda, (size_t)n * 16, cudaMemcpyHostToDevice));
    if (pdp) { CUX(cudaMalloc((void**)&d_p, (size_t)n * 36)); CUX(cudaMemcpy(d_p, pdp, (size_t)n * 36, cudaMemcpyHostToDevice)); }
    if (nrm) { CUX(cudaMalloc((void**)&d_n, (size_t)n * 12)); CUX(cudaMemcpy(d_n, nrm, (size_t)n * 12, cudaMemcpyHostToDevice)); }
    k_texture_eval<<<(unsigned)((n + 127) / 128), 128, 0, g_stream>>>(s->d, tex, as_float, (long long)n, d_q, d_p, d_n, d_l, d_o);
    CUX(cudaStreamSynchronize(g_stream));
    CUX(cudaMemcpy(out, d_o, (size_t)n * 16, cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return SG_OK;
}
int sg_texture_eval_p(SgScene* s, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* lambda, float* out) {
    return sg_texture_eval_ctx(s, tex, as_float, n, q, pdp, nullptr, lambda, out);
}
int sg_texture_eval(SgScene* s, int tex, int as_float, int64_t n, const float* q, const float* lambda, float* out) {
    return sg_texture_eval_ctx(s, tex, as_float, n, q, nullptr, nullptr, lambda, out);
}

int sg_film_develop(SgScene* s, const SgFilmPixel* film, int64_t n, float* out_rgb) {
    ENTER(0);
    if (!s || n < 0 || (n > 0 && (!film || !out_rgb))) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return SG_OK;
    double* d_f = nullptr; float* d_o = nullptr;
    auto cleanup = [&]() { cudaFree(d_f); cudaFree(d_o); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CUX(cudaMalloc((void**)&d_f, (size_t)n * 32)); CUX(cudaMalloc((void**)&d_o, (size_t)n * 12));
    CUX(cudaMemcpy(d_f, film, (size_t)n * 32, cudaMemcpyHostToDevice));
    k_film_develop<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(s->d, d_f, (long long)n, d_o);
    CUX(cudaStreamSynchronize(g_stream));
    CUX(cudaMemcpy(out_rgb, d_o, (size_t)n * 12, cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return SG_OK;
}

int sg_film_get_image(SgScene* s, const SgFilmPixel* film, int32_t w, int32_t h, uint32_t flags, float* out_rgb) {
    ENTER(0);
    if (!s || w < 0 || h < 0 || (flags & ~3u)) return fail(SG_ERR_INVALID_ARGUMENT, "sg_film_get_image: bad size or flags");
    const size_t n = (size_t)w * (size_t)h;
    if (n == 0) return SG_OK;
    if (!film || !out_rgb) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    double* d_f = nullptr; float* d_o = nullptr;
    auto cleanup = [&]() { cudaFree(d_f); cudaFree(d_o); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    CUX(cudaMalloc((void**)&d_f, n * 32)); CUX(cudaMalloc((void**)&d_o, n * 12));
    CUX(cudaMemcpy(d_f, film, n * 32, cudaMemcpyHostToDevice));
    k_film_image<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(s->d, d_f, w, h, flags, d_o);
    CUX(cudaStreamSynchronize(g_stream));
    CUX(cudaMemcpy(out_rgb, d_o, n * 12, cudaMemcpyDeviceToHost));
#undef CUX
    cleanup();
    return SG_OK;
}

static uint32_t next_pow2_u32(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
static bool is_pow2_u32(uint32_t v) { return v && !(v & (v - 1)); }

int sg_image_pyramid_layout(int32_t width, int32_t height, int32_t nc, int32_t* n_levels, SgImageLevel* levels, uint64_t* n_texels) {
    if (width < 1 || height < 1 || nc < 1 || nc > 4 || !n_levels || !n_texels) return fail(SG_ERR_INVALID_ARGUMENT, "sg_image_pyramid_layout: bad arguments");
    uint32_t w = (uint32_t)width, h = (uint32_t)height;
    if (!is_pow2_u32(w) || !is_pow2_u32(h)) {
        const uint32_t nw = next_pow2_u32(w), nh = next_pow2_u32(h);
        if (!(nw > w && nh > h)) return fail(SG_ERR_UNSUPPORTED, "float_resize_up needs both dimensions to grow (image.rs:1009-1010 asserts)");
        w = nw; h = nh;
    }
    const int32_t n = 1 + (int32_t)std::log2((float)std::max(w, h));                                // image.rs:719
    if (n > 32) return fail(SG_ERR_UNSUPPORTED, "image too large");
    uint64_t off = 0; int32_t rx = (int32_t)w, ry = (int32_t)h;
    for (int32_t l = 0; l < n; ++l) {
        if (levels) { levels[l].offset = (uint32_t)off; levels[l].res[0] = rx; levels[l].res[1] = ry; levels[l].pad = 0; }
        off += (uint64_t)rx * ry * nc;
        rx = std::max(1, (rx + 1) / 2); ry = std::max(1, (ry + 1) / 2);
    }
    if (off >= (1ull << 32)) return fail(SG_ERR_UNSUPPORTED, "pyramid exceeds the 32-bit texel offsets of SgImageLevel");
    *n_levels = n; *n_texels = off;
    return SG_OK;
}

int sg_image_generate_pyramid(const float* image, int32_t width, int32_t height, int32_t nc, int32_t wrap, float* out_texels) {
    ENTER(0);
    if (!image || !out_texels) return fail(SG_ERR_INVALID_ARGUMENT, "null argument");
    if (wrap != SG_WRAP_REPEAT && wrap != SG_WRAP_CLAMP) return fail(SG_ERR_UNSUPPORTED, "pyramid construction supports the repeat and clamp wrap modes (image.rs:826-828 asserts on black)");
    int32_t n_levels = 0; uint64_t n_texels = 0; SgImageLevel levels[32];
    int rc = sg_image_pyramid_layout(width, height, nc, &n_levels, levels, &n_texels);
    if (rc != SG_OK) return rc;
    float *d_in = nullptr, *d_out = nullptr; ResampleWeight *d_xw = nullptr, *d_yw = nullptr;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_out); cudaFree(d_xw); cudaFree(d_yw); };
#define CUX(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(e__ == cudaErrorMemoryAllocation ? SG_ERR_OUT_OF_MEMORY : SG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
    const size_t n_in = (size_t)width * height * nc;
    CUX(cudaMalloc((void**)&d_in, n_in * 4)); CUX(cudaMalloc((void**)&d_out, n_texels * 4));
    CUX(cudaMemcpyAsync(d_in, image, n_in * 4, cudaMemcpyHostToDevice, g_stream));
    const int rx0 = levels[0].res[0], ry0 = levels[0].res[1];
    if (rx0 != width || ry0 != height) {                                                              // float_resize_up
        CUX(cudaMalloc((void**)&d_xw, (size_t)rx0 * sizeof(ResampleWeight))); CUX(cudaMalloc((void**)&d_yw, (size_t)ry0 * sizeof(ResampleWeight)));
        k_resample_weights<<<(rx0 + 127) / 128, 128, 0, g_stream>>>(width, rx0, d_xw);
        k_resample_weights<<<(ry0 + 127) / 128, 128, 0, g_stream>>>(height, ry0, d_yw);
        const long long n0 = (long long)rx0 * ry0 * nc;
        k_resize_up<<<(unsigned)((n0 + 255) / 256), 256, 0, g_stream>>>(d_in, width, height, nc, wrap, d_xw, d_yw, rx0, ry0, d_out);
    } else CUX(cudaMemcpyAsync(d_out, d_in, n_in * 4, cudaMemcpyDeviceToDevice, g_stream));
    for (int32_t l = 0; l + 1 < n_levels; ++l) {
        const long long nn = (long long)levels[l + 1].res[0] * levels[l + 1].res[1] * nc;
        k_downsample<<<(unsigned)((nn + 255) / 256), 256, 0, g_stream>>>(d_out + levels[l].offset, levels[l].res[0], levels[l].res[1], nc,
                                                                          levels[l + 1].res[0], levels[l + 1].res[1], d_out + levels[l + 1].offset);
    }
    CUX(cudaGetLastError());
    CUX(cudaMemcpyAsync(out_texels, d_out, n_texels * 4, cudaMemcpyDeviceToHost, g_stream));
    CUX(cudaStreamSynchronize(g_stream));
#undef CUX
    cleanup();
    return SG_OK;
}

}  // extern "C"
