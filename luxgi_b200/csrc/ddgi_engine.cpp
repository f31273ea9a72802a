// ddgi_engine.cpp — C++ host side of libluxddgi.so: the B200 replacement for the SDF-traced branch of the Maple
// renderer's DDGI pass, behind the C ABI declared in include/luxddgi.h.
//
// The host layer mirrors the reference's systems one to one (Code/Maple/src/Engine/DDGI/DDGIRenderer.cpp):
//   lux::ddgi::init::initializeProbeGrid   :163-211   atlas sizing, ray buffers, 2x2 ping-pong atlases
//   lux::ddgi::trace_rays::system          :215-331   (SDF branch :302-328)
//   lux::ddgi::probe_update::system        :345-414
//   lux::ddgi::border_update::system       :416-465
//   lux::ddgi::end_frame::system           :333-343
// There is no CPU fallback: creation fails with LUX_ERR_NO_DEVICE when no sm_100-class GPU is usable.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <new>
#include <string>

#include "ddgi_kernels.h"

namespace {

thread_local std::string g_lastError;

int fail(int code, const char* fmt, ...)
{
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define LUX_CUDA(expr)                                                                                          \
    do                                                                                                          \
    {                                                                                                           \
        cudaError_t e__ = (expr);                                                                               \
        if (e__ != cudaSuccess)                                                                                 \
            return fail(e__ == cudaErrorMemoryAllocation ? LUX_ERR_OUT_OF_MEMORY : LUX_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                           \
    } while (0)

struct DeviceBuffer
{
    void*  ptr      = nullptr;
    size_t bytes    = 0;
    bool   borrowed = false;

    void release()
    {
        if (ptr && !borrowed)
            cudaFree(ptr);
        ptr      = nullptr;
        bytes    = 0;
        borrowed = false;
    }
};

} // namespace

// The opaque context = the reference's per-entity components for this pass:
// DDGIUniform + DDGIPipelineInternal (DDGIRenderer.cpp:85-97) + the bound inputs of RaytracePass.sdfDescriptor.
struct LuxDDGIContext
{
    int          device    = 0;
    cudaStream_t stream    = nullptr;
    bool         ownStream = false;
    uint32_t     flags     = 0;
    int          rank = 0, world = 1;

    LuxDDGIUniform uniform{};
    int            totalProbes = 0, probeBegin = 0, probeCount = 0;
    int            layerProbes = 1, layerStride = 1; // LuxDDGIState: 1 = one z-slab, world = interleaved layers
    int            raysPadded  = 0;

    // DDGIPipelineInternal
    DeviceBuffer radiance, directionDepth;
    DeviceBuffer irradiance[2], depth[2];
    int32_t      frames      = 0;
    int32_t      pingPong    = 0;
    int32_t      lastWritten = 1; // index of the atlas pair most recently written ("current")
    bool         raysValid   = false;

    // per-frame tables
    DeviceBuffer dirs, wIrr, wDepth, scaleIrr, scaleDepth, nzIrr, nzDepth;
    DeviceBuffer origins, records, meta, chunkCounter; // wavefront trace scratch
    DeviceBuffer sortTicket, binCounts, binBlockSums, sortedIdx; // sorted shade scratch (the hit count lives in chunkCounter[1])
    DeviceBuffer dirsHalf;                                       // [R] fp16 directions for the blend weights (pipelined update)
    DeviceBuffer blendLists;                                     // live-ray lists of the list blend (lux::BlendLists layout)
    bool         useBlendLists = false;
    DeviceBuffer ummaW[2];                                       // tcgen05 irradiance GEMM: B' hi / lo [192][kPad] fp16
    DeviceBuffer tcW[4];                                         // LUX_DDGI_FLAG_BLEND_TC: irradiance hi / lo [64][kPad], depth hi / lo [256][kPad] fp16
    DeviceBuffer unitOrder, unitIndex, rayOrder, raySlot;        // march order tables (init::marchOrder)
    int          probeUnits = 0, rayClusters = 0;
    int          marchBeam  = -1;                                // chunk shape the tables were built for (-1 = none yet)
    DeviceBuffer probeTaps;                                      // [probeCount] float2: the first step's taps at each probe position (march step 0)
    bool         probeTapsDirty = true;                          // the volume or the probe positions changed

    cudaStream_t auxStream = nullptr;
    cudaEvent_t  evFork = nullptr, evJoin = nullptr, evWeights = nullptr;

    // uGlobalSDF / uGlobalMipSDF / sdfData
    bool             hasSdf = false;
    LuxGlobalSDFData sdfData{};
    DeviceBuffer     sdf, mip;
    cudaArray_t         sdfArray = nullptr, mipArray = nullptr;
    DeviceBuffer        mipScratch;        // one cascade's mip: flood ping-pong of lux_ddgi_update_global_sdf_region (kept between calls)
    cudaTextureObject_t sdfTex = 0, mipTex = 0;

    // surface cache
    bool                      hasAtlas = false;
    LuxGlobalSurfaceAtlasData atlasData{};
    DeviceBuffer              chunks, cull, objects, objectInverse, tiles, tileZRow, light, atlasDepth, chunkMasks;
    bool                      masksDirty = true;

    // sky
    int          skyFace = 0;
    DeviceBuffer sky;

    // timers
    cudaEvent_t   ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // copy engine overlap: host<->device transfers run on their own stream, ordered against the kernels by events
    cudaStream_t  copyStream = nullptr;   // host -> device (light cache)
    cudaStream_t  downStream = nullptr;   // device -> host (atlas rows): its own stream so both copy engines run at once
    // exchange step: in-place NCCL all-gather of the updated atlas rows, on its own stream
    void*         ncclComm = nullptr;
    cudaStream_t  gatherStream = nullptr;
    cudaEvent_t   evBlendDone = nullptr, evGather[2] = {nullptr, nullptr};
    bool          gatherPending[2] = {false, false};
    cudaEvent_t   fences[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t      fenceSeq  = 0;
    cudaEvent_t   evLightReady = nullptr, evShadeDone = nullptr, evIrrDone = nullptr, evDepthDone = nullptr, evCopyDone = nullptr;
    bool          lightPending = false;
    bool          timed = false;
    uint64_t      launches = 0;
};

// NCCL is bound at run time (dlopen): the library has no link-time dependency on it and single-GPU users never load it.
namespace {
struct NcclApi
{
    void* lib = nullptr;
    int (*allGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*groupStart)()                                                     = nullptr;
    int (*groupEnd)()                                                       = nullptr;
    const char* (*getErrorString)(int)                                      = nullptr;
};
NcclApi* ncclApi()
{
    static NcclApi api;
    static bool    tried = false;
    if (!tried)
    {
        tried = true;
        // an already loaded libnccl.so.2 (e.g. the one a host framework brought) is reused: same soname
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (lib)
        {
            api.allGather      = reinterpret_cast<decltype(api.allGather)>(dlsym(lib, "ncclAllGather"));
            api.groupStart     = reinterpret_cast<decltype(api.groupStart)>(dlsym(lib, "ncclGroupStart"));
            api.groupEnd       = reinterpret_cast<decltype(api.groupEnd)>(dlsym(lib, "ncclGroupEnd"));
            api.getErrorString = reinterpret_cast<decltype(api.getErrorString)>(dlsym(lib, "ncclGetErrorString"));
            if (api.allGather && api.groupStart && api.groupEnd && api.getErrorString)
                api.lib = lib;
        }
    }
    return api.lib ? &api : nullptr;
}
} // namespace

namespace lux {
namespace ddgi {

static int upload(LuxDDGIContext& c, DeviceBuffer& dst, const void* src, size_t bytes, LuxMemKind kind)
{
    if (kind == LUX_MEM_DEVICE)
    {
        dst.release();
        dst.ptr      = const_cast<void*>(src);
        dst.bytes    = bytes;
        dst.borrowed = true;
        return LUX_OK;
    }
    if (dst.borrowed || dst.bytes != bytes)
    {
        dst.release();
        LUX_CUDA(cudaMalloc(&dst.ptr, bytes ? bytes : 1));
        dst.bytes = bytes;
    }
    if (bytes)
        LUX_CUDA(cudaMemcpyAsync(dst.ptr, src, bytes, cudaMemcpyHostToDevice, c.stream));
    return LUX_OK;
}

static int allocZero(LuxDDGIContext& c, DeviceBuffer& dst, size_t bytes)
{
    dst.release();
    LUX_CUDA(cudaMalloc(&dst.ptr, bytes ? bytes : 1));
    dst.bytes = bytes;
    LUX_CUDA(cudaMemsetAsync(dst.ptr, 0, bytes, c.stream));
    return LUX_OK;
}

static void releaseSdfTextures(LuxDDGIContext& c)
{
    if (c.sdfTex) cudaDestroyTextureObject(c.sdfTex);
    if (c.mipTex) cudaDestroyTextureObject(c.mipTex);
    if (c.sdfArray) cudaFreeArray(c.sdfArray);
    if (c.mipArray) cudaFreeArray(c.mipArray);
    c.sdfTex = c.mipTex = 0;
    c.sdfArray = c.mipArray = nullptr;
}

// R16F volume [d][h][w] in linear device memory -> layered 2-D cudaArray (one layer per z) + texture object:
// unnormalised coordinates, clamp addressing, point sampling (only tld4 gathers are issued against it).
static int makeLayeredTexture(LuxDDGIContext& c, const void* dev, int w, int h, int d, cudaArray_t* arr, cudaTextureObject_t* tex)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDescHalf();
    LUX_CUDA(cudaMalloc3DArray(arr, &desc, make_cudaExtent((size_t)w, (size_t)h, (size_t)d), cudaArrayLayered));
    cudaMemcpy3DParms cp{};
    cp.srcPtr   = make_cudaPitchedPtr(const_cast<void*>(dev), (size_t)w * 2, (size_t)w, (size_t)h);
    cp.dstArray = *arr;
    cp.extent   = make_cudaExtent((size_t)w, (size_t)h, (size_t)d);
    cp.kind     = cudaMemcpyDeviceToDevice;
    LUX_CUDA(cudaMemcpy3DAsync(&cp, c.stream));
    cudaResourceDesc rd{};
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = *arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode       = cudaFilterModePoint;
    td.readMode         = cudaReadModeElementType;
    td.normalizedCoords = 0;
    LUX_CUDA(cudaCreateTextureObject(tex, &rd, &td, nullptr));
    return LUX_OK;
}

namespace init {

// Atlas sizing, DDGIRenderer.cpp:181-191
static void atlasSizes(LuxDDGIUniform& u)
{
    u.irradianceProbeSideLength = LUX_IRRADIANCE_OCT_SIZE;
    u.depthProbeSideLength      = LUX_DEPTH_OCT_SIZE;
    const int xy                = u.probeCounts[0] * u.probeCounts[1];
    u.irradianceTextureWidth    = (LUX_IRRADIANCE_OCT_SIZE + 2) * xy + 2;
    u.irradianceTextureHeight   = (LUX_IRRADIANCE_OCT_SIZE + 2) * u.probeCounts[2] + 2;
    u.depthTextureWidth         = (LUX_DEPTH_OCT_SIZE + 2) * xy + 2;
    u.depthTextureHeight        = (LUX_DEPTH_OCT_SIZE + 2) * u.probeCounts[2] + 2;
}

static int validate(const LuxDDGIUniform& u)
{
    if (u.probeCounts[0] <= 0 || u.probeCounts[1] <= 0 || u.probeCounts[2] <= 0)
        return fail(LUX_ERR_INVALID_ARG, "probeCounts must be positive (%d,%d,%d)", u.probeCounts[0], u.probeCounts[1], u.probeCounts[2]);
    if ((long long)u.probeCounts[0] * u.probeCounts[1] * u.probeCounts[2] > (1ll << 30))
        return fail(LUX_ERR_INVALID_ARG, "too many probes");
    if (u.raysPerProbe <= 0 || u.raysPerProbe > 65535)
        return fail(LUX_ERR_INVALID_ARG, "raysPerProbe %d out of range", u.raysPerProbe);
    if (u.irradianceProbeSideLength != LUX_IRRADIANCE_OCT_SIZE || u.depthProbeSideLength != LUX_DEPTH_OCT_SIZE)
        return fail(LUX_ERR_UNSUPPORTED, "probe side lengths must be %d / %d", LUX_IRRADIANCE_OCT_SIZE, LUX_DEPTH_OCT_SIZE);
    LuxDDGIUniform t = u;
    atlasSizes(t);
    if (t.irradianceTextureWidth != u.irradianceTextureWidth || t.irradianceTextureHeight != u.irradianceTextureHeight ||
        t.depthTextureWidth != u.depthTextureWidth || t.depthTextureHeight != u.depthTextureHeight)
        return fail(LUX_ERR_INVALID_ARG, "atlas sizes in DDGIUniform do not match probeCounts (call lux_ddgi_uniform_finalize)");
    if (!(u.ddgiGamma > 0.0f))
        return fail(LUX_ERR_INVALID_ARG, "ddgiGamma must be positive");
    return LUX_OK;
}

// Work order of the wavefront march (csrc/march_kernel.inc): which probes and which ray directions the k-th chunk holds.  Any order gives
// the same results (every ray is independent); this one makes the rays in flight a narrow cone leaving a compact block of probes:
//  * probe units (beam chunks: 2 consecutive probe ids; row chunks: 32, a run along x) are visited tile by tile: tiles of 32 x 16 x 16 probes
//    in Morton order of the tile coordinates, units inside a tile by (z, y, x);
//  * ray ids are clustered by direction: the un-rotated spherical-Fibonacci directions (DDGICommon.glsl:42-52) are sorted along a Hilbert curve
//    over their octahedral image, and every run of MARCH_CLUSTER_RAYS (32) slots of that order is one cluster.  The per-frame rotation is rigid,
//    so the clusters stay clusters.  (Consecutive Fibonacci indices are 137.5 degrees apart in azimuth: the id order itself is as incoherent
//    as it gets.)
static uint32_t hilbertIndex(uint32_t n, uint32_t x, uint32_t y)
{
    uint32_t d = 0;
    for (uint32_t s = n / 2; s > 0; s /= 2)
    {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (!ry)
        {
            if (rx)
            {
                x = n - 1 - x;
                y = n - 1 - y;
            }
            std::swap(x, y);
        }
    }
    return d;
}
static uint32_t spread3(uint32_t v) // 10 bits -> every third bit
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
static int marchOrder(LuxDDGIContext& c, bool beam)
{
    const LuxDDGIUniform& u = c.uniform;
    const int X = u.probeCounts[0], Y = u.probeCounts[1];
    const int unit = beam ? 2 : 32;
    const int PG = (c.probeCount + unit - 1) / unit; // probe units
    const int R = u.raysPerProbe, slots = (R + lux::MARCH_CLUSTER_RAYS - 1) / lux::MARCH_CLUSTER_RAYS * lux::MARCH_CLUSTER_RAYS;
    const bool identity = (c.flags & LUX_DDGI_FLAG_MARCH_PROBE_MAJOR) != 0; // ids as they come, probe units outermost
    std::vector<uint32_t> order(PG), index(PG);
    std::vector<std::pair<uint64_t, uint32_t>> keyed(PG);
    for (int g = 0; g < PG; g++)
    {
        const int id = g * unit, x = id % X, y = (id / X) % Y, z = id / (X * Y); // first probe of the unit, shard-local
        // interleaved shards (LuxDDGIState::layerStride = world): consecutive LOCAL layers are `world` layers apart, so a tile takes fewer of them to
        // stay as compact in the scene (measured at N = 2: 16 local layers per tile made the march 2.7 % slower than z-slabs)
        const int B  = c.layerStride > 1 ? c.layerProbes / (X * Y) : 16; // layers per interleaved block
        const int tz = c.layerStride > 1 ? (B >= 2 ? std::min(16, B) : std::max(1, 16 / c.layerStride)) : 16;
        const uint64_t tile = spread3((uint32_t)(x / 32)) | (spread3((uint32_t)(y / 16)) << 1) | ((uint64_t)spread3((uint32_t)(z / tz)) << 2);
        const uint64_t in = ((uint64_t)(z % tz) << 16) | ((uint64_t)(y % 16) << 8) | (uint64_t)(x % 32);
        keyed[g] = {identity ? (uint64_t)g : ((tile << 24) | in), (uint32_t)g};
    }
    std::sort(keyed.begin(), keyed.end());
    for (int i = 0; i < PG; i++)
    {
        order[i]               = keyed[i].second;
        index[keyed[i].second] = (uint32_t)i;
    }
    std::vector<uint16_t> rayOrder(slots), raySlot(slots);
    std::vector<std::pair<uint32_t, uint16_t>> rk(R);
    for (int i = 0; i < R; i++)
    {
        const double ab = i * 0.6180339887498949, phi = 6.283185307179586 * (ab - std::floor(ab));
        const double ct = 1.0 - (2.0 * i + 1.0) / R, st = std::sqrt(std::max(0.0, 1.0 - ct * ct));
        double x = std::cos(phi) * st, y = std::sin(phi) * st, z = ct;
        const double n = std::fabs(x) + std::fabs(y) + std::fabs(z);
        x /= n; y /= n;
        if (z < 0.0)
        { // fold the lower hemisphere outwards (octahedral map)
            const double fx = (1.0 - std::fabs(y)) * (x >= 0.0 ? 1.0 : -1.0), fy = (1.0 - std::fabs(x)) * (y >= 0.0 ? 1.0 : -1.0);
            x = fx; y = fy;
        }
        const uint32_t N = 1024;
        const uint32_t qx = (uint32_t)std::min<double>(N - 1, std::max(0.0, (x * 0.5 + 0.5) * N)), qy = (uint32_t)std::min<double>(N - 1, std::max(0.0, (y * 0.5 + 0.5) * N));
        rk[i] = {identity ? (uint32_t)i : hilbertIndex(N, qx, qy), (uint16_t)i};
    }
    std::sort(rk.begin(), rk.end());
    for (int s = 0; s < slots; s++)
        rayOrder[s] = (uint16_t)(s < R ? rk[s].second : s); // padding slots keep ids >= R: never traced
    for (int s = 0; s < slots; s++)
        raySlot[rayOrder[s]] = (uint16_t)s;
    int rc;
    if ((rc = allocZero(c, c.unitOrder, (size_t)PG * sizeof(uint32_t))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.unitIndex, (size_t)PG * sizeof(uint32_t))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.rayOrder, (size_t)slots * sizeof(uint16_t))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.raySlot, (size_t)slots * sizeof(uint16_t))) != LUX_OK) return rc;
    LUX_CUDA(cudaMemcpyAsync(c.unitOrder.ptr, order.data(), (size_t)PG * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    LUX_CUDA(cudaMemcpyAsync(c.unitIndex.ptr, index.data(), (size_t)PG * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream));
    LUX_CUDA(cudaMemcpyAsync(c.rayOrder.ptr, rayOrder.data(), (size_t)slots * sizeof(uint16_t), cudaMemcpyHostToDevice, c.stream));
    LUX_CUDA(cudaMemcpyAsync(c.raySlot.ptr, raySlot.data(), (size_t)slots * sizeof(uint16_t), cudaMemcpyHostToDevice, c.stream));
    LUX_CUDA(cudaStreamSynchronize(c.stream)); // the host vectors go out of scope
    c.probeUnits = PG;
    c.marchBeam  = beam ? 1 : 0;
    if (c.sortTicket.ptr) // padding records (lanes / slots beyond the volume) are never marched, so nothing ever writes their ticket: "not a hit", and
                          // which records are padding depends on the chunk shape
        LUX_CUDA(cudaMemsetAsync(c.sortTicket.ptr, 0xff, c.sortTicket.bytes, c.stream));
    c.rayClusters = slots / lux::MARCH_CLUSTER_RAYS;
    return LUX_OK;
}

// init::initializeProbeGrid, DDGIRenderer.cpp:163-211.  The reference caps probes at int16 max (:169); this engine does not.
static int initializeProbeGrid(LuxDDGIContext& c)
{
    const LuxDDGIUniform& u = c.uniform;
    const size_t rayBytes   = (size_t)c.probeCount * u.raysPerProbe * 8; // RGBA16F
    int rc;
    if ((rc = allocZero(c, c.radiance, rayBytes)) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.directionDepth, rayBytes)) != LUX_OK) return rc;
    const size_t irrBytes   = (size_t)u.irradianceTextureWidth * u.irradianceTextureHeight * 8; // RGBA16F
    const size_t depthBytes = (size_t)u.depthTextureWidth * u.depthTextureHeight * 4;           // RG16F
    for (int i = 0; i < 2; i++)
    {
        if ((rc = allocZero(c, c.irradiance[i], irrBytes)) != LUX_OK) return rc;
        if ((rc = allocZero(c, c.depth[i], depthBytes)) != LUX_OK) return rc;
    }
    c.raysPadded = (u.raysPerProbe + 31) / 32 * 32;
    if ((rc = allocZero(c, c.dirs, (size_t)u.raysPerProbe * sizeof(float4))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.wIrr, (size_t)c.raysPadded * 64 * sizeof(float))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.wDepth, (size_t)c.raysPadded * 256 * sizeof(float))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.scaleIrr, 64 * sizeof(float))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.scaleDepth, 256 * sizeof(float))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.nzIrr, (size_t)c.raysPadded * sizeof(uint32_t))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.nzDepth, (size_t)c.raysPadded * sizeof(uint32_t))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.origins, (size_t)c.probeCount * sizeof(float4))) != LUX_OK) return rc;
    if ((rc = allocZero(c, c.chunkCounter, 64 + lux::trace_shade_pool_bytes())) != LUX_OK) return rc; // 16 counters + the shade's per-SM pools
    if (!(c.flags & LUX_DDGI_FLAG_TRACE_SIMPLE))
    {
        const size_t nrec = lux::trace_record_capacity(c.probeCount, u.raysPerProbe);
        if ((rc = allocZero(c, c.records, nrec * sizeof(float4))) != LUX_OK) return rc;
        if ((rc = allocZero(c, c.meta, nrec * sizeof(uint32_t))) != LUX_OK) return rc;
        if (!(c.flags & LUX_DDGI_FLAG_SHADE_UNSORTED) && nrec < (size_t(1) << 30) /* record index + 2 cascade bits in one u32 */)
        {
            if ((rc = allocZero(c, c.sortTicket, nrec * sizeof(uint2))) != LUX_OK) return rc;
            if ((rc = allocZero(c, c.sortedIdx, nrec * sizeof(uint32_t))) != LUX_OK) return rc;
            if ((rc = allocZero(c, c.binCounts, lux::trace_sort_bins() * sizeof(uint32_t))) != LUX_OK) return rc;
            if ((rc = allocZero(c, c.binBlockSums, lux::trace_sort_blocks() * sizeof(uint32_t))) != LUX_OK) return rc;
        }
        c.marchBeam = -1; // tables are built by the first trace, when the bound volume decides the chunk shape (trace_rays::setup)
    }
    if ((rc = allocZero(c, c.dirsHalf, (size_t)u.raysPerProbe * sizeof(uint2))) != LUX_OK) return rc;
    c.useBlendLists = !(c.flags & (LUX_DDGI_FLAG_BLEND_TC | LUX_DDGI_FLAG_BLEND_TILES)) &&
                      ((c.flags & LUX_DDGI_FLAG_BLEND_LISTS) || lux::blend_lists_preferred(c.probeCount));
    if (c.useBlendLists && (rc = allocZero(c, c.blendLists, lux::blend_lists_bytes(u.raysPerProbe, c.raysPadded))) != LUX_OK) return rc;
    if (c.flags & LUX_DDGI_FLAG_BLEND_TC)
    {
        const size_t kPad = (size_t)lux::blend_tc_kpad(u.raysPerProbe);
        for (int i = 0; i < 4; i++)
            if ((rc = allocZero(c, c.tcW[i], (i < 2 ? 64 : 256) * kPad * sizeof(uint16_t))) != LUX_OK) return rc;
        if (!(c.flags & LUX_DDGI_FLAG_BLEND_TC_MMA_SYNC))
            for (int i = 0; i < 2; i++)
                if ((rc = allocZero(c, c.ummaW[i], (size_t)192 * lux::blend_umma_irr_kpad(u.raysPerProbe) * sizeof(uint16_t))) != LUX_OK) return rc;
    }
    c.frames      = 0;
    c.pingPong    = 0;
    c.lastWritten = 1;
    c.raysValid   = false;
    return LUX_OK;
}

} // namespace init

static void fillVolume(const LuxDDGIContext& c, lux::TraceParams& p)
{
    const LuxDDGIUniform& u = c.uniform;
    for (int i = 0; i < 3; i++)
    {
        p.start[i] = u.startPosition[i];
        p.step[i]  = u.step[i];
    }
    p.countX       = u.probeCounts[0];
    p.countY       = u.probeCounts[1];
    p.raysPerProbe = u.raysPerProbe;
    p.probeBegin   = c.probeBegin;
    p.probeCount   = c.probeCount;
    p.layerProbes  = c.layerProbes;
    p.layerStride  = c.layerStride;
    p.origins      = (const float4*)c.origins.ptr;
}

// probe positions only change with the uniform: computed once per (re)configuration
static int updateOrigins(LuxDDGIContext& c)
{
    lux::TraceParams p{};
    fillVolume(c, p);
    lux::launch_probe_origins(p, c.stream);
    c.launches += 1;
    c.probeTapsDirty = true;
    LUX_CUDA(cudaGetLastError());
    return LUX_OK;
}

static void mark(LuxDDGIContext& c, int i)
{
    if (c.flags & LUX_DDGI_FLAG_STAGE_TIMERS)
        cudaEventRecord(c.ev[i], c.stream);
}

namespace trace_rays {

// Per-frame setup of the trace: surface-cache prefilter (when the lists changed) and the frame's ray directions.
static int setup(LuxDDGIContext& c, const LuxTracePushConstants& push)
{
    if (!c.hasSdf)
        return fail(LUX_ERR_NOT_READY, "trace_rays: no global SDF bound (lux_ddgi_set_global_sdf)");
    const LuxDDGIUniform& u = c.uniform;

    if (c.hasAtlas && c.masksDirty && !(c.flags & LUX_DDGI_FLAG_NO_PREFILTER))
    { // surface-cache prefilter: depends on the object lists and on the largest surfaceThreshold (1.05 * voxel)
        float vmax = 0.0f;
        for (uint32_t i = 0; i < c.sdfData.cascadesCount; i++)
            vmax = std::fmax(vmax, c.sdfData.cascadeVoxelSize[i]);
        const size_t bytes = (size_t)64000 * 64 * sizeof(unsigned long long);
        if (c.chunkMasks.bytes != bytes)
        {
            c.chunkMasks.release();
            LUX_CUDA(cudaMalloc(&c.chunkMasks.ptr, bytes));
            c.chunkMasks.bytes = bytes;
        }
        launch_chunk_masks((const uint32_t*)c.chunks.ptr, (const uint32_t*)c.cull.ptr, (const LuxObjectBuffer*)c.objects.ptr,
                           (const float*)c.objectInverse.ptr, c.atlasData.objectsCount, c.atlasData.chunkSize, 1.05f * vmax * 1.001f,
                           (unsigned long long*)c.chunkMasks.ptr, c.stream);
        c.launches += 1;
        LUX_CUDA(cudaGetLastError());
        c.masksDirty = false;
    }
    if (!(c.flags & LUX_DDGI_FLAG_TRACE_SIMPLE))
    { // Chunk shape of the march.  Beams (a warp = one probe's rays into a narrow cone) need far fewer sector requests per gather, but their
      // lanes spread over many z-slices, and a slice of a large volume is a 2 MiB page of its own: measured on B200, beams win while the
      // full-resolution volume is within the TLB's reach (C4, 256 MiB: 3.28 vs 3.99 ms) and lose beyond it (C5, 2 GiB: 118 vs 91 ms).
        const size_t sdfBytes = (size_t)c.sdfData.resolution * c.sdfData.resolution * c.sdfData.resolution * c.sdfData.cascadesCount * 2;
        bool beam = sdfBytes <= (size_t(256) << 20);
        if (c.flags & LUX_DDGI_FLAG_MARCH_ROWS) beam = false;
        if (c.flags & LUX_DDGI_FLAG_MARCH_BEAMS) beam = true;
        if (c.marchBeam != (beam ? 1 : 0))
        {
            int rc = init::marchOrder(c, beam);
            if (rc != LUX_OK)
                return rc;
        }
    }
    mark(c, 0);
    launch_ray_dirs(push.randomOrientation, u.raysPerProbe, (float4*)c.dirs.ptr, (uint2*)c.dirsHalf.ptr, c.stream);
    c.launches += 1;
    mark(c, 1);
    return LUX_OK;
}

// The trace kernels of the shard on stream `s`.
static int launch(LuxDDGIContext& c, cudaStream_t s, bool timers)
{
    TraceParams p{};
    fillVolume(c, p);
    p.sdf          = c.sdfData;
    p.tex          = (const uint16_t*)c.sdf.ptr;
    p.mip          = (const uint16_t*)c.mip.ptr;
    p.res          = (int)c.sdfData.resolution;
    p.mipRes       = p.res / 4;
    p.cascades     = (int)c.sdfData.cascadesCount;
    p.texObj       = c.sdfTex;
    p.mipObj       = c.mipTex;
    p.hasAtlas     = c.hasAtlas ? 1 : 0;
    if (c.hasAtlas)
    {
        p.chunkSize     = c.atlasData.chunkSize;
        p.atlasRes      = c.atlasData.resolution;
        p.objectsCount  = c.atlasData.objectsCount;
        p.chunks        = (const uint32_t*)c.chunks.ptr;
        p.cull          = (const uint32_t*)c.cull.ptr;
        p.objects       = (const LuxObjectBuffer*)c.objects.ptr;
        p.objectInverse = (const float*)c.objectInverse.ptr;
        p.chunkMasks    = (c.flags & LUX_DDGI_FLAG_NO_PREFILTER) ? nullptr : (const unsigned long long*)c.chunkMasks.ptr;
        p.tiles         = (const LuxTileBuffer*)c.tiles.ptr;
        p.tileZRow      = (c.flags & LUX_DDGI_FLAG_NO_PREFILTER) ? nullptr : (const float4*)c.tileZRow.ptr;
        p.light         = (const uint2*)c.light.ptr;
        p.depth         = (const float*)c.atlasDepth.ptr;
    }
    p.skyFace  = c.skyFace;
    p.sky      = (const uint2*)c.sky.ptr;
    p.dirs     = (const float4*)c.dirs.ptr;
    p.radiance = (uint2*)c.radiance.ptr;
    p.dirDist  = (uint2*)c.directionDepth.ptr;
    p.steps    = nullptr;
    p.records  = (float4*)c.records.ptr;
    p.meta     = (uint32_t*)c.meta.ptr;
    p.unitOrder = (const uint32_t*)c.unitOrder.ptr;
    p.unitIndex = (const uint32_t*)c.unitIndex.ptr;
    p.beam      = c.marchBeam;
    p.rayOrder = (const uint16_t*)c.rayOrder.ptr;
    p.raySlot  = (const uint16_t*)c.raySlot.ptr;
    p.probeUnits  = c.probeUnits;
    p.rayClusters = c.rayClusters;
    p.probeMajor  = (c.flags & LUX_DDGI_FLAG_MARCH_PROBE_MAJOR) ? 1 : 0;
    if (!(c.flags & LUX_DDGI_FLAG_TRACE_SIMPLE))
    { // per-probe cache of the first step's taps (march_kernel.inc): rebuilt when the volume or the probe positions changed
        if (c.probeTaps.bytes != (size_t)c.probeCount * sizeof(float2))
        {
            c.probeTaps.release();
            LUX_CUDA(cudaMalloc(&c.probeTaps.ptr, (size_t)c.probeCount * sizeof(float2)));
            c.probeTaps.bytes = (size_t)c.probeCount * sizeof(float2);
            c.probeTapsDirty  = true;
        }
        if (c.probeTapsDirty)
        {
            lux::launch_probe_taps(p, c.sdfTex != 0, (float2*)c.probeTaps.ptr, s);
            c.launches += 1;
            c.probeTapsDirty = false;
        }
        p.probeTaps = (const float2*)c.probeTaps.ptr;
    }
    unsigned int* counters = (unsigned int*)c.chunkCounter.ptr; // [0] march chunk counter, [1] hit count, [2] non-finite ray value seen
    p.nonFinite = counters + 2;
    cudaMemsetAsync(counters + 2, 0, sizeof(unsigned int), s);
    if (c.sortedIdx.ptr)
    {
        p.invChunkSize = c.hasAtlas ? 1.0f / c.atlasData.chunkSize : 0.0f;
        p.sortTicket   = (uint2*)c.sortTicket.ptr;
        p.binCounts    = (uint32_t*)c.binCounts.ptr;
        p.binBlockSums = (uint32_t*)c.binBlockSums.ptr;
        p.hitCount     = counters + 1;
        p.shadeCounter = counters + 3;
        p.shadePools   = reinterpret_cast<unsigned long long*>(counters + 16);
        p.sortedIdx    = (uint32_t*)c.sortedIdx.ptr;
    }
    const int variant = (c.flags & LUX_DDGI_FLAG_TRACE_SIMPLE) ? 0 : (c.sdfTex ? 2 : 1);
    c.launches += launch_trace(p, variant, counters, s, c.lightPending ? c.evLightReady : nullptr,
                               (timers && (c.flags & LUX_DDGI_FLAG_STAGE_TIMERS)) ? c.ev[5] : nullptr);
    return LUX_OK;
}

// trace_rays::system, SDF branch (DDGIRenderer.cpp:235-240, 302-328): the whole shard as one batch on the context's stream
static int system(LuxDDGIContext& c, const LuxTracePushConstants& push)
{
    int rc = setup(c, push);
    if (rc != LUX_OK)
        return rc;
    launch(c, c.stream, true);
    c.lightPending = false;
    cudaEventRecord(c.evShadeDone, c.stream);
    mark(c, 2);
    LUX_CUDA(cudaGetLastError());
    c.raysValid = true;
    return LUX_OK;
}

} // namespace trace_rays

namespace probe_update {

// The frame's probe-independent blend weights from a row of fp16 ray directions
static void weights(LuxDDGIContext& c, const uint2* dirsHalf, cudaStream_t s)
{
    const LuxDDGIUniform& u = c.uniform;
    c.launches += launch_blend_weights(dirsHalf, u.raysPerProbe, c.raysPadded, u.sharpness, (float*)c.wIrr.ptr, (float*)c.wDepth.ptr,
                                       (float*)c.scaleIrr.ptr, (float*)c.scaleDepth.ptr, (uint32_t*)c.nzIrr.ptr, (uint32_t*)c.nzDepth.ptr, s);
    if (c.useBlendLists)
        c.launches += lux::launch_blend_lists((const float*)c.wIrr.ptr, (const float*)c.wDepth.ptr, u.raysPerProbe,
                                              lux::blend_lists_layout(c.blendLists.ptr, u.raysPerProbe, c.raysPadded), s);
    if (c.flags & LUX_DDGI_FLAG_BLEND_TC)
    {
        lux::launch_blend_tc_weights((const float*)c.wIrr.ptr, (const float*)c.wDepth.ptr, c.raysPadded, lux::blend_tc_kpad(u.raysPerProbe), (uint16_t*)c.tcW[0].ptr,
                                     (uint16_t*)c.tcW[1].ptr, (uint16_t*)c.tcW[2].ptr, (uint16_t*)c.tcW[3].ptr, s);
        c.launches += 2;
        if (c.ummaW[0].ptr)
        {
            lux::launch_blend_umma_irr_weights((const float*)c.wIrr.ptr, u.raysPerProbe, (uint16_t*)c.ummaW[0].ptr, (uint16_t*)c.ummaW[1].ptr, s);
            c.launches += 1;
        }
    }
}

// Blend (+ fused border) of the shard on stream `s`; records evIrr / evDepth after the respective kernel when given.
static void launch(LuxDDGIContext& c, cudaStream_t s, cudaEvent_t evIrr, cudaEvent_t evDepth)
{
    const LuxDDGIUniform& u = c.uniform;
    const int    writeIdx = 1 - c.pingPong;
    BlendParams p{};
    p.probeBegin   = c.probeBegin;
    p.probeCount   = c.probeCount;
    p.layerProbes  = c.layerProbes;
    p.layerStride  = c.layerStride;
    p.raysPerProbe = u.raysPerProbe;
    p.raysPadded   = c.raysPadded;
    p.probesPerRow = u.probeCounts[0] * u.probeCounts[1];
    p.irrWidth     = u.irradianceTextureWidth;
    p.depthWidth   = u.depthTextureWidth;
    p.hysteresis   = u.hysteresis;
    p.invGamma     = 1.0f / u.ddgiGamma; // ProbeUpdate.glsl:141
    p.maxDistance  = u.maxDistance;
    p.firstFrame   = (c.frames == 0) ? 1 : 0; // DDGIRenderer.cpp:362
    p.fuseBorder   = (c.flags & LUX_DDGI_FLAG_UNFUSED_BORDER) ? 0 : 1;
    p.radiance     = (const uint2*)c.radiance.ptr;
    p.dirDist      = (const uint2*)c.directionDepth.ptr;
    p.wIrr         = (const float*)c.wIrr.ptr;
    p.wDepth       = (const float*)c.wDepth.ptr;
    p.scaleIrr     = (const float*)c.scaleIrr.ptr;
    p.scaleDepth   = (const float*)c.scaleDepth.ptr;
    p.nzIrr        = (const uint32_t*)c.nzIrr.ptr;
    p.nzDepth      = (const uint32_t*)c.nzDepth.ptr;
    p.nonFinite    = (const uint32_t*)c.chunkCounter.ptr + 2;
    p.prevIrr      = (const uint2*)c.irradiance[c.pingPong].ptr;
    p.outIrr       = (uint2*)c.irradiance[writeIdx].ptr;
    p.prevDepth    = (const uint32_t*)c.depth[c.pingPong].ptr;
    p.outDepth     = (uint32_t*)c.depth[writeIdx].ptr;
    const bool tc   = (c.flags & LUX_DDGI_FLAG_BLEND_TC) != 0;
    const int  kPad = lux::blend_tc_kpad(u.raysPerProbe);
    const lux::BlendLists lists = c.useBlendLists ? lux::blend_lists_layout(c.blendLists.ptr, u.raysPerProbe, c.raysPadded) : lux::BlendLists{};
    if (tc)
    {
        if (!c.ummaW[0].ptr || !lux::launch_blend_irradiance_umma(p, (const uint16_t*)c.ummaW[0].ptr, (const uint16_t*)c.ummaW[1].ptr, s))
            lux::launch_blend_irradiance_tc(p, (const uint16_t*)c.tcW[0].ptr, (const uint16_t*)c.tcW[1].ptr, kPad, s);
    }
    else if (c.useBlendLists)
        lux::launch_blend_irradiance_lists(p, lists, s);
    else
        launch_blend_irradiance(p, s);
    if (evIrr)
        cudaEventRecord(evIrr, s);
    if (c.useBlendLists)
        lux::launch_blend_depth_lists(p, lists, s);
    else if (tc && c.ummaW[0].ptr && lux::launch_blend_depth_umma(p, (const uint16_t*)c.tcW[2].ptr, (const uint16_t*)c.tcW[3].ptr, kPad, s))
        ;
    else if (!tc || !lux::launch_blend_depth_tc(p, (const uint16_t*)c.tcW[2].ptr, (const uint16_t*)c.tcW[3].ptr, kPad, s))
        launch_blend_depth(p, s);
    if (evDepth)
        cudaEventRecord(evDepth, s);
    c.launches += 2;
}

// probe_update::system (DDGIRenderer.cpp:377-413): writeIdx = 1 - pingPong, firstFrame = (frames == 0)
static int system(LuxDDGIContext& c)
{
    if (!c.raysValid)
        return fail(LUX_ERR_NOT_READY, "probe_update: ray buffers are empty (call lux_ddgi_trace_rays or lux_ddgi_set_ray_buffers)");
    weights(c, (const uint2*)c.directionDepth.ptr, c.stream); // the directions as stored in the ray buffer (row of the first probe)
    cudaStreamWaitEvent(c.stream, c.evCopyDone, 0); // row downloads of earlier frames must have left the atlases
    if (c.gatherPending[1 - c.pingPong])               // ... and so must the all-gather that last read / wrote this pair
        cudaStreamWaitEvent(c.stream, c.evGather[1 - c.pingPong], 0);
    launch(c, c.stream, c.evIrrDone, c.evDepthDone);
    mark(c, 3);
    LUX_CUDA(cudaGetLastError());
    c.lastWritten = 1 - c.pingPong;
    return LUX_OK;
}

} // namespace probe_update

namespace border_update {

// border_update::system (DDGIRenderer.cpp:441-464): in place on tex[1 - pingPong]
static int system(LuxDDGIContext& c)
{
    const LuxDDGIUniform& u = c.uniform;
    const int writeIdx = 1 - c.pingPong;
    launch_border((uint2*)c.irradiance[writeIdx].ptr, u.irradianceTextureWidth, (uint32_t*)c.depth[writeIdx].ptr, u.depthTextureWidth,
                  u.probeCounts[0] * u.probeCounts[1], c.probeBegin, c.probeCount, c.layerProbes, c.layerStride, c.stream);
    c.launches += 2;
    LUX_CUDA(cudaGetLastError());
    return LUX_OK;
}

} // namespace border_update

namespace end_frame {

// end_frame::system (DDGIRenderer.cpp:333-343)
static int system(LuxDDGIContext& c)
{
    c.pingPong = 1 - c.pingPong;
    c.frames++;
    return LUX_OK;
}

} // namespace end_frame

} // namespace ddgi
} // namespace lux

using namespace lux::ddgi;

// Shard of `rank`: one z-slab, or (LUX_DDGI_FLAG_SHARD_INTERLEAVED) the z-layers rank, rank + world, ...  See LuxDDGIState.
static void shardLayout(const LuxDDGIUniform& u, int rank, int world, uint32_t flags, LuxDDGIState* out)
{
    const int  xy     = u.probeCounts[0] * u.probeCounts[1];
    const int  zCount = u.probeCounts[2] / world;
    const bool inter  = (flags & LUX_DDGI_FLAG_SHARD_INTERLEAVED) != 0 && world > 1;
    const int  B      = inter ? LUX_DDGI_SHARD_BLOCK_LAYERS(flags) : 1; // layers per interleave unit (validated by the callers: world * B divides Z)
    const int  zBegin = inter ? rank * B : zCount * rank;
    out->probeBegin         = zBegin * xy;
    out->probeCount         = zCount * xy;
    out->irradianceRowBegin = 1 + zBegin * (LUX_IRRADIANCE_OCT_SIZE + 2);
    out->irradianceRowCount = zCount * (LUX_IRRADIANCE_OCT_SIZE + 2);
    out->depthRowBegin      = 1 + zBegin * (LUX_DEPTH_OCT_SIZE + 2);
    out->depthRowCount      = zCount * (LUX_DEPTH_OCT_SIZE + 2);
    out->layerProbes        = B * xy;
    out->layerStride        = inter ? world : 1;
    out->unitLayers         = B;
}
static bool shardFlagsValid(const LuxDDGIUniform& u, int world, uint32_t flags)
{
    if (!(flags & LUX_DDGI_FLAG_SHARD_INTERLEAVED) || world <= 1)
        return true;
    return u.probeCounts[2] % (world * LUX_DDGI_SHARD_BLOCK_LAYERS(flags)) == 0;
}
// does this shard own every row of [rowBegin, rowBegin + rowCount) of an atlas with `side` + 2 rows per z-layer?
static bool ownsRows(const LuxDDGIContext* c, int side, int rowBegin, int rowCount)
{
    LuxDDGIState st{};
    shardLayout(c->uniform, c->rank, c->world, c->flags, &st);
    const int S = side + 2;
    if (rowCount <= 0)
        return true;
    if (rowBegin < 1)
        return false; // the outer pad row belongs to nobody
    const int xy = c->uniform.probeCounts[0] * c->uniform.probeCounts[1];
    const int first = (rowBegin - 1) / S, last = (rowBegin + rowCount - 2) / S, zBegin = st.probeBegin / xy, zCount = st.probeCount / xy;
    for (int z = first; z <= last; z++)
    {
        const int k = z - zBegin; // layers past the shard's first one
        if (k < 0)
            return false;
        const int unit = k / st.unitLayers; // in units of B layers: own units are every layerStride-th
        if (unit % st.layerStride != 0 || (unit / st.layerStride) * st.unitLayers + k % st.unitLayers >= zCount)
            return false;
    }
    return true;
}


#define CHECK_CTX(ctx)                                                  \
    do                                                                  \
    {                                                                   \
        if (!(ctx))                                                     \
            return fail(LUX_ERR_INVALID_ARG, "null context");           \
        LUX_CUDA(cudaSetDevice((ctx)->device));                         \
    } while (0)

// Ordering of the surface light cache between the copy stream (host uploads) and its users on the context's stream.
//   lightAcquire: before any kernel / copy on c->stream that reads or writes c->light - waits for a pending upload.
//   lightRelease: after it - the next upload on the copy stream waits for this point (evShadeDone = "last use of the light cache on c->stream").
static void lightAcquire(LuxDDGIContext* c)
{
    if (c->lightPending)
    {
        cudaStreamWaitEvent(c->stream, c->evLightReady, 0);
        c->lightPending = false;
    }
}
static void lightRelease(LuxDDGIContext* c) { cudaEventRecord(c->evShadeDone, c->stream); }

extern "C" {

uint32_t lux_ddgi_version(void) { return LUXDDGI_VERSION; }

const char* lux_ddgi_last_error(void) { return g_lastError.c_str(); }

int lux_ddgi_uniform_finalize(LuxDDGIUniform* u)
{
    if (!u)
        return fail(LUX_ERR_INVALID_ARG, "null uniform");
    init::atlasSizes(*u);
    return LUX_OK;
}

int lux_ddgi_uniform_from_volume(const LuxIrradianceVolume* v, const float aabbMin[3], const float aabbMax[3], LuxDDGIUniform* out)
{
    if (!v || !aabbMin || !aabbMax || !out)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    if (!(v->probeDistance > 0.0f))
        return fail(LUX_ERR_INVALID_ARG, "probeDistance must be positive");
    LuxDDGIUniform u{};
    for (int i = 0; i < 3; i++)
    {
        float sceneLength  = aabbMax[i] - aabbMin[i];
        u.probeCounts[i]   = (int)(sceneLength / v->probeDistance) + 2; // DDGIRenderer.cpp:669 "Add 2 more probes to fully cover scene"
        u.startPosition[i] = aabbMin[i];
        u.step[i]          = v->probeDistance;
    }
    u.probeCounts[3]   = 1;
    u.startPosition[3] = 1.0f;
    u.step[3]          = v->probeDistance;
    u.maxDistance      = v->probeDistance * 1.5f; // :674
    u.sharpness        = v->depthSharpness;
    u.hysteresis       = v->hysteresis;
    u.normalBias       = v->normalBias;
    u.ddgiGamma        = v->ddgiGamma;
    u.raysPerProbe     = v->raysPerProbe; // the reference forgets this copy (SURVEY finding 8); the engine does not
    init::atlasSizes(u);
    *out = u;
    return LUX_OK;
}

int lux_ddgi_create(const LuxDDGIUniform* uniform, const LuxDDGICreateInfo* info, LuxDDGIContext** out)
{
    if (!uniform || !out)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    int rc = init::validate(*uniform);
    if (rc != LUX_OK)
        return rc;
    LuxDDGICreateInfo ci{};
    if (info)
        ci = *info;
    if (ci.world <= 0)
        ci.world = 1;
    if (ci.rank < 0 || ci.rank >= ci.world)
        return fail(LUX_ERR_INVALID_ARG, "rank %d outside world %d", ci.rank, ci.world);
    if (uniform->probeCounts[2] % ci.world != 0)
        return fail(LUX_ERR_INVALID_ARG, "world %d must divide probeCounts.z %d (z-slab sharding)", ci.world, uniform->probeCounts[2]);
    if (!shardFlagsValid(*uniform, ci.world, ci.flags))
        return fail(LUX_ERR_INVALID_ARG, "world %d x %d layers per interleaved block must divide probeCounts.z %d", ci.world, LUX_DDGI_SHARD_BLOCK_LAYERS(ci.flags),
                    uniform->probeCounts[2]);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(LUX_ERR_NO_DEVICE, "no CUDA device available (%s); the engine has no CPU fallback", cudaGetErrorString(e));
    if (ci.device < 0 || ci.device >= ndev)
        return fail(LUX_ERR_INVALID_ARG, "device %d out of range (%d devices)", ci.device, ndev);
    cudaDeviceProp prop{};
    LUX_CUDA(cudaGetDeviceProperties(&prop, ci.device));
    if (prop.major != 10)
        return fail(LUX_ERR_NO_DEVICE, "device %d is sm_%d%d; this build contains sm_100a code only", ci.device, prop.major, prop.minor);
    LUX_CUDA(cudaSetDevice(ci.device));

    LuxDDGIContext* c = new (std::nothrow) LuxDDGIContext();
    if (!c)
        return fail(LUX_ERR_OUT_OF_MEMORY, "host allocation failed");
    c->device  = ci.device;
    c->flags   = ci.flags;
    c->rank    = ci.rank;
    c->world   = ci.world;
    c->uniform = *uniform;
    c->totalProbes = uniform->probeCounts[0] * uniform->probeCounts[1] * uniform->probeCounts[2];
    {
        LuxDDGIState lay{};
        shardLayout(*uniform, ci.rank, ci.world, ci.flags, &lay);
        c->probeCount  = lay.probeCount;
        c->probeBegin  = lay.probeBegin;
        c->layerProbes = lay.layerProbes;
        c->layerStride = lay.layerStride;
    }
    if (ci.stream)
        c->stream = (cudaStream_t)ci.stream;
    else
    {
        cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess)
        {
            delete c;
            return fail(LUX_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(se));
        }
        c->ownStream = true;
    }
    for (auto& ev : c->ev)
        cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->downStream, cudaStreamNonBlocking);
    for (cudaEvent_t& e : c->fences)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (cudaEvent_t* e : {&c->evLightReady, &c->evShadeDone, &c->evIrrDone, &c->evDepthDone, &c->evCopyDone, &c->evFork, &c->evJoin, &c->evWeights})
        cudaEventCreateWithFlags(e, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&c->auxStream, cudaStreamNonBlocking);
    rc = init::initializeProbeGrid(*c);
    if (rc == LUX_OK)
        rc = updateOrigins(*c);
    if (rc != LUX_OK)
    {
        lux_ddgi_destroy(c);
        return rc;
    }
    // default sky: the reference's 1x1 black fallback cube (DDGIRenderer.cpp:308)
    c->skyFace = 0;
    *out = c;
    return LUX_OK;
}

int lux_ddgi_destroy(LuxDDGIContext* c)
{
    if (!c)
        return LUX_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copyStream)
        cudaStreamSynchronize(c->copyStream);
    if (c->downStream)
        cudaStreamSynchronize(c->downStream);
    if (c->auxStream)
        cudaStreamSynchronize(c->auxStream);
    DeviceBuffer* all[] = {&c->radiance, &c->directionDepth, &c->irradiance[0], &c->irradiance[1], &c->depth[0], &c->depth[1],
                           &c->dirs, &c->blendLists, &c->ummaW[0], &c->ummaW[1], &c->wIrr, &c->wDepth, &c->scaleIrr, &c->scaleDepth, &c->nzIrr, &c->nzDepth, &c->origins, &c->records, &c->meta, &c->chunkCounter, &c->sortTicket, &c->binCounts, &c->binBlockSums, &c->sortedIdx, &c->dirsHalf, &c->sdf, &c->mip, &c->chunks, &c->cull,
                           &c->objects, &c->objectInverse, &c->chunkMasks, &c->tiles, &c->tileZRow, &c->light, &c->atlasDepth, &c->sky,
                           &c->unitOrder, &c->unitIndex, &c->rayOrder, &c->raySlot, &c->mipScratch, &c->probeTaps, &c->tcW[0], &c->tcW[1], &c->tcW[2], &c->tcW[3]};
    for (DeviceBuffer* b : all)
        b->release();
    releaseSdfTextures(*c);
    for (auto& ev : c->ev)
        if (ev)
            cudaEventDestroy(ev);
    for (cudaEvent_t e : {c->evLightReady, c->evShadeDone, c->evIrrDone, c->evDepthDone, c->evCopyDone, c->evFork, c->evJoin, c->evWeights})
        if (e)
            cudaEventDestroy(e);
    if (c->copyStream)
        cudaStreamDestroy(c->copyStream);
    if (c->downStream)
        cudaStreamDestroy(c->downStream);
    if (c->gatherStream)
    {
        cudaStreamSynchronize(c->gatherStream);
        cudaStreamDestroy(c->gatherStream);
        for (cudaEvent_t e : {c->evBlendDone, c->evGather[0], c->evGather[1]})
            if (e)
                cudaEventDestroy(e);
    }
    for (cudaEvent_t e : c->fences)
        if (e)
            cudaEventDestroy(e);
    if (c->auxStream)
        cudaStreamDestroy(c->auxStream);
    if (c->ownStream)
        cudaStreamDestroy(c->stream);
    delete c;
    return LUX_OK;
}

int lux_ddgi_set_uniform(LuxDDGIContext* c, const LuxDDGIUniform* u)
{
    CHECK_CTX(c);
    if (!u)
        return fail(LUX_ERR_INVALID_ARG, "null uniform");
    for (int i = 0; i < 3; i++)
        if (u->probeCounts[i] != c->uniform.probeCounts[i])
            return fail(LUX_ERR_UNSUPPORTED, "probeCounts changed; create a new context");
    if (u->raysPerProbe != c->uniform.raysPerProbe)
        return fail(LUX_ERR_UNSUPPORTED, "raysPerProbe changed; create a new context");
    int rc = init::validate(*u);
    if (rc != LUX_OK)
        return rc;
    c->uniform = *u;
    return updateOrigins(*c);
}

// (Re)creates the layered SDF textures over c->sdf / c->mip and publishes `data` as the bound GlobalSDFData.
static int bindSdfTextures(LuxDDGIContext* c, const LuxGlobalSDFData* data)
{
    const int res = (int)data->resolution, mres = res / 4;
    releaseSdfTextures(*c);
    // Default SDF path = layered-texture gathers (chosen by ncu, profiles/r1_*): falls back to explicit loads when asked to
    // (LUX_DDGI_FLAG_SDF_LOADS), for the simple kernel, or when the volume has more z-slices than a layered array allows.
    const bool wantTex = !(c->flags & (LUX_DDGI_FLAG_SDF_LOADS | LUX_DDGI_FLAG_TRACE_SIMPLE)) && res <= 2048;
    if ((c->flags & LUX_DDGI_FLAG_SDF_TEXTURE) && res > 2048)
        return fail(LUX_ERR_UNSUPPORTED, "layered SDF textures support at most 2048 z-slices");
    if (wantTex)
    {
        int rc;
        const int w = res * (int)data->cascadesCount, mw = mres * (int)data->cascadesCount;
        if ((rc = makeLayeredTexture(*c, c->sdf.ptr, w, res, res, &c->sdfArray, &c->sdfTex)) != LUX_OK) return rc;
        if ((rc = makeLayeredTexture(*c, c->mip.ptr, mw, mres, mres, &c->mipArray, &c->mipTex)) != LUX_OK) return rc;
    }
    c->sdfData    = *data;
    c->hasSdf     = true;
    c->masksDirty = true;
    c->probeTapsDirty = true;
    return LUX_OK;
}

int lux_ddgi_set_global_sdf(LuxDDGIContext* c, const LuxGlobalSDFData* data, const void* sdf, const void* mip, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!data || !sdf || !mip)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    if (data->cascadesCount < 1 || data->cascadesCount > LUX_MAX_CASCADES)
        return fail(LUX_ERR_INVALID_ARG, "cascadesCount %u out of range", data->cascadesCount);
    const int res = (int)data->resolution;
    if (res < 4 || (float)res != data->resolution || res % 4 != 0)
        return fail(LUX_ERR_INVALID_ARG, "resolution %g must be a positive multiple of 4", data->resolution);
    const size_t n    = (size_t)res * res * res * data->cascadesCount;
    const size_t mres = res / 4;
    const size_t nm   = mres * mres * mres * data->cascadesCount;
    int rc;
    if ((rc = upload(*c, c->sdf, sdf, n * 2, kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->mip, mip, nm * 2, kind)) != LUX_OK) return rc;
    rc = bindSdfTextures(c, data);
    if (rc != LUX_OK)
        return rc;
    if (kind == LUX_MEM_HOST)
        LUX_CUDA(cudaStreamSynchronize(c->stream)); // the caller may free its buffers on return
    return LUX_OK;
}

static int stageToDevice(LuxDDGIContext* c, const void* src, size_t bytes, LuxMemKind kind, void** dev, bool* owned);

// Copies the box [x0, x0+dx) x [y0, y0+dy) x [z0, z0+dz) of a linear R16F volume of row length `w` texels and `h` rows per slice into the same
// box of a layered array (layer = z).
static int refreshLayeredRegion(LuxDDGIContext* c, cudaArray_t arr, const void* linear, int w, int h, int x0, int y0, int z0, int dx, int dy, int dz)
{
    cudaMemcpy3DParms cp{};
    // the source box starts at the offset pointer (same pitch and slice height), so no srcPos is involved: the runtime documents pointer positions
    // in bytes but scales them by the array's element size once an array takes part in the copy (measured: a byte offset read the wrong columns)
    const uint16_t* src = (const uint16_t*)linear + ((size_t)z0 * h + y0) * w + x0;
    cp.srcPtr   = make_cudaPitchedPtr(const_cast<uint16_t*>(src), (size_t)w * 2, (size_t)w, (size_t)h);
    cp.dstArray = arr;
    cp.dstPos   = make_cudaPos((size_t)x0, (size_t)y0, (size_t)z0);     // array positions are in elements
    cp.extent   = make_cudaExtent((size_t)dx, (size_t)dy, (size_t)dz);
    cp.kind     = cudaMemcpyDeviceToDevice;
    LUX_CUDA(cudaMemcpy3DAsync(&cp, c->stream));
    return LUX_OK;
}

int lux_ddgi_update_global_sdf_region(LuxDDGIContext* c, uint32_t cascade, const int32_t chunkMin[3], const int32_t chunkMax[3], const void* texels,
                                      LuxMemKind kind, int32_t rebuildMip)
{
    CHECK_CTX(c);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound (lux_ddgi_set_global_sdf)");
    if (!chunkMin || !chunkMax || cascade >= c->sdfData.cascadesCount)
        return fail(LUX_ERR_INVALID_ARG, "bad region arguments");
    if (!texels && !c->sdf.borrowed)
        return fail(LUX_ERR_INVALID_ARG, "texels == NULL needs a caller-owned (LUX_MEM_DEVICE) volume that already holds the new texels");
    const int res = (int)c->sdfData.resolution, casc = (int)c->sdfData.cascadesCount, mres = res / 4, w = res * casc, mw = mres * casc;
    int lo[3], n[3];
    for (int i = 0; i < 3; i++)
    {
        if (chunkMin[i] > chunkMax[i])
            return fail(LUX_ERR_INVALID_ARG, "empty chunk range on axis %d", i);
        lo[i] = std::max(chunkMin[i], 0) * LUX_SDF_RASTERIZE_CHUNK_SIZE;
        const int hi = std::min((chunkMax[i] + 1) * LUX_SDF_RASTERIZE_CHUNK_SIZE, res);
        n[i] = hi - lo[i];
        if (n[i] <= 0)
            return LUX_OK; // wholly outside the volume: nothing to do (the reference's out-of-range chunks store nothing either)
    }
    const int x0 = (int)cascade * res + lo[0];
    // the trace of the previous frame may still read the volume: everything below is ordered behind it on the context's stream
    if (texels)
    {
        void* dT = nullptr;
        bool  owned = false;
        int   rc = stageToDevice(c, texels, (size_t)n[0] * n[1] * n[2] * 2, kind, &dT, &owned);
        if (rc != LUX_OK)
            return rc;
        if (c->sdf.borrowed)
        { // never write into a caller-owned volume: continue on a private copy
            void* own = nullptr;
            cudaError_t e = cudaMalloc(&own, c->sdf.bytes);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(own, c->sdf.ptr, c->sdf.bytes, cudaMemcpyDeviceToDevice, c->stream);
            if (e != cudaSuccess)
            {
                if (owned) cudaFree(dT);
                return fail(LUX_ERR_OUT_OF_MEMORY, "private copy of the global SDF: %s", cudaGetErrorString(e));
            }
            c->sdf.ptr      = own;
            c->sdf.borrowed = false;
        }
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(dT, (size_t)n[0] * 2, (size_t)n[0], (size_t)n[1]);
        cp.dstPtr = make_cudaPitchedPtr(c->sdf.ptr, (size_t)w * 2, (size_t)w, (size_t)res);
        cp.dstPos = make_cudaPos((size_t)x0 * 2, (size_t)lo[1], (size_t)lo[2]);
        cp.extent = make_cudaExtent((size_t)n[0] * 2, (size_t)n[1], (size_t)n[2]);
        cp.kind   = cudaMemcpyDeviceToDevice;
        cudaError_t e = cudaMemcpy3DAsync(&cp, c->stream);
        if (e == cudaSuccess && owned)
            e = cudaStreamSynchronize(c->stream);
        if (owned)
            cudaFree(dT);
        if (e != cudaSuccess)
            return fail(LUX_ERR_CUDA, "update_global_sdf_region: %s", cudaGetErrorString(e));
    }
    int rc;
    if (c->sdfArray && (rc = refreshLayeredRegion(c, c->sdfArray, c->sdf.ptr, w, res, x0, lo[1], lo[2], n[0], n[1], n[2])) != LUX_OK)
        return rc;
    if (rebuildMip)
    {
        if (c->mip.borrowed)
        { // the mip becomes library-owned the first time the library rebuilds it
            void* own = nullptr;
            LUX_CUDA(cudaMalloc(&own, c->mip.bytes));
            LUX_CUDA(cudaMemcpyAsync(own, c->mip.ptr, c->mip.bytes, cudaMemcpyDeviceToDevice, c->stream));
            c->mip.ptr      = own;
            c->mip.borrowed = false;
        }
        const size_t scratchBytes = (size_t)mres * mres * mres * 2;
        if (c->mipScratch.bytes != scratchBytes)
        {
            c->mipScratch.release();
            LUX_CUDA(cudaMalloc(&c->mipScratch.ptr, scratchBytes));
            c->mipScratch.bytes = scratchBytes;
        }
        uint16_t*   tmp = (uint16_t*)c->mipScratch.ptr;
        const int   k   = (int)cascade;
        const float cascadeMaxDistance = c->sdfData.cascadePosDistance[k][3] * 2.0f;
        lux::launch_sdf_fill(tmp, (size_t)mres * mres * mres, 0x3c00, c->stream);
        lux::launch_sdf_mip_pass((const uint16_t*)c->sdf.ptr, w, res, (uint16_t*)c->mip.ptr, mw, mres, mres, res, 4, k * res, k * mres, cascadeMaxDistance, c->stream);
        for (int i = 1; i < 5; i++)
        {
            if (i & 1)
                lux::launch_sdf_mip_pass((const uint16_t*)c->mip.ptr, mw, mres, tmp, mres, mres, mres, mres, 1, k * mres, 0, cascadeMaxDistance, c->stream);
            else
                lux::launch_sdf_mip_pass(tmp, mres, mres, (uint16_t*)c->mip.ptr, mw, mres, mres, mres, 1, 0, k * mres, cascadeMaxDistance, c->stream);
        }
        c->launches += 6;
        LUX_CUDA(cudaGetLastError());
        if (c->mipArray && (rc = refreshLayeredRegion(c, c->mipArray, c->mip.ptr, mw, mres, k * mres, 0, 0, mres, mres, mres)) != LUX_OK)
            return rc;
    }
    c->probeTapsDirty = true; // a probe may sit in the patched chunks
    return LUX_OK;
}

// ---- global SDF build ---------------------------------------------------------------------------
// Host sequencing of merge_sdf::system's first frame (GlobalDistanceField.cpp:575-848) with chunkCalculate (:460-533),
// getChunkId (:193-210) and fillFlood (:537-573).  Reference behaviours kept on purpose are listed in DESIGN.md §10.
namespace {

struct ChunkEntry
{
    int coord[3];
    int count;
    uint32_t models[LUX_SDF_RASTERIZE_MODEL_MAX_COUNT];
};

inline float clampf(float x, float lo, float hi) { float t = (x < lo) ? lo : x; return (hi < t) ? hi : t; } // glm::clamp = min(max(x, lo), hi)

// mip of every cascade of the bound (c->sdf) volume into c->mip: one 4x min-downsample + 4 flood passes through a temporary
int buildMipOnDevice(LuxDDGIContext* c, const LuxGlobalSDFData& data)
{
    const int res = (int)data.resolution, casc = (int)data.cascadesCount, mres = res / 4;
    uint16_t* tmp = nullptr;
    LUX_CUDA(cudaMalloc(&tmp, (size_t)mres * mres * mres * 2));
    lux::launch_sdf_fill(tmp, (size_t)mres * mres * mres, 0x3c00, c->stream); // cleared to 1.0 (:633-635)
    c->launches += 1;
    for (int k = 0; k < casc; k++)
    {
        const float cascadeMaxDistance = data.cascadePosDistance[k][3] * 2.0f;
        lux::launch_sdf_mip_pass((const uint16_t*)c->sdf.ptr, res * casc, res, (uint16_t*)c->mip.ptr, mres * casc, mres, mres, res, 4, k * res, k * mres,
                                 cascadeMaxDistance, c->stream);
        for (int i = 1; i < 5; i++)
        {
            if (i & 1) // Mip -> Tmp
                lux::launch_sdf_mip_pass((const uint16_t*)c->mip.ptr, mres * casc, mres, tmp, mres, mres, mres, mres, 1, k * mres, 0, cascadeMaxDistance, c->stream);
            else // Tmp -> Mip
                lux::launch_sdf_mip_pass(tmp, mres, mres, (uint16_t*)c->mip.ptr, mres * casc, mres, mres, mres, 1, 0, k * mres, cascadeMaxDistance, c->stream);
        }
        c->launches += 5;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(tmp);
    LUX_CUDA(e);
    LUX_CUDA(cudaGetLastError());
    return LUX_OK;
}

int validateSdfData(const LuxGlobalSDFData* data)
{
    if (!data)
        return fail(LUX_ERR_INVALID_ARG, "null GlobalSDFData");
    if (data->cascadesCount < 1 || data->cascadesCount > LUX_MAX_CASCADES)
        return fail(LUX_ERR_INVALID_ARG, "cascadesCount %u out of range", data->cascadesCount);
    const int res = (int)data->resolution;
    if (res < 4 || (float)res != data->resolution || res % 4 != 0)
        return fail(LUX_ERR_INVALID_ARG, "resolution %g must be a positive multiple of 4", data->resolution);
    return LUX_OK;
}

} // namespace

int lux_ddgi_build_sdf_mip(LuxDDGIContext* c)
{
    CHECK_CTX(c);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound");
    if (c->mip.borrowed)
        return fail(LUX_ERR_UNSUPPORTED, "the bound mip is caller-owned device memory; upload it from the host or build the SDF with lux_ddgi_build_global_sdf");
    int rc = buildMipOnDevice(c, c->sdfData);
    if (rc != LUX_OK)
        return rc;
    const LuxGlobalSDFData data = c->sdfData;
    return bindSdfTextures(c, &data);
}

int lux_ddgi_build_global_sdf(LuxDDGIContext* c, const LuxGlobalSDFData* data, const LuxMeshSDF* meshes, int32_t meshCount, float minObjectRadius)
{
    CHECK_CTX(c);
    int rc = validateSdfData(data);
    if (rc != LUX_OK)
        return rc;
    if (meshCount < 0 || (meshCount > 0 && !meshes))
        return fail(LUX_ERR_INVALID_ARG, "bad mesh list");
    const int res = (int)data->resolution, casc = (int)data->cascadesCount, texWidth = res * casc, mres = res / 4;
    const int rasterizeChunks = (res + LUX_SDF_RASTERIZE_CHUNK_SIZE - 1) / LUX_SDF_RASTERIZE_CHUNK_SIZE;
    for (int m = 0; m < meshCount; m++)
    {
        if (meshes[m].mipCount < 1 || meshes[m].mipCount > LUX_SDF_MESH_MAX_MIPS || meshes[m].mipCount < std::min(casc, 3))
            return fail(LUX_ERR_INVALID_ARG, "mesh %d has %d mips; cascade level min(cascade, 2) must exist", m, meshes[m].mipCount);
        for (int l = 0; l < meshes[m].mipCount; l++)
            if (!meshes[m].mips[l])
                return fail(LUX_ERR_INVALID_ARG, "mesh %d mip %d is null", m, l);
    }
    // volumes: caller-owned buffers are replaced by engine-owned ones, cleared to 1.0 (:633-635)
    const size_t n = (size_t)res * res * texWidth, nm = (size_t)mres * mres * mres * casc;
    for (DeviceBuffer* b : {&c->sdf, &c->mip})
    {
        const size_t bytes = (b == &c->sdf ? n : nm) * 2;
        if (b->borrowed || b->bytes != bytes)
        {
            b->release();
            LUX_CUDA(cudaMalloc(&b->ptr, bytes));
            b->bytes = bytes;
        }
    }
    lux::launch_sdf_fill((uint16_t*)c->sdf.ptr, n, 0x3c00, c->stream);
    lux::launch_sdf_fill((uint16_t*)c->mip.ptr, nm, 0x3c00, c->stream);
    c->launches += 2;

    // mesh volumes (all mips, back to back) and records to the device
    std::vector<lux::SdfMeshRecord> records((size_t)meshCount);
    std::vector<size_t>             levelOffset((size_t)meshCount * LUX_SDF_MESH_MAX_MIPS, 0);
    size_t texels = 0;
    for (int m = 0; m < meshCount; m++)
    {
        const LuxMeshSDF& ms = meshes[m];
        lux::SdfMeshRecord& r = records[(size_t)m];
        std::memcpy(r.aabbMin, ms.aabbMin, 12); std::memcpy(r.aabbMax, ms.aabbMax, 12);
        std::memcpy(r.localToUVWMul, ms.localToUVWMul, 12); std::memcpy(r.localToUVWAdd, ms.localToUVWAdd, 12);
        r.maxDistance = ms.maxDistance;
        std::memcpy(r.worldMatrix, ms.worldMatrix, 64);
        for (int l = 0; l < ms.mipCount; l++)
        {
            levelOffset[(size_t)m * LUX_SDF_MESH_MAX_MIPS + l] = texels;
            texels += (size_t)std::max(ms.size[0] >> l, 1u) * std::max(ms.size[1] >> l, 1u) * std::max(ms.size[2] >> l, 1u);
        }
    }
    uint16_t* dVolumes = nullptr;
    lux::SdfMeshRecord* dRecords = nullptr;
    LuxObjectRasterizeData* dObjects = nullptr;
    lux::SdfMeshLevel* dLevels = nullptr;
    lux::SdfChunkDispatch* dDispatch = nullptr;
    auto cleanup = [&]() { cudaFree(dVolumes); cudaFree(dRecords); cudaFree(dObjects); cudaFree(dLevels); cudaFree(dDispatch); };
#define LUX_CUDA_CLEAN(call)                                                              \
    do                                                                                    \
    {                                                                                     \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
        {                                                                                 \
            cleanup();                                                                    \
            return fail(LUX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));           \
        }                                                                                 \
    } while (0)
    LUX_CUDA_CLEAN(cudaMalloc(&dVolumes, std::max<size_t>(texels, 1) * 2));
    LUX_CUDA_CLEAN(cudaMalloc(&dRecords, std::max<size_t>(records.size(), 1) * sizeof(lux::SdfMeshRecord)));
    LUX_CUDA_CLEAN(cudaMalloc(&dObjects, std::max<size_t>(records.size(), 1) * sizeof(LuxObjectRasterizeData)));
    LUX_CUDA_CLEAN(cudaMalloc(&dLevels, std::max<size_t>(records.size(), 1) * sizeof(lux::SdfMeshLevel)));
    for (int m = 0; m < meshCount; m++)
        for (int l = 0; l < meshes[m].mipCount; l++)
        {
            const size_t off = levelOffset[(size_t)m * LUX_SDF_MESH_MAX_MIPS + l];
            const size_t cnt = (size_t)std::max(meshes[m].size[0] >> l, 1u) * std::max(meshes[m].size[1] >> l, 1u) * std::max(meshes[m].size[2] >> l, 1u);
            LUX_CUDA_CLEAN(cudaMemcpyAsync(dVolumes + off, meshes[m].mips[l], cnt * 2, cudaMemcpyHostToDevice, c->stream));
        }

    for (int k = 0; k < casc; k++)
    {
        const float D = data->cascadePosDistance[k][3], cascadeMaxDistance = D * 2.0f, voxel = data->cascadeVoxelSize[k];
        const float center[3] = {data->cascadePosDistance[k][0], data->cascadePosDistance[k][1], data->cascadePosDistance[k][2]};
        const float bmin[3] = {center[0] - D, center[1] - D, center[2] - D}, bmax[3] = {center[0] + D, center[1] + D, center[2] + D};
        const int   mipLevel = std::min(k, 2);
        std::vector<ChunkEntry>          chunks;
        std::vector<int>                 chunkSlot((size_t)rasterizeChunks * rasterizeChunks * rasterizeChunks, -1); // dense (z, y, x) -> index into `chunks`
        std::vector<int>                 included; // mesh index per object index
        for (int m = 0; m < meshCount; m++)
        {
            const LuxMeshSDF& ms = meshes[m];
            const float* t = ms.worldMatrix;
            // objectBounds = sdf.aabb.transform(world) (BoundingBox.cpp:10-22; abs() of the product in the third term is the reference's)
            float cen[3], oldEdge[3], newCenter[3], newEdge[3], omn[3], omx[3];
            for (int i = 0; i < 3; i++)
            {
                cen[i]     = (ms.aabbMax[i] + ms.aabbMin[i]) * 0.5f;
                oldEdge[i] = (ms.aabbMax[i] - ms.aabbMin[i]) * 0.5f;
            }
            for (int r = 0; r < 3; r++)
            {
                newCenter[r] = ((t[0 + r] * cen[0] + t[4 + r] * cen[1]) + t[8 + r] * cen[2]) + t[12 + r] * 1.0f;
                newEdge[r]   = std::fabs(t[0 + r]) * oldEdge[0] + std::fabs(t[4 + r]) * oldEdge[1] + std::fabs(t[8 + r] * oldEdge[2]);
                omn[r]       = newCenter[r] - newEdge[r];
                omx[r]       = newCenter[r] + newEdge[r];
            }
            // BoundingSphere(objectBounds) against the cascade box, minimum radius (:692-701)
            float sc[3], dv[3], diag[3];
            for (int i = 0; i < 3; i++)
            {
                sc[i]   = (omx[i] + omn[i]) * 0.5f;
                dv[i]   = sc[i] - clampf(sc[i], bmin[i], bmax[i]);
                diag[i] = omx[i] - omn[i];
            }
            const float radius = std::sqrt((diag[0] * diag[0] + diag[1] * diag[1]) + diag[2] * diag[2]) / 2.0f;
            const float dist2  = (dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2];
            if (!(dist2 <= radius * radius && radius >= minObjectRadius))
                continue;
            // getChunkId (:193-210): margin, bias, truncating division; its clamps are no-ops
            const float objectMargin = voxel * (float)LUX_SDF_RASTERIZE_CHUNK_MARGIN, chunkSize = voxel * (float)LUX_SDF_RASTERIZE_CHUNK_SIZE;
            int cmin[3], cmax[3];
            for (int i = 0; i < 3; i++)
            {
                const float biasMin = bmin[i] + 0.1f;
                cmin[i] = (int)(((omn[i] - objectMargin) - biasMin) / chunkSize);
                cmax[i] = (int)(((omx[i] + objectMargin) - biasMin) / chunkSize);
            }
            const uint32_t objectIndex = (uint32_t)included.size();
            included.push_back(m);
            // The reference registers the object with every chunk of the UNCLAMPED range (its clamps are no-ops); chunks outside the volume are
            // dropped below and do not influence the lists of the chunks inside it, so only the in-range part is walked (an object much larger
            // than the cascade would otherwise cost (extent / chunkSize)^3 entries).
            for (int i = 0; i < 3; i++)
            {
                cmin[i] = std::max(cmin[i], 0);
                cmax[i] = std::min(cmax[i], rasterizeChunks - 1);
            }
            for (int z = cmin[2]; z <= cmax[2]; z++)
                for (int y = cmin[1]; y <= cmax[1]; y++)
                    for (int x = cmin[0]; x <= cmax[0]; x++)
                    {
                        int& slot = chunkSlot[((size_t)z * rasterizeChunks + y) * rasterizeChunks + x];
                        if (slot < 0)
                        {
                            slot = (int)chunks.size();
                            chunks.push_back(ChunkEntry{{x, y, z}, 0, {}});
                        }
                        ChunkEntry* ch = &chunks[(size_t)slot];
                        if (ch->count == LUX_SDF_RASTERIZE_MODEL_MAX_COUNT)
                            ch->count = 0; // :515-519 copies the empty next-layer entry over the full one: the list restarts
                        ch->models[ch->count++] = objectIndex;
                    }
        }
        // per-object device data of this cascade: records of the included meshes in object order, their mip level
        std::vector<lux::SdfMeshRecord> objRecords(included.size());
        std::vector<lux::SdfMeshLevel>  levels(included.size());
        for (size_t i = 0; i < included.size(); i++)
        {
            const int m = included[i];
            objRecords[i] = records[(size_t)m];
            levels[i]     = lux::SdfMeshLevel{dVolumes + levelOffset[(size_t)m * LUX_SDF_MESH_MAX_MIPS + mipLevel], (int)std::max(meshes[m].size[0] >> mipLevel, 1u),
                                              (int)std::max(meshes[m].size[1] >> mipLevel, 1u), (int)std::max(meshes[m].size[2] >> mipLevel, 1u)};
        }
        std::vector<lux::SdfChunkDispatch> dispatches;
        for (const ChunkEntry& e : chunks)
        {
            if (e.coord[0] < 0 || e.coord[1] < 0 || e.coord[2] < 0 || e.coord[0] >= rasterizeChunks || e.coord[1] >= rasterizeChunks || e.coord[2] >= rasterizeChunks)
                continue; // wholly outside the volume: every store would be out of bounds
            lux::SdfChunkDispatch d{};
            for (int i = 0; i < 3; i++)
                d.coord[i] = e.coord[i] * LUX_SDF_RASTERIZE_CHUNK_SIZE;
            d.count = e.count;
            std::memcpy(d.models, e.models, sizeof(d.models));
            d.read = 0; // layer 0 = SDFRasterizeModelNoRead; additive layers never receive models (see above)
            dispatches.push_back(d);
        }
        if (!included.empty())
        {
            LUX_CUDA_CLEAN(cudaMemcpyAsync(dRecords, objRecords.data(), objRecords.size() * sizeof(lux::SdfMeshRecord), cudaMemcpyHostToDevice, c->stream));
            LUX_CUDA_CLEAN(cudaMemcpyAsync(dLevels, levels.data(), levels.size() * sizeof(lux::SdfMeshLevel), cudaMemcpyHostToDevice, c->stream));
            lux::launch_sdf_object_data(dRecords, (int)included.size(), k, dObjects, c->stream);
            c->launches += 1;
        }
        if (!dispatches.empty())
        {
            cudaFree(dDispatch);
            dDispatch = nullptr;
            LUX_CUDA_CLEAN(cudaMalloc(&dDispatch, dispatches.size() * sizeof(lux::SdfChunkDispatch)));
            LUX_CUDA_CLEAN(cudaMemcpyAsync(dDispatch, dispatches.data(), dispatches.size() * sizeof(lux::SdfChunkDispatch), cudaMemcpyHostToDevice, c->stream));
            lux::SdfRasterizeParams rp{};
            for (int i = 0; i < 3; i++)
            {
                rp.mul[i] = (bmax[i] - bmin[i]) / (float)res;
                rp.add[i] = bmin[i] + voxel * 0.5f;
            }
            rp.maxDistance  = cascadeMaxDistance;
            rp.res          = res;
            rp.cascadeIndex = k;
            rp.texWidth     = texWidth;
            rp.objects      = dObjects;
            rp.levels       = dLevels;
            rp.dispatches   = dDispatch;
            rp.sdf          = (uint16_t*)c->sdf.ptr;
            lux::launch_sdf_rasterize(rp, (int)dispatches.size(), c->stream);
            c->launches += 1;
        }
        LUX_CUDA_CLEAN(cudaStreamSynchronize(c->stream)); // host vectors of this cascade go out of scope
    }
#undef LUX_CUDA_CLEAN
    cleanup();
    LUX_CUDA(cudaGetLastError());
    rc = buildMipOnDevice(c, *data);
    if (rc != LUX_OK)
        return rc;
    return bindSdfTextures(c, data);
}

// .sdf reader (cereal BinaryOutputArchive of `uvec3 size, int32 mipCount, vector<uint8_t>` + one vector per further mip, SDFBaker.cpp:158-204)
int lux_ddgi_sdf_file_read(const char* path, uint32_t size[3], int32_t* mipCount, uint64_t* texels, uint16_t* out)
{
    if (!path || !size || !mipCount || !texels)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return fail(LUX_ERR_INVALID_ARG, "cannot open %s", path);
    auto bad = [&](const char* why) {
        std::fclose(f);
        return fail(LUX_ERR_INVALID_ARG, "%s: %s", path, why);
    };
    std::fseek(f, 0, SEEK_END);
    const long fileBytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    uint32_t hdr[4];
    if (std::fread(hdr, 4, 4, f) != 4)
        return bad("truncated header");
    size[0] = hdr[0]; size[1] = hdr[1]; size[2] = hdr[2];
    *mipCount = (int32_t)hdr[3];
    if (*mipCount < 1 || *mipCount > LUX_SDF_MESH_MAX_MIPS || !size[0] || !size[1] || !size[2])
        return bad("implausible size / mip count");
    uint64_t total = 0;
    for (int l = 0; l < *mipCount; l++)
    {
        uint64_t n = 0;
        if (std::fread(&n, 8, 1, f) != 1)
            return bad("truncated mip header");
        const uint64_t want = (uint64_t)std::max(size[0] >> l, 1u) * std::max(size[1] >> l, 1u) * std::max(size[2] >> l, 1u);
        if (n != want * 2)
            return bad("mip byte count does not match the volume size");
        if (out)
        {
            if (std::fread(out + total, 2, want, f) != want)
                return bad("truncated mip data");
        }
        else if (std::fseek(f, (long)n, SEEK_CUR) != 0 || std::ftell(f) > fileBytes)
            return bad("truncated mip data");
        total += want;
    }
    *texels = total;
    std::fclose(f);
    return LUX_OK;
}

int lux_ddgi_set_surface_atlas(LuxDDGIContext* c, const LuxGlobalSurfaceAtlasData* data, const uint32_t* chunks,
                               const uint32_t* cullObjects, size_t cullObjectsCount, const LuxObjectBuffer* objects, size_t objectsCount,
                               const LuxTileBuffer* tiles, size_t tilesCount, const void* light, const float* depth, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!data)
    {
        c->hasAtlas = false; // unbind: hits return zero radiance
        return LUX_OK;
    }
    if (!chunks || !cullObjects || !objects || !tiles || !light || !depth)
        return fail(LUX_ERR_INVALID_ARG, "null surface-cache buffer");
    if (data->resolution == 0 || !(data->chunkSize > 0.0f))
        return fail(LUX_ERR_INVALID_ARG, "bad GlobalSurfaceAtlasData (resolution %u, chunkSize %g)", data->resolution, data->chunkSize);
    if (data->objectsCount > objectsCount)
        return fail(LUX_ERR_INVALID_ARG, "objectsCount %u exceeds the object buffer (%zu)", data->objectsCount, objectsCount);
    const size_t nchunks = (size_t)LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    const size_t texels  = (size_t)data->resolution * data->resolution;
    int rc;
    if ((rc = upload(*c, c->chunks, chunks, nchunks * 4, kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->cull, cullObjects, cullObjectsCount * 4, kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->objects, objects, objectsCount * sizeof(LuxObjectBuffer), kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->tiles, tiles, tilesCount * sizeof(LuxTileBuffer), kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->light, light, texels * 8, kind)) != LUX_OK) return rc;
    if ((rc = upload(*c, c->atlasDepth, depth, texels * 4, kind)) != LUX_OK) return rc;
    if (c->objectInverse.bytes != objectsCount * 64)
    {
        c->objectInverse.release();
        LUX_CUDA(cudaMalloc(&c->objectInverse.ptr, objectsCount ? objectsCount * 64 : 1));
        c->objectInverse.bytes = objectsCount * 64;
    }
    lux::launch_object_inverse((const LuxObjectBuffer*)c->objects.ptr, (int)objectsCount, (float*)c->objectInverse.ptr, c->stream);
    c->launches += objectsCount ? 1 : 0;
    if (c->tileZRow.bytes != tilesCount * 16)
    {
        c->tileZRow.release();
        LUX_CUDA(cudaMalloc(&c->tileZRow.ptr, tilesCount ? tilesCount * 16 : 1));
        c->tileZRow.bytes = tilesCount * 16;
    }
    lux::launch_tile_zrow((const LuxTileBuffer*)c->tiles.ptr, (int)tilesCount, (float4*)c->tileZRow.ptr, c->stream);
    c->launches += tilesCount ? 1 : 0;
    LUX_CUDA(cudaGetLastError());
    if (kind == LUX_MEM_HOST)
        LUX_CUDA(cudaStreamSynchronize(c->stream));
    c->atlasData  = *data;
    c->hasAtlas   = true;
    c->masksDirty = true;
    return LUX_OK;
}

int lux_ddgi_update_surface_light_cache(LuxDDGIContext* c, const void* light, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!c->hasAtlas)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    if (!light)
        return fail(LUX_ERR_INVALID_ARG, "null light cache");
    const size_t texels = (size_t)c->atlasData.resolution * c->atlasData.resolution;
    if (kind == LUX_MEM_DEVICE)
        return upload(*c, c->light, light, texels * 8, kind);
    if (c->light.borrowed || c->light.bytes != texels * 8)
    {
        c->light.release();
        LUX_CUDA(cudaMalloc(&c->light.ptr, texels * 8));
        c->light.bytes = texels * 8;
    }
    // Copy-engine overlap: the upload runs on the copy stream once the previous frame's shade kernel (the only reader) is
    // done, and only the NEXT shade kernel waits for it — the march of that frame overlaps the transfer.
    LUX_CUDA(cudaStreamWaitEvent(c->copyStream, c->evShadeDone, 0));
    LUX_CUDA(cudaMemcpyAsync(c->light.ptr, light, texels * 8, cudaMemcpyHostToDevice, c->copyStream));
    LUX_CUDA(cudaEventRecord(c->evLightReady, c->copyStream));
    c->lightPending = true;
    return LUX_OK;
}

// Sharded relight: every rank uploads only ITS rows of the (replicated) light cache and the ranks exchange them over NVLink, so the
// host -> device traffic of a frame is one light cache in total instead of one per GPU.
int lux_ddgi_update_surface_light_cache_rows(LuxDDGIContext* c, const void* lightRows, int32_t rowBegin, int32_t rowCount, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!c->hasAtlas)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    const int32_t res = (int32_t)c->atlasData.resolution;
    if (!lightRows || rowBegin < 0 || rowCount <= 0 || rowBegin + rowCount > res)
        return fail(LUX_ERR_INVALID_ARG, "bad light-cache row range [%d,%d) of %d", rowBegin, rowBegin + rowCount, res);
    const size_t rowBytes = (size_t)res * 8;
    if (c->light.borrowed)
    { // never write into a caller-owned buffer: continue on a private copy
        lightAcquire(c);
        void* own = nullptr;
        LUX_CUDA(cudaMalloc(&own, (size_t)res * rowBytes));
        LUX_CUDA(cudaMemcpyAsync(own, c->light.ptr, (size_t)res * rowBytes, cudaMemcpyDeviceToDevice, c->stream));
        LUX_CUDA(cudaStreamSynchronize(c->stream));
        c->light.ptr      = own;
        c->light.borrowed = false;
        c->light.bytes    = (size_t)res * rowBytes;
    }
    if (c->ncclComm && (rowCount * c->world != res || rowBegin != c->rank * rowCount))
        return fail(LUX_ERR_INVALID_ARG, "with a communicator bound, rank r must pass rows [r*res/world, (r+1)*res/world)");
    char* base = (char*)c->light.ptr;
    LUX_CUDA(cudaStreamWaitEvent(c->copyStream, c->evShadeDone, 0)); // the previous frame's shade is the only reader
    LUX_CUDA(cudaMemcpyAsync(base + (size_t)rowBegin * rowBytes, lightRows, (size_t)rowCount * rowBytes,
                             kind == LUX_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->copyStream));
    if (c->ncclComm)
    {
        NcclApi* n = ncclApi();
        int rc = n->allGather(base + (size_t)rowBegin * rowBytes, base, (size_t)rowCount * rowBytes, /*ncclUint8*/ 1, c->ncclComm, c->copyStream);
        if (rc != 0)
            return fail(LUX_ERR_CUDA, "ncclAllGather(light cache): %s", n->getErrorString(rc));
    }
    LUX_CUDA(cudaEventRecord(c->evLightReady, c->copyStream));
    c->lightPending = true;
    return LUX_OK;
}

// surface::culling (GlobalSurfaceAtlas.cpp:607-641) + SDFCulling.comp: rebuilds the chunk / culled-object lists of the bound surface
// cache ON DEVICE from its object buffer.  `capacityWords` = GlobalSurfaceAtlasData.culledObjectsCapacity of the shader (lists that
// do not fit are dropped, as there); 0 = size the buffer so that every list fits.
int lux_ddgi_cull_surface_objects(LuxDDGIContext* c, uint32_t capacityWords)
{
    CHECK_CTX(c);
    if (!c->hasAtlas)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    const size_t nchunks = (size_t)LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    uint32_t* scratch = nullptr; // 65 536 sizes + 16 block sums + 1 total
    LUX_CUDA(cudaMalloc(&scratch, (65536 + 16 + 1) * sizeof(uint32_t)));
    auto done = [&](int rc) { cudaFree(scratch); return rc; };
    // worst case: every chunk lists every object
    const size_t worst = 1 + nchunks * ((size_t)c->atlasData.objectsCount + 1);
    size_t words = capacityWords ? (size_t)capacityWords : worst;
    if (capacityWords == 0 && worst > (size_t(1) << 26))
    { // size it from the counts instead of the worst case: one extra counting pass
        std::vector<uint32_t> tmp(1);
        DeviceBuffer probe;
        uint32_t*    dummyCull = nullptr;
        if (cudaMalloc(&probe.ptr, nchunks * 4) != cudaSuccess || cudaMalloc(&dummyCull, 16) != cudaSuccess)
        {
            probe.release();
            return done(fail(LUX_ERR_OUT_OF_MEMORY, "surface cull sizing scratch"));
        }
        probe.bytes = nchunks * 4;
        lux::launch_surface_cull((const LuxObjectBuffer*)c->objects.ptr, c->atlasData.objectsCount, c->atlasData.chunkSize, 0u, scratch, scratch + 65536,
                                 scratch + 65536 + 16, (uint32_t*)probe.ptr, dummyCull, 1u, c->stream); // capacity 0: nothing is written but cull[0]
        c->launches += 5;
        cudaError_t e = cudaMemcpyAsync(tmp.data(), dummyCull, 4, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
        cudaFree(dummyCull);
        probe.release();
        if (e != cudaSuccess)
            return done(fail(LUX_ERR_CUDA, "surface cull sizing: %s", cudaGetErrorString(e)));
        words = tmp[0];
    }
    for (DeviceBuffer* b : {&c->chunks, &c->cull})
    {
        const size_t bytes = (b == &c->chunks ? nchunks : words) * 4;
        if (b->borrowed || b->bytes != bytes)
        {
            b->release();
            cudaError_t e = cudaMalloc(&b->ptr, bytes);
            if (e != cudaSuccess)
                return done(fail(LUX_ERR_OUT_OF_MEMORY, "surface cull lists (%zu bytes): %s", bytes, cudaGetErrorString(e)));
            b->bytes = bytes;
        }
    }
    cudaStreamWaitEvent(c->stream, c->evShadeDone, 0);
    lux::launch_surface_cull((const LuxObjectBuffer*)c->objects.ptr, c->atlasData.objectsCount, c->atlasData.chunkSize, (uint32_t)std::min<size_t>(words, 0xffffffffu),
                             scratch, scratch + 65536, scratch + 65536 + 16, (uint32_t*)c->chunks.ptr, (uint32_t*)c->cull.ptr,
                             (uint32_t)std::min<size_t>(words, 0xffffffffu), c->stream);
    c->launches += 5;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return done(fail(LUX_ERR_CUDA, "surface cull: %s", cudaGetErrorString(e)));
    c->atlasData.culledObjectsCapacity = (uint32_t)std::min<size_t>(words, 0xffffffffu);
    c->masksDirty = true;
    return done(LUX_OK);
}

int lux_ddgi_get_surface_cull_lists(LuxDDGIContext* c, void** chunksDevice, void** cullDevice, size_t* cullWords)
{
    CHECK_CTX(c);
    if (!c->hasAtlas || !chunksDevice || !cullDevice)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    *chunksDevice = c->chunks.ptr;
    *cullDevice   = c->cull.ptr;
    if (cullWords)
        *cullWords = c->cull.bytes / 4;
    return LUX_OK;
}

int lux_ddgi_set_skybox(LuxDDGIContext* c, int32_t faceSize, const void* faces, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (faceSize <= 0 || !faces)
    {
        c->skyFace = 0;
        c->sky.release();
        return LUX_OK;
    }
    int rc = upload(*c, c->sky, faces, (size_t)6 * faceSize * faceSize * 8, kind);
    if (rc != LUX_OK)
        return rc;
    if (kind == LUX_MEM_HOST)
        LUX_CUDA(cudaStreamSynchronize(c->stream));
    c->skyFace = faceSize;
    return LUX_OK;
}

int lux_ddgi_set_ray_buffers(LuxDDGIContext* c, const void* radiance, const void* directionDistance, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!radiance || !directionDistance)
        return fail(LUX_ERR_INVALID_ARG, "null ray buffer");
    const size_t bytes = (size_t)c->probeCount * c->uniform.raysPerProbe * 8;
    cudaMemcpyKind k = kind == LUX_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    LUX_CUDA(cudaMemcpyAsync(c->radiance.ptr, radiance, bytes, k, c->stream));
    LUX_CUDA(cudaMemcpyAsync(c->directionDepth.ptr, directionDistance, bytes, k, c->stream));
    // caller-provided rays are not scanned: the blend takes its guarded loop (exact either way, ~4 % slower), see blend_irradiance_kernel
    LUX_CUDA(cudaMemsetAsync((unsigned int*)c->chunkCounter.ptr + 2, 0xff, sizeof(unsigned int), c->stream));
    LUX_CUDA(cudaStreamSynchronize(c->stream));
    c->raysValid = true;
    return LUX_OK;
}

int lux_ddgi_trace_rays(LuxDDGIContext* c, const LuxTracePushConstants* push)
{
    CHECK_CTX(c);
    if (!push)
        return fail(LUX_ERR_INVALID_ARG, "null push constants");
    return trace_rays::system(*c, *push);
}

int lux_ddgi_probe_update(LuxDDGIContext* c)
{
    CHECK_CTX(c);
    return probe_update::system(*c);
}

int lux_ddgi_border_update(LuxDDGIContext* c)
{
    CHECK_CTX(c);
    return border_update::system(*c);
}

int lux_ddgi_end_frame(LuxDDGIContext* c)
{
    CHECK_CTX(c);
    return end_frame::system(*c);
}

// Exchange step: one in-place all-gather per atlas (own slab rows -> every rank's full atlas) on the gather stream, right after the
// blend.  It overlaps the next frame's trace, which never reads the atlases; the blend that next overwrites this pair waits for it.
static int enqueueAllGather(LuxDDGIContext* c)
{
    if (!c->ncclComm)
        return LUX_OK;
    NcclApi* n = ncclApi();
    LuxDDGIState st{};
    shardLayout(c->uniform, c->rank, c->world, c->flags, &st);
    const int    w        = c->lastWritten;
    const size_t irrRow   = (size_t)c->uniform.irradianceTextureWidth * 8, depRow = (size_t)c->uniform.depthTextureWidth * 4;
    char*        irr      = (char*)c->irradiance[w].ptr;
    char*        dep      = (char*)c->depth[w].ptr;
    LUX_CUDA(cudaEventRecord(c->evBlendDone, c->stream));
    LUX_CUDA(cudaStreamWaitEvent(c->gatherStream, c->evBlendDone, 0));
    int rc = n->groupStart();
    // z-slabs: one all-gather per atlas.  Interleaved blocks: round k gathers the blocks k * world .. k * world + world - 1 (contiguous rows), this rank
    // contributing its k-th block; all rounds of both atlases in one group.
    const int rounds = st.layerStride == 1 ? 1 : st.probeCount / st.layerProbes;
    const int irrS = (LUX_IRRADIANCE_OCT_SIZE + 2) * st.unitLayers, depS = (LUX_DEPTH_OCT_SIZE + 2) * st.unitLayers; // rows per interleave unit
    for (int k = 0; k < rounds && rc == 0; k++)
    {
        const size_t irrOwn = st.layerStride == 1 ? (size_t)st.irradianceRowBegin : (size_t)st.irradianceRowBegin + (size_t)k * st.layerStride * irrS;
        const size_t irrAll = st.layerStride == 1 ? 1 : 1 + (size_t)k * st.layerStride * irrS;
        const size_t irrCnt = st.layerStride == 1 ? (size_t)st.irradianceRowCount : (size_t)irrS;
        rc = n->allGather(irr + irrOwn * irrRow, irr + irrAll * irrRow, irrCnt * irrRow, /*ncclUint8*/ 1, c->ncclComm, c->gatherStream);
        if (rc != 0)
            break;
        const size_t depOwn = st.layerStride == 1 ? (size_t)st.depthRowBegin : (size_t)st.depthRowBegin + (size_t)k * st.layerStride * depS;
        const size_t depAll = st.layerStride == 1 ? 1 : 1 + (size_t)k * st.layerStride * depS;
        const size_t depCnt = st.layerStride == 1 ? (size_t)st.depthRowCount : (size_t)depS;
        rc = n->allGather(dep + depOwn * depRow, dep + depAll * depRow, depCnt * depRow, 1, c->ncclComm, c->gatherStream);
    }
    const int rcEnd = n->groupEnd();
    if (rc != 0 || rcEnd != 0)
        return fail(LUX_ERR_CUDA, "ncclAllGather: %s", n->getErrorString(rc != 0 ? rc : rcEnd));
    LUX_CUDA(cudaEventRecord(c->evGather[w], c->gatherStream));
    c->gatherPending[w] = true;
    return LUX_OK;
}

// Readers of the WHOLE current atlas (consumers, full downloads) run after the exchange of the frame that wrote it.
static void waitAllGather(LuxDDGIContext* c, cudaStream_t s)
{
    for (int w = 0; w < 2; w++)
        if (c->gatherPending[w])
            cudaStreamWaitEvent(s, c->evGather[w], 0);
}

int lux_ddgi_update(LuxDDGIContext* c, const float orientation[16])
{
    CHECK_CTX(c);
    if (!orientation)
        return fail(LUX_ERR_INVALID_ARG, "null orientation");
    LuxTracePushConstants push{};
    std::memcpy(push.randomOrientation, orientation, sizeof(float) * 16);
    push.numFrames       = (uint32_t)c->frames;
    push.infiniteBounces = c->frames != 0 ? 1u : 0u; // DDGIRenderer.cpp:263
    push.numLights       = 0;
    push.intensity       = 1.0f;
    if (!(c->flags & (LUX_DDGI_FLAG_STAGE_TIMERS | LUX_DDGI_FLAG_UNFUSED_BORDER | LUX_DDGI_FLAG_NO_PIPELINE | LUX_DDGI_FLAG_TRACE_SIMPLE)))
    { // Shipped form of the update.  The blend weights depend on the frame's ray directions only, so they are computed on the
      // auxiliary stream while the march runs.
        int rc = trace_rays::setup(*c, push);
        if (rc != LUX_OK)
            return rc;
        cudaEventRecord(c->evFork, c->stream);
        cudaStreamWaitEvent(c->auxStream, c->evFork, 0);
        probe_update::weights(*c, (const uint2*)c->dirsHalf.ptr, c->auxStream);
        cudaEventRecord(c->evWeights, c->auxStream);
        trace_rays::launch(*c, c->stream, false);
        cudaEventRecord(c->evShadeDone, c->stream); // the copy engine may start on each output as soon as its kernel is done
        cudaStreamWaitEvent(c->stream, c->evWeights, 0);
        cudaStreamWaitEvent(c->stream, c->evCopyDone, 0); // only the blend overwrites atlas rows an earlier frame's download may still read
        if (c->gatherPending[1 - c->pingPong])            // ... or the all-gather of two frames ago
            cudaStreamWaitEvent(c->stream, c->evGather[1 - c->pingPong], 0);
        probe_update::launch(*c, c->stream, c->evIrrDone, c->evDepthDone);
        c->lightPending = false;
        LUX_CUDA(cudaGetLastError());
        c->raysValid   = true;
        c->lastWritten = 1 - c->pingPong;
        c->timed       = false;
        rc = enqueueAllGather(c);
        if (rc != LUX_OK)
            return rc;
        return end_frame::system(*c);
    }
    int rc = trace_rays::system(*c, push);
    if (rc != LUX_OK)
        return rc;
    rc = probe_update::system(*c);
    if (rc != LUX_OK)
        return rc;
    if (c->flags & LUX_DDGI_FLAG_UNFUSED_BORDER)
    {
        rc = border_update::system(*c);
        if (rc != LUX_OK)
            return rc;
    }
    mark(*c, 4);
    c->timed = (c->flags & LUX_DDGI_FLAG_STAGE_TIMERS) != 0;
    rc = enqueueAllGather(c);
    if (rc != LUX_OK)
        return rc;
    return end_frame::system(*c);
}

int lux_ddgi_set_nccl_comm(LuxDDGIContext* c, void* ncclComm)
{
    CHECK_CTX(c);
    if (!ncclComm)
    {
        c->ncclComm = nullptr;
        return LUX_OK;
    }
    if (c->world < 2)
        return fail(LUX_ERR_INVALID_ARG, "a communicator needs a context created with world > 1");
    if (!ncclApi())
        return fail(LUX_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "missing symbols");
    if (!c->gatherStream)
    {
        LUX_CUDA(cudaStreamCreateWithFlags(&c->gatherStream, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&c->evBlendDone, &c->evGather[0], &c->evGather[1]})
            LUX_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    c->ncclComm = ncclComm;
    return LUX_OK;
}

int lux_ddgi_synchronize(LuxDDGIContext* c)
{
    CHECK_CTX(c);
    LUX_CUDA(cudaStreamSynchronize(c->stream));
    LUX_CUDA(cudaStreamSynchronize(c->copyStream));
    LUX_CUDA(cudaStreamSynchronize(c->downStream));
    if (c->gatherStream)
        LUX_CUDA(cudaStreamSynchronize(c->gatherStream));
    return LUX_OK;
}

static int bufferOf(LuxDDGIContext* c, LuxBufferId id, DeviceBuffer** out)
{
    switch (id)
    {
    case LUX_BUF_RADIANCE: *out = &c->radiance; break;
    case LUX_BUF_DIRECTION_DISTANCE: *out = &c->directionDepth; break;
    case LUX_BUF_IRRADIANCE: *out = &c->irradiance[c->lastWritten]; break;
    case LUX_BUF_DEPTH: *out = &c->depth[c->lastWritten]; break;
    case LUX_BUF_IRRADIANCE_PREV: *out = &c->irradiance[1 - c->lastWritten]; break;
    case LUX_BUF_DEPTH_PREV: *out = &c->depth[1 - c->lastWritten]; break;
    case LUX_BUF_GLOBAL_SDF: *out = &c->sdf; break;
    case LUX_BUF_GLOBAL_SDF_MIP: *out = &c->mip; break;
    default: return fail(LUX_ERR_INVALID_ARG, "unknown buffer id %d", (int)id);
    }
    return LUX_OK;
}

int lux_ddgi_get_buffer(LuxDDGIContext* c, LuxBufferId id, void** devicePtr, size_t* bytes)
{
    CHECK_CTX(c);
    if (!devicePtr)
        return fail(LUX_ERR_INVALID_ARG, "null output pointer");
    DeviceBuffer* b = nullptr;
    int rc = bufferOf(c, id, &b);
    if (rc != LUX_OK)
        return rc;
    *devicePtr = b->ptr;
    if (bytes)
        *bytes = b->bytes;
    return LUX_OK;
}

int lux_ddgi_download(LuxDDGIContext* c, LuxBufferId id, void* host, size_t bytes)
{
    CHECK_CTX(c);
    if (!host)
        return fail(LUX_ERR_INVALID_ARG, "null host pointer");
    DeviceBuffer* b = nullptr;
    int rc = bufferOf(c, id, &b);
    if (rc != LUX_OK)
        return rc;
    if (bytes != b->bytes)
        return fail(LUX_ERR_INVALID_ARG, "size mismatch: buffer holds %zu bytes, caller passed %zu", b->bytes, bytes);
    waitAllGather(c, c->stream);
    LUX_CUDA(cudaMemcpyAsync(host, b->ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    LUX_CUDA(cudaStreamSynchronize(c->stream));
    return LUX_OK;
}

int lux_ddgi_download_async(LuxDDGIContext* c, LuxBufferId id, void* pinnedHost, size_t bytes)
{
    CHECK_CTX(c);
    if (!pinnedHost)
        return fail(LUX_ERR_INVALID_ARG, "null host pointer");
    DeviceBuffer* b = nullptr;
    int rc = bufferOf(c, id, &b);
    if (rc != LUX_OK)
        return rc;
    if (bytes != b->bytes)
        return fail(LUX_ERR_INVALID_ARG, "size mismatch: buffer holds %zu bytes, caller passed %zu", b->bytes, bytes);
    waitAllGather(c, c->stream);
    LUX_CUDA(cudaMemcpyAsync(pinnedHost, b->ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    return LUX_OK;
}

int lux_ddgi_download_rows_async(LuxDDGIContext* c, LuxBufferId id, int32_t rowBegin, int32_t rowCount, void* pinnedHost)
{
    CHECK_CTX(c);
    if (!pinnedHost)
        return fail(LUX_ERR_INVALID_ARG, "null host pointer");
    DeviceBuffer* b = nullptr;
    int rc = bufferOf(c, id, &b);
    if (rc != LUX_OK)
        return rc;
    size_t rowBytes;
    int    rows;
    if (id == LUX_BUF_IRRADIANCE || id == LUX_BUF_IRRADIANCE_PREV)
    {
        rowBytes = (size_t)c->uniform.irradianceTextureWidth * 8;
        rows     = c->uniform.irradianceTextureHeight;
    }
    else if (id == LUX_BUF_DEPTH || id == LUX_BUF_DEPTH_PREV)
    {
        rowBytes = (size_t)c->uniform.depthTextureWidth * 4;
        rows     = c->uniform.depthTextureHeight;
    }
    else
        return fail(LUX_ERR_INVALID_ARG, "row download is for atlases only");
    if (rowBegin < 0 || rowCount < 0 || rowBegin + rowCount > rows)
        return fail(LUX_ERR_INVALID_ARG, "rows [%d,%d) outside the atlas (%d rows)", rowBegin, rowBegin + rowCount, rows);
    // on the copy stream, as soon as the blend kernel that produced this atlas has finished
    const bool isIrr = (id == LUX_BUF_IRRADIANCE || id == LUX_BUF_IRRADIANCE_PREV);
    LUX_CUDA(cudaStreamWaitEvent(c->downStream, isIrr ? c->evIrrDone : c->evDepthDone, 0));
    if (!ownsRows(c, isIrr ? LUX_IRRADIANCE_OCT_SIZE : LUX_DEPTH_OCT_SIZE, rowBegin, rowCount))
        waitAllGather(c, c->downStream); // rows of other ranks arrive with the exchange
    LUX_CUDA(cudaMemcpyAsync(pinnedHost, (const char*)b->ptr + rowBegin * rowBytes, rowCount * rowBytes, cudaMemcpyDeviceToHost, c->downStream));
    LUX_CUDA(cudaEventRecord(c->evCopyDone, c->downStream));
    return LUX_OK;
}

int lux_ddgi_download_shard_async(LuxDDGIContext* c, LuxBufferId id, void* pinnedHost)
{
    CHECK_CTX(c);
    if (!pinnedHost)
        return fail(LUX_ERR_INVALID_ARG, "null host pointer");
    const bool isIrr = (id == LUX_BUF_IRRADIANCE || id == LUX_BUF_IRRADIANCE_PREV);
    if (!isIrr && id != LUX_BUF_DEPTH && id != LUX_BUF_DEPTH_PREV)
        return fail(LUX_ERR_INVALID_ARG, "shard download is for atlases only");
    DeviceBuffer* b = nullptr;
    int rc = bufferOf(c, id, &b);
    if (rc != LUX_OK)
        return rc;
    LuxDDGIState st{};
    shardLayout(c->uniform, c->rank, c->world, c->flags, &st);
    const size_t rowBytes = isIrr ? (size_t)c->uniform.irradianceTextureWidth * 8 : (size_t)c->uniform.depthTextureWidth * 4;
    const int    S = ((isIrr ? LUX_IRRADIANCE_OCT_SIZE : LUX_DEPTH_OCT_SIZE) + 2) * st.unitLayers, layers = st.probeCount / st.layerProbes; // rows per unit, units
    const int    rowBegin = isIrr ? st.irradianceRowBegin : st.depthRowBegin;
    LUX_CUDA(cudaStreamWaitEvent(c->downStream, isIrr ? c->evIrrDone : c->evDepthDone, 0));
    // `layers` units of S rows each, layerStride * S rows apart in the atlas, packed on the host: one strided copy
    LUX_CUDA(cudaMemcpy2DAsync(pinnedHost, (size_t)S * rowBytes, (const char*)b->ptr + (size_t)rowBegin * rowBytes, (size_t)st.layerStride * S * rowBytes,
                               (size_t)S * rowBytes, (size_t)layers, cudaMemcpyDeviceToHost, c->downStream));
    LUX_CUDA(cudaEventRecord(c->evCopyDone, c->downStream));
    return LUX_OK;
}

int lux_ddgi_download_fence(LuxDDGIContext* c, uint64_t* fence)
{
    CHECK_CTX(c);
    if (!fence)
        return fail(LUX_ERR_INVALID_ARG, "null fence");
    const uint64_t id = ++c->fenceSeq;
    LUX_CUDA(cudaEventRecord(c->fences[id % 8], c->downStream));
    *fence = id;
    return LUX_OK;
}

int lux_ddgi_wait_fence(LuxDDGIContext* c, uint64_t fence)
{
    CHECK_CTX(c);
    if (fence == 0 || fence > c->fenceSeq)
        return fail(LUX_ERR_INVALID_ARG, "unknown fence %llu", (unsigned long long)fence);
    // The slot may have been re-recorded by a later fence (8 slots).  That fence sits later on the same in-order stream, so waiting for it
    // implies this one: recorded is not the same as executed, hence never "return without waiting".
    LUX_CUDA(cudaEventSynchronize(c->fences[fence % 8]));
    return LUX_OK;
}

int lux_ddgi_restore(LuxDDGIContext* c, const void* irradiance, const void* depth, int32_t frames, int32_t pingPong)
{
    CHECK_CTX(c);
    if (!irradiance || !depth || frames < 0 || (pingPong != 0 && pingPong != 1))
        return fail(LUX_ERR_INVALID_ARG, "bad restore arguments");
    // "current" after `frames` frames is tex[pingPong] (the pair the next frame reads as prev)
    LUX_CUDA(cudaMemcpyAsync(c->irradiance[pingPong].ptr, irradiance, c->irradiance[pingPong].bytes, cudaMemcpyHostToDevice, c->stream));
    LUX_CUDA(cudaMemcpyAsync(c->depth[pingPong].ptr, depth, c->depth[pingPong].bytes, cudaMemcpyHostToDevice, c->stream));
    LUX_CUDA(cudaStreamSynchronize(c->stream));
    c->frames      = frames;
    c->pingPong    = pingPong;
    c->lastWritten = pingPong;
    return LUX_OK;
}

int lux_ddgi_shard_layout(const LuxDDGIUniform* u, int32_t rank, int32_t world, LuxDDGIState* out)
{
    return lux_ddgi_shard_layout_ex(u, rank, world, 0u, out);
}

int lux_ddgi_shard_layout_ex(const LuxDDGIUniform* u, int32_t rank, int32_t world, uint32_t flags, LuxDDGIState* out)
{
    if (!u || !out)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    if (world <= 0 || rank < 0 || rank >= world)
        return fail(LUX_ERR_INVALID_ARG, "rank %d outside world %d", rank, world);
    if (u->probeCounts[2] <= 0 || u->probeCounts[2] % world != 0)
        return fail(LUX_ERR_INVALID_ARG, "world %d must divide probeCounts.z %d (z-slab sharding)", world, u->probeCounts[2]);
    if (!shardFlagsValid(*u, world, flags))
        return fail(LUX_ERR_INVALID_ARG, "world %d x %d layers per interleaved block must divide probeCounts.z %d", world, LUX_DDGI_SHARD_BLOCK_LAYERS(flags), u->probeCounts[2]);
    *out = LuxDDGIState{};
    shardLayout(*u, rank, world, flags, out);
    return LUX_OK;
}

int lux_ddgi_get_state(LuxDDGIContext* c, LuxDDGIState* out)
{
    if (!c || !out)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    *out = LuxDDGIState{};
    shardLayout(c->uniform, c->rank, c->world, c->flags, out);
    out->frames         = c->frames;
    out->pingPong       = c->pingPong;
    out->kernelLaunches = c->launches;
    return LUX_OK;
}

// ---- consumer side: the step after the path (sample_probe::system, DDGIRenderer.cpp:467-503; SampleProbe.comp) ----
static int stageToDevice(LuxDDGIContext* c, const void* src, size_t bytes, LuxMemKind kind, void** dev, bool* owned)
{
    *owned = false;
    if (kind == LUX_MEM_DEVICE)
    {
        *dev = const_cast<void*>(src);
        return LUX_OK;
    }
    LUX_CUDA(cudaMalloc(dev, bytes ? bytes : 1));
    *owned = true;
    LUX_CUDA(cudaMemcpyAsync(*dev, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return LUX_OK;
}

int lux_ddgi_sample_irradiance(LuxDDGIContext* c, int32_t count, const float* P, const float* N, const float* Wo, float* out, LuxMemKind kind)
{
    CHECK_CTX(c);
    waitAllGather(c, c->stream);
    if (count < 0 || !P || !N || !Wo || !out)
        return fail(LUX_ERR_INVALID_ARG, "bad sample_irradiance arguments");
    if (c->frames == 0)
        return fail(LUX_ERR_NOT_READY, "no atlas has been written yet");
    const size_t bytes = (size_t)count * 3 * sizeof(float);
    void *dP, *dN, *dW, *dO;
    bool  oP, oN, oW, oO;
    int   rc;
    if ((rc = stageToDevice(c, P, bytes, kind, &dP, &oP)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, N, bytes, kind, &dN, &oN)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, Wo, bytes, kind, &dW, &oW)) != LUX_OK) return rc;
    if (kind == LUX_MEM_DEVICE) { dO = out; oO = false; }
    else { LUX_CUDA(cudaMalloc(&dO, bytes ? bytes : 1)); oO = true; }
    lux::launch_sample_irradiance(c->uniform, c->irradiance[c->lastWritten].ptr, c->depth[c->lastWritten].ptr, count, (const float*)dP,
                                  (const float*)dN, (const float*)dW, (float*)dO, c->stream);
    c->launches += count > 0 ? 1 : 0;
    LUX_CUDA(cudaGetLastError());
    if (oO)
    {
        LUX_CUDA(cudaMemcpyAsync(out, dO, bytes, cudaMemcpyDeviceToHost, c->stream));
        LUX_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(dO);
    }
    if (oP) cudaFree(dP);
    if (oN) cudaFree(dN);
    if (oW) cudaFree(dW);
    return LUX_OK;
}

int lux_ddgi_sample_probe(LuxDDGIContext* c, int32_t width, int32_t height, const float* depthD32F, const float* normalsRGBA32F,
                          const float cameraPosition[4], const float viewProjInv[16], float* outRGBA32F, LuxMemKind kind)
{
    CHECK_CTX(c);
    waitAllGather(c, c->stream);
    if (width <= 0 || height <= 0 || !depthD32F || !normalsRGBA32F || !cameraPosition || !viewProjInv || !outRGBA32F)
        return fail(LUX_ERR_INVALID_ARG, "bad sample_probe arguments");
    if (c->frames == 0)
        return fail(LUX_ERR_NOT_READY, "no atlas has been written yet");
    const size_t px = (size_t)width * height;
    void *dD, *dN, *dO;
    bool  oD, oN, oO;
    int   rc;
    if ((rc = stageToDevice(c, depthD32F, px * 4, kind, &dD, &oD)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, normalsRGBA32F, px * 16, kind, &dN, &oN)) != LUX_OK) return rc;
    if (kind == LUX_MEM_DEVICE) { dO = outRGBA32F; oO = false; }
    else { LUX_CUDA(cudaMalloc(&dO, px * 16)); oO = true; }
    lux::launch_sample_probe(c->uniform, c->irradiance[c->lastWritten].ptr, c->depth[c->lastWritten].ptr, width, height, (const float*)dD,
                             (const float*)dN, cameraPosition, viewProjInv, (float*)dO, c->stream);
    c->launches += 1;
    LUX_CUDA(cudaGetLastError());
    if (oO)
    {
        LUX_CUDA(cudaMemcpyAsync(outRGBA32F, dO, px * 16, cudaMemcpyDeviceToHost, c->stream));
        LUX_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(dO);
    }
    if (oD) cudaFree(dD);
    if (oN) cudaFree(dN);
    return LUX_OK;
}

int lux_ddgi_indirect_light(LuxDDGIContext* c, const void* baseLightRGBA16F, int32_t count, const uint32_t* texelIndex, const float* worldPos,
                            const float* normal, const float* albedo, const float* metallic, float intensity, const float cameraPos[3],
                            LuxMemKind kind)
{
    CHECK_CTX(c);
    waitAllGather(c, c->stream);
    if (!c->hasAtlas)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    if (c->frames == 0)
        return fail(LUX_ERR_NOT_READY, "no atlas has been written yet");
    if (count < 0 || !texelIndex || !worldPos || !normal || !albedo || !metallic || !cameraPos)
        return fail(LUX_ERR_INVALID_ARG, "bad indirect_light arguments");
    const size_t texels = (size_t)c->atlasData.resolution * c->atlasData.resolution;
    lightAcquire(c);
    if (c->light.borrowed)
    { // never write into a caller-owned buffer: take a private copy first
        void* own = nullptr;
        LUX_CUDA(cudaMalloc(&own, texels * 8));
        LUX_CUDA(cudaMemcpyAsync(own, c->light.ptr, texels * 8, cudaMemcpyDeviceToDevice, c->stream));
        c->light.ptr      = own;
        c->light.borrowed = false;
        c->light.bytes    = texels * 8;
    }
    void *dB = nullptr, *dT, *dP, *dN, *dA, *dM;
    bool  oB = false, oT, oP, oN, oA, oM;
    int   rc;
    if (baseLightRGBA16F && (rc = stageToDevice(c, baseLightRGBA16F, texels * 8, kind, &dB, &oB)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, texelIndex, (size_t)count * 4, kind, &dT, &oT)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, worldPos, (size_t)count * 12, kind, &dP, &oP)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, normal, (size_t)count * 12, kind, &dN, &oN)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, albedo, (size_t)count * 12, kind, &dA, &oA)) != LUX_OK) return rc;
    if ((rc = stageToDevice(c, metallic, (size_t)count * 4, kind, &dM, &oM)) != LUX_OK) return rc;
    if (dB) // light = base everywhere, then the listed texels receive base + indirect
        LUX_CUDA(cudaMemcpyAsync(c->light.ptr, dB, texels * 8, cudaMemcpyDeviceToDevice, c->stream));
    lux::launch_indirect_light(c->uniform, c->irradiance[c->lastWritten].ptr, c->depth[c->lastWritten].ptr, c->light.ptr, dB, count,
                               (const uint32_t*)dT, (const float*)dP, (const float*)dN, (const float*)dA, (const float*)dM, intensity, cameraPos,
                               c->stream);
    lightRelease(c);
    c->launches += count > 0 ? 1 : 0;
    LUX_CUDA(cudaGetLastError());
    if (oB || oT || oP || oN || oA || oM)
        LUX_CUDA(cudaStreamSynchronize(c->stream));
    if (oB) cudaFree(dB);
    if (oT) cudaFree(dT);
    if (oP) cudaFree(dP);
    if (oN) cudaFree(dN);
    if (oA) cudaFree(dA);
    if (oM) cudaFree(dM);
    return LUX_OK;
}

int lux_ddgi_surface_direct_light(LuxDDGIContext* c, const LuxLight* light, const float cameraPosBias[4], int32_t count, const uint32_t* texelIndex,
                                  const float* worldPos, const float* normal, const float* albedo, const float* metallicRoughness, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound");
    if (!c->hasAtlas)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    if (count < 0 || !light || !cameraPosBias || (count > 0 && (!texelIndex || !worldPos || !normal || !albedo || !metallicRoughness)))
        return fail(LUX_ERR_INVALID_ARG, "bad surface_direct_light arguments");
    if (light->type != LUX_LIGHT_DIRECTIONAL && light->type != LUX_LIGHT_SPOT && light->type != LUX_LIGHT_POINT)
        return fail(LUX_ERR_INVALID_ARG, "unknown light type");
    if (count == 0)
        return LUX_OK;
    const size_t texels = (size_t)c->atlasData.resolution * c->atlasData.resolution;
    lightAcquire(c);
    if (c->light.borrowed)
    { // never write into a caller-owned buffer: take a private copy first
        void* own = nullptr;
        LUX_CUDA(cudaMalloc(&own, texels * 8));
        LUX_CUDA(cudaMemcpyAsync(own, c->light.ptr, texels * 8, cudaMemcpyDeviceToDevice, c->stream));
        c->light.ptr      = own;
        c->light.borrowed = false;
        c->light.bytes    = texels * 8;
    }
    lux::TraceParams p{};
    p.sdf      = c->sdfData;
    p.tex      = (const uint16_t*)c->sdf.ptr;
    p.mip      = (const uint16_t*)c->mip.ptr;
    p.res      = (int)c->sdfData.resolution;
    p.mipRes   = p.res / 4;
    p.cascades = (int)c->sdfData.cascadesCount;
    p.texObj   = c->sdfTex;
    p.mipObj   = c->mipTex;
    void *dT = nullptr, *dP = nullptr, *dN = nullptr, *dA = nullptr, *dM = nullptr;
    bool  oT = false, oP = false, oN = false, oA = false, oM = false;
    int   rc;
    if ((rc = stageToDevice(c, texelIndex, (size_t)count * 4, kind, &dT, &oT)) == LUX_OK &&
        (rc = stageToDevice(c, worldPos, (size_t)count * 12, kind, &dP, &oP)) == LUX_OK &&
        (rc = stageToDevice(c, normal, (size_t)count * 12, kind, &dN, &oN)) == LUX_OK &&
        (rc = stageToDevice(c, albedo, (size_t)count * 12, kind, &dA, &oA)) == LUX_OK &&
        (rc = stageToDevice(c, metallicRoughness, (size_t)count * 8, kind, &dM, &oM)) == LUX_OK)
    {
        lux::launch_direct_light(p, c->sdfTex != 0, *light, cameraPosBias, c->light.ptr, count, (const uint32_t*)dT, (const float*)dP, (const float*)dN,
                                 (const float*)dA, (const float*)dM, c->stream);
        lightRelease(c);
        c->launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && (oT || oP || oN || oA || oM))
            e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess)
            rc = fail(LUX_ERR_CUDA, "surface_direct_light: %s", cudaGetErrorString(e));
    }
    if (oT) cudaFree(dT);
    if (oP) cudaFree(dP);
    if (oN) cudaFree(dN);
    if (oA) cudaFree(dA);
    if (oM) cudaFree(dM);
    return rc;
}

// SDF + surface cache + sky fields of TraceParams, for the kernels that trace rays outside the probe update
static void fillSceneParams(const LuxDDGIContext* c, lux::TraceParams& p)
{
    p.sdf      = c->sdfData;
    p.tex      = (const uint16_t*)c->sdf.ptr;
    p.mip      = (const uint16_t*)c->mip.ptr;
    p.res      = (int)c->sdfData.resolution;
    p.mipRes   = p.res / 4;
    p.cascades = (int)c->sdfData.cascadesCount;
    p.texObj   = c->sdfTex;
    p.mipObj   = c->mipTex;
    p.hasAtlas = c->hasAtlas ? 1 : 0;
    if (c->hasAtlas)
    {
        p.chunkSize     = c->atlasData.chunkSize;
        p.atlasRes      = c->atlasData.resolution;
        p.objectsCount  = c->atlasData.objectsCount;
        p.chunks        = (const uint32_t*)c->chunks.ptr;
        p.cull          = (const uint32_t*)c->cull.ptr;
        p.objects       = (const LuxObjectBuffer*)c->objects.ptr;
        p.objectInverse = (const float*)c->objectInverse.ptr;
        p.tiles         = (const LuxTileBuffer*)c->tiles.ptr;
        p.light         = (const uint2*)c->light.ptr;
        p.depth         = (const float*)c->atlasDepth.ptr;
    }
    p.skyFace = c->skyFace;
    p.sky     = (const uint2*)c->sky.ptr;
}

namespace {
// host or device inputs of one call staged on the context's stream; frees what it allocated when it goes out of scope
struct Staged
{
    LuxDDGIContext* c;
    LuxMemKind      kind;
    void*           owned[8];
    int             n = 0;
    Staged(LuxDDGIContext* ctx, LuxMemKind k) : c(ctx), kind(k) {}
    ~Staged()
    {
        for (int i = 0; i < n; i++)
            cudaFree(owned[i]);
    }
    int in(const void* src, size_t bytes, void** dev)
    {
        bool o  = false;
        int  rc = stageToDevice(c, src, bytes, kind, dev, &o);
        if (o) // also when the copy failed after the allocation succeeded
            owned[n++] = *dev;
        return rc;
    }
};
} // namespace

int lux_ddgi_sdf_reflection(LuxDDGIContext* c, const LuxReflectionPushConstants* push, int32_t width, int32_t height, const float* depthD32F,
                            const float* normalsRGBA32F, const float* pbrRGBA32F, const uint8_t* sobolRGBA8, const uint8_t* scramblingRankingRGBA8,
                            void* outRGBA16F, LuxMemKind kind)
{
    CHECK_CTX(c);
    waitAllGather(c, c->stream);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound");
    if (!push || width <= 0 || height <= 0 || !depthD32F || !normalsRGBA32F || !pbrRGBA32F || !sobolRGBA8 || !scramblingRankingRGBA8 || !outRGBA16F)
        return fail(LUX_ERR_INVALID_ARG, "bad sdf_reflection arguments");
    if (push->approximateWithDDGI == 1u && c->frames == 0)
        return fail(LUX_ERR_NOT_READY, "approximateWithDDGI: no atlas has been written yet");
    const size_t px = (size_t)width * height;
    Staged st(c, kind);
    void *dD, *dN, *dP, *dS, *dR, *dO;
    int   rc;
    if ((rc = st.in(depthD32F, px * 4, &dD)) != LUX_OK || (rc = st.in(normalsRGBA32F, px * 16, &dN)) != LUX_OK || (rc = st.in(pbrRGBA32F, px * 16, &dP)) != LUX_OK ||
        (rc = st.in(sobolRGBA8, 256 * 4, &dS)) != LUX_OK || (rc = st.in(scramblingRankingRGBA8, 128 * 128 * 4, &dR)) != LUX_OK ||
        (rc = st.in(outRGBA16F, px * 8, &dO)) != LUX_OK) // read-modify-write: depth == 1 pixels keep their contents
        return rc;
    lux::TraceParams p{};
    fillSceneParams(c, p);
    lightAcquire(c);
    lux::launch_sdf_reflection(p, c->sdfTex != 0, c->uniform, *push, c->irradiance[c->lastWritten].ptr, c->depth[c->lastWritten].ptr, width, height,
                               (const float*)dD, (const float*)dN, (const float*)dP, (const uint32_t*)dS, (const uint32_t*)dR, dO, c->stream);
    lightRelease(c);
    c->launches += 1;
    LUX_CUDA(cudaGetLastError());
    if (kind != LUX_MEM_DEVICE)
    {
        LUX_CUDA(cudaMemcpyAsync(outRGBA16F, dO, px * 8, cudaMemcpyDeviceToHost, c->stream));
        LUX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return LUX_OK;
}

int lux_ddgi_sdf_shadow(LuxDDGIContext* c, const LuxLight* light, const float viewProjInv[16], uint32_t numFrames, float shadowBias, int32_t width,
                        int32_t height, const float* depthD32F, const float* normalsRGBA32F, const uint8_t* sobolRGBA8,
                        const uint8_t* scramblingRankingRGBA8, uint32_t* outMaskR32UI, LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound");
    if (!light || !viewProjInv || width <= 0 || height <= 0 || !depthD32F || !normalsRGBA32F || !sobolRGBA8 || !scramblingRankingRGBA8 || !outMaskR32UI)
        return fail(LUX_ERR_INVALID_ARG, "bad sdf_shadow arguments");
    if (width % 8 != 0 || height % 4 != 0)
        return fail(LUX_ERR_INVALID_ARG, "sdf_shadow: width must be a multiple of 8 and height of 4 (one word per 8x4 workgroup), got %d x %d", width, height);
    const size_t px = (size_t)width * height, words = (size_t)(width / 8) * (height / 4);
    Staged st(c, kind);
    void *dD, *dN, *dS, *dR, *dO;
    int   rc;
    if ((rc = st.in(depthD32F, px * 4, &dD)) != LUX_OK || (rc = st.in(normalsRGBA32F, px * 16, &dN)) != LUX_OK || (rc = st.in(sobolRGBA8, 256 * 4, &dS)) != LUX_OK ||
        (rc = st.in(scramblingRankingRGBA8, 128 * 128 * 4, &dR)) != LUX_OK || (rc = st.in(outMaskR32UI, words * 4, &dO)) != LUX_OK)
        return rc;
    lux::TraceParams p{};
    fillSceneParams(c, p);
    lux::launch_sdf_shadow(p, c->sdfTex != 0, *light, viewProjInv, numFrames, shadowBias, width, height, (const float*)dD, (const float*)dN,
                           (const uint32_t*)dS, (const uint32_t*)dR, (uint32_t*)dO, c->stream);
    c->launches += 1;
    LUX_CUDA(cudaGetLastError());
    if (kind != LUX_MEM_DEVICE)
    {
        LUX_CUDA(cudaMemcpyAsync(outMaskR32UI, dO, words * 4, cudaMemcpyDeviceToHost, c->stream));
        LUX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return LUX_OK;
}

int lux_ddgi_get_surface_light_cache(LuxDDGIContext* c, void** devicePtr, size_t* bytes)
{
    CHECK_CTX(c);
    if (!c->hasAtlas || !devicePtr)
        return fail(LUX_ERR_NOT_READY, "no surface cache bound");
    if (c->lightPending) // a host upload may still be in flight on the copy stream: the pointer handed out is valid for immediate use on ANY stream
    {
        LUX_CUDA(cudaEventSynchronize(c->evLightReady));
        c->lightPending = false;
    }
    LUX_CUDA(cudaStreamSynchronize(c->stream)); // ... and so are this library's own writers (direct / indirect light)
    *devicePtr = c->light.ptr;
    if (bytes)
        *bytes = c->light.bytes;
    return LUX_OK;
}

int lux_ddgi_trace_global_sdf(LuxDDGIContext* c, int32_t count, const LuxGlobalSDFTrace* traces, float cascadeTraceStartBias, LuxGlobalSDFHit* hits,
                              LuxMemKind kind)
{
    CHECK_CTX(c);
    if (!c->hasSdf)
        return fail(LUX_ERR_NOT_READY, "no global SDF bound");
    if (count < 0 || (count > 0 && (!traces || !hits)))
        return fail(LUX_ERR_INVALID_ARG, "bad trace_global_sdf arguments");
    if (count == 0)
        return LUX_OK;
    lux::TraceParams p{};
    p.sdf      = c->sdfData;
    p.tex      = (const uint16_t*)c->sdf.ptr;
    p.mip      = (const uint16_t*)c->mip.ptr;
    p.res      = (int)c->sdfData.resolution;
    p.mipRes   = p.res / 4;
    p.cascades = (int)c->sdfData.cascadesCount;
    p.texObj   = c->sdfTex;
    p.mipObj   = c->mipTex;
    void *dT = nullptr, *dH = nullptr;
    bool  oT = false;
    int   rc = stageToDevice(c, traces, (size_t)count * sizeof(LuxGlobalSDFTrace), kind, &dT, &oT);
    if (rc != LUX_OK)
        return rc;
    if (kind == LUX_MEM_DEVICE)
        dH = hits;
    else
        LUX_CUDA(cudaMalloc(&dH, (size_t)count * sizeof(LuxGlobalSDFHit)));
    lux::launch_sdf_rays(p, c->sdfTex != 0, count, (const LuxGlobalSDFTrace*)dT, cascadeTraceStartBias, (LuxGlobalSDFHit*)dH, c->stream);
    c->launches += 1;
    LUX_CUDA(cudaGetLastError());
    if (kind != LUX_MEM_DEVICE)
    {
        cudaError_t e = cudaMemcpyAsync(hits, dH, (size_t)count * sizeof(LuxGlobalSDFHit), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
        cudaFree(dH);
        if (oT)
            cudaFree(dT);
        LUX_CUDA(e);
    }
    return LUX_OK;
}

int lux_ddgi_measure_l2_read_bandwidth(LuxDDGIContext* c, size_t bytes, int32_t repeats, float* gbPerSecond)
{
    CHECK_CTX(c);
    if (!gbPerSecond || repeats < 1)
        return fail(LUX_ERR_INVALID_ARG, "bad measure_l2_read_bandwidth arguments");
    if (bytes == 0)
        bytes = (size_t)64 << 20;
    bytes &= ~(size_t)16383;
    if (bytes == 0 || bytes > ((size_t)1 << 32))
        return fail(LUX_ERR_INVALID_ARG, "measure_l2_read_bandwidth: buffer size out of range");
    int sms = 0, dev = 0;
    LUX_CUDA(cudaGetDevice(&dev));
    LUX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8; // 8 resident blocks of 256 threads per SM
    void*     buf    = nullptr;
    LUX_CUDA(cudaMalloc(&buf, bytes + 16));
    cudaError_t e = cudaMemsetAsync(buf, 0, bytes + 16, c->stream);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    float best = 0.0f;
    for (int r = 0; e == cudaSuccess && r <= repeats; r++) // r = 0 warms the L2
    {
        cudaEventRecord(e0, c->stream);
        lux::launch_l2_sweep(buf, bytes, blocks, (uint32_t*)((char*)buf + bytes), c->stream);
        cudaEventRecord(e1, c->stream);
        e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess)
            e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess && r > 0 && ms > 0.0f)
            best = std::fmax(best, (float)((double)bytes * blocks / (ms * 1e-3) / 1e9));
    }
    c->launches += repeats + 1;
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf);
    if (e != cudaSuccess)
        return fail(LUX_ERR_CUDA, "measure_l2_read_bandwidth: %s", cudaGetErrorString(e));
    *gbPerSecond = best;
    return LUX_OK;
}

int lux_ddgi_get_stage_ms(LuxDDGIContext* c, LuxStageTimes* out)
{
    CHECK_CTX(c);
    if (!out)
        return fail(LUX_ERR_INVALID_ARG, "null argument");
    if (!c->timed)
        return fail(LUX_ERR_NOT_READY, "no timed update yet (create with LUX_DDGI_FLAG_STAGE_TIMERS and call lux_ddgi_update)");
    LUX_CUDA(cudaEventSynchronize(c->ev[4]));
    cudaEventElapsedTime(&out->setup_ms, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&out->trace_ms, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&out->blend_ms, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&out->border_ms, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&out->total_ms, c->ev[0], c->ev[4]);
    out->march_ms = out->shade_ms = 0.0f;
    if (!(c->flags & LUX_DDGI_FLAG_TRACE_SIMPLE))
    {
        cudaEventElapsedTime(&out->march_ms, c->ev[1], c->ev[5]);
        cudaEventElapsedTime(&out->shade_ms, c->ev[5], c->ev[2]);
    }
    return LUX_OK;
}

} // extern "C"
