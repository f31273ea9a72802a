// ddgi_kernels.h — launch interface between the C++ host (ddgi_engine.cpp) and the sm_100a kernels (ddgi_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/luxddgi.h"

namespace lux {

// Everything the trace kernel reads, passed as one __grid_constant__ parameter block.
struct TraceParams
{
    // probe volume (DDGIUniform)
    float start[3];
    float step[3];
    int   countX, countY;
    int   raysPerProbe;
    int   probeBegin, probeCount; // shard: probeCount probes from probeBegin on, see shard_probe()
    int   layerProbes, layerStride; // probes per z-layer; 1 = one z-slab, world = interleaved layers (LuxDDGIState)
    // global SDF
    LuxGlobalSDFData sdf;
    const uint16_t*  tex; // R16F [res][res][res*cascades]
    const uint16_t*  mip; // R16F [res/4][res/4][res/4*cascades]
    int              res, mipRes, cascades;
    cudaTextureObject_t texObj, mipObj; // optional layered-2D texture objects (LUX_DDGI_FLAG_SDF_TEXTURE)
    // surface cache
    int                    hasAtlas;
    float                  chunkSize;
    uint32_t               atlasRes, objectsCount;
    const uint32_t*        chunks;
    const uint32_t*        cull;
    const LuxObjectBuffer* objects;
    const float*           objectInverse; // [objects][16], inverse(object.transform) precomputed at upload
    const unsigned long long* chunkMasks; // [40^3][64] conservative per-sub-cell candidate masks, or null
    const LuxTileBuffer*   tiles;
    const float4*          tileZRow;      // [tiles] (m[2], m[6], m[10], m[14]) of each tile transform, or null (exact early reject in the shade)
    const uint2*           light; // RGBA16F
    const float*           depth; // D32F
    // sky
    int          skyFace;
    const uint2* sky; // [6][n][n] RGBA16F or null
    // rays
    const float4* dirs;     // [R] normalize(mat3(rot) * sphericalFibonacci(r, R))
    const float4* origins;  // [probeCount] probeLocation of this shard's probes
    float4*       records;  // wavefront trace: per ray (hitTime, hit u, v, w), indexed in MARCH ORDER (see MarchOrder)
    uint32_t*     meta;     // wavefront trace: cascade | kind << 2 | steps << 4
    // march order.  A chunk = 64 records = [j = 0..1][lane = 0..31], walked [direction cluster][probe unit] over spatially
    // tiled probe units; record index = chunk * 64 + j * 32 + lane.  Two shapes of a chunk:
    //   beam = 1: 2 x-adjacent probes (j) x one cluster of 32 angularly adjacent directions (lane): a warp holds rays that leave one point into
    //             a narrow cone and share the sectors of their first taps.  Used while the volume's pages fit the TLB (see ddgi_engine.cpp).
    //   beam = 0: 32 x-adjacent probes (lane) x 2 directions of a cluster (j): a warp holds parallel rays, whose taps stay in two z-slices = two
    //             2 MiB pages per gather.  Used for volumes beyond the TLB's reach (C5).
    // Tables are built on the host (init::marchOrder).
    const uint32_t* unitOrder; // [probeUnits] visiting order of the probe units (pairs or groups of 32), spatially tiled
    const uint32_t* unitIndex; // [probeUnits] its inverse
    const uint16_t* rayOrder;  // [raySlots]   ray id in each slot: runs of MARCH_CLUSTER_RAYS slots are angularly adjacent directions; >= raysPerProbe = padding
    const uint16_t* raySlot;   // [raySlots]   its inverse
    int             probeUnits, rayClusters; // rayClusters = raySlots / MARCH_CLUSTER_RAYS
    int             beam;
    // kernel-uniform values of the march, computed once on the host with the same IEEE operations the device would use: they are read straight
    // from the constant bank instead of occupying a register per lane (march_consts())
    struct MarchConsts
    {
        float traceMaxDistance, chunkSizeDistance, chunkMarginDistance2;
        float cascadesCountF, cascadesInv; // (float)cascadesCount and its reciprocal when the division is exact (power of two), else 0
        float mipW, mipH, texW, texH;      // texture extents as floats
        int   mipDm1, texDm1;              // depth - 1
        float cc0[3], cd0, m0, minv0, v0, vinv0; // cascade 0: centre, half extent, 2*cd, 1/(2*cd) or 0, voxel, 1/voxel or 0
    } mc;
    int             probeMajor;               // 1 = [probe unit][cluster] loop nest over ids as they come (LUX_DDGI_FLAG_MARCH_PROBE_MAJOR), 0 = [cluster][probe unit]
    uint2*        radiance; // [probeCount][R] RGBA16F
    uint2*        dirDist;  // [probeCount][R] RGBA16F
    uint16_t*     steps;    // optional [probeCount][R] march-step counts (debug / roofline counters)
    const float2* probeTaps; // [probeCount] (mip tap, full-resolution tap) at each probe position in cascade 0, or null (march_kernel.inc: step 0)
    uint32_t*     nonFinite; // [1] set when a non-finite fp16 value is written to the ray buffers (the blend then guards gated weights)
    // sorted shade (null sortedIdx = shade in ray order)
    float     invChunkSize;
    uint2*    sortTicket;   // [records] (bin, ticket within bin | cascade << 30) or (0xffffffff, -) for rays that need no shading; written by the march
    uint32_t* binCounts;    // [trace_sort_bins] hit histogram, turned into its exclusive prefix sum in place
    uint32_t* binBlockSums; // [trace_sort_blocks]
    uint32_t* hitCount;     // [1] number of hits = length of sortedIdx in use
    uint32_t* sortedIdx;    // [records] record indices in bin order
    unsigned long long* shadePools; // [256][1 + 32] per-SM segment tickets of shade_sorted_kernel (zeroed per launch)
    unsigned int*       shadeCounter; // [1] next group of segments
};

struct BlendParams
{
    int   probeBegin, probeCount;
    int   layerProbes, layerStride; // see shard_probe()
    int   raysPerProbe, raysPadded; // raysPadded = round_up(R, 32): weight rows beyond R are zero
    int   probesPerRow;             // X*Y
    int   irrWidth, depthWidth;     // atlas widths in texels
    float hysteresis, invGamma, maxDistance;
    int   firstFrame;
    int   fuseBorder;
    const uint2* radiance; // [probeCount][R]
    const uint2* dirDist;  // [probeCount][R]
    const float* wIrr;     // [raysPadded][64]   gated weights, probe independent
    const float* wDepth;   // [raysPadded][256]
    const float* scaleIrr; // [64]  1/(2*sum w) or 1
    const float* scaleDepth; // [256]
    const uint32_t* nzIrr;   // [raysPadded] bit g: texel group g (8 consecutive texels) has a non-zero weight for this ray
    const uint32_t* nzDepth; // [raysPadded]
    const uint32_t* nonFinite; // [1] (nullable) != 0: the ray buffers may hold fp16 Inf / NaN
    const uint2* prevIrr;  // RGBA16F atlas
    uint2*       outIrr;
    const uint32_t* prevDepth; // RG16F atlas
    uint32_t*       outDepth;
};

// Live-ray lists of the list blend (blend_lists.inc): per group of 2 x 2 texels the rays with a non-zero weight, in ray order, with their weights
struct BlendLists
{
    float4*   irrW;     // [16][cap]  weights of the group's four texels
    uint16_t* irrIdx;   // [16][cap]  byte offset of the ray's values within its staged phase
    uint32_t* irrOff;   // [16][irrPhases + 1]  list position at which each phase of rays begins
    float4*   depthW;   // [64][cap]
    uint16_t* depthIdx; // [64][cap]
    uint32_t* depthOff; // [64][depthPhases + 1]
    float*    irrMean;     // [16]  mean id of the group's live rays
    float*    depthMean;   // [64]
    uint8_t*  irrAssign;   // [8 warps][2]   groups owned by each warp of a block this frame
    uint8_t*  depthAssign; // [16 warps][4]
    int       cap, irrPhases, depthPhases;
};
size_t     blend_lists_bytes(int raysPerProbe, int raysPadded);
BlendLists blend_lists_layout(void* base, int raysPerProbe, int raysPadded);
int        launch_blend_lists(const float* wIrr, const float* wDepth, int raysPerProbe, const BlendLists& lists, cudaStream_t s); // returns the launch count
bool       blend_lists_preferred(int probeCount); // the list kernels pay off once every SM gets full 64-probe tiles
void       launch_blend_irradiance_lists(const BlendParams& p, const BlendLists& lists, cudaStream_t s);
void       launch_blend_depth_lists(const BlendParams& p, const BlendLists& lists, cudaStream_t s);

// per-frame setup: ray directions and the probe-independent blend weights
// rotation travels as a kernel argument; dirsHalf (nullable) = the same directions rounded to fp16, as the blend reads them
void launch_ray_dirs(const float* rot16Host, int raysPerProbe, float4* dirs, uint2* dirsHalf, cudaStream_t s);
int  launch_blend_weights(const uint2* dirDistRow0, int raysPerProbe, int raysPadded, float sharpness, float* wIrr, float* wDepth,
                          float* scaleIrr, float* scaleDepth, uint32_t* nzIrr, uint32_t* nzDepth, cudaStream_t s);
void launch_chunk_masks(const uint32_t* chunks, const uint32_t* cull, const LuxObjectBuffer* objects, const float* objectInverse,
                        uint32_t objectsCount, float chunkSize, float thrMax, unsigned long long* masks, cudaStream_t s);
void launch_object_inverse(const LuxObjectBuffer* objects, int count, float* inv, cudaStream_t s);
void launch_tile_zrow(const LuxTileBuffer* tiles, int count, float4* out, cudaStream_t s);

// variant 0: one thread per ray (simple); 1: wavefront (march + shade), explicit fp16 loads; 2: wavefront, layered-texture
// gathers.  Returns the number of kernels launched.
// `beforeShade` (nullable): event the stream waits on before the first kernel that reads the surface cache.
// `afterMarch` (nullable): recorded between the march and the shade kernel (stage timers).
int    launch_trace(const TraceParams& p, int variant, unsigned int* chunkCounter, cudaStream_t s, cudaEvent_t beforeShade,
                    cudaEvent_t afterMarch);
constexpr int MARCH_CLUSTER_RAYS = 32; // directions per cluster = lanes of a warp
size_t trace_record_count(int probeCount, int raysPerProbe, bool beam);
size_t trace_record_capacity(int probeCount, int raysPerProbe); // enough for either chunk shape
size_t trace_shade_pool_bytes();
size_t trace_sort_bins();   // bins of the sorted shade's counting sort (culling chunks x octants, padded to the scan's block size)
size_t trace_sort_blocks(); // scan blocks over those bins
void   launch_probe_origins(const TraceParams& p, cudaStream_t s);
void   launch_probe_taps(const TraceParams& p, bool useTextures, float2* taps, cudaStream_t s); // per-probe cache of the first step's taps
void launch_blend_irradiance(const BlendParams& p, cudaStream_t s);
// tensor-core blend (LUX_DDGI_FLAG_BLEND_TC): weights transposed / scaled / split into fp16 hi + lo [n][kPad], then the two GEMM kernels
int  blend_tc_kpad(int raysPerProbe);
void launch_blend_tc_weights(const float* wIrr, const float* wDepth, int rowsPadded, int kPad, uint16_t* irrHi, uint16_t* irrLo, uint16_t* depthHi,
                             uint16_t* depthLo, cudaStream_t s);
void launch_blend_irradiance_tc(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s);
bool launch_blend_depth_tc(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s); // false: not applicable, use the FP32 kernel
// tcgen05 / TMA form (blend_umma.inc); false: not applicable (odd ray count, no tensor-map encoder), use the mma.sync kernel
int  blend_umma_irr_kpad(int raysPerProbe);
void launch_blend_umma_irr_weights(const float* wIrr, int raysPerProbe, uint16_t* hi, uint16_t* lo, cudaStream_t s); // [192][kPad] each
bool launch_blend_irradiance_umma(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, cudaStream_t s);
bool launch_blend_depth_umma(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s); // hi / lo: the [256][kPad] matrices of blend_tc.inc
void launch_blend_depth(const BlendParams& p, cudaStream_t s);
void launch_border(uint2* irr, int irrWidth, uint32_t* depth, int depthWidth, int probesPerRow, int probeBegin, int probeCount, int layerProbes, int layerStride,
                   cudaStream_t s);


// consumer side
void launch_sample_irradiance(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, int count, const float* P, const float* N,
                              const float* Wo, float* out, cudaStream_t s);
void launch_sample_probe(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, int width, int height, const float* gDepth,
                         const float* gNormal, const float* cameraPosition, const float* viewProjInv, float* out, cudaStream_t s);
void launch_indirect_light(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, void* light, const void* base, int count,
                           const uint32_t* texel, const float* P, const float* N, const float* albedo, const float* metallic, float intensity,
                           const float* cameraPos, cudaStream_t s);

// tracyGlobalSDF for arbitrary rays; only the SDF fields of TraceParams are read
void launch_sdf_rays(const TraceParams& p, bool useTextures, int count, const LuxGlobalSDFTrace* traces, float cascadeTraceStartBias,
                     LuxGlobalSDFHit* hits, cudaStream_t s);

// SDFDeferredLight.frag for a list of surface-cache texels, additive into the RGBA16F light cache
void launch_direct_light(const TraceParams& p, bool useTextures, const LuxLight& l, const float* cameraPosBias, void* light, int count,
                         const uint32_t* texel, const float* P, const float* N, const float* albedo, const float* metallicRoughness, cudaStream_t s);

// SDFReflection.comp / SDFShadow.comp over a G-buffer; sobol / scr = RGBA8 texels as uint32
void launch_sdf_reflection(const TraceParams& p, bool useTextures, const LuxDDGIUniform& ddgi, const LuxReflectionPushConstants& push, const void* irr,
                           const void* dep, int width, int height, const float* gDepth, const float* gNormal, const float* gPbr, const uint32_t* sobol,
                           const uint32_t* scr, void* out, cudaStream_t s);
void launch_sdf_shadow(const TraceParams& p, bool useTextures, const LuxLight& light, const float* viewProjInv, uint32_t numFrames, float shadowBias,
                       int width, int height, const float* gDepth, const float* gNormal, const uint32_t* sobol, const uint32_t* scr, uint32_t* out,
                       cudaStream_t s);

// measurement aid: one sweep of `bytes` (multiple of 16 KiB) by each of `blocks` blocks through L2
void launch_l2_sweep(const void* buf, size_t bytes, int blocks, uint32_t* sink, cudaStream_t s);

// ---- global SDF build ----
struct SdfMeshRecord // device copy of LuxMeshSDF without the host pointers
{
    float aabbMin[3], aabbMax[3];
    float localToUVWMul[3], localToUVWAdd[3];
    float maxDistance;
    float worldMatrix[16];
};
struct SdfMeshLevel { const uint16_t* data; int w, h, d; }; // the mip level a cascade samples
struct SdfChunkDispatch                                      // RasterizeConsts (GlobalDistanceField.cpp:161-166) + which pipeline
{
    int32_t  coord[3]; // first voxel of the chunk
    int32_t  count;
    uint32_t models[LUX_SDF_RASTERIZE_MODEL_MAX_COUNT];
    int32_t  read;     // 1 = READ_DISTANCE variant (additive layer)
};
struct SdfRasterizeParams // ModelsRasterizeData (SDFRasterizeModel.glsl:16-24) + bindings
{
    float mul[3], add[3];
    float maxDistance;
    int   res, cascadeIndex, texWidth;
    const LuxObjectRasterizeData* objects;
    const SdfMeshLevel*           levels;
    const SdfChunkDispatch*       dispatches;
    uint16_t*                     sdf;
};
void launch_sdf_object_data(const SdfMeshRecord* meshes, int count, int cascadeLevel, LuxObjectRasterizeData* out, cudaStream_t s);
void launch_sdf_rasterize(const SdfRasterizeParams& p, int dispatchCount, cudaStream_t s);
void launch_sdf_fill(uint16_t* p, size_t n, uint16_t value, cudaStream_t s);
void launch_sdf_mip_pass(const uint16_t* src, int srcWidth, int srcHeight, uint16_t* dst, int dstWidth, int dstHeight, int outRes, int globalSDFResolution,
                         int mipmapCoordScale, int cascadeTexOffsetX, int cascadeMipMapOffsetX, float maxDistance, cudaStream_t s);

// ---- surface-cache culling: SDFCulling.comp with a deterministic (ascending chunk address) list layout ----
// sizesPadded: 65 536 words of scratch, blockSums: 16 words, totalWords: 1 word.  Returns through `chunks` (64 000 words) and `cull`.
void launch_surface_cull(const LuxObjectBuffer* objects, uint32_t objectsCount, float chunkSize, uint32_t capacity, uint32_t* sizesPadded,
                         uint32_t* blockSums, uint32_t* totalWords, uint32_t* chunks, uint32_t* cull, uint32_t cullWords, cudaStream_t s);

} // namespace lux
