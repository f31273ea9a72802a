/*
 * luxddgi.h — C ABI of the B200-native DDGI probe-update engine (libluxddgi.so).
 *
 * This is the drop-in boundary for the SDF-traced branch of the Maple renderer's DDGI pass
 * (flwmxd/LuxGI).  The reference has no C ABI; its "operator API" for this path is the three IoC
 * systems registered by ddgi::registerDDGI (Engine/DDGI/DDGIRenderer.cpp:1016-1037):
 *
 *     trace_rays::system     DDGIRenderer.cpp:215-331   -> lux_ddgi_trace_rays
 *     probe_update::system   DDGIRenderer.cpp:345-414   -> lux_ddgi_probe_update
 *     border_update::system  DDGIRenderer.cpp:416-465   -> lux_ddgi_border_update
 *     end_frame::system      DDGIRenderer.cpp:333-343   -> lux_ddgi_end_frame
 *     ddgi::on_game_start    DDGIRenderer.cpp:563-690   -> lux_ddgi_uniform_from_volume + lux_ddgi_create
 *     ddgi::on_game_end      DDGIRenderer.cpp:692-714   -> lux_ddgi_destroy
 *
 * Every struct below is byte-identical to the GLSL block / host twin it cites, so a maintainer can pass the
 * engine's own uniform blocks and SSBO contents through unchanged.
 *
 * Plain C: pointers and sizes only.  All functions return LUX_OK (0) or a negative LuxStatus; none of them
 * aborts.  lux_ddgi_last_error() returns a thread-local description of the last failure.
 * A context is bound to one CUDA device and one stream; calls on one context are not thread-safe
 * (the reference runs its systems sequentially on the render thread, IoC/SystemBuilder.h:202-208).
 */
#ifndef LUXDDGI_H
#define LUXDDGI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUXDDGI_VERSION 0x00010000

#if defined(__GNUC__)
#define LUX_API __attribute__((visibility("default")))
#else
#define LUX_API
#endif

/* ---------------------------------------------------------------------------------------------------------
 * Status codes
 * ------------------------------------------------------------------------------------------------------ */
typedef enum LuxStatus {
    LUX_OK                 = 0,
    LUX_ERR_INVALID_ARG    = -1, /* null pointer, bad size, inconsistent uniform                           */
    LUX_ERR_NOT_READY      = -2, /* stage called before its inputs were set                                */
    LUX_ERR_CUDA           = -3, /* a CUDA runtime call failed; text in lux_ddgi_last_error()              */
    LUX_ERR_NO_DEVICE      = -4, /* no usable sm_100 device: the engine has NO CPU fallback, it fails here */
    LUX_ERR_OUT_OF_MEMORY  = -5,
    LUX_ERR_UNSUPPORTED    = -6
} LuxStatus;

/* ---------------------------------------------------------------------------------------------------------
 * Constants of the pass (Engine/DDGI/DDGIRenderer.h:17-18, Shaders/SDF/SDFCommon.glsl:8-10,
 * Shaders/SDF/AtlasCommon.glsl:5-6, Shaders/DDGI/ProbeUpdate.glsl:7)
 * ------------------------------------------------------------------------------------------------------ */
#define LUX_IRRADIANCE_OCT_SIZE                 8      /* interior texels per probe side, irradiance      */
#define LUX_DEPTH_OCT_SIZE                      16     /* interior texels per probe side, depth           */
#define LUX_GLOBAL_SDF_WORLD_SIZE               60000.0f
#define LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE     32
#define LUX_GLOBAL_SDF_RASTERIZE_CHUNK_MARGIN   4
#define LUX_GLOBAL_SDF_MAX_STEPS                250
#define LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION     40
#define LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD 0.05f
#define LUX_MAX_CASCADES                        4

/* ---------------------------------------------------------------------------------------------------------
 * a1. DDGIUniform — probe-volume parameter block, `scalar` layout, 96 bytes.
 *     Shaders/DDGI/DDGICommon.glsl:11-31; host twin Engine/DDGI/DDGIRenderer.h:22-42.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxDDGIUniform {
    float   startPosition[4];          /* @0   world position of probe (0,0,0); w unused                  */
    float   step[4];                   /* @16  probe spacing per axis; w unused                           */
    int32_t probeCounts[4];            /* @32  X, Y, Z probes; w unused                                   */
    float   maxDistance;               /* @48  depth clamp = 1.5 * probeDistance                          */
    float   sharpness;                 /* @52  depth weight exponent (host twin: depthSharpness)          */
    float   hysteresis;                /* @56                                                             */
    float   normalBias;                /* @60  consumer only                                              */
    float   ddgiGamma;                 /* @64                                                             */
    int32_t irradianceProbeSideLength; /* @68  must be 8                                                  */
    int32_t irradianceTextureWidth;    /* @72  10*X*Y + 2                                                 */
    int32_t irradianceTextureHeight;   /* @76  10*Z + 2                                                   */
    int32_t depthProbeSideLength;      /* @80  must be 16                                                 */
    int32_t depthTextureWidth;         /* @84  18*X*Y + 2                                                 */
    int32_t depthTextureHeight;        /* @88  18*Z + 2                                                   */
    int32_t raysPerProbe;              /* @92                                                             */
} LuxDDGIUniform;

/* IrradianceVolume — the serialised user-facing component (Engine/DDGI/DDGIRenderer.h:44-66). */
typedef struct LuxIrradianceVolume {
    float   probeDistance;   /* 1.5  */
    int32_t infiniteBounce;  /* 1    */
    int32_t raysPerProbe;    /* 256  */
    float   hysteresis;      /* 0.98 */
    float   intensity;       /* 1.0  */
    float   normalBias;      /* 0.1  */
    float   depthSharpness;  /* 50   */
    float   ddgiGamma;       /* 5    */
} LuxIrradianceVolume;

/* ---------------------------------------------------------------------------------------------------------
 * a2. Trace push constants, 80 bytes (Shaders/DDGI/GISDFRays.comp:53-60; host DDGIRenderer.cpp:111-118).
 *     Only the upper-left 3x3 of randomOrientation (column-major) is used by the SDF path.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxTracePushConstants {
    float    randomOrientation[16]; /* column-major mat4: element (row r, col c) = m[c*4 + r]              */
    uint32_t numFrames;             /* declared, unused by GISDFRays.comp                                  */
    uint32_t infiniteBounces;       /* declared, unused by GISDFRays.comp                                  */
    int32_t  numLights;             /* declared, unused by GISDFRays.comp                                  */
    float    intensity;             /* declared, unused by GISDFRays.comp                                  */
} LuxTracePushConstants;

/* ---------------------------------------------------------------------------------------------------------
 * a5. GlobalSDFData, std140, 96 bytes (Shaders/SDF/GlobalSDFData.glsl:4-12; host GlobalDistanceField.h:19-27).
 *     SDF texture: R16F, [z][y][x] with x in [0, resolution*cascadesCount) — cascade c occupies
 *     x in [c*res, (c+1)*res); value = world distance / (2*cascadePosDistance[c].w) clamped to [-1,1].
 *     Mip texture: same with resolution/4 per axis.  Both sampled trilinear, clamp-to-edge
 *     (GlobalDistanceField.cpp:615,621).
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxGlobalSDFData {
    float    cascadePosDistance[LUX_MAX_CASCADES][4]; /* @0   centre.xyz, half-extent.w                  */
    float    cascadeVoxelSize[4];                     /* @64                                              */
    uint32_t cascadesCount;                           /* @80                                              */
    float    resolution;                              /* @84  voxels per cascade axis (float, as in GLSL) */
    float    nearPlane;                               /* @88  unused on this path                         */
    float    farPlane;                                /* @92  unused on this path                         */
} LuxGlobalSDFData;

/* ---------------------------------------------------------------------------------------------------------
 * f3. Global SDF build: the step BEFORE the path.  Mesh distance fields are min-merged into the
 *     cascade volume chunk by chunk (Shaders/SDF/SDFRasterizeModel.glsl:42-63 + SDFCommon.glsl:18-62, host
 *     Engine/DDGI/GlobalDistanceField.cpp:460-848), then the quarter-resolution min-mip is built and flooded
 *     (Shaders/SDF/GlobalSDFMipmap.comp:32-68, GlobalDistanceField.cpp:537-573,825-841).
 * ------------------------------------------------------------------------------------------------------ */
#define LUX_SDF_RASTERIZE_MODEL_MAX_COUNT 28 /* GlobalDistanceField.cpp:84, SDFRasterizeModel.glsl:8 */
#define LUX_SDF_RASTERIZE_CHUNK_SIZE      32 /* GlobalDistanceField.cpp:88                            */
#define LUX_SDF_RASTERIZE_CHUNK_MARGIN     4 /* GlobalDistanceField.cpp:91                            */
#define LUX_SDF_MESH_MAX_MIPS              3 /* SDFBaker.cpp:158                                      */

/* ObjectRasterizeData, std430, 176 bytes (Shaders/SDF/ObjectRasterizeData.glsl:7-17; host GlobalDistanceField.cpp:101-111) */
typedef struct LuxObjectRasterizeData {
    float worldToVolume[16];          /* @0   column-major */
    float volumeToWorld[16];          /* @64               */
    float volumeToUVWMul[3];          /* @128              */
    float mipOffset;                  /* @140              */
    float volumeToUVWAdd[3];          /* @144              */
    float decodeMul;                  /* @156              */
    float volumeLocalBoundsExtent[3]; /* @160              */
    float decodeAdd;                  /* @172              */
} LuxObjectRasterizeData;

/* One mesh distance field = component::MeshDistanceField + the entity's world matrix (MeshDistanceField.h:15-26).
 * Volume: R16F [z][y][x], value (d/maxDistance + 1)/2, mip m has size max(size >> m, 1), box-filtered (SDFBaker.cpp:96-204).
 * Sampled trilinear inside ONE mip level with REPEAT addressing (Texture3D default wrap, RHI/Definitions.h:152-174). */
typedef struct LuxMeshSDF {
    const void* mips[LUX_SDF_MESH_MAX_MIPS];
    uint32_t    size[3];
    int32_t     mipCount;
    float       aabbMin[3], aabbMax[3];
    float       localToUVWMul[3], localToUVWAdd[3];
    float       maxDistance;
    float       worldMatrix[16]; /* column-major */
} LuxMeshSDF;

/* ---------------------------------------------------------------------------------------------------------
 * a5b. GlobalSDFTrace / GlobalSDFHit (Shaders/SDF/GlobalSDFTrace.glsl:4-12, GlobalSDFHit.glsl:4-11): the argument and result of
 *      tracyGlobalSDF (SDFCommon.glsl:98-194), for its other users (row f4: SDFShadow.comp:140-148, SDFReflection.comp,
 *      SDFDeferredLight.frag:101-109).  Plain C layout, 40 / 28 bytes.  minDistance is carried but, as in the reference, never read.
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxGlobalSDFTrace {
    float    worldPosition[3];
    float    minDistance;
    float    worldDirection[3];
    float    maxDistance;
    float    stepScale;
    uint32_t needsHitNormal;
} LuxGlobalSDFTrace;

typedef struct LuxGlobalSDFHit {
    float    hitNormal[3]; /* (0,0,0) unless needsHitNormal and hit */
    float    hitTime;      /* < 0: miss (isHit, SDFCommon.glsl:73-76) */
    uint32_t hitCascade;
    uint32_t stepsCount;
    float    hitSDF;
} LuxGlobalSDFHit;

/* ---------------------------------------------------------------------------------------------------------
 * Light, 64 bytes (Shaders/Common/Light.glsl:19-29; host twin Scene/Component/Light.h:21-33 LightData), as SDFDeferredLight.frag's UniformBufferObject carries it (row f4).
 * type: 0 directional, 1 spot, 2 point (Light.glsl:13-17).  direction.w is the soft-shadow radius (unused by the surface-cache pass).
 * ------------------------------------------------------------------------------------------------------ */
#define LUX_LIGHT_DIRECTIONAL 0.0f
#define LUX_LIGHT_SPOT        1.0f
#define LUX_LIGHT_POINT       2.0f
typedef struct LuxLight {
    float color[4];
    float position[4];
    float direction[4];
    float intensity;
    float radius;
    float type;
    float angle;
} LuxLight;

/* ---------------------------------------------------------------------------------------------------------
 * a6. Surface-cache records (Shaders/SDF/AtlasCommon.glsl:8-32; host GlobalSurfaceAtlas.cpp:59-73,
 *     SurfaceAtlasTile.h:117-125).
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxGlobalSurfaceAtlasData { /* 32 bytes */
    float    cameraPos[3];           /* @0  unused by sampling                                            */
    float    chunkSize;              /* @12 world size of one of the 40^3 culling chunks                  */
    uint32_t culledObjectsCapacity;  /* @16                                                               */
    uint32_t resolution;             /* @20 atlas side in texels                                          */
    uint32_t objectsCount;           /* @24                                                               */
    uint32_t padding;                /* @28                                                               */
} LuxGlobalSurfaceAtlasData;

typedef struct LuxObjectBuffer { /* std430, 128 bytes */
    float    objectBounds[4];  /* @0   bounding sphere centre.xyz, radius.w                               */
    uint32_t tileOffset[6];    /* @16  index into tiles[]; 0 = no tile                                    */
    int32_t  padding[2];       /* @40                                                                     */
    float    transform[16];    /* @48  column-major OBB local->world                                      */
    float    extends[4];       /* @112 OBB half extents xyz, 1                                            */
} LuxObjectBuffer;

typedef struct LuxTileBuffer { /* std430, 96 bytes */
    float extends[4];      /* @0   (tile x, tile y, width-1, height-1) / atlas resolution                 */
    float transform[16];   /* @16  column-major tile view rotation (translation zeroed)                   */
    float objectBounds[4]; /* @80  view-space bounds size xyz, 0                                          */
} LuxTileBuffer;

/* ---------------------------------------------------------------------------------------------------------
 * Engine-side types
 * ------------------------------------------------------------------------------------------------------ */
typedef struct LuxDDGIContext LuxDDGIContext; /* opaque */

typedef enum LuxMemKind {
    LUX_MEM_HOST   = 0, /* pointer is host memory; the engine copies it to the device                      */
    LUX_MEM_DEVICE = 1  /* pointer is device memory on the context's device; the engine BORROWS it         */
} LuxMemKind;

enum {
    LUX_DDGI_FLAG_NONE          = 0,
    LUX_DDGI_FLAG_STAGE_TIMERS  = 1u << 0, /* record CUDA events around every stage (lux_ddgi_get_stage_ms)  */
    LUX_DDGI_FLAG_UNFUSED_BORDER= 1u << 1, /* probe_update writes interiors only; border_update does borders */
    LUX_DDGI_FLAG_SDF_TEXTURE   = 1u << 2, /* require the layered-texture gather (tld4) SDF path (the default) */
    LUX_DDGI_FLAG_SDF_LOADS     = 1u << 3, /* read the SDF with explicit fp16 global loads instead             */
    LUX_DDGI_FLAG_TRACE_SIMPLE  = 1u << 4, /* one-thread-per-ray trace kernel (no wavefront scheduling), for A/B */
    LUX_DDGI_FLAG_NO_PREFILTER  = 1u << 5, /* walk the full per-chunk object lists (no sub-cell candidate masks), for A/B */
    LUX_DDGI_FLAG_SHADE_UNSORTED= 1u << 6, /* shade hits in ray order instead of culling-chunk order (no counting sort), for A/B */
    LUX_DDGI_FLAG_NO_PIPELINE   = 1u << 7, /* lux_ddgi_update computes the blend weights on the context's stream instead of a second one, for A/B */
    LUX_DDGI_FLAG_MARCH_ROWS    = 1u << 9, /* force row chunks in the march (a warp = 32 x-adjacent probes x one direction); default: by volume size */
    LUX_DDGI_FLAG_MARCH_BEAMS   = 1u << 10,/* force beam chunks (a warp = one probe x 32 angularly adjacent directions)                        */
    LUX_DDGI_FLAG_BLEND_TC      = 1u << 11,/* opt-in: the blend as a tensor-core GEMM (fp16 hi / lo split operands, fp32 accumulation).  NOT the reference's
                                            * summation order: held to the north-star tolerance (1e-3 relative / 1e-4 absolute), not bit-exact         */
    LUX_DDGI_FLAG_BLEND_LISTS   = 1u << 12,/* force the list form of the FP32 blend (per texel group, the frame's rays with a non-zero weight; bit-identical to
                                            * the tiled form).  Default: by probe count (>= 148 x 64 probes in the shard)                               */
    LUX_DDGI_FLAG_BLEND_TILES   = 1u << 13,/* force the tiled form (32-ray chunks, zero skipping per 8-texel group), for A/B                            */
    LUX_DDGI_FLAG_BLEND_TC_MMA_SYNC = 1u << 14,/* with BLEND_TC: the mma.sync kernels instead of the tcgen05 / TMA ones, for A/B                   */
    LUX_DDGI_FLAG_SHARD_INTERLEAVED = 1u << 15,/* multi-GPU: the probe z-layers are dealt out to the ranks in blocks of B = LUX_DDGI_SHARD_BLOCK_LAYERS(flags) layers - rank g
                                            * owns the blocks g, g + world, g + 2 world, ... - instead of one contiguous z-slab per rank.  The cost of a probe layer
                                            * depends on its height in the scene (layers that look at open sky or ground march shorter rays), so slabs are unevenly
                                            * loaded: C5 at 8 GPUs 13.2 - 15.7 ms of march per rank.  The exchange stays contiguous: round k all-gathers the blocks
                                            * k world .. k world + world - 1, rank g contributing the g-th.  Results are identical.  B = 1 balances perfectly (16.0 -
                                            * 16.2 ms) but a shard's probes are then 8 x sparser in space and reuse less of the SDF they pull through L2: 18.5 ms per
                                            * update against 17.7 ms with slabs; larger blocks keep the locality                                     */
    LUX_DDGI_FLAG_SHARD_BLOCK_SHIFT = 16,      /* bits 16..19: log2 of the layers per interleaved block (0 = single layers)                                            */
    LUX_DDGI_FLAG_MARCH_PROBE_MAJOR = 1u << 8 /* wavefront march in the round-1 work order (probe groups outermost, ray ids as they come) instead of direction
                                            * clusters outermost over spatially tiled probe groups; same results, for A/B of the DRAM traffic */
};

#define LUX_DDGI_SHARD_BLOCK_LAYERS(flags) (1 << (((flags) >> 16) & 0xF))
#define LUX_DDGI_FLAG_SHARD_BLOCKS(log2Layers) (LUX_DDGI_FLAG_SHARD_INTERLEAVED | ((uint32_t)(log2Layers) << 16))

typedef struct LuxDDGICreateInfo {
    int32_t  device;  /* CUDA device ordinal                                                              */
    int32_t  rank;    /* z-slab shard owned by this context, 0 <= rank < world                            */
    int32_t  world;   /* number of z-slab shards (1 = whole volume); must divide probeCounts.z            */
    uint32_t flags;   /* LUX_DDGI_FLAG_*                                                                  */
    void*    stream;  /* cudaStream_t to run on, or NULL for an engine-owned stream                       */
} LuxDDGICreateInfo;

typedef enum LuxBufferId {
    LUX_BUF_RADIANCE            = 0, /* RGBA16F [probe_count(shard)][raysPerProbe]: rgb, 0                */
    LUX_BUF_DIRECTION_DISTANCE  = 1, /* RGBA16F [probe_count(shard)][raysPerProbe]: dir.xyz, hit distance */
    LUX_BUF_IRRADIANCE          = 2, /* RGBA16F [10Z+2][10XY+2], the atlas most recently written          */
    LUX_BUF_DEPTH               = 3, /* RG16F   [18Z+2][18XY+2], the atlas most recently written          */
    LUX_BUF_IRRADIANCE_PREV     = 4, /* the other half of the ping-pong pair                              */
    LUX_BUF_DEPTH_PREV          = 5,
    LUX_BUF_GLOBAL_SDF          = 6, /* R16F [res][res][res*cascades], the bound / built global SDF            */
    LUX_BUF_GLOBAL_SDF_MIP      = 7  /* R16F [res/4][res/4][res/4*cascades]                                    */
} LuxBufferId;

typedef struct LuxDDGIState {
    int32_t  frames;          /* frames completed (DDGIPipelineInternal::frames)                          */
    int32_t  pingPong;        /* DDGIPipelineInternal::pingPong                                           */
    int32_t  probeBegin;      /* first probe id of this shard                                             */
    int32_t  probeCount;      /* probes in this shard                                                     */
    int32_t  irradianceRowBegin, irradianceRowCount; /* atlas rows owned by this shard (all-gather unit)  */
    int32_t  depthRowBegin, depthRowCount;
    uint64_t kernelLaunches;  /* kernels launched by this context since creation                          */
    int32_t  layerProbes;     /* probes per interleave unit: B z-layers of X * Y probes (B = 1 for a z-slab)   */
    int32_t  layerStride;     /* 1: the shard is one z-slab.  world (LUX_DDGI_FLAG_SHARD_INTERLEAVED): its k-th unit is unit rank + k * world, i.e.
                               * shard-local probe l is probe probeBegin + (l / layerProbes) * layerStride * layerProbes + l % layerProbes and the
                               * atlas rows of its k-th unit start at *RowBegin + k * layerStride * unitLayers * (10 | 18); *RowCount = all own rows */
    int32_t  unitLayers;      /* B: z-layers per interleave unit                                             */
} LuxDDGIState;

typedef struct LuxStageTimes { /* milliseconds of the last lux_ddgi_update, needs FLAG_STAGE_TIMERS */
    float setup_ms, trace_ms, blend_ms, border_ms, total_ms;
    float march_ms, shade_ms; /* the two kernels of the wavefront trace (0 for the simple kernel) */
} LuxStageTimes;

/* ---------------------------------------------------------------------------------------------------------
 * Entry points
 * ------------------------------------------------------------------------------------------------------ */
LUX_API uint32_t    lux_ddgi_version(void);
LUX_API const char* lux_ddgi_last_error(void);

/* Grid derivation of ddgi::on_game_start (DDGIRenderer.cpp:663-682) and delegates::uniformChanged (:975-1013):
 * probeCounts = ivec3(sceneLength / probeDistance) + 2, start = aabb.min, step = probeDistance,
 * maxDistance = 1.5 * probeDistance, atlas sizes.  Unlike the reference (SURVEY finding 8) raysPerProbe IS copied. */
LUX_API int lux_ddgi_uniform_from_volume(const LuxIrradianceVolume* volume, const float aabbMin[3], const float aabbMax[3],
                                 LuxDDGIUniform* out);
/* Atlas sizing of init::initializeProbeGrid (DDGIRenderer.cpp:181-191): fills side lengths + texture sizes. */
LUX_API int lux_ddgi_uniform_finalize(LuxDDGIUniform* uniform);

/* Allocates ray buffers and the 2x2 ping-pong atlases (zero-filled), DDGIRenderer.cpp:163-211. */
LUX_API int lux_ddgi_create(const LuxDDGIUniform* uniform, const LuxDDGICreateInfo* info, LuxDDGIContext** out);
LUX_API int lux_ddgi_destroy(LuxDDGIContext* ctx);

/* Runtime-editable fields only (hysteresis, sharpness, gamma, maxDistance, normalBias, start, step);
 * probeCounts / raysPerProbe changes need a new context (the reference rebuilds its textures too, :505-561). */
LUX_API int lux_ddgi_set_uniform(LuxDDGIContext* ctx, const LuxDDGIUniform* uniform);

/* uGlobalSDF / uGlobalMipSDF + UniformBufferObject.sdfData (DDGIRenderer.cpp:304-305,314). fp16 texels. */
LUX_API int lux_ddgi_set_global_sdf(LuxDDGIContext* ctx, const LuxGlobalSDFData* data,
                            const void* sdfR16F, const void* mipR16F, LuxMemKind kind);

/* f4: direct lighting of surface-cache texels = Shaders/SDF/SDFDeferredLight.frag:44-129 (fetchLight, shadow ray through the global SDF with
 * start bias 2, BRDF of Raytraced/BRDF.glsl:65-83), blended ADDITIVELY into the RGBA16F light cache as the reference's pipeline does
 * (GlobalSurfaceAtlas.cpp:950-972: BlendMode::Add, alpha += 1).  One call = one light over the listed atlas texels; the per-texel arrays are
 * what SDFDeferredColor.frag captured there (world position, decoded normal, albedo, (metallic, roughness)).  cameraPos[3] = shadowBias.
 * Every texelIndex must be < resolution^2 of the bound surface cache (not checked on the device). */
LUX_API int lux_ddgi_surface_direct_light(LuxDDGIContext* ctx, const LuxLight* light, const float cameraPosBias[4], int32_t count,
                                          const uint32_t* texelIndex, const float* worldPos, const float* normal, const float* albedo,
                                          const float* metallicRoughness, LuxMemKind kind);

/* f4: tracyGlobalSDF for arbitrary rays through the bound global SDF (shadow / reflection / surface-cache light rays).
 * cascadeTraceStartBias = the shader call's last argument (0 in GISDFRays / SDFShadow, 2 in SDFDeferredLight). */
LUX_API int lux_ddgi_trace_global_sdf(LuxDDGIContext* ctx, int32_t count, const LuxGlobalSDFTrace* traces, float cascadeTraceStartBias,
                                      LuxGlobalSDFHit* hits, LuxMemKind kind);

/* f3: builds the global SDF and its mip ON DEVICE from mesh distance fields and binds them like lux_ddgi_set_global_sdf.
 * `data` gives the cascades (centre, half extent, voxel size = 2*extent/resolution, resolution, count); objects whose
 * bounding sphere misses a cascade or is smaller than minObjectRadius are skipped (GlobalDistanceField.cpp:692-701).
 * One-shot build (the reference's first frame: every chunk rasterized; its static-chunk caching across frames is not modelled).
 * Mesh volumes are HOST pointers (they come from .sdf files). */
LUX_API int lux_ddgi_build_global_sdf(LuxDDGIContext* ctx, const LuxGlobalSDFData* data, const LuxMeshSDF* meshes, int32_t meshCount,
                                      float minObjectRadius);
/* Rebuilds only the mip of the bound SDF (GlobalSDFMipmap.comp: one 4x min-downsample + 4 flood passes per cascade). */
LUX_API int lux_ddgi_build_sdf_mip(LuxDDGIContext* ctx);
/* Per-frame partial refresh of the bound global SDF = the cached path of merge_sdf::system (GlobalDistanceField.cpp:652-657: a cascade is
 * revisited every 2 / 3 / 5 / 11 frames; :775-848: only chunks whose object lists changed are re-rasterized, then the cascade's mip is rebuilt
 * and flooded).  The texels of the 32^3-voxel rasterize chunks [chunkMin, chunkMax] (inclusive chunk coordinates, clipped to the volume) of
 * `cascade` replace the bound ones - in the linear volume AND in the layered-texture copy the trace reads - without re-allocating or re-copying
 * anything else.  texelsR16F = the region as a dense box [dz][dy][dx] of fp16 texels, dx = 32 * (chunkMax[0] - chunkMin[0] + 1) etc. (clipped
 * the same way); NULL = "the bound volume is caller-owned device memory (LUX_MEM_DEVICE) and already holds the new texels: refresh the texture
 * copy of the region only".  rebuildMip != 0 rebuilds the mip of that cascade on device (downsample + 4 flood passes, as the reference does after
 * any chunk dispatch) and refreshes its texture copy.  Cost is proportional to the region (plus the cascade's mip when asked). */
LUX_API int lux_ddgi_update_global_sdf_region(LuxDDGIContext* ctx, uint32_t cascade, const int32_t chunkMin[3], const int32_t chunkMax[3],
                                              const void* texelsR16F, LuxMemKind kind, int32_t rebuildMip);
/* Reader of the reference's baked .sdf files (cereal binary, SDFBaker.cpp:158-204 / :207-240).  Call with out == NULL to get the sizes:
 * size[3], mipCount and the total number of fp16 texels over all mips; then with a buffer of that many uint16_t (mips back to back). */
LUX_API int lux_ddgi_sdf_file_read(const char* path, uint32_t size[3], int32_t* mipCount, uint64_t* texels, uint16_t* out);

/* SDFAtlasChunkBuffer, SDFCullObjectBuffer, SDFObjectBuffer, SDFAtlasTileBuffer, uSurfaceAtlasTex (RGBA16F,
 * linear/repeat), uSurfaceAtlasDepth (D32F, clamp), UniformBufferObject.data (DDGIRenderer.cpp:306-313).
 * chunks has 40^3 entries. */
LUX_API int lux_ddgi_set_surface_atlas(LuxDDGIContext* ctx, const LuxGlobalSurfaceAtlasData* data,
                               const uint32_t* chunks, const uint32_t* cullObjects, size_t cullObjectsCount,
                               const LuxObjectBuffer* objects, size_t objectsCount,
                               const LuxTileBuffer* tiles, size_t tilesCount,
                               const void* lightCacheRGBA16F, const float* depthD32F, LuxMemKind kind);
/* Per-frame relight: replaces the light-cache texels only (same resolution). */
LUX_API int lux_ddgi_update_surface_light_cache(LuxDDGIContext* ctx, const void* lightCacheRGBA16F, LuxMemKind kind);
/* Same for a row range of the atlas.  With a communicator bound (lux_ddgi_set_nccl_comm) rank r passes rows [r*res/world, (r+1)*res/world)
 * and the ranks all-gather them in place over NVLink: one light cache crosses PCIe per frame in total, not one per GPU. */
LUX_API int lux_ddgi_update_surface_light_cache_rows(LuxDDGIContext* ctx, const void* lightRowsRGBA16F, int32_t rowBegin, int32_t rowCount, LuxMemKind kind);

/* uSkybox: 6 faces (+X,-X,+Y,-Y,+Z,-Z) of faceSize^2 RGBA16F texels.  Default = the reference's 1x1 black
 * fallback cube (DDGIRenderer.cpp:308). */
/* f4 (first half): surface::culling + Shaders/SDF/SDFCulling.comp:36-101 on device.  Rebuilds SDFAtlasChunkBuffer / SDFCullObjectBuffer of
 * the bound surface cache from its object buffer: per chunk the ids (ascending) of the objects whose bounding sphere touches the chunk.
 * Lists are laid out in ascending chunk address (the shader's atomic allocation makes its own layout scheduling dependent);
 * capacityWords = the shader's culledObjectsCapacity (lists that do not fit are dropped), 0 = make everything fit. */
LUX_API int lux_ddgi_cull_surface_objects(LuxDDGIContext* ctx, uint32_t capacityWords);
LUX_API int lux_ddgi_get_surface_cull_lists(LuxDDGIContext* ctx, void** chunksDevice, void** cullDevice, size_t* cullWords);
LUX_API int lux_ddgi_set_skybox(LuxDDGIContext* ctx, int32_t faceSize, const void* facesRGBA16F, LuxMemKind kind);

/* The three stages + frame bookkeeping, in the order RenderGraph.cpp:98-114 runs them. */
LUX_API int lux_ddgi_trace_rays(LuxDDGIContext* ctx, const LuxTracePushConstants* pushConsts);
LUX_API int lux_ddgi_probe_update(LuxDDGIContext* ctx);  /* blend both atlases; borders fused unless UNFUSED_BORDER */
LUX_API int lux_ddgi_border_update(LuxDDGIContext* ctx); /* idempotent; a no-op cost-wise when borders were fused   */
LUX_API int lux_ddgi_end_frame(LuxDDGIContext* ctx);     /* pingPong ^= 1; frames++                                 */

/* trace_rays + probe_update (+ border_update if unfused) + end_frame with rotation `orientation`
 * (column-major mat4, as pushed at DDGIRenderer.cpp:271-272).  Asynchronous on the context's stream. */
LUX_API int lux_ddgi_update(LuxDDGIContext* ctx, const float orientation[16]);

/* Exchange step of a sharded volume.  `ncclComm` is the caller's ncclComm_t over the `world` ranks given at create time
 * (rank order = shard order).  With a communicator bound, lux_ddgi_update ends with one in-place ncclAllGather per atlas (own slab rows ->
 * every rank's full atlas) on an internal stream: it overlaps the next update's trace, the next blend into that atlas pair waits for it, and
 * every reader of whole atlases (downloads, lux_ddgi_sample_*, lux_ddgi_indirect_light, lux_ddgi_synchronize) is ordered after it.
 * libnccl.so.2 is bound with dlopen on first use.  NULL unbinds (the host then exchanges rows itself, see lux_ddgi_get_state). */
LUX_API int lux_ddgi_set_nccl_comm(LuxDDGIContext* ctx, void* ncclComm);
LUX_API int lux_ddgi_synchronize(LuxDDGIContext* ctx);

/* Outputs.  Device pointers stay valid until destroy; `bytes` may be NULL. */
LUX_API int lux_ddgi_get_buffer(LuxDDGIContext* ctx, LuxBufferId id, void** devicePtr, size_t* bytes);
LUX_API int lux_ddgi_download(LuxDDGIContext* ctx, LuxBufferId id, void* host, size_t bytes);
/* Same, without the trailing synchronize: `pinnedHost` must be page-locked; order with lux_ddgi_synchronize. */
LUX_API int lux_ddgi_download_async(LuxDDGIContext* ctx, LuxBufferId id, void* pinnedHost, size_t bytes);
/* Rows [rowBegin, rowBegin + rowCount) of an atlas (e.g. the shard's own rows from lux_ddgi_get_state) into pinned memory. */
LUX_API int lux_ddgi_download_rows_async(LuxDDGIContext* ctx, LuxBufferId id, int32_t rowBegin, int32_t rowCount, void* pinnedHost);
/* Frame-pipelined consumers: a fence marks "every download enqueued so far"; waiting on the fence of frame f-1 while frame f computes
 * keeps the device busy and still delivers every frame's atlases to the host (one frame of latency).  Up to 8 fences may be pending. */
LUX_API int lux_ddgi_download_fence(LuxDDGIContext* ctx, uint64_t* fence);
LUX_API int lux_ddgi_wait_fence(LuxDDGIContext* ctx, uint64_t fence);
/* Overwrite the ray buffers (this shard's rows) — what probe_update consumes is whatever these hold, exactly as the
 * reference's blend reads the iRadiance / iDirectionDistance images (ProbeUpdate.glsl:53-64).  Lets a host (or a test)
 * run the blend stage on rays produced elsewhere. */
LUX_API int lux_ddgi_set_ray_buffers(LuxDDGIContext* ctx, const void* radianceRGBA16F, const void* directionDistanceRGBA16F,
                             LuxMemKind kind);
/* Checkpoint/resume of the probe state (the reference never saves it): load both atlases + counters. */
LUX_API int lux_ddgi_restore(LuxDDGIContext* ctx, const void* irradianceRGBA16F, const void* depthRG16F,
                     int32_t frames, int32_t pingPong);

/* ---- consumer side (the step after the path) ----
 * sampleIrradiance() of Shaders/DDGI/DDGICommon.glsl:163-233 against the atlases most recently written, for `count` points:
 * P, N, Wo and out are [count][3] floats.  Both atlas taps are bilinear through the 1-texel borders (linear / repeat sampler). */
LUX_API int lux_ddgi_sample_irradiance(LuxDDGIContext* ctx, int32_t count, const float* P, const float* N, const float* Wo, float* out,
                                       LuxMemKind kind);
/* sample_probe::system (DDGIRenderer.cpp:467-503) = Shaders/DDGI/SampleProbe.comp: per pixel of a width x height G-buffer
 * (depth D32F [h][w]; normals RGBA32F [h][w][4] with xy = octahedral normal, GBuffer.cpp:17) reconstruct the world position with
 * viewProjInv (column-major), sample the probe volume and write RGBA32F [h][w][4] (the INDIRECT_LIGHTING target, GBuffer.cpp:24). */
LUX_API int lux_ddgi_sample_probe(LuxDDGIContext* ctx, int32_t width, int32_t height, const float* depthD32F, const float* normalsRGBA32F,
                                  const float cameraPosition[4], const float viewProjInv[16], float* outRGBA32F, LuxMemKind kind);

/* Measurement aid (the L2 bandwidth the request-level roofline of the trace is quoted against is measured, not assumed): read-only
 * sweeps of a `bytes`-sized device buffer (rounded down to a multiple of 16 KiB; default 64 MiB when 0) that fits in L2, every SM reading the
 * whole buffer, timed with CUDA events on the context's stream; best of `repeats` (>= 1) launches after one warm-up.  *gbPerSecond = bytes read / s / 1e9. */
LUX_API int lux_ddgi_measure_l2_read_bandwidth(LuxDDGIContext* ctx, size_t bytes, int32_t repeats, float* gbPerSecond);

/* ---- the other tracyGlobalSDF users: screen-space passes over a width x height G-buffer ----
 * Blue-noise inputs (Raytraced/BlueNoise.glsl:8-19): `sobolRGBA8` = the 256 x 1 RGBA8 texels of textures/blue_noise/sobol_256_4d.png,
 * `scramblingRankingRGBA8` = the 128 x 128 RGBA8 texels of scrambling_ranking_128x128_2d_1spp.png (Engine/Noise/BlueNoise.h:15-26); a texel
 * channel c reads as float(c) / 255 (UNORM). */

/* Push constants of Shaders/SDF/SDFReflection.comp:66-78 (host: Engine/Raytrace/RaytracedReflection.cpp), 112 bytes. */
typedef struct LuxReflectionPushConstants {
    float    bias;               /* unused by the SDF branch */
    float    trim;
    float    intensity;          /* unused by the SDF branch */
    float    roughDDGIIntensity;
    uint32_t numLights;          /* unused */
    uint32_t numFrames;
    uint32_t sampleGI;           /* unused */
    uint32_t approximateWithDDGI;
    float    cameraPosition[4];
    float    viewProjInv[16];    /* column-major */
} LuxReflectionPushConstants;

/* Shaders/SDF/SDFReflection.comp:84-163 (dispatch: RaytracedReflection.cpp:335-339): per pixel with depth != 1 one reflection ray from the
 * G-buffer surface (pushed out by one cascade-0 voxel) through the bound global SDF.  roughness < 0.05: mirror ray; roughness > 0.45 and
 * approximateWithDDGI == 1: roughDDGIIntensity * sampleIrradiance along the mirror direction (alpha 0); otherwise a GGX half vector from the
 * blue-noise sample.  A hit returns the surface cache sampled with normal = -R (rgb normalised, alpha = weight sum), a miss the skybox texel.
 * depth D32F [h][w]; normals RGBA32F [h][w][4] (xy = octahedral normal); pbr RGBA32F [h][w][4] (g = roughness); outRGBA16F [h][w][4] is
 * read-modify-write: pixels with depth == 1 keep their previous contents, as the shader does not store them. */
LUX_API int lux_ddgi_sdf_reflection(LuxDDGIContext* ctx, const LuxReflectionPushConstants* push, int32_t width, int32_t height, const float* depthD32F,
                                    const float* normalsRGBA32F, const float* pbrRGBA32F, const uint8_t* sobolRGBA8,
                                    const uint8_t* scramblingRankingRGBA8, void* outRGBA16F, LuxMemKind kind);

/* Shaders/SDF/SDFShadow.comp:121-157 (UBO :25-32; host: Engine/Raytrace/RaytracedShadow.cpp): per pixel with depth != 1 one soft-shadow ray
 * towards `light` (disk sample from the blue noise, light->direction[3] = light radius) through the bound global SDF, start bias 0, max
 * distance tMax - bias.  Output: one uint32 per 8 x 4 pixel workgroup, bit (y % 4) * 8 + (x % 8) set = visible, [height/4][width/8]
 * (R32UI image); width must be a multiple of 8 and height of 4.  As in the shader, a workgroup whose first pixel has depth == 1 stores
 * nothing (its word keeps its previous contents) and pixels with depth == 1 contribute no bit: outMaskR32UI is read-modify-write. */
LUX_API int lux_ddgi_sdf_shadow(LuxDDGIContext* ctx, const LuxLight* light, const float viewProjInv[16], uint32_t numFrames, float shadowBias,
                                int32_t width, int32_t height, const float* depthD32F, const float* normalsRGBA32F, const uint8_t* sobolRGBA8,
                                const uint8_t* scramblingRankingRGBA8, uint32_t* outMaskR32UI, LuxMemKind kind);

/* Infinite-bounce feedback: surface::indirect_light::system (GlobalSurfaceAtlas.cpp:1004-1085) =
 * Shaders/SDF/SDFAtlasIndirectLight.frag:44-67, additive into the RGBA16F light cache.  For each listed atlas texel t:
 *   light[t].rgb = fp16( base[t].rgb + intensity * (min(albedo,0.9) - min(albedo,0.9)*metallic)/PI * sampleIrradiance(P, N, normalize(cameraPos-P)) )
 *   light[t].a   = fp16( base[t].a + 1 )   (the shader writes alpha 1 and the pass blends ONE + ONE on alpha as well, RHI/Vulkan/VulkanPipeline.cpp:160-165)
 * `baseLightRGBA16F` (full atlas; emissive + direct light, i.e. the cache after CopyEmissive + SDFDeferredLight) may be NULL to
 * add onto the current contents.  The rasterisation of tiles into texel lists stays with the caller. */
LUX_API int lux_ddgi_indirect_light(LuxDDGIContext* ctx, const void* baseLightRGBA16F, int32_t count, const uint32_t* texelIndex,
                                    const float* worldPos, const float* normal, const float* albedo, const float* metallic, float intensity,
                                    const float cameraPos[3], LuxMemKind kind);
LUX_API int lux_ddgi_get_surface_light_cache(LuxDDGIContext* ctx, void** devicePtr, size_t* bytes);

LUX_API int lux_ddgi_get_state(LuxDDGIContext* ctx, LuxDDGIState* out);
/* z-slab layout of shard `rank` of `world` without a context (pure host arithmetic, usable on a machine with no GPU):
 * fills probeBegin/Count and the atlas row ranges of `out`; the other fields are zero. */
LUX_API int lux_ddgi_shard_layout(const LuxDDGIUniform* uniform, int32_t rank, int32_t world, LuxDDGIState* out);
/* ... for the layout `flags` select (LUX_DDGI_FLAG_SHARD_INTERLEAVED) */
LUX_API int lux_ddgi_shard_layout_ex(const LuxDDGIUniform* uniform, int32_t rank, int32_t world, uint32_t flags, LuxDDGIState* out);
/* The shard's OWN rows of an atlas, packed in shard-local layer order (irradianceRowCount | depthRowCount rows), to pinned host memory on the
 * download stream as soon as the blend that wrote them has finished (one strided copy; for a z-slab the same as lux_ddgi_download_rows_async
 * of the own range).  Completion: lux_ddgi_download_fence / lux_ddgi_wait_fence. */
LUX_API int lux_ddgi_download_shard_async(LuxDDGIContext* ctx, LuxBufferId id, void* pinnedHost);
LUX_API int lux_ddgi_get_stage_ms(LuxDDGIContext* ctx, LuxStageTimes* out);

#ifdef __cplusplus
}
#endif
#endif /* LUXDDGI_H */
