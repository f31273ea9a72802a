// ddgi_kernels.cu — sm_100a kernels of the DDGI probe update: per-frame setup, SDF sphere trace with surface-cache
// radiance, irradiance / depth blend with hysteresis and fused border, standalone border.
//
// What each kernel replaces in the reference (paths relative to Code/Maple/src/):
//   ray_dirs_kernel          Shaders/DDGI/GISDFRays.comp:73 + DDGICommon.glsl:42-52 (hoisted: directions depend on rayId only)
//   blend_weights_kernel     Shaders/DDGI/ProbeUpdate.glsl:83-91 (hoisted: weights depend on (texel, rayId) only, SURVEY finding 5)
//   trace_kernel             Shaders/DDGI/GISDFRays.comp:63-128, Shaders/SDF/SDFCommon.glsl:64-199, Shaders/SDF/AtlasCommon.glsl:40-157
//   blend_irradiance_kernel  Shaders/DDGI/ProbeUpdate.glsl:105-152 (+ BorderUpdate.glsl fused in the epilogue)
//   blend_depth_kernel       same with DEPTHPROBE_UPDATE
//   border_kernel            Shaders/DDGI/BorderUpdate.glsl:136-156
//
// Compiled with -fmad=false; see ddgi_math.cuh for the numerics contract.
#include <cstdlib>
#include <cstring>

#include <cuda.h> // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: libcuda is not linked)

#include "ddgi_kernels.h"
#include "ddgi_math.cuh"

namespace lux {

struct RotationArg { float m[16]; };

// probe id of shard-local probe `local` (LuxDDGIState::layerStride): a z-slab, or every world-th z-layer
__device__ __forceinline__ int shard_probe(int probeBegin, int layerProbes, int layerStride, int local)
{
    return layerStride == 1 ? probeBegin + local : probeBegin + (local / layerProbes) * layerStride * layerProbes + local % layerProbes;
}
template <class Params>
__device__ __forceinline__ int shard_probe(const Params& P, int local) { return shard_probe(P.probeBegin, P.layerProbes, P.layerStride, local); }

// =====================================================================================================================
// Per-frame setup
// =====================================================================================================================

// DDGICommon.glsl:42-52 with the constants glslang folded into the shipped GISDFRays.comp.spv.
__device__ __forceinline__ f3 spherical_fibonacci(float i, float raysPerProbe)
{
    const float PHI_M1 = 0.61803400516510009765625f;
    const float TWO_PI = 6.283185482025146484375f;
    float ab       = i * PHI_M1;
    float phi      = TWO_PI * (ab - floorf(ab));
    float cosTheta = 1.0f - (2.0f * i + 1.0f) * __fdiv_rn(1.0f, raysPerProbe);
    float sinTheta = __fsqrt_rn(gclamp(1.0f - cosTheta * cosTheta, 0.0f, 1.0f));
    return {cos_rn(phi) * sinTheta, sin_rn(phi) * sinTheta, cosTheta};
}

__global__ void ray_dirs_kernel(const RotationArg rot, int R, float4* __restrict__ dirs, uint2* __restrict__ dirsHalf)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R)
        return;
    f3 d = normalize3(mat3_mul(rot.m, spherical_fibonacci((float)r, (float)R)));
    dirs[r] = make_float4(d.x, d.y, d.z, 0.0f);
    if (dirsHalf) // the direction as the blend reads it back from the RGBA16F ray buffer (ProbeUpdate.glsl:57)
        dirsHalf[r] = make_uint2((uint32_t)f2h_bits(d.x) | ((uint32_t)f2h_bits(d.y) << 16), (uint32_t)f2h_bits(d.z));
}

__device__ __forceinline__ float sign_not_zero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; }

// Column order of the depth weight matrix: column n = block*8 + (j%2)*4 + (i%4) with block = (j/2)*4 + i/4, i.e. the 16x16
// octahedral map is cut into 4x2-texel blocks.  A block spans ~13 degrees, so "all 8 weights of this ray are zero" holds
// for ~3/4 of the (ray, block) pairs (depth weights vanish beyond 46 degrees); rows of the map, in contrast, wind across a
// whole great arc and are almost always live.  Only the ORDER OF COLUMNS changes: each texel still sums its rays in order.
__device__ __forceinline__ int depth_column_of_texel(int i, int j) { return (((j >> 1) << 2) + (i >> 2)) * 8 + ((j & 1) << 2) + (i & 3); }
__device__ __forceinline__ void depth_texel_of_column(int n, int& i, int& j)
{
    int block = n >> 3, w = n & 7;
    i = ((block & 3) << 2) + (w & 3);
    j = ((block >> 2) << 1) + (w >> 2);
}

// DDGICommon.glsl:74-92 for interior texel (i, j) of a probe with `side` texels per side
__device__ __forceinline__ f3 texel_direction(int i, int j, int side)
{
    float s  = __fdiv_rn(2.0f, (float)side);
    float ox = ((float)i + 0.5f) * s - 1.0f;
    float oy = ((float)j + 0.5f) * s - 1.0f;
    f3    v  = {ox, oy, (1.0f - fabsf(ox)) - fabsf(oy)};
    if (v.z < 0.0f)
    {
        float nx = (1.0f - fabsf(v.y)) * sign_not_zero(v.x);
        float ny = (1.0f - fabsf(v.x)) * sign_not_zero(v.y);
        v.x = nx;
        v.y = ny;
    }
    return normalize3(v);
}

// One thread per (texel, ray).  Texels 0..63 are irradiance texels, 64..319 depth texels.
// Ray directions are read back from row 0 of the direction/distance buffer, i.e. already quantised to fp16 exactly as
// the reference's blend sees them (texelFetch of an RGBA16F image, ProbeUpdate.glsl:59).
// Weights below the reference's gate (ProbeUpdate.glsl:93, w >= 1e-8) are stored as exact zeros.
__global__ void blend_weights_kernel(const uint2* __restrict__ dirsHalf, int R, int Rpad, float sharpness,
                                     float* __restrict__ wIrr, float* __restrict__ wDepth)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x; // texel 0..319
    int r = blockIdx.y;
    if (t >= 320)
        return;
    const float FLT_EPS = 0.00000001f;
    float       w       = 0.0f;
    if (r < R)
    {
        uint2 dh = dirsHalf[r];
        f3    rd = {h2f_bits((uint16_t)(dh.x & 0xffffu)), h2f_bits((uint16_t)(dh.x >> 16)), h2f_bits((uint16_t)(dh.y & 0xffffu))};
        if (t < 64)
        {
            f3 td = texel_direction(t & 7, t >> 3, 8);
            w     = gmax(0.0f, dot3(td, rd));
        }
        else
        {
            int u  = t - 64;
            f3  td = texel_direction(u & 15, u >> 4, 16);
            w      = pow_rn(gmax(0.0f, dot3(td, rd)), sharpness);
        }
        if (!(w >= FLT_EPS))
            w = 0.0f;
    }
    if (t < 64)
        wIrr[(size_t)r * 64 + t] = w;
    else
        wDepth[(size_t)r * 256 + depth_column_of_texel((t - 64) & 15, (t - 64) >> 4)] = w;
}

// Sequential sum over rays in ray order (the reference's accumulation order), then 1/(2*sum) (ProbeUpdate.glsl:133-134).
// Per ray, which groups of 8 consecutive texels carry any non-zero weight (bit g): lets the blend skip exact zeros.
__global__ void blend_nonzero_kernel(const float* __restrict__ wIrr, const float* __restrict__ wDepth, int Rpad,
                                     uint32_t* __restrict__ nzIrr, uint32_t* __restrict__ nzDepth)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Rpad)
        return;
    uint32_t zi = 0, zd = 0;
    for (int g = 0; g < 8; g++)
    {
        bool any = false;
        for (int j = 0; j < 8; j++)
            any |= wIrr[(size_t)r * 64 + g * 8 + j] != 0.0f;
        zi |= (any ? 1u : 0u) << g;
    }
    for (int g = 0; g < 32; g++)
    {
        bool any = false;
        for (int j = 0; j < 8; j++)
            any |= wDepth[(size_t)r * 256 + g * 8 + j] != 0.0f;
        zd |= (any ? 1u : 0u) << g;
    }
    nzIrr[r]   = zi;
    nzDepth[r] = zd;
}

__global__ void blend_scales_kernel(const float* __restrict__ wIrr, const float* __restrict__ wDepth, int R,
                                    float* __restrict__ scaleIrr, float* __restrict__ scaleDepth)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 320)
        return;
    const float FLT_EPS = 0.00000001f;
    float       total   = 0.0f;
    if (t < 64)
        for (int r = 0; r < R; r++)
            total += wIrr[(size_t)r * 64 + t];
    else
        for (int r = 0; r < R; r++)
            total += wDepth[(size_t)r * 256 + depth_column_of_texel((t - 64) & 15, (t - 64) >> 4)];
    float s = (total > FLT_EPS) ? __fdiv_rn(1.0f, 2.0f * total) : 1.0f;
    if (t < 64)
        scaleIrr[t] = s;
    else
        scaleDepth[t - 64] = s;
}

// inverse(object.transform) (AtlasCommon.glsl:133), hoisted to upload time.  Cofactor expansion fixed by the contract.
// inverse(mat4) of the numerics contract: cofactor expansion over 2x2 sub-determinants (DESIGN.md §4 rule 4)
__device__ __forceinline__ void inverse4_dev(const float* __restrict__ m, float* __restrict__ o)
{
#define A(r, c) m[(c)*4 + (r)]
#define B(r, c) o[(c)*4 + (r)]
    float s0 = A(0, 0) * A(1, 1) - A(1, 0) * A(0, 1);
    float s1 = A(0, 0) * A(1, 2) - A(1, 0) * A(0, 2);
    float s2 = A(0, 0) * A(1, 3) - A(1, 0) * A(0, 3);
    float s3 = A(0, 1) * A(1, 2) - A(1, 1) * A(0, 2);
    float s4 = A(0, 1) * A(1, 3) - A(1, 1) * A(0, 3);
    float s5 = A(0, 2) * A(1, 3) - A(1, 2) * A(0, 3);
    float c5 = A(2, 2) * A(3, 3) - A(3, 2) * A(2, 3);
    float c4 = A(2, 1) * A(3, 3) - A(3, 1) * A(2, 3);
    float c3 = A(2, 1) * A(3, 2) - A(3, 1) * A(2, 2);
    float c2 = A(2, 0) * A(3, 3) - A(3, 0) * A(2, 3);
    float c1 = A(2, 0) * A(3, 2) - A(3, 0) * A(2, 2);
    float c0 = A(2, 0) * A(3, 1) - A(3, 0) * A(2, 1);
    float det = ((((s0 * c5 - s1 * c4) + s2 * c3) + s3 * c2) - s4 * c1) + s5 * c0;
    float id  = __fdiv_rn(1.0f, det);
    B(0, 0) = ((A(1, 1) * c5 - A(1, 2) * c4) + A(1, 3) * c3) * id;
    B(0, 1) = ((-A(0, 1) * c5 + A(0, 2) * c4) - A(0, 3) * c3) * id;
    B(0, 2) = ((A(3, 1) * s5 - A(3, 2) * s4) + A(3, 3) * s3) * id;
    B(0, 3) = ((-A(2, 1) * s5 + A(2, 2) * s4) - A(2, 3) * s3) * id;
    B(1, 0) = ((-A(1, 0) * c5 + A(1, 2) * c2) - A(1, 3) * c1) * id;
    B(1, 1) = ((A(0, 0) * c5 - A(0, 2) * c2) + A(0, 3) * c1) * id;
    B(1, 2) = ((-A(3, 0) * s5 + A(3, 2) * s2) - A(3, 3) * s1) * id;
    B(1, 3) = ((A(2, 0) * s5 - A(2, 2) * s2) + A(2, 3) * s1) * id;
    B(2, 0) = ((A(1, 0) * c4 - A(1, 1) * c2) + A(1, 3) * c0) * id;
    B(2, 1) = ((-A(0, 0) * c4 + A(0, 1) * c2) - A(0, 3) * c0) * id;
    B(2, 2) = ((A(3, 0) * s4 - A(3, 1) * s2) + A(3, 3) * s0) * id;
    B(2, 3) = ((-A(2, 0) * s4 + A(2, 1) * s2) - A(2, 3) * s0) * id;
    B(3, 0) = ((-A(1, 0) * c3 + A(1, 1) * c1) - A(1, 2) * c0) * id;
    B(3, 1) = ((A(0, 0) * c3 - A(0, 1) * c1) + A(0, 2) * c0) * id;
    B(3, 2) = ((-A(3, 0) * s3 + A(3, 1) * s1) - A(3, 2) * s0) * id;
    B(3, 3) = ((A(2, 0) * s3 - A(2, 1) * s1) + A(2, 2) * s0) * id;
#undef A
#undef B
}

__global__ void object_inverse_kernel(const LuxObjectBuffer* __restrict__ objects, int count, float* __restrict__ inv)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    inverse4_dev(objects[k].transform, inv + (size_t)k * 16);
}

// Prefilter for the surface-cache object loop (AtlasCommon.glsl:125-139), built once per surface-cache upload.
// One block per culling chunk, one thread per quarter-chunk sub-cell (4x4x4): bit j of the thread's 64-bit mask stays
// set iff list entry j could pass BOTH the bounding-sphere test and the OBB-extent test for SOME point of the sub-cell
// (expanded by eps; sub-cells on the outer faces of the 40^3 grid are unbounded because chunk coordinates are clamped).
// thrMax = 1.05 * the largest cascade voxel size, the largest surfaceThreshold the trace can pass.
__global__ void chunk_masks_kernel(const uint32_t* __restrict__ chunks, const uint32_t* __restrict__ cull,
                                   const LuxObjectBuffer* __restrict__ objects, const float* __restrict__ objectInverse,
                                   uint32_t objectsCountLimit, float chunkSize, float thrMax, unsigned long long* __restrict__ masks)
{
    const int N = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    const int chunk = blockIdx.x, sub = threadIdx.x;
    const int cx = chunk % N, cy = (chunk / N) % N, cz = chunk / (N * N);
    const int sx = sub & 3, sy = (sub >> 2) & 3, sz = sub >> 4;
    unsigned long long mask = 0ull;
    uint32_t start = chunks[chunk];
    if (start != 0)
    {
        uint32_t count = cull[start];
        if (count <= objectsCountLimit)
        {
            const float BIG = 1e18f, eps = 1e-3f * chunkSize, cell = chunkSize * 0.25f;
            int   ci[3] = {cx, cy, cz}, si[3] = {sx, sy, sz};
            float bc[3], hb[3];
            for (int a = 0; a < 3; a++)
            {
                float lo = ((float)ci[a] - 0.5f * (float)N) * chunkSize + (float)si[a] * cell;
                float hi = lo + cell;
                bc[a] = 0.5f * (lo + hi);
                hb[a] = 0.5f * cell + eps;
                if ((ci[a] == 0 && si[a] == 0) || (ci[a] == N - 1 && si[a] == 3))
                    hb[a] = BIG; // clamped chunk coordinate: the cell is unbounded outwards (inwards too, harmlessly)
            }
            uint32_t n = count < 64u ? count : 64u;
            for (uint32_t j = 0; j < n; j++)
            {
                uint32_t     obj = cull[start + 1 + j];
                const float* ob  = objects[obj].objectBounds;
                float d2 = 0.0f;
                for (int a = 0; a < 3; a++)
                {
                    float d = fabsf(ob[a] - bc[a]) - hb[a];
                    d = d > 0.0f ? d : 0.0f;
                    d2 += d * d;
                }
                float r = ob[3] + eps;
                bool keep = d2 <= r * r;
                if (keep)
                {
                    const float* W  = objectInverse + (size_t)obj * 16;
                    const float* ex = objects[obj].extends;
                    for (int i = 0; i < 3 && keep; i++)
                    {
                        float c = W[0 + i] * bc[0] + W[4 + i] * bc[1] + W[8 + i] * bc[2] + W[12 + i];
                        float rr = fabsf(W[0 + i]) * hb[0] + fabsf(W[4 + i]) * hb[1] + fabsf(W[8 + i]) * hb[2];
                        rr += 1e-4f * (fabsf(c) + rr) + eps; // slack for fp32 rounding of the exact test
                        if (fabsf(c) - rr > ex[i] + thrMax)
                            keep = false;
                    }
                }
                if (keep)
                    mask |= 1ull << j;
            }
        }
    }
    masks[(size_t)chunk * 64 + sub] = mask;
}

// =====================================================================================================================
// Trace
// =====================================================================================================================

struct SdfVolume
{
    const uint16_t* d;
    int             w, h, dep;
};

__device__ __forceinline__ float ld_h(const uint16_t* p) { return h2f_bits(__ldg(p)); }
__device__ __forceinline__ float lerp1(float a, float b, float t) { return a + t * (b - a); }

// texture(sampler3D, uvw).r: trilinear, clamp-to-edge, LOD 0.  Explicit fp16 loads, fp32 nested lerp x -> y -> z.
__device__ __forceinline__ float sample3D(const SdfVolume& t, float u, float v, float w)
{
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f, z = w * (float)t.dep - 0.5f;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float ax = x - fx, ay = y - fy, az = z - fz;
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;
    int   x0 = iclamp(ix, 0, t.w - 1), x1 = iclamp(ix + 1, 0, t.w - 1);
    int   y0 = iclamp(iy, 0, t.h - 1), y1 = iclamp(iy + 1, 0, t.h - 1);
    int   z0 = iclamp(iz, 0, t.dep - 1), z1 = iclamp(iz + 1, 0, t.dep - 1);
    const uint16_t* r00 = t.d + ((size_t)z0 * t.h + y0) * t.w;
    const uint16_t* r10 = t.d + ((size_t)z0 * t.h + y1) * t.w;
    const uint16_t* r01 = t.d + ((size_t)z1 * t.h + y0) * t.w;
    const uint16_t* r11 = t.d + ((size_t)z1 * t.h + y1) * t.w;
    float v000 = ld_h(r00 + x0), v100 = ld_h(r00 + x1);
    float v010 = ld_h(r10 + x0), v110 = ld_h(r10 + x1);
    float v001 = ld_h(r01 + x0), v101 = ld_h(r01 + x1);
    float v011 = ld_h(r11 + x0), v111 = ld_h(r11 + x1);
    float c00 = lerp1(v000, v100, ax);
    float c10 = lerp1(v010, v110, ax);
    float c01 = lerp1(v001, v101, ax);
    float c11 = lerp1(v011, v111, ax);
    float c0  = lerp1(c00, c10, ay);
    float c1  = lerp1(c01, c11, ay);
    return lerp1(c0, c1, az);
}

struct Hit
{
    f3       normal;
    float    time;
    uint32_t cascade;
    uint32_t steps;
    float    sdf;
};

// SDFCommon.glsl:84-95
__device__ __forceinline__ void line_hit_aabb(f3 s, f3 e, f3 bmin, f3 bmax, float& nearT, float& farT)
{
    f3 inv   = {__fdiv_rn(1.0f, e.x - s.x), __fdiv_rn(1.0f, e.y - s.y), __fdiv_rn(1.0f, e.z - s.z)};
    f3 enter = (bmin - s) * inv;
    f3 exit_ = (bmax - s) * inv;
    f3 mn    = {gmin(enter.x, exit_.x), gmin(enter.y, exit_.y), gmin(enter.z, exit_.z)};
    f3 mx    = {gmax(enter.x, exit_.x), gmax(enter.y, exit_.y), gmax(enter.z, exit_.z)};
    nearT    = gclamp(gmax(mn.x, gmax(mn.y, mn.z)), 0.0f, 1.0f);
    farT     = gclamp(gmin(mx.x, gmin(mx.y, mx.z)), 0.0f, 1.0f);
}

// SDFCommon.glsl:98-194 (tracyGlobalSDF) with minDistance 0, stepScale 1, needsHitNormal true, start bias 0
__device__ __forceinline__ Hit trace_global_sdf(const TraceParams& P, f3 origin, f3 dir, float maxDistance)
{
    const LuxGlobalSDFData& data = P.sdf;
    Hit hit;
    hit.steps   = 0;
    hit.time    = -1.0f;
    hit.normal  = {0.0f, 0.0f, 0.0f};
    hit.cascade = 0;
    hit.sdf     = 0.0f;

    const float stepScale = 1.0f, cascadeTraceStartBias = 0.0f;
    float traceMaxDistance    = gmin(maxDistance, data.cascadePosDistance[data.cascadesCount - 1][3] * 2.0f);
    float chunkSizeDistance   = __fdiv_rn((float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE, data.resolution);
    float chunkMarginDistance = __fdiv_rn((float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_MARGIN, data.resolution);
    float nextIntersectionStart = 0.0f;
    f3    traceEnd        = origin + dir * traceMaxDistance;
    float cascadesCountF  = (float)data.cascadesCount;
    SdfVolume tex = {P.tex, P.res * P.cascades, P.res, P.res};
    SdfVolume mip = {P.mip, P.mipRes * P.cascades, P.mipRes, P.mipRes};

    for (uint32_t cascade = 0; cascade < data.cascadesCount && hit.time < 0.0f; cascade++)
    {
        f3    c         = {data.cascadePosDistance[cascade][0], data.cascadePosDistance[cascade][1], data.cascadePosDistance[cascade][2]};
        float cd        = data.cascadePosDistance[cascade][3];
        float voxelSize = data.cascadeVoxelSize[cascade];
        float voxelHalf = voxelSize * 0.5f;
        f3    worldPosition = origin + dir * (voxelSize * cascadeTraceStartBias);
        f3    ext = {cd, cd, cd};

        float nearT, farT;
        line_hit_aabb(worldPosition, traceEnd, c - ext, c + ext, nearT, farT);
        nearT *= traceMaxDistance;
        farT *= traceMaxDistance;
        nearT = gmax(nearT, nextIntersectionStart);

        float stepTime = nearT;
        if (nearT >= farT)
            stepTime = farT;
        else
            nextIntersectionStart = farT;

        float    cascadeMaxDistance = cd * 2.0f;
        uint32_t step = 0;
        for (; step < LUX_GLOBAL_SDF_MAX_STEPS && stepTime < farT; step++)
        {
            f3 stepPosition = worldPosition + dir * stepTime;
            f3 pc           = stepPosition - c;
            f3 cuv = {gclamp(__fdiv_rn(pc.x, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f),
                      gclamp(__fdiv_rn(pc.y, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f),
                      gclamp(__fdiv_rn(pc.z, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f)};
            f3 uvw = {__fdiv_rn((float)cascade + cuv.x, cascadesCountF), cuv.y, cuv.z};

            float stepDistance = sample3D(mip, uvw.x, uvw.y, uvw.z);
            if (stepDistance < chunkSizeDistance)
            {
                float stepDistanceTex = sample3D(tex, uvw.x, uvw.y, uvw.z);
                if (stepDistanceTex < chunkMarginDistance * 2.0f)
                    stepDistance = stepDistanceTex;
            }
            else
                stepDistance = chunkSizeDistance;

            stepDistance *= cascadeMaxDistance;

            float minSurfaceThickness = voxelHalf * gclamp(__fdiv_rn(stepTime, voxelSize), 0.0f, 1.0f);
            if (stepDistance < minSurfaceThickness)
            {
                hit.time    = gmax((stepTime + stepDistance) - minSurfaceThickness, 0.0f);
                hit.cascade = cascade;
                hit.sdf     = stepDistance;
                float o  = __fdiv_rn(1.0f, data.resolution);
                float xp = sample3D(tex, uvw.x + o, uvw.y, uvw.z);
                float xn = sample3D(tex, uvw.x - o, uvw.y, uvw.z);
                float yp = sample3D(tex, uvw.x, uvw.y + o, uvw.z);
                float yn = sample3D(tex, uvw.x, uvw.y - o, uvw.z);
                float zp = sample3D(tex, uvw.x, uvw.y, uvw.z + o);
                float zn = sample3D(tex, uvw.x, uvw.y, uvw.z - o);
                hit.normal = normalize3({xp - xn, yp - yn, zp - zn});
                break;
            }
            stepTime += gmax(stepDistance * stepScale, voxelSize);
        }
        hit.steps += step;
    }
    return hit;
}

__device__ __forceinline__ int wrap_texel(int a, int n)
{
    if (a >= -n && a < 2 * n)
    {
        a = a < 0 ? a + n : a;
        return a >= n ? a - n : a;
    }
    return ((a % n) + n) % n;
}

__device__ __forceinline__ void gather_coords(float u, float v, int W, int H, bool repeat, int& i0, int& i1, int& j0, int& j1)
{
    int a = (int)floorf(u * (float)W - 0.5f);
    int b = (int)floorf(v * (float)H - 0.5f);
    if (repeat)
    { // repeat addressing; atlas coordinates live in [0, 1], so the texel index is within one period of the range: a conditional add /
      // subtract gives the same residue as the two integer modulos (~20 instructions each), which remain for anything further out
        i0 = wrap_texel(a, W); i1 = wrap_texel(a + 1, W);
        j0 = wrap_texel(b, H); j1 = wrap_texel(b + 1, H);
    }
    else
    {
        i0 = iclamp(a, 0, W - 1); i1 = iclamp(a + 1, 0, W - 1);
        j0 = iclamp(b, 0, H - 1); j1 = iclamp(b + 1, 0, H - 1);
    }
}

__device__ __forceinline__ f4 unpack_rgba16f(uint2 t)
{
    return {h2f_bits((uint16_t)(t.x & 0xffffu)), h2f_bits((uint16_t)(t.x >> 16)), h2f_bits((uint16_t)(t.y & 0xffffu)),
            h2f_bits((uint16_t)(t.y >> 16))};
}

// AtlasCommon.glsl:58-112: one tile of one object
__device__ __forceinline__ f4 sample_atlas_tile(const TraceParams& P, const LuxTileBuffer* __restrict__ tile, f3 localPosition,
                                                f3 normal, float surfaceThreshold)
{
    float tm[16];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float4 c = __ldg(reinterpret_cast<const float4*>(tile->transform) + i);
        tm[i * 4 + 0] = c.x; tm[i * 4 + 1] = c.y; tm[i * 4 + 2] = c.z; tm[i * 4 + 3] = c.w;
    }
    f3 nt = normalize3(mat4_mul_point(tm, normal, 1.0f));
    float normalWeight = gclamp(nt.z, 0.0f, 1.0f);
    normalWeight = __fdiv_rn(normalWeight - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD, 1.0f - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD);
    if (normalWeight <= 0.0f)
        return {0.0f, 0.0f, 0.0f, 0.0f};

    float4 ext = __ldg(reinterpret_cast<const float4*>(tile->extends));
    float4 ob  = __ldg(reinterpret_cast<const float4*>(tile->objectBounds));
    f3     tp  = mat4_mul_point(tm, localPosition, 1.0f);
    float  tileDepth = __fdiv_rn(tp.z, ob.z);
    float  tu = gclamp(__fdiv_rn(tp.x, ob.x) + 0.5f, 0.0f, 1.0f);
    float  tv = gclamp(__fdiv_rn(tp.y, ob.y) + 0.5f, 0.0f, 1.0f);
    float  au = tu * ext.z + ext.x, av = tv * ext.w + ext.y;
    float  res = (float)P.atlasRes;
    float  fx = gfract(au * res + 0.5f), fy = gfract(av * res + 0.5f);
    f4     bw = {(1.0f - fx) * fy, fx * fy, fx * (1.0f - fy), (1.0f - fx) * (1.0f - fy)};

    int R = (int)P.atlasRes, i0, i1, j0, j1;
    gather_coords(au, av, R, R, false, i0, i1, j0, j1);
    float z4[4] = {__ldg(P.depth + (size_t)j1 * R + i0), __ldg(P.depth + (size_t)j1 * R + i1), __ldg(P.depth + (size_t)j0 * R + i1),
                   __ldg(P.depth + (size_t)j0 * R + i0)};
    float depthThreshold = __fdiv_rn(2.0f * surfaceThreshold, ob.z);
    float vis[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        vis[i] = 1.0f - gclamp(__fdiv_rn(fabsf(tileDepth - z4[i]) - depthThreshold, 0.5f * depthThreshold), 0.0f, 1.0f);
        if (z4[i] >= 1.0f)
            vis[i] = 0.0f;
    }
    f4    visv = {vis[0], vis[1], vis[2], vis[3]};
    float sampleWeight = dot4(visv, bw);
    sampleWeight *= normalWeight;
    if (sampleWeight <= 0.0f)
        return {0.0f, 0.0f, 0.0f, 0.0f};

    bw = {bw.x * visv.x, bw.y * visv.y, bw.z * visv.z, bw.w * visv.w};
    gather_coords(au, av, R, R, true, i0, i1, j0, j1);
    f4 t0 = unpack_rgba16f(__ldg(P.light + (size_t)j1 * R + i0));
    f4 t1 = unpack_rgba16f(__ldg(P.light + (size_t)j1 * R + i1));
    f4 t2 = unpack_rgba16f(__ldg(P.light + (size_t)j0 * R + i1));
    f4 t3 = unpack_rgba16f(__ldg(P.light + (size_t)j0 * R + i0));
    float cr = dot4({t0.x, t1.x, t2.x, t3.x}, bw);
    float cg = dot4({t0.y, t1.y, t2.y, t3.y}, bw);
    float cb = dot4({t0.z, t1.z, t2.z, t3.z}, bw);
    return {cr * sampleWeight, cg * sampleWeight, cb * sampleWeight, sampleWeight};
}

// AtlasCommon.glsl:115-157 (sampleGlobalSurfaceAtlas, debug = false)
__device__ __forceinline__ f4 sample_global_surface_atlas(const TraceParams& P, f3 worldPosition, f3 worldNormal, float surfaceThreshold)
{
    f4 result = {0.0f, 0.0f, 0.0f, 0.0f};
    if (!P.hasAtlas)
        return result;
    const float half = (float)LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * 0.5f;
    const int   N    = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    int cx = iclamp((int)floorf(__fdiv_rn(worldPosition.x, P.chunkSize) + half), 0, N - 1);
    int cy = iclamp((int)floorf(__fdiv_rn(worldPosition.y, P.chunkSize) + half), 0, N - 1);
    int cz = iclamp((int)floorf(__fdiv_rn(worldPosition.z, P.chunkSize) + half), 0, N - 1);
    uint32_t objectsStart = __ldg(P.chunks + (cz * N * N + cy * N + cx));
    if (objectsStart == 0)
        return result;
    uint32_t objectsCount = __ldg(P.cull + objectsStart);
    if (objectsCount > P.objectsCount)
        return result;
    objectsStart++;
    for (uint32_t k = 0; k < objectsCount; k++)
    {
        uint32_t               objectAddress = __ldg(P.cull + objectsStart++);
        const LuxObjectBuffer* object        = P.objects + objectAddress;
        float4                 ob            = __ldg(reinterpret_cast<const float4*>(object->objectBounds));
        f3                     bc            = {ob.x, ob.y, ob.z};
        if (length3(bc - worldPosition) > ob.w)
            continue;
        float wl[16];
        const float4* invp = reinterpret_cast<const float4*>(P.objectInverse + (size_t)objectAddress * 16);
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            float4 c = __ldg(invp + i);
            wl[i * 4 + 0] = c.x; wl[i * 4 + 1] = c.y; wl[i * 4 + 2] = c.z; wl[i * 4 + 3] = c.w;
        }
        f3     localPosition = mat4_mul_point(wl, worldPosition, 1.0f);
        float4 ex            = __ldg(reinterpret_cast<const float4*>(object->extends));
        if (fabsf(localPosition.x) > ex.x + surfaceThreshold || fabsf(localPosition.y) > ex.y + surfaceThreshold ||
            fabsf(localPosition.z) > ex.z + surfaceThreshold)
            continue;
        f3 normal = normalize3(mat3_mul(wl, worldNormal));
#pragma unroll 1
        for (int i = 0; i < 6; i++)
        {
            uint32_t tileOffset = __ldg(object->tileOffset + i);
            if (tileOffset != 0)
            {
                f4 s = sample_atlas_tile(P, P.tiles + tileOffset, localPosition, normal, surfaceThreshold);
                result.x += s.x; result.y += s.y; result.z += s.z; result.w += s.w;
            }
        }
    }
    float d = gmax(result.w, 0.0001f);
    result.x = __fdiv_rn(result.x, d);
    result.y = __fdiv_rn(result.y, d);
    result.z = __fdiv_rn(result.z, d);
    return result;
}

// texture(samplerCube, dir).rgb: Vulkan face selection, bilinear inside the face, clamp at the face edge
__device__ __forceinline__ f3 sample_sky(const TraceParams& P, f3 d)
{
    if (P.sky == nullptr || P.skyFace <= 0)
        return {0.0f, 0.0f, 0.0f};
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int   face;
    float sc, tc, ma;
    if (az >= ax && az >= ay) { face = d.z >= 0.0f ? 4 : 5; sc = d.z >= 0.0f ? d.x : -d.x; tc = -d.y; ma = az; }
    else if (ay >= ax)        { face = d.y >= 0.0f ? 2 : 3; sc = d.x; tc = d.y >= 0.0f ? d.z : -d.z; ma = ay; }
    else                      { face = d.x >= 0.0f ? 0 : 1; sc = d.x >= 0.0f ? -d.z : d.z; tc = -d.y; ma = ax; }
    float u = 0.5f * __fdiv_rn(sc, ma) + 0.5f, v = 0.5f * __fdiv_rn(tc, ma) + 0.5f;
    int   N = P.skyFace;
    float x = u * (float)N - 0.5f, y = v * (float)N - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float axw = x - fx, ayw = y - fy;
    int   x0 = iclamp((int)fx, 0, N - 1), x1 = iclamp((int)fx + 1, 0, N - 1);
    int   y0 = iclamp((int)fy, 0, N - 1), y1 = iclamp((int)fy + 1, 0, N - 1);
    const uint2* base = P.sky + (size_t)face * N * N;
    f4 t00 = unpack_rgba16f(__ldg(base + (size_t)y0 * N + x0)), t10 = unpack_rgba16f(__ldg(base + (size_t)y0 * N + x1));
    f4 t01 = unpack_rgba16f(__ldg(base + (size_t)y1 * N + x0)), t11 = unpack_rgba16f(__ldg(base + (size_t)y1 * N + x1));
    f3 out;
    out.x = lerp1(lerp1(t00.x, t10.x, axw), lerp1(t01.x, t11.x, axw), ayw);
    out.y = lerp1(lerp1(t00.y, t10.y, axw), lerp1(t01.y, t11.y, axw), ayw);
    out.z = lerp1(lerp1(t00.z, t10.z, axw), lerp1(t01.z, t11.z, axw), ayw);
    return out;
}

// Thread (x = probe lane, y = ray in group).  A warp holds 32 consecutive probes tracing the SAME ray direction:
// parallel rays from a row of probes stay spatially coherent (adjacent SDF rows, same surface-cache objects) and
// terminate after similar step counts, unlike the 32 divergent directions of one probe.
// fp16 Inf / NaN in any of the packed halves of an RGBA16F texel about to be stored: raise the frame's flag (read by the blend)
__device__ __forceinline__ void flag_non_finite(const TraceParams& P, uint2 texel)
{
    const bool bad = ((texel.x & 0x7c00u) == 0x7c00u) | ((texel.x & 0x7c000000u) == 0x7c000000u) | ((texel.y & 0x7c00u) == 0x7c00u) |
                     ((texel.y & 0x7c000000u) == 0x7c000000u);
    if (bad && P.nonFinite)
        atomicOr(P.nonFinite, 1u);
}

constexpr int TRACE_RAYS_PER_BLOCK = 8;

__global__ void __launch_bounds__(32 * TRACE_RAYS_PER_BLOCK) trace_kernel(const __grid_constant__ TraceParams P)
{
    __shared__ uint2 sRad[32][TRACE_RAYS_PER_BLOCK + 1];
    __shared__ uint2 sDir[32][TRACE_RAYS_PER_BLOCK + 1];

    const int lane       = threadIdx.x;
    const int rayInGroup = threadIdx.y;
    const int rayId      = blockIdx.x * TRACE_RAYS_PER_BLOCK + rayInGroup;
    const int probeLocal = blockIdx.y * 32 + lane;
    const bool active    = (rayId < P.raysPerProbe) && (probeLocal < P.probeCount);

    if (active)
    {
        const int probeId = shard_probe(P, probeLocal);
        // probeLocation, DDGICommon.glsl:101-114
        int cx = probeId % P.countX;
        int cy = (probeId % (P.countX * P.countY)) / P.countX;
        int cz = probeId / (P.countX * P.countY);
        f3  rayOrigin = {P.step[0] * (float)cx + P.start[0], P.step[1] * (float)cy + P.start[1], P.step[2] * (float)cz + P.start[2]};
        float4 d4 = __ldg(P.dirs + rayId);
        f3     direction = {d4.x, d4.y, d4.z};

        Hit hit = trace_global_sdf(P, rayOrigin, direction, LUX_GLOBAL_SDF_WORLD_SIZE);

        f4 radiance = {0.0f, 0.0f, 0.0f, 0.0f};
        if (hit.time >= 0.0f)
        {
            if (hit.sdf <= 0.0f && hit.time <= P.sdf.cascadeVoxelSize[0])
                radiance = {0.0f, 0.0f, 0.0f, LUX_GLOBAL_SDF_WORLD_SIZE};
            else
            {
                f3    hitPosition      = rayOrigin + direction * hit.time;
                float surfaceThreshold = P.sdf.cascadeVoxelSize[hit.cascade] * 1.05f;
                f4    sc = sample_global_surface_atlas(P, hitPosition, hit.normal, surfaceThreshold);
                radiance   = {sc.x, sc.y, sc.z, hit.time};
                radiance.w = gmax(radiance.w + P.sdf.cascadeVoxelSize[hit.cascade] * 0.5f, 0.0f);
            }
        }
        else
        {
            f3 s     = sample_sky(P, direction);
            radiance = {s.x, s.y, s.z, LUX_GLOBAL_SDF_WORLD_SIZE};
        }
        uint32_t r0 = f2h_bits(radiance.x), r1 = f2h_bits(radiance.y), r2 = f2h_bits(radiance.z);
        uint32_t d0 = f2h_bits(direction.x), d1 = f2h_bits(direction.y), d2 = f2h_bits(direction.z), d3 = f2h_bits(radiance.w);
        sRad[lane][rayInGroup] = make_uint2(r0 | (r1 << 16), r2); // alpha = 0
        sDir[lane][rayInGroup] = make_uint2(d0 | (d1 << 16), d2 | (d3 << 16));
        flag_non_finite(P, make_uint2(r0 | (r1 << 16), r2 | (d3 << 16)));
        if (P.steps)
            P.steps[(size_t)probeLocal * P.raysPerProbe + rayId] = (uint16_t)hit.steps;
    }
    __syncthreads();
    // transposed write-out: consecutive threads -> consecutive rays of one probe (64 contiguous bytes per probe)
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int pl  = tid / TRACE_RAYS_PER_BLOCK, rl = tid % TRACE_RAYS_PER_BLOCK;
    const int ray = blockIdx.x * TRACE_RAYS_PER_BLOCK + rl, probe = blockIdx.y * 32 + pl;
    if (ray < P.raysPerProbe && probe < P.probeCount)
    {
        size_t o = (size_t)probe * P.raysPerProbe + ray;
        P.radiance[o] = sRad[pl][rl];
        P.dirDist[o]  = sDir[pl][rl];
    }
}

// =====================================================================================================================
// Trace, wavefront form
// =====================================================================================================================
//
// Same arithmetic as trace_kernel, restructured for SIMT convergence (ncu on the simple kernel, C4: 7 of 32 lanes
// active on average; 11/32 in the march loop, 2/32 in the surface-cache loops — profiles/r1_v1_*):
//   * MARCH (march_kernel.inc): persistent warps pull 64-ray chunks (2 probes x 32 angularly adjacent directions) and their lanes REFILL from
//     the chunk as their rays terminate, so the step loop always runs (almost) full; a finished ray leaves a record;
//   * hits are not shaded where they occur: they are counting-sorted by culling chunk and shaded in that order (shade_sorted_kernel);
//   * surface-cache sampling first SCANS the chunk's object list for candidates and then loops over candidates and
//     tiles, so the expensive tile code is entered by all lanes together.  Tile contributions are accumulated in the
//     reference's (object, tile) order; skipped terms are exact zeros, so the sums are bit-identical.

constexpr int TW_MAX_CAND      = 8;
constexpr int CAND_STRIDE      = 256; // candidate scratch is [TW_MAX_CAND][256 threads]: conflict-free per lane

template <bool TEX>
struct SdfSampler;

template <>
struct SdfSampler<false>
{
    SdfVolume tex, mip;
    __device__ __forceinline__ SdfSampler(const TraceParams& P)
        : tex{P.tex, P.res * P.cascades, P.res, P.res}, mip{P.mip, P.mipRes * P.cascades, P.mipRes, P.mipRes} {}
    __device__ __forceinline__ float sampleTex(float u, float v, float w) const { return sample3D(tex, u, v, w); }
    __device__ __forceinline__ float sampleMip(float u, float v, float w) const { return sample3D(mip, u, v, w); }
};

// tld4 on a layered 2-D texture: the four texels of the bilinear footprint at (x, y) of `layer`, as
// .x=(i0,j1) .y=(i1,j1) .z=(i1,j0) .w=(i0,j0).  Coordinates are unnormalised; clamp addressing.
__device__ __forceinline__ float4 gather_layer(cudaTextureObject_t tex, float x, float y, int layer)
{
    float4 r;
    asm("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(tex), "r"(layer), "f"(x), "f"(y));
    return r;
}

// Trilinear tap through two layered gathers: the texture unit does addressing, clamping and fp16 decode, the lerps
// stay in fp32 ALU with full-precision weights (hardware filtering would quantise them to 8 bits).
__device__ __forceinline__ float sample3D_tex(cudaTextureObject_t obj, int W, int H, int D, float u, float v, float w)
{
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f, z = w * (float)D - 0.5f;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float ax = x - fx, ay = y - fy, az = z - fz;
    int   iz = (int)fz;
    int   z0 = iclamp(iz, 0, D - 1), z1 = iclamp(iz + 1, 0, D - 1);
    // the corner shared by texels (ix, ix+1) x (iy, iy+1): an exact coordinate, so the footprint is unambiguous
    float gx = fx + 1.0f, gy = fy + 1.0f;
    float4 a = gather_layer(obj, gx, gy, z0);
    float4 b = gather_layer(obj, gx, gy, z1);
    float c00 = lerp1(a.w, a.z, ax);
    float c10 = lerp1(a.x, a.y, ax);
    float c01 = lerp1(b.w, b.z, ax);
    float c11 = lerp1(b.x, b.y, ax);
    float c0  = lerp1(c00, c10, ay);
    float c1  = lerp1(c01, c11, ay);
    return lerp1(c0, c1, az);
}

template <>
struct SdfSampler<true>
{
    cudaTextureObject_t tex, mip;
    int                 tw, th, mw, mh;
    __device__ __forceinline__ SdfSampler(const TraceParams& P)
        : tex(P.texObj), mip(P.mipObj), tw(P.res * P.cascades), th(P.res), mw(P.mipRes * P.cascades), mh(P.mipRes) {}
    __device__ __forceinline__ float sampleTex(float u, float v, float w) const { return sample3D_tex(tex, tw, th, th, u, v, w); }
    __device__ __forceinline__ float sampleMip(float u, float v, float w) const { return sample3D_tex(mip, mw, mh, mh, u, v, w); }
};

// tracyGlobalSDF (SDFCommon.glsl:98-194) in full generality, for its other users (row f4): shadow, reflection and surface-cache light rays.
template <bool TEX>
__device__ __forceinline__ LuxGlobalSDFHit trace_global_sdf_general(const TraceParams& P, const SdfSampler<TEX>& sdf, const LuxGlobalSDFTrace& tr,
                                                                     float cascadeTraceStartBias)
{
    const LuxGlobalSDFData& data = P.sdf;
    const f3 origin = {tr.worldPosition[0], tr.worldPosition[1], tr.worldPosition[2]}, dir = {tr.worldDirection[0], tr.worldDirection[1], tr.worldDirection[2]};
    LuxGlobalSDFHit hit;
    hit.hitNormal[0] = hit.hitNormal[1] = hit.hitNormal[2] = 0.0f;
    hit.hitTime = -1.0f; hit.hitCascade = 0; hit.stepsCount = 0; hit.hitSDF = 0.0f;

    const float traceMaxDistance    = gmin(tr.maxDistance, data.cascadePosDistance[data.cascadesCount - 1][3] * 2.0f);
    const float chunkSizeDistance   = __fdiv_rn((float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE, data.resolution);
    const float chunkMarginDistance = __fdiv_rn((float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_MARGIN, data.resolution);
    const float cascadesCountF      = (float)data.cascadesCount;
    float nextIntersectionStart = 0.0f;
    const f3 traceEnd = origin + dir * traceMaxDistance;
    for (uint32_t cascade = 0; cascade < data.cascadesCount && hit.hitTime < 0.0f; cascade++)
    {
        f3    c         = {data.cascadePosDistance[cascade][0], data.cascadePosDistance[cascade][1], data.cascadePosDistance[cascade][2]};
        float cd        = data.cascadePosDistance[cascade][3];
        float voxelSize = data.cascadeVoxelSize[cascade];
        float voxelHalf = voxelSize * 0.5f;
        f3    worldPosition = origin + dir * (voxelSize * cascadeTraceStartBias);
        f3    ext = {cd, cd, cd};
        float nearT, farT;
        line_hit_aabb(worldPosition, traceEnd, c - ext, c + ext, nearT, farT);
        nearT *= traceMaxDistance;
        farT *= traceMaxDistance;
        nearT = gmax(nearT, nextIntersectionStart);
        float stepTime = nearT;
        if (nearT >= farT)
            stepTime = farT;
        else
            nextIntersectionStart = farT;
        const float cascadeMaxDistance = cd * 2.0f;
        uint32_t step = 0;
        for (; step < LUX_GLOBAL_SDF_MAX_STEPS && stepTime < farT; step++)
        {
            f3 stepPosition = worldPosition + dir * stepTime;
            f3 pc           = stepPosition - c;
            f3 cuv = {gclamp(__fdiv_rn(pc.x, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f), gclamp(__fdiv_rn(pc.y, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f),
                      gclamp(__fdiv_rn(pc.z, cascadeMaxDistance) + 0.5f, 0.0f, 1.0f)};
            f3 uvw = {__fdiv_rn((float)cascade + cuv.x, cascadesCountF), cuv.y, cuv.z};
            float stepDistance = sdf.sampleMip(uvw.x, uvw.y, uvw.z);
            if (stepDistance < chunkSizeDistance)
            {
                float stepDistanceTex = sdf.sampleTex(uvw.x, uvw.y, uvw.z);
                if (stepDistanceTex < chunkMarginDistance * 2.0f)
                    stepDistance = stepDistanceTex;
            }
            else
                stepDistance = chunkSizeDistance;
            stepDistance *= cascadeMaxDistance;
            float minSurfaceThickness = voxelHalf * gclamp(__fdiv_rn(stepTime, voxelSize), 0.0f, 1.0f);
            if (stepDistance < minSurfaceThickness)
            {
                hit.hitTime    = gmax((stepTime + stepDistance) - minSurfaceThickness, 0.0f);
                hit.hitCascade = cascade;
                hit.hitSDF     = stepDistance;
                if (tr.needsHitNormal)
                {
                    float o  = __fdiv_rn(1.0f, data.resolution);
                    float xp = sdf.sampleTex(uvw.x + o, uvw.y, uvw.z), xn = sdf.sampleTex(uvw.x - o, uvw.y, uvw.z);
                    float yp = sdf.sampleTex(uvw.x, uvw.y + o, uvw.z), yn = sdf.sampleTex(uvw.x, uvw.y - o, uvw.z);
                    float zp = sdf.sampleTex(uvw.x, uvw.y, uvw.z + o), zn = sdf.sampleTex(uvw.x, uvw.y, uvw.z - o);
                    f3 n = normalize3({xp - xn, yp - yn, zp - zn});
                    hit.hitNormal[0] = n.x; hit.hitNormal[1] = n.y; hit.hitNormal[2] = n.z;
                }
                break;
            }
            stepTime += gmax(stepDistance * tr.stepScale, voxelSize);
        }
        hit.stepsCount += step;
    }
    return hit;
}

// one thread per ray of a caller-provided list (lux_ddgi_trace_global_sdf)
template <bool TEX>
__global__ void __launch_bounds__(128) sdf_rays_kernel(const __grid_constant__ TraceParams P, int count, const LuxGlobalSDFTrace* __restrict__ traces,
                                                       float cascadeTraceStartBias, LuxGlobalSDFHit* __restrict__ hits)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    const SdfSampler<TEX> sdf(P);
    const LuxGlobalSDFTrace tr = traces[k];
    hits[k] = trace_global_sdf_general<TEX>(P, sdf, tr, cascadeTraceStartBias);
}

// One tile of one candidate object, split in two: the normal weight (cheap, rejects most tiles) ...
__device__ __forceinline__ void load_tile_transform(const LuxTileBuffer* __restrict__ tile, float* tm)
{
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float4 c = __ldg(reinterpret_cast<const float4*>(tile->transform) + i);
        tm[i * 4 + 0] = c.x; tm[i * 4 + 1] = c.y; tm[i * 4 + 2] = c.z; tm[i * 4 + 3] = c.w;
    }
}
__device__ __forceinline__ float tile_normal_weight(const LuxTileBuffer* __restrict__ tile, f3 normal, float* tm)
{
    load_tile_transform(tile, tm);
    f3    nt = normalize3(mat4_mul_point(tm, normal, 1.0f));
    float nw = gclamp(nt.z, 0.0f, 1.0f);
    return __fdiv_rn(nw - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD, 1.0f - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD);
}

// ... and the depth-tested bilinear sample (AtlasCommon.glsl:62-96)
__device__ __forceinline__ f4 tile_sample(const TraceParams& P, const LuxTileBuffer* __restrict__ tile, const float* tm,
                                          f3 localPosition, float normalWeight, float surfaceThreshold)
{
    float4 ext = __ldg(reinterpret_cast<const float4*>(tile->extends));
    float4 ob  = __ldg(reinterpret_cast<const float4*>(tile->objectBounds));
    f3     tp  = mat4_mul_point(tm, localPosition, 1.0f);
    float  tileDepth = __fdiv_rn(tp.z, ob.z);
    float  tu = gclamp(__fdiv_rn(tp.x, ob.x) + 0.5f, 0.0f, 1.0f);
    float  tv = gclamp(__fdiv_rn(tp.y, ob.y) + 0.5f, 0.0f, 1.0f);
    float  au = tu * ext.z + ext.x, av = tv * ext.w + ext.y;
    float  res = (float)P.atlasRes;
    float  fx = gfract(au * res + 0.5f), fy = gfract(av * res + 0.5f);
    f4     bw = {(1.0f - fx) * fy, fx * fy, fx * (1.0f - fy), (1.0f - fx) * (1.0f - fy)};

    int R = (int)P.atlasRes, i0, i1, j0, j1;
    gather_coords(au, av, R, R, false, i0, i1, j0, j1);
    float z4[4] = {__ldg(P.depth + (size_t)j1 * R + i0), __ldg(P.depth + (size_t)j1 * R + i1), __ldg(P.depth + (size_t)j0 * R + i1),
                   __ldg(P.depth + (size_t)j0 * R + i0)};
    float depthThreshold = __fdiv_rn(2.0f * surfaceThreshold, ob.z);
    float vis[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        vis[i] = 1.0f - gclamp(__fdiv_rn(fabsf(tileDepth - z4[i]) - depthThreshold, 0.5f * depthThreshold), 0.0f, 1.0f);
        if (z4[i] >= 1.0f)
            vis[i] = 0.0f;
    }
    f4    visv = {vis[0], vis[1], vis[2], vis[3]};
    float sampleWeight = dot4(visv, bw);
    sampleWeight *= normalWeight;
    if (sampleWeight <= 0.0f)
        return {0.0f, 0.0f, 0.0f, 0.0f};
    bw = {bw.x * visv.x, bw.y * visv.y, bw.z * visv.z, bw.w * visv.w};
    gather_coords(au, av, R, R, true, i0, i1, j0, j1);
    f4 t0 = unpack_rgba16f(__ldg(P.light + (size_t)j1 * R + i0));
    f4 t1 = unpack_rgba16f(__ldg(P.light + (size_t)j1 * R + i1));
    f4 t2 = unpack_rgba16f(__ldg(P.light + (size_t)j0 * R + i1));
    f4 t3 = unpack_rgba16f(__ldg(P.light + (size_t)j0 * R + i0));
    float cr = dot4({t0.x, t1.x, t2.x, t3.x}, bw);
    float cg = dot4({t0.y, t1.y, t2.y, t3.y}, bw);
    float cb = dot4({t0.z, t1.z, t2.z, t3.z}, bw);
    return {cr * sampleWeight, cg * sampleWeight, cb * sampleWeight, sampleWeight};
}

__device__ __forceinline__ void load_object_inverse(const TraceParams& P, uint32_t objectAddress, float* wl)
{
    const float4* invp = reinterpret_cast<const float4*>(P.objectInverse + (size_t)objectAddress * 16);
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float4 c = __ldg(invp + i);
        wl[i * 4 + 0] = c.x; wl[i * 4 + 1] = c.y; wl[i * 4 + 2] = c.z; wl[i * 4 + 3] = c.w;
    }
}

// AtlasCommon.glsl:115-157 in scan-then-shade form.  `cand` is this lane's candidate scratch in shared memory.
__device__ __forceinline__ f4 sample_global_surface_atlas_2p(const TraceParams& P, f3 worldPosition, f3 worldNormal,
                                                             float surfaceThreshold, uint32_t* cand)
{
    f4 result = {0.0f, 0.0f, 0.0f, 0.0f};
    if (!P.hasAtlas)
        return result;
    const float half = (float)LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * 0.5f;
    const int   N    = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    int cx = iclamp((int)floorf(__fdiv_rn(worldPosition.x, P.chunkSize) + half), 0, N - 1);
    int cy = iclamp((int)floorf(__fdiv_rn(worldPosition.y, P.chunkSize) + half), 0, N - 1);
    int cz = iclamp((int)floorf(__fdiv_rn(worldPosition.z, P.chunkSize) + half), 0, N - 1);
    uint32_t objectsStart = __ldg(P.chunks + (cz * N * N + cy * N + cx));
    if (objectsStart == 0)
        return result;
    uint32_t objectsCount = __ldg(P.cull + objectsStart);
    if (objectsCount > P.objectsCount)
        return result;
    objectsStart++;
    // Conservative prefilter (chunk_masks_kernel): bit j of the sub-cell mask is clear only if list entry j < 64 cannot pass
    // the sphere or the OBB test anywhere in this quarter-chunk cell, so skipping it never changes the result.
    unsigned long long mask = ~0ull;
    if (P.chunkMasks)
    {
        const float q = __fdiv_rn(4.0f, P.chunkSize);
        int sx = iclamp((int)floorf((worldPosition.x - ((float)cx - half) * P.chunkSize) * q), 0, 3);
        int sy = iclamp((int)floorf((worldPosition.y - ((float)cy - half) * P.chunkSize) * q), 0, 3);
        int sz = iclamp((int)floorf((worldPosition.z - ((float)cz - half) * P.chunkSize) * q), 0, 3);
        mask = __ldg(P.chunkMasks + (size_t)(cz * N * N + cy * N + cx) * 64 + (sz * 16 + sy * 4 + sx));
    }
    uint32_t k = 0;
    while (k < objectsCount)
    {
        // ---- scan: bounding sphere + OBB tests only, in list order ----
        int nc = 0;
        while (k < objectsCount && nc < TW_MAX_CAND)
        {
            if (k < 64)
            { // jump to the next list position the prefilter kept
                unsigned long long rest = mask >> k;
                if (rest == 0ull)
                {
                    k = 64;
                    continue;
                }
                k += __ffsll((long long)rest) - 1;
                if (k >= objectsCount)
                    break;
            }
            uint32_t objectAddress = __ldg(P.cull + objectsStart + k);
            k++;
            const LuxObjectBuffer* object = P.objects + objectAddress;
            float4 ob = __ldg(reinterpret_cast<const float4*>(object->objectBounds));
            f3     bc = {ob.x, ob.y, ob.z};
            if (length3(bc - worldPosition) > ob.w)
                continue;
            float wl[16];
            load_object_inverse(P, objectAddress, wl);
            f3     lp = mat4_mul_point(wl, worldPosition, 1.0f);
            float4 ex = __ldg(reinterpret_cast<const float4*>(object->extends));
            if (fabsf(lp.x) > ex.x + surfaceThreshold || fabsf(lp.y) > ex.y + surfaceThreshold || fabsf(lp.z) > ex.z + surfaceThreshold)
                continue;
            cand[(nc++) * CAND_STRIDE] = objectAddress;
        }
        // ---- shade the candidates, tiles in order ----
        for (int c = 0; c < nc; c++)
        {
            uint32_t objectAddress = cand[c * CAND_STRIDE];
            const LuxObjectBuffer* object = P.objects + objectAddress;
            float wl[16];
            load_object_inverse(P, objectAddress, wl);
            f3 localPosition = mat4_mul_point(wl, worldPosition, 1.0f);
            f3 normal        = normalize3(mat3_mul(wl, worldNormal));
            // pass A: normal weights of the six tiles (cheap; most tiles face away).  pass B: lanes walk their passing tiles
            // together, so the expensive depth-tested sample is entered by many lanes at once.  Order is preserved.
            uint32_t passing = 0;
#pragma unroll 1
            for (int i = 0; i < 6; i++)
            {
                uint32_t tileOffset = __ldg(object->tileOffset + i);
                if (tileOffset == 0)
                    continue;
                if (P.tileZRow)
                { // Exact early reject: the weight is positive only if the z component of the transformed normal is; its sign is that
                  // of the un-normalised z (normalize multiplies by a positive number, NaN compares false like the full path), and the
                  // expression below is character for character the r.z of mat4_mul_point.  Three of a box's six faces leave here.
                    float4 zr = __ldg(P.tileZRow + tileOffset);
                    float  rz = ((zr.x * normal.x + zr.y * normal.y) + zr.z * normal.z) + zr.w * 1.0f;
                    if (!(rz > 0.0f))
                        continue;
                }
                float tm[16];
                float nw = tile_normal_weight(P.tiles + tileOffset, normal, tm);
                if (nw > 0.0f)
                {
                    passing |= 1u << i;
                    cand[(TW_MAX_CAND + i) * CAND_STRIDE] = __float_as_uint(nw); // reused by pass B
                }
            }
            while (passing)
            {
                int i = __ffs(passing) - 1;
                passing &= passing - 1;
                const LuxTileBuffer* tile = P.tiles + __ldg(object->tileOffset + i);
                float tm[16];
                load_tile_transform(tile, tm);
                float nw = __uint_as_float(cand[(TW_MAX_CAND + i) * CAND_STRIDE]);
                f4 s = tile_sample(P, tile, tm, localPosition, nw, surfaceThreshold);
                result.x += s.x; result.y += s.y; result.z += s.z; result.w += s.w;
            }
        }
    }
    float d = gmax(result.w, 0.0001f);
    result.x = __fdiv_rn(result.x, d);
    result.y = __fdiv_rn(result.y, d);
    result.z = __fdiv_rn(result.z, d);
    return result;
}

__device__ __forceinline__ f3 probe_origin(const TraceParams& P, int probeId)
{
    int cx = probeId % P.countX;
    int cy = (probeId % (P.countX * P.countY)) / P.countX;
    int cz = probeId / (P.countX * P.countY);
    return {P.step[0] * (float)cx + P.start[0], P.step[1] * (float)cy + P.start[1], P.step[2] * (float)cz + P.start[2]};
}

// ---------------------------------------------------------------------------------------------------------------------
// Stage 1 of the wavefront trace: MARCH (march_kernel.inc).  Persistent warps pull chunks of rays from a global counter; lanes refill
// from the warp's chunk in batches and every loop iteration is one sphere-trace step for all active lanes.  A finished ray leaves a 20-byte
// record (hit time, hit uvw, cascade, kind, steps) and, if it hit, a ticket in the shade's counting sort; nothing is shaded here.
// ---------------------------------------------------------------------------------------------------------------------
// Occupancy: 8 warps x 4 blocks = 32 warps per SM at 64 registers.  ptxas keeps the four gathers of a step together from 56 registers up (at 48
// it serialises the two taps to save registers, SASS inspected); measured on B200 (profiles/r2_march_occupancy.md): 48 / 56 / 64 registers =
// 4.12 / 4.16 / 3.99 ms on C4 and 91.8 / 94.5 / 91.2 ms on C5.
#ifndef MARCH_WARPS_N
#define MARCH_WARPS_N 8
#endif
#ifndef MARCH_BLOCKS_PER_SM
#define MARCH_BLOCKS_PER_SM 4
#endif
#ifndef MARCH_BATCH_N
#define MARCH_BATCH_N 16
#endif
constexpr int MARCH_WARPS = MARCH_WARPS_N;
#ifndef SHADE_BLOCKS_PER_SM
#define SHADE_BLOCKS_PER_SM 5
#endif
constexpr int MARCH_CHUNK_RAYS  = 64;                                      // rays per pool fetch (2 probes x 32 directions): small, so
                                                                           // that shards with few rays per warp still balance
// refill when this many lanes are idle.  Measured on B200 (profiles/r2_experiments.md): rows (C5) 6 / 8 / 12 / 16 / 20 / 24 idle lanes = 82.3 / 81.4 / 80.4 /
// 79.8 / 82.7 / 86.2 ms, beams (C4) 3.26 / 3.21 / 3.17 / 3.18 / 3.30 / 3.60 ms: a refill (flush of the parked records, tickets, pool bookkeeping) costs
// about as many instructions as a step, so fewer and fuller refills win until the idle lanes outweigh them
constexpr int MARCH_REFILL_MIN_ROWS = 16, MARCH_REFILL_MIN_BEAMS = 12;
constexpr unsigned int MARCH_BATCH       = MARCH_BATCH_N;                             // consecutive chunks a block draws at a time (rows: one unit x one cluster)
constexpr unsigned int MARCH_BATCH_SLOTS = 32;                             // published batch ids kept per block (a waiter reads its slot at once)

enum RayKind : uint32_t { RAY_MISS = 0, RAY_INSIDE = 1, RAY_HIT = 2 };

constexpr int SORT_CELLS_PER_CHUNK = 64;
constexpr unsigned int SHADE_SM_GROUP = 8, SHADE_POOL_SLOTS = 32, SHADE_POOL_SMS = 256; // see shade_sorted_kernel
// Bin of a hit position.  Only groups work, never enters a result: plain (approximate) arithmetic is fine.
__device__ __forceinline__ uint32_t shade_bin(const TraceParams& P, f3 pos)
{
    const int   N    = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    const float half = (float)N * 0.5f, inv = P.invChunkSize;
    float fx = pos.x * inv + half, fy = pos.y * inv + half, fz = pos.z * inv + half;
    float gx = floorf(fx), gy = floorf(fy), gz = floorf(fz);
    int cx = iclamp((int)gx, 0, N - 1), cy = iclamp((int)gy, 0, N - 1), cz = iclamp((int)gz, 0, N - 1);
    // sub-cell of the chunk = the 4 x 4 x 4 grid of the prefilter masks (chunk_masks_kernel): the hits of a bin share the candidate mask
    int sx = iclamp((int)((fx - gx) * 4.0f), 0, 3), sy = iclamp((int)((fy - gy) * 4.0f), 0, 3), sz = iclamp((int)((fz - gz) * 4.0f), 0, 3);
    return (uint32_t)(((cz * N + cy) * N + cx) * SORT_CELLS_PER_CHUNK + ((sz * 4 + sy) * 4 + sx));
}

// MARCH ORDER (TraceParams::beam).  Chunk c (64 records = [j][lane]) <-> (probe unit, direction cluster[, direction pair]); record index =
// c * 64 + j * 32 + lane.  Everything downstream of the march (classify, scatter, shade) addresses records through these functions.
struct MarchChunk { uint32_t probe0, slot0; }; // first probe (shard-local id) and first ray slot of the chunk
__device__ __forceinline__ void march_split(const TraceParams& P, unsigned int q, unsigned int& cluster, unsigned int& unitIdx)
{
    if (P.probeMajor) { cluster = q % (unsigned)P.rayClusters; unitIdx = q / (unsigned)P.rayClusters; }
    else              { unitIdx = q % (unsigned)P.probeUnits; cluster = q / (unsigned)P.probeUnits; }
}
__device__ __forceinline__ unsigned int march_join(const TraceParams& P, unsigned int cluster, unsigned int unitIdx)
{
    return P.probeMajor ? unitIdx * (unsigned)P.rayClusters + cluster : cluster * (unsigned)P.probeUnits + unitIdx;
}
__device__ __forceinline__ MarchChunk march_chunk(const TraceParams& P, unsigned int c)
{
    unsigned int cluster, unitIdx;
    if (P.beam)
    {
        march_split(P, c, cluster, unitIdx);
        return {__ldg(P.unitOrder + unitIdx) * 2u, cluster * MARCH_CLUSTER_RAYS};
    }
    march_split(P, c / (MARCH_CLUSTER_RAYS / 2), cluster, unitIdx);
    return {__ldg(P.unitOrder + unitIdx) * 32u, cluster * MARCH_CLUSTER_RAYS + (c % (MARCH_CLUSTER_RAYS / 2)) * 2u};
}
// record index of (shard-local probe, ray id)
__device__ __forceinline__ uint32_t march_record(const TraceParams& P, uint32_t probeLocal, uint32_t rayId)
{
    const unsigned int slot = __ldg(P.raySlot + rayId), cluster = slot / MARCH_CLUSTER_RAYS, s = slot % MARCH_CLUSTER_RAYS;
    if (P.beam)
        return march_join(P, cluster, __ldg(P.unitIndex + (probeLocal >> 1))) * 64u + (probeLocal & 1u) * 32u + s;
    const unsigned int c = march_join(P, cluster, __ldg(P.unitIndex + (probeLocal >> 5))) * (MARCH_CLUSTER_RAYS / 2) + (s >> 1);
    return c * 64u + (s & 1u) * 32u + (probeLocal & 31u);
}
struct RayOfRecord { int probeLocal, rayId; };
__device__ __forceinline__ RayOfRecord march_ray(const TraceParams& P, uint32_t g)
{
    const MarchChunk mc = march_chunk(P, g >> 6);
    const uint32_t j = (g >> 5) & 1u, lane = g & 31u;
    return {(int)(mc.probe0 + (P.beam ? j : lane)), (int)__ldg(P.rayOrder + mc.slot0 + (P.beam ? lane : j))};
}

#include "march_kernel.inc"

// ---------------------------------------------------------------------------------------------------------------------
// Stage 2: SHADE.  One thread per ray record, in the same [32 probes] x [8 rays] tiling as the simple kernel, so a warp
// holds hits of parallel rays from adjacent probes (same building face, same culling chunk most of the time).
// Misses take the sky, inside-geometry rays are black, hits get their normal (six taps) and surface-cache radiance.
// ---------------------------------------------------------------------------------------------------------------------
template <bool TEX>
__global__ void __launch_bounds__(256, SHADE_BLOCKS_PER_SM) shade_kernel(const __grid_constant__ TraceParams P)
{
    __shared__ uint32_t sCand[TW_MAX_CAND + 6][CAND_STRIDE]; // candidates + the six tile normal weights
    // one thread per ray in [probe][ray] order: consecutive threads write consecutive rays of a probe
    const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
    if (item >= (long long)P.probeCount * P.raysPerProbe)
        return;
    const int probeLocal = (int)(item / P.raysPerProbe), rayId = (int)(item % P.raysPerProbe);
    const uint32_t g = march_record(P, (uint32_t)probeLocal, (uint32_t)rayId);
    const SdfSampler<TEX> sdf(P);
    const LuxGlobalSDFData& data = P.sdf;

    const float4   rec  = __ldg(P.records + g);
    const uint32_t meta = __ldg(P.meta + g);
    const uint32_t hc = meta & 3u, kind = (meta >> 2) & 3u;
    float4 d4 = __ldg(P.dirs + rayId);
    f3     d  = {d4.x, d4.y, d4.z};
    f4     radiance;
    if (kind == RAY_HIT)
    {
        float4 o4 = __ldg(P.origins + probeLocal);
        f3     o  = {o4.x, o4.y, o4.z};
        const float texelOffset = __fdiv_rn(1.0f, data.resolution);
        float xp = sdf.sampleTex(rec.y + texelOffset, rec.z, rec.w);
        float xn = sdf.sampleTex(rec.y - texelOffset, rec.z, rec.w);
        float yp = sdf.sampleTex(rec.y, rec.z + texelOffset, rec.w);
        float yn = sdf.sampleTex(rec.y, rec.z - texelOffset, rec.w);
        float zp = sdf.sampleTex(rec.y, rec.z, rec.w + texelOffset);
        float zn = sdf.sampleTex(rec.y, rec.z, rec.w - texelOffset);
        f3    normal = normalize3({xp - xn, yp - yn, zp - zn});
        f3    hitPosition      = o + d * rec.x;
        float surfaceThreshold = data.cascadeVoxelSize[hc] * 1.05f;
        f4    sc = sample_global_surface_atlas_2p(P, hitPosition, normal, surfaceThreshold, &sCand[0][threadIdx.x]);
        radiance   = {sc.x, sc.y, sc.z, rec.x};
        radiance.w = gmax(radiance.w + data.cascadeVoxelSize[hc] * 0.5f, 0.0f);
    }
    else if (kind == RAY_INSIDE)
        radiance = {0.0f, 0.0f, 0.0f, LUX_GLOBAL_SDF_WORLD_SIZE};
    else
    {
        f3 s     = sample_sky(P, d);
        radiance = {s.x, s.y, s.z, LUX_GLOBAL_SDF_WORLD_SIZE};
    }
    uint32_t r0 = f2h_bits(radiance.x), r1 = f2h_bits(radiance.y), r2 = f2h_bits(radiance.z);
    uint32_t d0 = f2h_bits(d.x), d1 = f2h_bits(d.y), d2 = f2h_bits(d.z), d3 = f2h_bits(radiance.w);
    P.radiance[item] = make_uint2(r0 | (r1 << 16), r2);
    P.dirDist[item]  = make_uint2(d0 | (d1 << 16), d2 | (d3 << 16));
    flag_non_finite(P, make_uint2(r0 | (r1 << 16), r2 | (d3 << 16)));
    if (P.steps)
        P.steps[item] = (uint16_t)(meta >> 4);
}

// ---------------------------------------------------------------------------------------------------------------------
// Stage 2, sorted form (default): CLASSIFY, SCAN -> SCATTER -> SHADE_SORTED.
//
// In ray order a warp holds hits of 32 parallel rays from a 240-voxel-long row of probes: they fall into ~12 different
// culling chunks, so the surface-cache loops ran with 13-17 of 32 lanes (ncu, profiles/r1_v5_*).  The sorted form bins every
// hit by (culling chunk, 4 x 4 x 4 sub-cell of the chunk = the cell of the prefilter masks) with a counting sort and shades in bin order: the lanes of a warp then walk the SAME
// culled object list with the same prefilter mask and mostly the same tiles, and the whole pass reads each part of the SDF
// shell and of the surface-cache atlases once (one sweep over the scene; round 1 keyed the sort by 4 M-record output windows
// first and swept the scene once per window: 105 GB of DRAM traffic on C5).  Per-ray arithmetic is untouched (each ray's
// result depends on no other ray), so results stay bit-identical to the ray-order kernel.
//   march         every hit takes a ticket in its bin when its record is written (one returning atomic, hidden by the march)
//   classify      one thread per record, ray-order tiling: writes direction+distance for every ray and the final radiance of
//                 misses / inside rays (coalesced 128-byte row segments)
//   scan          exclusive prefix sum over the 4 096 000 bins (3 small kernels)
//   scatter       sortedIdx[prefix[bin] + ticket] = record index; streams the tickets in march order, in which the hits of a bin
//                 that were ticketed together also sit together, so the 4-byte stores of a sector meet in L2
//   shade_sorted  persistent grid over the sorted hit list: normal (6 taps) + surface cache; 8-byte radiance store per hit.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SORT_BINS             = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * SORT_CELLS_PER_CHUNK;
constexpr int SCAN_THREADS          = 1024;
constexpr int SCAN_BINS_PER_BLOCK   = SCAN_THREADS * 4;

// One thread per ray in [probe][ray] order, CLASSIFY_RPT rays per thread with all loads issued before the first use.  Reads are a gather (a
// probe's rays sit in its clusters' records, 32 at a time), but a block covers whole rows of a probe, so every sector it touches is used completely
// while it is in L1 / L2; writes are consecutive rays of a probe.  A streaming pass: 20 bytes in, 16 bytes out per ray.
constexpr int CLASSIFY_RPT = 4;
__global__ void __launch_bounds__(256) classify_kernel(const __grid_constant__ TraceParams P)
{
    const LuxGlobalSDFData& data = P.sdf;
    const long long total = (long long)P.probeCount * P.raysPerProbe;
    const long long base  = (long long)blockIdx.x * (256 * CLASSIFY_RPT) + threadIdx.x;
    float4   rec[CLASSIFY_RPT];
    uint32_t meta[CLASSIFY_RPT];
    int      ray[CLASSIFY_RPT];
#pragma unroll
    for (int k = 0; k < CLASSIFY_RPT; k++)
    {
        const long long item = base + k * 256;
        const bool valid = item < total;
        const int probeLocal = valid ? (int)(item / P.raysPerProbe) : 0;
        ray[k] = valid ? (int)(item % P.raysPerProbe) : -1;
        const uint32_t g = valid ? march_record(P, (uint32_t)probeLocal, (uint32_t)ray[k]) : 0u;
        rec[k]  = valid ? __ldg(P.records + g) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        meta[k] = valid ? __ldg(P.meta + g) : 0u;
    }
#pragma unroll
    for (int k = 0; k < CLASSIFY_RPT; k++)
    {
        if (ray[k] < 0)
            continue;
        const long long item = base + k * 256;
        const float4 d4 = __ldg(P.dirs + ray[k]);
        const uint32_t hc = meta[k] & 3u, kind = (meta[k] >> 2) & 3u;
        f3 d = {d4.x, d4.y, d4.z};
        f4 radiance;
        if (kind == RAY_HIT) // rgb comes from the shade kernel (zero without a surface cache)
            radiance = {0.0f, 0.0f, 0.0f, gmax(rec[k].x + data.cascadeVoxelSize[hc] * 0.5f, 0.0f)};
        else if (kind == RAY_INSIDE)
            radiance = {0.0f, 0.0f, 0.0f, LUX_GLOBAL_SDF_WORLD_SIZE};
        else
        {
            f3 s     = sample_sky(P, d);
            radiance = {s.x, s.y, s.z, LUX_GLOBAL_SDF_WORLD_SIZE};
        }
        uint32_t r0 = f2h_bits(radiance.x), r1 = f2h_bits(radiance.y), r2 = f2h_bits(radiance.z);
        uint32_t d0 = f2h_bits(d.x), d1 = f2h_bits(d.y), d2 = f2h_bits(d.z), d3 = f2h_bits(radiance.w);
        P.radiance[item] = make_uint2(r0 | (r1 << 16), r2);
        P.dirDist[item]  = make_uint2(d0 | (d1 << 16), d2 | (d3 << 16));
        flag_non_finite(P, make_uint2(r0 | (r1 << 16), r2 | (d3 << 16)));
        if (P.steps)
            P.steps[item] = (uint16_t)(meta[k] >> 4);
    }
}

// The same pass for ROW chunks (TraceParams::beam == 0): there the 32 lanes of a record row are 32 x-adjacent probes of one ray, so a block takes
// 32 probes x 16 consecutive ray ids, reads whole 512-byte record rows and transposes through shared memory to write 128-byte row segments.
#ifndef CLASSIFY_ROW_RAYS_N
#define CLASSIFY_ROW_RAYS_N 64
#define CLASSIFY_ROW_WARPS_N 8
#define CLASSIFY_ROW_BLOCKS_N 4
#endif
constexpr int CLASSIFY_ROW_RAYS    = CLASSIFY_ROW_RAYS_N;  // ray ids per block: a probe's output is one 512-byte segment per buffer (16 rays = 128-byte bursts ran at 62 % of the copy bandwidth)
constexpr int CLASSIFY_ROW_WARPS   = CLASSIFY_ROW_WARPS_N;
constexpr int CLASSIFY_ROW_RPT     = CLASSIFY_ROW_RAYS / CLASSIFY_ROW_WARPS; // record rows per warp
__global__ void __launch_bounds__(32 * CLASSIFY_ROW_WARPS, CLASSIFY_ROW_BLOCKS_N) classify_rows_kernel(const __grid_constant__ TraceParams P, int rayGroups)
{
    __shared__ uint2 sRad[32][CLASSIFY_ROW_RAYS + 1];
    __shared__ uint2 sDir[32][CLASSIFY_ROW_RAYS + 1];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = blockIdx.x;
    const uint32_t probeGroup = (uint32_t)(unit / rayGroups);
    const int probeLocal = (int)probeGroup * 32 + lane;
    const int rayBase    = (int)(unit % rayGroups) * CLASSIFY_ROW_RAYS;
    const LuxGlobalSDFData& data = P.sdf;
    const bool probeValid = probeLocal < P.probeCount;

    float4   rec[CLASSIFY_ROW_RPT];
    uint32_t meta[CLASSIFY_ROW_RPT];
    bool     valid[CLASSIFY_ROW_RPT];
#pragma unroll
    for (int k = 0; k < CLASSIFY_ROW_RPT; k++)
    {
        const int rayId = rayBase + warp + k * CLASSIFY_ROW_WARPS;
        valid[k] = probeValid && rayId < P.raysPerProbe;
        const uint32_t g = valid[k] ? march_record(P, (uint32_t)probeLocal, (uint32_t)rayId) : 0u;
        rec[k]   = valid[k] ? __ldcs(P.records + g) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        meta[k]  = valid[k] ? __ldcs(P.meta + g) : 0u;
    }
#pragma unroll
    for (int k = 0; k < CLASSIFY_ROW_RPT; k++)
    {
        const int rayInUnit = warp + k * CLASSIFY_ROW_WARPS;
        if (valid[k])
        {
            const uint32_t hc = meta[k] & 3u, kind = (meta[k] >> 2) & 3u;
            const float4 d4 = __ldg(P.dirs + rayBase + rayInUnit); // the frame's directions: 16 KB, cache resident
            f3 d = {d4.x, d4.y, d4.z};
            f4 radiance;
            if (kind == RAY_HIT) // rgb comes from the shade kernel (zero without a surface cache)
                radiance = {0.0f, 0.0f, 0.0f, gmax(rec[k].x + data.cascadeVoxelSize[hc] * 0.5f, 0.0f)};
            else if (kind == RAY_INSIDE)
                radiance = {0.0f, 0.0f, 0.0f, LUX_GLOBAL_SDF_WORLD_SIZE};
            else
            {
                f3 s     = sample_sky(P, d);
                radiance = {s.x, s.y, s.z, LUX_GLOBAL_SDF_WORLD_SIZE};
            }
            uint32_t r0 = f2h_bits(radiance.x), r1 = f2h_bits(radiance.y), r2 = f2h_bits(radiance.z);
            uint32_t d0 = f2h_bits(d.x), d1 = f2h_bits(d.y), d2 = f2h_bits(d.z), d3 = f2h_bits(radiance.w);
            sRad[lane][rayInUnit] = make_uint2(r0 | (r1 << 16), r2);
            sDir[lane][rayInUnit] = make_uint2(d0 | (d1 << 16), d2 | (d3 << 16));
            flag_non_finite(P, make_uint2(r0 | (r1 << 16), r2 | (d3 << 16)));
            if (P.steps)
                P.steps[(size_t)probeLocal * P.raysPerProbe + rayBase + rayInUnit] = (uint16_t)(meta[k] >> 4);
        }
    }
    __syncthreads();
    // transposed write-out: 64 consecutive rays (512 bytes) per probe and buffer
#pragma unroll
    for (int k = 0; k < CLASSIFY_ROW_RPT; k++)
    {
        const int idx = threadIdx.x + k * (32 * CLASSIFY_ROW_WARPS);
        const int pl = idx / CLASSIFY_ROW_RAYS, rl = idx % CLASSIFY_ROW_RAYS;
        const int oProbe = (int)probeGroup * 32 + pl;
        const int oRay   = rayBase + rl;
        if (oProbe < P.probeCount && oRay < P.raysPerProbe)
        {
            size_t o = (size_t)oProbe * P.raysPerProbe + oRay;
            __stcs(P.radiance + o, sRad[pl][rl]);
            __stcs(P.dirDist + o, sDir[pl][rl]);
        }
    }
}

// exclusive scan of one value per thread over a 1024-thread block; returns the prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* sWarp, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
            inc += n;
    }
    if (lane == 31)
        sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = sWarp[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o)
                winc += n;
        }
        sWarp[lane] = winc - w;
        if (lane == 31)
            sWarp[32] = winc;
    }
    __syncthreads();
    uint32_t r = sWarp[warp] + inc - v;
    *total     = sWarp[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint4* __restrict__ counts, uint32_t* __restrict__ blockSums)
{
    __shared__ uint32_t sWarp[33];
    uint4    v = counts[(size_t)blockIdx.x * SCAN_THREADS + threadIdx.x];
    uint32_t total;
    block_exclusive_scan(v.x + v.y + v.z + v.w, sWarp, &total);
    if (threadIdx.x == 0)
        blockSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_top_kernel(uint32_t* __restrict__ blockSums, int nb, uint32_t* __restrict__ total)
{
    __shared__ uint32_t sWarp[33];
    uint32_t carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS)
    {
        int      i = base + threadIdx.x;
        uint32_t v = i < nb ? blockSums[i] : 0u, t;
        uint32_t e = block_exclusive_scan(v, sWarp, &t);
        if (i < nb)
            blockSums[i] = carry + e;
        carry += t;
    }
    if (threadIdx.x == 0)
        *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint4* __restrict__ counts, const uint32_t* __restrict__ blockSums)
{
    __shared__ uint32_t sWarp[33];
    const size_t i = (size_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    uint4    v = counts[i];
    uint32_t total;
    uint32_t b = blockSums[blockIdx.x] + block_exclusive_scan(v.x + v.y + v.z + v.w, sWarp, &total);
    counts[i]  = make_uint4(b, b + v.x, b + v.x + v.y, b + v.x + v.y + v.z);
}

constexpr int SCATTER_RPT = 8; // records per thread (four 16-byte loads of two tickets each): the pass is a chain ticket -> prefix -> store per record, so loads are batched
__global__ void __launch_bounds__(256) scatter_kernel(const uint2* __restrict__ ticket, const uint32_t* __restrict__ prefix,
                                                      uint32_t* __restrict__ sortedIdx, size_t records)
{
    // thread -> records g0 + k * 512, g0 + k * 512 + 1 (k = 0..3); the record capacity is a multiple of 64, so a pair never straddles the end
    const size_t g0 = (size_t)blockIdx.x * (256 * SCATTER_RPT) + (size_t)threadIdx.x * 2;
    uint4    t[SCATTER_RPT / 2];
    uint32_t base[SCATTER_RPT];
#pragma unroll
    for (int k = 0; k < SCATTER_RPT / 2; k++)
    {
        const size_t g = g0 + (size_t)k * 512;
        t[k] = g + 1 < records ? __ldcs(reinterpret_cast<const uint4*>(ticket + g)) : make_uint4(0xffffffffu, 0u, 0xffffffffu, 0u);
    }
#pragma unroll
    for (int k = 0; k < SCATTER_RPT / 2; k++)
    {
        base[2 * k]     = t[k].x != 0xffffffffu ? __ldg(prefix + t[k].x) : 0u;
        base[2 * k + 1] = t[k].z != 0xffffffffu ? __ldg(prefix + t[k].z) : 0u;
    }
#pragma unroll
    for (int k = 0; k < SCATTER_RPT / 2; k++)
    {
        const uint32_t g = (uint32_t)(g0 + (size_t)k * 512);
        if (t[k].x != 0xffffffffu)
            sortedIdx[base[2 * k] + (t[k].y & 0x3fffffffu)] = g | (t[k].y & 0xc0000000u);
        if (t[k].z != 0xffffffffu)
            sortedIdx[base[2 * k + 1] + (t[k].w & 0x3fffffffu)] = (g + 1u) | (t[k].w & 0xc0000000u);
    }
}

template <bool TEX>
__global__ void __launch_bounds__(256, SHADE_BLOCKS_PER_SM) shade_sorted_kernel(const __grid_constant__ TraceParams P)
{
    __shared__ uint32_t sCand[TW_MAX_CAND + 6][CAND_STRIDE]; // candidates + the six tile normal weights
    const SdfSampler<TEX> sdf(P);
    const LuxGlobalSDFData& data = P.sdf;
    const uint32_t hits = *P.hitCount;
    const float texelOffset = __fdiv_rn(1.0f, data.resolution);
    // Segments of 256 consecutive hits are handed out per SM: the blocks resident on one SM draw tickets from that SM's counter (in global memory, indexed
    // by %smid) and every SHADE_SM_GROUP consecutive tickets share one group of adjacent segments fetched from the global counter, so co-resident blocks
    // shade neighbouring hits - same bins, same objects, tiles and SDF neighbourhood in L1 - instead of segments 148 x 256 hits apart (static striding).
    __shared__ uint32_t sSegment;
    unsigned int smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* pool = P.shadePools + (size_t)(smid % SHADE_POOL_SMS) * (1 + SHADE_POOL_SLOTS);
    while (true)
    {
        if (threadIdx.x == 0)
        {
            const unsigned int tk = (unsigned int)atomicAdd(pool, 1ull), seq = tk / SHADE_SM_GROUP, slot = seq % SHADE_POOL_SLOTS;
            unsigned int grp;
            if (tk % SHADE_SM_GROUP == 0)
            {
                grp = atomicAdd(P.shadeCounter, 1u);
                atomicExch(pool + 1 + slot, ((unsigned long long)(seq + 1u) << 32) | grp);
            }
            else
            {
                unsigned long long v;
                while ((unsigned int)((v = atomicAdd(pool + 1 + slot, 0ull)) >> 32) != seq + 1u)
                    __nanosleep(40);
                grp = (unsigned int)v;
            }
            sSegment = grp * SHADE_SM_GROUP + tk % SHADE_SM_GROUP;
        }
        __syncthreads();
        const unsigned long long t64 = (unsigned long long)sSegment * 256ull + threadIdx.x;
        __syncthreads();
        if (t64 - threadIdx.x >= hits)
            break;
        if (t64 >= hits)
            continue;
        const uint32_t t = (uint32_t)t64;
        const uint32_t gi   = __ldg(P.sortedIdx + t);
        const uint32_t g    = gi & 0x3fffffffu, hc = gi >> 30;
        const RayOfRecord rr = march_ray(P, g);
        const int probeLocal = rr.probeLocal, rayId = rr.rayId;
        const float4   rec = __ldg(P.records + g);
        float4 d4 = __ldg(P.dirs + rayId);
        float4 o4 = __ldg(P.origins + probeLocal);
        f3     d  = {d4.x, d4.y, d4.z}, o = {o4.x, o4.y, o4.z};
        float xp = sdf.sampleTex(rec.y + texelOffset, rec.z, rec.w);
        float xn = sdf.sampleTex(rec.y - texelOffset, rec.z, rec.w);
        float yp = sdf.sampleTex(rec.y, rec.z + texelOffset, rec.w);
        float yn = sdf.sampleTex(rec.y, rec.z - texelOffset, rec.w);
        float zp = sdf.sampleTex(rec.y, rec.z, rec.w + texelOffset);
        float zn = sdf.sampleTex(rec.y, rec.z, rec.w - texelOffset);
        f3    normal = normalize3({xp - xn, yp - yn, zp - zn});
        f3    hitPosition      = o + d * rec.x;
        float surfaceThreshold = data.cascadeVoxelSize[hc] * 1.05f;
        f4    sc = sample_global_surface_atlas_2p(P, hitPosition, normal, surfaceThreshold, &sCand[0][threadIdx.x]);
        uint32_t r0 = f2h_bits(sc.x), r1 = f2h_bits(sc.y), r2 = f2h_bits(sc.z);
        P.radiance[(size_t)probeLocal * P.raysPerProbe + rayId] = make_uint2(r0 | (r1 << 16), r2);
        flag_non_finite(P, make_uint2(r0 | (r1 << 16), r2));
    }
}

__global__ void probe_origins_kernel(const TraceParams P, float4* __restrict__ origins)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.probeCount)
        return;
    f3 o = probe_origin(P, shard_probe(P, i));
    origins[i] = make_float4(o.x, o.y, o.z, 0.0f);
}

// =====================================================================================================================
// Blend (+ fused border)
// =====================================================================================================================
//
// Because the weights are probe independent the blend is   Out[(probe,channel)][texel] = sum_r Val[(probe,channel)][r] * W[r][texel]
// i.e. a small-N SGEMM.  It is evaluated on the FP32 pipe with one FFMA per (value, weight) pair and each accumulator
// walks r = 0..R-1 in order, which is exactly the reference's sequential loop (ProbeUpdate.glsl:68-102) with the
// gated weights stored as zeros.  Register tile 8 (rows) x 8 (texels) per thread; A = values staged in shared memory
// as [row][k] (+1 pad), B = weights staged as [k][texel].

constexpr int KC = 32; // rays per shared-memory chunk (== warp size: the per-chunk live-ray mask is one ballot)

// mirrored border stores for interior texel (i, j) of a probe whose ring origin is (bx, by)  (BorderUpdate.glsl:25-133)
template <typename T>
__device__ __forceinline__ void store_with_border(T* __restrict__ img, int W, int bx, int by, int side, int i, int j, T v, bool border)
{
    const int x = i + 1, y = j + 1;
    img[(size_t)(by + y) * W + (bx + x)] = v;
    if (!border)
        return;
    if (y == 1)    img[(size_t)(by + 0) * W + (bx + side + 1 - x)] = v;
    if (y == side) img[(size_t)(by + side + 1) * W + (bx + side + 1 - x)] = v;
    if (x == 1)    img[(size_t)(by + side + 1 - y) * W + (bx + 0)] = v;
    if (x == side) img[(size_t)(by + side + 1 - y) * W + (bx + side + 1)] = v;
    if (x == side && y == side) img[(size_t)(by + 0) * W + (bx + 0)] = v;
    if (x == 1 && y == side)    img[(size_t)(by + 0) * W + (bx + side + 1)] = v;
    if (x == side && y == 1)    img[(size_t)(by + side + 1) * W + (bx + 0)] = v;
    if (x == 1 && y == 1)       img[(size_t)(by + side + 1) * W + (bx + side + 1)] = v;
}

// cp.async (LDGSTS) helpers: global -> shared without a register round trip, completion tracked per thread group.
__device__ __forceinline__ void cp_async_16(void* smemDst, const void* gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_8_zfill(void* smemDst, const void* gsrc, bool valid)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
    int      n = valid ? 8 : 0; // src-size 0: the 8 destination bytes are zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(n));
}
__device__ __forceinline__ void cp_async_4(void* smemDst, const void* gsrc)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Load TR consecutive floats (TR in {2,3,4,6}) from shared memory with the widest aligned vectors.
template <int TR>
__device__ __forceinline__ void lds_row(const float* __restrict__ p, float* a)
{
    if constexpr (TR == 4)
    {
        float4 v = *reinterpret_cast<const float4*>(p);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    }
    else if constexpr (TR % 2 == 0)
    {
#pragma unroll
        for (int i = 0; i < TR / 2; i++)
        {
            float2 v = *reinterpret_cast<const float2*>(p + 2 * i);
            a[2 * i] = v.x; a[2 * i + 1] = v.y;
        }
    }
    else
    {
#pragma unroll
        for (int i = 0; i < TR; i++)
            a[i] = p[i];
    }
}

// Blend tiling, second form (profiles/r1_*: the 8x8 tiles reached 28 % / 44 % of the FP32 peak and did all the work on
// weights that are exact zeros).  Now a WARP owns one group of 8 texels and its 32 lanes own 32 disjoint row groups
// (rows = (probe, channel)), so "all 8 weights of this ray are zero" is a warp-uniform condition and the ray is skipped
// for that warp: depth weights pow(cos, 50) are zero below the 1e-8 gate on ~85 % of the sphere, irradiance weights
// max(0, cos) on half of it.  Skipped terms are exact zeros, so every accumulator still sees the reference's sum in ray
// order.  A fragments are lane-contiguous (conflict-free), B fragments are warp broadcasts.

// Epilogue of the irradiance blend (ProbeUpdate.glsl:133-151 + BorderUpdate.glsl): Cs = [3 * PB rows (probe, channel)][64 texels] weighted sums.
template <int PB, int NT>
__device__ __forceinline__ void irradiance_epilogue(const BlendParams& P, const float* Cs, int probe0, int tid)
{
    constexpr int N = 64;
    const uint32_t one = 0x3c00u; // fp16 1.0 (alpha, ProbeUpdate.glsl:150)
    for (int idx = tid; idx < PB * N; idx += NT)
    {
        int p = idx / N, t = idx % N;
        int probeLocal = probe0 + p;
        if (probeLocal >= P.probeCount)
            continue;
        int   probe = shard_probe(P, probeLocal);
        float s = __ldg(P.scaleIrr + t);
        float r = Cs[(p * 3 + 0) * N + t] * s, g = Cs[(p * 3 + 1) * N + t] * s, b = Cs[(p * 3 + 2) * N + t] * s;
        r = pow_rn(r, P.invGamma);
        g = pow_rn(g, P.invGamma);
        b = pow_rn(b, P.invGamma);
        int i = t & 7, j = t >> 3;
        int bx = (probe % P.probesPerRow) * 10 + 1, by = (probe / P.probesPerRow) * 10 + 1;
        if (!P.firstFrame)
        {
            f4 prev = unpack_rgba16f(__ldg(P.prevIrr + (size_t)(by + j + 1) * P.irrWidth + (bx + i + 1)));
            r = mixh(r, prev.x, P.hysteresis);
            g = mixh(g, prev.y, P.hysteresis);
            b = mixh(b, prev.z, P.hysteresis);
        }
        uint2 o = make_uint2((uint32_t)f2h_bits(r) | ((uint32_t)f2h_bits(g) << 16), (uint32_t)f2h_bits(b) | (one << 16));
        store_with_border<uint2>(P.outIrr, P.irrWidth, bx, by, 8, i, j, o, P.fuseBorder != 0);
    }
}

// Irradiance: PB probes per block, rows m = p*3 + channel (M = 3*PB), 8 warps = the 8 rows of the 8x8 octahedral map.
template <int PB>
__global__ void __launch_bounds__(256) blend_irradiance_kernel(const __grid_constant__ BlendParams P)
{
    constexpr int M = 3 * PB, MS = M + 2, N = 64, NT = 256, TR = M / 32;
    static_assert(M % 32 == 0, "rows must split evenly over 32 lanes");
    extern __shared__ __align__(16) float smem[];
    // double-buffered by cp.async: raw fp16 ray texels + weights + zero masks of chunk c+1 arrive while chunk c is computed
    float*    As  = smem;                                   // [KC][MS] fp32 values of the current chunk
    float*    Bs  = As + KC * MS;                           // [2][KC][N]
    uint2*    Raw = reinterpret_cast<uint2*>(Bs + 2 * KC * N); // [2][PB][KC] RGBA16F texels
    uint32_t* Zs  = reinterpret_cast<uint32_t*>(Raw + 2 * PB * KC); // [2][KC] bit g = texel group g has a non-zero weight
    float*    Cs  = smem;                                   // epilogue overlay [M][N]

    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
    const int probe0 = blockIdx.x * PB;
    const int R = P.raysPerProbe;
    // A non-finite ray value (fp16 Inf / NaN: an HDR sky or light cache beyond 65504) somewhere in the ray buffers - flagged by whoever wrote them.
    // Inf * 0 would poison texels whose weight is gated to zero, where the reference SKIPS the ray (ProbeUpdate.glsl:93): the guarded loop then.
    const bool anyBad = P.nonFinite != nullptr && __ldg(P.nonFinite) != 0u;

    float acc[TR][8];
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
            acc[i][j] = 0.0f;

    auto prefetch = [&](int k0, int st) {
        for (int idx = tid; idx < PB * KC; idx += NT)
        {
            int  p = idx / KC, k = idx % KC; // consecutive threads -> consecutive rays of one probe (coalesced)
            int  probe = probe0 + p, ray = k0 + k;
            bool ok = probe < P.probeCount && ray < R;
            cp_async_8_zfill(Raw + (st * PB + p) * KC + k, P.radiance + (ok ? (size_t)probe * R + ray : 0), ok);
        }
        const float* src = P.wIrr + (size_t)k0 * N;
        for (int idx = tid; idx < KC * N / 4; idx += NT)
            cp_async_16(Bs + st * KC * N + idx * 4, src + idx * 4);
        if (tid < KC)
            cp_async_4(Zs + st * KC + tid, P.nzIrr + k0 + tid);
        cp_async_commit();
    };

    const int nChunks = P.raysPadded / KC;
    prefetch(0, 0);
    for (int c = 0; c < nChunks; c++)
    {
        const int st = c & 1;
        cp_async_wait_all();
        __syncthreads(); // chunk c landed; everybody is done computing chunk c-1
        for (int idx = tid; idx < PB * KC; idx += NT)
        {
            int   p = idx / KC, k = idx % KC;
            uint2 t = Raw[(st * PB + p) * KC + k];
            float* dst = As + k * MS + p * 3;
            dst[0] = h2f_bits((uint16_t)(t.x & 0xffffu));
            dst[1] = h2f_bits((uint16_t)(t.x >> 16));
            dst[2] = h2f_bits((uint16_t)(t.y & 0xffffu));
        }
        if (c + 1 < nChunks)
            prefetch((c + 1) * KC, st ^ 1);
        __syncthreads();
        const float*    B = Bs + st * KC * N;
        const uint32_t* Z = Zs + st * KC;
        // rays of this chunk with a non-zero weight for the warp's texel group (KC == 32: one ballot), visited in ray order
        uint32_t live = __ballot_sync(0xffffffffu, (Z[lane] >> grp) & 1u);
        while (live)
        {
            const int k = __ffs(live) - 1;
            live &= live - 1;
            float a[TR], b[8];
            lds_row<TR>(As + k * MS + lane * TR, a);
            float4 b0 = *reinterpret_cast<const float4*>(B + k * N + grp * 8);
            float4 b1 = *reinterpret_cast<const float4*>(B + k * N + grp * 8 + 4);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            if (!anyBad)
            {
#pragma unroll
                for (int i = 0; i < TR; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
            }
            else
            { // the reference SKIPS a ray whose weight is below the gate (ProbeUpdate.glsl:93): a gated (zero) weight must not meet an Inf
#pragma unroll
                for (int i = 0; i < TR; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        acc[i][j] = b[j] != 0.0f ? __fmaf_rn(a[i], b[j], acc[i][j]) : acc[i][j];
            }
        }
    }
    __syncthreads(); // before the epilogue overlays the staging buffers

#pragma unroll
    for (int i = 0; i < TR; i++)
    {
        float4* dst = reinterpret_cast<float4*>(Cs + (lane * TR + i) * N + grp * 8);
        dst[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        dst[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();

    irradiance_epilogue<PB, NT>(P, Cs, probe0, tid);
}

// Epilogue of the depth blend: Cs = [2 * PB rows (probe, {d, d*d})][256 weight-matrix columns].
template <int PB, int NT>
__device__ __forceinline__ void depth_epilogue(const BlendParams& P, const float* Cs, int probe0, int tid)
{
    constexpr int N = 256;
    for (int idx = tid; idx < PB * N; idx += NT)
    {
        int p = idx / N, t = idx % N;
        int probeLocal = probe0 + p;
        if (probeLocal >= P.probeCount)
            continue;
        int   probe = shard_probe(P, probeLocal);
        int i, j; // column t of the weight matrix -> texel (i, j)
        depth_texel_of_column(t, i, j);
        float s = __ldg(P.scaleDepth + (j * 16 + i));
        float r = Cs[(p * 2 + 0) * N + t] * s, g = Cs[(p * 2 + 1) * N + t] * s;
        int bx = (probe % P.probesPerRow) * 18 + 1, by = (probe / P.probesPerRow) * 18 + 1;
        if (!P.firstFrame)
        {
            uint32_t pv = __ldg(P.prevDepth + (size_t)(by + j + 1) * P.depthWidth + (bx + i + 1));
            r = mixh(r, h2f_bits((uint16_t)(pv & 0xffffu)), P.hysteresis);
            g = mixh(g, h2f_bits((uint16_t)(pv >> 16)), P.hysteresis);
        }
        uint32_t o = (uint32_t)f2h_bits(r) | ((uint32_t)f2h_bits(g) << 16);
        store_with_border<uint32_t>(P.outDepth, P.depthWidth, bx, by, 16, i, j, o, P.fuseBorder != 0);
    }
}

// Depth: PB probes per block, rows m = p*2 + {d, d*d} (M = 2*PB), 16 warps x 2 texel groups x 8 texels = 256 texels.
template <int PB>
__global__ void __launch_bounds__(512) blend_depth_kernel(const __grid_constant__ BlendParams P)
{
    constexpr int M = 2 * PB, MS = M + 4, N = 256, NT = 512, TR = M / 32;
    static_assert(M % 32 == 0, "rows must split evenly over 32 lanes");
    extern __shared__ __align__(16) float smem[];
    float*    As  = smem;                                   // [KC][MS] (d, d*d) of the current chunk
    float*    Bs  = As + KC * MS;                           // [2][KC][N]
    uint2*    Raw = reinterpret_cast<uint2*>(Bs + 2 * KC * N); // [2][PB][KC] direction/distance texels
    uint32_t* Zs  = reinterpret_cast<uint32_t*>(Raw + 2 * PB * KC); // [2][KC]
    float*    Cs  = smem;                                   // epilogue overlay [M][N]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int probe0 = blockIdx.x * PB;
    const int R = P.raysPerProbe;
    const bool anyBad = P.nonFinite != nullptr && __ldg(P.nonFinite) != 0u; // see blend_irradiance_kernel

    float acc[2][TR][8];
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int i = 0; i < TR; i++)
#pragma unroll
            for (int j = 0; j < 8; j++)
                acc[g][i][j] = 0.0f;

    auto prefetch = [&](int k0, int st) {
        for (int idx = tid; idx < PB * KC; idx += NT)
        {
            int  p = idx / KC, k = idx % KC;
            int  probe = probe0 + p, ray = k0 + k;
            bool ok = probe < P.probeCount && ray < R;
            cp_async_8_zfill(Raw + (st * PB + p) * KC + k, P.dirDist + (ok ? (size_t)probe * R + ray : 0), ok);
        }
        const float* src = P.wDepth + (size_t)k0 * N;
        for (int idx = tid; idx < KC * N / 4; idx += NT)
            cp_async_16(Bs + st * KC * N + idx * 4, src + idx * 4);
        if (tid < KC)
            cp_async_4(Zs + st * KC + tid, P.nzDepth + k0 + tid);
        cp_async_commit();
    };

    const int nChunks = P.raysPadded / KC;
    prefetch(0, 0);
    for (int c = 0; c < nChunks; c++)
    {
        const int st = c & 1, k0 = c * KC;
        cp_async_wait_all();
        __syncthreads();
        for (int idx = tid; idx < PB * KC; idx += NT)
        {
            int   p = idx / KC, k = idx % KC;
            float d = 0.0f;
            if (probe0 + p < P.probeCount && k0 + k < R)
            {
                uint2 t = Raw[(st * PB + p) * KC + k];
                d = gmin(P.maxDistance, h2f_bits((uint16_t)(t.y >> 16)) - 0.01f); // ProbeUpdate.glsl:75
                if (d == -1.0f)
                    d = P.maxDistance;
            }
            *reinterpret_cast<float2*>(As + k * MS + p * 2) = make_float2(d, d * d);
        }
        if (c + 1 < nChunks)
            prefetch((c + 1) * KC, st ^ 1);
        __syncthreads();
        const float*    B = Bs + st * KC * N;
        const uint32_t* Z = Zs + st * KC;
        uint32_t live = __ballot_sync(0xffffffffu, ((Z[lane] >> (warp * 2)) & 3u) != 0u);
        while (live)
        {
            const int k = __ffs(live) - 1;
            live &= live - 1;
            const uint32_t z = (Z[k] >> (warp * 2)) & 3u;
            float a[TR];
            lds_row<TR>(As + k * MS + lane * TR, a);
#pragma unroll
            for (int g = 0; g < 2; g++)
            {
                if (!((z >> g) & 1u))
                    continue;
                float  b[8];
                float4 b0 = *reinterpret_cast<const float4*>(B + k * N + (warp * 2 + g) * 8);
                float4 b1 = *reinterpret_cast<const float4*>(B + k * N + (warp * 2 + g) * 8 + 4);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
                if (!anyBad)
                {
#pragma unroll
                    for (int i = 0; i < TR; i++)
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            acc[g][i][j] = __fmaf_rn(a[i], b[j], acc[g][i][j]);
                }
                else
                {
#pragma unroll
                    for (int i = 0; i < TR; i++)
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            acc[g][i][j] = b[j] != 0.0f ? __fmaf_rn(a[i], b[j], acc[g][i][j]) : acc[g][i][j];
                }
            }
        }
    }
    __syncthreads();

#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int i = 0; i < TR; i++)
        {
            float4* dst = reinterpret_cast<float4*>(Cs + (lane * TR + i) * N + (warp * 2 + g) * 8);
            dst[0] = make_float4(acc[g][i][0], acc[g][i][1], acc[g][i][2], acc[g][i][3]);
            dst[1] = make_float4(acc[g][i][4], acc[g][i][5], acc[g][i][6], acc[g][i][7]);
        }
    __syncthreads();

    depth_epilogue<PB, NT>(P, Cs, probe0, tid);
}


#include "blend_lists.inc"
#include "blend_tc.inc"
#include "blend_umma.inc"

// Standalone border pass (BorderUpdate.glsl:136-156): one warp-sized group of threads per probe and atlas.
template <typename T, int SIDE>
__global__ void border_kernel(T* __restrict__ img, int W, int probesPerRow, int probeBegin, int probeCount, int layerProbes, int layerStride)
{
    constexpr int NB = 4 * SIDE + 4; // 36 / 68 copies
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int p = idx / NB, e = idx % NB;
    if (p >= probeCount)
        return;
    int probe = shard_probe(probeBegin, layerProbes, layerStride, p);
    int bx = (probe % probesPerRow) * (SIDE + 2) + 1, by = (probe / probesPerRow) * (SIDE + 2) + 1;
    int sx, sy, dx, dy;
    if (e < SIDE)            { int x = e + 1;            sx = SIDE + 1 - x; sy = 1;    dx = x; dy = 0; }
    else if (e < 2 * SIDE)   { int x = e - SIDE + 1;     sx = SIDE + 1 - x; sy = SIDE; dx = x; dy = SIDE + 1; }
    else if (e < 3 * SIDE)   { int y = e - 2 * SIDE + 1; sx = 1;    sy = SIDE + 1 - y; dx = 0;        dy = y; }
    else if (e < 4 * SIDE)   { int y = e - 3 * SIDE + 1; sx = SIDE; sy = SIDE + 1 - y; dx = SIDE + 1; dy = y; }
    else if (e == 4 * SIDE)     { sx = SIDE; sy = SIDE; dx = 0;        dy = 0; }
    else if (e == 4 * SIDE + 1) { sx = 1;    sy = SIDE; dx = SIDE + 1; dy = 0; }
    else if (e == 4 * SIDE + 2) { sx = SIDE; sy = 1;    dx = 0;        dy = SIDE + 1; }
    else                        { sx = 1;    sy = 1;    dx = SIDE + 1; dy = SIDE + 1; }
    img[(size_t)(by + dy) * W + (bx + dx)] = img[(size_t)(by + sy) * W + (bx + sx)];
}

// =====================================================================================================================
// Consumer ("next" row f2): sampleIrradiance (DDGICommon.glsl:163-233) and SampleProbe.comp
// =====================================================================================================================

struct AtlasView
{
    const uint16_t* d;
    int             w, h, c;
};

__device__ __forceinline__ int wrapi(int i, int n) { return ((i % n) + n) % n; }

// textureLod(sampler2D, uv, 0): bilinear, repeat wrap, fp16 texels, fp32 weights
__device__ __forceinline__ void sample_atlas(const AtlasView& t, float u, float v, float* out)
{
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int   x0 = wrapi((int)fx, t.w), x1 = wrapi((int)fx + 1, t.w), y0 = wrapi((int)fy, t.h), y1 = wrapi((int)fy + 1, t.h);
    for (int k = 0; k < t.c; k++)
    {
        float a = lerp1(ld_h(t.d + ((size_t)y0 * t.w + x0) * t.c + k), ld_h(t.d + ((size_t)y0 * t.w + x1) * t.c + k), ax);
        float b = lerp1(ld_h(t.d + ((size_t)y1 * t.w + x0) * t.c + k), ld_h(t.d + ((size_t)y1 * t.w + x1) * t.c + k), ax);
        out[k]  = lerp1(a, b, ay);
    }
}

__device__ __forceinline__ void oct_encode(f3 v, float& ox, float& oy)
{
    float l1  = (fabsf(v.x) + fabsf(v.y)) + fabsf(v.z);
    float inv = __fdiv_rn(1.0f, l1);
    ox = v.x * inv;
    oy = v.y * inv;
    if (v.z < 0.0f)
    {
        float qx = (1.0f - fabsf(oy)) * sign_not_zero(ox), qy = (1.0f - fabsf(ox)) * sign_not_zero(oy);
        ox = qx;
        oy = qy;
    }
}

// DDGICommon.glsl:141-158
__device__ __forceinline__ void texture_coord_from_direction(f3 dir, int probeIndex, int width, int height, int side, float& u, float& v)
{
    float ox, oy;
    oct_encode(normalize3(dir), ox, oy);
    float zx = (ox + 1.0f) * 0.5f, zy = (oy + 1.0f) * 0.5f;
    float withBorder = (float)side + 2.0f;
    float tx = __fdiv_rn(zx * (float)side, (float)width), ty = __fdiv_rn(zy * (float)side, (float)height);
    int   perRow = (width - 2) / (int)withBorder;
    float fi = (float)probeIndex, fp = (float)perRow;
    float modv = fi - fp * floorf(__fdiv_rn(fi, fp));
    float px = modv * withBorder + 2.0f, py = (float)(probeIndex / perRow) * withBorder + 2.0f;
    u = __fdiv_rn(px, (float)width) + tx;
    v = __fdiv_rn(py, (float)height) + ty;
}

__device__ f3 sample_irradiance(const LuxDDGIUniform& ddgi, f3 P, f3 N, f3 Wo, const AtlasView& irr, const AtlasView& dep)
{
    int   bg[3], cnt[3] = {ddgi.probeCounts[0], ddgi.probeCounts[1], ddgi.probeCounts[2]};
    float Pv[3] = {P.x, P.y, P.z};
#pragma unroll
    for (int a = 0; a < 3; a++)
        bg[a] = iclamp((int)__fdiv_rn(Pv[a] - ddgi.startPosition[a], ddgi.step[a]), 0, cnt[a] - 1);
    f3 base = {ddgi.step[0] * (float)bg[0] + ddgi.startPosition[0], ddgi.step[1] * (float)bg[1] + ddgi.startPosition[1],
               ddgi.step[2] * (float)bg[2] + ddgi.startPosition[2]};
    f3    sum = {0.0f, 0.0f, 0.0f};
    float sumWeight = 0.0f;
    float al[3] = {gclamp(__fdiv_rn(P.x - base.x, ddgi.step[0]), 0.0f, 1.0f), gclamp(__fdiv_rn(P.y - base.y, ddgi.step[1]), 0.0f, 1.0f),
                   gclamp(__fdiv_rn(P.z - base.z, ddgi.step[2]), 0.0f, 1.0f)};
#pragma unroll 1
    for (int i = 0; i < 8; ++i)
    {
        int off[3] = {i & 1, (i >> 1) & 1, (i >> 2) & 1}, pg[3];
        float tri[3];
#pragma unroll
        for (int a = 0; a < 3; a++)
        {
            pg[a]  = iclamp(bg[a] + off[a], 0, cnt[a] - 1);
            tri[a] = (1.0f - al[a]) * (1.0f - (float)off[a]) + al[a] * (float)off[a];
        }
        f3 probePos = {ddgi.step[0] * (float)pg[0] + ddgi.startPosition[0], ddgi.step[1] * (float)pg[1] + ddgi.startPosition[1],
                       ddgi.step[2] * (float)pg[2] + ddgi.startPosition[2]};
        float weight = 1.0f;
        f3    dirToProbe = normalize3(probePos - P);
        float bf = gmax(0.0001f, (dot3(dirToProbe, N) + 1.0f) * 0.5f);
        weight *= bf * bf + 0.2f;
        int probeIdx = pg[0] + pg[1] * cnt[0] + pg[2] * cnt[0] * cnt[1];

        f3    vBias        = (N + Wo * 3.0f) * ddgi.normalBias;
        f3    probeToPoint = (P - probePos) + vBias;
        f3    dir          = normalize3({-probeToPoint.x, -probeToPoint.y, -probeToPoint.z});
        float u, v, tmp[4];
        texture_coord_from_direction({-dir.x, -dir.y, -dir.z}, probeIdx, ddgi.depthTextureWidth, ddgi.depthTextureHeight, ddgi.depthProbeSideLength, u, v);
        float dist = length3(probeToPoint);
        sample_atlas(dep, u, v, tmp);
        float mean = tmp[0];
        float variance = fabsf(tmp[0] * tmp[0] - tmp[1]);
        float dm = gmax(dist - mean, 0.0f);
        float cheb = __fdiv_rn(variance, variance + dm * dm);
        cheb = gmax(cheb * cheb * cheb, 0.0f);
        weight *= (dist <= mean) ? 1.0f : cheb;
        weight = gmax(0.000001f, weight);

        texture_coord_from_direction(normalize3(N), probeIdx, ddgi.irradianceTextureWidth, ddgi.irradianceTextureHeight, ddgi.irradianceProbeSideLength, u, v);
        sample_atlas(irr, u, v, tmp);
        float e = ddgi.ddgiGamma * 0.5f;
        f3    pi = {pow_rn(tmp[0], e), pow_rn(tmp[1], e), pow_rn(tmp[2], e)};
        const float crush = 0.2f;
        if (weight < crush)
            weight *= weight * weight * __fdiv_rn(1.0f, crush * crush);
        weight *= tri[0] * tri[1] * tri[2];
        sum = sum + pi * weight;
        sumWeight += weight;
    }
    f3 net = {__fdiv_rn(sum.x, sumWeight), __fdiv_rn(sum.y, sumWeight), __fdiv_rn(sum.z, sumWeight)};
    net    = net * net;
    return net * 6.283185482025146484375f;
}

__global__ void sample_irradiance_kernel(const __grid_constant__ LuxDDGIUniform ddgi, const uint16_t* __restrict__ irr,
                                         const uint16_t* __restrict__ dep, int count, const float* __restrict__ P, const float* __restrict__ N,
                                         const float* __restrict__ Wo, float* __restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    AtlasView ai{irr, ddgi.irradianceTextureWidth, ddgi.irradianceTextureHeight, 4}, ad{dep, ddgi.depthTextureWidth, ddgi.depthTextureHeight, 2};
    f3 r = sample_irradiance(ddgi, {P[3 * k], P[3 * k + 1], P[3 * k + 2]}, {N[3 * k], N[3 * k + 1], N[3 * k + 2]},
                             {Wo[3 * k], Wo[3 * k + 1], Wo[3 * k + 2]}, ai, ad);
    out[3 * k] = r.x; out[3 * k + 1] = r.y; out[3 * k + 2] = r.z;
}

struct SampleProbeArgs
{
    LuxDDGIUniform ddgi;
    float          cameraPosition[4];
    float          viewProjInv[16];
    int            width, height;
};

// SampleProbe.comp:36-60
__global__ void sample_probe_kernel(const __grid_constant__ SampleProbeArgs A, const uint16_t* __restrict__ irr, const uint16_t* __restrict__ dep,
                                    const float* __restrict__ gDepth, const float4* __restrict__ gNormal, float4* __restrict__ out)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= A.width || y >= A.height)
        return;
    size_t o = (size_t)y * A.width + x;
    float  d = __ldg(gDepth + o);
    if (d == 1.0f)
    {
        out[o] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    float tx = __fdiv_rn((float)x + 0.5f, (float)A.width), ty = __fdiv_rn((float)y + 0.5f, (float)A.height);
    float sx = tx * 2.0f - 1.0f, sy = ty * 2.0f - 1.0f;
    const float* m = A.viewProjInv;
    float wx = ((m[0] * sx + m[4] * sy) + m[8] * d) + m[12] * 1.0f;
    float wy = ((m[1] * sx + m[5] * sy) + m[9] * d) + m[13] * 1.0f;
    float wz = ((m[2] * sx + m[6] * sy) + m[10] * d) + m[14] * 1.0f;
    float ww = ((m[3] * sx + m[7] * sy) + m[11] * d) + m[15] * 1.0f;
    f3 Pw = {__fdiv_rn(wx, ww), __fdiv_rn(wy, ww), __fdiv_rn(wz, ww)};
    float4 n4 = __ldg(gNormal + o);
    // octohedralToDirection, Common/Math.glsl:27-33
    f3 v = {n4.x, n4.y, (1.0f - fabsf(n4.x)) - fabsf(n4.y)};
    if (v.z < 0.0f)
    {
        float s0 = (v.x >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f, s1 = (v.y >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f;
        float nx = (1.0f - fabsf(v.y)) * s0, ny = (1.0f - fabsf(v.x)) * s1;
        v.x = nx;
        v.y = ny;
    }
    f3 Nn = normalize3(v);
    f3 cam = {A.cameraPosition[0], A.cameraPosition[1], A.cameraPosition[2]};
    f3 Wo = normalize3(cam - Pw);
    AtlasView ai{irr, A.ddgi.irradianceTextureWidth, A.ddgi.irradianceTextureHeight, 4}, ad{dep, A.ddgi.depthTextureWidth, A.ddgi.depthTextureHeight, 2};
    f3 r = sample_irradiance(A.ddgi, Pw, Nn, Wo, ai, ad);
    out[o] = make_float4(r.x, r.y, r.z, 1.0f);
}

// SDFAtlasIndirectLight.frag:44-67 + additive blend into the RGBA16F light cache, for a list of atlas texels
__global__ void indirect_light_kernel(const __grid_constant__ LuxDDGIUniform ddgi, const uint16_t* __restrict__ irr, const uint16_t* __restrict__ dep,
                                      uint2* __restrict__ light, const uint2* __restrict__ base, int count, const uint32_t* __restrict__ texel,
                                      const float* __restrict__ P, const float* __restrict__ N, const float* __restrict__ albedo,
                                      const float* __restrict__ metallic, float intensity, float camX, float camY, float camZ)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    AtlasView ai{irr, ddgi.irradianceTextureWidth, ddgi.irradianceTextureHeight, 4}, ad{dep, ddgi.depthTextureWidth, ddgi.depthTextureHeight, 2};
    f3 Pw = {P[3 * k], P[3 * k + 1], P[3 * k + 2]}, Nn = {N[3 * k], N[3 * k + 1], N[3 * k + 2]};
    f3 cam = {camX, camY, camZ};
    f3 Wo = normalize3(cam - Pw);
    f3 E  = sample_irradiance(ddgi, Pw, Nn, Wo, ai, ad);
    const float PI_F = 3.1415926535897932384626433832795f;
    float m = metallic[k];
    float Ev[3] = {E.x, E.y, E.z}, outv[3];
    uint32_t idx = texel[k];
    uint2 b = base ? base[idx] : light[idx];
    f4 dst = unpack_rgba16f(b);
    float dv[3] = {dst.x, dst.y, dst.z};
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        float a       = gmin(albedo[3 * k + c], 0.9f);
        float diffuse = __fdiv_rn(a - a * m, PI_F);
        outv[c]       = dv[c] + (intensity * diffuse) * Ev[c];
    }
    uint32_t h0 = f2h_bits(outv[0]), h1 = f2h_bits(outv[1]), h2 = f2h_bits(outv[2]);
    uint32_t h3 = f2h_bits(dst.w + 1.0f); // the shader writes alpha 1 and the pass blends ONE + ONE on alpha too (VulkanPipeline.cpp:160-165)
    light[idx] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
}

// SDFDeferredLight.frag:44-129 (fetchLight for directional / point / spot lights, shadow ray through the global SDF with start bias 2,
// BRDF of Raytraced/BRDF.glsl:8-36,65-83) for a list of surface-cache texels, blended additively into the RGBA16F light cache as the
// reference's pipeline state does (GlobalSurfaceAtlas.cpp:950-972: BlendMode::Add, so alpha accumulates the 1.0 the shader writes).
// One thread per listed texel; the shadow march reuses trace_global_sdf_general.  pow() is binary64, rounded once (contract §4).
__device__ __forceinline__ float powd_rn(float x, float y) { return (float)pow((double)x, (double)y); }

template <bool TEX>
__global__ void __launch_bounds__(128) direct_light_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ LuxLight L, float camX, float camY,
                                                           float camZ, float shadowBias, uint2* __restrict__ light, int count,
                                                           const uint32_t* __restrict__ texel, const float* __restrict__ Pw, const float* __restrict__ N,
                                                           const float* __restrict__ albedo, const float* __restrict__ metallicRoughness)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    const SdfSampler<TEX> sdf(P);
    const float PI_F = 3.14159265358979323846f;
    const f3    worldPos = {Pw[3 * k], Pw[3 * k + 1], Pw[3 * k + 2]}, normal = {N[3 * k], N[3 * k + 1], N[3 * k + 2]};
    const float alb[3] = {albedo[3 * k], albedo[3 * k + 1], albedo[3 * k + 2]};
    const float metallic = metallicRoughness[2 * k], roughness = metallicRoughness[2 * k + 1];
    // fetchLight, SDFDeferredLight.frag:44-81
    f3    Wi = {0.0f, 0.0f, 0.0f};
    float dist = LUX_GLOBAL_SDF_WORLD_SIZE, atten = 1.0f;
    const f3 lightPos = {L.position[0], L.position[1], L.position[2]};
    if (L.type == LUX_LIGHT_DIRECTIONAL)
        Wi = {-L.direction[0], -L.direction[1], -L.direction[2]};
    else if (L.type == LUX_LIGHT_POINT)
    {
        f3    dir = lightPos - worldPos;
        float d   = length3(dir);
        Wi        = normalize3(dir);
        atten     = __fdiv_rn(L.radius, powd_rn(d, 2.0f) + 1.0f);
        dist      = d;
    }
    else if (L.type == LUX_LIGHT_SPOT)
    {
        f3    Lv          = lightPos - worldPos;
        float cutoffAngle = 1.0f - L.angle;
        f3    lightDir    = normalize3(Lv);
        float d           = length3(Lv);
        float theta       = dot3(lightDir, {L.direction[0], L.direction[1], L.direction[2]});
        float epsilon     = cutoffAngle - cutoffAngle * 0.9f;
        atten             = __fdiv_rn(theta - cutoffAngle, epsilon);
        atten *= __fdiv_rn(L.radius, powd_rn(d, 2.0f) + 1.0f);
        atten = gclamp(atten, 0.0f, 1.0f);
        Wi    = lightDir;
        dist  = d;
    }
    float shadowMask = 1.0f;
    const float NoL  = dot3(normal, Wi);
    const float bias = (2.0f * shadowBias) * gclamp(1.0f - NoL, 0.0f, 1.0f) + shadowBias;
    if (NoL > 0.0f)
    {
        if (atten > 0.0f)
        {
            const f3 origin = worldPos + normal * shadowBias;
            LuxGlobalSDFTrace tr;
            tr.worldPosition[0] = origin.x; tr.worldPosition[1] = origin.y; tr.worldPosition[2] = origin.z;
            tr.minDistance = 0.0f;
            tr.worldDirection[0] = Wi.x; tr.worldDirection[1] = Wi.y; tr.worldDirection[2] = Wi.z;
            tr.maxDistance = dist - bias;
            tr.stepScale = 1.0f;
            tr.needsHitNormal = 0u;
            LuxGlobalSDFHit hit = trace_global_sdf_general<TEX>(P, sdf, tr, 2.0f);
            shadowMask = hit.hitTime >= 0.0f ? 0.0f : 1.0f;
        }
    }
    else
        shadowMask = 0.0f;
    const f3    cam  = {camX, camY, camZ};
    const f3    view = normalize3(cam - worldPos);
    const float intensity = powd_rn(L.intensity, 1.4f) + 0.1f;
    const float Lrad[3] = {L.color[0] * intensity, L.color[1] * intensity, L.color[2] * intensity};
    const f3    Lh = normalize3(Wi + view);
    // BRDF, Raytraced/BRDF.glsl:65-83
    const float Fd = 0.04f, EPSILON = 0.00001f;
    const float cosLi = gmax(0.0f, dot3(normal, Wi)), cosLh = gmax(0.0f, dot3(normal, Lh)), NdotV = gmax(0.0f, dot3(normal, view));
    const float ct = gmax(dot3(Lh, view), 0.0f);
    const float p5 = powd_rn(gclamp(1.0f - ct, 0.0f, 1.0f), 5.0f);
    const float alpha = roughness * roughness, alphaSq = alpha * alpha;
    const float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    const float D     = __fdiv_rn(alphaSq, (PI_F * denom) * denom);
    const float r1 = roughness + 1.0f, kk = __fdiv_rn(r1 * r1, 8.0f);
    const float G   = __fdiv_rn(cosLi, cosLi * (1.0f - kk) + kk) * __fdiv_rn(NdotV, NdotV * (1.0f - kk) + kk);
    const float den = gmax(EPSILON, (4.0f * cosLi) * NdotV);
    const uint32_t idx = texel[k];
    const uint2    b   = light[idx];
    const f4       dst = unpack_rgba16f(b);
    const float    dv[3] = {dst.x, dst.y, dst.z};
    uint32_t h[4];
#pragma unroll
    for (int c = 0; c < 3; c++)
    {
        float F0 = Fd * (1.0f - metallic) + alb[c] * metallic;
        float F  = F0 + (1.0f - F0) * p5;
        float kd = (1.0f - F) * (1.0f - metallic);
        float br = __fdiv_rn(kd * alb[c], PI_F) + __fdiv_rn((F * D) * G, den);
        float o  = (((br * Lrad[c]) * cosLi) * shadowMask) * atten;
        h[c]     = f2h_bits(o + dv[c]);
    }
    h[3]       = f2h_bits(1.0f + dst.w);
    light[idx] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
}

// =====================================================================================================================
// The screen-space tracyGlobalSDF users: SDFReflection.comp and SDFShadow.comp over a G-buffer.
// cross / reflect / the UNORM8 blue-noise fetch under the numerics contract: every product, sum and difference is rounded.
// =====================================================================================================================
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
__device__ __forceinline__ f3 reflect3(f3 i, f3 n)
{
    float t = 2.0f * dot3(n, i);
    return {i.x - t * n.x, i.y - t * n.y, i.z - t * n.z};
}
__device__ __forceinline__ float unorm8(uint32_t b) { return __fdiv_rn((float)b, 255.0f); }
// sampleBlueNoise, Raytraced/BlueNoise.glsl:8-19; sobol = 256 x 1 RGBA8, scr = 128 x 128 RGBA8 (one uint32 per texel, r in the low byte)
__device__ __forceinline__ float sample_blue_noise(int cx, int cy, int samplerIndex, int dimension, const uint32_t* __restrict__ sobol,
                                                   const uint32_t* __restrict__ scr)
{
    cx           = cx % 128;
    cy           = cy % 128;
    samplerIndex = samplerIndex % 256;
    dimension    = dimension % 4;
    const uint32_t t = __ldg(scr + cy * 128 + cx);
    int rankedIndex = samplerIndex ^ (int)gclamp(unorm8((t >> 16) & 0xffu) * 256.0f, 0.0f, 255.0f);
    int value       = (int)gclamp(unorm8((__ldg(sobol + rankedIndex) >> (8 * dimension)) & 0xffu) * 256.0f, 0.0f, 255.0f);
    value           = value ^ (int)gclamp(unorm8((t >> (8 * (dimension % 2))) & 0xffu) * 256.0f, 0.0f, 255.0f);
    return __fdiv_rn(0.5f + (float)value, 256.0f);
}
__device__ __forceinline__ f3 world_position_from_depth(float tx, float ty, float d, const float* __restrict__ m) // Common/Math.glsl:35-42
{
    float sx = tx * 2.0f - 1.0f, sy = ty * 2.0f - 1.0f;
    float wx = ((m[0] * sx + m[4] * sy) + m[8] * d) + m[12] * 1.0f;
    float wy = ((m[1] * sx + m[5] * sy) + m[9] * d) + m[13] * 1.0f;
    float wz = ((m[2] * sx + m[6] * sy) + m[10] * d) + m[14] * 1.0f;
    float ww = ((m[3] * sx + m[7] * sy) + m[11] * d) + m[15] * 1.0f;
    return {__fdiv_rn(wx, ww), __fdiv_rn(wy, ww), __fdiv_rn(wz, ww)};
}
__device__ __forceinline__ f3 octohedral_to_direction(float ex, float ey) // Common/Math.glsl:27-33
{
    f3 v = {ex, ey, (1.0f - fabsf(ex)) - fabsf(ey)};
    if (v.z < 0.0f)
    {
        float s0 = (v.x >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f, s1 = (v.y >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f;
        float nx = (1.0f - fabsf(v.y)) * s0, ny = (1.0f - fabsf(v.x)) * s1;
        v.x = nx;
        v.y = ny;
    }
    return normalize3(v);
}
// importanceSampleGGX(...).xyz, Raytraced/BRDF.glsl:176-201
__device__ __forceinline__ f3 importance_sample_ggx(float ex, float ey, f3 N, float roughness)
{
    const float TWO_PI = 6.283185482025146484375f;
    float a = roughness * roughness, m2 = a * a;
    float phi      = TWO_PI * ex;
    float cosTheta = __fsqrt_rn(__fdiv_rn(1.0f - ey, 1.0f + (m2 - 1.0f) * ey));
    float sinTheta = __fsqrt_rn(1.0f - cosTheta * cosTheta);
    f3    H  = {cos_rn(phi) * sinTheta, sin_rn(phi) * sinTheta, cosTheta};
    f3    up = fabsf(N.z) < 0.999f ? f3{0.0f, 0.0f, 1.0f} : f3{1.0f, 0.0f, 0.0f};
    f3    tangent   = normalize3(cross3(up, N));
    f3    bitangent = cross3(N, tangent);
    f3    sv = {(tangent.x * H.x + bitangent.x * H.y) + N.x * H.z, (tangent.y * H.x + bitangent.y * H.y) + N.y * H.z,
                (tangent.z * H.x + bitangent.z * H.y) + N.z * H.z};
    return normalize3(sv);
}
// texture(samplerCube, dir) with its alpha (sample_sky returns .rgb for the probe trace)
__device__ __forceinline__ f4 sample_sky4(const TraceParams& P, f3 d)
{
    if (P.sky == nullptr || P.skyFace <= 0)
        return {0.0f, 0.0f, 0.0f, 0.0f};
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int   face;
    float sc, tc, ma;
    if (az >= ax && az >= ay) { face = d.z >= 0.0f ? 4 : 5; sc = d.z >= 0.0f ? d.x : -d.x; tc = -d.y; ma = az; }
    else if (ay >= ax)        { face = d.y >= 0.0f ? 2 : 3; sc = d.x; tc = d.y >= 0.0f ? d.z : -d.z; ma = ay; }
    else                      { face = d.x >= 0.0f ? 0 : 1; sc = d.x >= 0.0f ? -d.z : d.z; tc = -d.y; ma = ax; }
    float u = 0.5f * __fdiv_rn(sc, ma) + 0.5f, v = 0.5f * __fdiv_rn(tc, ma) + 0.5f;
    int   N = P.skyFace;
    float x = u * (float)N - 0.5f, y = v * (float)N - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float axw = x - fx, ayw = y - fy;
    int   x0 = iclamp((int)fx, 0, N - 1), x1 = iclamp((int)fx + 1, 0, N - 1);
    int   y0 = iclamp((int)fy, 0, N - 1), y1 = iclamp((int)fy + 1, 0, N - 1);
    const uint2* base = P.sky + (size_t)face * N * N;
    f4 t00 = unpack_rgba16f(__ldg(base + (size_t)y0 * N + x0)), t10 = unpack_rgba16f(__ldg(base + (size_t)y0 * N + x1));
    f4 t01 = unpack_rgba16f(__ldg(base + (size_t)y1 * N + x0)), t11 = unpack_rgba16f(__ldg(base + (size_t)y1 * N + x1));
    f4 out;
    out.x = lerp1(lerp1(t00.x, t10.x, axw), lerp1(t01.x, t11.x, axw), ayw);
    out.y = lerp1(lerp1(t00.y, t10.y, axw), lerp1(t01.y, t11.y, axw), ayw);
    out.z = lerp1(lerp1(t00.z, t10.z, axw), lerp1(t01.z, t11.z, axw), ayw);
    out.w = lerp1(lerp1(t00.w, t10.w, axw), lerp1(t01.w, t11.w, axw), ayw);
    return out;
}

struct ReflectionArgs
{
    LuxDDGIUniform             ddgi;
    LuxReflectionPushConstants push;
    const uint16_t *           irr, *dep;
    int                        width, height;
    const float *              gDepth, *gNormal, *gPbr;
    const uint32_t *           sobol, *scr;
    uint2*                     out;
};

// SDFReflection.comp:84-163, one thread per pixel
template <bool TEX>
__global__ void __launch_bounds__(128) sdf_reflection_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ ReflectionArgs A)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= A.width * A.height)
        return;
    const int   x = o % A.width, y = o / A.width;
    const float depth = __ldg(A.gDepth + o);
    if (!(depth != 1.0f))
        return;
    const SdfSampler<TEX> sdf(P);
    const float tx = __fdiv_rn((float)x + 0.5f, (float)A.width), ty = __fdiv_rn((float)y + 0.5f, (float)A.height);
    f3           worldPos  = world_position_from_depth(tx, ty, depth, A.push.viewProjInv);
    const float  roughness = __ldg(A.gPbr + 4 * (size_t)o + 1);
    const float4 n4        = __ldg(reinterpret_cast<const float4*>(A.gNormal) + o);
    const f3     normal    = octohedral_to_direction(n4.x, n4.y);
    const f3     cam       = {A.push.cameraPosition[0], A.push.cameraPosition[1], A.push.cameraPosition[2]};
    const f3     Wo        = normalize3(cam - worldPos);
    worldPos               = worldPos + normal * P.sdf.cascadeVoxelSize[0];
    const f3 negWo = {-Wo.x, -Wo.y, -Wo.z};
    f4   color = {0.0f, 0.0f, 0.0f, 0.0f};
    f3   R;
    bool traceRay = true;
    if (roughness < 0.05f) // MIRROR_REFLECTIONS_ROUGHNESS_THRESHOLD
        R = reflect3(negWo, normal);
    else if (roughness > 0.45f && A.push.approximateWithDDGI == 1u) // DDGI_REFLECTIONS_ROUGHNESS_THRESHOLD
    {
        R = reflect3(negWo, normal);
        AtlasView ai{A.irr, A.ddgi.irradianceTextureWidth, A.ddgi.irradianceTextureHeight, 4}, ad{A.dep, A.ddgi.depthTextureWidth, A.ddgi.depthTextureHeight, 2};
        f3 e     = sample_irradiance(A.ddgi, worldPos, R, Wo, ai, ad);
        color    = {A.push.roughDDGIIntensity * e.x, A.push.roughDDGIIntensity * e.y, A.push.roughDDGIIntensity * e.z, 0.0f};
        traceRay = false;
    }
    else
    {
        float xi0 = sample_blue_noise(x, y, (int)A.push.numFrames, 0, A.sobol, A.scr) * A.push.trim;
        float xi1 = sample_blue_noise(x, y, (int)A.push.numFrames, 1, A.sobol, A.scr) * A.push.trim;
        f3    Wh  = importance_sample_ggx(xi0, xi1, normal, roughness);
        R         = reflect3(negWo, Wh);
    }
    if (traceRay)
    { // trace(), SDFReflection.comp:86-117
        LuxGlobalSDFTrace tr;
        tr.worldPosition[0] = worldPos.x; tr.worldPosition[1] = worldPos.y; tr.worldPosition[2] = worldPos.z;
        tr.minDistance = 0.0f;
        tr.worldDirection[0] = R.x; tr.worldDirection[1] = R.y; tr.worldDirection[2] = R.z;
        tr.maxDistance = LUX_GLOBAL_SDF_WORLD_SIZE;
        tr.stepScale = 1.0f;
        tr.needsHitNormal = 0u;
        const LuxGlobalSDFHit hit = trace_global_sdf_general<TEX>(P, sdf, tr, 0.0f);
        if (hit.hitTime >= 0.0f)
        {
            float surfaceThreshold = P.sdf.cascadeVoxelSize[hit.hitCascade] * 1.05f;
            color = sample_global_surface_atlas(P, worldPos + R * hit.hitTime, {-R.x, -R.y, -R.z}, surfaceThreshold);
        }
        else
            color = sample_sky4(P, R);
    }
    const uint32_t h0 = f2h_bits(color.x), h1 = f2h_bits(color.y), h2 = f2h_bits(color.z), h3 = f2h_bits(color.w);
    A.out[o] = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
}

struct ShadowArgs
{
    LuxLight         light;
    float            viewProjInv[16];
    uint32_t         numFrames;
    float            shadowBias;
    int              width, height;
    const float *    gDepth, *gNormal;
    const uint32_t * sobol, *scr;
    uint32_t*        out;
};

// SDFShadow.comp:121-157 with fetchLight :42-117 (softShadow = true): one warp = one 8 x 4 workgroup, the visibility word is a ballot
template <bool TEX>
__global__ void __launch_bounds__(128) sdf_shadow_kernel(const __grid_constant__ TraceParams P, const __grid_constant__ ShadowArgs A)
{
    const int gw = A.width / 8, gh = A.height / 4;
    const int g  = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (g >= gw * gh) // whole warps leave together
        return;
    const int li = threadIdx.x & 31;
    const int x = (g % gw) * 8 + (li % 8), y = (g / gw) * 4 + (li / 8);
    const size_t o     = (size_t)y * A.width + x;
    const float  depth = __ldg(A.gDepth + o);
    const bool   live  = depth != 1.0f;
    uint32_t     result = 0;
    if (live)
    {
        const SdfSampler<TEX> sdf(P);
        const LuxLight&       L = A.light;
        const float PI_F = 3.1415926535897932384626433832795f;
        const float tx = __fdiv_rn((float)x + 0.5f, (float)A.width), ty = __fdiv_rn((float)y + 0.5f, (float)A.height);
        const f3     worldPos = world_position_from_depth(tx, ty, depth, A.viewProjInv);
        const float4 n4       = __ldg(reinterpret_cast<const float4*>(A.gNormal) + o);
        const f3     normal   = octohedral_to_direction(n4.x, n4.y);
        const float  rx = sample_blue_noise(x, y, (int)A.numFrames, 0, A.sobol, A.scr), ry = sample_blue_noise(x, y, (int)A.numFrames, 1, A.sobol, A.scr);
        f3    lightDir = {0.0f, 0.0f, 0.0f};
        float lightRadius = 0.0f, tMax = 0.0f, attenuation = 0.0f;
        const f3 lightPos = {L.position[0], L.position[1], L.position[2]};
        if (L.type == LUX_LIGHT_DIRECTIONAL)
        {
            lightDir    = {-L.direction[0], -L.direction[1], -L.direction[2]};
            tMax        = LUX_GLOBAL_SDF_WORLD_SIZE;
            lightRadius = L.direction[3];
            attenuation = 1.0f;
        }
        else if (L.type == LUX_LIGHT_POINT)
        {
            f3    dir  = lightPos - worldPos;
            float dist = length3(dir);
            lightDir    = normalize3(dir);
            attenuation = 1.0f;
            tMax        = dist;
            lightRadius = __fdiv_rn(L.direction[3], dist);
        }
        else if (L.type == LUX_LIGHT_SPOT)
        {
            f3    Lv          = lightPos - worldPos;
            float cutoffAngle = 1.0f - L.angle;
            lightDir          = normalize3(Lv);
            float dist        = length3(Lv);
            float theta       = dot3(lightDir, {L.direction[0], L.direction[1], L.direction[2]});
            float epsilon     = cutoffAngle - cutoffAngle * 0.9f;
            attenuation       = __fdiv_rn(theta - cutoffAngle, epsilon);
            attenuation *= __fdiv_rn(L.radius, powd_rn(dist, 2.0f) + 1.0f);
            attenuation = gclamp(attenuation, 0.0f, 1.0f);
            tMax        = dist;
            lightRadius = __fdiv_rn(L.direction[3], dist);
        }
        const f3    lightTangent   = normalize3(cross3(lightDir, {0.0f, 1.0f, 0.0f}));
        const f3    lightBitangent = normalize3(cross3(lightTangent, lightDir));
        const float pointRadius = lightRadius * __fsqrt_rn(rx);
        const float pointAngle  = (ry * 2.0f) * PI_F;
        const float dx = pointRadius * cos_rn(pointAngle), dy = pointRadius * sin_rn(pointAngle);
        const f3    wv = {(lightDir.x + dx * lightTangent.x) + dy * lightBitangent.x, (lightDir.y + dx * lightTangent.y) + dy * lightBitangent.y,
                          (lightDir.z + dx * lightTangent.z) + dy * lightBitangent.z};
        const f3    Wi = normalize3(wv);
        attenuation *= gclamp(dot3(normal, Wi), 0.0f, 1.0f);
        const float NoL = dot3(normal, Wi);
        if (NoL > 0.0f && attenuation > 0.0f)
        {
            const float bias   = (2.0f * A.shadowBias) * gclamp(1.0f - NoL, 0.0f, 1.0f) + A.shadowBias;
            const f3    origin = worldPos + normal * A.shadowBias;
            LuxGlobalSDFTrace tr;
            tr.worldPosition[0] = origin.x; tr.worldPosition[1] = origin.y; tr.worldPosition[2] = origin.z;
            tr.minDistance = bias;
            tr.worldDirection[0] = Wi.x; tr.worldDirection[1] = Wi.y; tr.worldDirection[2] = Wi.z;
            tr.maxDistance = tMax - bias;
            tr.stepScale = 1.0f;
            tr.needsHitNormal = 0u;
            const LuxGlobalSDFHit hit = trace_global_sdf_general<TEX>(P, sdf, tr, 0.0f);
            result = hit.hitTime >= 0.0f ? 0u : 1u;
        }
    }
    const uint32_t mask   = __ballot_sync(0xffffffffu, result != 0u);
    const bool     stored = __shfl_sync(0xffffffffu, live ? 1 : 0, 0) != 0; // the shader's store sits inside lane 0's `depth != 1` branch
    if (li == 0 && stored)
        A.out[g] = mask;
}

// =====================================================================================================================
// Global SDF build: mesh distance fields -> cascade volume -> min-mip.
//   sdf_object_data_kernel   chunkCalculate's ObjectRasterizeData (GlobalDistanceField.cpp:484-508), one thread per mesh
//   sdf_rasterize_kernel     SDFRasterizeModel.glsl:42-63 + SDFCommon.glsl:18-62; one block = one 8x8x8 group of a chunk dispatch,
//                            all chunk dispatches of a pipeline pass in ONE launch (chunks are independent)
//   sdf_mip_kernel           GlobalSDFMipmap.comp:32-68
// Mesh volumes are sampled in software (8 fp16 loads, fp32 nested lerps, REPEAT addressing at one integer mip level): the
// texture unit's 8-bit filter weights would break parity, as for the global SDF.
// =====================================================================================================================
__global__ void sdf_object_data_kernel(const SdfMeshRecord* __restrict__ meshes, int count, int cascadeLevel, LuxObjectRasterizeData* __restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count)
        return;
    const SdfMeshRecord& ms = meshes[k];
    LuxObjectRasterizeData o;
    f3 mn = {ms.aabbMin[0], ms.aabbMin[1], ms.aabbMin[2]}, mx = {ms.aabbMax[0], ms.aabbMax[1], ms.aabbMax[2]};
    f3 volumeCenter = (mx + mn) * 0.5f;
    float worldToLocal[16];
    inverse4_dev(ms.worldMatrix, worldToLocal);
    const float tr[16] = {1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, -volumeCenter.x, -volumeCenter.y, -volumeCenter.z, 1.0f};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            o.worldToVolume[c * 4 + r] = ((worldToLocal[0 * 4 + r] * tr[c * 4 + 0] + worldToLocal[1 * 4 + r] * tr[c * 4 + 1]) + worldToLocal[2 * 4 + r] * tr[c * 4 + 2]) +
                                         worldToLocal[3 * 4 + r] * tr[c * 4 + 3];
    inverse4_dev(o.worldToVolume, o.volumeToWorld);
    f3 size = mx - mn;
    o.volumeLocalBoundsExtent[0] = __fdiv_rn(size.x, 2.0f);
    o.volumeLocalBoundsExtent[1] = __fdiv_rn(size.y, 2.0f);
    o.volumeLocalBoundsExtent[2] = __fdiv_rn(size.z, 2.0f);
    const float vc[3] = {volumeCenter.x, volumeCenter.y, volumeCenter.z};
    for (int i = 0; i < 3; i++)
    {
        o.volumeToUVWMul[i] = ms.localToUVWMul[i];
        o.volumeToUVWAdd[i] = ms.localToUVWAdd[i] + vc[i] * ms.localToUVWMul[i];
    }
    o.mipOffset = (float)min(cascadeLevel, 2);
    o.decodeMul = ms.maxDistance;
    o.decodeAdd = -ms.maxDistance;
    out[k] = o;
}

__device__ __forceinline__ int wrap_repeat(int i, int n) { return ((i % n) + n) % n; }

__device__ __forceinline__ float sample_mesh_sdf(const uint16_t* __restrict__ d, int W, int H, int D, float u, float v, float w)
{
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f, z = w * (float)D - 0.5f;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float ax = x - fx, ay = y - fy, az = z - fz;
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;
    int   x0 = wrap_repeat(ix, W), x1 = wrap_repeat(ix + 1, W), y0 = wrap_repeat(iy, H), y1 = wrap_repeat(iy + 1, H);
    int   z0 = wrap_repeat(iz, D), z1 = wrap_repeat(iz + 1, D);
    const uint16_t* p00 = d + ((size_t)z0 * H + y0) * W;
    const uint16_t* p10 = d + ((size_t)z0 * H + y1) * W;
    const uint16_t* p01 = d + ((size_t)z1 * H + y0) * W;
    const uint16_t* p11 = d + ((size_t)z1 * H + y1) * W;
    float c00 = lerp1(ld_h(p00 + x0), ld_h(p00 + x1), ax), c10 = lerp1(ld_h(p10 + x0), ld_h(p10 + x1), ax);
    float c01 = lerp1(ld_h(p01 + x0), ld_h(p01 + x1), ax), c11 = lerp1(ld_h(p11 + x0), ld_h(p11 + x1), ax);
    return lerp1(lerp1(c00, c10, ay), lerp1(c01, c11, ay), az);
}

__device__ __forceinline__ float combine_distance_to_sdf(float sdf, float distanceToSDF)
{
    if (sdf <= 0.0f && distanceToSDF <= 0.0f)
        return sdf;
    float maxSDF = gmax(sdf, 0.0f);
    return __fsqrt_rn(maxSDF * maxSDF + distanceToSDF * distanceToSDF);
}

__global__ void __launch_bounds__(512) sdf_rasterize_kernel(const __grid_constant__ SdfRasterizeParams P)
{
    __shared__ SdfChunkDispatch sd;
    const SdfChunkDispatch& gd = P.dispatches[blockIdx.x >> 6];
    if (threadIdx.x < sizeof(SdfChunkDispatch) / 4)
        reinterpret_cast<uint32_t*>(&sd)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&gd)[threadIdx.x];
    __syncthreads();
    const int group = blockIdx.x & 63; // 4x4x4 groups of 8x8x8 voxels per chunk
    const int x = ((group & 3) << 3) + (threadIdx.x & 7), y = (((group >> 2) & 3) << 3) + ((threadIdx.x >> 3) & 7), z = ((group >> 4) << 3) + (threadIdx.x >> 6);
    const int vx = sd.coord[0] + x, vy = sd.coord[1] + y, vz = sd.coord[2] + z;
    if (vx < 0 || vy < 0 || vz < 0 || vx >= P.res || vy >= P.res || vz >= P.res)
        return; // out-of-bounds image access
    f3 worldPos = {(float)vx * P.mul[0] + P.add[0], (float)vy * P.mul[1] + P.add[1], (float)vz * P.mul[2] + P.add[2]};
    const size_t o = ((size_t)vz * P.res + vy) * P.texWidth + (vx + P.cascadeIndex * P.res);
    float minDistance = P.maxDistance;
    if (sd.read)
        minDistance *= h2f_bits(P.sdf[o]);
    for (int i = 0; i < sd.count; i++)
    {
        const uint32_t id = sd.models[i];
        const LuxObjectRasterizeData& m = P.objects[id];
        float w2v[16];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            float4 c = __ldg(reinterpret_cast<const float4*>(m.worldToVolume) + k);
            w2v[k * 4 + 0] = c.x; w2v[k * 4 + 1] = c.y; w2v[k * 4 + 2] = c.z; w2v[k * 4 + 3] = c.w;
        }
        f3 volumePos = mat4_mul_point(w2v, worldPos, 1.0f);
        f3 e = {__ldg(m.volumeLocalBoundsExtent + 0), __ldg(m.volumeLocalBoundsExtent + 1), __ldg(m.volumeLocalBoundsExtent + 2)};
        f3 clamped = {gclamp(volumePos.x, -e.x, e.x), gclamp(volumePos.y, -e.y, e.y), gclamp(volumePos.z, -e.z, e.z)};
        float v2w[16];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            float4 c = __ldg(reinterpret_cast<const float4*>(m.volumeToWorld) + k);
            v2w[k * 4 + 0] = c.x; v2w[k * 4 + 1] = c.y; v2w[k * 4 + 2] = c.z; v2w[k * 4 + 3] = c.w;
        }
        f3    worldPosClamped  = mat4_mul_point(v2w, clamped, 1.0f);
        float distanceToVolume = length3(worldPos - worldPosClamped);
        if (distanceToVolume < 0.01f)
            distanceToVolume = length3(volumePos - clamped);
        distanceToVolume = gmax(distanceToVolume, 0.0f);
        float objectDistance = distanceToVolume;
        if (!(minDistance <= distanceToVolume))
        {
            const SdfMeshLevel lv = P.levels[id];
            float u = volumePos.x * __ldg(m.volumeToUVWMul + 0) + __ldg(m.volumeToUVWAdd + 0);
            float v = volumePos.y * __ldg(m.volumeToUVWMul + 1) + __ldg(m.volumeToUVWAdd + 1);
            float w = volumePos.z * __ldg(m.volumeToUVWMul + 2) + __ldg(m.volumeToUVWAdd + 2);
            float volumeDistance = (sample_mesh_sdf(lv.data, lv.w, lv.h, lv.d, u, v, w) * 2.0f - 1.0f) * __ldg(&m.decodeMul);
            float result = combine_distance_to_sdf(volumeDistance, distanceToVolume);
            if (distanceToVolume > 0.0f)
                result = gmax(distanceToVolume, result);
            objectDistance = result;
        }
        minDistance = gmin(minDistance, objectDistance);
    }
    P.sdf[o] = f2h_bits(gclamp(__fdiv_rn(minDistance, P.maxDistance), -1.0f, 1.0f));
}

__global__ void sdf_fill_kernel(uint16_t* __restrict__ p, size_t n, uint16_t v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void __launch_bounds__(64) sdf_mip_kernel(const uint16_t* __restrict__ src, int srcWidth, int srcHeight, uint16_t* __restrict__ dst, int dstWidth,
                                                     int dstHeight, int outRes, int globalSDFResolution, int mipmapCoordScale, int cascadeTexOffsetX,
                                                     int cascadeMipMapOffsetX, float maxDistance)
{
    const int x = blockIdx.x * 4 + (threadIdx.x & 3), y = blockIdx.y * 4 + ((threadIdx.x >> 2) & 3), z = blockIdx.z * 4 + (threadIdx.x >> 4);
    if (x >= outRes || y >= outRes || z >= outRes)
        return;
    const int off[7][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
    float minDistance = 0.0f;
#pragma unroll
    for (int k = 0; k < 7; k++)
    {
        int cx = iclamp(x * mipmapCoordScale + off[k][0], 0, globalSDFResolution - 1);
        int cy = iclamp(y * mipmapCoordScale + off[k][1], 0, globalSDFResolution - 1);
        int cz = iclamp(z * mipmapCoordScale + off[k][2], 0, globalSDFResolution - 1);
        float result = h2f_bits(src[((size_t)cz * srcHeight + cy) * srcWidth + cx + cascadeTexOffsetX]);
        float len = length3({(float)off[k][0], (float)off[k][1], (float)off[k][2]});
        float distanceToVoxel = len * __fdiv_rn(maxDistance, (float)globalSDFResolution);
        result = combine_distance_to_sdf(result, distanceToVoxel);
        minDistance = k == 0 ? result : gmin(minDistance, result);
    }
    dst[((size_t)z * dstHeight + y) * dstWidth + x + cascadeMipMapOffsetX] = f2h_bits(minDistance);
}

void launch_sdf_object_data(const SdfMeshRecord* meshes, int count, int cascadeLevel, LuxObjectRasterizeData* out, cudaStream_t s)
{
    if (count > 0)
        sdf_object_data_kernel<<<(count + 63) / 64, 64, 0, s>>>(meshes, count, cascadeLevel, out);
}

void launch_sdf_rasterize(const SdfRasterizeParams& p, int dispatchCount, cudaStream_t s)
{
    if (dispatchCount > 0)
        sdf_rasterize_kernel<<<(unsigned)dispatchCount * 64u, 512, 0, s>>>(p);
}

void launch_sdf_fill(uint16_t* p, size_t n, uint16_t value, cudaStream_t s) { sdf_fill_kernel<<<148 * 8, 256, 0, s>>>(p, n, value); }

void launch_sdf_mip_pass(const uint16_t* src, int srcWidth, int srcHeight, uint16_t* dst, int dstWidth, int dstHeight, int outRes, int globalSDFResolution,
                         int mipmapCoordScale, int cascadeTexOffsetX, int cascadeMipMapOffsetX, float maxDistance, cudaStream_t s)
{
    const unsigned g = (unsigned)(outRes + 3) / 4;
    sdf_mip_kernel<<<dim3(g, g, g), 64, 0, s>>>(src, srcWidth, srcHeight, dst, dstWidth, dstHeight, outRes, globalSDFResolution, mipmapCoordScale,
                                                cascadeTexOffsetX, cascadeMipMapOffsetX, maxDistance);
}

// =====================================================================================================================
// Surface-cache culling: SDFCulling.comp:36-101, one thread per culling chunk.
// The shader allocates list space with a returning atomic, which makes the layout of the cull buffer depend on scheduling.  Here
// the same lists are laid out deterministically in ascending chunk address: count -> exclusive scan of (count + 1) -> fill.  The
// capacity rule is the shader's (the counter advances even for a list that then does not fit).  Element 0 of the chunk buffer only
// ever holds chunk 0's own list start: the shader's `atlasChunks.data[0] = chunkAddress` store of empty chunks is a data race in
// the reference with no defined outcome and would let hits in chunk 0 read an arbitrary list.
// =====================================================================================================================
__device__ __forceinline__ bool chunk_intersects_object(const LuxObjectBuffer* __restrict__ objects, uint32_t i, f3 chunkMin, f3 chunkMax)
{
    float4 b  = __ldg(reinterpret_cast<const float4*>(objects[i].objectBounds));
    f3     c  = {b.x, b.y, b.z};
    f3     cl = {gclamp(c.x, chunkMin.x, chunkMax.x), gclamp(c.y, chunkMin.y, chunkMax.y), gclamp(c.z, chunkMin.z, chunkMax.z)};
    return length3(c - cl) <= b.w;
}

__device__ __forceinline__ void cull_chunk_bounds(int chunk, float chunkSize, f3& mn, f3& mx)
{
    const int   N    = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    const float half = (float)N * 0.5f;
    int cx = chunk % N, cy = (chunk / N) % N, cz = chunk / (N * N);
    mn = {((float)cx - half) * chunkSize, ((float)cy - half) * chunkSize, ((float)cz - half) * chunkSize};
    mx = {mn.x + chunkSize, mn.y + chunkSize, mn.z + chunkSize};
}

// sizes[chunk] = objectsCount + 1, or 0 for an empty chunk (padded entries stay 0)
__global__ void cull_count_kernel(const LuxObjectBuffer* __restrict__ objects, uint32_t objectsCount, float chunkSize, uint32_t* __restrict__ sizes)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    if (chunk >= total)
        return;
    f3 mn, mx;
    cull_chunk_bounds(chunk, chunkSize, mn, mx);
    uint32_t n = 0;
    for (uint32_t i = 0; i < objectsCount; i++)
        n += chunk_intersects_object(objects, i, mn, mx) ? 1u : 0u;
    sizes[chunk] = n ? n + 1u : 0u;
}

// prefix = exclusive scan of sizes (in place); *totalWords = sum.  cull[0] = 1 + sum is the shader's final counter.
__global__ void cull_fill_kernel(const LuxObjectBuffer* __restrict__ objects, uint32_t objectsCount, float chunkSize, uint32_t capacity,
                                 const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ totalWords, uint32_t* __restrict__ chunks,
                                 uint32_t* __restrict__ cull, uint32_t cullWords)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    if (chunk >= total)
        return;
    if (chunk == 0)
        cull[0] = 1u + *totalWords;
    f3 mn, mx;
    cull_chunk_bounds(chunk, chunkSize, mn, mx);
    uint32_t n = 0;
    for (uint32_t i = 0; i < objectsCount; i++)
        n += chunk_intersects_object(objects, i, mn, mx) ? 1u : 0u;
    uint32_t start = 1u + prefix[chunk];
    if (n == 0 || start + n + 1u > capacity || start + n + 1u > cullWords)
    {
        chunks[chunk] = 0u;
        return;
    }
    chunks[chunk] = start;
    cull[start]   = n;
    for (uint32_t i = 0; i < objectsCount; i++)
        if (chunk_intersects_object(objects, i, mn, mx))
            cull[++start] = i;
}

void launch_surface_cull(const LuxObjectBuffer* objects, uint32_t objectsCount, float chunkSize, uint32_t capacity, uint32_t* sizesPadded,
                         uint32_t* blockSums, uint32_t* totalWords, uint32_t* chunks, uint32_t* cull, uint32_t cullWords, cudaStream_t s)
{
    const int total = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    const int nb    = (total + SCAN_BINS_PER_BLOCK - 1) / SCAN_BINS_PER_BLOCK;
    cudaMemsetAsync(sizesPadded, 0, (size_t)nb * SCAN_BINS_PER_BLOCK * sizeof(uint32_t), s);
    cull_count_kernel<<<(total + 127) / 128, 128, 0, s>>>(objects, objectsCount, chunkSize, sizesPadded);
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>((const uint4*)sizesPadded, blockSums);
    scan_top_kernel<<<1, SCAN_THREADS, 0, s>>>(blockSums, nb, totalWords);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, s>>>((uint4*)sizesPadded, blockSums);
    cull_fill_kernel<<<(total + 127) / 128, 128, 0, s>>>(objects, objectsCount, chunkSize, capacity, sizesPadded, totalWords, chunks, cull, cullWords);
}

void launch_sdf_rays(const TraceParams& p, bool useTextures, int count, const LuxGlobalSDFTrace* traces, float cascadeTraceStartBias,
                     LuxGlobalSDFHit* hits, cudaStream_t s)
{
    if (count <= 0)
        return;
    if (useTextures)
        sdf_rays_kernel<true><<<(count + 127) / 128, 128, 0, s>>>(p, count, traces, cascadeTraceStartBias, hits);
    else
        sdf_rays_kernel<false><<<(count + 127) / 128, 128, 0, s>>>(p, count, traces, cascadeTraceStartBias, hits);
}

void launch_sdf_reflection(const TraceParams& p, bool useTextures, const LuxDDGIUniform& ddgi, const LuxReflectionPushConstants& push, const void* irr,
                           const void* dep, int width, int height, const float* gDepth, const float* gNormal, const float* gPbr, const uint32_t* sobol,
                           const uint32_t* scr, void* out, cudaStream_t s)
{
    const int px = width * height;
    if (px <= 0)
        return;
    ReflectionArgs a{ddgi, push, (const uint16_t*)irr, (const uint16_t*)dep, width, height, gDepth, gNormal, gPbr, sobol, scr, (uint2*)out};
    if (useTextures)
        sdf_reflection_kernel<true><<<(px + 127) / 128, 128, 0, s>>>(p, a);
    else
        sdf_reflection_kernel<false><<<(px + 127) / 128, 128, 0, s>>>(p, a);
}

void launch_sdf_shadow(const TraceParams& p, bool useTextures, const LuxLight& light, const float* viewProjInv, uint32_t numFrames, float shadowBias,
                       int width, int height, const float* gDepth, const float* gNormal, const uint32_t* sobol, const uint32_t* scr, uint32_t* out,
                       cudaStream_t s)
{
    const int groups = (width / 8) * (height / 4);
    if (groups <= 0)
        return;
    ShadowArgs a{};
    a.light = light;
    for (int i = 0; i < 16; i++)
        a.viewProjInv[i] = viewProjInv[i];
    a.numFrames = numFrames; a.shadowBias = shadowBias; a.width = width; a.height = height;
    a.gDepth = gDepth; a.gNormal = gNormal; a.sobol = sobol; a.scr = scr; a.out = out;
    if (useTextures)
        sdf_shadow_kernel<true><<<(groups + 3) / 4, 128, 0, s>>>(p, a);
    else
        sdf_shadow_kernel<false><<<(groups + 3) / 4, 128, 0, s>>>(p, a);
}

// L2 read bandwidth (the denominator of the request-level roofline): every block sweeps the whole buffer once, from a
// block-dependent start so that the SMs are spread over the buffer at any moment, with ld.global.cg (no L1 allocation) and four independent
// 16-byte loads in flight per thread.  After the first touch every line is an L2 hit as long as the buffer fits in L2.
__global__ void __launch_bounds__(256) l2_sweep_kernel(const uint4* __restrict__ buf, unsigned int n16, uint32_t* __restrict__ sink)
{
    const unsigned int start = (unsigned int)(((unsigned long long)blockIdx.x * n16) / gridDim.x) & ~1023u;
    uint32_t acc = 0;
    for (unsigned int base = 0; base < n16; base += 1024)
    {
        unsigned int i = start + base + threadIdx.x;
        if (i >= n16)
            i -= n16;
        // n16 is a multiple of 1024, so the four loads of one thread stay inside [0, n16)
        uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + 256), c = __ldcg(buf + i + 512), d = __ldcg(buf + i + 768);
        acc ^= a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0x9e3779b9u) // never true for the zero-filled buffer; keeps the loads alive
        sink[0] = acc;
}

void launch_l2_sweep(const void* buf, size_t bytes, int blocks, uint32_t* sink, cudaStream_t s)
{
    l2_sweep_kernel<<<blocks, 256, 0, s>>>((const uint4*)buf, (unsigned int)(bytes / 16), sink);
}

void launch_direct_light(const TraceParams& p, bool useTextures, const LuxLight& l, const float* cameraPosBias, void* light, int count,
                         const uint32_t* texel, const float* P, const float* N, const float* albedo, const float* metallicRoughness, cudaStream_t s)
{
    if (count <= 0)
        return;
    const int grid = (count + 127) / 128;
    if (useTextures)
        direct_light_kernel<true><<<grid, 128, 0, s>>>(p, l, cameraPosBias[0], cameraPosBias[1], cameraPosBias[2], cameraPosBias[3], (uint2*)light, count, texel, P,
                                                        N, albedo, metallicRoughness);
    else
        direct_light_kernel<false><<<grid, 128, 0, s>>>(p, l, cameraPosBias[0], cameraPosBias[1], cameraPosBias[2], cameraPosBias[3], (uint2*)light, count, texel, P,
                                                         N, albedo, metallicRoughness);
}

void launch_indirect_light(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, void* light, const void* base, int count,
                           const uint32_t* texel, const float* P, const float* N, const float* albedo, const float* metallic, float intensity,
                           const float* cameraPos, cudaStream_t s)
{
    if (count > 0)
        indirect_light_kernel<<<(count + 127) / 128, 128, 0, s>>>(ddgi, (const uint16_t*)irr, (const uint16_t*)dep, (uint2*)light, (const uint2*)base, count,
                                                                  texel, P, N, albedo, metallic, intensity, cameraPos[0], cameraPos[1], cameraPos[2]);
}

void launch_sample_irradiance(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, int count, const float* P, const float* N,
                              const float* Wo, float* out, cudaStream_t s)
{
    if (count > 0)
        sample_irradiance_kernel<<<(count + 127) / 128, 128, 0, s>>>(ddgi, (const uint16_t*)irr, (const uint16_t*)dep, count, P, N, Wo, out);
}

void launch_sample_probe(const LuxDDGIUniform& ddgi, const void* irr, const void* dep, int width, int height, const float* gDepth,
                         const float* gNormal, const float* cameraPosition, const float* viewProjInv, float* out, cudaStream_t s)
{
    SampleProbeArgs a;
    a.ddgi = ddgi;
    for (int i = 0; i < 4; i++) a.cameraPosition[i] = cameraPosition[i];
    for (int i = 0; i < 16; i++) a.viewProjInv[i] = viewProjInv[i];
    a.width = width;
    a.height = height;
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    sample_probe_kernel<<<grid, block, 0, s>>>(a, (const uint16_t*)irr, (const uint16_t*)dep, gDepth, (const float4*)gNormal, (float4*)out);
}

// =====================================================================================================================
// Launchers
// =====================================================================================================================

void launch_ray_dirs(const float* rot16Host, int R, float4* dirs, uint2* dirsHalf, cudaStream_t s)
{
    RotationArg rot;
    for (int i = 0; i < 16; i++)
        rot.m[i] = rot16Host[i];
    ray_dirs_kernel<<<(R + 127) / 128, 128, 0, s>>>(rot, R, dirs, dirsHalf);
}

int launch_blend_weights(const uint2* dirsHalf, int R, int Rpad, float sharpness, float* wIrr, float* wDepth, float* scaleIrr,
                         float* scaleDepth, uint32_t* nzIrr, uint32_t* nzDepth, cudaStream_t s)
{
    blend_weights_kernel<<<dim3(5, Rpad), 64, 0, s>>>(dirsHalf, R, Rpad, sharpness, wIrr, wDepth);
    blend_scales_kernel<<<5, 64, 0, s>>>(wIrr, wDepth, R, scaleIrr, scaleDepth);
    blend_nonzero_kernel<<<(Rpad + 63) / 64, 64, 0, s>>>(wIrr, wDepth, Rpad, nzIrr, nzDepth);
    return 3;
}

// third row + translation of every tile transform, (m[2], m[6], m[10], m[14]): what the sign of a tile's normal weight depends on
__global__ void tile_zrow_kernel(const LuxTileBuffer* __restrict__ tiles, int count, float4* __restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count)
        out[k] = make_float4(tiles[k].transform[2], tiles[k].transform[6], tiles[k].transform[10], tiles[k].transform[14]);
}

void launch_tile_zrow(const LuxTileBuffer* tiles, int count, float4* out, cudaStream_t s)
{
    if (count > 0)
        tile_zrow_kernel<<<(count + 127) / 128, 128, 0, s>>>(tiles, count, out);
}

void launch_object_inverse(const LuxObjectBuffer* objects, int count, float* inv, cudaStream_t s)
{
    if (count > 0)
        object_inverse_kernel<<<(count + 127) / 128, 128, 0, s>>>(objects, count, inv);
}

void launch_chunk_masks(const uint32_t* chunks, const uint32_t* cull, const LuxObjectBuffer* objects, const float* objectInverse,
                        uint32_t objectsCount, float chunkSize, float thrMax, unsigned long long* masks, cudaStream_t s)
{
    const int N = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION;
    chunk_masks_kernel<<<N * N * N, 64, 0, s>>>(chunks, cull, objects, objectInverse, objectsCount, chunkSize, thrMax, masks);
}

size_t trace_record_count(int probeCount, int raysPerProbe, bool beam)
{
    const size_t slots = (size_t)(raysPerProbe + MARCH_CLUSTER_RAYS - 1) / MARCH_CLUSTER_RAYS * MARCH_CLUSTER_RAYS;
    const size_t unit  = beam ? 2 : 32;
    return ((size_t)probeCount + unit - 1) / unit * unit * slots; // chunks x 64
}
size_t trace_record_capacity(int probeCount, int raysPerProbe) { return trace_record_count(probeCount, raysPerProbe, false); }
size_t trace_shade_pool_bytes() { return (size_t)SHADE_POOL_SMS * (1 + SHADE_POOL_SLOTS) * sizeof(unsigned long long); }

void launch_probe_origins(const TraceParams& p, cudaStream_t s)
{
    probe_origins_kernel<<<(p.probeCount + 255) / 256, 256, 0, s>>>(p, const_cast<float4*>(p.origins));
}

size_t trace_sort_bins() { return ((size_t)SORT_BINS + SCAN_BINS_PER_BLOCK - 1) / SCAN_BINS_PER_BLOCK * SCAN_BINS_PER_BLOCK; }
size_t trace_sort_blocks() { return trace_sort_bins() / SCAN_BINS_PER_BLOCK; }

// classify, scan -> scatter -> shade_sorted (see the comment above classify_kernel); returns the number of launches
template <bool TEX>
static int launch_shade_sorted(const TraceParams& p, cudaStream_t s)
{
    const size_t records = trace_record_count(p.probeCount, p.raysPerProbe, p.beam != 0);
    const size_t rays    = (size_t)p.probeCount * p.raysPerProbe;
    const int    nb      = (int)trace_sort_blocks();
    if (p.beam)
        classify_kernel<<<(unsigned)((rays + 256 * CLASSIFY_RPT - 1) / (256 * CLASSIFY_RPT)), 256, 0, s>>>(p);
    else
    {
        const int rayGroups = (p.raysPerProbe + CLASSIFY_ROW_RAYS - 1) / CLASSIFY_ROW_RAYS;
        classify_rows_kernel<<<(unsigned)((size_t)rayGroups * ((p.probeCount + 31) / 32)), 32 * CLASSIFY_ROW_WARPS, 0, s>>>(p, rayGroups);
    }
    if (!p.hasAtlas)
        return 1; // hits carry no radiance without a surface cache: classify wrote the final values
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>((const uint4*)p.binCounts, p.binBlockSums);
    scan_top_kernel<<<1, SCAN_THREADS, 0, s>>>(p.binBlockSums, nb, p.hitCount);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, s>>>((uint4*)p.binCounts, p.binBlockSums);
    scatter_kernel<<<(unsigned)((records + 256 * SCATTER_RPT - 1) / (256 * SCATTER_RPT)), 256, 0, s>>>(p.sortTicket, p.binCounts, p.sortedIdx, records);
    long long blocks = (long long)(records + 255) / 256;
    if (blocks > 148ll * SHADE_BLOCKS_PER_SM)
        blocks = 148ll * SHADE_BLOCKS_PER_SM;
    cudaMemsetAsync(p.shadePools, 0, trace_shade_pool_bytes(), s);
    cudaMemsetAsync(p.shadeCounter, 0, sizeof(unsigned int), s);
    shade_sorted_kernel<TEX><<<(unsigned)blocks, 256, 0, s>>>(p);
    return 6;
}

// x / d as ExactDivisor does it on the device: the reciprocal when d is a power of two (then x * (1/d) is the correctly rounded quotient), else 0
static float exact_reciprocal(float d)
{
    uint32_t b;
    memcpy(&b, &d, 4);
    const uint32_t e = (b >> 23) & 0xffu;
    return (((b & 0x7fffffu) == 0u) && e > 2u && e < 252u) ? 1.0f / d : 0.0f;
}
// TraceParams::mc (host float arithmetic is IEEE binary32 here: no contraction, no extended precision; see build.py)
static void march_consts(TraceParams& p)
{
    const LuxGlobalSDFData& d = p.sdf;
    TraceParams::MarchConsts& m = p.mc;
    const float last2 = d.cascadePosDistance[d.cascadesCount - 1][3] * 2.0f;
    m.traceMaxDistance     = last2 < LUX_GLOBAL_SDF_WORLD_SIZE ? last2 : LUX_GLOBAL_SDF_WORLD_SIZE; // gmin(WORLD_SIZE, last2)
    m.chunkSizeDistance    = (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE / d.resolution;
    m.chunkMarginDistance2 = ((float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_MARGIN / d.resolution) * 2.0f;
    m.cascadesCountF       = (float)d.cascadesCount;
    m.cascadesInv          = exact_reciprocal(m.cascadesCountF);
    m.mipW = (float)(p.mipRes * p.cascades); m.mipH = (float)p.mipRes;
    m.texW = (float)(p.res * p.cascades);    m.texH = (float)p.res;
    m.mipDm1 = p.mipRes - 1; m.texDm1 = p.res - 1;
    for (int i = 0; i < 3; i++)
        m.cc0[i] = d.cascadePosDistance[0][i];
    m.cd0 = d.cascadePosDistance[0][3];
    m.m0 = m.cd0 * 2.0f; m.minv0 = exact_reciprocal(m.m0);
    m.v0 = d.cascadeVoxelSize[0]; m.vinv0 = exact_reciprocal(m.v0);
}

void launch_probe_taps(const TraceParams& pIn, bool useTextures, float2* taps, cudaStream_t s)
{
    TraceParams p = pIn;
    march_consts(p);
    const int grid = (p.probeCount + 127) / 128;
    if (useTextures)
        probe_taps_kernel<true><<<grid, 128, 0, s>>>(p, taps);
    else
        probe_taps_kernel<false><<<grid, 128, 0, s>>>(p, taps);
}

template <bool TEX>
static int launch_wavefront(const TraceParams& pIn, unsigned int* chunkCounter, cudaStream_t s, cudaEvent_t beforeShade, cudaEvent_t afterMarch)
{
    TraceParams p = pIn;
    march_consts(p);
    const int chunks = (int)(trace_record_count(p.probeCount, p.raysPerProbe, p.beam != 0) / MARCH_CHUNK_RAYS);
    cudaMemsetAsync(chunkCounter, 0, sizeof(unsigned int), s);
    if (p.sortTicket)
        cudaMemsetAsync(p.binCounts, 0, trace_sort_bins() * sizeof(uint32_t), s);
    long long blocks = ((long long)chunks + MARCH_WARPS - 1) / MARCH_WARPS;
    const long long persistent = 148ll * MARCH_BLOCKS_PER_SM; // one resident generation per SM
    if (blocks > persistent)
        blocks = persistent;
    if (p.cascades > 1)
    {
        if (p.beam) march_kernel<TEX, true, true><<<(unsigned)blocks, 32 * MARCH_WARPS, 0, s>>>(p, chunks, chunkCounter);
        else        march_kernel<TEX, true, false><<<(unsigned)blocks, 32 * MARCH_WARPS, 0, s>>>(p, chunks, chunkCounter);
    }
    else
    {
        if (p.beam) march_kernel<TEX, false, true><<<(unsigned)blocks, 32 * MARCH_WARPS, 0, s>>>(p, chunks, chunkCounter);
        else        march_kernel<TEX, false, false><<<(unsigned)blocks, 32 * MARCH_WARPS, 0, s>>>(p, chunks, chunkCounter);
    }
    if (afterMarch)
        cudaEventRecord(afterMarch, s);
    if (beforeShade) // the march never reads the surface cache: a pending light-cache upload only gates the shade
        cudaStreamWaitEvent(s, beforeShade, 0);
    if (p.sortedIdx)
        return 1 + launch_shade_sorted<TEX>(p, s);
    const size_t rays = (size_t)p.probeCount * p.raysPerProbe;
    shade_kernel<TEX><<<(unsigned)((rays + 255) / 256), 256, 0, s>>>(p);
    return 2;
}

int launch_trace(const TraceParams& p, int variant, unsigned int* chunkCounter, cudaStream_t s, cudaEvent_t beforeShade, cudaEvent_t afterMarch)
{
    if (variant == 0)
    {
        dim3 block(32, TRACE_RAYS_PER_BLOCK);
        dim3 grid((p.raysPerProbe + TRACE_RAYS_PER_BLOCK - 1) / TRACE_RAYS_PER_BLOCK, (p.probeCount + 31) / 32);
        if (beforeShade)
            cudaStreamWaitEvent(s, beforeShade, 0);
        trace_kernel<<<grid, block, 0, s>>>(p);
        return 1;
    }
    return variant == 2 ? launch_wavefront<true>(p, chunkCounter, s, beforeShade, afterMarch) : launch_wavefront<false>(p, chunkCounter, s, beforeShade, afterMarch);
}

template <int PB>
static void launch_blend_irradiance_t(const BlendParams& p, cudaStream_t s)
{
    constexpr int M = 3 * PB;
    size_t main = (size_t)(KC * (M + 2) + 2 * KC * 64 + 2 * KC) * sizeof(float) + (size_t)2 * PB * KC * sizeof(uint2);
    size_t epi  = (size_t)M * 64 * sizeof(float);
    size_t smem = main > epi ? main : epi;
    // per launch: the attribute is per device (and per context), a process-wide "done" flag would leave a second GPU's context without it
    cudaFuncSetAttribute(blend_irradiance_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    blend_irradiance_kernel<PB><<<(p.probeCount + PB - 1) / PB, 256, smem, s>>>(p);
}

void launch_blend_irradiance(const BlendParams& p, cudaStream_t s)
{
    // 64-probe tiles only when they still give every SM two blocks; measured on a 1/8 C4 shard (8192 probes): 0.264 -> 0.229 ms
    if (p.probeCount >= 2 * 148 * 64)
        launch_blend_irradiance_t<64>(p, s);
    else
        launch_blend_irradiance_t<32>(p, s);
}

template <int PB>
static void launch_blend_depth_t(const BlendParams& p, cudaStream_t s)
{
    constexpr int M = 2 * PB;
    size_t main = (size_t)(KC * (M + 4) + 2 * KC * 256 + 2 * KC) * sizeof(float) + (size_t)2 * PB * KC * sizeof(uint2);
    size_t epi  = (size_t)M * 256 * sizeof(float);
    size_t smem = main > epi ? main : epi;
    cudaFuncSetAttribute(blend_depth_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    blend_depth_kernel<PB><<<(p.probeCount + PB - 1) / PB, 512, smem, s>>>(p);
}

void launch_blend_depth(const BlendParams& p, cudaStream_t s)
{
    // 32-probe tiles (two blocks per SM, 90 KB each) beat 64-probe tiles at every size measured: C4 full volume 0.83 -> 0.72 ms, 1/8 shard
    // 0.123 -> 0.10 ms; the barrier between the staged ray chunks is hidden by the second block.
    if (p.probeCount >= 64 * 64)
        launch_blend_depth_t<32>(p, s);
    else
        launch_blend_depth_t<16>(p, s);
}

// ---- list blend (blend_lists.inc) ----
static size_t align16(size_t v) { return (v + 15) & ~size_t(15); }
BlendLists blend_lists_layout(void* base, int raysPerProbe, int raysPadded)
{
    BlendLists L{};
    L.cap         = raysPadded + LB_PAD;
    L.irrPhases   = (raysPerProbe + LBI_KP - 1) / LBI_KP;
    L.depthPhases = (raysPerProbe + LBD_KP - 1) / LBD_KP;
    unsigned char* p = static_cast<unsigned char*>(base);
    L.irrW     = reinterpret_cast<float4*>(p);   p += align16((size_t)LBI_GROUPS * L.cap * sizeof(float4));
    L.depthW   = reinterpret_cast<float4*>(p);   p += align16((size_t)LBD_GROUPS * L.cap * sizeof(float4));
    L.irrIdx   = reinterpret_cast<uint16_t*>(p); p += align16((size_t)LBI_GROUPS * L.cap * sizeof(uint16_t));
    L.depthIdx = reinterpret_cast<uint16_t*>(p); p += align16((size_t)LBD_GROUPS * L.cap * sizeof(uint16_t));
    L.irrOff   = reinterpret_cast<uint32_t*>(p); p += align16((size_t)LBI_GROUPS * (L.irrPhases + 1) * sizeof(uint32_t));
    L.depthOff = reinterpret_cast<uint32_t*>(p); p += align16((size_t)LBD_GROUPS * (L.depthPhases + 1) * sizeof(uint32_t));
    L.irrMean     = reinterpret_cast<float*>(p); p += align16(LBI_GROUPS * sizeof(float));
    L.depthMean   = reinterpret_cast<float*>(p); p += align16(LBD_GROUPS * sizeof(float));
    L.irrAssign   = reinterpret_cast<uint8_t*>(p); p += align16(LBI_GROUPS);
    L.depthAssign = reinterpret_cast<uint8_t*>(p);
    return L;
}
size_t blend_lists_bytes(int raysPerProbe, int raysPadded)
{
    const BlendLists L = blend_lists_layout(nullptr, raysPerProbe, raysPadded);
    return (size_t)(reinterpret_cast<unsigned char*>(L.depthAssign) - static_cast<unsigned char*>(nullptr)) + align16(LBD_GROUPS);
}
int launch_blend_lists(const float* wIrr, const float* wDepth, int raysPerProbe, const BlendLists& lists, cudaStream_t s)
{
    blend_lists_kernel<<<LBI_GROUPS + LBD_GROUPS, 32, 0, s>>>(wIrr, wDepth, raysPerProbe, lists);
    blend_assign_kernel<<<1, 64, 0, s>>>(lists);
    return 2;
}
bool blend_lists_preferred(int probeCount) { return probeCount >= 148 * LB_PB; }
void launch_blend_irradiance_lists(const BlendParams& p, const BlendLists& lists, cudaStream_t s)
{
    const size_t smem = (size_t)LBI_KP * 3 * LB_PB * sizeof(float) + (size_t)LB_PB * LBI_RAW_LD * sizeof(uint2) +
                        (LBI_THREADS / 32) * (2 * sizeof(LbStage) + 2 * sizeof(uint64_t) + 6 * sizeof(int));
    cudaFuncSetAttribute(blend_irradiance_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); // per launch: the attribute is per device
    blend_irradiance_lists_kernel<<<(p.probeCount + LB_PB - 1) / LB_PB, LBI_THREADS, smem, s>>>(p, lists);
}
void launch_blend_depth_lists(const BlendParams& p, const BlendLists& lists, cudaStream_t s)
{
    const size_t smem = (size_t)LBD_KP * LB_PB * sizeof(float) + (size_t)LB_PB * LBD_RAW_LD * sizeof(uint32_t) +
                        (LBD_THREADS / 32) * (2 * sizeof(LbStage) + 2 * sizeof(uint64_t) + 12 * sizeof(int));
    cudaFuncSetAttribute(blend_depth_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    blend_depth_lists_kernel<<<(p.probeCount + LB_PB - 1) / LB_PB, LBD_THREADS, smem, s>>>(p, lists);
}

// ---- tensor-core blend (LUX_DDGI_FLAG_BLEND_TC, blend_tc.inc) ----
int blend_tc_kpad(int raysPerProbe) { return (raysPerProbe + TC_KC - 1) / TC_KC * TC_KC; }

void launch_blend_tc_weights(const float* wIrr, const float* wDepth, int rowsPadded, int kPad, uint16_t* irrHi, uint16_t* irrLo, uint16_t* depthHi,
                             uint16_t* depthLo, cudaStream_t s)
{
    blend_tc_weights_kernel<<<(64 * kPad + 255) / 256, 256, 0, s>>>(wIrr, rowsPadded, kPad, 64, irrHi, irrLo);
    blend_tc_weights_kernel<<<(256 * kPad + 255) / 256, 256, 0, s>>>(wDepth, rowsPadded, kPad, 256, depthHi, depthLo);
}

void launch_blend_irradiance_tc(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s)
{
    const size_t smem = 192 * 64 * sizeof(float); // the epilogue overlay is the larger of the two uses
    cudaFuncSetAttribute(blend_irradiance_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    blend_irradiance_tc_kernel<<<(p.probeCount + 63) / 64, 256, smem, s>>>(p, hi, lo, kPad);
}

bool launch_blend_depth_tc(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s)
{
    if (!(p.maxDistance * p.maxDistance < 60000.0f)) // d * d must stay inside fp16
        return false;
    const size_t smem = (size_t)(2 * 64 + 2 * 256) * TC_LD * sizeof(uint16_t);
    cudaFuncSetAttribute(blend_depth_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    blend_depth_tc_kernel<<<(p.probeCount + 31) / 32, 256, smem, s>>>(p, hi, lo, kPad);
    return true;
}

// ---- tcgen05 / TMA form of the tensor-core blend (blend_umma.inc) ----
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeTiledFn tensor_map_encoder()
{
    static TensorMapEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }();
    return fn;
}
// fp16 matrix [outer][inner] with a row pitch in bytes, boxes of boxOuter x boxInner elements, 128-byte swizzle, zero fill outside
static bool make_tensor_map_f16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t pitchBytes, uint32_t boxInner, uint32_t boxOuter)
{
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15) || (pitchBytes & 15))
        return false;
    const cuuint64_t dims[2] = {inner, outer}, strides[1] = {pitchBytes};
    const cuuint32_t box[2] = {boxInner, boxOuter}, estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
int blend_umma_irr_kpad(int raysPerProbe) { return (raysPerProbe * 4 + umma::BK - 1) / umma::BK * umma::BK; }
void launch_blend_umma_irr_weights(const float* wIrr, int raysPerProbe, uint16_t* hi, uint16_t* lo, cudaStream_t s)
{
    const int kPad = blend_umma_irr_kpad(raysPerProbe);
    umma::blend_umma_irr_weights_kernel<<<(umma::IRR_BN * kPad + 255) / 256, 256, 0, s>>>(wIrr, raysPerProbe, kPad, hi, lo);
}
bool launch_blend_irradiance_umma(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, cudaStream_t s)
{
    const int kPad = blend_umma_irr_kpad(p.raysPerProbe);
    CUtensorMap mapA, mapBhi, mapBlo;
    if (!make_tensor_map_f16(&mapA, p.radiance, (uint64_t)p.raysPerProbe * 4, (uint64_t)p.probeCount, (uint64_t)p.raysPerProbe * 8, umma::BK, umma::BM) ||
        !make_tensor_map_f16(&mapBhi, hi, (uint64_t)kPad, umma::IRR_BN, (uint64_t)kPad * 2, umma::BK, umma::IRR_BN) ||
        !make_tensor_map_f16(&mapBlo, lo, (uint64_t)kPad, umma::IRR_BN, (uint64_t)kPad * 2, umma::BK, umma::IRR_BN))
        return false; // odd ray count (row pitch not a multiple of 16 bytes) or no encoder: the mma.sync kernel takes over
    const size_t smem = (size_t)umma::STAGES * umma::IRR_STAGE_BYTES + 1024 + 256;
    cudaFuncSetAttribute(umma::blend_irradiance_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma::blend_irradiance_umma_kernel<<<(p.probeCount + umma::BM - 1) / umma::BM, 256, smem, s>>>(mapA, mapBhi, mapBlo, p, kPad / umma::BK);
    return true;
}

bool launch_blend_depth_umma(const BlendParams& p, const uint16_t* hi, const uint16_t* lo, int kPad, cudaStream_t s)
{
    if (!(p.maxDistance * p.maxDistance < 60000.0f) || (p.raysPerProbe & 1)) // d * d must stay inside fp16; rays are loaded in pairs
        return false;
    CUtensorMap mapBhi, mapBlo;
    if (!make_tensor_map_f16(&mapBhi, hi, (uint64_t)kPad, umma::DEP_BN, (uint64_t)kPad * 2, umma::BK, umma::DEP_BN) ||
        !make_tensor_map_f16(&mapBlo, lo, (uint64_t)kPad, umma::DEP_BN, (uint64_t)kPad * 2, umma::BK, umma::DEP_BN))
        return false;
    const size_t smem = (size_t)umma::DEP_STAGES * umma::DEP_STAGE_BYTES + 1024 + 256;
    cudaFuncSetAttribute(umma::blend_depth_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma::blend_depth_umma_kernel<<<(p.probeCount + umma::DEP_PROBES - 1) / umma::DEP_PROBES, 320, smem, s>>>(mapBhi, mapBlo, p, kPad / umma::BK);
    return true;
}

void launch_border(uint2* irr, int irrWidth, uint32_t* depth, int depthWidth, int probesPerRow, int probeBegin, int probeCount, int layerProbes, int layerStride,
                   cudaStream_t s)
{
    if (probeCount <= 0)
        return;
    if (irr)
    {
        long long n = (long long)probeCount * 36;
        border_kernel<uint2, 8><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(irr, irrWidth, probesPerRow, probeBegin, probeCount, layerProbes, layerStride);
    }
    if (depth)
    {
        long long n = (long long)probeCount * 68;
        border_kernel<uint32_t, 16><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(depth, depthWidth, probesPerRow, probeBegin, probeCount, layerProbes, layerStride);
    }
}



} // namespace lux
