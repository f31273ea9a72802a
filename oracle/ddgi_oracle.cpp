// ddgi_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's SDF-traced DDGI probe update (flwmxd/LuxGI, Maple engine) used as the
// parity oracle and as the timed CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; libluxddgi.so never does.
//
// PARITY PINNING: the reference ships no tests, golden vectors or KATs for this path and no
// Vulkan toolchain exists in this image, but it does ship its compiled shaders.  What pins this restatement:
//   * the reference's own border offset tables (BorderUpdate.glsl:25-133), extracted to
//     tests/golden/border_offsets.json and compared entry by entry;
//   * golden vectors produced by EXECUTING the reference's shipped SPIR-V binaries (Assets/shaders/spv/DDGI/*.comp.spv)
//     with the interpreter in oracle/spirv/ (tests/golden/README.md): trace and border reproduce them bit for bit, blend
//     bit for bit in `unfused` mode and within 1 fp16 ulp in the contract's FMA mode (tests/test_spirv_golden.py); the trace also on
//     2- and 4-cascade volumes and on an open scene (sky path);
//   * the rows either side of the path the same way, from Assets/shaders/spv/{DDGI/SampleProbe.comp, SDF/SDFRasterizeModel*.comp,
//     SDF/GlobalSDFMipmap.comp, SDF/SDFCulling.comp, SDF/SDFDeferredLight.frag, SDF/SDFAtlasIndirectLight.frag, SDF/SDFReflection.comp,
//     SDF/SDFShadow.comp}.spv, all bit for bit;
//   * closed-form known-answer tests (tests/test_oracle_kat.py).
//
// Every function cites the reference file:line it follows (paths relative to Code/Maple/src/).
//
// NUMERICS CONTRACT (shared with the CUDA engine by specification, not by code — see DESIGN.md §4):
//   * every +,-,*,/,sqrt is IEEE-754 binary32 round-to-nearest-even, evaluated in source order; no contraction
//     (built with -ffp-contract=off) except where fmaf() is written explicitly (blend accumulation, mix);
//   * GLSL min/max/clamp are select-based: min(x,y) = y<x ? y : x, max(x,y) = x<y ? y : x;
//   * sin, cos, pow are evaluated in binary64 and rounded once to binary32;
//   * normalize(v) = v * (1 / sqrt((v.x*v.x + v.y*v.y) + v.z*v.z));
//   * textures: fp16 texels decoded exactly, fp32 filter weights, nested lerp x→y→z, lerp(a,b,t) = a + t*(b-a);
//   * imageStore to *16F formats rounds to nearest even (overflow → inf).
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp -shared).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/luxddgi.h" // POD layouts of the boundary only (no product code)

namespace {

// ---------------------------------------------------------------------------------------------------------
// binary16 <-> binary32 (Vulkan image load/store conversion for R16F/RG16F/RGBA16F, VulkanHelper.cpp:1113-1130)
// ---------------------------------------------------------------------------------------------------------
inline float h2f(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1fu;
    uint32_t man  = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0)
    {
        if (man == 0)
            bits = sign;
        else
        { // subnormal: value = man * 2^-24
            float    f = (float)man * 5.9604644775390625e-08f;
            uint32_t b;
            std::memcpy(&b, &f, 4);
            bits = b | sign;
        }
    }
    else if (exp == 31)
        bits = sign | 0x7f800000u | (man << 13);
    else
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    float out;
    std::memcpy(&out, &bits, 4);
    return out;
}

inline uint16_t f2h(float f)
{
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    x &= 0x7fffffffu;
    if (x > 0x7f800000u)
        return 0x7fffu; // NaN (canonical, as cvt.rn.f16.f32)
    if (x >= 0x477ff000u)
        return sign | 0x7c00u; // >= 65520 rounds to inf
    if (x < 0x38800000u)
    { // below 2^-14: half subnormal, integer mantissa = RNE(|f| * 2^24)
        float a;
        std::memcpy(&a, &x, 4);
        uint32_t m = (uint32_t)std::lrintf(a * 16777216.0f); // exact scaling, RNE under the default mode
        return sign | (uint16_t)m;                           // m == 1024 is the smallest normal, same bits
    }
    uint32_t man = x & 0x7fffffu;
    uint32_t e   = (x >> 23) - 112u;
    uint32_t h   = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u)))
        h++;
    return sign | (uint16_t)h;
}

// ---------------------------------------------------------------------------------------------------------
// GLSL built-ins under the numerics contract
// ---------------------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline int   iclamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
inline float gfract(float x) { return x - std::floor(x); }

inline vec3 add(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 sub(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 mul(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 mul(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot3(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot4(vec4 a, vec4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
inline vec3 normalize3(vec3 v)
{
    float inv = 1.0f / std::sqrt(dot3(v, v));
    return {v.x * inv, v.y * inv, v.z * inv};
}
inline float length3(vec3 v) { return std::sqrt(dot3(v, v)); }
inline float sin_rn(float x) { return (float)std::sin((double)x); }
inline float cos_rn(float x) { return (float)std::cos((double)x); }
inline float pow_rn(float x, float y) { return (float)std::pow((double)x, (double)y); }

// column-major mat4 m[c*4+r] times (v,1) / times (v, w)
inline vec3 mat4_mul_point(const float* m, vec3 v, float w)
{
    vec3 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w;
    return r;
}
inline vec3 mat3_mul(const float* m, vec3 v) // upper-left 3x3 of a column-major mat4
{
    vec3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + m[8] * v.z;
    r.y = (m[1] * v.x + m[5] * v.y) + m[9] * v.z;
    r.z = (m[2] * v.x + m[6] * v.y) + m[10] * v.z;
    return r;
}

// inverse(mat4) — GLSL leaves the algorithm to the implementation; the contract fixes it to the cofactor
// expansion over 2x2 sub-determinants below (used at AtlasCommon.glsl:133).
void inverse4(const float* m, float* o)
{
#define A(r, c) m[(c)*4 + (r)]
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
    float inv = 1.0f / det;
#define B(r, c) o[(c)*4 + (r)]
    B(0, 0) = ((A(1, 1) * c5 - A(1, 2) * c4) + A(1, 3) * c3) * inv;
    B(0, 1) = ((-A(0, 1) * c5 + A(0, 2) * c4) - A(0, 3) * c3) * inv;
    B(0, 2) = ((A(3, 1) * s5 - A(3, 2) * s4) + A(3, 3) * s3) * inv;
    B(0, 3) = ((-A(2, 1) * s5 + A(2, 2) * s4) - A(2, 3) * s3) * inv;
    B(1, 0) = ((-A(1, 0) * c5 + A(1, 2) * c2) - A(1, 3) * c1) * inv;
    B(1, 1) = ((A(0, 0) * c5 - A(0, 2) * c2) + A(0, 3) * c1) * inv;
    B(1, 2) = ((-A(3, 0) * s5 + A(3, 2) * s2) - A(3, 3) * s1) * inv;
    B(1, 3) = ((A(2, 0) * s5 - A(2, 2) * s2) + A(2, 3) * s1) * inv;
    B(2, 0) = ((A(1, 0) * c4 - A(1, 1) * c2) + A(1, 3) * c0) * inv;
    B(2, 1) = ((-A(0, 0) * c4 + A(0, 1) * c2) - A(0, 3) * c0) * inv;
    B(2, 2) = ((A(3, 0) * s4 - A(3, 1) * s2) + A(3, 3) * s0) * inv;
    B(2, 3) = ((-A(2, 0) * s4 + A(2, 1) * s2) - A(2, 3) * s0) * inv;
    B(3, 0) = ((-A(1, 0) * c3 + A(1, 1) * c1) - A(1, 2) * c0) * inv;
    B(3, 1) = ((A(0, 0) * c3 - A(0, 1) * c1) + A(0, 2) * c0) * inv;
    B(3, 2) = ((-A(3, 0) * s3 + A(3, 1) * s1) - A(3, 2) * s0) * inv;
    B(3, 3) = ((A(2, 0) * s3 - A(2, 1) * s1) + A(2, 2) * s0) * inv;
#undef A
#undef B
}

// ---------------------------------------------------------------------------------------------------------
// Software texture units
// ---------------------------------------------------------------------------------------------------------
struct Tex3D // R16F, [z][y][x], linear filter, clamp-to-edge (GlobalDistanceField.cpp:615,621)
{
    const uint16_t* data;
    int             w, h, d;
    inline float texel(int x, int y, int z) const { return h2f(data[((size_t)z * h + y) * w + x]); }
};

inline float lerp1(float a, float b, float t) { return a + t * (b - a); }

// texture(sampler3D, uvw).r at LOD 0
float sample3D(const Tex3D& t, float u, float v, float w, uint64_t* taps)
{
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f, z = w * (float)t.d - 0.5f;
    float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    float ax = x - fx, ay = y - fy, az = z - fz;
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;
    int   x0 = iclamp(ix, 0, t.w - 1), x1 = iclamp(ix + 1, 0, t.w - 1);
    int   y0 = iclamp(iy, 0, t.h - 1), y1 = iclamp(iy + 1, 0, t.h - 1);
    int   z0 = iclamp(iz, 0, t.d - 1), z1 = iclamp(iz + 1, 0, t.d - 1);
    float c00 = lerp1(t.texel(x0, y0, z0), t.texel(x1, y0, z0), ax);
    float c10 = lerp1(t.texel(x0, y1, z0), t.texel(x1, y1, z0), ax);
    float c01 = lerp1(t.texel(x0, y0, z1), t.texel(x1, y0, z1), ax);
    float c11 = lerp1(t.texel(x0, y1, z1), t.texel(x1, y1, z1), ax);
    float c0  = lerp1(c00, c10, ay);
    float c1  = lerp1(c01, c11, ay);
    if (taps)
        ++*taps;
    return lerp1(c0, c1, az);
}

// textureGather footprint: i0 = floor(u*W - 0.5), order (i0,j1),(i1,j1),(i1,j0),(i0,j0) (Vulkan spec §16.9)
struct GatherCoords { int i0, i1, j0, j1; };
inline GatherCoords gatherCoords(float u, float v, int W, int H, bool repeat)
{
    int i0 = (int)std::floor(u * (float)W - 0.5f);
    int j0 = (int)std::floor(v * (float)H - 0.5f);
    int i1 = i0 + 1, j1 = j0 + 1;
    GatherCoords g;
    if (repeat)
    {
        g.i0 = ((i0 % W) + W) % W; g.i1 = ((i1 % W) + W) % W;
        g.j0 = ((j0 % H) + H) % H; g.j1 = ((j1 % H) + H) % H;
    }
    else
    {
        g.i0 = iclamp(i0, 0, W - 1); g.i1 = iclamp(i1, 0, W - 1);
        g.j0 = iclamp(j0, 0, H - 1); g.j1 = iclamp(j1, 0, H - 1);
    }
    return g;
}

// Conservative "open space" table over the mip volume, the CPU twin of the engine's optional mip-tap skip (LUX_DDGI_FLAG_OPEN_SKIP): one bit per
// cell of 8x8x8 mip texels, set when every texel a trilinear tap placed anywhere in the cell can touch (the cell dilated by one texel, clamped
// at the volume's edges, cascade seams included because cells tile the whole side-by-side volume) is >= threshold.  A set bit proves that
// tracyGlobalSDF's mip tap there returns >= chunkSizeDistance, i.e. the march takes its `stepDistance = chunkSizeDistance` branch without
// needing the tap.  It never changes a result; the oracle only uses it to CHECK that claim (Counters::openViolations must stay 0).
struct OpenTable
{
    int                   cw = 0, ch = 0, cd = 0; // cells per axis of the whole mip volume
    int                   cell = 8;               // mip texels per cell side (the engine uses 8; other sizes only for the share statistics)
    std::vector<uint32_t> bits;
    std::vector<uint32_t> nearBits; // statistics only: every texel of the dilated cell < chunkSizeDistance * (1 - 2^-10)
    bool open(int cx, int cy, int cz) const
    {
        size_t i = ((size_t)cz * ch + cy) * cw + cx;
        return (bits[i >> 5] >> (i & 31)) & 1u;
    }
    bool isNear(int cx, int cy, int cz) const
    {
        size_t i = ((size_t)cz * ch + cy) * cw + cx;
        return (nearBits[i >> 5] >> (i & 31)) & 1u;
    }
};

struct Scene
{
    const OpenTable*           open = nullptr;
    bool                       openEmulate = false; // decide the mip test FROM the table, statement for statement as march_open_kernel does
    LuxDDGIUniform             ddgi;
    LuxGlobalSDFData           sdfData;
    Tex3D                      tex, mip;
    bool                       hasAtlas;
    LuxGlobalSurfaceAtlasData  atlasData;
    const uint32_t*            chunks;
    const uint32_t*            cull;
    const LuxObjectBuffer*     objects;
    const LuxTileBuffer*       tiles;
    const uint16_t*            light; // RGBA16F res^2, linear/repeat (GlobalSurfaceAtlas.cpp:411, Definitions.h:152-159)
    const float*               depth; // D32F res^2, linear/clamp (VulkanTexture.cpp:643-668)
    int                        skyFace;
    const uint16_t*            sky; // 6 faces RGBA16F, or null = black 1x1 fallback (DDGIRenderer.cpp:308)
};

struct Counters
{
    uint64_t mipTaps = 0, texTaps = 0, hits = 0, tileSamples = 0, steps = 0, objectsVisited = 0;
    uint64_t openSteps = 0, openViolations = 0; // only with Scene::open (validation of the engine's open-space table, see buildOpenTable)
    uint64_t nearSteps = 0, nearViolations = 0, nearTexUsed = 0, texUsed = 0; // statistics for a possible "near" table (mip tap provably < chunkSizeDistance)
    uint8_t* classOut = nullptr; // optional per-step class trace of one ray: 1 open, 2 near + full-resolution tap used, 3 near + both taps, 4 undecided (both)
    int      classCap = 0, classN = 0;
};

// ---------------------------------------------------------------------------------------------------------
// DDGICommon.glsl
// ---------------------------------------------------------------------------------------------------------
// DDGICommon.glsl:42-52.  Constants as folded by glslang in the shipped GISDFRays.comp.spv.
vec3 sphericalFibonacci(float i, float raysPerProbe)
{
    const float PHI_M1 = 0.61803400516510009765625f;
    const float TWO_PI = 6.283185482025146484375f;
    float       ab       = i * PHI_M1;
    float       phi      = TWO_PI * (ab - std::floor(ab));
    float       cosTheta = 1.0f - (2.0f * i + 1.0f) * (1.0f / raysPerProbe);
    float       sinTheta = std::sqrt(gclamp(1.0f - cosTheta * cosTheta, 0.0f, 1.0f));
    return {cos_rn(phi) * sinTheta, sin_rn(phi) * sinTheta, cosTheta};
}

inline float signNotZero(float k) { return (k >= 0.0f) ? 1.0f : -1.0f; } // DDGICommon.glsl:55-58

// DDGICommon.glsl:74-82
vec3 octDecode(vec2 o)
{
    vec3 v = {o.x, o.y, (1.0f - std::fabs(o.x)) - std::fabs(o.y)};
    if (v.z < 0.0f)
    {
        float nx = (1.0f - std::fabs(v.y)) * signNotZero(v.x);
        float ny = (1.0f - std::fabs(v.x)) * signNotZero(v.y);
        v.x = nx;
        v.y = ny;
    }
    return normalize3(v);
}

// DDGICommon.glsl:85-92 with fragCoord = probe base + 2 + (i,j): (fragCoord-2) % (side+2) = (i,j)
vec2 normalizedOctCoordLocal(int i, int j, int side)
{
    float s = 2.0f / (float)side;
    return {((float)i + 0.5f) * s - 1.0f, ((float)j + 0.5f) * s - 1.0f};
}

// DDGICommon.glsl:101-114
vec3 probeLocation(const LuxDDGIUniform& d, int index)
{
    int X = d.probeCounts[0], Y = d.probeCounts[1];
    int cx = index % X;
    int cy = (index % (X * Y)) / X;
    int cz = index / (X * Y);
    return {d.step[0] * (float)cx + d.startPosition[0], d.step[1] * (float)cy + d.startPosition[1],
            d.step[2] * (float)cz + d.startPosition[2]};
}

// ---------------------------------------------------------------------------------------------------------
// SDFCommon.glsl
// ---------------------------------------------------------------------------------------------------------
struct Hit // GlobalSDFHit.glsl:4-11
{
    vec3     hitNormal;
    float    hitTime;
    uint32_t hitCascade;
    uint32_t stepsCount;
    float    hitSDF;
};

// SDFCommon.glsl:84-95
vec2 lineHitAABB(vec3 s, vec3 e, vec3 bmin, vec3 bmax)
{
    vec3 inv   = {1.0f / (e.x - s.x), 1.0f / (e.y - s.y), 1.0f / (e.z - s.z)};
    vec3 enter = mul(sub(bmin, s), inv);
    vec3 exit_ = mul(sub(bmax, s), inv);
    vec3 mn    = {gmin(enter.x, exit_.x), gmin(enter.y, exit_.y), gmin(enter.z, exit_.z)};
    vec3 mx    = {gmax(enter.x, exit_.x), gmax(enter.y, exit_.y), gmax(enter.z, exit_.z)};
    vec2 r;
    r.x = gmax(mn.x, gmax(mn.y, mn.z));
    r.y = gmin(mx.x, gmin(mx.y, mx.z));
    r.x = gclamp(r.x, 0.0f, 1.0f);
    r.y = gclamp(r.y, 0.0f, 1.0f);
    return r;
}

// SDFCommon.glsl:98-194, with trace = {origin, dir, minDistance 0, maxDistance, stepScale, needsHitNormal true}
Hit tracyGlobalSDF(const Scene& sc, vec3 origin, vec3 dir, float maxDistance, float stepScale,
                   float cascadeTraceStartBias, Counters& cn)
{
    const LuxGlobalSDFData& data = sc.sdfData;
    Hit                     hit;
    hit.stepsCount = 0;
    hit.hitTime    = -1.0f;
    hit.hitNormal  = {0, 0, 0};
    hit.hitCascade = 0;
    hit.hitSDF     = 0.0f;

    float traceMaxDistance    = gmin(maxDistance, data.cascadePosDistance[data.cascadesCount - 1][3] * 2.0f);
    float chunkSizeDistance   = (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE / data.resolution;
    float chunkMarginDistance = (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_MARGIN / data.resolution;
    float nextIntersectionStart = 0.0f;
    vec3  traceEndPosition      = add(origin, mul(dir, traceMaxDistance));
    float cascadesCountF        = (float)data.cascadesCount;

    for (uint32_t cascade = 0; cascade < data.cascadesCount && hit.hitTime < 0.0f; cascade++)
    {
        const float* cpd       = data.cascadePosDistance[cascade];
        float        voxelSize = data.cascadeVoxelSize[cascade];
        float        voxelHalf = voxelSize * 0.5f;
        vec3         worldPosition = add(origin, mul(dir, voxelSize * cascadeTraceStartBias));

        vec3 c    = {cpd[0], cpd[1], cpd[2]};
        vec3 ext  = {cpd[3], cpd[3], cpd[3]};
        vec2 isec = lineHitAABB(worldPosition, traceEndPosition, sub(c, ext), add(c, ext));
        isec.x *= traceMaxDistance;
        isec.y *= traceMaxDistance;
        isec.x = gmax(isec.x, nextIntersectionStart);

        float stepTime = isec.x;
        if (isec.x >= isec.y)
            stepTime = isec.y;
        else
            nextIntersectionStart = isec.y;

        uint32_t step = 0;
        for (; step < LUX_GLOBAL_SDF_MAX_STEPS && stepTime < isec.y; step++)
        {
            vec3 stepPosition = add(worldPosition, mul(dir, stepTime));

            // getGlobalSDFCascadeUV, SDFCommon.glsl:64-71
            vec3  posInCascade       = sub(stepPosition, c);
            float cascadeMaxDistance = cpd[3] * 2.0f;
            vec3  cascadeUV = {gclamp(posInCascade.x / cascadeMaxDistance + 0.5f, 0.0f, 1.0f),
                               gclamp(posInCascade.y / cascadeMaxDistance + 0.5f, 0.0f, 1.0f),
                               gclamp(posInCascade.z / cascadeMaxDistance + 0.5f, 0.0f, 1.0f)};
            vec3  textureUV = {((float)cascade + cascadeUV.x) / cascadesCountF, cascadeUV.y, cascadeUV.z};

            if (sc.open && sc.openEmulate)
            { // the experimental march's control flow (march_kernel.inc, MARCH_OPEN_SKIP): same table, same decisions, same taps
                const OpenTable& ot = *sc.open;
                const float inv = 1.0f / (float)ot.cell;
                int cx = std::min((int)((textureUV.x * (float)sc.mip.w) * inv), ot.cw - 1);
                int cy = std::min((int)((textureUV.y * (float)sc.mip.h) * inv), ot.ch - 1);
                int cz = std::min((int)((textureUV.z * (float)sc.mip.d) * inv), ot.cd - 1);
                float d;
                if (ot.open(cx, cy, cz))
                    d = chunkSizeDistance;
                else if (ot.isNear(cx, cy, cz))
                {
                    float t = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z, &cn.texTaps);
                    d = t;
                    if (!(t < chunkMarginDistance * 2.0f))
                        d = sample3D(sc.mip, textureUV.x, textureUV.y, textureUV.z, &cn.mipTaps);
                }
                else
                {
                    d = sample3D(sc.mip, textureUV.x, textureUV.y, textureUV.z, &cn.mipTaps);
                    if (d < chunkSizeDistance)
                    {
                        float t = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z, &cn.texTaps);
                        if (t < chunkMarginDistance * 2.0f)
                            d = t;
                    }
                    else
                        d = chunkSizeDistance;
                }
                d *= cascadeMaxDistance;
                float thick = voxelHalf * gclamp(stepTime / voxelSize, 0.0f, 1.0f);
                if (d < thick)
                {
                    hit.hitTime    = gmax((stepTime + d) - thick, 0.0f);
                    hit.hitCascade = cascade;
                    hit.hitSDF     = d;
                    float o = 1.0f / data.resolution;
                    float xp = sample3D(sc.tex, textureUV.x + o, textureUV.y, textureUV.z, &cn.texTaps), xn = sample3D(sc.tex, textureUV.x - o, textureUV.y, textureUV.z, &cn.texTaps);
                    float yp = sample3D(sc.tex, textureUV.x, textureUV.y + o, textureUV.z, &cn.texTaps), yn = sample3D(sc.tex, textureUV.x, textureUV.y - o, textureUV.z, &cn.texTaps);
                    float zp = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z + o, &cn.texTaps), zn = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z - o, &cn.texTaps);
                    hit.hitNormal = normalize3({xp - xn, yp - yn, zp - zn});
                    break;
                }
                stepTime += gmax(d * stepScale, voxelSize);
                continue;
            }
            float stepDistance = sample3D(sc.mip, textureUV.x, textureUV.y, textureUV.z, &cn.mipTaps);
            if (sc.open)
            { // the cell the engine would look up: floor(fl(u * W) / 4) per axis, clamped
                const OpenTable& ot = *sc.open;
                const float inv = 1.0f / (float)ot.cell; // a power of two: exact
                int cx = iclamp((int)((textureUV.x * (float)sc.mip.w) * inv), 0, ot.cw - 1);
                int cy = iclamp((int)((textureUV.y * (float)sc.mip.h) * inv), 0, ot.ch - 1);
                int cz = iclamp((int)((textureUV.z * (float)sc.mip.d) * inv), 0, ot.cd - 1);
                if (ot.open(cx, cy, cz))
                {
                    cn.openSteps++;
                    if (stepDistance < chunkSizeDistance)
                        cn.openViolations++;
                }
                uint8_t cls = ot.open(cx, cy, cz) ? 1 : 4;
                if (ot.isNear(cx, cy, cz))
                {
                    cn.nearSteps++;
                    cls = 3;
                    if (!(stepDistance < chunkSizeDistance))
                        cn.nearViolations++;
                    else if (sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z, nullptr) < chunkMarginDistance * 2.0f)
                    {
                        cn.nearTexUsed++;
                        cls = 2;
                    }
                }
                if (cn.classOut && cn.classN < cn.classCap)
                    cn.classOut[cn.classN++] = cls;
            }
            if (stepDistance < chunkSizeDistance)
            {
                float stepDistanceTex = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z, &cn.texTaps);
                if (stepDistanceTex < chunkMarginDistance * 2.0f)
                {
                    stepDistance = stepDistanceTex;
                    cn.texUsed++;
                }
            }
            else
                stepDistance = chunkSizeDistance;

            stepDistance *= cascadeMaxDistance;

            float minSurfaceThickness = voxelHalf * gclamp(stepTime / voxelSize, 0.0f, 1.0f);
            if (stepDistance < minSurfaceThickness)
            {
                hit.hitTime    = gmax((stepTime + stepDistance) - minSurfaceThickness, 0.0f);
                hit.hitCascade = cascade;
                hit.hitSDF     = stepDistance;
                float texelOffset = 1.0f / data.resolution;
                float xp = sample3D(sc.tex, textureUV.x + texelOffset, textureUV.y, textureUV.z, &cn.texTaps);
                float xn = sample3D(sc.tex, textureUV.x - texelOffset, textureUV.y, textureUV.z, &cn.texTaps);
                float yp = sample3D(sc.tex, textureUV.x, textureUV.y + texelOffset, textureUV.z, &cn.texTaps);
                float yn = sample3D(sc.tex, textureUV.x, textureUV.y - texelOffset, textureUV.z, &cn.texTaps);
                float zp = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z + texelOffset, &cn.texTaps);
                float zn = sample3D(sc.tex, textureUV.x, textureUV.y, textureUV.z - texelOffset, &cn.texTaps);
                hit.hitNormal = normalize3({xp - xn, yp - yn, zp - zn});
                break;
            }
            stepTime += gmax(stepDistance * stepScale, voxelSize);
        }
        hit.stepsCount += step;
    }
    cn.steps += hit.stepsCount;
    return hit;
}

// ---------------------------------------------------------------------------------------------------------
// AtlasCommon.glsl
// ---------------------------------------------------------------------------------------------------------
// AtlasCommon.glsl:58-97 (tile sampling) + :100-112 (normal weight)
vec4 sampleGlobalSurfaceAtlasTile(const Scene& sc, const LuxTileBuffer& tile, vec3 localPosition, vec3 normal,
                                  float surfaceThreshold, Counters& cn)
{
    // :104-110
    vec3 nt = mat4_mul_point(tile.transform, normal, 1.0f);
    nt      = normalize3(nt);
    float normalWeight = gclamp(nt.z, 0.0f, 1.0f);
    normalWeight = (normalWeight - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD) / (1.0f - LUX_SURFACE_ATLAS_TILE_NORMAL_THRESHOLD);
    if (normalWeight <= 0.0f)
        return {0, 0, 0, 0};

    cn.tileSamples++;
    // :62-67
    vec3  tp        = mat4_mul_point(tile.transform, localPosition, 1.0f);
    float tileDepth = tp.z / tile.objectBounds[2];
    vec2  tileUV    = {gclamp(tp.x / tile.objectBounds[0] + 0.5f, 0.0f, 1.0f),
                       gclamp(tp.y / tile.objectBounds[1] + 0.5f, 0.0f, 1.0f)};
    vec2  atlasUV   = {tileUV.x * tile.extends[2] + tile.extends[0], tileUV.y * tile.extends[3] + tile.extends[1]};
    // :68-74
    float res = (float)sc.atlasData.resolution;
    vec2  f   = {gfract(atlasUV.x * res + 0.5f), gfract(atlasUV.y * res + 0.5f)};
    vec4  bw  = {(1.0f - f.x) * f.y, f.x * f.y, f.x * (1.0f - f.y), (1.0f - f.x) * (1.0f - f.y)};
    // :76-84
    int          R  = (int)sc.atlasData.resolution;
    GatherCoords gd = gatherCoords(atlasUV.x, atlasUV.y, R, R, false);
    float        z4[4] = {sc.depth[(size_t)gd.j1 * R + gd.i0], sc.depth[(size_t)gd.j1 * R + gd.i1],
                          sc.depth[(size_t)gd.j0 * R + gd.i1], sc.depth[(size_t)gd.j0 * R + gd.i0]};
    float depthThreshold = 2.0f * surfaceThreshold / tile.objectBounds[2];
    float vis[4];
    for (int i = 0; i < 4; i++)
    {
        vis[i] = 1.0f - gclamp((std::fabs(tileDepth - z4[i]) - depthThreshold) / (0.5f * depthThreshold), 0.0f, 1.0f);
        if (z4[i] >= 1.0f)
            vis[i] = 0.0f;
    }
    vec4 visv = {vis[0], vis[1], vis[2], vis[3]};
    // :86-91
    float sampleWeight = dot4(visv, bw);
    sampleWeight *= normalWeight;
    if (sampleWeight <= 0.0f)
        return {0, 0, 0, 0};
    // :93-96 (sampleGlobalSurfaceAtlasTex :50-56)
    bw = {bw.x * visv.x, bw.y * visv.y, bw.z * visv.z, bw.w * visv.w};
    GatherCoords gc = gatherCoords(atlasUV.x, atlasUV.y, R, R, true);
    const uint16_t* t0 = sc.light + ((size_t)gc.j1 * R + gc.i0) * 4;
    const uint16_t* t1 = sc.light + ((size_t)gc.j1 * R + gc.i1) * 4;
    const uint16_t* t2 = sc.light + ((size_t)gc.j0 * R + gc.i1) * 4;
    const uint16_t* t3 = sc.light + ((size_t)gc.j0 * R + gc.i0) * 4;
    vec3 col;
    col.x = dot4({h2f(t0[0]), h2f(t1[0]), h2f(t2[0]), h2f(t3[0])}, bw);
    col.y = dot4({h2f(t0[1]), h2f(t1[1]), h2f(t2[1]), h2f(t3[1])}, bw);
    col.z = dot4({h2f(t0[2]), h2f(t1[2]), h2f(t2[2]), h2f(t3[2])}, bw);
    return {col.x * sampleWeight, col.y * sampleWeight, col.z * sampleWeight, sampleWeight};
}

// AtlasCommon.glsl:115-157 (macro sampleGlobalSurfaceAtlas, debug = false)
vec4 sampleGlobalSurfaceAtlas(const Scene& sc, vec3 worldPosition, vec3 worldNormal, float surfaceThreshold, Counters& cn)
{
    vec4 result = {0, 0, 0, 0};
    if (!sc.hasAtlas)
        return result;
    const LuxGlobalSurfaceAtlasData& data = sc.atlasData;
    const float half = (float)LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * 0.5f;
    int cx = iclamp((int)std::floor(worldPosition.x / data.chunkSize + half), 0, LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION - 1);
    int cy = iclamp((int)std::floor(worldPosition.y / data.chunkSize + half), 0, LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION - 1);
    int cz = iclamp((int)std::floor(worldPosition.z / data.chunkSize + half), 0, LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION - 1);
    uint32_t chunkAddress = (uint32_t)(cz * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION +
                                       cy * LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION + cx); // flattenId :40-48
    uint32_t objectsStart = sc.chunks[chunkAddress];
    if (objectsStart == 0)
        return result;
    uint32_t objectsCount = sc.cull[objectsStart];
    if (objectsCount > data.objectsCount)
        return result;
    objectsStart++;
    for (uint32_t k = 0; k < objectsCount; k++)
    {
        uint32_t               objectAddress = sc.cull[objectsStart++];
        const LuxObjectBuffer& object        = sc.objects[objectAddress];
        cn.objectsVisited++;
        vec3 bc = {object.objectBounds[0], object.objectBounds[1], object.objectBounds[2]};
        if (length3(sub(bc, worldPosition)) > object.objectBounds[3])
            continue;
        float worldToLocal[16];
        inverse4(object.transform, worldToLocal);
        vec3 localPosition = mat4_mul_point(worldToLocal, worldPosition, 1.0f);
        vec3 localExtents  = {object.extends[0] + surfaceThreshold, object.extends[1] + surfaceThreshold,
                              object.extends[2] + surfaceThreshold};
        if (std::fabs(localPosition.x) > localExtents.x || std::fabs(localPosition.y) > localExtents.y ||
            std::fabs(localPosition.z) > localExtents.z)
            continue;
        vec3 normal = normalize3(mat3_mul(worldToLocal, worldNormal));
        for (int i = 0; i < 6; i++)
        {
            uint32_t tileOffset = object.tileOffset[i];
            if (tileOffset != 0)
            {
                vec4 s = sampleGlobalSurfaceAtlasTile(sc, sc.tiles[tileOffset], localPosition, normal, surfaceThreshold, cn);
                result.x += s.x; result.y += s.y; result.z += s.z; result.w += s.w;
            }
        }
    }
    float d = gmax(result.w, 0.0001f);
    result.x /= d; result.y /= d; result.z /= d;
    return result;
}

// texture(samplerCube, dir).rgb — Vulkan cube face selection (spec §16.5.4), bilinear inside the face,
// clamp at face edges.  The reference binds a 1x1 fallback cube when the scene has no skybox.
vec4 sampleSky4(const Scene& sc, vec3 d)
{
    if (!sc.sky || sc.skyFace <= 0)
        return {0, 0, 0, 0};
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int   face;
    float sc_, tc, ma;
    if (az >= ax && az >= ay) { face = d.z >= 0 ? 4 : 5; sc_ = d.z >= 0 ? d.x : -d.x; tc = -d.y; ma = az; }
    else if (ay >= ax)        { face = d.y >= 0 ? 2 : 3; sc_ = d.x; tc = d.y >= 0 ? d.z : -d.z; ma = ay; }
    else                      { face = d.x >= 0 ? 0 : 1; sc_ = d.x >= 0 ? -d.z : d.z; tc = -d.y; ma = ax; }
    float u = 0.5f * (sc_ / ma) + 0.5f, v = 0.5f * (tc / ma) + 0.5f;
    int   N = sc.skyFace;
    float x = u * (float)N - 0.5f, y = v * (float)N - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float axw = x - fx, ayw = y - fy;
    int   x0 = iclamp((int)fx, 0, N - 1), x1 = iclamp((int)fx + 1, 0, N - 1);
    int   y0 = iclamp((int)fy, 0, N - 1), y1 = iclamp((int)fy + 1, 0, N - 1);
    const uint16_t* base = sc.sky + (size_t)face * N * N * 4;
    float out[4];
    for (int ch = 0; ch < 4; ch++)
    {
        float a = lerp1(h2f(base[((size_t)y0 * N + x0) * 4 + ch]), h2f(base[((size_t)y0 * N + x1) * 4 + ch]), axw);
        float b = lerp1(h2f(base[((size_t)y1 * N + x0) * 4 + ch]), h2f(base[((size_t)y1 * N + x1) * 4 + ch]), axw);
        out[ch] = lerp1(a, b, ayw);
    }
    return {out[0], out[1], out[2], out[3]};
}
inline vec3 sampleSky(const Scene& sc, vec3 d)
{
    vec4 s = sampleSky4(sc, d);
    return {s.x, s.y, s.z};
}

// ---------------------------------------------------------------------------------------------------------
// GISDFRays.comp:63-128 — one probe ray
// ---------------------------------------------------------------------------------------------------------
void traceOneRay(const Scene& sc, const float* rot, int rayId, int probeId, uint16_t* radianceOut, uint16_t* dirDistOut,
                 uint16_t* stepsOut, Counters& cn)
{
    const LuxDDGIUniform& ddgi = sc.ddgi;
    vec3 rayOrigin = probeLocation(ddgi, probeId);
    vec3 direction = normalize3(mat3_mul(rot, sphericalFibonacci((float)rayId, (float)ddgi.raysPerProbe)));

    Hit hit = tracyGlobalSDF(sc, rayOrigin, direction, LUX_GLOBAL_SDF_WORLD_SIZE, 1.0f, 0.0f, cn);

    vec4 radiance = {0, 0, 0, 0};
    if (hit.hitTime >= 0.0f) // isHit, SDFCommon.glsl:73-76
    {
        cn.hits++;
        if (hit.hitSDF <= 0.0f && hit.hitTime <= sc.sdfData.cascadeVoxelSize[0])
            radiance = {0, 0, 0, LUX_GLOBAL_SDF_WORLD_SIZE};
        else
        {
            vec3  hitPosition      = add(rayOrigin, mul(direction, hit.hitTime)); // getHitPosition :78-81
            float surfaceThreshold = sc.sdfData.cascadeVoxelSize[hit.hitCascade] * 1.05f; // :196-199
            vec4  surfaceColor     = sampleGlobalSurfaceAtlas(sc, hitPosition, hit.hitNormal, surfaceThreshold, cn);
            radiance   = {surfaceColor.x, surfaceColor.y, surfaceColor.z, hit.hitTime};
            radiance.w = gmax(radiance.w + sc.sdfData.cascadeVoxelSize[hit.hitCascade] * 0.5f, 0.0f);
        }
    }
    else
    {
        vec3 s   = sampleSky(sc, direction);
        radiance = {s.x, s.y, s.z, LUX_GLOBAL_SDF_WORLD_SIZE};
    }
    radianceOut[0] = f2h(radiance.x); radianceOut[1] = f2h(radiance.y); radianceOut[2] = f2h(radiance.z); radianceOut[3] = f2h(0.0f);
    dirDistOut[0] = f2h(direction.x); dirDistOut[1] = f2h(direction.y); dirDistOut[2] = f2h(direction.z); dirDistOut[3] = f2h(radiance.w);
    if (stepsOut)
        *stepsOut = (uint16_t)hit.stepsCount;
}

const float FLT_EPS = 0.00000001f; // ProbeUpdate.glsl:51

// Contract form: blend accumulation and mix() use one explicit FMA each.  `g_unfused` switches both to the
// literal two-rounding form of the shipped SPIR-V (OpVectorTimesScalar + OpFAdd; FMix = x*(1-a) + y*a) so that the
// restatement can be compared bit for bit with the interpreter-executed reference binaries (tests/test_spirv_golden.py).
bool g_unfused = false;
inline float mixh(float x, float y, float a) { return g_unfused ? x * (1.0f - a) + y * a : std::fmaf(y, a, x * (1.0f - a)); }
inline float madd(float a, float b, float c) { return g_unfused ? a * b + c : std::fmaf(a, b, c); }

// ---------------------------------------------------------------------------------------------------------
// ProbeUpdate.glsl:66-152 for one interior texel of one probe.  `naive` keeps octDecode/pow inside the ray loop
// exactly as written; the hoisted form (weights precomputed per (texel, ray)) is bit-identical because both
// are pure functions of (texel, fp16 ray direction).
// ---------------------------------------------------------------------------------------------------------
struct BlendArgs
{
    const LuxDDGIUniform* ddgi;
    const uint16_t*       radiance; // [nprobes][R][4]
    const uint16_t*       dirDist;  // [nprobes][R][4]
    int                   rayRowOffset; // probe id of row 0 of the ray buffers
    const uint16_t*       prevIrr;
    const uint16_t*       prevDepth;
    uint16_t*             outIrr;
    uint16_t*             outDepth;
    int                   firstFrame;
};

void blendIrradianceTexel(const BlendArgs& a, int probe, int i, int j, const float* wRow /*nullable: hoisted weights [R]*/)
{
    const LuxDDGIUniform& d = *a.ddgi;
    const int side = d.irradianceProbeSideLength, S = side + 2, W = d.irradianceTextureWidth;
    const int perRow = (W - 2) / S;
    const int px = probe % perRow, py = probe / perRow;
    const int cx = 2 + px * S + i, cy = 2 + py * S + j; // ProbeUpdate.glsl:107
    const int R = d.raysPerProbe;
    const uint16_t* rad = a.radiance + (size_t)(probe - a.rayRowOffset) * R * 4;
    const uint16_t* dd  = a.dirDist + (size_t)(probe - a.rayRowOffset) * R * 4;

    float rx = 0, ry = 0, rz = 0, total = 0;
    for (int r = 0; r < R; r++)
    {
        float weight;
        if (wRow)
            weight = wRow[r];
        else
        {
            vec3 rayDirection   = {h2f(dd[r * 4 + 0]), h2f(dd[r * 4 + 1]), h2f(dd[r * 4 + 2])};
            vec3 texelDirection = octDecode(normalizedOctCoordLocal(i, j, side));
            weight              = gmax(0.0f, dot3(texelDirection, rayDirection));
        }
        if (weight >= FLT_EPS)
        {
            rx = madd(h2f(rad[r * 4 + 0]), weight, rx);
            ry = madd(h2f(rad[r * 4 + 1]), weight, ry);
            rz = madd(h2f(rad[r * 4 + 2]), weight, rz);
            total += weight;
        }
    }
    if (total > FLT_EPS)
    {
        float s = 1.0f / (2.0f * total);
        rx *= s; ry *= s; rz *= s;
    }
    float ig = 1.0f / d.ddgiGamma;
    rx = pow_rn(rx, ig); ry = pow_rn(ry, ig); rz = pow_rn(rz, ig);
    size_t o = ((size_t)cy * W + cx) * 4;
    if (!a.firstFrame)
    {
        rx = mixh(rx, h2f(a.prevIrr[o + 0]), d.hysteresis);
        ry = mixh(ry, h2f(a.prevIrr[o + 1]), d.hysteresis);
        rz = mixh(rz, h2f(a.prevIrr[o + 2]), d.hysteresis);
    }
    a.outIrr[o + 0] = f2h(rx); a.outIrr[o + 1] = f2h(ry); a.outIrr[o + 2] = f2h(rz); a.outIrr[o + 3] = f2h(1.0f);
}

void blendDepthTexel(const BlendArgs& a, int probe, int i, int j, const float* wRow)
{
    const LuxDDGIUniform& d = *a.ddgi;
    const int side = d.depthProbeSideLength, S = side + 2, W = d.depthTextureWidth;
    const int perRow = (W - 2) / S;
    const int px = probe % perRow, py = probe / perRow;
    const int cx = 2 + px * S + i, cy = 2 + py * S + j;
    const int R = d.raysPerProbe;
    const uint16_t* dd = a.dirDist + (size_t)(probe - a.rayRowOffset) * R * 4;

    float rx = 0, ry = 0, total = 0;
    for (int r = 0; r < R; r++)
    {
        float rayProbeDistance = gmin(d.maxDistance, h2f(dd[r * 4 + 3]) - 0.01f);
        if (rayProbeDistance == -1.0f)
            rayProbeDistance = d.maxDistance;
        float weight;
        if (wRow)
            weight = wRow[r];
        else
        {
            vec3 rayDirection   = {h2f(dd[r * 4 + 0]), h2f(dd[r * 4 + 1]), h2f(dd[r * 4 + 2])};
            vec3 texelDirection = octDecode(normalizedOctCoordLocal(i, j, side));
            weight              = pow_rn(gmax(0.0f, dot3(texelDirection, rayDirection)), d.sharpness);
        }
        if (weight >= FLT_EPS)
        {
            rx = madd(rayProbeDistance, weight, rx);
            ry = madd(rayProbeDistance * rayProbeDistance, weight, ry);
            total += weight;
        }
    }
    if (total > FLT_EPS)
    {
        float s = 1.0f / (2.0f * total);
        rx *= s; ry *= s;
    }
    size_t o = ((size_t)cy * W + cx) * 2;
    if (!a.firstFrame)
    {
        rx = mixh(rx, h2f(a.prevDepth[o + 0]), d.hysteresis);
        ry = mixh(ry, h2f(a.prevDepth[o + 1]), d.hysteresis);
    }
    a.outDepth[o + 0] = f2h(rx); a.outDepth[o + 1] = f2h(ry);
}

// BorderUpdate.glsl:25-133,136-156 — the offset tables reduce to this mirror rule (checked entry by entry
// against tests/golden/border_offsets.json): for x,y in 1..side (coordinates relative to the probe's ring origin)
//   (x,0) <- (side+1-x, 1)   (x,side+1) <- (side+1-x, side)   (0,y) <- (1, side+1-y)   (side+1,y) <- (side, side+1-y)
//   corners: (0,0)<-(side,side) (side+1,0)<-(1,side) (0,side+1)<-(side,1) (side+1,side+1)<-(1,1)
void borderProbe(uint16_t* img, int W, int channels, int side, int probe)
{
    const int S = side + 2, perRow = (W - 2) / S;
    const int bx = (probe % perRow) * S + 1, by = (probe / perRow) * S + 1; // BorderUpdate.glsl:150
    auto cp = [&](int sx, int sy, int dx, int dy) {
        std::memcpy(img + ((size_t)(by + dy) * W + (bx + dx)) * channels, img + ((size_t)(by + sy) * W + (bx + sx)) * channels,
                    sizeof(uint16_t) * channels);
    };
    for (int x = 1; x <= side; x++)
    {
        cp(side + 1 - x, 1, x, 0);
        cp(side + 1 - x, side, x, side + 1);
    }
    for (int y = 1; y <= side; y++)
    {
        cp(1, side + 1 - y, 0, y);
        cp(side, side + 1 - y, side + 1, y);
    }
    cp(side, side, 0, 0);
    cp(1, side, side + 1, 0);
    cp(side, 1, 0, side + 1);
    cp(1, 1, side + 1, side + 1);
}

// ---------------------------------------------------------------------------------------------------------
// Consumer side ("next" row f2): sampleIrradiance + SampleProbe.comp
// ---------------------------------------------------------------------------------------------------------
// textureLod(sampler2D, uv, 0) on an fp16 atlas: bilinear, REPEAT wrap (default Texture2D sampler, RHI/Definitions.h:152-159)
struct Atlas2D
{
    const uint16_t* d;
    int             w, h, c;
};

inline int wrapi(int i, int n) { return ((i % n) + n) % n; }

void sampleAtlas(const Atlas2D& t, float u, float v, float* out)
{
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float ax = x - fx, ay = y - fy;
    int   x0 = wrapi((int)fx, t.w), x1 = wrapi((int)fx + 1, t.w), y0 = wrapi((int)fy, t.h), y1 = wrapi((int)fy + 1, t.h);
    for (int k = 0; k < t.c; k++)
    {
        float a = lerp1(h2f(t.d[((size_t)y0 * t.w + x0) * t.c + k]), h2f(t.d[((size_t)y0 * t.w + x1) * t.c + k]), ax);
        float b = lerp1(h2f(t.d[((size_t)y1 * t.w + x0) * t.c + k]), h2f(t.d[((size_t)y1 * t.w + x1) * t.c + k]), ax);
        out[k]  = lerp1(a, b, ay);
    }
}

// DDGICommon.glsl:60-72
inline vec2 octEncode(vec3 v)
{
    float l1norm = (std::fabs(v.x) + std::fabs(v.y)) + std::fabs(v.z);
    float inv    = 1.0f / l1norm;
    vec2  r      = {v.x * inv, v.y * inv};
    if (v.z < 0.0f)
    {
        vec2 q = {(1.0f - std::fabs(r.y)) * signNotZero(r.x), (1.0f - std::fabs(r.x)) * signNotZero(r.y)};
        r      = q;
    }
    return r;
}

// DDGICommon.glsl:141-158
inline vec2 textureCoordFromDirection(vec3 dir, int probeIndex, int width, int height, int probeSideLength)
{
    vec2  oc  = octEncode(normalize3(dir));
    vec2  o01 = {(oc.x + 1.0f) * 0.5f, (oc.y + 1.0f) * 0.5f};
    float probeWithBorderSide = (float)probeSideLength + 2.0f;
    vec2  octTex = {(o01.x * (float)probeSideLength) / (float)width, (o01.y * (float)probeSideLength) / (float)height};
    int   probesPerRow = (width - 2) / (int)probeWithBorderSide;
    float fi = (float)probeIndex, fp = (float)probesPerRow;
    float modv = fi - fp * std::floor(fi / fp); // mod(probeIndex, probesPerRow) on floats
    vec2  topLeft = {modv * probeWithBorderSide + 2.0f, (float)(probeIndex / probesPerRow) * probeWithBorderSide + 2.0f};
    vec2  topLeftN = {topLeft.x / (float)width, topLeft.y / (float)height};
    return {topLeftN.x + octTex.x, topLeftN.y + octTex.y};
}

inline float square1(float v) { return v * v; }

// DDGICommon.glsl:163-233
vec3 sampleIrradiance(const LuxDDGIUniform& ddgi, vec3 P, vec3 N, vec3 Wo, const Atlas2D& irr, const Atlas2D& dep)
{
    int   bg[3], cnt[3] = {ddgi.probeCounts[0], ddgi.probeCounts[1], ddgi.probeCounts[2]};
    float Pv[3] = {P.x, P.y, P.z};
    for (int a = 0; a < 3; a++) // baseGridCoord :124-127: clamp(ivec3((X - start) / step), 0, counts - 1)
        bg[a] = iclamp((int)((Pv[a] - ddgi.startPosition[a]) / ddgi.step[a]), 0, cnt[a] - 1);
    vec3 baseProbePos = {ddgi.step[0] * (float)bg[0] + ddgi.startPosition[0], ddgi.step[1] * (float)bg[1] + ddgi.startPosition[1],
                         ddgi.step[2] * (float)bg[2] + ddgi.startPosition[2]};
    vec3 sumIrradiance = {0, 0, 0};
    float sumWeight = 0.0f;
    vec3 alpha = {gclamp((P.x - baseProbePos.x) / ddgi.step[0], 0.0f, 1.0f), gclamp((P.y - baseProbePos.y) / ddgi.step[1], 0.0f, 1.0f),
                  gclamp((P.z - baseProbePos.z) / ddgi.step[2], 0.0f, 1.0f)};
    for (int i = 0; i < 8; ++i)
    {
        int  off[3] = {i & 1, (i >> 1) & 1, (i >> 2) & 1};
        int  pg[3];
        for (int a = 0; a < 3; a++)
            pg[a] = iclamp(bg[a] + off[a], 0, cnt[a] - 1);
        vec3 probePos = {ddgi.step[0] * (float)pg[0] + ddgi.startPosition[0], ddgi.step[1] * (float)pg[1] + ddgi.startPosition[1],
                         ddgi.step[2] * (float)pg[2] + ddgi.startPosition[2]};
        // mix(1 - alpha, alpha, offset) = x*(1-a) + y*a with a in {0, 1}
        float al[3] = {alpha.x, alpha.y, alpha.z}, tri[3];
        for (int a = 0; a < 3; a++)
            tri[a] = (1.0f - al[a]) * (1.0f - (float)off[a]) + al[a] * (float)off[a];
        float weight = 1.0f;
        vec3  dirToProbe = normalize3(sub(probePos, P));
        weight *= square1(gmax(0.0001f, (dot3(dirToProbe, N) + 1.0f) * 0.5f)) + 0.2f;
        int probeIdx = pg[0] + pg[1] * cnt[0] + pg[2] * cnt[0] * cnt[1];

        vec3  vBias        = mul(add(N, mul(Wo, 3.0f)), ddgi.normalBias);
        vec3  probeToPoint = add(sub(P, probePos), vBias);
        vec3  dir          = normalize3({-probeToPoint.x, -probeToPoint.y, -probeToPoint.z});
        vec2  tc           = textureCoordFromDirection({-dir.x, -dir.y, -dir.z}, probeIdx, ddgi.depthTextureWidth, ddgi.depthTextureHeight,
                                                       ddgi.depthProbeSideLength);
        float dist = length3(probeToPoint);
        float tmp[4];
        sampleAtlas(dep, tc.x, tc.y, tmp);
        float mean     = tmp[0];
        float variance = std::fabs(square1(tmp[0]) - tmp[1]);
        float cheb     = variance / (variance + square1(gmax(dist - mean, 0.0f)));
        cheb           = gmax(cheb * cheb * cheb, 0.0f);
        weight *= (dist <= mean) ? 1.0f : cheb;
        weight = gmax(0.000001f, weight);

        tc = textureCoordFromDirection(normalize3(N), probeIdx, ddgi.irradianceTextureWidth, ddgi.irradianceTextureHeight,
                                       ddgi.irradianceProbeSideLength);
        sampleAtlas(irr, tc.x, tc.y, tmp);
        float e = ddgi.ddgiGamma * 0.5f;
        vec3  probeIrradiance = {pow_rn(tmp[0], e), pow_rn(tmp[1], e), pow_rn(tmp[2], e)};
        const float crushThreshold = 0.2f;
        if (weight < crushThreshold)
            weight *= weight * weight * (1.0f / square1(crushThreshold));
        weight *= tri[0] * tri[1] * tri[2];
        sumIrradiance = add(sumIrradiance, mul(probeIrradiance, weight));
        sumWeight += weight;
    }
    vec3 net = {sumIrradiance.x / sumWeight, sumIrradiance.y / sumWeight, sumIrradiance.z / sumWeight};
    net      = mul(net, net);
    const float TWO_PI_F = 6.283185482025146484375f; // 2 * PI folded to float
    return mul(net, TWO_PI_F);
}

// Common/Math.glsl:27-33
inline vec3 octohedralToDirection(vec2 e)
{
    vec3 v = {e.x, e.y, (1.0f - std::fabs(e.x)) - std::fabs(e.y)};
    if (v.z < 0.0f)
    {
        float sx = (v.x >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f, sy = (v.y >= 0.0f ? 1.0f : 0.0f) * 2.0f - 1.0f; // step(0, v) * 2 - 1
        float nx = (1.0f - std::fabs(v.y)) * sx, ny = (1.0f - std::fabs(v.x)) * sy;
        v.x = nx;
        v.y = ny;
    }
    return normalize3(v);
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench only)
// ---------------------------------------------------------------------------------------------------------
extern "C" {

struct OracleSceneDesc
{
    const LuxDDGIUniform*            ddgi;
    const LuxGlobalSDFData*          sdfData;
    const uint16_t*                  sdf;
    const uint16_t*                  mip;
    const LuxGlobalSurfaceAtlasData* atlasData; // null = no surface cache (radiance 0 on hit)
    const uint32_t*                  chunks;
    const uint32_t*                  cull;
    const LuxObjectBuffer*           objects;
    const LuxTileBuffer*             tiles;
    const uint16_t*                  light;
    const float*                     depth;
    int32_t                          skyFace;
    const uint16_t*                  sky;
};

struct OracleCounters
{
    uint64_t mipTaps, texTaps, hits, tileSamples, steps, objectsVisited;
};

int oracle_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void oracle_set_unfused(int on) { g_unfused = on != 0; }

uint16_t oracle_f2h(float f) { return f2h(f); }
float    oracle_h2f(uint16_t h) { return h2f(h); }

void oracle_spherical_fibonacci(int i, int R, const float* rot, float* out3)
{
    vec3 v = sphericalFibonacci((float)i, (float)R);
    if (rot)
        v = normalize3(mat3_mul(rot, v));
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

void oracle_oct_decode(int i, int j, int side, float* out3)
{
    vec3 v = octDecode(normalizedOctCoordLocal(i, j, side));
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

void oracle_probe_location(const LuxDDGIUniform* d, int index, float* out3)
{
    vec3 v = probeLocation(*d, index);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

void oracle_inverse4(const float* m, float* o) { inverse4(m, o); }

float oracle_sample3d(const uint16_t* data, int w, int h, int d, float u, float v, float ww)
{
    Tex3D t{data, w, h, d};
    return sample3D(t, u, v, ww, nullptr);
}

// Trace `count` probes.  probeIds == null: probes probeBegin .. probeBegin+count-1; else the listed ids.
// Output row k (of `count`) holds probe k's R rays.  stepsOut (nullable) [count][R] u16.
int oracle_trace(const OracleSceneDesc* s, const float* rot16, int probeBegin, int count, const int32_t* probeIds,
                 uint16_t* radiance, uint16_t* dirDist, uint16_t* stepsOut, OracleCounters* counters)
{
    if (!s || !s->ddgi || !s->sdfData || !s->sdf || !s->mip || !rot16 || !radiance || !dirDist)
        return -1;
    Scene sc;
    sc.ddgi    = *s->ddgi;
    sc.sdfData = *s->sdfData;
    int res    = (int)s->sdfData->resolution;
    int casc   = (int)s->sdfData->cascadesCount;
    sc.tex     = Tex3D{s->sdf, res * casc, res, res};
    sc.mip     = Tex3D{s->mip, (res / 4) * casc, res / 4, res / 4};
    sc.hasAtlas = s->atlasData != nullptr;
    if (sc.hasAtlas)
        sc.atlasData = *s->atlasData;
    sc.chunks = s->chunks; sc.cull = s->cull; sc.objects = s->objects; sc.tiles = s->tiles;
    sc.light = s->light; sc.depth = s->depth;
    sc.skyFace = s->skyFace; sc.sky = s->sky;
    const int R = sc.ddgi.raysPerProbe;

    Counters total;
#pragma omp parallel
    {
        Counters cn;
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < count; k++)
        {
            int probe = probeIds ? probeIds[k] : probeBegin + k;
            for (int r = 0; r < R; r++)
            {
                size_t o = ((size_t)k * R + r);
                traceOneRay(sc, rot16, r, probe, radiance + o * 4, dirDist + o * 4, stepsOut ? stepsOut + o : nullptr, cn);
            }
        }
#pragma omp critical
        {
            total.mipTaps += cn.mipTaps; total.texTaps += cn.texTaps; total.hits += cn.hits;
            total.tileSamples += cn.tileSamples; total.steps += cn.steps; total.objectsVisited += cn.objectsVisited;
        }
    }
    if (counters)
    {
        counters->mipTaps = total.mipTaps; counters->texTaps = total.texTaps; counters->hits = total.hits;
        counters->tileSamples = total.tileSamples; counters->steps = total.steps; counters->objectsVisited = total.objectsVisited;
    }
    return 0;
}

// Blend probes [probeBegin, probeBegin+count), or, when probeIds != null, the listed probes: ray-buffer row k then belongs to
// probe probeIds[k] (stratified subsamples of large volumes).  Otherwise the ray buffers hold rows for probes rayRowOffset...
// naive != 0: literal per-(texel, ray) evaluation as in the shader (this is the timed CPU baseline);
// naive == 0: weights hoisted per (texel, ray) once (bit-identical; see test_oracle_kat.py).
int oracle_blend_ids(const LuxDDGIUniform* ddgi, const uint16_t* radiance, const uint16_t* dirDist, int rayRowOffset,
                     const uint16_t* prevIrr, const uint16_t* prevDepth, uint16_t* outIrr, uint16_t* outDepth, int firstFrame,
                     int probeBegin, int count, const int32_t* probeIds, int naive)
{
    if (!ddgi || !radiance || !dirDist || !outIrr || !outDepth)
        return -1;
    if (!firstFrame && (!prevIrr || !prevDepth))
        return -1;
    const int R = ddgi->raysPerProbe, si = ddgi->irradianceProbeSideLength, sd = ddgi->depthProbeSideLength;

    std::vector<float> wi, wd;
    if (!naive && count > 0)
    {
        // Ray directions are probe-independent (GISDFRays.comp:73): take them from the first row.
        const uint16_t* dd = probeIds ? dirDist : dirDist + (size_t)(probeBegin - rayRowOffset) * R * 4;
        wi.resize((size_t)si * si * R);
        wd.resize((size_t)sd * sd * R);
        for (int j = 0; j < si; j++)
            for (int i = 0; i < si; i++)
            {
                vec3 t = octDecode(normalizedOctCoordLocal(i, j, si));
                for (int r = 0; r < R; r++)
                    wi[((size_t)j * si + i) * R + r] = gmax(0.0f, dot3(t, {h2f(dd[r * 4]), h2f(dd[r * 4 + 1]), h2f(dd[r * 4 + 2])}));
            }
#pragma omp parallel for schedule(static)
        for (int j = 0; j < sd; j++)
            for (int i = 0; i < sd; i++)
            {
                vec3 t = octDecode(normalizedOctCoordLocal(i, j, sd));
                for (int r = 0; r < R; r++)
                    wd[((size_t)j * sd + i) * R + r] =
                        pow_rn(gmax(0.0f, dot3(t, {h2f(dd[r * 4]), h2f(dd[r * 4 + 1]), h2f(dd[r * 4 + 2])})), ddgi->sharpness);
            }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < count; k++)
    {
        int probe = probeIds ? probeIds[k] : probeBegin + k;
        // with a probe list, row k of the ray buffers is probe probeIds[k]: express that through a per-probe row offset
        BlendArgs a{ddgi, radiance, dirDist, probeIds ? probe - k : rayRowOffset, prevIrr, prevDepth, outIrr, outDepth, firstFrame};
        for (int j = 0; j < si; j++)
            for (int i = 0; i < si; i++)
                blendIrradianceTexel(a, probe, i, j, naive ? nullptr : &wi[((size_t)j * si + i) * R]);
        for (int j = 0; j < sd; j++)
            for (int i = 0; i < sd; i++)
                blendDepthTexel(a, probe, i, j, naive ? nullptr : &wd[((size_t)j * sd + i) * R]);
    }
    return 0;
}

int oracle_blend(const LuxDDGIUniform* ddgi, const uint16_t* radiance, const uint16_t* dirDist, int rayRowOffset,
                 const uint16_t* prevIrr, const uint16_t* prevDepth, uint16_t* outIrr, uint16_t* outDepth, int firstFrame,
                 int probeBegin, int count, int naive)
{
    return oracle_blend_ids(ddgi, radiance, dirDist, rayRowOffset, prevIrr, prevDepth, outIrr, outDepth, firstFrame, probeBegin, count,
                            nullptr, naive);
}

int oracle_border_ids(const LuxDDGIUniform* ddgi, uint16_t* irr, uint16_t* depth, const int32_t* probeIds, int count)
{
    if (!ddgi || !probeIds)
        return -1;
    for (int k = 0; k < count; k++)
    {
        if (irr)
            borderProbe(irr, ddgi->irradianceTextureWidth, 4, ddgi->irradianceProbeSideLength, probeIds[k]);
        if (depth)
            borderProbe(depth, ddgi->depthTextureWidth, 2, ddgi->depthProbeSideLength, probeIds[k]);
    }
    return 0;
}

int oracle_border(const LuxDDGIUniform* ddgi, uint16_t* irr, uint16_t* depth, int probeBegin, int count)
{
    if (!ddgi)
        return -1;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < count; k++)
    {
        if (irr)
            borderProbe(irr, ddgi->irradianceTextureWidth, 4, ddgi->irradianceProbeSideLength, probeBegin + k);
        if (depth)
            borderProbe(depth, ddgi->depthTextureWidth, 2, ddgi->depthProbeSideLength, probeBegin + k);
    }
    return 0;
}

// sampleIrradiance (DDGICommon.glsl:163-233) for `count` points; P, N, Wo, out are [count][3] floats.
int oracle_sample_irradiance(const LuxDDGIUniform* ddgi, const uint16_t* irr, const uint16_t* depth, int count, const float* P, const float* N,
                             const float* Wo, float* out)
{
    if (!ddgi || !irr || !depth || !P || !N || !Wo || !out)
        return -1;
    Atlas2D ai{irr, ddgi->irradianceTextureWidth, ddgi->irradianceTextureHeight, 4};
    Atlas2D ad{depth, ddgi->depthTextureWidth, ddgi->depthTextureHeight, 2};
#pragma omp parallel for schedule(static)
    for (int k = 0; k < count; k++)
    {
        vec3 r = sampleIrradiance(*ddgi, {P[3 * k], P[3 * k + 1], P[3 * k + 2]}, {N[3 * k], N[3 * k + 1], N[3 * k + 2]},
                                  {Wo[3 * k], Wo[3 * k + 1], Wo[3 * k + 2]}, ai, ad);
        out[3 * k] = r.x; out[3 * k + 1] = r.y; out[3 * k + 2] = r.z;
    }
    return 0;
}

// SampleProbe.comp:36-60 over a width x height G-buffer: depth [h][w] (D32F), normals [h][w][4] (RGBA32F, xy = octahedral normal),
// cameraPosition[4], viewProjInv (column-major mat4); out [h][w][4] floats (the INDIRECT_LIGHTING target is RGBA32F, GBuffer.cpp:24).
int oracle_sample_probe(const LuxDDGIUniform* ddgi, const uint16_t* irr, const uint16_t* depthAtlas, int width, int height,
                        const float* gDepth, const float* gNormal, const float* cameraPosition, const float* viewProjInv, float* out)
{
    if (!ddgi || !irr || !depthAtlas || !gDepth || !gNormal || !cameraPosition || !viewProjInv || !out)
        return -1;
    Atlas2D ai{irr, ddgi->irradianceTextureWidth, ddgi->irradianceTextureHeight, 4};
    Atlas2D ad{depthAtlas, ddgi->depthTextureWidth, ddgi->depthTextureHeight, 2};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            size_t o = (size_t)y * width + x;
            float  d = gDepth[o];
            if (d == 1.0f)
            {
                out[4 * o] = out[4 * o + 1] = out[4 * o + 2] = out[4 * o + 3] = 0.0f;
                continue;
            }
            // texCoord = (coord + 0.5) / size; worldPositionFromDepth (Common/Math.glsl:35-42)
            float tx = ((float)x + 0.5f) / (float)width, ty = ((float)y + 0.5f) / (float)height;
            float sx = tx * 2.0f - 1.0f, sy = ty * 2.0f - 1.0f;
            const float* m = viewProjInv;
            float wx = ((m[0] * sx + m[4] * sy) + m[8] * d) + m[12] * 1.0f;
            float wy = ((m[1] * sx + m[5] * sy) + m[9] * d) + m[13] * 1.0f;
            float wz = ((m[2] * sx + m[6] * sy) + m[10] * d) + m[14] * 1.0f;
            float ww = ((m[3] * sx + m[7] * sy) + m[11] * d) + m[15] * 1.0f;
            vec3  Pw = {wx / ww, wy / ww, wz / ww};
            vec3  Nn = octohedralToDirection({gNormal[4 * o], gNormal[4 * o + 1]});
            vec3  Wo = normalize3(sub({cameraPosition[0], cameraPosition[1], cameraPosition[2]}, Pw));
            vec3  r  = sampleIrradiance(*ddgi, Pw, Nn, Wo, ai, ad);
            out[4 * o] = r.x; out[4 * o + 1] = r.y; out[4 * o + 2] = r.z; out[4 * o + 3] = 1.0f;
        }
    return 0;
}

// Indirect-light refresh of the surface light cache ("next" row f1): for each listed atlas texel
//   light.rgb = fp16( base.rgb + intensity * (min(albedo, 0.9) - min(albedo, 0.9) * metallic) / PI * sampleIrradiance(P, N, Wo) )
// with Wo = normalize(cameraPos - P)  (Shaders/SDF/SDFAtlasIndirectLight.frag:44-67; additive blend into the RGBA16F light cache,
// GlobalSurfaceAtlas.cpp:1004-1085).  `light` is updated in place; base == null means "add to the current contents".
int oracle_indirect_light(const LuxDDGIUniform* ddgi, const uint16_t* irr, const uint16_t* depth, uint16_t* light, const uint16_t* base, int count,
                          const uint32_t* texel, const float* P, const float* N, const float* albedo, const float* metallic, float intensity,
                          const float* cameraPos)
{
    if (!ddgi || !irr || !depth || !light || !texel || !P || !N || !albedo || !metallic || !cameraPos)
        return -1;
    Atlas2D ai{irr, ddgi->irradianceTextureWidth, ddgi->irradianceTextureHeight, 4};
    Atlas2D ad{depth, ddgi->depthTextureWidth, ddgi->depthTextureHeight, 2};
    const float PI_F = 3.1415926535897932384626433832795f;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < count; k++)
    {
        vec3 Pw = {P[3 * k], P[3 * k + 1], P[3 * k + 2]}, Nn = {N[3 * k], N[3 * k + 1], N[3 * k + 2]};
        vec3 Wo = normalize3(sub({cameraPos[0], cameraPos[1], cameraPos[2]}, Pw));
        vec3 E  = sampleIrradiance(*ddgi, Pw, Nn, Wo, ai, ad);
        float Ev[3] = {E.x, E.y, E.z};
        size_t o = (size_t)texel[k] * 4;
        for (int c = 0; c < 3; c++)
        {
            float a       = gmin(albedo[3 * k + c], 0.9f);
            float diffuse = (a - a * metallic[k]) / PI_F;
            float out     = (intensity * diffuse) * Ev[c];
            float dst     = h2f(base ? base[o + c] : light[o + c]);
            light[o + c]  = f2h(dst + out);
        }
        light[o + 3] = f2h(h2f(base ? base[o + 3] : light[o + 3]) + 1.0f); // the shader writes alpha 1 and the pass blends ONE + ONE on alpha too (VulkanPipeline.cpp:160-165)
    }
    return 0;
}

} // extern "C"

// =====================================================================================================================
// Global SDF build ("next" row f3): the step before the path.
//   device side   Shaders/SDF/SDFRasterizeModel.glsl:42-63 (main), Shaders/SDF/SDFCommon.glsl:18-62 (combineDistanceToSDF,
//                 distanceToModelSDF), Shaders/SDF/GlobalSDFMipmap.comp:32-68
//   host side     Engine/DDGI/GlobalDistanceField.cpp:193-210 (getChunkId), :460-533 (chunkCalculate), :537-573 (fillFlood),
//                 :575-848 (merge_sdf::system, first frame: nothing cached), Math/BoundingBox.cpp:10-36, Math/BoundingSphere.cpp:9-13
// Reference behaviours kept on purpose (each changes the output, so "same inputs -> same volume" needs them):
//   * chunkCalculate's overflow loop assigns through a reference (`chunk = chunksCache[key]`, :515-519): when a chunk already
//     holds 28 models the layer-0 entry is overwritten by the (empty) next-layer entry, i.e. the list restarts; additive layers
//     never receive a model.  A chunk therefore keeps the LAST ((n-1) mod 28)+1 models registered for it.
//   * getChunkId discards its glm::clamp results (:202-203): chunk ranges are not clamped to the cascade.  Chunks outside the
//     volume would be out-of-bounds image stores (discarded under robust access); they are skipped here.
//   * BoundingBox::transform takes abs() of the product in its third term (BoundingBox.cpp:17-19).
//   * the flood passes add a WORLD-space voxel step to a NORMALISED distance (GlobalSDFMipmap.comp:44-45 with :831 maxDistance).
//   * mesh volumes are sampled with REPEAT addressing (Texture3D default wrap) at an integer LOD.
// =====================================================================================================================
namespace {

struct MeshTex // one mip level of a mesh distance field, R16F [z][y][x], trilinear, repeat
{
    const uint16_t* d;
    int             w, h, dd;
};

inline float sampleMesh(const MeshTex& t, float u, float v, float w)
{
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f, z = w * (float)t.dd - 0.5f;
    float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    float ax = x - fx, ay = y - fy, az = z - fz;
    int   ix = (int)fx, iy = (int)fy, iz = (int)fz;
    int   x0 = wrapi(ix, t.w), x1 = wrapi(ix + 1, t.w), y0 = wrapi(iy, t.h), y1 = wrapi(iy + 1, t.h), z0 = wrapi(iz, t.dd), z1 = wrapi(iz + 1, t.dd);
    auto  T = [&](int xx, int yy, int zz) { return h2f(t.d[((size_t)zz * t.h + yy) * t.w + xx]); };
    float c00 = lerp1(T(x0, y0, z0), T(x1, y0, z0), ax), c10 = lerp1(T(x0, y1, z0), T(x1, y1, z0), ax);
    float c01 = lerp1(T(x0, y0, z1), T(x1, y0, z1), ax), c11 = lerp1(T(x0, y1, z1), T(x1, y1, z1), ax);
    return lerp1(lerp1(c00, c10, ay), lerp1(c01, c11, ay), az);
}

// SDFCommon.glsl:18-39
inline float combineDistanceToSDF(float sdf, float distanceToSDF)
{
    if (sdf <= 0.0f && distanceToSDF <= 0.0f)
        return sdf;
    float maxSDF = gmax(sdf, 0.0f);
    return std::sqrt(maxSDF * maxSDF + distanceToSDF * distanceToSDF);
}

// SDFCommon.glsl:41-62
inline float distanceToModelSDF(float minDistance, const LuxObjectRasterizeData& m, const MeshTex& tex, vec3 worldPos)
{
    vec3 volumePos = mat4_mul_point(m.worldToVolume, worldPos, 1.0f);
    vec3 volumeUV  = {volumePos.x * m.volumeToUVWMul[0] + m.volumeToUVWAdd[0], volumePos.y * m.volumeToUVWMul[1] + m.volumeToUVWAdd[1],
                      volumePos.z * m.volumeToUVWMul[2] + m.volumeToUVWAdd[2]};
    vec3 e = {m.volumeLocalBoundsExtent[0], m.volumeLocalBoundsExtent[1], m.volumeLocalBoundsExtent[2]};
    vec3 volumePosClamped = {gclamp(volumePos.x, -e.x, e.x), gclamp(volumePos.y, -e.y, e.y), gclamp(volumePos.z, -e.z, e.z)};
    vec3 worldPosClamped  = mat4_mul_point(m.volumeToWorld, volumePosClamped, 1.0f);
    float distanceToVolume = length3(sub(worldPos, worldPosClamped));
    if (distanceToVolume < 0.01f)
        distanceToVolume = length3(sub(volumePos, volumePosClamped));
    distanceToVolume = gmax(distanceToVolume, 0.0f);
    if (minDistance <= distanceToVolume)
        return distanceToVolume;
    float volumeDistance = (sampleMesh(tex, volumeUV.x, volumeUV.y, volumeUV.z) * 2.0f - 1.0f) * m.decodeMul;
    float result = combineDistanceToSDF(volumeDistance, distanceToVolume);
    if (distanceToVolume > 0.0f)
        result = gmax(distanceToVolume, result);
    return result;
}


inline void mat4_mul(const float* a, const float* b, float* o) // column-major a*b, sums in k order
{
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            o[c * 4 + r] = ((a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1]) + a[2 * 4 + r] * b[c * 4 + 2]) + a[3 * 4 + r] * b[c * 4 + 3];
}

struct ChunkList { int coord[3]; int count; uint32_t models[LUX_SDF_RASTERIZE_MODEL_MAX_COUNT]; };

} // namespace

extern "C" {

// One SDFRasterizeModel dispatch: 32^3 voxels of chunk `chunkCoord` (voxel units) of cascade `cascadeIndex`.
// ubo = ModelsRasterizeData (SDFRasterizeModel.glsl:16-24).  meshTex[i] must already be the mip level objects[i].mipOffset selects.
int oracle_sdf_rasterize_chunk(const float* coordToPosMul3, const float* coordToPosAdd3, float maxDistance, int cascadeResolution, int cascadeIndex,
                               const int32_t* chunkCoord3, int objectsCount, const uint32_t* objectIds, int readDistance,
                               const LuxObjectRasterizeData* objects, const uint16_t* const* meshData, const int32_t* meshSizes /*[n][3]*/,
                               uint16_t* globalSDF, int texWidth)
{
    const int C = LUX_SDF_RASTERIZE_CHUNK_SIZE, res = cascadeResolution;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < C; z++)
        for (int y = 0; y < C; y++)
            for (int x = 0; x < C; x++)
            {
                int vx = chunkCoord3[0] + x, vy = chunkCoord3[1] + y, vz = chunkCoord3[2] + z;
                vec3 worldPos = {(float)vx * coordToPosMul3[0] + coordToPosAdd3[0], (float)vy * coordToPosMul3[1] + coordToPosAdd3[1],
                                 (float)vz * coordToPosMul3[2] + coordToPosAdd3[2]};
                int tx = vx + cascadeIndex * res;
                if (vx < 0 || vy < 0 || vz < 0 || vx >= res || vy >= res || vz >= res)
                    continue; // out-of-bounds image access
                size_t o = ((size_t)vz * res + vy) * texWidth + tx;
                float minDistance = maxDistance;
                if (readDistance)
                    minDistance *= h2f(globalSDF[o]);
                for (int i = 0; i < objectsCount; i++)
                {
                    uint32_t id = objectIds[i];
                    MeshTex  t{meshData[id], meshSizes[id * 3 + 0], meshSizes[id * 3 + 1], meshSizes[id * 3 + 2]};
                    float d = distanceToModelSDF(minDistance, objects[id], t, worldPos);
                    minDistance = gmin(minDistance, d);
                }
                globalSDF[o] = f2h(gclamp(minDistance / maxDistance, -1.0f, 1.0f));
            }
    return 0;
}

// One GlobalSDFMipmap dispatch over mipRes^3 outputs (GlobalSDFMipmap.comp:32-68).  src / dst are R16F [z][y][x] with the given row widths.
int oracle_sdf_mip_pass(const uint16_t* src, int srcWidth, int srcHeight, uint16_t* dst, int dstWidth, int dstHeight, int outRes, int globalSDFResolution,
                        int mipmapCoordScale, int cascadeTexOffsetX, int cascadeMipMapOffsetX, float maxDistance)
{
    static const int off[7][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < outRes; z++)
        for (int y = 0; y < outRes; y++)
            for (int x = 0; x < outRes; x++)
            {
                float minDistance = 0.0f;
                for (int k = 0; k < 7; k++)
                {
                    int cx = iclamp(x * mipmapCoordScale + off[k][0], 0, globalSDFResolution - 1);
                    int cy = iclamp(y * mipmapCoordScale + off[k][1], 0, globalSDFResolution - 1);
                    int cz = iclamp(z * mipmapCoordScale + off[k][2], 0, globalSDFResolution - 1);
                    float result = h2f(src[((size_t)cz * srcHeight + cy) * srcWidth + cx + cascadeTexOffsetX]);
                    float len = length3({(float)off[k][0], (float)off[k][1], (float)off[k][2]});
                    float distanceToVoxel = len * (maxDistance / (float)globalSDFResolution);
                    result = combineDistanceToSDF(result, distanceToVoxel);
                    minDistance = k == 0 ? result : gmin(minDistance, result);
                }
                dst[((size_t)z * dstHeight + y) * dstWidth + x + cascadeMipMapOffsetX] = f2h(minDistance);
            }
    return 0;
}

// Mip of every cascade: one downsample + 4 flood passes ping-ponging through a temporary (GlobalDistanceField.cpp:537-573, 825-841)
int oracle_sdf_build_mip(const LuxGlobalSDFData* data, const uint16_t* sdf, uint16_t* mip)
{
    const int res = (int)data->resolution, casc = (int)data->cascadesCount, mres = res / 4;
    std::vector<uint16_t> tmp((size_t)mres * mres * mres, f2h(1.0f));
    for (int c = 0; c < casc; c++)
    {
        float cascadeMaxDistance = data->cascadePosDistance[c][3] * 2.0f;
        oracle_sdf_mip_pass(sdf, res * casc, res, mip, mres * casc, mres, mres, res, 4, c * res, c * mres, cascadeMaxDistance);
        for (int i = 1; i < 5; i++)
        {
            if (i & 1)
                oracle_sdf_mip_pass(mip, mres * casc, mres, tmp.data(), mres, mres, mres, mres, 1, c * mres, 0, cascadeMaxDistance);
            else
                oracle_sdf_mip_pass(tmp.data(), mres, mres, mip, mres * casc, mres, mres, mres, 1, 0, c * mres, cascadeMaxDistance);
        }
    }
    return 0;
}

// ObjectRasterizeData of one mesh for cascade level `cascadeLevel` (chunkCalculate, GlobalDistanceField.cpp:484-508)
void oracle_sdf_object_data(const LuxMeshSDF* mesh, int cascadeLevel, LuxObjectRasterizeData* out)
{
    vec3 mn = {mesh->aabbMin[0], mesh->aabbMin[1], mesh->aabbMin[2]}, mx = {mesh->aabbMax[0], mesh->aabbMax[1], mesh->aabbMax[2]};
    vec3 volumeCenter = mul(add(mx, mn), 0.5f);
    float worldToLocal[16], tr[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, -volumeCenter.x, -volumeCenter.y, -volumeCenter.z, 1};
    inverse4(mesh->worldMatrix, worldToLocal);
    mat4_mul(worldToLocal, tr, out->worldToVolume);
    inverse4(out->worldToVolume, out->volumeToWorld);
    vec3 size = sub(mx, mn);
    out->volumeLocalBoundsExtent[0] = size.x / 2.0f; out->volumeLocalBoundsExtent[1] = size.y / 2.0f; out->volumeLocalBoundsExtent[2] = size.z / 2.0f;
    for (int i = 0; i < 3; i++)
    {
        out->volumeToUVWMul[i] = mesh->localToUVWMul[i];
        out->volumeToUVWAdd[i] = mesh->localToUVWAdd[i] + (&volumeCenter.x)[i] * mesh->localToUVWMul[i];
    }
    out->mipOffset = (float)(cascadeLevel < 2 ? cascadeLevel : 2);
    out->decodeMul = mesh->maxDistance;
    out->decodeAdd = -mesh->maxDistance;
}

// The whole one-shot build: global SDF [res][res][res*cascades] (cleared to 1.0 first) and its mip.
// chunkStats (nullable, 4 ints): chunks dispatched, models referenced, models dropped by the overflow bug, chunks skipped out of range
int oracle_sdf_build(const LuxGlobalSDFData* data, const LuxMeshSDF* meshes, int meshCount, float minObjectRadius, uint16_t* sdf, uint16_t* mip,
                     int32_t* chunkStats)
{
    if (!data || !meshes || !sdf || !mip)
        return -1;
    const int res = (int)data->resolution, casc = (int)data->cascadesCount, texWidth = res * casc;
    const int rasterizeChunks = (res + LUX_SDF_RASTERIZE_CHUNK_SIZE - 1) / LUX_SDF_RASTERIZE_CHUNK_SIZE;
    for (size_t i = 0; i < (size_t)res * res * texWidth; i++)
        sdf[i] = f2h(1.0f);
    int32_t stats[4] = {0, 0, 0, 0};
    for (int c = 0; c < casc; c++)
    {
        const float D = data->cascadePosDistance[c][3], cascadeMaxDistance = D * 2.0f, voxel = data->cascadeVoxelSize[c];
        const vec3  center = {data->cascadePosDistance[c][0], data->cascadePosDistance[c][1], data->cascadePosDistance[c][2]};
        const vec3  bmin = {center.x - D, center.y - D, center.z - D}, bmax = {center.x + D, center.y + D, center.z + D};
        std::vector<LuxObjectRasterizeData> objects;
        std::vector<const uint16_t*>        meshData;
        std::vector<int32_t>                meshSizes;
        std::vector<ChunkList>              chunks; // insertion order; lookup by coordinate
        auto findChunk = [&](int x, int y, int z) -> ChunkList& {
            for (auto& ch : chunks)
                if (ch.coord[0] == x && ch.coord[1] == y && ch.coord[2] == z)
                    return ch;
            chunks.push_back(ChunkList{{x, y, z}, 0, {}});
            return chunks.back();
        };
        for (int m = 0; m < meshCount; m++)
        {
            const LuxMeshSDF& ms = meshes[m];
            // BoundingBox::transform (BoundingBox.cpp:10-22)
            const float* t = ms.worldMatrix;
            vec3 amn = {ms.aabbMin[0], ms.aabbMin[1], ms.aabbMin[2]}, amx = {ms.aabbMax[0], ms.aabbMax[1], ms.aabbMax[2]};
            vec3 newCenter = mat4_mul_point(t, mul(add(amx, amn), 0.5f), 1.0f);
            vec3 oldEdge   = mul(sub(amx, amn), 0.5f);
            vec3 newEdge   = {std::fabs(t[0]) * oldEdge.x + std::fabs(t[4]) * oldEdge.y + std::fabs(t[8] * oldEdge.z),
                              std::fabs(t[1]) * oldEdge.x + std::fabs(t[5]) * oldEdge.y + std::fabs(t[9] * oldEdge.z),
                              std::fabs(t[2]) * oldEdge.x + std::fabs(t[6]) * oldEdge.y + std::fabs(t[10] * oldEdge.z)};
            vec3 omn = sub(newCenter, newEdge), omx = add(newCenter, newEdge);
            // BoundingSphere(box) + intersectsWithSphere (BoundingSphere.cpp:9-13, BoundingBox.cpp:31-36)
            vec3  sc = mul(add(omx, omn), 0.5f);
            float radius = length3(sub(omx, omn)) / 2.0f;
            vec3  cl = {gclamp(sc.x, bmin.x, bmax.x), gclamp(sc.y, bmin.y, bmax.y), gclamp(sc.z, bmin.z, bmax.z)};
            vec3  dv = sub(sc, cl);
            if (!(dot3(dv, dv) <= radius * radius && radius >= minObjectRadius))
                continue;
            // getChunkId (:193-210)
            const float objectMargin = voxel * (float)LUX_SDF_RASTERIZE_CHUNK_MARGIN;
            vec3 biasMin = {bmin.x + 0.1f, bmin.y + 0.1f, bmin.z + 0.1f};
            vec3 lo = {(omn.x - objectMargin) - biasMin.x, (omn.y - objectMargin) - biasMin.y, (omn.z - objectMargin) - biasMin.z};
            vec3 hi = {(omx.x + objectMargin) - biasMin.x, (omx.y + objectMargin) - biasMin.y, (omx.z + objectMargin) - biasMin.z};
            const float chunkSize = voxel * (float)LUX_SDF_RASTERIZE_CHUNK_SIZE;
            int cmin[3] = {(int)(lo.x / chunkSize), (int)(lo.y / chunkSize), (int)(lo.z / chunkSize)};
            int cmax[3] = {(int)(hi.x / chunkSize), (int)(hi.y / chunkSize), (int)(hi.z / chunkSize)};
            uint32_t objectIndex = (uint32_t)objects.size();
            objects.emplace_back();
            oracle_sdf_object_data(&ms, c, &objects.back());
            int mipLevel = c < 2 ? c : 2;
            if (mipLevel >= ms.mipCount)
                return -2;
            meshData.push_back((const uint16_t*)ms.mips[mipLevel]);
            for (int i = 0; i < 3; i++)
            {
                uint32_t sz = ms.size[i] >> mipLevel;
                meshSizes.push_back((int32_t)(sz ? sz : 1));
            }
            for (int z = cmin[2]; z <= cmax[2]; z++)
                for (int y = cmin[1]; y <= cmax[1]; y++)
                    for (int x = cmin[0]; x <= cmax[0]; x++)
                    {
                        ChunkList& ch = findChunk(x, y, z);
                        if (ch.count == LUX_SDF_RASTERIZE_MODEL_MAX_COUNT)
                        { // `chunk = chunksCache[nextLayerKey]` copies the empty next-layer entry over this one
                            stats[2] += ch.count;
                            ch.count = 0;
                        }
                        ch.models[ch.count++] = objectIndex;
                    }
        }
        const float cmul[3] = {(bmax.x - bmin.x) / (float)res, (bmax.y - bmin.y) / (float)res, (bmax.z - bmin.z) / (float)res};
        const float cadd[3] = {bmin.x + voxel * 0.5f, bmin.y + voxel * 0.5f, bmin.z + voxel * 0.5f};
        for (const ChunkList& ch : chunks)
        {
            if (ch.coord[0] < 0 || ch.coord[1] < 0 || ch.coord[2] < 0 || ch.coord[0] >= rasterizeChunks || ch.coord[1] >= rasterizeChunks ||
                ch.coord[2] >= rasterizeChunks)
            {
                stats[3]++;
                continue;
            }
            int32_t cc[3] = {ch.coord[0] * LUX_SDF_RASTERIZE_CHUNK_SIZE, ch.coord[1] * LUX_SDF_RASTERIZE_CHUNK_SIZE, ch.coord[2] * LUX_SDF_RASTERIZE_CHUNK_SIZE};
            oracle_sdf_rasterize_chunk(cmul, cadd, cascadeMaxDistance, res, c, cc, ch.count, ch.models, 0, objects.data(), meshData.data(),
                                       meshSizes.data(), sdf, texWidth);
            stats[0]++;
            stats[1] += ch.count;
        }
    }
    oracle_sdf_build_mip(data, sdf, mip);
    if (chunkStats)
        for (int i = 0; i < 4; i++)
            chunkStats[i] = stats[i];
    return 0;
}

} // extern "C"

extern "C" {
// =====================================================================================================================
// Surface-cache culling ("next" row f4, first half): Shaders/SDF/SDFCulling.comp:36-101 + boxIntersectsSphere / flattenId
// (AtlasCommon.glsl:34-47), host GlobalSurfaceAtlas.cpp:607-641 (counter reset to 1).  One invocation per culling chunk.
// The shader allocates list space with an atomic, so WHERE a chunk's list lands depends on execution order; `order` fixes one
// (an array of chunk addresses executed one after the other, null = ascending addresses).  emulateSlot0 keeps the shader's
// `atlasChunks.data[0] = chunkAddress` store of empty / overflowing chunks (a data race in the reference: the last writer wins);
// without it element 0 only ever holds chunk 0's own list start, which is what the engine builds.
// =====================================================================================================================
int oracle_surface_cull(const LuxGlobalSurfaceAtlasData* data, const LuxObjectBuffer* objects, const int32_t* order, int orderCount,
                        int emulateSlot0, uint32_t* chunks, uint32_t* cull, uint32_t cullCapacityWords)
{
    if (!data || !objects || !chunks || !cull || cullCapacityWords < 1)
        return -1;
    const int N = LUX_SURFACE_ATLAS_CHUNKS_RESOLUTION, total = N * N * N;
    cull[0] = 1; // GlobalSurfaceAtlas.cpp:617-618
    const int count = order ? orderCount : total;
    for (int k = 0; k < count; k++)
    {
        const uint32_t chunkAddress = order ? (uint32_t)order[k] : (uint32_t)k;
        const int cx = (int)(chunkAddress % N), cy = (int)((chunkAddress / N) % N), cz = (int)(chunkAddress / (N * N));
        const float half = (float)N * 0.5f;
        vec3 chunkMin = {((float)cx - half) * data->chunkSize, ((float)cy - half) * data->chunkSize, ((float)cz - half) * data->chunkSize};
        vec3 chunkMax = {chunkMin.x + data->chunkSize, chunkMin.y + data->chunkSize, chunkMin.z + data->chunkSize};
        auto hits = [&](uint32_t i) {
            const float* b = objects[i].objectBounds;
            vec3 c = {b[0], b[1], b[2]};
            vec3 cl = {gclamp(c.x, chunkMin.x, chunkMax.x), gclamp(c.y, chunkMin.y, chunkMax.y), gclamp(c.z, chunkMin.z, chunkMax.z)};
            return length3(sub(c, cl)) <= b[3];
        };
        uint32_t objectsCount = 0;
        for (uint32_t i = 0; i < data->objectsCount; i++)
            if (hits(i))
                objectsCount++;
        if (objectsCount == 0)
        {
            if (emulateSlot0)
                chunks[0] = chunkAddress;
            continue;
        }
        const uint32_t objectsSize = objectsCount + 1;
        uint32_t objectsStart = cull[0];
        cull[0] += objectsSize; // atomicAdd
        if (objectsStart + objectsSize > data->culledObjectsCapacity)
        {
            if (emulateSlot0)
                chunks[0] = chunkAddress;
            continue;
        }
        if ((size_t)objectsStart + objectsSize > cullCapacityWords)
            return -2;
        cull[objectsStart]   = objectsCount;
        chunks[chunkAddress] = objectsStart;
        for (uint32_t i = 0; i < data->objectsCount; i++) // both shader branches (local array / second scan) list ascending ids
            if (hits(i))
                cull[++objectsStart] = i;
    }
    return 0;
}

// tracyGlobalSDF for arbitrary rays (row f4: the shadow / reflection / surface-cache light rays call the same function with other
// arguments).  needsHitNormal = false leaves the normal at (0,0,0) (SDFCommon.glsl:165-177); minDistance is never read, as in the reference.
} // extern "C"

// threshold = chunkSizeDistance * (1 + 2^-10): three nested fp32 lerps of values in [-1, 1] err by < 1e-6, the margin is >= 3e-5 at res 1024
OpenTable buildOpenTable(const Tex3D& mip, float chunkSizeDistance, int cell = 8)
{
    OpenTable t;
    if (cell < 1 || (cell & (cell - 1)) || mip.w % cell || mip.h % cell || mip.d % cell)
        return t;
    t.cell = cell;
    t.cw = mip.w / cell; t.ch = mip.h / cell; t.cd = mip.d / cell;
    const float threshold = chunkSizeDistance * (1.0f + 0.0009765625f);
    const size_t cells = (size_t)t.cw * t.ch * t.cd;
    t.bits.assign((cells + 31) / 32, 0u);
    t.nearBits.assign((cells + 31) / 32, 0u);
    const float nearThreshold = chunkSizeDistance * (1.0f - 0.0009765625f);
#pragma omp parallel for schedule(static)
    for (long long w = 0; w < (long long)t.bits.size(); w++)
    {
        uint32_t word = 0, nearWord = 0;
        for (int b = 0; b < 32; b++)
        {
            size_t i = (size_t)w * 32 + b;
            if (i >= cells)
                break;
            int cx = (int)(i % t.cw), cy = (int)((i / t.cw) % t.ch), cz = (int)(i / ((size_t)t.cw * t.ch));
            bool nearAll = true;
            for (int z = std::max(cell * cz - 1, 0); z <= std::min(cell * cz + cell, mip.d - 1) && nearAll; z++)
                for (int y = std::max(cell * cy - 1, 0); y <= std::min(cell * cy + cell, mip.h - 1) && nearAll; y++)
                    for (int x = std::max(cell * cx - 1, 0); x <= std::min(cell * cx + cell, mip.w - 1); x++)
                        if (!(mip.texel(x, y, z) < nearThreshold))
                        {
                            nearAll = false;
                            break;
                        }
            nearWord |= (nearAll ? 1u : 0u) << b;
            bool open = true;
            for (int z = std::max(cell * cz - 1, 0); z <= std::min(cell * cz + cell, mip.d - 1) && open; z++)
                for (int y = std::max(cell * cy - 1, 0); y <= std::min(cell * cy + cell, mip.h - 1) && open; y++)
                    for (int x = std::max(cell * cx - 1, 0); x <= std::min(cell * cx + cell, mip.w - 1); x++)
                        if (!(mip.texel(x, y, z) >= threshold))
                        {
                            open = false;
                            break;
                        }
            word |= (open ? 1u : 0u) << b;
        }
        t.bits[w] = word;
        t.nearBits[w] = nearWord;
    }
    return t;
}

extern "C" {
// tracyGlobalSDF with the experimental march's table-driven control flow (Scene::openEmulate): hits must equal oracle_trace_global_sdf's bit for bit;
// tapsOut (optional) = {mip taps, full-resolution taps} actually taken.
int oracle_trace_global_sdf_open_skip(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, int count, const LuxGlobalSDFTrace* traces,
                                      float cascadeTraceStartBias, LuxGlobalSDFHit* hits, uint64_t* tapsOut, int cell)
{
    if (!sdfData || !sdf || !mip || count < 0 || (count > 0 && (!traces || !hits)))
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
    OpenTable table = buildOpenTable(sc.mip, (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE / sdfData->resolution, cell);
    if (table.bits.empty())
        return -2;
    sc.open = &table;
    sc.openEmulate = true;
    uint64_t mipTaps = 0, texTaps = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : mipTaps, texTaps)
    for (int k = 0; k < count; k++)
    {
        const LuxGlobalSDFTrace& t = traces[k];
        Counters cn;
        Hit h = tracyGlobalSDF(sc, {t.worldPosition[0], t.worldPosition[1], t.worldPosition[2]}, {t.worldDirection[0], t.worldDirection[1], t.worldDirection[2]},
                               t.maxDistance, t.stepScale, cascadeTraceStartBias, cn);
        LuxGlobalSDFHit& o = hits[k];
        bool wantN = t.needsHitNormal != 0 && h.hitTime >= 0.0f;
        o.hitNormal[0] = wantN ? h.hitNormal.x : 0.0f; o.hitNormal[1] = wantN ? h.hitNormal.y : 0.0f; o.hitNormal[2] = wantN ? h.hitNormal.z : 0.0f;
        o.hitTime = h.hitTime; o.hitCascade = h.hitCascade; o.stepsCount = h.stepsCount; o.hitSDF = h.hitSDF;
        mipTaps += cn.mipTaps; texTaps += cn.texTaps;
    }
    if (tapsOut)
    {
        tapsOut[0] = mipTaps; tapsOut[1] = texTaps;
    }
    return 0;
}

// Per-step class trace of every ray (see Counters::classOut), [count][maxSteps] bytes, 0 = the ray has ended: input of the warp-coherence
// estimate in DESIGN §11 (a SIMT warp pays for a tap as soon as ONE of its lanes needs it).
int oracle_step_classes(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, int count, const LuxGlobalSDFTrace* traces,
                        float cascadeTraceStartBias, int maxSteps, uint8_t* out, int cell)
{
    if (!sdfData || !sdf || !mip || count < 0 || (count > 0 && (!traces || !out)) || maxSteps < 1)
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
    OpenTable table = buildOpenTable(sc.mip, (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE / sdfData->resolution, cell);
    if (table.bits.empty())
        return -2;
    sc.open = &table;
    std::memset(out, 0, (size_t)count * maxSteps);
#pragma omp parallel for schedule(dynamic, 64)
    for (int k = 0; k < count; k++)
    {
        const LuxGlobalSDFTrace& t = traces[k];
        Counters cn;
        cn.classOut = out + (size_t)k * maxSteps;
        cn.classCap = maxSteps;
        tracyGlobalSDF(sc, {t.worldPosition[0], t.worldPosition[1], t.worldPosition[2]}, {t.worldDirection[0], t.worldDirection[1], t.worldDirection[2]},
                       t.maxDistance, t.stepScale, cascadeTraceStartBias, cn);
    }
    return 0;
}

// Validation of the open-space table on a ray list: out = {march steps, steps in open cells, violations (open cell but mip tap < chunkSizeDistance,
// must be 0), open cells, cells, steps in "near" cells, near violations, near steps whose full-resolution tap is the one used, steps that use it}.  bitsOut (optional, ceil(cells / 32) words) receives the table for comparison with the engine's.
int oracle_open_space_stats(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, int count, const LuxGlobalSDFTrace* traces,
                            float cascadeTraceStartBias, uint64_t* out, uint32_t* bitsOut, int cell)
{
    if (!sdfData || !sdf || !mip || count < 0 || (count > 0 && !traces) || !out)
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
    OpenTable table = buildOpenTable(sc.mip, (float)LUX_GLOBAL_SDF_RASTERIZE_CHUNK_SIZE / sdfData->resolution, cell);
    if (table.bits.empty())
        return -2;
    sc.open = &table;
    uint64_t steps = 0, openSteps = 0, violations = 0, nearSteps = 0, nearViolations = 0, nearTexUsed = 0, texUsed = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : steps, openSteps, violations, nearSteps, nearViolations, nearTexUsed, texUsed)
    for (int k = 0; k < count; k++)
    {
        const LuxGlobalSDFTrace& t = traces[k];
        Counters cn;
        tracyGlobalSDF(sc, {t.worldPosition[0], t.worldPosition[1], t.worldPosition[2]}, {t.worldDirection[0], t.worldDirection[1], t.worldDirection[2]},
                       t.maxDistance, t.stepScale, cascadeTraceStartBias, cn);
        steps += cn.mipTaps; openSteps += cn.openSteps; violations += cn.openViolations;
        nearSteps += cn.nearSteps; nearViolations += cn.nearViolations; nearTexUsed += cn.nearTexUsed; texUsed += cn.texUsed;
    }
    uint64_t openCells = 0;
    for (uint32_t w : table.bits)
        openCells += (uint64_t)__builtin_popcount(w);
    out[0] = steps; out[1] = openSteps; out[2] = violations; out[3] = openCells; out[4] = (uint64_t)table.cw * table.ch * table.cd;
    out[5] = nearSteps; out[6] = nearViolations; out[7] = nearTexUsed; out[8] = texUsed;
    if (bitsOut)
        std::memcpy(bitsOut, table.bits.data(), table.bits.size() * 4);
    return 0;
}

int oracle_trace_global_sdf(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, int count, const LuxGlobalSDFTrace* traces,
                            float cascadeTraceStartBias, LuxGlobalSDFHit* hits)
{
    if (!sdfData || !sdf || !mip || !traces || !hits)
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
#pragma omp parallel for schedule(dynamic, 64)
    for (int k = 0; k < count; k++)
    {
        const LuxGlobalSDFTrace& t = traces[k];
        Counters cn;
        Hit h = tracyGlobalSDF(sc, {t.worldPosition[0], t.worldPosition[1], t.worldPosition[2]}, {t.worldDirection[0], t.worldDirection[1], t.worldDirection[2]},
                               t.maxDistance, t.stepScale, cascadeTraceStartBias, cn);
        LuxGlobalSDFHit& o = hits[k];
        const bool n = t.needsHitNormal && h.hitTime >= 0.0f;
        o.hitNormal[0] = n ? h.hitNormal.x : 0.0f; o.hitNormal[1] = n ? h.hitNormal.y : 0.0f; o.hitNormal[2] = n ? h.hitNormal.z : 0.0f;
        o.hitTime = h.hitTime; o.hitCascade = h.hitCascade; o.stepsCount = h.stepsCount; o.hitSDF = h.hitSDF;
    }
    return 0;
}

} // extern "C"

// Direct lighting of surface-cache texels ("next" row f4): Shaders/SDF/SDFDeferredLight.frag:44-129 with Raytraced/BRDF.glsl:8-36,65-83
// and Common/Light.glsl:13-29; additive blend into the RGBA16F light cache (GlobalSurfaceAtlas.cpp:950-972).  pow() follows the contract
// (binary64, rounded once); mix() is the literal x*(1-a) + y*a of the shipped binary's FMix.
namespace {
const float LUX_PI_F = 3.14159265358979323846f; // M_PI as Common.glsl defines it, rounded to binary32
inline float ndfGGX(float cosLh, float roughness)
{
    float alpha = roughness * roughness, alphaSq = alpha * alpha;
    float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    return alphaSq / ((LUX_PI_F * denom) * denom);
}
inline float gaSchlickG1(float cosTheta, float k) { return cosTheta / (cosTheta * (1.0f - k) + k); }
inline float gaSchlickGGX(float cosLi, float NdotV, float roughness)
{
    float r = roughness + 1.0f, k = (r * r) / 8.0f;
    return gaSchlickG1(cosLi, k) * gaSchlickG1(NdotV, k);
}
inline vec3 brdf(vec3 albedo, vec3 normal, float roughness, float metallic, vec3 view, vec3 halfV, vec3 lightDir)
{
    const float Fd = 0.04f, EPSILON = 0.00001f;
    vec3  F0 = {Fd * (1.0f - metallic) + albedo.x * metallic, Fd * (1.0f - metallic) + albedo.y * metallic, Fd * (1.0f - metallic) + albedo.z * metallic};
    float cosLi = gmax(0.0f, dot3(normal, lightDir)), cosLh = gmax(0.0f, dot3(normal, halfV)), NdotV = gmax(0.0f, dot3(normal, view));
    float ct = gmax(dot3(halfV, view), 0.0f);
    float p5 = pow_rn(gclamp(1.0f - ct, 0.0f, 1.0f), 5.0f);
    vec3  F  = {F0.x + (1.0f - F0.x) * p5, F0.y + (1.0f - F0.y) * p5, F0.z + (1.0f - F0.z) * p5};
    float D = ndfGGX(cosLh, roughness), G = gaSchlickGGX(cosLi, NdotV, roughness);
    vec3  kd = {(1.0f - F.x) * (1.0f - metallic), (1.0f - F.y) * (1.0f - metallic), (1.0f - F.z) * (1.0f - metallic)};
    float den = gmax(EPSILON, (4.0f * cosLi) * NdotV);
    return {(kd.x * albedo.x) / LUX_PI_F + ((F.x * D) * G) / den, (kd.y * albedo.y) / LUX_PI_F + ((F.y * D) * G) / den, (kd.z * albedo.z) / LUX_PI_F + ((F.z * D) * G) / den};
}
} // namespace

extern "C" int oracle_surface_direct_light(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, const LuxLight* light,
                                           const float* cameraPosBias, uint16_t* lightCache, int count, const uint32_t* texel, const float* P,
                                           const float* N, const float* albedo, const float* metallicRoughness)
{
    if (!sdfData || !sdf || !mip || !light || !cameraPosBias || !lightCache || !texel || !P || !N || !albedo || !metallicRoughness)
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
    const LuxLight L = *light;
    const float shadowBias = cameraPosBias[3];
#pragma omp parallel for schedule(dynamic, 64)
    for (int k = 0; k < count; k++)
    {
        vec3 worldPos = {P[3 * k], P[3 * k + 1], P[3 * k + 2]}, normal = {N[3 * k], N[3 * k + 1], N[3 * k + 2]};
        vec3 alb = {albedo[3 * k], albedo[3 * k + 1], albedo[3 * k + 2]};
        float metallic = metallicRoughness[2 * k], roughness = metallicRoughness[2 * k + 1];
        // fetchLight, SDFDeferredLight.frag:44-81
        vec3  Wi = {0, 0, 0};
        float dist = LUX_GLOBAL_SDF_WORLD_SIZE, atten = 1.0f;
        if (L.type == LUX_LIGHT_DIRECTIONAL)
        {
            Wi = {-L.direction[0], -L.direction[1], -L.direction[2]};
            dist = LUX_GLOBAL_SDF_WORLD_SIZE;
            atten = 1.0f;
        }
        else if (L.type == LUX_LIGHT_POINT)
        {
            vec3 dir = sub({L.position[0], L.position[1], L.position[2]}, worldPos);
            float d = length3(dir);
            Wi = normalize3(dir);
            atten = L.radius / (pow_rn(d, 2.0f) + 1.0f);
            dist = d;
        }
        else if (L.type == LUX_LIGHT_SPOT)
        {
            vec3  Lv = sub({L.position[0], L.position[1], L.position[2]}, worldPos);
            float cutoffAngle = 1.0f - L.angle;
            vec3  lightDir = normalize3(Lv);
            float d = length3(Lv);
            float theta = dot3(lightDir, {L.direction[0], L.direction[1], L.direction[2]});
            float epsilon = cutoffAngle - cutoffAngle * 0.9f;
            atten = (theta - cutoffAngle) / epsilon;
            atten *= L.radius / (pow_rn(d, 2.0f) + 1.0f);
            atten = gclamp(atten, 0.0f, 1.0f);
            Wi = lightDir;
            dist = d;
        }
        float shadowMask = 1.0f;
        float NoL = dot3(normal, Wi);
        float bias = (2.0f * shadowBias) * gclamp(1.0f - NoL, 0.0f, 1.0f) + shadowBias;
        if (NoL > 0.0f)
        {
            if (atten > 0.0f)
            {
                Counters cn;
                vec3 origin = add(worldPos, mul(normal, shadowBias));
                Hit  hit = tracyGlobalSDF(sc, origin, Wi, dist - bias, 1.0f, 2.0f, cn);
                shadowMask = hit.hitTime >= 0.0f ? 0.0f : 1.0f;
            }
        }
        else
            shadowMask = 0.0f;
        vec3  view = normalize3(sub({cameraPosBias[0], cameraPosBias[1], cameraPosBias[2]}, worldPos));
        float intensity = pow_rn(L.intensity, 1.4f) + 0.1f;
        vec3  Lrad = {L.color[0] * intensity, L.color[1] * intensity, L.color[2] * intensity};
        vec3  Lh = normalize3(add(Wi, view));
        float cosLi = gmax(0.0f, dot3(normal, Wi));
        vec3  b = brdf(alb, normal, roughness, metallic, view, Lh, Wi);
        float out[4] = {(((b.x * Lrad.x) * cosLi) * shadowMask) * atten, (((b.y * Lrad.y) * cosLi) * shadowMask) * atten,
                        (((b.z * Lrad.z) * cosLi) * shadowMask) * atten, 1.0f};
        size_t o = (size_t)texel[k] * 4;
        for (int c = 0; c < 4; c++)
            lightCache[o + c] = f2h(out[c] + h2f(lightCache[o + c])); // BlendMode::Add: src + dst
    }
    return 0;
}

extern "C" void oracle_octohedral_to_direction(float ex, float ey, float* out3)
{
    vec3 v = octohedralToDirection({ex, ey});
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

// ---------------------------------------------------------------------------------------------------------
// The screen-space tracyGlobalSDF users ("next" row f4): Shaders/SDF/SDFReflection.comp and Shaders/SDF/SDFShadow.comp with
// Raytraced/BlueNoise.glsl:8-19, Raytraced/BRDF.glsl:176-201 (importanceSampleGGX), Common/Math.glsl:27-42.
// cross / reflect follow the contract's reading of the GLSL.std.450 instructions: every product, sum and difference rounded.
// ---------------------------------------------------------------------------------------------------------
namespace {
inline vec3 cross3(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3 reflect3(vec3 i, vec3 n)
{
    float t = 2.0f * dot3(n, i);
    return {i.x - t * n.x, i.y - t * n.y, i.z - t * n.z};
}
inline float unorm8(uint8_t b) { return (float)b / 255.0f; }
// sampleBlueNoise, BlueNoise.glsl:8-19: sobol = 256 x 1 RGBA8, scramblingRanking = 128 x 128 RGBA8
float sampleBlueNoise(int cx, int cy, int samplerIndex, int dimension, const uint8_t* sobol, const uint8_t* scr)
{
    cx = cx % 128;
    cy = cy % 128;
    samplerIndex = samplerIndex % 256;
    dimension    = dimension % 4;
    const uint8_t* t = scr + ((size_t)cy * 128 + cx) * 4;
    int rankedIndex = samplerIndex ^ (int)gclamp(unorm8(t[2]) * 256.0f, 0.0f, 255.0f);
    int value       = (int)gclamp(unorm8(sobol[(size_t)rankedIndex * 4 + dimension]) * 256.0f, 0.0f, 255.0f);
    value           = value ^ (int)gclamp(unorm8(t[dimension % 2]) * 256.0f, 0.0f, 255.0f);
    return (0.5f + (float)value) / 256.0f;
}
inline vec3 worldPositionFromDepth(float tx, float ty, float d, const float* m) // Common/Math.glsl:35-42
{
    float sx = tx * 2.0f - 1.0f, sy = ty * 2.0f - 1.0f;
    float wx = ((m[0] * sx + m[4] * sy) + m[8] * d) + m[12] * 1.0f;
    float wy = ((m[1] * sx + m[5] * sy) + m[9] * d) + m[13] * 1.0f;
    float wz = ((m[2] * sx + m[6] * sy) + m[10] * d) + m[14] * 1.0f;
    float ww = ((m[3] * sx + m[7] * sy) + m[11] * d) + m[15] * 1.0f;
    return {wx / ww, wy / ww, wz / ww};
}
// importanceSampleGGX(...).xyz, BRDF.glsl:176-201 (the pdf is computed and dropped by the caller)
vec3 importanceSampleGGX(vec2 E, vec3 N, float roughness)
{
    const float TWO_PI = 6.283185482025146484375f; // 2.0f * M_PI folded by glslang
    float a = roughness * roughness, m2 = a * a;
    float phi      = TWO_PI * E.x;
    float cosTheta = std::sqrt((1.0f - E.y) / (1.0f + (m2 - 1.0f) * E.y));
    float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
    vec3  H  = {cos_rn(phi) * sinTheta, sin_rn(phi) * sinTheta, cosTheta};
    vec3  up = std::fabs(N.z) < 0.999f ? vec3{0.0f, 0.0f, 1.0f} : vec3{1.0f, 0.0f, 0.0f};
    vec3  tangent   = normalize3(cross3(up, N));
    vec3  bitangent = cross3(N, tangent);
    vec3  sv = {(tangent.x * H.x + bitangent.x * H.y) + N.x * H.z, (tangent.y * H.x + bitangent.y * H.y) + N.y * H.z,
                (tangent.z * H.x + bitangent.z * H.y) + N.z * H.z};
    return normalize3(sv);
}
void fillScene(Scene& sc, const OracleSceneDesc* s)
{
    if (s->ddgi)
        sc.ddgi = *s->ddgi;
    sc.sdfData = *s->sdfData;
    int res = (int)s->sdfData->resolution, casc = (int)s->sdfData->cascadesCount;
    sc.tex = Tex3D{s->sdf, res * casc, res, res};
    sc.mip = Tex3D{s->mip, (res / 4) * casc, res / 4, res / 4};
    sc.hasAtlas = s->atlasData != nullptr;
    if (sc.hasAtlas)
        sc.atlasData = *s->atlasData;
    sc.chunks = s->chunks; sc.cull = s->cull; sc.objects = s->objects; sc.tiles = s->tiles;
    sc.light = s->light; sc.depth = s->depth;
    sc.skyFace = s->skyFace; sc.sky = s->sky;
}
} // namespace

// SDFReflection.comp:84-163.  gDepth [h][w] D32F, gNormal / gPbr [h][w][4] RGBA32F, out [h][w][4] RGBA16F (pixels with depth == 1 untouched).
extern "C" int oracle_sdf_reflection(const OracleSceneDesc* s, const uint16_t* irr, const uint16_t* depthAtlas, const LuxReflectionPushConstants* push,
                                     int width, int height, const float* gDepth, const float* gNormal, const float* gPbr, const uint8_t* sobol,
                                     const uint8_t* scramblingRanking, uint16_t* out)
{
    if (!s || !s->ddgi || !s->sdfData || !s->sdf || !s->mip || !irr || !depthAtlas || !push || !gDepth || !gNormal || !gPbr || !sobol || !scramblingRanking || !out)
        return -1;
    Scene sc{};
    fillScene(sc, s);
    Atlas2D ai{irr, sc.ddgi.irradianceTextureWidth, sc.ddgi.irradianceTextureHeight, 4};
    Atlas2D ad{depthAtlas, sc.ddgi.depthTextureWidth, sc.ddgi.depthTextureHeight, 2};
    const float MIRROR = 0.05f, DDGI_ROUGH = 0.45f; // Reflection/ReflectionCommon.glsl:7-8
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++)
        {
            size_t o = (size_t)y * width + x;
            float  depth = gDepth[o];
            if (!(depth != 1.0f))
                continue;
            float tx = ((float)x + 0.5f) / (float)width, ty = ((float)y + 0.5f) / (float)height;
            vec3  worldPos  = worldPositionFromDepth(tx, ty, depth, push->viewProjInv);
            float roughness = gPbr[4 * o + 1];
            vec3  normal    = octohedralToDirection({gNormal[4 * o], gNormal[4 * o + 1]});
            vec3  Wo        = normalize3(sub({push->cameraPosition[0], push->cameraPosition[1], push->cameraPosition[2]}, worldPos));
            worldPos        = add(worldPos, mul(normal, sc.sdfData.cascadeVoxelSize[0]));
            vec3 negWo = {-Wo.x, -Wo.y, -Wo.z};
            vec4 color = {0, 0, 0, 0};
            Counters cn;
            auto trace = [&](vec3 R) -> vec4 { // trace(), SDFReflection.comp:86-117
                Hit hit = tracyGlobalSDF(sc, worldPos, R, LUX_GLOBAL_SDF_WORLD_SIZE, 1.0f, 0.0f, cn);
                if (hit.hitTime >= 0.0f)
                {
                    float surfaceThreshold = sc.sdfData.cascadeVoxelSize[hit.hitCascade] * 1.05f;
                    return sampleGlobalSurfaceAtlas(sc, add(worldPos, mul(R, hit.hitTime)), {-R.x, -R.y, -R.z}, surfaceThreshold, cn);
                }
                return sampleSky4(sc, R);
            };
            if (roughness < MIRROR)
                color = trace(reflect3(negWo, normal));
            else if (roughness > DDGI_ROUGH && push->approximateWithDDGI == 1u)
            {
                vec3 R = reflect3(negWo, normal);
                vec3 e = sampleIrradiance(sc.ddgi, worldPos, R, Wo, ai, ad);
                color  = {push->roughDDGIIntensity * e.x, push->roughDDGIIntensity * e.y, push->roughDDGIIntensity * e.z, 0.0f};
            }
            else
            {
                vec2 Xi = {sampleBlueNoise(x, y, (int)push->numFrames, 0, sobol, scramblingRanking) * push->trim,
                           sampleBlueNoise(x, y, (int)push->numFrames, 1, sobol, scramblingRanking) * push->trim};
                vec3 Wh = importanceSampleGGX(Xi, normal, roughness);
                color   = trace(reflect3(negWo, Wh));
            }
            out[4 * o] = f2h(color.x); out[4 * o + 1] = f2h(color.y); out[4 * o + 2] = f2h(color.z); out[4 * o + 3] = f2h(color.w);
        }
    return 0;
}

// SDFShadow.comp:121-157 with fetchLight :42-117 (softShadow = true).  outMask [height/4][width/8] uint32, read-modify-write.
extern "C" int oracle_sdf_shadow(const LuxGlobalSDFData* sdfData, const uint16_t* sdf, const uint16_t* mip, const LuxLight* light, const float* viewProjInv,
                                 uint32_t numFrames, float shadowBias, int width, int height, const float* gDepth, const float* gNormal,
                                 const uint8_t* sobol, const uint8_t* scramblingRanking, uint32_t* outMask)
{
    if (!sdfData || !sdf || !mip || !light || !viewProjInv || !gDepth || !gNormal || !sobol || !scramblingRanking || !outMask || width % 8 || height % 4)
        return -1;
    Scene sc{};
    sc.sdfData = *sdfData;
    const int res = (int)sdfData->resolution, casc = (int)sdfData->cascadesCount;
    sc.tex = Tex3D{sdf, res * casc, res, res};
    sc.mip = Tex3D{mip, (res / 4) * casc, res / 4, res / 4};
    const LuxLight L = *light;
    const float PI_F = 3.1415926535897932384626433832795f; // Common/Math.glsl:6
    const int gw = width / 8, gh = height / 4;
#pragma omp parallel for schedule(dynamic, 1)
    for (int g = 0; g < gw * gh; g++)
    {
        const int gx = g % gw, gy = g / gw;
        uint32_t visibility = 0;
        bool     stored = false;
        for (int li = 0; li < 32; li++)
        {
            const int x = gx * 8 + (li % 8), y = gy * 4 + (li / 8);
            size_t o = (size_t)y * width + x;
            float  depth = gDepth[o];
            if (!(depth != 1.0f))
                continue;
            if (li == 0)
                stored = true;
            float tx = ((float)x + 0.5f) / (float)width, ty = ((float)y + 0.5f) / (float)height;
            vec3  worldPos = worldPositionFromDepth(tx, ty, depth, viewProjInv);
            vec3  normal   = octohedralToDirection({gNormal[4 * o], gNormal[4 * o + 1]});
            vec2  rnd = {sampleBlueNoise(x, y, (int)numFrames, 0, sobol, scramblingRanking), sampleBlueNoise(x, y, (int)numFrames, 1, sobol, scramblingRanking)};
            // fetchLight(light, worldPos, normal, rnd, Wi, tMax, attenuation, true)
            vec3  lightDir = {0, 0, 0}, Wi = {0, 0, 0};
            float lightRadius = 0.0f, tMax = 0.0f, attenuation = 0.0f;
            if (L.type == LUX_LIGHT_DIRECTIONAL)
            {
                lightDir = {-L.direction[0], -L.direction[1], -L.direction[2]};
                tMax = LUX_GLOBAL_SDF_WORLD_SIZE;
                Wi = lightDir;
                lightRadius = L.direction[3];
                attenuation = 1.0f;
            }
            else if (L.type == LUX_LIGHT_POINT)
            {
                vec3  dir  = sub({L.position[0], L.position[1], L.position[2]}, worldPos);
                float dist = length3(dir);
                lightDir = normalize3(dir);
                attenuation = 1.0f;
                Wi = lightDir;
                tMax = dist;
                lightRadius = L.direction[3] / dist;
            }
            else if (L.type == LUX_LIGHT_SPOT)
            {
                vec3  Lv = sub({L.position[0], L.position[1], L.position[2]}, worldPos);
                float cutoffAngle = 1.0f - L.angle;
                lightDir = normalize3(Lv);
                float dist    = length3(Lv);
                float theta   = dot3(lightDir, {L.direction[0], L.direction[1], L.direction[2]});
                float epsilon = cutoffAngle - cutoffAngle * 0.9f;
                attenuation = (theta - cutoffAngle) / epsilon;
                attenuation *= L.radius / (pow_rn(dist, 2.0f) + 1.0f);
                attenuation = gclamp(attenuation, 0.0f, 1.0f);
                Wi = lightDir;
                tMax = dist;
                lightRadius = L.direction[3] / dist;
            }
            {
                vec3  lightTangent   = normalize3(cross3(lightDir, {0.0f, 1.0f, 0.0f}));
                vec3  lightBitangent = normalize3(cross3(lightTangent, lightDir));
                float pointRadius = lightRadius * std::sqrt(rnd.x);
                float pointAngle  = (rnd.y * 2.0f) * PI_F;
                float dx = pointRadius * cos_rn(pointAngle), dy = pointRadius * sin_rn(pointAngle);
                vec3  w = {(lightDir.x + dx * lightTangent.x) + dy * lightBitangent.x, (lightDir.y + dx * lightTangent.y) + dy * lightBitangent.y,
                           (lightDir.z + dx * lightTangent.z) + dy * lightBitangent.z};
                Wi = normalize3(w);
            }
            attenuation *= gclamp(dot3(normal, Wi), 0.0f, 1.0f);
            float    NoL = dot3(normal, Wi);
            uint32_t result = 0;
            if (NoL > 0.0f)
            {
                if (attenuation > 0.0f)
                {
                    float bias = (2.0f * shadowBias) * gclamp(1.0f - NoL, 0.0f, 1.0f) + shadowBias;
                    vec3  rayOrigin = add(worldPos, mul(normal, shadowBias));
                    Counters cn;
                    Hit hit = tracyGlobalSDF(sc, rayOrigin, Wi, tMax - bias, 1.0f, 0.0f, cn);
                    result  = hit.hitTime >= 0.0f ? 0u : 1u;
                }
            }
            visibility |= result << li;
        }
        if (stored)
            outMask[(size_t)gy * gw + gx] = visibility;
    }
    return 0;
}

extern "C" {
// Border copy list of one probe in ring-relative coordinates: out[n][4] = (srcx, srcy, dstx, dsty); returns n.
int oracle_border_offsets(int side, int32_t* out)
{
    int n = 0;
    auto push = [&](int sx, int sy, int dx, int dy) { out[n * 4] = sx; out[n * 4 + 1] = sy; out[n * 4 + 2] = dx; out[n * 4 + 3] = dy; n++; };
    for (int x = 1; x <= side; x++) push(side + 1 - x, 1, x, 0);
    for (int x = 1; x <= side; x++) push(side + 1 - x, side, x, side + 1);
    for (int y = 1; y <= side; y++) push(1, side + 1 - y, 0, y);
    for (int y = 1; y <= side; y++) push(side, side + 1 - y, side + 1, y);
    push(side, side, 0, 0); push(1, side, side + 1, 0); push(side, 1, 0, side + 1); push(1, 1, side + 1, side + 1);
    return n;
}

} // extern "C"
