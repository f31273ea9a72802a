// ddgi_math.cuh — device-side arithmetic of the DDGI engine under the numerics contract.
//
// The translation unit is compiled with -fmad=false: every +,-,*,/ and sqrtf below is a single IEEE-754 binary32
// round-to-nearest operation in source order.  A fused multiply-add happens only where __fmaf_rn is written.
// GLSL min/max/clamp are select-based (NaN-transparent in the first operand), matching SPIR-V's latitude for
// FMin/FMax.  sin/cos/pow go through binary64 and are rounded once.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lux {

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

__device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
__device__ __forceinline__ int   iclamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
__device__ __forceinline__ float gfract(float x) { return x - floorf(x); }

__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float dot4(f4 a, f4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
__device__ __forceinline__ f3 normalize3(f3 v)
{
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(v, v)));
    return {v.x * inv, v.y * inv, v.z * inv};
}
__device__ __forceinline__ float length3(f3 v) { return __fsqrt_rn(dot3(v, v)); }

__device__ __forceinline__ float sin_rn(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos_rn(float x) { return (float)cos((double)x); }
// pow(x, y) for x >= 0 rounded once from binary64: exp(y*log(x)) carries ~1e-15 relative error, far below half
// an ulp of binary32, so the result equals the correctly rounded pow except on ~1e-8 of inputs.
__device__ __forceinline__ float pow_rn(float x, float y) { return (float)exp((double)y * log((double)x)); }

// column-major mat4 (m[c*4+r]) times (v, w)
__device__ __forceinline__ f3 mat4_mul_point(const float* __restrict__ m, f3 v, float w)
{
    f3 r;
    r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * w;
    r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * w;
    r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * w;
    return r;
}
__device__ __forceinline__ f3 mat3_mul(const float* __restrict__ m, f3 v)
{
    f3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + m[8] * v.z;
    r.y = (m[1] * v.x + m[5] * v.y) + m[9] * v.z;
    r.z = (m[2] * v.x + m[6] * v.y) + m[10] * v.z;
    return r;
}

__device__ __forceinline__ float h2f_bits(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }

// x / d for a divisor that is reused many times.  When d is a normal power of two, x * (1/d) is the correctly rounded
// quotient too (pure exponent shift), so the IEEE division (~17 SASS instructions) can be replaced by one multiply
// without changing a single bit; any other divisor keeps the true division.
struct ExactDivisor
{
    float d, inv;
    bool  pow2;
    __device__ __forceinline__ ExactDivisor() : d(1.0f), inv(1.0f), pow2(true) {}
    __device__ __forceinline__ explicit ExactDivisor(float v) : d(v)
    {
        uint32_t b = __float_as_uint(v);
        uint32_t e = (b >> 23) & 0xffu;
        pow2 = ((b & 0x7fffffu) == 0u) && e > 2u && e < 252u; // normal power of two whose reciprocal is normal as well
        inv  = pow2 ? __fdiv_rn(1.0f, v) : 0.0f;
    }
    __device__ __forceinline__ float div(float x) const { return pow2 ? x * inv : __fdiv_rn(x, d); }
};

// mix(x, y, a) in the contract's form: fma(y, a, x*(1-a))
__device__ __forceinline__ float mixh(float x, float y, float a) { return __fmaf_rn(y, a, x * (1.0f - a)); }

} // namespace lux
