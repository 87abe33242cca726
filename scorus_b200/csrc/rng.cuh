// rng.cuh -- counter-based random streams of the device sampler ("stream spec v1").
//
// The reference consumes one sequential rand::Rng on the calling thread
// (src/mcmc/ensemble_sample.rs:139,149,152,185; src/mcmc/utils.rs:97,104).  Here every draw
// is a pure function of (seed, step, walker, stream id), so any walker can be processed by
// any thread / GPU and a run is reproducible from (seed, counters).  DESIGN.md section 5
// states the spec; oracle/philox_streams.c restates it independently on the CPU for replay.
#pragma once
#include <cstdint>

namespace scorus {

enum : uint32_t {
    ST_ZU = 1,         // idx = walker, sub = 0: (r0,r1) -> z uniform, (r2,r3) -> accept uniform
    ST_FLAGS = 2,      // idx = walker, sub = attempt*nblk + blk: Bernoulli flag bits
    ST_PERM = 3,       // idx = rung, sub = 0,1: keys of the pairing bijection (npb > 1024)
    ST_SWAP_PERM = 4,  // idx = hot rung, sub = 0,1: keys of the jvec bijection (npb > 1024)
    ST_SWAP_U = 5,     // idx = cold slot j2, sub = hot rung: swap uniform
    ST_PERM_KEY = 6,   // idx = walker: 64-bit sort key of the exact shuffle (npb <= 1024)
    ST_SWAP_KEY = 7,   // idx = hot walker slot: 64-bit sort key of the exact jvec (npb <= 1024)
    ST_INIT = 8        // idx = walker, sub = d/2: box-uniform start of coordinates d, d+1 (step = 0)
};

constexpr uint32_t kExactShuffleMax = 1024;  // rungs up to this size get an exact uniform shuffle

struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Philox4x32-10 (Salmon et al. 2011)
__host__ __device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__host__ __device__ __forceinline__ U4 draw(uint64_t seed, uint64_t step, uint32_t stream,
                                            uint32_t idx, uint32_t sub)
{
    U4 c;
    c.x = idx;
    c.y = sub;
    c.z = (uint32_t)step;
    c.w = ((uint32_t)(step >> 32) & 0x00FFFFFFu) | (stream << 24);
    return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 53-bit uniform in [0,1)
__host__ __device__ __forceinline__ double u53(uint32_t lo, uint32_t hi)
{
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * 0x1.0p-53;
}

struct BijKeys { uint32_t k[8]; };

// sub0: first of the two draws the keys come from -- 0 for a whole-rung matching, 2c for matching block c of the
// rung (blocked matchings)
__host__ __device__ __forceinline__ BijKeys bij_keys(uint64_t seed, uint64_t step, uint32_t stream,
                                                     uint32_t idx, uint32_t sub0 = 0)
{
    U4 a = draw(seed, step, stream, idx, sub0), b = draw(seed, step, stream, idx, sub0 + 1u);
    BijKeys r;
    r.k[0] = a.x; r.k[1] = a.y; r.k[2] = a.z; r.k[3] = a.w;
    r.k[4] = b.x; r.k[5] = b.y; r.k[6] = b.z; r.k[7] = b.w;
    return r;
}

__host__ __device__ __forceinline__ uint32_t bij_bits(uint32_t n)
{
    uint32_t b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    return b;
}

// keyed bijection of [0, n): four add / odd-multiply / xorshift rounds on b = ceil(log2 n)
// bits (each a bijection of the b-bit integers), cycle-walked back into range.
__host__ __device__ __forceinline__ uint32_t bij(uint32_t x, uint32_t n, uint32_t b, const BijKeys &K)
{
    const uint32_t C0 = 0x9E3779B1u, C1 = 0x85EBCA6Bu, C2 = 0xC2B2AE35u, C3 = 0x27D4EB2Fu;
    const uint32_t mask = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    uint32_t s1 = (b + 1) / 2, s2 = b / 3;
    if (s1 < 1) s1 = 1;
    if (s2 < 1) s2 = 1;
    do {
        x = (x + K.k[0]) & mask; x = (x * (K.k[1] | 1u)) & mask; x ^= x >> s1; x = (x * C0) & mask; x ^= x >> s2;
        x = (x + K.k[2]) & mask; x = (x * (K.k[3] | 1u)) & mask; x ^= x >> s1; x = (x * C1) & mask; x ^= x >> s2;
        x = (x + K.k[4]) & mask; x = (x * (K.k[5] | 1u)) & mask; x ^= x >> s1; x = (x * C2) & mask; x ^= x >> s2;
        x = (x + K.k[6]) & mask; x = (x * (K.k[7] | 1u)) & mask; x ^= x >> s1; x = (x * C3) & mask; x ^= x >> s2;
    } while (x >= n);
    return x;
}

// ---- the inverse bijection: position of a walker in the matching order ----
// Every round of bij is invertible on b-bit integers (add, multiply by an odd number, xorshift), and the inverse of a
// cycle-walked permutation is the cycle-walked inverse, so a rank of a sharded run finds the matching position -- and
// with it the partner -- of each of ITS walkers in O(1) instead of scanning the whole matching (api_shard.cu).
__host__ __device__ constexpr uint32_t inv_odd32(uint32_t a)
{
    uint32_t x = a;  // a*a = 1 mod 8: three correct bits, doubled by every Newton step
    for (int i = 0; i < 4; ++i) x *= 2u - a * x;
    return x;
}
__host__ __device__ __forceinline__ uint32_t unxorshift(uint32_t y, uint32_t s, uint32_t b)
{
    uint32_t x = y;
    for (uint32_t t = s; t < b; t += s) x ^= y >> t;
    return x;
}
__host__ __device__ __forceinline__ uint32_t bij_inv(uint32_t y, uint32_t n, uint32_t b, const BijKeys &K)
{
    constexpr uint32_t I0 = inv_odd32(0x9E3779B1u), I1 = inv_odd32(0x85EBCA6Bu), I2 = inv_odd32(0xC2B2AE35u),
                       I3 = inv_odd32(0x27D4EB2Fu);
    const uint32_t mask = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    uint32_t s1 = (b + 1) / 2, s2 = b / 3;
    if (s1 < 1) s1 = 1;
    if (s2 < 1) s2 = 1;
    const uint32_t m0 = inv_odd32(K.k[1] | 1u), m1 = inv_odd32(K.k[3] | 1u), m2 = inv_odd32(K.k[5] | 1u),
                   m3 = inv_odd32(K.k[7] | 1u);
    do {
        y = unxorshift(y, s2, b); y = (y * I3) & mask; y = unxorshift(y, s1, b); y = (y * m3) & mask; y = (y - K.k[6]) & mask;
        y = unxorshift(y, s2, b); y = (y * I2) & mask; y = unxorshift(y, s1, b); y = (y * m2) & mask; y = (y - K.k[4]) & mask;
        y = unxorshift(y, s2, b); y = (y * I1) & mask; y = unxorshift(y, s1, b); y = (y * m1) & mask; y = (y - K.k[2]) & mask;
        y = unxorshift(y, s2, b); y = (y * I0) & mask; y = unxorshift(y, s1, b); y = (y * m0) & mask; y = (y - K.k[0]) & mask;
    } while (y >= n);
    return y;
}

// ---- the pieces that depend on the scalar type T of the walkers (the reference is generic over T: Float,
// src/mcmc/ensemble_sample.rs:95; the f32 instantiation of the kernels is produced from the same sources, gen_f32.py) ----
// unit uniform in [0,1) at the type's full mantissa: 53 bits of (hi, lo) for f64, the top 24 bits of hi for f32
template <typename T> __host__ __device__ __forceinline__ T to_unit(uint32_t lo, uint32_t hi);
template <> __host__ __device__ __forceinline__ double to_unit<double>(uint32_t lo, uint32_t hi) { return u53(lo, hi); }
template <> __host__ __device__ __forceinline__ float to_unit<float>(uint32_t, uint32_t hi) { return (float)(hi >> 8) * 0x1.0p-24f; }
template <typename T> __host__ __device__ __forceinline__ T real_nan();
template <> __host__ __device__ __forceinline__ double real_nan<double>()
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff8000000000000ll);
#else
    return __builtin_nan("");
#endif
}
template <> __host__ __device__ __forceinline__ float real_nan<float>()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7fc00000);
#else
    return __builtin_nanf("");
#endif
}
// guard band of the exp-free swap decision (swap_kernel.cuh): far above the rounding of exp / log in T, far below
// anything that matters
template <typename T> __host__ __device__ __forceinline__ T swap_guard();
template <> __host__ __device__ __forceinline__ double swap_guard<double>() { return 1e-5; }
template <> __host__ __device__ __forceinline__ float swap_guard<float>() { return 1e-4f; }
// ... and what is added to the band per unit of the magnitudes that enter the threshold, against the rounding of the
// threshold itself (thousands of ulps of T)
template <typename T> __host__ __device__ __forceinline__ T swap_round_guard();
template <> __host__ __device__ __forceinline__ double swap_round_guard<double>() { return 1e-12; }
template <> __host__ __device__ __forceinline__ float swap_round_guard<float>() { return 1e-5f; }

// Goodman-Weare stretch factor from a unit uniform: src/mcmc/utils.rs:39-44 with
// p = u * 2(sqrt a - 1/sqrt a);  inv_sqrt_a and range are computed once on the host.
template <typename T>
__host__ __device__ __forceinline__ T z_from_uniform(T u01, T inv_sqrt_a, T range)
{
    T p = u01 * range;
    T y = inv_sqrt_a + p / T(2);
    return y * y;
}

#ifdef __CUDACC__
// fused multiply-add in T, whatever -fmad says (the kernels are built with -fmad=false: the reference rounds every
// multiply and add; polynomial evaluations inside a transcendental are not the reference's arithmetic)
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// streaming loads: read once, do not keep in L1 (the shared-memory carve-out leaves little).  One 16-byte (f64) or
// 8-byte (f32) chunk = two coordinates.
__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
// scalar read of a peer's memory: not through the non-coherent path (the value was written by another GPU)
__device__ __forceinline__ double ld_peer(const double *p)
{
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_peer(const float *p)
{
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
#endif

}  // namespace scorus
