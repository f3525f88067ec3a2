// Device self-test of the invariant-divisor division used by the fused quantiser / dequantiser
// (common.cuh: div_invariant) against IEEE division (__fdiv_rn), over a dense operand set.
#include "common.cuh"

namespace beast {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Divisor d: a float in [2^-27, 2^100) (the quantiser's divisor is clamped to >= 1e-8 > 2^-27);
// a few special significands (1.0, all ones, 1+ulp) are forced.
__device__ float selftest_divisor(int d, uint64_t seed) {
    const uint64_t h = splitmix64(seed ^ (uint64_t)d * 0x100000001B3ull);
    uint32_t mant = (uint32_t)(h & 0x7FFFFFu);
    const int sel = d % 16;
    if (sel == 0) mant = 0;
    else if (sel == 1) mant = 0x7FFFFFu;
    else if (sel == 2) mant = 1;
    else if (sel == 3) mant = 0x7FFFFEu;
    else if (sel == 4) mant = 0x400000u;
    const int e = (int)((h >> 32) % 127u) - 27 + 127;           // exponent 2^-27 .. 2^99
    return __uint_as_float(((uint32_t)e << 23) | mant);
}

// Quantiser form: numerators = every float in the two binades below b, b itself, 0, and random
// floats in [0, b] over all exponents.
__global__ void selftest_div_kernel(int n_div, uint64_t seed, unsigned long long* mism) {
    const int d = blockIdx.y;
    if (d >= n_div) return;
    const float b = selftest_divisor(d, seed);
    const float y = __frcp_rn(b);
    const uint32_t bb = __float_as_uint(b);
    unsigned long long bad = 0;
    const uint32_t total = (1u << 24) + (1u << 22);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        float a;
        if (j < (1u << 24)) {
            a = __uint_as_float(bb - j);                        // walks down from b through two binades
        } else {
            const uint64_t h = splitmix64(seed + ((uint64_t)d << 32) + j);
            const uint32_t bits = (uint32_t)(h % (uint64_t)(bb + 1u));   // any non-negative float <= b (incl. 0, subnormals)
            a = __uint_as_float(bits);
        }
        const float want = __fdiv_rn(a, b);
        const float got = div_invariant(a, b, y);
        // Bit-equal whenever the quotient is >= 2^-60 (then a >= 2^-87 and every residual is a normal
        // number).  A smaller quotient can never reach bin 1 ((V-1) * 2^-60 < 0.5 for V <= 2^31), so
        // there only "non-negative and < 2^-60" is required.
        const bool ok = (__float_as_uint(want) == __float_as_uint(got)) ||
                        (want < 8.6736174e-19f && got < 8.6736174e-19f && got >= 0.0f);
        if (!ok) {
            ++bad;
            const unsigned long long slot = atomicAdd(&mism[2], 1ull);
            if (slot < 8) {                                     // keep a few examples: (a, b, want, got) bit patterns
                mism[4 + 2 * slot] = ((unsigned long long)__float_as_uint(a) << 32) | bb;
                mism[5 + 2 * slot] = ((unsigned long long)__float_as_uint(want) << 32) | __float_as_uint(got);
            }
        }
    }
    if (bad) atomicAdd(&mism[0], bad);
}

// Dequantiser form: float(tok) / (V-1) for every V in [2, vmax] and every tok in [-2, V+1],
// plus large random int64 tokens.
__global__ void selftest_deq_kernel(int vmax, uint64_t seed, unsigned long long* mism) {
    unsigned long long bad = 0;
    for (int V = 2 + blockIdx.x; V <= vmax; V += gridDim.x) {
        const float vm1 = (float)(V - 1);
        const float y = __frcp_rn(vm1);
        for (long long t = -2 + threadIdx.x; t <= V + 1; t += blockDim.x) {
            const float a = __ll2float_rn(t);
            bad += (__float_as_uint(__fdiv_rn(a, vm1)) == __float_as_uint(div_invariant(a, vm1, y))) ? 0 : 1;
        }
        for (int r = threadIdx.x; r < 64; r += blockDim.x) {
            const long long t = (long long)splitmix64(seed + (uint64_t)V * 64 + r);
            const float a = __ll2float_rn(t);
            bad += (__float_as_uint(__fdiv_rn(a, vm1)) == __float_as_uint(div_invariant(a, vm1, y))) ? 0 : 1;
        }
    }
    if (bad) atomicAdd(&mism[1], bad);
}

}  // namespace beast

using namespace beast;

extern "C" int beast_selftest_div(int32_t n_divisors, int32_t vmax, uint64_t seed, unsigned long long* mismatches,
                                  void* stream) {
    if (!mismatches) return BEAST_E_NULL;
    if (n_divisors < 0 || n_divisors > 65535 || vmax < 2) return BEAST_E_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(mismatches, 0, 20 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    if (n_divisors > 0) {
        selftest_div_kernel<<<dim3(296, n_divisors), 256, 0, st>>>(n_divisors, seed, mismatches);
        count_launch();
        BEAST_CHECK_LAUNCH();
    }
    selftest_deq_kernel<<<592, 256, 0, st>>>(vmax, seed, mismatches);
    count_launch();
    BEAST_CHECK_LAUNCH();
    return BEAST_OK;
}
