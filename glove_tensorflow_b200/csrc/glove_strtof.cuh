// Decimal text -> float32, correctly rounded (round-to-nearest-even), for the CSV ingest kernels.
//
// The reference reads interaction.csv with tf.data make_csv_dataset [ref src/models/data_utils.py:4-26]; its float
// columns are inferred as tf.float32 and DecodeCSV converts the field text straight to float (one rounding, not
// text -> double -> float).  This header restates that contract with the Eisel-Lemire algorithm (D. Lemire, "Number
// parsing at a gigabyte per second", SPE 2021; with a 128-bit power-of-five table it needs no fallback for significands
// of up to 19 digits -- Mushtak & Lemire, "Fast number parsing without fallback", SPE 2023).  Host + device: the same
// routine backs the ingest kernels and the host entry glove_parse_float32 that the CPU tests hammer against an
// exact-rational rounding.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GLOVE_HD __host__ __device__ __forceinline__
#else
#define GLOVE_HD inline
#endif

namespace glove {

constexpr int kPow5Min = -65, kPow5Max = 38;
struct Pow5 { uint64_t hi, lo; };
#if defined(__CUDACC__)
static __device__ const Pow5 kPow5Dev[kPow5Max - kPow5Min + 1] = {
#include "glove_pow5_table.inc"
};
#endif
static const Pow5 kPow5Host[kPow5Max - kPow5Min + 1] = {
#include "glove_pow5_table.inc"
};

GLOVE_HD void mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}
GLOVE_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// bits of the float32 nearest to w * 10^q (w != 0 handled, sign applied by the caller)
GLOVE_HD uint32_t decimal_to_f32_bits(uint64_t w, int32_t q) {
    constexpr int kMantBits = 23, kMinExp = -127, kInfPower = 0xFF;
    if (w == 0 || q < -64) return 0u;            // < 1.9e19 * 1e-65: below half the smallest subnormal
    if (q > kPow5Max) return 0x7f800000u;        // >= 1e39 > FLT_MAX
    const int lz = clz64(w);
    w <<= lz;
#if defined(__CUDA_ARCH__)
    const Pow5 t = kPow5Dev[q - kPow5Min];
#else
    const Pow5 t = kPow5Host[q - kPow5Min];
#endif
    uint64_t hi, lo;
    mul64(w, t.hi, hi, lo);
    constexpr uint64_t kPrecisionMask = 0xFFFFFFFFFFFFFFFFull >> (kMantBits + 3);
    if ((hi & kPrecisionMask) == kPrecisionMask) {   // the truncated product cannot decide: add the low-word product
        uint64_t hi2, lo2;
        mul64(w, t.lo, hi2, lo2);
        lo += hi2;
        if (hi2 > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 64 - kMantBits - 3;
    uint64_t mant = hi >> shift;
    // floor(log2(10^q)) + 63, 217706 / 65536 ~ log2(10)
    int power2 = (int)(((int64_t)217706 * q) >> 16) + 63 + upperbit - lz - kMinExp;
    if (power2 <= 0) {                            // subnormal result
        if (-power2 + 1 >= 64) return 0u;
        mant >>= -power2 + 1;
        mant += (mant & 1);
        mant >>= 1;
        return (uint32_t)mant;                    // mant == 2^23 encodes the smallest normal correctly
    }
    // exact halfway between two floats can only happen for small |q| (5^q must divide / fit the significand)
    if (lo <= 1 && q >= -17 && q <= 10 && (mant & 3) == 1 && (mant << shift) == hi) mant &= ~1ull;
    mant += (mant & 1);
    mant >>= 1;
    if (mant >= (2ull << kMantBits)) { mant = 1ull << kMantBits; ++power2; }
    mant &= ~(1ull << kMantBits);
    if (power2 >= kInfPower) return 0x7f800000u;
    return ((uint32_t)power2 << kMantBits) | (uint32_t)mant;
}

enum { STRTOF_OK = 0, STRTOF_BAD = 1, STRTOF_TOO_LONG = 2 };

GLOVE_HD bool ieq3(const uint8_t *p, char a, char b, char c) {
    return (p[0] | 32) == a && (p[1] | 32) == b && (p[2] | 32) == c;
}

// [sign] digits [. digits] [(e|E) [sign] digits] | [sign] inf | [sign] infinity | [sign] nan   (whole field, no blanks).
// Empty field -> +0.0 (make_csv_dataset's inferred column default).  More than 19 significant digits are accepted when
// dropping the tail provably does not change the rounding; otherwise STRTOF_TOO_LONG.
template <typename GetByte>
GLOVE_HD int parse_f32(GetByte at, int n, uint32_t &bits) {
    bits = 0;
    if (n == 0) return STRTOF_OK;
    int i = 0;
    uint32_t sign = 0;
    if (at(0) == '-' || at(0) == '+') { sign = at(0) == '-' ? 0x80000000u : 0u; i = 1; }
    if (i >= n) return STRTOF_BAD;
    const int c0 = at(i) | 32;
    if (c0 == 'i' || c0 == 'n') {
        uint8_t w[8];
        const int m = n - i;
        if (m != 3 && m != 8) return STRTOF_BAD;
        for (int k = 0; k < m; ++k) w[k] = at(i + k);
        if (m == 3 && ieq3(w, 'n', 'a', 'n')) { bits = sign | 0x7fc00000u; return STRTOF_OK; }
        if (ieq3(w, 'i', 'n', 'f') && (m == 3 || (ieq3(w + 3, 'i', 'n', 'i') && (w[6] | 32) == 't' && (w[7] | 32) == 'y'))) {
            bits = sign | 0x7f800000u;
            return STRTOF_OK;
        }
        return STRTOF_BAD;
    }
    uint64_t w = 0;
    int digits = 0, sig = 0, dropped_exp = 0, frac_exp = 0;
    bool truncated = false, seen_dot = false;
    for (; i < n; ++i) {
        const int c = at(i);
        if (c == '.') {
            if (seen_dot) return STRTOF_BAD;
            seen_dot = true;
            continue;
        }
        const unsigned dgt = (unsigned)(c - '0');
        if (dgt > 9) break;
        ++digits;
        if (seen_dot) --frac_exp;
        if (sig == 0 && dgt == 0) continue;          // leading zeros carry no significance
        if (sig < 19) { w = w * 10 + dgt; ++sig; }
        else { ++dropped_exp; truncated |= dgt != 0; }
    }
    if (digits == 0) return STRTOF_BAD;
    int64_t e10 = 0;
    if (i < n) {
        if ((at(i) | 32) != 'e') return STRTOF_BAD;
        ++i;
        bool eneg = false;
        if (i < n && (at(i) == '-' || at(i) == '+')) { eneg = at(i) == '-'; ++i; }
        if (i >= n) return STRTOF_BAD;
        for (; i < n; ++i) {
            const unsigned dgt = (unsigned)(at(i) - '0');
            if (dgt > 9) return STRTOF_BAD;
            if (e10 < 100000) e10 = e10 * 10 + dgt;
        }
        if (eneg) e10 = -e10;
    }
    int64_t q = e10 + frac_exp + dropped_exp;
    if (q < -1000) q = -1000;
    if (q > 1000) q = 1000;
    uint32_t r = decimal_to_f32_bits(w, (int32_t)q);
    if (truncated && r != decimal_to_f32_bits(w + 1, (int32_t)q)) return STRTOF_TOO_LONG;
    bits = sign | r;
    return STRTOF_OK;
}

}  // namespace glove
