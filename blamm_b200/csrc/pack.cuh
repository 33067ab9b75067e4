// ASCII -> 2-bit pack (+ zero-mask plane).  Replaces the host-side one-hot fill of
// SeqMatrix::getNextSeqMatrix (sequence.cpp:306-337: 16 B of FP32 written per character, 18 B/nt over PCIe)
// with 0.25 B/nt (+0.125 B/nt mask) produced on the device.  HBM-bound: 1 B read, 0.375 B written per char.
#pragma once
#include "common.cuh"

namespace b200 {

// Per character: code in ACGT order and a "contributes zero" bit.
//   upper-case ACGT           -> code, zero = 0
//   lower-case acgt           -> code, zero = (mode == LOWER_ZERO)   (sequence.cpp:312-319 tests upper case only)
//   anything else             -> code 0, zero = 1                    (cannot occur in a filtered block)
__device__ __forceinline__ void classify(uint32_t ch, bool fold_lower, uint32_t& code, uint32_t& zero)
{
    uint32_t up = ch & 0xDFu;
    uint32_t k = (up >> 1) & 3u;           // A(0x41)->0  C(0x43)->1  G(0x47)->3  T(0x54)->2
    code = k ^ (k >> 1);                   // -> A0 C1 G2 T3
    bool valid = (up == 0x41u) | (up == 0x43u) | (up == 0x47u) | (up == 0x54u);
    bool lower = (ch & 0x20u) != 0;
    zero = (!valid | (lower & !fold_lower)) ? 1u : 0u;
    if (!valid) code = 0;
}

// One thread packs 32 characters: two uint4 loads -> two code words + one mask word.
// `ascii` must be 16-byte aligned and readable up to a multiple of 32 characters (the staging buffers are
// padded); characters at or beyond n are packed as code 0 / zero 1.
__global__ void __launch_bounds__(256)
pack_ascii_kernel(const uint8_t* __restrict__ ascii, uint32_t n, int fold_lower,
                  uint32_t* __restrict__ codes, uint32_t* __restrict__ zmask, uint32_t* __restrict__ has_zero)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t base = t * 32u;
    uint32_t anyz = 0;
    if (base < n) {
        const uint4* src = reinterpret_cast<const uint4*>(ascii + base);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t c0 = 0, c1 = 0, zm = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            uint32_t ch = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
            uint32_t code, zero;
            classify(ch, fold_lower != 0, code, zero);
            if (base + i >= n) { code = 0; zero = 1; }
            if (i < 16) c0 |= code << (2 * i); else c1 |= code << (2 * (i - 16));
            zm |= zero << i;
        }
        codes[2 * t] = c0;
        codes[2 * t + 1] = c1;
        zmask[t] = zm;
        // the tail padding (>= n) does not count as "block has zero-contribution characters"
        uint32_t live = (n - base >= 32u) ? 0xffffffffu : ((1u << (n - base)) - 1u);
        anyz = zm & live;
    }
    anyz = __any_sync(0xffffffffu, anyz != 0);
    if (anyz && (threadIdx.x & 31) == 0) atomicOr(has_zero, 1u);
}

} // namespace b200
