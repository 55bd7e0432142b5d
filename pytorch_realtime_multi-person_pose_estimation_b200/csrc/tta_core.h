// Flip test-time averaging, host/device shared core.
//   handle_paf_and_heat      /root/reference/evaluate/coco_eval.py:197-242
// averaged[c](y, x) = (normal[c](y, x) +/- flipped[swap[c]](y, w-1-x)) / 2, the sign being minus for the PAF x
// components (the even PAF channels; the left/right permutation keeps the parity of a channel).
// The same functions are compiled for the device (tta.cu) and for the host (tests/cuda/post_core_host.cpp), so the
// index logic is checked against the oracle and the reference's golden vector on a machine without a GPU.
#pragma once

#if defined(__CUDACC__)
#define B2P_TTA_HD __host__ __device__ __forceinline__
#else
#define B2P_TTA_HD inline
#endif

namespace b2p {

constexpr int kTtaHeat = 19, kTtaPaf = 38;

// left/right partner of every heat-map channel (coco_eval.py:207-208) and PAF channel (:228-230)
#define B2P_SWAP_HEAT {0, 1, 5, 6, 7, 2, 3, 4, 11, 12, 13, 8, 9, 10, 15, 14, 17, 16, 18}
#define B2P_SWAP_PAF                                                                                                  \
    {6, 7, 8, 9, 10, 11, 0, 1, 2, 3, 4, 5, 20, 21, 22, 23, 24, 25, 26, 27, 12, 13, 14, 15, 16, 17, 18, 19, 28, 29, 32, \
     33, 30, 31, 36, 37, 34, 35}

B2P_TTA_HD int tta_swap_channel(bool paf, int c) {
    // spelled as arithmetic so that no table has to live in constant memory on the device
    if (!paf) {
        if (c >= 2 && c <= 4) return c + 3;          // right arm  -> left arm
        if (c >= 5 && c <= 7) return c - 3;
        if (c >= 8 && c <= 10) return c + 3;         // right leg  -> left leg
        if (c >= 11 && c <= 13) return c - 3;
        if (c >= 14 && c <= 17) return c ^ 1;        // eyes, ears
        return c;                                    // nose, neck, background
    }
    if (c < 6) return c + 6;                         // neck-hip-knee-ankle chains
    if (c < 12) return c - 6;
    if (c < 20) return c + 8;                        // neck-shoulder-elbow-wrist, shoulder-ear
    if (c < 28) return c - 8;
    if (c < 30) return c;                            // neck-nose
    if (c < 32) return c + 2;                        // nose-eye
    if (c < 34) return c - 2;
    if (c < 36) return c + 2;                        // eye-ear
    return c - 2;
}

// One output element.  `normal` / `flipped` address one image: value(c, y, x) = p[c*sc + y*sy + x*sx].
B2P_TTA_HD float tta_flip_merge_at(const float* normal, const float* flipped, bool paf, int c, int y, int x, int w,
                                   long sc, long sy, long sx) {
    const float a = normal[c * sc + y * sy + x * sx];
    float b = flipped[tta_swap_channel(paf, c) * sc + y * sy + (w - 1 - x) * sx];
    if (paf && (c & 1) == 0) b = -b;
    return (a + b) * 0.5f;      // float32 add, then an exact halving: identical to numpy's (a + b) / 2.
}

}  // namespace b2p
