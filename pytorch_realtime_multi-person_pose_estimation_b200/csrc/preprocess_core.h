// Image normalisations of get_outputs, host/device shared core.
//   rtpose_preprocess / vgg_preprocess / inception_preprocess / ssd_preprocess
//                                              /root/reference/lib/datasets/preprocessing.py:16-21, 32-43, 46-52, 77-86
// Every mode is a per-channel affine map of the uint8 BGR pixel, optionally with the channel order reversed (the
// reference's `[:, :, ::-1]`), evaluated in float32 with exactly the operations numpy performs (no FMA):
//   rtpose     out[c] = x[c] / 256 - 0.5
//   vgg        out[i] = (x[2-i] / 255 - mean[i]) / std[i]
//   inception  out[i] = x[2-i] / 128 - 1
//   ssd        out[c] = x[c] - (123, 117, 104)[c]          (R-104, G-117, B-123 computed in RGB and flipped back to BGR)
// conv_first_kernel applies it while it loads its input tile (zero padding is applied AFTER the normalisation, as
// nn.Conv2d pads the normalised tensor).  tests/test_host.py checks the host build against the numpy functions and the
// reference's golden output bit for bit.
#pragma once

#if defined(__CUDACC__)
#define B2P_PP_HD __host__ __device__ __forceinline__
#else
#define B2P_PP_HD inline
#endif

namespace b2p {

enum : int { kPreNone = 0, kPreRtpose = 1, kPreVgg = 2, kPreInception = 3, kPreSsd = 4 };

// input (BGR) channel that feeds output channel c
B2P_PP_HD int pre_src_channel(int mode, int c) { return (mode == kPreVgg || mode == kPreInception) ? 2 - c : c; }

B2P_PP_HD float pre_value(int mode, unsigned char u, int c /*output channel*/) {
    const float x = (float)u;
#if defined(__CUDA_ARCH__)
#define B2P_PP_DIV(a, b) __fdiv_rn(a, b)
#define B2P_PP_SUB(a, b) __fsub_rn(a, b)
#else
#define B2P_PP_DIV(a, b) ((a) / (b))
#define B2P_PP_SUB(a, b) ((a) - (b))
#endif
    if (mode == kPreRtpose) return B2P_PP_SUB(B2P_PP_DIV(x, 256.f), 0.5f);
    if (mode == kPreVgg) {
        const float m = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
        const float s = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        return B2P_PP_DIV(B2P_PP_SUB(B2P_PP_DIV(x, 255.f), m), s);
    }
    if (mode == kPreInception) return B2P_PP_SUB(B2P_PP_DIV(x, 128.f), 1.f);
    if (mode == kPreSsd) return B2P_PP_SUB(x, c == 0 ? 123.f : (c == 1 ? 117.f : 104.f));
    return x;
#undef B2P_PP_DIV
#undef B2P_PP_SUB
}

}  // namespace b2p
