// TEST INFRASTRUCTURE: standalone check of the flip test-time-averaging and crop_with_factor entry points through the
// C ABI only (host pointers; no torch, starts in a second).  The fused b200pose_infer*_flip calls are compared bit for bit with
// the composition of already-validated calls: forward(frames) + forward(host-mirrored frames) + host merge with
// csrc/tta_core.h + b200pose_post_run.
// Build: g++ -O2 -std=c++17 tests/cuda/test_flip.cpp -o build/test_flip -L<pkg> -lb200pose -Wl,-rpath,<pkg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../include/b200pose.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/preprocess_core.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/resize_core.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/tta_core.h"

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
    g_state ^= g_state << 13; g_state ^= g_state >> 7; g_state ^= g_state << 17;
    return (uint32_t)(g_state >> 32);
}
static inline float urand() { return (rnd() >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

#define CHECK(call)                                                                                      \
    do {                                                                                                 \
        int rc_ = (call);                                                                                \
        if (rc_) { printf("FAIL %s -> %d: %s\n", #call, rc_, b200pose_last_error()); return 1; }         \
    } while (0)

static int fetch(b200pose_post* post, int n, std::vector<std::vector<float>>& rows) {
    if (b200pose_post_sync(post)) { printf("FAIL sync: %s\n", b200pose_last_error()); return 1; }
    rows.assign(n, {});
    for (int i = 0; i < n; ++i) {
        const int st = b200pose_post_status(post, i);
        if (st < 0 || (st & 0xF)) { printf("FAIL status image %d = %d\n", i, st); return 1; }
        const int k = b200pose_post_num_humans(post, i);
        rows[i].resize((size_t)k * B200POSE_HUMAN_FLOATS);
        if (k && b200pose_post_get_humans(post, i, rows[i].data(), k) != k) { printf("FAIL get_humans\n"); return 1; }
    }
    return 0;
}

static bool same(const std::vector<std::vector<float>>& a, const std::vector<std::vector<float>>& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (a[i].size() != b[i].size() || (a[i].size() && memcmp(a[i].data(), b[i].data(), a[i].size() * 4))) return false;
    return true;
}

// host crop_with_factor with the shared core (the kernel's own functions compiled for the host)
static std::vector<unsigned char> host_crop(const std::vector<unsigned char>& raw, int n, int sh, int sw, int dest, int factor,
                                            b2p::CropGeom* geom) {
    const b2p::CropGeom g = b2p::crop_geometry(sh, sw, dest, factor);
    *geom = g;
    std::vector<unsigned char> out((size_t)n * g.pad_h * g.pad_w * 3);
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < g.pad_h; ++y)
            for (int x = 0; x < g.pad_w; ++x)
                for (int c = 0; c < 3; ++c)
                    out[(((size_t)i * g.pad_h + y) * g.pad_w + x) * 3 + c] =
                        b2p::crop_px(raw.data() + (size_t)i * sh * sw * 3, sh, sw, 3, g, y, x, c);
    return out;
}

static int raw_section(b200pose_net* net, b200pose_post* post) {
    int failures = 0;
    const int n = 2;
    // 1. the crop kernel alone: general bilinear (up- and down-scaling), the exact-2x INTER_AREA switch, a cut 2x box
    const int shapes[5][3] = {{150, 211, 184}, {480, 640, 368}, {368, 400, 184}, {368, 403, 184}, {97, 64, 200}};
    for (auto& sp : shapes) {
        const int sh = sp[0], sw = sp[1], dest = sp[2];
        std::vector<unsigned char> raw((size_t)n * sh * sw * 3);
        for (auto& b : raw) b = (unsigned char)(rnd() & 255);
        b2p::CropGeom g;
        const std::vector<unsigned char> want = host_crop(raw, n, sh, sw, dest, 8, &g);
        double sc; int rh, rw, ph, pw;
        CHECK(b200pose_crop_geometry(sh, sw, dest, 8, &sc, &rh, &rw, &ph, &pw));
        std::vector<unsigned char> got(want.size(), 0xAA);
        CHECK(b200pose_net_crop_with_factor(net, raw.data(), 0, n, sh, sw, dest, 8, got.data(), 0, nullptr));
        const bool ok = sc == g.im_scale && rh == g.res_h && rw == g.res_w && ph == g.pad_h && pw == g.pad_w && got == want;
        printf("%s  crop_with_factor kernel %dx%d -> %dx%d (pad %dx%d%s) vs host core\n", ok ? "PASS" : "FAIL", sh, sw, rh, rw,
               ph, pw, g.area2 ? ", 2x area" : "");
        failures += !ok;
    }
    // 2. fused raw path == host crop + the validated uint8 entry points, without and with flip averaging
    const int sh = 150, sw = 211, dest = 184;
    std::vector<unsigned char> raw((size_t)n * sh * sw * 3), raw_m(raw.size());
    for (auto& b : raw) b = (unsigned char)(rnd() & 255);
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < sh; ++y)
            for (int x = 0; x < sw; ++x)
                for (int c = 0; c < 3; ++c)
                    raw_m[(((size_t)i * sh + y) * sw + x) * 3 + c] = raw[(((size_t)i * sh + y) * sw + (sw - 1 - x)) * 3 + c];
    b2p::CropGeom g;
    const std::vector<unsigned char> fr = host_crop(raw, n, sh, sw, dest, 8, &g), fr_m = host_crop(raw_m, n, sh, sw, dest, 8, &g);
    const int H = g.pad_h, W = g.pad_w, h = H / 8, w = W / 8;
    std::vector<std::vector<float>> want, got;
    CHECK(b200pose_infer_u8(net, post, fr.data(), 0, n, H, W, 0, 0.1f, nullptr));
    if (fetch(post, n, want)) return failures + 1;
    CHECK(b200pose_infer_raw_u8(net, post, raw.data(), 0, n, sh, sw, dest, 8, 0, 0.1f, 0, nullptr));
    if (fetch(post, n, got)) return failures + 1;
    size_t persons = 0;
    for (auto& r : want) persons += r.size() / B200POSE_HUMAN_FLOATS;
    bool ok = same(want, got);
    printf("%s  b200pose_infer_raw_u8 == host crop + b200pose_infer_u8 [%dx%d -> %dx%d, %zu persons]\n", ok ? "PASS" : "FAIL", sh,
           sw, H, W, persons);
    failures += !ok;
    const size_t eh = (size_t)n * 19 * h * w, ep = (size_t)n * 38 * h * w;
    std::vector<float> paf_n(ep), heat_n(eh), paf_f(ep), heat_f(eh), avg_p(ep), avg_h(eh);
    float* outs[12] = {nullptr};
    outs[10] = paf_n.data(); outs[11] = heat_n.data();
    CHECK(b200pose_net_forward_u8(net, fr.data(), 0, n, H, W, 0, outs, 0, nullptr));
    outs[10] = paf_f.data(); outs[11] = heat_f.data();
    CHECK(b200pose_net_forward_u8(net, fr_m.data(), 0, n, H, W, 0, outs, 0, nullptr));
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                for (int c = 0; c < 19; ++c)
                    avg_h[(size_t)i * 19 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                        heat_n.data() + (size_t)i * 19 * h * w, heat_f.data() + (size_t)i * 19 * h * w, false, c, y, x, w,
                        (long)h * w, w, 1);
                for (int c = 0; c < 38; ++c)
                    avg_p[(size_t)i * 38 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                        paf_n.data() + (size_t)i * 38 * h * w, paf_f.data() + (size_t)i * 38 * h * w, true, c, y, x, w,
                        (long)h * w, w, 1);
            }
    CHECK(b200pose_post_run(post, avg_h.data(), avg_p.data(), 0, 0, n, h, w, 0.1f, nullptr));
    if (fetch(post, n, want)) return failures + 1;
    CHECK(b200pose_infer_raw_u8(net, post, raw.data(), 0, n, sh, sw, dest, 8, 0, 0.1f, 1, nullptr));
    if (fetch(post, n, got)) return failures + 1;
    ok = same(want, got);
    printf("%s  b200pose_infer_raw_u8(flip) == mirror raw + host crop + forward x2 + merge + post\n", ok ? "PASS" : "FAIL");
    failures += !ok;
    return failures;
}

// vgg / inception / ssd normalisation fused into conv1_1 == the same normalisation on the host (shared core) + the fp32
// entry point, bit for bit (not yet run on hardware in round 1: part of the "multiscale" group of sections)
static int preprocess_section(b200pose_net* net) {
    const int n = 2, H = 64, W = 72, h = H / 8, w = W / 8;
    std::vector<unsigned char> fr((size_t)n * H * W * 3);
    for (auto& b : fr) b = (unsigned char)(rnd() & 255);
    int failures = 0;
    const char* names[5] = {"", "rtpose", "vgg", "inception", "ssd"};
    for (int pm = 4; pm >= 1; --pm) {      // ends on rtpose, the default of every other section
        if (b200pose_net_set_preprocess(net, pm)) { printf("FAIL set_preprocess %d\n", pm); return failures + 1; }
        std::vector<float> x((size_t)n * 3 * H * W);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < 3; ++c)
                for (int y = 0; y < H; ++y)
                    for (int xx = 0; xx < W; ++xx)
                        x[(((size_t)i * 3 + c) * H + y) * W + xx] =
                            b2p::pre_value(pm, fr[(((size_t)i * H + y) * W + xx) * 3 + b2p::pre_src_channel(pm, c)], c);
        for (int mode = 0; mode < 2; ++mode) {
            std::vector<float> pa((size_t)n * 38 * h * w), ha((size_t)n * 19 * h * w), pb(pa.size()), hb(ha.size());
            float* outs[12] = {nullptr};
            outs[10] = pa.data(); outs[11] = ha.data();
            CHECK(b200pose_net_forward_u8(net, fr.data(), 0, n, H, W, mode, outs, 0, nullptr));
            outs[10] = pb.data(); outs[11] = hb.data();
            CHECK(b200pose_net_forward(net, x.data(), 0, n, H, W, mode, outs, 0, nullptr));
            const bool ok = !memcmp(pa.data(), pb.data(), pa.size() * 4) && !memcmp(ha.data(), hb.data(), ha.size() * 4);
            printf("%s  preprocess %-9s fused into conv1_1 == host normalisation + fp32 entry [%s]\n", ok ? "PASS" : "FAIL",
                   names[pm], mode ? "fp32" : "bf16");
            failures += !ok;
        }
    }
    return failures;
}

// Multi-scale (+ flip) averaging == per scale: host crop -> validated forward (-> host flip merge) -> host bicubic resize
// (shared core) -> float32 running sum -> / n_scales -> validated post_run.
static int multiscale_section(b200pose_net* net, b200pose_post* post, int flip) {
    const int n = 2, sh = 90, sw = 123, base = 96, ns = 4;
    const double scales[ns] = {0.5, 1.0, 1.5, 2.0};
    std::vector<unsigned char> raw((size_t)n * sh * sw * 3), raw_m(raw.size());
    g_state = 0xD1B54A32D192ED03ull;
    for (auto& b : raw) b = (unsigned char)(rnd() & 255);
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < sh; ++y)
            for (int x = 0; x < sw; ++x)
                for (int c = 0; c < 3; ++c)
                    raw_m[(((size_t)i * sh + y) * sw + x) * 3 + c] = raw[(((size_t)i * sh + y) * sw + (sw - 1 - x)) * 3 + c];
    b2p::CropGeom g1 = b2p::crop_geometry(sh, sw, base, 8);
    const int h1 = g1.pad_h / 8, w1 = g1.pad_w / 8;
    std::vector<float> acc_h((size_t)n * 19 * h1 * w1), acc_p((size_t)n * 38 * h1 * w1);
    for (int k = 0; k < ns; ++k) {
        b2p::CropGeom g;
        const int dest = (int)((double)base * scales[k]);
        const std::vector<unsigned char> fr = host_crop(raw, n, sh, sw, dest, 8, &g);
        const int H = g.pad_h, W = g.pad_w, h = H / 8, w = W / 8;
        const size_t eh = (size_t)n * 19 * h * w, ep = (size_t)n * 38 * h * w;
        std::vector<float> heat(eh), paf(ep);
        float* outs[12] = {nullptr};
        outs[10] = paf.data(); outs[11] = heat.data();
        CHECK(b200pose_net_forward_u8(net, fr.data(), 0, n, H, W, 0, outs, 0, nullptr));
        if (flip) {
            const std::vector<unsigned char> fr_m = host_crop(raw_m, n, sh, sw, dest, 8, &g);
            std::vector<float> heat_f(eh), paf_f(ep), mh(eh), mp(ep);
            outs[10] = paf_f.data(); outs[11] = heat_f.data();
            CHECK(b200pose_net_forward_u8(net, fr_m.data(), 0, n, H, W, 0, outs, 0, nullptr));
            for (int i = 0; i < n; ++i)
                for (int y = 0; y < h; ++y)
                    for (int x = 0; x < w; ++x) {
                        for (int c = 0; c < 19; ++c)
                            mh[(size_t)i * 19 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                                heat.data() + (size_t)i * 19 * h * w, heat_f.data() + (size_t)i * 19 * h * w, false, c, y, x, w,
                                (long)h * w, w, 1);
                        for (int c = 0; c < 38; ++c)
                            mp[(size_t)i * 38 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                                paf.data() + (size_t)i * 38 * h * w, paf_f.data() + (size_t)i * 38 * h * w, true, c, y, x, w,
                                (long)h * w, w, 1);
                    }
            heat.swap(mh); paf.swap(mp);
        }
        const double sy = b2p::rs_step(h1, h), sx = b2p::rs_step(w1, w);
        auto accumulate = [&](const std::vector<float>& src, std::vector<float>& acc, int C) {
            for (long pl = 0; pl < (long)n * C; ++pl)
                for (int y = 0; y < h1; ++y) {
                    const b2p::CubCoef cy = b2p::rs_cubic_coef(y, h, sy);
                    for (int x = 0; x < w1; ++x) {
                        const b2p::CubCoef cx = b2p::rs_cubic_coef(x, w, sx);
                        float v = b2p::rs_cubic_at(src.data() + pl * h * w, w, 1, cx, cy);
                        float& a = acc[(size_t)pl * h1 * w1 + (size_t)y * w1 + x];
                        if (k > 0) v = a + v;
                        if (k == ns - 1) v = v / (float)ns;
                        a = v;
                    }
                }
        };
        accumulate(heat, acc_h, 19);
        accumulate(paf, acc_p, 38);
    }
    std::vector<std::vector<float>> want, got;
    CHECK(b200pose_post_run(post, acc_h.data(), acc_p.data(), 0, 0, n, h1, w1, 0.1f, nullptr));
    if (fetch(post, n, want)) return 1;
    CHECK(b200pose_infer_raw_u8_multiscale(net, post, raw.data(), 0, n, sh, sw, base, 8, scales, ns, 0, 0.1f, flip, nullptr));
    if (fetch(post, n, got)) return 1;
    size_t persons = 0;
    for (auto& r : want) persons += r.size() / B200POSE_HUMAN_FLOATS;
    const bool ok = same(want, got);
    printf("%s  b200pose_infer_raw_u8_multiscale(flip=%d) == composed path [4 scales, base %dx%d, %zu persons]\n",
           ok ? "PASS" : "FAIL", flip, g1.pad_h, g1.pad_w, persons);
    return !ok;
}

int main(int argc, char** argv) {
    const int n = 2, H = argc > 1 ? atoi(argv[1]) : 184, W = argc > 2 ? atoi(argv[2]) : 248;
    const bool with_multiscale = !(argc > 3 && !strcmp(argv[3], "no-multiscale"));   // usage: test_flip [H W [no-multiscale]]
    const int h = H / 8, w = W / 8;
    b200pose_net* net = nullptr;
    b200pose_post* post = nullptr;
    CHECK(b200pose_net_create(&net, 0));
    for (int t = 0; t < B200POSE_NUM_TENSORS; ++t) {
        long d[4] = {1, 1, 1, 1};
        const int nd = b200pose_net_tensor_shape(t, d);
        long cnt = 1;
        for (int i = 0; i < nd; ++i) cnt *= d[i];
        std::vector<float> v(cnt);
        if (nd == 4) {   // uniform with the He variance 2 / fan_in
            const float a = sqrtf(6.0f / (float)(d[1] * d[2] * d[3]));
            for (auto& x : v) x = (2.f * urand() - 1.f) * a;
        } else
            for (auto& x : v) x = 0.2f * urand() - 0.1f;
        CHECK(b200pose_net_set_tensor(net, t, v.data(), cnt));
    }
    CHECK(b200pose_net_finalize(net));
    CHECK(b200pose_post_create(&post, 0, 4, 1024, 2048));

    const size_t fe = (size_t)n * H * W * 3;
    std::vector<unsigned char> frames(fe), mirrored(fe);
    for (auto& b : frames) b = (unsigned char)(rnd() & 255);
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                for (int c = 0; c < 3; ++c)
                    mirrored[(((size_t)i * H + y) * W + x) * 3 + c] = frames[(((size_t)i * H + y) * W + (W - 1 - x)) * 3 + c];
    std::vector<float> pre((size_t)n * 3 * H * W);     // rtpose_preprocess of `frames`, NCHW
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    pre[(((size_t)i * 3 + c) * H + y) * W + x] = frames[(((size_t)i * H + y) * W + x) * 3 + c] / 256.0f - 0.5f;

    const size_t eh = (size_t)n * 19 * h * w, ep = (size_t)n * 38 * h * w;
    int failures = 0;
    const char* mode_name[3] = {"bf16", "fp32", "bf16x3"};
    for (int mode = 0; mode < 3; ++mode) {
        std::vector<float> paf_n(ep), heat_n(eh), paf_f(ep), heat_f(eh), avg_p(ep), avg_h(eh), dev_p(ep), dev_h(eh);
        float* outs[12] = {nullptr};
        outs[10] = paf_n.data(); outs[11] = heat_n.data();
        CHECK(b200pose_net_forward_u8(net, frames.data(), 0, n, H, W, mode, outs, 0, nullptr));
        outs[10] = paf_f.data(); outs[11] = heat_f.data();
        CHECK(b200pose_net_forward_u8(net, mirrored.data(), 0, n, H, W, mode, outs, 0, nullptr));
        for (int i = 0; i < n; ++i)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    for (int c = 0; c < 19; ++c)
                        avg_h[(size_t)i * 19 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                            heat_n.data() + (size_t)i * 19 * h * w, heat_f.data() + (size_t)i * 19 * h * w, false, c, y, x, w,
                            (long)h * w, w, 1);
                    for (int c = 0; c < 38; ++c)
                        avg_p[(size_t)i * 38 * h * w + ((size_t)c * h + y) * w + x] = b2p::tta_flip_merge_at(
                            paf_n.data() + (size_t)i * 38 * h * w, paf_f.data() + (size_t)i * 38 * h * w, true, c, y, x, w,
                            (long)h * w, w, 1);
                }
        // 1. the merge kernel alone (host pointers, NCHW)
        CHECK(b200pose_flip_merge(post, heat_n.data(), heat_f.data(), paf_n.data(), paf_f.data(), 0, 0, n, h, w, dev_h.data(),
                                  dev_p.data(), nullptr));
        const bool m_ok = !memcmp(dev_h.data(), avg_h.data(), eh * 4) && !memcmp(dev_p.data(), avg_p.data(), ep * 4);
        printf("%s  flip_merge kernel (NCHW) vs host core [%s]\n", m_ok ? "PASS" : "FAIL", mode_name[mode]);
        failures += !m_ok;
        if (mode == 0) {   // NHWC layout of the same data
            std::vector<float> a(eh), b(eh), c(ep), d(ep), oh(eh), op(ep);
            auto to_hwc = [&](const std::vector<float>& src, std::vector<float>& dst, int ch) {
                for (int i = 0; i < n; ++i)
                    for (int k = 0; k < ch; ++k)
                        for (int y = 0; y < h; ++y)
                            for (int x = 0; x < w; ++x)
                                dst[(((size_t)i * h + y) * w + x) * ch + k] = src[(((size_t)i * ch + k) * h + y) * w + x];
            };
            to_hwc(heat_n, a, 19); to_hwc(heat_f, b, 19); to_hwc(paf_n, c, 38); to_hwc(paf_f, d, 38);
            CHECK(b200pose_flip_merge(post, a.data(), b.data(), c.data(), d.data(), 0, 1, n, h, w, oh.data(), op.data(), nullptr));
            std::vector<float> eh_(eh), ep_(ep);
            to_hwc(avg_h, eh_, 19); to_hwc(avg_p, ep_, 38);
            const bool ok = !memcmp(oh.data(), eh_.data(), eh * 4) && !memcmp(op.data(), ep_.data(), ep * 4);
            printf("%s  flip_merge kernel (NHWC) vs host core\n", ok ? "PASS" : "FAIL");
            failures += !ok;
        }
        // 2. composition of validated calls -> person rows
        std::vector<std::vector<float>> want, got_u8, got_f32;
        CHECK(b200pose_post_run(post, avg_h.data(), avg_p.data(), 0, 0, n, h, w, 0.1f, nullptr));
        if (fetch(post, n, want)) return 1;
        // 3. fused paths
        CHECK(b200pose_infer_u8_flip(net, post, frames.data(), 0, n, H, W, mode, 0.1f, nullptr));
        if (fetch(post, n, got_u8)) return 1;
        CHECK(b200pose_infer_flip(net, post, pre.data(), 0, n, H, W, mode, 0.1f, nullptr));
        if (fetch(post, n, got_f32)) return 1;
        size_t persons = 0;
        for (auto& r : want) persons += r.size() / B200POSE_HUMAN_FLOATS;
        const bool ok_u8 = same(want, got_u8), ok_f32 = same(want, got_f32);
        printf("%s  b200pose_infer_u8_flip == composed path [%s, %zu persons]\n", ok_u8 ? "PASS" : "FAIL", mode_name[mode], persons);
        printf("%s  b200pose_infer_flip    == composed path [%s]\n", ok_f32 ? "PASS" : "FAIL", mode_name[mode]);
        failures += !ok_u8 + !ok_f32;
        // the un-flipped fused path must differ from the averaged one on random maps (the test would be vacuous otherwise)
        if (mode == 0) {
            std::vector<std::vector<float>> plain;
            CHECK(b200pose_infer_u8(net, post, frames.data(), 0, n, H, W, mode, 0.1f, nullptr));
            if (fetch(post, n, plain)) return 1;
            printf("%s  averaged result differs from the single-orientation result\n", same(want, plain) ? "WARN" : "PASS");
        }
    }
    failures += raw_section(net, post);
    if (with_multiscale) {
        failures += preprocess_section(net);
        failures += multiscale_section(net, post, 0);
        failures += multiscale_section(net, post, 1);
    }
    b200pose_post_destroy(post);
    b200pose_net_destroy(net);
    printf("%s (%d failures), kernels launched: %ld\n", failures ? "FLIP TEST FAILED" : "FLIP TEST OK", failures,
           b200pose_launch_count());
    return failures ? 1 : 0;
}
