// TEST INFRASTRUCTURE: host-only driver around csrc/post_core.h (the sequential-exact cores the CUDA kernels
// run on the device) so they can be checked against the oracle on a machine without a GPU.
// Build: g++ -O2 -ffp-contract=off -shared -fPIC post_core_host.cpp -o build/libpostcore_host.so
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/post_core.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/preprocess_core.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/resize_core.h"
#include "../../pytorch_realtime_multi-person_pose_estimation_b200/csrc/tta_core.h"

using namespace b2p;

static const int LIMB_PARTS[19][2] = B2P_LIMB_TABLES;
static const int LIMB_PAF[19][2] = B2P_LIMB_PAF_TABLES;

static std::vector<float> g_out;   // per human: score, then 18 x (x, y, peak score, cid)  (cid < 0 = missing)
static int g_degraded = 0, g_ties = 0, g_sort_mismatch = 0;
static long g_cand_total = 0, g_cand_needed = 0;   // statistics: candidates per frame / candidates up to the last accepted one

extern "C" int core_process(int n_peaks, const float* peaks /*[P][5] sorted by part*/, int h_up, const float* paf,
                            long sc, long sy, long sx, int shift, int list_cap_stress) {
    (void)list_cap_stress;
    std::vector<int> px[18], py[18];
    std::vector<float> ps[18];
    for (int i = 0; i < n_peaks; ++i) {
        int part = (int)peaks[i * 5 + 4];
        px[part].push_back((int)peaks[i * 5 + 0]);
        py[part].push_back((int)peaks[i * 5 + 1]);
        ps[part].push_back(peaks[i * 5 + 2]);
    }
    int part_base[19];
    part_base[0] = 0;
    for (int p = 0; p < 18; ++p) part_base[p + 1] = part_base[p] + (int)px[p].size();
    std::vector<float> peak_score(n_peaks + 1);
    std::vector<int> peak_x(n_peaks + 1), peak_y(n_peaks + 1);
    for (int p = 0; p < 18; ++p)
        for (size_t i = 0; i < px[p].size(); ++i) {
            peak_score[part_base[p] + i] = ps[p][i];
            peak_x[part_base[p] + i] = px[p][i];
            peak_y[part_base[p] + i] = py[p][i];
        }
    PafView pv{paf, sc, sy, sx, shift};
    std::vector<int> ca[19], cb[19];
    std::vector<float> cs[19];
    g_ties = 0;
    g_sort_mismatch = 0;
    g_cand_total = g_cand_needed = 0;
    int total_conn = 0;
    for (int l = 0; l < 19; ++l) {
        const int pa = LIMB_PARTS[l][0], pb = LIMB_PARTS[l][1];
        const int na = (int)px[pa].size(), nb = (int)px[pb].size();
        if (na == 0 || nb == 0) continue;
        std::vector<uint64_t> keys;
        for (int a = 0; a < na; ++a)
            for (int b = 0; b < nb; ++b) {
                float s;
                if (pair_score(pv, LIMB_PAF[l][0], LIMB_PAF[l][1], px[pa][a], py[pa][a], px[pb][b], py[pb][b], h_up, &s))
                    keys.push_back(cand_key(s, (uint32_t)(a * nb + b)));
            }
        std::vector<uint64_t> sorted = keys;
        std::sort(sorted.begin(), sorted.end());   // what the device bitonic sort produces (keys are unique)
        bool ties = false;
        for (size_t i = 0; i + 1 < sorted.size(); ++i)
            if ((sorted[i] >> 32) == (sorted[i + 1] >> 32)) ties = true;
        if (ties) ++g_ties;
        {   // the product always runs the exact std::sort emulation; check both formulations agree
            std::vector<uint64_t> seq = keys, par = keys;
            seq_std_sort(seq.data(), (int)seq.size());
            par_std_sort_host(par.data(), (int)par.size());
            if (seq != par) ++g_sort_mismatch;
            if (!ties && seq != sorted) ++g_sort_mismatch;   // without ties every correct sort gives this order
            sorted = par;
        }
        const int maxc = na < nb ? na : nb;
        std::vector<uint32_t> ua((na + 31) / 32, 0), ub((nb + 31) / 32, 0);
        ca[l].resize(maxc); cb[l].resize(maxc); cs[l].resize(maxc);
        int nc = greedy_match(sorted.data(), (int)sorted.size(), nb, ua.data(), ub.data(), maxc, ca[l].data(),
                              cb[l].data(), cs[l].data());
        ca[l].resize(nc); cb[l].resize(nc); cs[l].resize(nc);
        total_conn += nc;
        {   // how far down the sorted list does the greedy rule have to look?  (until min(na, nb) connections exist, else
            // to the end) - the potential of sorting lazily, left to right
            std::vector<char> fa(na, 0), fb(nb, 0);
            long last = (long)sorted.size();
            int made = 0;
            for (size_t i = 0; i < sorted.size(); ++i) {
                const uint32_t pair = (uint32_t)sorted[i];
                const int a = pair / nb, b = pair % nb;
                if (!fa[a] && !fb[b]) { fa[a] = fb[b] = 1; if (++made == maxc) { last = (long)i + 1; break; } }
            }
            g_cand_total += (long)sorted.size();
            g_cand_needed += last;
        }
    }
    std::vector<float> rows((size_t)(total_conn + 1) * kRowFloats);
    std::vector<uint8_t> alive(total_conn + 1, 0), list_n(n_peaks + 1, 0);
    std::vector<int32_t> lists((size_t)(n_peaks + 1) * kListCap);
    Assembler as;
    as.rows = rows.data(); as.alive = alive.data(); as.lists = lists.data(); as.list_n = list_n.data();
    as.part_base = part_base; as.peak_score = peak_score.data();
    as.row_cap = total_conn + 1; as.nrows = 0; as.degraded = 0; as.overflow = 0;
    for (int l = 0; l < 19; ++l)
        for (size_t c = 0; c < ca[l].size(); ++c)
            as.add_connection(l, LIMB_PARTS[l][0], LIMB_PARTS[l][1], part_base[LIMB_PARTS[l][0]] + ca[l][c],
                              part_base[LIMB_PARTS[l][1]] + cb[l][c], cs[l][c]);
    g_degraded = as.degraded;
    g_out.clear();
    int nh = 0;
    for (int r = 0; r < as.nrows; ++r) {
        if (!as.keep(r)) continue;
        const float* row = rows.data() + r * kRowFloats;
        g_out.push_back(f_div(row[18], row[19]));
        for (int p = 0; p < 18; ++p) {
            int cid = (int)row[p];
            if (cid < 0) { g_out.insert(g_out.end(), {0.f, 0.f, 0.f, -1.f}); continue; }
            g_out.insert(g_out.end(), {(float)peak_x[cid], (float)peak_y[cid], peak_score[cid], (float)cid});
        }
        ++nh;
    }
    return nh;
}
extern "C" const float* core_result() { return g_out.data(); }
extern "C" int core_degraded() { return g_degraded; }
extern "C" int core_ties() { return g_ties; }
extern "C" long core_cand_total() { return g_cand_total; }
extern "C" long core_cand_needed() { return g_cand_needed; }
extern "C" int core_sort_mismatch() { return g_sort_mismatch; }
// libstdc++'s std::sort order of `n` keys (the sequential restatement), for comparisons with the device kernels
extern "C" void core_seq_sort(const uint64_t* keys, int n, uint64_t* out) {
    for (int i = 0; i < n; ++i) out[i] = keys[i];
    seq_std_sort(out, n);
}
// direct test hook: sort `n` keys with both formulations, return 0 when identical
extern "C" int core_sort_check(uint64_t* keys, int n, uint64_t* out) {
    std::vector<uint64_t> seq(keys, keys + n), par(keys, keys + n), blk(keys, keys + n), blk2(keys, keys + n);
    seq_std_sort(seq.data(), n);
    par_std_sort_host(par.data(), n);                 // warp-chunked partition everywhere
    par_std_sort_host(blk.data(), n, 64, 7);          // rank-based block partition for ranges > 64, 7 virtual threads
    par_std_sort_host(blk2.data(), n, 300, 512);      // ... > 300 with 512 virtual threads (mostly empty slices)
    std::vector<uint64_t> dev(keys, keys + n);        // the device configuration: > 512 block-level, <= 512 warp rank tables
    par_std_sort_host(dev.data(), n, 512, 512, 512, 64);   // + ranges <= 64 by the per-lane sequential sort
    for (int i = 0; i < n; ++i) out[i] = par[i];
    return (seq == par ? 0 : 1) | (seq == blk ? 0 : 2) | (seq == blk2 ? 0 : 4) | (seq == dev ? 0 : 8);
}

// flip test-time averaging with the exact functions the CUDA kernel calls (csrc/tta_core.h); layout 0 = [n,C,h,w],
// 1 = [n,h,w,C].  Returns the number of channel-table mismatches between the arithmetic permutation and the tables.
extern "C" int core_flip_merge(const float* normal, const float* flipped, float* out, int n, int channels, int h, int w,
                               int layout) {
    static const int SH[19] = B2P_SWAP_HEAT;
    static const int SP[38] = B2P_SWAP_PAF;
    int bad = 0;
    for (int c = 0; c < 19; ++c) bad += tta_swap_channel(false, c) != SH[c];
    for (int c = 0; c < 38; ++c) bad += tta_swap_channel(true, c) != SP[c];
    const bool paf = channels == kTtaPaf;
    const long per = (long)channels * h * w;
    const long sc = layout == 0 ? (long)h * w : 1, sy = layout == 0 ? w : (long)w * channels, sx = layout == 0 ? 1 : channels;
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < channels; ++c)
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x)
                    out[i * per + c * sc + y * sy + x * sx] =
                        tta_flip_merge_at(normal + i * per, flipped + i * per, paf, c, y, x, w, sc, sy, sx);
    return bad;
}

// crop_with_factor with the exact functions the CUDA kernel calls (csrc/resize_core.h).  geom: im_scale, res_h, res_w,
// pad_h, pad_w, area2.  out (may be NULL to query the geometry only): uint8 [pad_h, pad_w, 3].
extern "C" int core_crop_with_factor(const unsigned char* src, int src_h, int src_w, int dest_size, int factor,
                                     unsigned char* out, double* geom) {
    const CropGeom g = crop_geometry(src_h, src_w, dest_size, factor);
    geom[0] = g.im_scale; geom[1] = g.res_h; geom[2] = g.res_w; geom[3] = g.pad_h; geom[4] = g.pad_w; geom[5] = g.area2;
    if (!out) return 0;
    for (int y = 0; y < g.pad_h; ++y)
        for (int x = 0; x < g.pad_w; ++x)
            for (int c = 0; c < 3; ++c) out[((long)y * g.pad_w + x) * 3 + c] = crop_px(src, src_h, src_w, 3, g, y, x, c);
    return 0;
}

// bicubic map resize with the exact functions the CUDA kernel calls (csrc/resize_core.h); src [sh, sw, C] HWC float32
// -> out [dh, dw, C].
extern "C" int core_resize_cubic(const float* src, int sh, int sw, int C, float* out, int dh, int dw) {
    const double step_y = rs_step(dh, sh), step_x = rs_step(dw, sw);
    for (int y = 0; y < dh; ++y) {
        const CubCoef cy = rs_cubic_coef(y, sh, step_y);
        for (int x = 0; x < dw; ++x) {
            const CubCoef cx = rs_cubic_coef(x, sw, step_x);
            for (int c = 0; c < C; ++c) out[((long)y * dw + x) * C + c] = rs_cubic_at(src + c, (long)sw * C, C, cx, cy);
        }
    }
    return 0;
}

// image normalisation with the exact functions conv_first_kernel calls (csrc/preprocess_core.h): uint8 HWC BGR [h,w,3]
// -> float32 CHW [3,h,w]; mode 1 rtpose, 2 vgg, 3 inception, 4 ssd.
extern "C" int core_preprocess(int mode, const unsigned char* img, int h, int w, float* out) {
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                out[((long)c * h + y) * w + x] = pre_value(mode, img[((long)y * w + x) * 3 + pre_src_channel(mode, c)], c);
    return 0;
}
