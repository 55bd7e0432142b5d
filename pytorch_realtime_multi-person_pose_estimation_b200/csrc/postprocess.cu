// Fused post-processing kernel family for sm_100a: heat-map NMS + bicubic refinement, PAF line-integral scoring
// + greedy bipartite matching, person assembly.  See postprocess.cuh for the reference lines each kernel replaces.
// HBM-side this path touches 56 x h x w x 4 B per image once (18 heat + 38 PAF planes); it is latency bound, so
// the design goal is: everything stays on the device, one block per (image, part) / (image, limb) / image, warp
// shuffles for ordered compaction and arg-max, shared memory for the plane / candidate keys.
#include <cstdio>
#include <cstdlib>

#include "host_util.h"
#include "postprocess.cuh"

namespace b2p {

namespace {

__constant__ int c_limb_parts[kNumLimb][2] = B2P_LIMB_TABLES;
__constant__ int c_limb_paf[kNumLimb][2] = B2P_LIMB_PAF_TABLES;

// OpenCV interpolateCubic (A = -0.75) for the 8 destination phases of an x8 up-sampling; exact float32 values
// (the 8 x 4 table is pinned against OpenCV in tests/test_oracle.py).  Phases 0-3 start at tap floor(src)-1 = X/8 - 2, phases 4-7 at X/8 - 1.
__constant__ float c_cubic[8][4] = {
    {-0x1.4acp-4f, 0x1.0568p-1f, 0x1.5918p-1f, -0x1.a94p-4f}, {-0x1.9c8p-5f, 0x1.5efp-2f, 0x1.a308p-1f, -0x1.c5cp-4f},
    {-0x1.5fp-6f, 0x1.7b2p-3f, 0x1.dbb8p-1f, -0x1.7c4p-4f},   {-0x1.68p-9f, 0x1.ad8p-5f, 0x1.fba8p-1f, -0x1.518p-5f},
    {-0x1.518p-5f, 0x1.fba8p-1f, 0x1.ad8p-5f, -0x1.68p-9f},   {-0x1.7c4p-4f, 0x1.dbb8p-1f, 0x1.7b2p-3f, -0x1.5fp-6f},
    {-0x1.c5cp-4f, 0x1.a308p-1f, 0x1.5efp-2f, -0x1.9c8p-5f},  {-0x1.a94p-4f, 0x1.5918p-1f, 0x1.0568p-1f, -0x1.4acp-4f}};

constexpr int kPeakThreads = 256;
#ifndef B2P_LIMB_THREADS
#define B2P_LIMB_THREADS 512
#endif
#ifndef B2P_SMEM_RANGE
#define B2P_SMEM_RANGE 4096
#endif
constexpr int kLimbThreads = B2P_LIMB_THREADS;
constexpr int kAsmThreads = 128;
constexpr int kHorStride = 40;   // (2*2+1) * 8

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------ peaks
__global__ void __launch_bounds__(kPeakThreads) peaks_kernel(PostBuffers pb, const float* __restrict__ heat, long h_img,
                                                             long h_ch, long h_y, long h_x, int h, int w, float thresh) {
    extern __shared__ float sm_f[];
    float* plane = sm_f;                       // [h*w]
    float* hor_all = sm_f + h * w;             // [8 warps][5][40]
    __shared__ int warp_cnt[kPeakThreads / 32];
    __shared__ int s_base;
    const int part = blockIdx.x, img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = h * w;
    const float* src = heat + img * h_img + part * h_ch;
    for (int i = tid; i < hw; i += kPeakThreads) plane[i] = src[(long)(i / w) * h_y + (long)(i % w) * h_x];
    if (tid == 0) s_base = 0;
    __syncthreads();

    const int cap = pb.peak_cap;
    int* px = pb.peak_x + ((long)img * kNumPart + part) * cap;
    int* py = pb.peak_y + ((long)img * kNumPart + part) * cap;
    float* ps = pb.peak_s + ((long)img * kNumPart + part) * cap;

    // find_peaks: ordered (raster) compaction
    for (int start = 0; start < hw; start += kPeakThreads) {
        const int i = start + tid;
        bool flag = false;
        int x = 0, y = 0;
        if (i < hw) {
            y = i / w;
            x = i - y * w;
            const float v = plane[i];
            flag = (v > thresh) && (y == 0 || v >= plane[i - w]) && (y == h - 1 || v >= plane[i + w]) &&
                   (x == 0 || v >= plane[i - 1]) && (x == w - 1 || v >= plane[i + 1]);
        }
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int pre = s_base, tot = 0;
        for (int k = 0; k < kPeakThreads / 32; ++k) {
            if (k < warp) pre += warp_cnt[k];
            tot += warp_cnt[k];
        }
        const int pos = pre + __popc(m & ((1u << lane) - 1u));
        if (flag && pos < cap) {
            px[pos] = x;
            py[pos] = y;
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    int n = s_base;
    if (n > cap) {
        if (tid == 0) atomicOr(&pb.status[img], 1);
        n = cap;
    }
    if (tid == 0) pb.counts[img * kNumPart + part] = n;

    // refinement: 5x5 window (clipped) -> x8 bicubic -> first arg-max        (one warp per peak)
    float* hor = hor_all + warp * 5 * kHorStride;
    for (int pk = warp; pk < n; pk += kPeakThreads / 32) {
        const int x = px[pk], y = py[pk];
        const int x_min = max(0, x - 2), x_max = min(w - 1, x + 2);
        const int y_min = max(0, y - 2), y_max = min(h - 1, y + 2);
        const int pw = x_max - x_min + 1, ph = y_max - y_min + 1;
        const int W8 = pw * 8, H8 = ph * 8;
        __syncwarp();
        // horizontal pass: lanes own columns X = lane and lane + 32 (W8 <= 40), rows are walked - no divisions
        for (int xi = 0; xi < 2; ++xi) {
            const int X = lane + 32 * xi;
            if (X < W8) {
                const int phase = X & 7;
                const int sx = (X >> 3) + (phase < 4 ? -2 : -1);
                const int t0 = clampi(sx, 0, pw - 1), t1 = clampi(sx + 1, 0, pw - 1);
                const int t2 = clampi(sx + 2, 0, pw - 1), t3 = clampi(sx + 3, 0, pw - 1);
                const float a0 = c_cubic[phase][0], a1 = c_cubic[phase][1], a2 = c_cubic[phase][2], a3 = c_cubic[phase][3];
                for (int r = 0; r < ph; ++r) {
                    const float* prow = plane + (y_min + r) * w + x_min;
                    float v = __fmul_rn(prow[t0], a0);
                    v = __fadd_rn(v, __fmul_rn(prow[t1], a1));
                    v = __fadd_rn(v, __fmul_rn(prow[t2], a2));
                    v = __fadd_rn(v, __fmul_rn(prow[t3], a3));
                    hor[r * kHorStride + X] = v;
                }
            }
        }
        __syncwarp();
        // vertical pass + first arg-max (row-major index Y*W8 + X; each lane visits its indices in increasing order)
        float best = -INFINITY;
        int best_idx = 0x7fffffff;
        for (int Y = 0; Y < H8; ++Y) {
            const int phase = Y & 7;
            const int sy = (Y >> 3) + (phase < 4 ? -2 : -1);
            const float* h0 = hor + clampi(sy, 0, ph - 1) * kHorStride;
            const float* h1 = hor + clampi(sy + 1, 0, ph - 1) * kHorStride;
            const float* h2 = hor + clampi(sy + 2, 0, ph - 1) * kHorStride;
            const float* h3 = hor + clampi(sy + 3, 0, ph - 1) * kHorStride;
            const float b0 = c_cubic[phase][0], b1 = c_cubic[phase][1], b2 = c_cubic[phase][2], b3 = c_cubic[phase][3];
#pragma unroll
            for (int xi = 0; xi < 2; ++xi) {
                const int X = lane + 32 * xi;
                if (X < W8) {
                    float v = __fmul_rn(h3[X], b3);
                    v = __fadd_rn(__fmul_rn(h2[X], b2), v);
                    v = __fadd_rn(__fmul_rn(h1[X], b1), v);
                    v = __fadd_rn(__fmul_rn(h0[X], b0), v);
                    if (v > best) { best = v; best_idx = Y * W8 + X; }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
            if (ov > best || (ov == best && oi < best_idx)) { best = ov; best_idx = oi; }
        }
        if (lane == 0) {
            const int ay = best_idx / W8, ax = best_idx - ay * W8;
            px[pk] = 8 * x_min + ax;     // == (x+0.5)*8-0.5 + (ax - ((x-x_min+0.5)*8-0.5)), paf_to_pose.py:126-139
            py[pk] = 8 * y_min + ay;
            ps[pk] = best;
        }
    }
}

// ------------------------------------------------------------------ limbs
__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* scratch /*[threads/32 + 1]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        int w = lane < nw ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) scratch[lane] = winc - w;
        if (lane == nw - 1) scratch[nw] = winc;
    }
    __syncthreads();
    const int res = scratch[warp] + inc - v;
    *total = scratch[blockDim.x >> 5];
    __syncthreads();
    return res;
}

// Exact std::sort (libstdc++ introsort) of `n` keys by the whole block, hierarchical:
//   level G (keys in global memory, n > kSmemRange): ranges are partitioned one at a time by all threads with the
//            rank-based exact partition (post_core.h bp_*), scratch (rank -> position) in global memory;
//   level S: every range of <= kSmemRange keys is copied into shared memory and sorted there completely -
//            block-level partitions down to <= kWarpRange keys, then a work queue of ranges, one warp per range with
//            the chunked exact partition (warp_partition), parts of <= 16 keys by a rank-based stable leaf sort -
//            and copied back.
// Together this is exactly __introsort_loop + __final_insertion_sort, including the order of equal keys.
constexpr int kSortQ = 128, kSortLocal = 40, kLimbWarps = kLimbThreads / 32;
constexpr int kSmemRange = B2P_SMEM_RANGE, kWarpRange = 512, kBigStack = 80;
// Per-warp rank tables of the warp phase live in the (then idle) block-phase scratch of 2 x (kSmemRange + 2) int32:
// two uint16 tables of kWarpRankRange entries per warp.  512 for the default 4096-key shared ranges; smaller shared
// ranges (tools/variants.py) shrink the tables, longer ranges then take the chunked warp_partition.
constexpr int kScratchBytes = 2 * (kSmemRange + 2) * (int)sizeof(int32_t);
constexpr int kWarpRankRange = kScratchBytes / (kLimbWarps * 4) >= 512 ? 512 : (kScratchBytes / (kLimbWarps * 4) >= 256 ? 256 : 128);
static_assert(kLimbWarps * 2 * kWarpRankRange * (int)sizeof(uint16_t) <= kScratchBytes, "warp rank tables do not fit the scratch");
#ifndef B2P_SEQ_RANGE
#define B2P_SEQ_RANGE 16
#endif
constexpr int kSeqRange = B2P_SEQ_RANGE;     // parts of <= kSeqRange keys are batched per warp and sorted one per lane (16 = leaves only
                                             // measured best: 2249 vs 2135 frames/s with 64; lane divergence eats larger values)
struct SortShared {
    int lock, top, pending;
    int sf[kSortQ], sl[kSortQ], sd[kSortQ];
    int ltop[kLimbWarps];
    int lf[kLimbWarps][kSortLocal], ll[kLimbWarps][kSortLocal], ld[kLimbWarps][kSortLocal];
    BlockPartState bp;
    int kind, cur_f, cur_l, cur_d;
    int s_top, s_f[kBigStack], s_l[kBigStack], s_d[kBigStack];    // level S block-phase stack
    int g_top, g_f[kBigStack], g_l[kBigStack], g_d[kBigStack];    // level G stack
    int scan[kLimbThreads / 32 + 1];
    unsigned long long scan2[kLimbThreads / 32 + 1];
    int ksum;
    unsigned char wscr[kLimbWarps][64];     // rank -> lane tables of warp_partition
    uint16_t* wtab;                         // kLimbWarps x 2 x kWarpRankRange uint16: per-warp rank -> position tables
    unsigned long long* dbg;                // optional diagnostics counters
    // per-warp batch of small ranges (<= kSeqRange keys): sorted one range per lane by seq_sort_range()
    int bn[kLimbWarps];
    int bsf[kLimbWarps][32], bsl[kLimbWarps][32], bsd[kLimbWarps][32];
};

// Exact partition of v[f, l) with the rank-based formulation of the Hoare loop (post_core.h: lo-stop #k from the left
// pairs with hi-stop #k from the right while posA[k] < posB[k]; cut = min(posA[K+1], posB[K])), evaluated ROW-WISE:
// the range is cut into one contiguous segment per cooperating warp, inside a segment the 32 lanes take consecutive
// keys (coalesced in global memory, conflict-free in shared memory) and the ranks come from ballots + running counts.
// kWarps = 1: a single warp (no block barriers); kWarps = kLimbWarps: the whole block.  Tables hold positions relative
// to f (uint16 for shared-memory ranges, int32 for global ones), 1-based ranks.
template <class PosT, int kWarps>
__device__ int rank_partition(uint64_t* v, int f, int l, PosT* tabA, PosT* tabB, SortShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int wq = (kWarps == 1) ? 0 : (tid >> 5);
    const uint32_t lt = (1u << lane) - 1u, gt = ~lt & ~(1u << lane);
    if ((kWarps == 1 ? lane : tid) == 0) {       // __move_median_to_first(first, first+1, mid, last-1)
        const long a = f + 1, b = f + (l - f) / 2, c = l - 1;
        long m;
        if (B2P_COMP(v[a], v[b])) m = B2P_COMP(v[b], v[c]) ? b : (B2P_COMP(v[a], v[c]) ? c : a);
        else m = B2P_COMP(v[a], v[c]) ? a : (B2P_COMP(v[b], v[c]) ? c : b);
        const uint64_t t = v[f]; v[f] = v[m]; v[m] = t;
        if (kWarps > 1) sh.ksum = 0;
    }
    if (kWarps == 1) __syncwarp(); else __syncthreads();
    const uint32_t pivot = (uint32_t)(v[f] >> 32);
    const int base = f + 1;
    const int rows = (l - base + 31) >> 5;
    const int rpw = (rows + kWarps - 1) / kWarps;             // rows per warp
    const int r0 = min(rows, wq * rpw), r1 = min(rows, r0 + rpw);
    // pass 1: stop counts of this warp's segment
    int cA = 0, cB = 0;
    for (int r = r0; r < r1; ++r) {
        const int p = base + (r << 5) + lane;
        const uint32_t k = p < l ? (uint32_t)(v[p] >> 32) : 0u;
        cA += __popc(__ballot_sync(0xffffffffu, p < l && k >= pivot));
        cB += __popc(__ballot_sync(0xffffffffu, p < l && k <= pivot));
    }
    int offA = 0, offB = 0, totA = cA, totB = cB;
    if (kWarps > 1) {
        if (lane == 0) sh.scan2[wq] = ((unsigned long long)(unsigned)cB << 32) | (unsigned)cA;
        __syncthreads();
        unsigned long long pre = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < kWarps; ++k) {
            const unsigned long long c = sh.scan2[k];
            if (k < wq) pre += c;
            tot += c;
        }
        offA = (int)(unsigned)pre; offB = (int)(pre >> 32);
        totA = (int)(unsigned)tot; totB = (int)(tot >> 32);
    }
    // pass 2: lo-stops, ranked from the left
    int run = offA;
    for (int r = r0; r < r1; ++r) {
        const int p = base + (r << 5) + lane;
        const bool st = p < l && (uint32_t)(v[p] >> 32) >= pivot;
        const uint32_t m = __ballot_sync(0xffffffffu, st);
        if (st) tabA[run + __popc(m & lt) + 1] = (PosT)(p - f);
        run += __popc(m);
    }
    // pass 3: hi-stops, ranked from the right
    run = totB - offB - cB;                                   // hi-stops in the segments to the right
    for (int r = r1 - 1; r >= r0; --r) {
        const int p = base + (r << 5) + lane;
        const bool st = p < l && (uint32_t)(v[p] >> 32) <= pivot;
        const uint32_t m = __ballot_sync(0xffffffffu, st);
        if (st) tabB[run + __popc(m & gt) + 1] = (PosT)(p - f);
        run += __popc(m);
    }
    if (kWarps == 1) __syncwarp(); else __syncthreads();
    // K = number of leading ranks with posA[k] < posB[k] (monotone)
    const int lim = totA < totB ? totA : totB;
    const int nthr = kWarps * 32, me = (kWarps == 1) ? lane : tid;
    int c = 0;
    for (int k = 1 + me; k <= lim; k += nthr) c += (tabA[k] < tabB[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    int K = c;
    if (kWarps > 1) {
        if (lane == 0 && c) atomicAdd(&sh.ksum, c);
        __syncthreads();
        K = sh.ksum;
    }
    for (int k = 1 + me; k <= K; k += nthr) {
        const int ia = f + tabA[k], ib = f + tabB[k];
        const uint64_t t = v[ia]; v[ia] = v[ib]; v[ib] = t;
    }
    const int a_next = (K + 1 <= totA) ? f + (int)tabA[K + 1] : l;
    const int cut = (K > 0 && f + (int)tabB[K] < a_next) ? f + (int)tabB[K] : a_next;
    if (kWarps == 1) __syncwarp(); else __syncthreads();
    return cut;
}

// Small ranges (<= kSeqRange keys) are not worth a warp-cooperative partition (a 32-key partition costs as many
// instructions as a 256-key one): the warp collects them and sorts 32 of them at once, one range per lane, with the
// sequential exact routine.
__device__ __forceinline__ void small_flush(SortShared& sh, uint64_t* v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = __shfl_sync(0xffffffffu, sh.bn[warp], 0);
    if (lane < n) seq_sort_range(v, sh.bsf[warp][lane], sh.bsl[warp][lane], sh.bsd[warp][lane]);
    __syncwarp();
    if (lane == 0) sh.bn[warp] = 0;
    __syncwarp();
}
__device__ __forceinline__ void small_add(SortShared& sh, uint64_t* v, int f, int l, int d) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = sh.bn[warp];
    __syncwarp();       // every lane holds the OLD count before lane 0 bumps it (lanes are not in lock step: a late reader
                        // would otherwise disagree on "batch full" and diverge around the barriers of small_flush)
    if (lane == 0) { sh.bsf[warp][n] = f; sh.bsl[warp][n] = l; sh.bsd[warp][n] = d; sh.bn[warp] = n + 1; }
    __syncwarp();
    if (n + 1 == 32) small_flush(sh, v);
}

// Left-descending introsort loop of one warp on v[f, l): partitions, hands the right parts to the shared stack (or its
// own local stack), batches parts of <= kSeqRange keys for the per-lane sequential sort.
__device__ void warp_descend(uint64_t* v, int f, int l, int d, SortShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (l - f > kSeqRange) {
        if (d == 0) {                      // depth limit exhausted: std::__partial_sort == heap sort
            if (lane == 0) seq_heap_sort(v, f, l);
            __syncwarp();
            return;
        }
        --d;
        const int cut = (l - f <= kWarpRankRange)
                            ? rank_partition<uint16_t, 1>(v, f, l, sh.wtab + warp * 2 * kWarpRankRange,
                                                          sh.wtab + warp * 2 * kWarpRankRange + kWarpRankRange, sh)
                            : (int)warp_partition(v, f, l, sh.wscr[warp]);
        if (l - cut > kSeqRange) {
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&sh.pending, 1);
                bool pushed = false;
                while (atomicCAS(&sh.lock, 0, 1) != 0) {}
                if (sh.top < kSortQ) { const int t = sh.top++; sh.sf[t] = cut; sh.sl[t] = l; sh.sd[t] = d; pushed = true; }
                __threadfence_block();
                atomicExch(&sh.lock, 0);
                if (!pushed) {
                    const int t = sh.ltop[warp];
                    if (t < kSortLocal) { sh.lf[warp][t] = cut; sh.ll[warp][t] = l; sh.ld[warp][t] = d; sh.ltop[warp] = t + 1; }
                    else { seq_sort_range(v, cut, l, d); atomicSub(&sh.pending, 1); }   // unreachable (depth bound)
                }
            }
        } else if (l - cut > 1) {
            small_add(sh, v, cut, l, d);
        }
        l = cut;
        __syncwarp();
    }
    if (l - f > 1) small_add(sh, v, f, l, d);
    __syncwarp();
}

// Level S: complete exact sort of w[f, l) (w = shared-memory resident keys, indices as given) with depth budget d.
// scrA/scrB: rank->position scratch for block partitions (may be null when l - f <= kWarpRange is guaranteed... it is
// only dereferenced for ranges > kWarpRange).
__device__ void smem_sort_range(uint64_t* w, int f, int l, int d, SortShared& sh, int32_t* scrA, int32_t* scrB,
                                unsigned long long* dbg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t0 = clock64();
    if (tid == 0) {
        sh.dbg = dbg;
        sh.lock = 0; sh.top = 0; sh.pending = 0; sh.s_top = 1; sh.s_f[0] = f; sh.s_l[0] = l; sh.s_d[0] = d;
        sh.wtab = reinterpret_cast<uint16_t*>(scrA);     // the block-phase scratch is idle during the warp phase
    }
    if (tid < kLimbWarps) { sh.ltop[tid] = 0; sh.bn[tid] = 0; }
    __syncthreads();
    // block phase: ranges > kWarpRange
    for (;;) {
        if (tid == 0) {
            if (sh.s_top == 0) sh.kind = 0;
            else {
                const int t = --sh.s_top;
                const int rf = sh.s_f[t], rl = sh.s_l[t], rd = sh.s_d[t];
                if (rl - rf <= 16) { leaf_insertion_sort(w, rf, rl); sh.kind = 2; }
                else if ((rl - rf <= kWarpRange || scrA == nullptr) && sh.top < kSortQ) {
                    const int q = sh.top++;
                    sh.sf[q] = rf; sh.sl[q] = rl; sh.sd[q] = rd; sh.pending += 1; sh.kind = 2;
                } else if (rd == 0) { seq_heap_sort(w, rf, rl); sh.kind = 2; }
                else if (scrA == nullptr) { seq_std_sort(w + rf, rl - rf); sh.kind = 2; }
                else { sh.cur_f = rf; sh.cur_l = rl; sh.cur_d = rd - 1; sh.kind = 1; }
            }
        }
        __syncthreads();
        const int kind = sh.kind;
        if (kind == 0) break;
        if (kind == 2) { __syncthreads(); continue; }
        const int cf = sh.cur_f, cl = sh.cur_l, cd = sh.cur_d;
        const int cut = rank_partition<int32_t, kLimbWarps>(w, cf, cl, scrA, scrB, sh);
        if (tid == 0) {
            int t = sh.s_top;
            sh.s_f[t] = cf; sh.s_l[t] = cut; sh.s_d[t] = cd; ++t;
            sh.s_f[t] = cut; sh.s_l[t] = cl; sh.s_d[t] = cd; ++t;
            sh.s_top = t;      // depth-first: <= 2 + depth entries, depth <= 2 log2(n) - log2(n / kSmemRange) ...
        }
        __syncthreads();
    }
    __syncthreads();
    const long long t1 = clock64();
    // warp phase: work queue
    unsigned idle = 0;
    for (;;) {
        int rf = 0, rl = 0, rd = 0, state = 0;    // state: 0 nothing yet, 1 got a range, 2 all done
        if (lane == 0) {
            if (sh.ltop[warp] > 0) {
                const int t = --sh.ltop[warp];
                rf = sh.lf[warp][t]; rl = sh.ll[warp][t]; rd = sh.ld[warp][t]; state = 1;
            } else {
                while (atomicCAS(&sh.lock, 0, 1) != 0) {}
                if (sh.top > 0) {
                    const int t = --sh.top;
                    rf = sh.sf[t]; rl = sh.sl[t]; rd = sh.sd[t]; state = 1;
                }
                __threadfence_block();
                atomicExch(&sh.lock, 0);
                if (state == 0 && atomicAdd(&sh.pending, 0) == 0) state = 2;
            }
        }
        state = __shfl_sync(0xffffffffu, state, 0);
        {
            const int have = __shfl_sync(0xffffffffu, sh.bn[warp], 0);   // one lane's view, broadcast
            if (state != 1 && have > 0) small_flush(sh, w);            // nothing else to do: sort the batched small ranges
        }
        if (state == 2) break;
        if (state == 0) {
            if (sh.dbg && lane == 0) atomicAdd(sh.dbg + 14, 1ull);
            if (++idle > (1u << 24)) { printf("[b200pose] exact sort: idle watchdog (block %d,%d)\n", (int)blockIdx.x, (int)blockIdx.y); __trap(); }
            __nanosleep(1000);
            continue;
        }
        idle = 0;
        rf = __shfl_sync(0xffffffffu, rf, 0);
        rl = __shfl_sync(0xffffffffu, rl, 0);
        rd = __shfl_sync(0xffffffffu, rd, 0);
        __threadfence_block();                 // see the swaps of the warp that published this range
        const long long td0 = clock64();
        warp_descend(w, rf, rl, rd, sh);
        if (sh.dbg && lane == 0) atomicAdd(sh.dbg + 12, (unsigned long long)(clock64() - td0));
        if (lane == 0) { __threadfence_block(); atomicSub(&sh.pending, 1); }
    }
    __syncthreads();
    if (dbg && tid == 0) { atomicAdd(dbg + 6, (unsigned long long)(t1 - t0)); atomicAdd(dbg + 7, (unsigned long long)(clock64() - t1)); }
}

// keys: n keys in generation order, either already in shared memory (n <= kSmemRange: smem_keys == keys) or in global
// memory with `gA/gB` scratch of n + 2 entries each.  smem_keys / smem_scr: shared buffers of kSmemRange keys and
// 2 x (kSmemRange + 2) ints.
__device__ void block_exact_sort(uint64_t* keys, int n, SortShared& sh, uint64_t* smem_keys, int32_t* smem_scr,
                                 int32_t* gA, int32_t* gB, unsigned long long* dbg) {
    const int tid = threadIdx.x;
    if (n <= 1) return;
    int lg = 0;
    for (int m = n; m > 1; m >>= 1) ++lg;
    int32_t* sA = smem_scr;
    int32_t* sB = smem_scr + kSmemRange + 2;
    if (keys == smem_keys) {
        smem_sort_range(keys, 0, n, 2 * lg, sh, sA, sB, dbg);
        return;
    }
    if (tid == 0) { sh.g_top = 1; sh.g_f[0] = 0; sh.g_l[0] = n; sh.g_d[0] = 2 * lg; }
    __syncthreads();
    for (;;) {
        __syncthreads();
        if (sh.g_top == 0) break;
        const int t = sh.g_top - 1;
        const int f = sh.g_f[t], l = sh.g_l[t], d = sh.g_d[t];
        __syncthreads();
        if (tid == 0) sh.g_top = t;
        if (l - f <= kSmemRange) {
            // level S: copy in, sort completely in shared memory, copy out
            for (int i = tid; i < l - f; i += kLimbThreads) smem_keys[i] = keys[f + i];
            __syncthreads();
            smem_sort_range(smem_keys - f, f, l, d, sh, sA, sB, dbg);
            for (int i = tid; i < l - f; i += kLimbThreads) keys[f + i] = smem_keys[i];
            __syncthreads();
        } else if (d == 0) {
            if (tid == 0) seq_heap_sort(keys, f, l);
            __syncthreads();
        } else {
            const long long tg = clock64();
            const int cut = rank_partition<int32_t, kLimbWarps>(keys, f, l, gA, gB, sh);
            if (dbg && tid == 0) { atomicAdd(dbg + 5, (unsigned long long)(clock64() - tg)); atomicAdd(dbg + 8, 1ull); }
            if (tid == 0) {
                int q = sh.g_top;
                sh.g_f[q] = f; sh.g_l[q] = cut; sh.g_d[q] = d - 1; ++q;
                sh.g_f[q] = cut; sh.g_l[q] = l; sh.g_d[q] = d - 1; ++q;
                sh.g_top = q;
            }
        }
    }
    __syncthreads();
}

// Greedy one-to-one assignment (pafprocess.cpp:98-124).  The sorted list is walked in segments of kSmemRange
// candidates: all threads first drop the candidates whose end points were already taken by earlier segments (ordered
// compaction into shared memory), then one warp runs the sequential rule over the survivors only, 32 per step, resolving
// conflicts inside a chunk in candidate order.  Identical to the sequential loop: a candidate rejected by the pre-filter
// would be rejected sequentially too, survivors are examined in order against the live used-sets.
__device__ int greedy_warp_chunked(const uint64_t* keys, int n, int nb, uint32_t* used_a, uint32_t* used_b, int max_conn,
                                   int nc, int* conn_a, int* conn_b, float* conn_s) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < n && nc < max_conn; base += 32) {
        const int i = base + lane;
        uint64_t k = 0;
        int a = -1, b = -1;
        bool free_ = false;
        if (i < n) {
            k = keys[i];
            const uint32_t pair = (uint32_t)k;
            a = pair / nb;
            b = pair - a * nb;
            free_ = !((used_a[a >> 5] >> (a & 31)) & 1u) && !((used_b[b >> 5] >> (b & 31)) & 1u);
        }
        uint32_t active = __ballot_sync(0xffffffffu, free_);
        while (active && nc < max_conn) {
            const int leader = __ffs(active) - 1;
            const int la = __shfl_sync(0xffffffffu, a, leader), lb = __shfl_sync(0xffffffffu, b, leader);
            if (lane == leader) {
                used_a[a >> 5] |= 1u << (a & 31);
                used_b[b >> 5] |= 1u << (b & 31);
                conn_a[nc] = a;
                conn_b[nc] = b;
                conn_s[nc] = key_score(k);
            }
            ++nc;
            active &= ~__ballot_sync(0xffffffffu, a == la || b == lb);
        }
        __syncwarp();
    }
    return nc;
}

constexpr int kGreedyPer = kSmemRange / kLimbThreads;    // candidates per thread and segment

__device__ int greedy_segmented(const uint64_t* keys, int n, int nb, uint32_t* used_a, uint32_t* used_b, int max_conn,
                                int* conn_a, int* conn_b, float* conn_s, uint64_t* seg /*smem, kSmemRange*/,
                                int* scan_scratch, int* s_nc) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_nc = 0;
    __syncthreads();
    for (int s0 = 0; s0 < n; s0 += kSmemRange) {
        if (*s_nc >= max_conn) break;
        uint64_t kk[kGreedyPer];
        int keep = 0, cnt = 0;
#pragma unroll
        for (int j = 0; j < kGreedyPer; ++j) {
            const int i = s0 + tid * kGreedyPer + j;
            kk[j] = 0;
            if (i < n) {
                kk[j] = keys[i];
                const uint32_t pair = (uint32_t)kk[j];
                const int a = pair / nb, b = pair - a * nb;
                if (!((used_a[a >> 5] >> (a & 31)) & 1u) && !((used_b[b >> 5] >> (b & 31)) & 1u)) { keep |= 1 << j; ++cnt; }
            }
        }
        int m;
        int off = block_exclusive_scan(cnt, &m, scan_scratch);     // (barriers inside: every key of the segment is in registers now)
#pragma unroll
        for (int j = 0; j < kGreedyPer; ++j)
            if ((keep >> j) & 1) seg[off++] = kk[j];
        __syncthreads();
        if (tid < 32) {
            const int nc = greedy_warp_chunked(seg, m, nb, used_a, used_b, max_conn, *s_nc, conn_a, conn_b, conn_s);
            __syncwarp();
            if (tid == 0) *s_nc = nc;
        }
        __syncthreads();
    }
    return *s_nc;
}

#define B2P_ITEM_DONE return
__global__ void __launch_bounds__(kLimbThreads) limbs_kernel(PostBuffers pb, PafView paf0, long p_img, int h_up, int lw,
                                                             int lh, int paf_in_smem) {
    // dynamic smem: [kSmemRange keys][2 x (kSmemRange + 2) int32 partition scratch][optional 2 PAF planes]
    extern __shared__ unsigned long long sm_keys[];
    int32_t* sm_scr = reinterpret_cast<int32_t*>(sm_keys + kSmemRange);
    __shared__ uint32_t used_a[64], used_b[64];            // peak_cap <= 2048
    __shared__ int scan_scratch[kLimbThreads / 32 + 1];
    __shared__ long s_pool_base;
    __shared__ SortShared s_sort;
    // Longest-job-first: grid = (image, rank); block `rank` of an image takes the limb with the rank-th largest number
    // of (a, b) pairs, so across the whole grid the heavy limbs are scheduled before the light ones (shorter tail).
    const int tid = threadIdx.x;
    const int img = blockIdx.x, rank_y = blockIdx.y;
    int limb = rank_y;
    {
        int my_pairs[kNumLimb];
#pragma unroll
        for (int l = 0; l < kNumLimb; ++l)
            my_pairs[l] = pb.counts[img * kNumPart + c_limb_parts[l][0]] * pb.counts[img * kNumPart + c_limb_parts[l][1]];
        for (int l = 0; l < kNumLimb; ++l) {
            int rank = 0;
            for (int m = 0; m < kNumLimb; ++m)
                rank += (my_pairs[m] > my_pairs[l]) || (my_pairs[m] == my_pairs[l] && m < l);
            if (rank == rank_y) limb = l;
        }
    }
    const int pa = c_limb_parts[limb][0], pbp = c_limb_parts[limb][1];
    const int cap = pb.peak_cap;
    const int na = pb.counts[img * kNumPart + pa], nb = pb.counts[img * kNumPart + pbp];
    int* out_cnt = pb.conn_cnt + img * kNumLimb + limb;
    if (na == 0 || nb == 0) {
        if (tid == 0) *out_cnt = 0;
        B2P_ITEM_DONE;
    }
    const int* ax = pb.peak_x + ((long)img * kNumPart + pa) * cap;
    const int* ay = pb.peak_y + ((long)img * kNumPart + pa) * cap;
    const int* bx = pb.peak_x + ((long)img * kNumPart + pbp) * cap;
    const int* by = pb.peak_y + ((long)img * kNumPart + pbp) * cap;
    PafView paf = paf0;
    paf.base += img * p_img;
    int c1 = c_limb_paf[limb][0], c2 = c_limb_paf[limb][1];
    if (paf_in_smem) {
        // the 10 x na x nb line-integral samples of this limb hit only its two PAF planes (2 x h x w floats):
        // stage them in shared memory once instead of gathering from L2 ~1e6 times
        float* planes = reinterpret_cast<float*>(sm_scr + 2 * (kSmemRange + 2));
        const int hw = lw * lh;
        for (int i = tid; i < 2 * hw; i += kLimbThreads) {
            const int ch = i >= hw, r = i - ch * hw;
            planes[i] = paf.base[(ch ? c2 : c1) * paf.sc + (long)(r / lw) * paf.sy + (long)(r % lw) * paf.sx];
        }
        __syncthreads();
        paf.base = planes; paf.sc = hw; paf.sy = lw; paf.sx = 1;
        c1 = 0; c2 = 1;
    }

    const int npairs = na * nb;
    const long long t_start = clock64();
    // Keys go to shared memory when even the upper bound (all pairs) fits, else to the pool: [npairs key slots]
    // [2 x (npairs + 2) int32 partition scratch].
    unsigned long long* keys = sm_keys;
    int32_t *posA = nullptr, *posB = nullptr;
    if (npairs > kSmemRange) {
        if (tid == 0) s_pool_base = (long)atomicAdd(pb.pool_cursor, (unsigned long long)npairs);
        __syncthreads();
        if (s_pool_base + (long)npairs > pb.pool_cap) {
            if (tid == 0) {
                atomicOr(&pb.status[img], 2);
                *out_cnt = 0;
            }
            B2P_ITEM_DONE;
        }
        keys = pb.pool + s_pool_base;
        __syncthreads();
    }
    // Single pass over the (a, b) pairs in generation order (a-major, b-minor = the order the reference pushes
    // candidates in), 512 pairs per step, ordered compaction by ballot + warp-count prefix.
    int n = 0;
    {
        const int lane = tid & 31, warp = tid >> 5;
        for (int base = 0; base < npairs; base += kLimbThreads) {
            const int p = base + tid;
            float sc = 0.f;
            bool pass = false;
            if (p < npairs) {
                const int a = p / nb, b = p - a * nb;
                pass = pair_score(paf, c1, c2, ax[a], ay[a], bx[b], by[b], h_up, &sc);
            }
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) scan_scratch[warp] = __popc(m);
            __syncthreads();
            int pre = n, tot = 0;
#pragma unroll
            for (int k = 0; k < kLimbThreads / 32; ++k) {
                const int c = scan_scratch[k];
                if (k < warp) pre += c;
                tot += c;
            }
            if (pass) keys[pre + __popc(m & ((1u << lane) - 1u))] = cand_key(sc, (uint32_t)p);
            n += tot;
            __syncthreads();
        }
    }
    if (n == 0) {
        if (tid == 0) *out_cnt = 0;
        B2P_ITEM_DONE;
    }
    if (npairs > kSmemRange && n > kSmemRange) {   // rank -> position scratch of the global-level partitions: n + 2 entries
        if (tid == 0) s_pool_base = (long)atomicAdd(pb.pool_cursor, (unsigned long long)n + 2);
        __syncthreads();
        if (s_pool_base + n + 2 > pb.pool_cap) {
            if (tid == 0) {
                atomicOr(&pb.status[img], 2);
                *out_cnt = 0;
            }
            B2P_ITEM_DONE;
        }
        posA = reinterpret_cast<int32_t*>(pb.pool + s_pool_base);
        posB = posA + n + 2;
    }
    for (int i = tid; i < 64; i += kLimbThreads) { used_a[i] = 0; used_b[i] = 0; }
    __threadfence();
    __syncthreads();
    const long long t_scored = clock64();
    block_exact_sort(reinterpret_cast<uint64_t*>(keys), n, s_sort, reinterpret_cast<uint64_t*>(sm_keys), sm_scr, posA,
                     posB, pb.dbg);     // std::sort, pafprocess.cpp:97
    const long long t_sorted = clock64();
    {
        __shared__ int s_nc;
        const long o = ((long)img * kNumLimb + limb) * cap;
        const int nc = greedy_segmented(reinterpret_cast<const uint64_t*>(keys), n, nb, used_a, used_b, min(na, nb),
                                        pb.conn_a + o, pb.conn_b + o, pb.conn_s + o, reinterpret_cast<uint64_t*>(sm_keys),
                                        scan_scratch, &s_nc);
        if (tid == 0) {
            *out_cnt = nc;
            if (pb.dbg) {   // phase maxima over blocks (cycles): scoring+compaction, sort, greedy; and max candidates
                atomicMax(pb.dbg + 0, (unsigned long long)(t_scored - t_start));
                atomicMax(pb.dbg + 1, (unsigned long long)(t_sorted - t_scored));
                atomicMax(pb.dbg + 2, (unsigned long long)(clock64() - t_sorted));
                atomicMax(pb.dbg + 3, (unsigned long long)n);
                atomicAdd(pb.dbg + 4, (unsigned long long)n);
                atomicAdd(pb.dbg + 9, (unsigned long long)(t_sorted - t_scored));
                atomicAdd(pb.dbg + 10, 1ull);
            }
        }
    }
}

// ------------------------------------------------------------------ assembly
__global__ void __launch_bounds__(kAsmThreads) assemble_kernel(PostBuffers pb) {
    extern __shared__ int sm_kept[];     // [human_cap]
    __shared__ int part_base[kNumPart + 1];
    __shared__ int s_nh, s_nrows;
    const int img = blockIdx.x, tid = threadIdx.x;
    const int cap = pb.peak_cap;
    const int id_cap = kNumPart * cap;
    if (tid == 0) {
        part_base[0] = 0;
        for (int p = 0; p < kNumPart; ++p) part_base[p + 1] = part_base[p] + pb.counts[img * kNumPart + p];
    }
    __syncthreads();
    float* id_score = pb.id_score + (long)img * id_cap;
    int* id_xy = pb.id_xy + (long)img * id_cap * 2;
    uint8_t* list_n = pb.list_n + (long)img * id_cap;
    uint8_t* alive = pb.alive + (long)img * pb.row_cap;
    const int total = part_base[kNumPart];
    for (int p = 0; p < kNumPart; ++p) {
        const int n = part_base[p + 1] - part_base[p];
        const long src = ((long)img * kNumPart + p) * cap;
        for (int i = tid; i < n; i += kAsmThreads) {
            const int id = part_base[p] + i;
            id_score[id] = pb.peak_s[src + i];
            id_xy[2 * id] = pb.peak_x[src + i];
            id_xy[2 * id + 1] = pb.peak_y[src + i];
        }
    }
    for (int i = tid; i < total; i += kAsmThreads) list_n[i] = 0;
    for (int i = tid; i < pb.row_cap; i += kAsmThreads) alive[i] = 0;
    __syncthreads();

    float* rows = pb.rows + (long)img * pb.row_cap * kRowFloats;
    __shared__ Assembler s_as;
    if (tid == 0) {
        s_as.rows = rows;
        s_as.alive = alive;
        s_as.lists = pb.lists + (long)img * id_cap * kListCap;
        s_as.list_n = list_n;
        s_as.part_base = part_base;
        s_as.peak_score = id_score;
        s_as.row_cap = pb.row_cap;
        s_as.nrows = 0;
        s_as.degraded = 0;
        s_as.overflow = 0;
    }
    __syncthreads();
    // The assembly itself is inherently sequential (pafprocess.cpp:127-185) and latency bound: one thread walks the
    // connections.  Before each limb all threads pull the lines that walk will touch (row lists of both end points,
    // the listed rows, peak scores) into L1 so the sequential part runs on L1 hits.
    for (int l = 0; l < kNumLimb; ++l) {
        const int p1 = c_limb_parts[l][0], p2 = c_limb_parts[l][1];
        const int nc = pb.conn_cnt[img * kNumLimb + l];
        const long o = ((long)img * kNumLimb + l) * cap;
        for (int c = tid; c < nc; c += kAsmThreads) {
            const int ids[2] = {part_base[p1] + pb.conn_a[o + c], part_base[p2] + pb.conn_b[o + c]};
            prefetch_l1(pb.conn_s + o + c);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int id = ids[e];
                prefetch_l1(id_score + id);
                const int n = s_as.list_n[id];
                for (int i = 0; i < n && i < kListCap; ++i) {
                    const int r = s_as.lists[id * kListCap + i];
                    prefetch_l1(rows + (long)r * kRowFloats);
                    prefetch_l1(rows + (long)r * kRowFloats + kRowFloats - 1);
                    prefetch_l1(alive + r);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            for (int c = 0; c < nc; ++c)
                s_as.add_connection(l, p1, p2, part_base[p1] + pb.conn_a[o + c], part_base[p2] + pb.conn_b[o + c],
                                    pb.conn_s[o + c]);
        }
        __syncthreads();
    }
    if (tid == 0) {
        Assembler& as = s_as;
        int nh = 0, st = 0;
        for (int r = 0; r < as.nrows; ++r)
            if (as.keep(r)) {
                if (nh < pb.human_cap) sm_kept[nh++] = r;
                else st |= 8;
            }
        if (as.overflow) st |= 4;
        if (as.degraded) st |= 16;
        // this block is the last writer of the image's status word: fold it into the sticky accumulator as well
        const int st_all = (st ? atomicOr(&pb.status[img], st) : pb.status[img]) | st;
        if (st_all) atomicOr(&pb.status_acc[img], st_all);
        s_nh = nh;
        s_nrows = as.nrows;
        pb.n_humans[img] = nh;
    }
    __syncthreads();
    const int nh = s_nh;
    float* out = pb.humans + (long)img * pb.human_cap * kHumanFloats;
    for (int e = tid; e < nh * kHumanFloats; e += kAsmThreads) {
        const int hi = e / kHumanFloats, f = e - hi * kHumanFloats;
        const float* row = rows + sm_kept[hi] * kRowFloats;
        float v;
        if (f == 0) v = __fdiv_rn(row[18], row[19]);     // get_score, pafprocess.cpp:204-206
        else {
            const int p = (f - 1) >> 2, k = (f - 1) & 3;
            const int cid = (int)row[p];                  // get_part_cid truncation, pafprocess.cpp:200-202
            if (cid < 0) v = (k == 3) ? -1.f : 0.f;
            else v = k == 0 ? (float)id_xy[2 * cid] : k == 1 ? (float)id_xy[2 * cid + 1] : k == 2 ? id_score[cid] : (float)cid;
        }
        out[e] = v;
    }
}

}  // namespace

#define B2P_TRY(x)                        \
    do {                                  \
        cudaError_t e_ = (x);             \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

cudaError_t post_alloc(PostBuffers& pb, int batch_cap, int peak_cap, int human_cap, long pool_cap) {
    memset(&pb, 0, sizeof(pb));
    if (peak_cap > 2048 || peak_cap < 1) return cudaErrorInvalidValue;
    pb.batch_cap = batch_cap;
    pb.peak_cap = peak_cap;
    pb.human_cap = human_cap;
    pb.cand_smem_cap = B2P_SMEM_RANGE;
    pb.pool_cap = pool_cap;
    pb.row_cap = 4 * peak_cap;
    const long B = batch_cap;
    B2P_TRY(cudaMalloc(&pb.counts, B * kNumPart * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.peak_x, B * kNumPart * peak_cap * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.peak_y, B * kNumPart * peak_cap * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.peak_s, B * kNumPart * peak_cap * sizeof(float)));
    B2P_TRY(cudaMalloc(&pb.conn_cnt, B * kNumLimb * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.conn_a, B * kNumLimb * peak_cap * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.conn_b, B * kNumLimb * peak_cap * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.conn_s, B * kNumLimb * peak_cap * sizeof(float)));
    B2P_TRY(cudaMalloc(&pb.rows, B * pb.row_cap * kRowFloats * sizeof(float)));
    B2P_TRY(cudaMalloc(&pb.alive, B * pb.row_cap));
    B2P_TRY(cudaMalloc(&pb.lists, B * kNumPart * peak_cap * kListCap * sizeof(int32_t)));
    B2P_TRY(cudaMalloc(&pb.list_n, B * kNumPart * peak_cap));
    B2P_TRY(cudaMalloc(&pb.id_score, B * kNumPart * peak_cap * sizeof(float)));
    B2P_TRY(cudaMalloc(&pb.id_xy, B * kNumPart * peak_cap * 2 * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.pool, pool_cap * sizeof(unsigned long long)));
    B2P_TRY(cudaMalloc(&pb.pool_cursor, sizeof(unsigned long long)));
    B2P_TRY(cudaMalloc(&pb.n_humans, B * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.humans, B * human_cap * kHumanFloats * sizeof(float)));
    B2P_TRY(cudaMalloc(&pb.status, B * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.status_acc, B * sizeof(int)));
    B2P_TRY(cudaMemset(pb.status_acc, 0, B * sizeof(int)));
    B2P_TRY(cudaMalloc(&pb.dbg, 16 * sizeof(unsigned long long)));
    B2P_TRY(cudaMemset(pb.dbg, 0, 16 * sizeof(unsigned long long)));
    B2P_TRY(cudaMemset(pb.status, 0, B * sizeof(int)));
    B2P_TRY(cudaMemset(pb.counts, 0, B * kNumPart * sizeof(int)));
    return cudaSuccess;
}

void post_free(PostBuffers& pb) {
    void* ptrs[] = {pb.counts, pb.peak_x, pb.peak_y, pb.peak_s,  pb.conn_cnt, pb.conn_a,      pb.conn_b,
                    pb.conn_s, pb.rows,   pb.alive,  pb.lists,   pb.list_n,   pb.id_score,    pb.id_xy,
                    pb.pool,   pb.pool_cursor, pb.n_humans, pb.humans, pb.status, pb.status_acc, pb.dbg};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    memset(&pb, 0, sizeof(pb));
}

cudaError_t post_peaks(const PostBuffers& pb, int batch, const float* heat, long h_img, long h_ch, long h_y, long h_x,
                       int h, int w, float thresh, cudaStream_t s) {
    if (batch > pb.batch_cap) return cudaErrorInvalidValue;
    const size_t smem = ((size_t)h * w + (kPeakThreads / 32) * 5 * kHorStride) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static DynSmemOptIn optin;
    B2P_TRY(optin.ensure(peaks_kernel, smem));
    B2P_TRY(cudaMemsetAsync(pb.status, 0, batch * sizeof(int), s));
    peaks_kernel<<<dim3(kNumPart, batch), kPeakThreads, smem, s>>>(pb, heat, h_img, h_ch, h_y, h_x, h, w, thresh);
    return cudaGetLastError();
}

cudaError_t post_limbs(const PostBuffers& pb, int batch, const float* paf, long p_img, long p_ch, long p_y,
                                    long p_x, int shift, int h_up, int lw, int lh, cudaStream_t s) {
    if (batch > pb.batch_cap) return cudaErrorInvalidValue;
    B2P_TRY(cudaMemsetAsync(pb.pool_cursor, 0, sizeof(unsigned long long), s));
    PafView pv{paf, p_ch, p_y, p_x, shift};
    if (pb.cand_smem_cap != kSmemRange) return cudaErrorInvalidValue;
    size_t smem = kSmemRange * sizeof(unsigned long long) + 2 * (kSmemRange + 2) * sizeof(int32_t);
    int in_smem = 0;
    if (shift == 3 && (size_t)2 * lw * lh * sizeof(float) <= 96 * 1024) {
        in_smem = 1;
        smem += (size_t)2 * lw * lh * sizeof(float);
    }
    static DynSmemOptIn optin;
    B2P_TRY(optin.ensure(limbs_kernel, smem));
    limbs_kernel<<<dim3(batch, kNumLimb), kLimbThreads, smem, s>>>(pb, pv, p_img, h_up, lw, lh, in_smem);
    return cudaGetLastError();
}

cudaError_t post_assemble(const PostBuffers& pb, int batch, cudaStream_t s) {
    if (batch > pb.batch_cap) return cudaErrorInvalidValue;
    assemble_kernel<<<batch, kAsmThreads, pb.human_cap * sizeof(int), s>>>(pb);
    return cudaGetLastError();
}

}  // namespace b2p
