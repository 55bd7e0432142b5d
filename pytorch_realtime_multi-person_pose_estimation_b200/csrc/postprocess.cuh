// Fused post-processing kernel family (device-resident replacement of NMS + pafprocess):
//   peaks_kernel    : find_peaks + NMS refinement      /root/reference/lib/utils/paf_to_pose.py:25-38, 67-145
//   limbs_kernel    : candidate scoring + greedy match  /root/reference/lib/pafprocess/pafprocess.cpp:47-125
//   assemble_kernel : person assembly + prune + getters /root/reference/lib/pafprocess/pafprocess.cpp:127-218
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "post_core.h"

namespace b2p {

struct PostBuffers {
    // capacities
    int batch_cap, peak_cap /*per part*/, human_cap, cand_smem_cap;
    long pool_cap;   // candidate key pool (entries) for limbs whose candidate list does not fit shared memory
    // peaks (per image, per part)
    int* counts;        // [B][18]
    int* peak_x;        // [B][18][peak_cap]  full-resolution integer coordinates
    int* peak_y;
    float* peak_s;      // refined score
    // connections (per image, per limb)
    int* conn_cnt;      // [B][19]
    int* conn_a;        // [B][19][peak_cap]
    int* conn_b;
    float* conn_s;
    // assembly scratch
    float* rows;        // [B][row_cap][20]
    uint8_t* alive;     // [B][row_cap]
    int32_t* lists;     // [B][18*peak_cap][kListCap]
    uint8_t* list_n;    // [B][18*peak_cap]
    float* id_score;    // [B][18*peak_cap]   peak score by id
    int* id_xy;         // [B][18*peak_cap][2]
    int row_cap;
    // candidate pool
    unsigned long long* pool;
    unsigned long long* pool_cursor;
    // results
    int* n_humans;      // [B]
    float* humans;      // [B][human_cap][1 + 18*4]: score, then per part (x, y, peak score, cid or -1)
    unsigned long long* dbg;   // [16] diagnostics: max cycles per limbs_kernel phase, candidate counts
    int* status_acc;    // [B] OR of `status` over every run since the last reset (b200pose_post_status_accum)
    int* status;        // [B] bit0 peak overflow, bit1 candidate pool overflow, bit2 row overflow, bit3 human overflow,
                        //     bit4 assembler used the slow scan, bit8.. number of limbs that needed the tie-exact sort
};

constexpr int kHumanFloats = 1 + 18 * 4;

cudaError_t post_alloc(PostBuffers& pb, int batch_cap, int peak_cap, int human_cap, long pool_cap);
void post_free(PostBuffers& pb);

// heat: fp32, value(img, part, y, x) = heat[img*h_img + part*h_ch + y*h_y + x*h_x]; low-res (h x w).
cudaError_t post_peaks(const PostBuffers& pb, int batch, const float* heat, long h_img, long h_ch, long h_y, long h_x,
                       int h, int w, float thresh, cudaStream_t s);
// paf: fp32 view per image: base + img*p_img, strides (p_ch, p_y, p_x), shift (3: low-res, 0: already upsampled).
cudaError_t post_limbs(const PostBuffers& pb, int batch, const float* paf, long p_img, long p_ch, long p_y,
                                    long p_x, int shift, int h_up, int lw, int lh, cudaStream_t s);
// Person assembly of the connections post_limbs() left in `pb`.  Small footprint (128 threads, human_cap*4 B smem per
// image): it can run on a second stream next to the convolutions of the following batch.
cudaError_t post_assemble(const PostBuffers& pb, int batch, cudaStream_t s);
// lw x lh: dimensions of the low-resolution PAF planes (used to stage them in shared memory when shift == 3)

}  // namespace b2p
