// Kernel-argument structs and launcher prototypes shared between the .cu files and api.cu.
#pragma once
#include "common.cuh"

struct LikArgs {
    int32_t likelihood, B, D;
    SlabMat F;                          // logits / means, ld = D
    const float* Y; int64_t y_cs;       // [B, D] | [B, 1]
    const float* lik_log_var; int64_t h_cs;   // Gaussian only
    float* ll_rows; float* aux_rows;    // [C][B] nullable
    float* ll_sum;                      // [C] nullable
    float* dF; int64_t df_cs;           // [C][B][D] nullable (needs inv_B)
    float* g_lik_log_var; int64_t g_cs; // nullable: dU/d lik_log_var (data term)
    float* probs;                       // softmax only, nullable [C][B][D]
    float* part;                        // [C][64][2] per-CTA partial sums for large batches (nullable: one CTA)
    float inv_B;
};

struct UpdArgs {
    float* theta; float* mom; int64_t cs; int64_t n;
    const float* grad; int64_t grad_cs; int32_t n_part; int64_t part_stride;
    int32_t n_seg;
    float h, hN, beta, noise_scale /* sqrt(2(1-beta)T) */, inv_N;
    int32_t resample;
    uint64_t seed, step; uint32_t stream_base;
    const float* eps_inject; const float* mom_inject;
    const unsigned long long* step_dev;   // nullable: device-resident base added to `step` (CUDA-graph replays draw fresh noise)
};

struct HypArgs {
    int32_t B, d, d_prev, d_x, g, ldx, has_mean;
    SlabMat Fprev; const float* X; int64_t x_cs;
    SlabMat T;      // raw T = dP z^T, ld = d
    SlabMat R;      // rowsum(dP), ld = 1
    SlabMat dF;     // dU/dF_l, ld = g
    SlabMat Fcur;   // F_l, ld = g
    const float* log_inv_ls; int64_t h_cs;
    float* gH; int64_t gh_cs;
    int64_t off_log_amp, off_log_inv_ls, off_mean;
    float* part; int64_t part_cs;     // [C][row blocks][2 d + 1] partial sums (workspace)
    int32_t n_rb, rows_per;           // filled by the launcher
};
int dgprf_hyper_row_blocks(int B);

// all-layer operand prep of the pipelined tensor-core kernels (k_prep_layers.cu)
struct PrepLayer {
    const float* z; int64_t z_cs; const float* log_inv_ls; const float* mean; int64_t h_cs;
    const float* W; int64_t w_cs;
    int32_t d, M, F, g, NG, Kp, has_mean;
    float* zt; float* ot; float* wt; float* wp;          // outputs (the ones whose task count is non-zero)
    int32_t n_zt, n_ot, n_wt, n_wp;                       // task (block) counts
};
struct PrepArgs { int32_t n_layers; PrepLayer L[DGPRF_MAX_LAYERS]; };
int dgprf_launch_prep_layers(const PrepArgs& a, int n_chains, cudaStream_t st);

int dgprf_launch_fwd_simt(const FwdArgs& a, int n_chains, cudaStream_t st);
bool dgprf_fwd_tc2_supported(const FwdArgs& a);
int dgprf_fwd_tc2_col_splits(int tile_cols, int B, int d, int M, int g, int n_chains);
int64_t dgprf_fwd_tc2_zt_floats(int M);
int64_t dgprf_fwd_tc2_at_floats(int B, int d);
int64_t dgprf_fwd_tc2_ot_floats(int M, int d);
int64_t dgprf_fwd_tc2_wt_floats(int F, int g);
int dgprf_launch_fwd_tc2(const FwdArgs& a, int n_chains, cudaStream_t st);
// Tile-blocked layout of the saved features between the pipelined forward and the pipelined backward of a layer:
//   [row block of 128][column tile of 64][cos b0 | cos b1 | sin b0 | sin b1 (arc-cosine: b0 | b1)][128 rows][32 floats]
// so that the four TMA boxes of a (128 x 64) tile are ONE contiguous 64 KB run in HBM instead of 4 x 128 pieces of 128
// bytes a row pitch (32 KB) apart: the same store pattern with no compute measures 6.25 TB/s blocked against 5.1-5.3 TB/s
// row-major (profiles/r02_tma_store_blocked.txt).  The buffer is internal to the workspace; nothing else reads it.
int64_t dgprf_phi_blocked_floats(int B, int M, int kind);
int dgprf_tc_tile_cols(int B, int M, int n_chains);
bool dgprf_bwd_tc2_shape_ok(int M, int g, int d, int d_prev, int CS, int hyper);
int64_t dgprf_bwd_tc2_wp_floats(int F);
int dgprf_bwd_tc2_pick_cs(int B, int M, int g, int d, int d_prev, int RS, int n_chains, int hyper);
bool dgprf_bwd_tc2_supported(const BwdArgs& a);
int dgprf_launch_bwd_tc2(const BwdArgs& a, int n_chains, cudaStream_t st);
int dgprf_launch_bwd_simt(const BwdArgs& a, int n_chains, cudaStream_t st);
int dgprf_launch_loglik(const LikArgs& a, int n_chains, cudaStream_t st);
int dgprf_launch_update(const UpdArgs& a, const dgprf_segment* segs, int n_seg, int n_chains, cudaStream_t st);
int dgprf_launch_sum_slabs(const SlabMat& m, int B, int ncol, float* out, int64_t out_cs, int n_chains, cudaStream_t st);
int dgprf_launch_grad_finalize(const float* part, int64_t part_cs, int64_t part_ss, int n_part,
                               const float* theta, int64_t theta_cs, float inv_N,
                               float* out, int64_t out_cs, int64_t n, int n_chains, cudaStream_t st);
int dgprf_launch_hyper_reduce(const HypArgs& a, int n_chains, cudaStream_t st);

// row-fused step (k9_step_rows.cu)
size_t dgprf_step_rows_smem(const dgprf_model* m);
int dgprf_step_rows_groups(int B);
bool dgprf_step_rows_eligible(const dgprf_model* m, int B);
int dgprf_launch_step_rows(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y, int64_t y_cs, int B,
                           float* gwpart, int64_t gw_cs, int64_t gw_ss, float* ll_part, int64_t ll_cs,
                           const UpdArgs* upd, const dgprf_segment* segs, int n_seg, unsigned int* bar, float* u_out,
                           bool* fused, cudaStream_t st);
// cluster-split tensor-pipe step (k10_step_cluster.cu): 0 tiles = not eligible
int dgprf_step_cluster_tiles(const dgprf_model* m, int B);
int dgprf_launch_step_cluster(const dgprf_model* m, const float* X, int64_t x_cs, const float* Y, int64_t y_cs, int B,
                              float* gwpart, int64_t gw_cs, int64_t gw_ss, float* ll_part, int64_t ll_cs,
                              const UpdArgs* upd, const dgprf_segment* segs, int n_seg, unsigned int* bar, float* u_out,
                              bool* fused, cudaStream_t st);
// lanes per 128-bit vector of the slab-summing update (K5 and the fused updates pick the same summation order)
int dgprf_update_lpv(int n_part, int64_t n4, int n_chains);
int dgprf_launch_sum_rows(const float* in, int64_t in_cs, int n, float* out, int n_chains, cudaStream_t st);
