// Shared device helpers and kernel-argument structs for libdgprf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/dgprf.h"

#define DGPRF_LOG_2PI 1.8378770664093453f

// ---- error plumbing (api.cu owns the thread-local message) -------------------------------
void dgprf_set_error(const char* fmt, ...);
#define DGPRF_CHECK_CUDA(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            dgprf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DGPRF_ECUDA;                                                             \
        }                                                                                   \
    } while (0)
#define DGPRF_REQUIRE(cond, ...)                                                            \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            dgprf_set_error(__VA_ARGS__);                                                   \
            return DGPRF_EINVAL;                                                            \
        }                                                                                   \
    } while (0)

// ---- measurement hook: CUDA events around every kernel launch (dgprf_profile_start/stop) ----
void dgprf_prof_begin(const char* name, cudaStream_t st);
void dgprf_prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    ProfScope(const char* name, cudaStream_t s) : st(s) { dgprf_prof_begin(name, s); }
    ~ProfScope() { dgprf_prof_end(st); }
};

// Opt-in dynamic shared-memory limit of a kernel, remembered per (kernel, device) and only ever raised (api.cu).
int dgprf_ensure_smem(const void* kernel, size_t smem);
// Signature of the process environment (api.cu): key of the per-thread plan caches, so that a debug switch set or cleared
// between two calls is seen by the next call.
uint64_t dgprf_env_signature(void);

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// The kernels of a layered step (operand prep, forward, slab sums, likelihood, backward, finalize, update: ~15-25 dependent
// launches) are launched with programmatic stream serialization: the next grid is scheduled while the previous one still
// runs and blocks in griddepcontrol.wait -- the FIRST instruction of every kernel of this library -- until that grid has
// completed and its writes are visible.  Nothing is ever touched early, so the only overlap is the launch latency and block
// scheduling (2-3 us per dependent launch), which is what a small layered step is made of.  A kernel launched without the
// attribute, or after a kernel that never signals, behaves exactly as before.  DGPRF_NO_PDL=1 switches the attribute off.
__device__ __forceinline__ void dgprf_pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool dgprf_pdl_enabled();       // api.cu (reads the environment through the per-thread signature cache)
template <typename... KArgs, typename... Args>
static inline cudaError_t dgprf_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = dgprf_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int64_t x, int64_t m) { return (int)((x + m - 1) / m); }

// ---- tile geometry of the SIMT (fp32) layer kernels ---------------------------------------
constexpr int kTM = 64;        // batch rows per tile
constexpr int kTN = 64;        // random-feature columns per tile
constexpr int kKC = 32;        // K chunk of GEMM #1 (input width)
constexpr int kMaxCS = 8;      // column splits of the SIMT kernels -> partial slabs of F / dF_prev
constexpr int kMaxSlabs = 16;  // most slabs any producer writes (tensor-core forward with 32-wide tiles)
constexpr int kMaxRS = 18;     // row splits     -> partial slabs of gW
constexpr int kThreads = 256;

static inline int col_splits(int M) { int t = ceil_div(M, kTN); return t < kMaxCS ? t : kMaxCS; }
static inline int row_splits(int B) { int t = ceil_div(B, kTM); return t < kMaxRS ? t : kMaxRS; }
static inline int pad_g(int g) { return g <= 4 ? 4 : g <= 16 ? 16 : g <= 32 ? 32 : g <= 64 ? 64 : -1; }


// A matrix given as a sum of `n_slabs` partial slabs: value(c,row,col) =
// sum_s ptr[c*cs + s*ss + row*ld + col].  n_slabs==1 is a plain dense matrix.
struct SlabMat {
    const float* ptr;
    int64_t cs;      // chain stride
    int64_t ss;      // slab stride
    int32_t ld;      // leading dimension
    int32_t n_slabs;
};

// All slab loads are issued before the (fixed-order) adds so they overlap: n_slabs <= kMaxSlabs.
__device__ __forceinline__ float slab_load(const SlabMat& m, int chain, int64_t row, int col) {
    const float* p = m.ptr + chain * m.cs + row * m.ld + col;
    float t[kMaxSlabs];
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s) t[s] = s < m.n_slabs ? __ldg(p + s * m.ss) : 0.f;
    float v = t[0];
#pragma unroll
    for (int s = 1; s < kMaxSlabs; ++s) v += t[s];
    return v;
}

// slab_load with a runtime loop over the slabs (same summation order, so bit-identical): for kernels whose per-element
// work is small, the 16-way predicated unroll of slab_load costs more issue slots than the loads it overlaps
__device__ __forceinline__ float slab_load_rt(const SlabMat& m, int chain, int64_t row, int col) {
    const float* p = m.ptr + chain * m.cs + row * m.ld + col;
    float v = __ldg(p);
    for (int s = 1; s < m.n_slabs; ++s) v += __ldg(p + s * m.ss);
    return v;
}

struct FwdArgs {
    int32_t kind, B, d_prev, d_x, d, M, g, F, CS, ldx, do_gemm2, has_mean, tile_cols;
    SlabMat Fprev;                       // previous GP-layer output (partial slabs), ld = d_prev
    const float* X;  int64_t x_cs;       // model input, ld = ldx
    const float* z;  int64_t z_cs;       // [d, M]
    const float* log_inv_ls; const float* log_amp; const float* mean; int64_t h_cs;
    const float* W;  int64_t w_cs;       // [F, g]
    float* Phi;      int64_t phi_cs;     // [B, F] nullable
    int32_t phi_blocked;                 // saved features in the tile-blocked layout (dgprf_phi_blocked_floats) instead of [B, F]
    float* Fpart;    int64_t fpart_cs;   // [CS][B][g]
    float* zt; int64_t zt_cs; float* wt; // pipelined TC forward: prepped z^T hi/lo [2][M][128] and W^T [NG][F] (workspace)
    float* at; float* ot;                // ... WIDE variant: input hi/lo [2][B][Kp] and Omega^T hi/lo [2][M][Kp]
    int32_t prepped;                     // zt / ot / wt were already written by k_prep_layers for this step
    float* Fsum; int64_t fsum_cs;        // pipelined TC forward, CS > 1: dense [B][g] sum of the CS slabs, written by the last column
    unsigned int* sum_ctr;               // split of a row block to finish (tickets [chains][row blocks], zero between launches); nullable
};

struct BwdArgs {
    int32_t kind, B, d_prev, d_x, d, M, g, F, CS, RS, ldx, has_mean, hyper;
    SlabMat dF;                          // dU/dF_l  (ld = g)
    const float* Phi; int64_t phi_cs;    // saved features [B, F]
    int32_t phi_blocked;                 // ... or in the tile-blocked layout the pipelined forward wrote
    const float* z;   int64_t z_cs;
    const float* log_inv_ls; const float* log_amp; const float* mean; int64_t h_cs;
    const float* W;   int64_t w_cs;
    float* gWpart;    int64_t gw_cs, gw_ss;   // [RS][w_len] already offset by off_W
    float* Dpart;     int64_t d_cs;           // [CS][B][d_prev]  dU/dF_{l-1} partials (nullable)
    float* Tpart;     int64_t t_cs;           // [CS][B][d]       raw T = dP z^T     (hyper)
    float* Rpart;     int64_t r_cs;           // [CS][B]          R = rowsum(dP)     (hyper | mean)
    float* wp;                                // pipelined TC backward: zero-padded W rows [F][32] (workspace)
    int32_t prepped;                          // wp was already written by k_prep_layers for this step
    float* Dsum; int64_t dsum_cs;             // pipelined TC backward, CS > 1: dense [B][d_prev] sum of the Dpart slabs, written by the last
    unsigned int* sum_ctr;                    // column split of a row tile to finish (tickets [chains][row tiles], zero between launches); nullable
};

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` needs >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = lane < nw ? red[lane] : 0.f;
        r = warp_sum(r);
    }
    return r;
}

// 4-byte async copy global -> shared (LDGSTS); !pred zero-fills the destination.
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// all but the newest committed group of this thread are complete
__device__ __forceinline__ void cp_async_wait_but_newest() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// sin/cos with a 2-constant Cody-Waite reduction to [-pi, pi] followed by the MUFU
// approximations (abs error ~5e-7 on the reduced range).  Random-feature phases reach
// |P| >> pi, where bare __sinf/__cosf lose all accuracy.
__device__ __forceinline__ void sincos_cw(float x, float* s, float* c) {
    // round(x / 2pi) by the 1.5 * 2^23 trick (two FADDs; rintf is an XU-pipe instruction like MUFU); exact for
    // |x / 2pi| < 2^22, far beyond any phase the features can produce before sin/cos lose meaning in fp32
    const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.f), -12582912.f);
    float r = fmaf(-k, 6.2831854820251465f, x);        // 2*pi rounded to fp32
    r = fmaf(-k, -1.7484555e-7f, r);                   // 2*pi - fp32(2*pi)
    *s = __sinf(r);
    *c = __cosf(r);
}
