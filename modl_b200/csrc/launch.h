// launch.h -- host-side launchers shared between the translation units of libmodl_b200.so.
// Each kernel family is compiled in its own .cu (parallel nvcc, shorter rebuilds); api.cu only
// sees these declarations.
#pragma once
#include "common.cuh"

namespace modl {

enum GemmLayout { A_KMAJOR = 0, A_MMAJOR = 1, B_KMAJOR = 0, B_NMAJOR = 1 };

// leading dimension of the gathered subset panels: padded so that rows start 16-byte aligned
static inline int64_t panel_ld(int64_t s) { return s > 0 ? round_up(s, 4) : 4; }

static inline int grid_for(modl_ctx *ctx, int64_t work, int per_sm = 8)
{
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    int64_t g = work < cap ? work : cap;
    return (int)(g < 1 ? 1 : g);
}

// gemm_simt.cu
template <typename T>
int gemm_simt(modl_ctx *ctx, int la, int lb, int64_t M, int64_t N, int64_t K, T alpha, const T *A,
              int64_t lda, const T *B, int64_t ldb, T beta, T *C, int64_t ldc, cudaStream_t st,
              WsSlot part_slot = WS_GEMM_PART);

// cd_launch.cu
template <typename T>
int cd_launch(modl_ctx *ctx, const T *G, int64_t g_stride, const T *Dx, const T *xnorm2, T *code,
              const int64_t *indices, T *code_batch, int64_t b, int64_t k, T alpha, T beta, T tol,
              int max_iter, int positive, int32_t *sweeps, cudaStream_t st);

// ridge_launch.cu
template <typename T>
int ridge_solve(modl_ctx *ctx, const T *G, int64_t g_stride, T *Dx, T *code, const int64_t *indices,
                T *code_batch, int64_t b, int64_t k, T alpha, cudaStream_t st);

// bcd_launch.cu
template <typename T>
int bcd_update(modl_ctx *ctx, T *Dp, const T *Bp, int64_t lds, const T *C, T *comp_norm,
               const int32_t *d_order, int64_t k, int64_t s, T l1_ratio, int positive, cudaStream_t st,
               unsigned *start_flag = nullptr, unsigned start_serial = 0);

// api.cu -- the phases of one minibatch step (modl_batch_fit_* without the context guard)
template <typename T>
int batch_fit_impl(modl_ctx *ctx, const modl_step_params *q, void *stream);

// tc_gemm.cu -- tcgen05 3xTF32 GEMM on packed split panels (float only)
size_t tc_packed_elems(int64_t rows, int64_t kd, int64_t rows_per_block = 128);
int tc_pick_bn(const modl_ctx *ctx, int64_t N, int64_t mtiles, int sm_avail = 0);
int64_t tc_rows_padded(int64_t rows);
int tc_pack_rows(modl_ctx *ctx, const float *src, int64_t ld, int64_t rows, int64_t p, const int64_t *subset,
                 int64_t kd, float *packed, int64_t row0, int64_t rows_pad_end, float *plain, int64_t ldp,
                 float *norm2, cudaStream_t st);
int tc_pack_cols(modl_ctx *ctx, const float *src, int64_t ld, int64_t kd, int64_t rows, float *packed,
                 int rows_per_block, cudaStream_t st);
// bn = accumulator tile width = rows per block of the packed B operand (multiple of 16, <= 256)
int tc_gemm(modl_ctx *ctx, const float *Apacked, const float *Bpacked, int64_t M, int64_t N, int64_t Kd, float alpha,
            float beta, float *C, int64_t ldc, int bn, cudaStream_t st, WsSlot part_slot = WS_GEMM_PART,
            float *C2 = nullptr, int64_t ldc2 = 0, int64_t n_split = 0, int64_t N1 = 0, const float *Braw = nullptr,
            int64_t ldb_raw = 0);

}  // namespace modl
