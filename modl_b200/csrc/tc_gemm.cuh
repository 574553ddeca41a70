// tc_gemm.cuh -- FP32-accurate dense contractions of the step on the 5th-generation tensor
// cores: tcgen05.mma (kind::tf32) with the error-compensated 3xTF32 split, operands staged by
// TMA bulk copies, accumulators in tensor memory.
//
// Which products  [ref: modl/decomposition/dict_fact.py]
//   [G ; Dx] = r [D_sub ; X_sub] . D_sub^T        :595, :604   contraction over the feature subset
//   C_ = (1-w) C_ + w/b code^T code               :573         contraction over the batch
//   B_ = (1-w) B_ + w/b code^T X                  :564         contraction over the batch
//
// Why a split.  Single-pass TF32 (10-bit mantissa) perturbs G and Dx at 1e-3, which flips the
// coordinate-descent stop decisions and breaks the 1e-4 parity bound.  Every fp32
// operand x is therefore stored as  hi = tf32(x)  (round to nearest) and  lo = tf32(x - hi)
// (x - hi is exact in fp32), and each product is accumulated as
//        lo_a . hi_b  +  hi_a . lo_b  +  hi_a . hi_b          (fp32 accumulation in TMEM)
// whose dropped terms are O(2^-24) relative: fp32-level accuracy at a third of the TF32 rate,
// which is still ~5x the FP32 FMA rate of the CUDA cores.
//
// Operand format ("packed split panels").  A pack kernel (one pass over the source, fused with
// the column gather of the feature subset or with the transposition the contraction needs)
// writes each operand in exactly the shared-memory image tcgen05 wants, so the GEMM CTA needs no
// register staging at all: one elected thread moves whole operand blocks with cp.async.bulk
// (TMA), one elected thread issues the MMAs, four warps drain TMEM.
//   operand = R rows (M or N side) x Kd contraction, zero padded to 128-row x 32-k blocks;
//   block (rb, kb) = 32 KB contiguous:  [hi | lo],  each  [8 chunks][128 rows][4 floats]
//   i.e. the canonical K-major, no-swizzle UMMA layout: 8-row x 16-byte core matrices,
//   SBO (next 8 rows) = 128 B, LBO (next 16-byte K chunk) = 2048 B.
#pragma once
#include "common.cuh"

namespace modl {

constexpr int TC_ROWS = 128;                 // rows of an operand block (= UMMA M, = UMMA N)
constexpr int TC_BK = 32;                    // contraction elements per block
constexpr int TC_CHUNKS = TC_BK / 4;         // 16-byte K chunks per block
constexpr int TC_HALF_FLOATS = TC_ROWS * TC_BK;            // floats of the hi (or lo) half
constexpr int TC_BLOCK_FLOATS = 2 * TC_HALF_FLOATS;        // hi + lo
constexpr unsigned TC_BLOCK_BYTES = TC_BLOCK_FLOATS * 4;   // 32 KB
constexpr unsigned TC_HALF_BYTES = TC_HALF_FLOATS * 4;     // 16 KB
constexpr int TC_MAX_STAGES = 3;
constexpr int TC_THREADS = 192;              // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int TC_MAX_BN = 256;               // widest accumulator tile (UMMA N)
// The tensor core adds every MMA into the TMEM accumulator with truncation, so the error of one
// accumulator grows linearly with the number of MMAs chained into it (measured: 1e-6 relative after
// 48, 6e-6 after 228 on same-sign data).  Long contractions therefore rotate over up to 4
// accumulators, one k-block (12 MMAs) at a time; the epilogue adds them in fp32 (RN).
constexpr int TC_CHAIN = 1;
constexpr size_t TC_SMEM_BUDGET = 226 * 1024;

// An operand is cut into blocks of `rpb` rows (rows-per-block: 128 on the M side; the UMMA N of
// the launch, a multiple of 16, on the N side) x 32 contraction elements.
__host__ __device__ inline int64_t tc_row_blocks(int64_t rows, int64_t rpb = TC_ROWS) { return ceil_div(rows, rpb); }
__host__ __device__ inline int64_t tc_k_blocks(int64_t kd) { return ceil_div(kd, TC_BK); }
__host__ __device__ inline size_t tc_block_floats(int64_t rpb) { return (size_t)(2 * rpb * TC_BK); }
// floats of a packed operand with `rows` rows and contraction length `kd`
__host__ __device__ inline size_t tc_packed_floats(int64_t rows, int64_t kd, int64_t rpb = TC_ROWS)
{
    return (size_t)tc_row_blocks(rows, rpb) * (size_t)tc_k_blocks(kd) * tc_block_floats(rpb);
}

// ---------------------------------------------------------------------------------------
// split + pack
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float tc_round_tf32(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));           // round to nearest, low 13 bits zero
    return __uint_as_float(r);
}
__device__ __forceinline__ void tc_split(float x, float &hi, float &lo)
{
    hi = tc_round_tf32(x);                 // |x - hi| <= 2^-12 |x|, and x - hi is exact in fp32
    lo = tc_round_tf32(x - hi);            // residual after both terms <= 2^-24 |x|
}

// position (in floats) of the 16-byte unit (row r, chunk c) inside the packed operand whose
// contraction spans `nkb` blocks; the lo copy sits TC_HALF_FLOATS further.
__device__ __forceinline__ size_t tc_unit_offset(int64_t r, int64_t c, int64_t nkb, int64_t rpb)
{
    const int64_t rb = r / rpb, rr = r % rpb, kb = c / TC_CHUNKS, cc = c % TC_CHUNKS;
    return ((size_t)(rb * nkb + kb)) * tc_block_floats(rpb) + (size_t)(cc * rpb + rr) * 4;
}

__device__ __forceinline__ void tc_store_unit(float *__restrict__ packed, size_t off, const float (&v)[4], int64_t rpb)
{
    float4 h, l;
    tc_split(v[0], h.x, l.x); tc_split(v[1], h.y, l.y); tc_split(v[2], h.z, l.z); tc_split(v[3], h.w, l.w);
    *reinterpret_cast<float4 *>(packed + off) = h;
    *reinterpret_cast<float4 *>(packed + off + rpb * TC_BK) = l;
}

// Rows of a row-major source become operand rows; the contraction runs along the source row,
// optionally through a column gather:  op[row0 + r, j] = src[r, subset ? subset[j] : j].
// grid.x: row r in [0, rows_pad) where rows_pad also covers the zero padding up to the block
// boundary (only when pad_rows), grid-stride; one thread per 16-byte unit.
// Optionally also emits the plain gathered panel (plain[r, j], ld = ldp) and the squared norm of
// the FULL source row, which the CD stop test and the dictionary update need anyway.
__global__ void __launch_bounds__(256)
tc_pack_rows_kernel(const float *__restrict__ src, int64_t ld, int rows, int p, const int64_t *__restrict__ subset,
                    int kd, float *__restrict__ packed, int64_t row0, int64_t rows_pad_end,
                    float *__restrict__ plain, int64_t ldp, float *__restrict__ norm2)
{
    __shared__ float scratch[33];
    const int64_t nkb = tc_k_blocks(kd);
    const int nunits = (int)(nkb * TC_CHUNKS);
    for (int64_t r = blockIdx.x; row0 + r < rows_pad_end; r += gridDim.x) {
        const bool real = r < rows;
        const float *row = src + r * ld;
        if (real && norm2 != nullptr) {
            float acc = 0.f, acc1 = 0.f;
            int j0 = 0;
            if (((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {     // 128-bit streaming loads
                const float4 *row4 = reinterpret_cast<const float4 *>(row);
                const int n4 = p >> 2;
#pragma unroll 4
                for (int j = threadIdx.x; j < n4; j += blockDim.x) {
                    const float4 v = __ldg(row4 + j);
                    acc = fmaf(v.x, v.x, acc); acc1 = fmaf(v.y, v.y, acc1);
                    acc = fmaf(v.z, v.z, acc); acc1 = fmaf(v.w, v.w, acc1);
                }
                j0 = n4 << 2;
            }
            for (int j = j0 + threadIdx.x; j < p; j += blockDim.x) {
                const float v = row[j];
                acc = fmaf(v, v, acc);
            }
            acc = block_sum(acc + acc1, scratch);
            if (threadIdx.x == 0) norm2[r] = acc;
        }
        for (int c = threadIdx.x; c < nunits; c += blockDim.x) {
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * c + i;
                v[i] = (real && j < kd) ? row[subset ? subset[j] : (int64_t)j] : 0.f;
            }
            tc_store_unit(packed, tc_unit_offset(row0 + r, c, nkb, TC_ROWS), v, TC_ROWS);
            if (real && plain != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * c + i < kd) plain[r * ldp + 4 * c + i] = v[i];
            }
        }
    }
}

// Columns of a row-major source become operand rows (the contraction runs DOWN the source):
//     op[r, j] = src[j, r],   src is kd x rows (ld).
// One thread per 16-byte unit, consecutive threads take consecutive r: coalesced 128 B reads of
// four source rows and coalesced 16-byte-per-thread writes.
__global__ void __launch_bounds__(256)
tc_pack_cols_kernel(const float *__restrict__ src, int64_t ld, int kd, int rows, float *__restrict__ packed, int rpb)
{
    const int64_t nkb = tc_k_blocks(kd);
    const int64_t rows_pad = tc_row_blocks(rows, rpb) * rpb;
    const int64_t nunits = nkb * TC_CHUNKS;
    const int64_t total = rows_pad * nunits;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = e / rows_pad, r = e % rows_pad;
        float v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t j = 4 * c + i;
            v[i] = (r < rows && j < kd) ? src[j * ld + r] : 0.f;
        }
        tc_store_unit(packed, tc_unit_offset(r, c, nkb, rpb), v, rpb);
    }
}

// ---------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol error becomes a trap (launch failure), never a hung GPU.
__device__ __forceinline__ void tc_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_tmem_alloc(unsigned smem_dst, unsigned cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tc_tmem_dealloc(unsigned taddr, unsigned cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, 128 x 128 x 8, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, uint64_t desc_a, uint64_t desc_b, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(unsigned bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld32(unsigned taddr, float (&v)[32])
{
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor of a [8 chunks][128 rows][16 B] half block
// starting at shared address `saddr` (+ the K offset of the MMA inside the block).
__device__ __forceinline__ uint64_t tc_smem_desc(unsigned saddr, unsigned lbo, unsigned sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);                      // start address,            bits [0,14)
    d |= (uint64_t)(lbo >> 4) << 16;                               // leading byte offset (K),  bits [16,30)
    d |= (uint64_t)(sbo >> 4) << 32;                               // stride byte offset (M/N), bits [32,46)
    d |= (uint64_t)1 << 46;                                        // descriptor version (sm_100)
    return d;                                                      // layout type 0 = no swizzle
}

// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = bn
__host__ __device__ inline unsigned tc_idesc(unsigned bn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((bn >> 3) << 17) | ((unsigned)(TC_ROWS >> 4) << 24);
}

// ---------------------------------------------------------------------------------------
// C[M x N] = alpha * A . B^T + beta * C      (A: M x Kd, B: N x Kd, packed split panels)
// grid = (n blocks, m blocks, split-K slices).  With part != NULL every slice writes its raw
// partial tile to part[z][M][N] and a fixed-order reduction applies alpha/beta afterwards.
// ---------------------------------------------------------------------------------------
struct TcGemmParams {
    const float *A;      // packed, 128-row blocks: tc_row_blocks(M) x nkb
    const float *B;      // packed, bn-row blocks:  tc_row_blocks(N, bn) x nkb
    int bn;              // accumulator tile width (UMMA N): multiple of 16, 16..256
    int stages;          // smem pipeline depth (<= TC_MAX_STAGES)
    int nacc;            // accumulators of bn columns each, used round-robin in chains of TC_CHAIN k-blocks
    unsigned tmem_cols;  // power of two >= nacc * bn
    float *C;
    int64_t ldc;
    float *part;         // split-K partials or NULL
    int M, N;
    int nkb;             // k blocks of the whole contraction
    int kb_per_split;    // k blocks per grid.z slice
    float alpha, beta;
    unsigned sbo;        // descriptor stride byte offset: next 8-row group (128)
    // optional second destination: column tiles at or beyond n_split (a multiple of bn) are written to C2
    // (column n - n_split, N2 valid columns, leading dimension ldc2); tiles below it keep N1 valid columns
    float *C2;
    int64_t ldc2;
    int n_split, N1, N2;
    // optional RAW B operand (B == NULL): Braw is the row-major source whose COLUMNS are the operand rows,
    // op[n, kk] = Braw[kk * ldb_raw + n] (kk < Kd, n < N) -- e.g. the batch X of  B_ += code^T X.  The four epilogue
    // warps then build every B block in shared memory themselves (coalesced loads along n, hi/lo TF32 split in
    // registers, the same 16-byte units the pack kernel writes), so the 8-byte-per-element packed copy of X never
    // goes through HBM.
    const float *Braw;
    int64_t ldb_raw;
    int Kd;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(TcGemmParams P)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) unsigned long long bars[3 * TC_MAX_STAGES + 1];
    __shared__ unsigned tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(tc_smem) + 1023u) & ~1023u;
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * (unsigned)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (unsigned)(TC_MAX_STAGES + s); };
    const unsigned done_bar = bar0 + 8u * (unsigned)(2 * TC_MAX_STAGES);
    auto fullb_bar = [&](int s) { return bar0 + 8u * (unsigned)(2 * TC_MAX_STAGES + 1 + s); };   // raw-B blocks converted
    const bool rawB = P.Braw != nullptr;
    const int STAGES = P.stages;
    const unsigned a_bytes = TC_BLOCK_BYTES, b_half = (unsigned)P.bn * TC_BK * 4u, b_bytes = 2u * b_half;
    const unsigned stage_bytes = a_bytes + b_bytes;
    const size_t b_block_floats = (size_t)2 * P.bn * TC_BK;

    const int nb = blockIdx.x, mb = blockIdx.y;
    const int kb0 = blockIdx.z * P.kb_per_split;
    const int kb1 = min(P.nkb, kb0 + P.kb_per_split);
    const int nk = kb1 - kb0;                 // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc_mbar_init(full_bar(s), 1); tc_mbar_init(empty_bar(s), 1); tc_mbar_init(fullb_bar(s), TC_THREADS - 64);
        }
        tc_mbar_init(done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) tc_tmem_alloc((unsigned)__cvta_generic_to_shared(&tmem_base_s), P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const float *Ab = P.A + ((size_t)mb * P.nkb + kb0) * TC_BLOCK_FLOATS;
            const float *Bb = P.B + ((size_t)nb * P.nkb + kb0) * b_block_floats;
            for (int i = 0; i < nk; ++i) {
                const int s = i % STAGES;
                const unsigned ph = (unsigned)(i / STAGES) & 1u;
                tc_mbar_wait(empty_bar(s), ph ^ 1u);                 // slot free (first pass: immediately)
                tc_mbar_expect_tx(full_bar(s), rawB ? a_bytes : stage_bytes);
                const unsigned sa = smem0 + (unsigned)s * stage_bytes;
                tc_bulk_g2s(sa, Ab + (size_t)i * TC_BLOCK_FLOATS, a_bytes, full_bar(s));
                if (!rawB) tc_bulk_g2s(sa + a_bytes, Bb + (size_t)i * b_block_floats, b_bytes, full_bar(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc = tc_idesc((unsigned)P.bn);
            const unsigned lbo_a = TC_ROWS * 16u, lbo_b = (unsigned)P.bn * 16u;   // next 16-byte K chunk of the block
            for (int i = 0; i < nk; ++i) {
                const int s = i % STAGES;
                const unsigned ph = (unsigned)(i / STAGES) & 1u;
                tc_mbar_wait(full_bar(s), ph);
                if (rawB) tc_mbar_wait(fullb_bar(s), ph);
                tc_fence_after();
                const unsigned sa = smem0 + (unsigned)s * stage_bytes;            // A: hi | lo
                const unsigned sb = sa + a_bytes;                                 // B: hi | lo
#pragma unroll
                const int chain = i / TC_CHAIN;
                const unsigned acc = tmem + (unsigned)((chain % P.nacc) * P.bn);  // column offset of this chain's accumulator
                const bool fresh = chain < P.nacc && (i % TC_CHAIN) == 0;         // first k-block ever written to it
#pragma unroll
                for (int k8 = 0; k8 < TC_BK / 8; ++k8) {
                    const unsigned ka = (unsigned)k8 * 2u * lbo_a, kb = (unsigned)k8 * 2u * lbo_b;   // two chunks per MMA
                    const uint64_t a_hi = tc_smem_desc(sa + ka, lbo_a, P.sbo), a_lo = tc_smem_desc(sa + TC_HALF_BYTES + ka, lbo_a, P.sbo);
                    const uint64_t b_hi = tc_smem_desc(sb + kb, lbo_b, P.sbo), b_lo = tc_smem_desc(sb + b_half + kb, lbo_b, P.sbo);
                    tc_mma_tf32(acc, a_lo, b_hi, idesc, !(fresh && k8 == 0));      // small terms first
                    tc_mma_tf32(acc, a_hi, b_lo, idesc, 1u);
                    tc_mma_tf32(acc, a_hi, b_hi, idesc, 1u);
                }
                tc_commit(empty_bar(s));                                          // smem slot reusable when these finish
            }
            tc_commit(done_bar);                                                  // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> shared (transpose) -> coalesced global =====
        // A thread of tcgen05.ld owns one accumulator ROW; storing rows straight to global would
        // touch 32 different rows per warp instruction.  The operand stages are free once the
        // last MMA has completed, so the tile is staged there and written out row by row, one
        // 512-byte segment per warp instruction (and C is read the same way when beta != 0).
        if (rawB) {
            // ===== converters: the B block of every stage, straight from the raw source =====
            // unit (n, c) = the four contraction elements 4c..4c+3 of operand row n; consecutive threads take consecutive n
            // (128-byte coalesced loads of four source rows, conflict-free 16-byte shared stores)
            const int ct = (int)threadIdx.x - 64, bn = P.bn;
            const int n_base = nb * bn;
            const int nunits = bn * TC_CHUNKS;
            constexpr int NCT = TC_THREADS - 64;
            constexpr int TC_RAW_MAX_BN = 160;                                  // widest tile of the raw-B form (launcher checks)
            constexpr int UMAX = (TC_RAW_MAX_BN * TC_CHUNKS + NCT - 1) / NCT;   // units per thread
            float v[2][UMAX][4];
            // loads of block i into v[i & 1]; every load of a block is in flight at once, and the loads of block i + 1 are
            // issued BEFORE block i is converted and stored: the global latency hides behind the conversion of the previous block
            auto issue = [&](int i, float (&vv)[UMAX][4]) {
                const int kk0 = (kb0 + i) * TC_BK;
#pragma unroll
                for (int u = 0; u < UMAX; ++u) {
                    const int e = ct + u * NCT;
                    const int c = e / bn, n = e - c * bn;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int kk = kk0 + 4 * c + w;
                        vv[u][w] = (e < nunits && n_base + n < P.N && kk < P.Kd) ? __ldg(P.Braw + (int64_t)kk * P.ldb_raw + n_base + n) : 0.f;
                    }
                }
            };
            auto convert = [&](int i, const float (&vv)[UMAX][4]) {
                const int s = i % STAGES;
                const unsigned ph = (unsigned)(i / STAGES) & 1u;
                tc_mbar_wait(empty_bar(s), ph ^ 1u);                 // the MMAs that read this slot are done
                const unsigned sb = smem0 + (unsigned)s * stage_bytes + a_bytes;
#pragma unroll
                for (int u = 0; u < UMAX; ++u) {
                    const int e = ct + u * NCT;
                    if (e < nunits) {
                        float4 h, l;
                        tc_split(vv[u][0], h.x, l.x); tc_split(vv[u][1], h.y, l.y); tc_split(vv[u][2], h.z, l.z); tc_split(vv[u][3], h.w, l.w);
                        const unsigned dst = sb + 16u * (unsigned)e;             // (c * bn + n) * 16 bytes
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + b_half), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic stores -> visible to the tensor core
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(fullb_bar(s)) : "memory");
            };
            issue(0, v[0]);
            for (int i = 0; i < nk; i += 2) {
                if (i + 1 < nk) issue(i + 1, v[1]);
                convert(i, v[0]);
                if (i + 1 < nk) {
                    if (i + 2 < nk) issue(i + 2, v[0]);
                    convert(i + 1, v[1]);
                }
            }
        }
        tc_mbar_wait(done_bar, 0);
        tc_fence_after();
        const int q = warp & 3;                        // TMEM lane quadrant this warp may read
        const unsigned trow = tmem + ((unsigned)(q * 32) << 16);
        const int pitch = P.bn + 4;                    // floats; 16-byte aligned rows, conflict-free quarter-warps
        const unsigned tile0 = smem0;
        const int nused = min(P.nacc, (nk + TC_CHAIN - 1) / TC_CHAIN);   // accumulators that received MMAs
        for (int c0 = 0; c0 < P.bn; c0 += 32) {
            float v[32];
            __syncwarp();                              // tcgen05.ld is warp-collective (.sync.aligned)
            tc_tmem_ld32(trow + (unsigned)c0, v);
            for (int a = 1; a < nused; ++a) {          // uniform trip count
                float u[32];
                __syncwarp();
                tc_tmem_ld32(trow + (unsigned)(a * P.bn + c0), u);
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) v[jj] += u[jj];
            }
            const unsigned dst = tile0 + 4u * (unsigned)((q * 32 + lane) * pitch + c0);
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
                if (c0 + jj < P.bn)                    // bn is a multiple of 16: the last strip may be half valid
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(dst + 4u * jj), "f"(v[jj]), "f"(v[jj + 1]),
                                 "f"(v[jj + 2]), "f"(v[jj + 3]) : "memory");
        }
        asm volatile("bar.sync 2, 128;\n" ::: "memory");        // the four epilogue warps
        const int ew = warp - 2;
        const bool raw = P.part != nullptr;            // split-K: raw partial tiles, at their column in the N-wide product
        const bool second = !raw && P.C2 != nullptr && nb * P.bn >= P.n_split;
        const int n_tile = second ? nb * P.bn - P.n_split : nb * P.bn;     // first column of the tile in its destination
        const int n_lim = raw ? P.N : (second ? P.N2 : P.N1);
        float *const c_base = second ? P.C2 : P.C;
        const int64_t c_ld = second ? P.ldc2 : P.ldc;
        const bool use_c = !raw && P.beta != 0.f;
        // 8 rows per batch (r0, r0 + 4, ..., r0 + 28): their C reads are all in flight before the first is consumed
        for (int r0 = ew; r0 < TC_ROWS; r0 += 32) {
            if (mb * TC_ROWS + r0 >= P.M) break;       // uniform per warp
            for (int c = lane * 4; c < P.bn; c += 128) {
                const int n = n_tile + c;
                float4 t[8], cc[8];
                float *dst[8];
                bool fast[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = r0 + 4 * u, m = mb * TC_ROWS + r;
                    float *grow = raw ? P.part + ((size_t)blockIdx.z * P.M + m) * P.N : c_base + (int64_t)m * c_ld;
                    dst[u] = grow + n;
                    const bool live = m < P.M && n < n_lim;
                    fast[u] = live && n + 4 <= n_lim && ((reinterpret_cast<uintptr_t>(dst[u]) & 15) == 0);
                    if (!live) dst[u] = nullptr;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(t[u].x), "=f"(t[u].y), "=f"(t[u].z), "=f"(t[u].w)
                                 : "r"(tile0 + 4u * (unsigned)(r * pitch + c)));
                    cc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (fast[u] && use_c) cc[u] = *reinterpret_cast<const float4 *>(dst[u]);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (dst[u] == nullptr) continue;
                    if (fast[u]) {
                        float4 o = t[u];
                        if (!raw) {
                            o.x = fmaf(P.beta, cc[u].x, P.alpha * o.x); o.y = fmaf(P.beta, cc[u].y, P.alpha * o.y);
                            o.z = fmaf(P.beta, cc[u].z, P.alpha * o.z); o.w = fmaf(P.beta, cc[u].w, P.alpha * o.w);
                        }
                        *reinterpret_cast<float4 *>(dst[u]) = o;
                    } else {
                        const float e[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            if (n + w >= n_lim) continue;
                            float o = e[w];
                            if (!raw) {
                                o *= P.alpha;
                                if (P.beta != 0.f) o = fmaf(P.beta, dst[u][w], o);
                            }
                            dst[u][w] = o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_tmem_dealloc(tmem, P.tmem_cols);
}

}  // namespace modl
