// K2: region x text similarity on the 5th-generation tensor cores.
//
// Replaces torch.matmul + alpha*s+beta (model/heads/text_contrastive.py:144,147) and, through
// the fused epilogue, similarity.max(dim=1) (model/yolo_clip.py:198-202).
//
//   S[b, m, n] = alpha * inv_norm_r[b, m] * sum_k A[b, m, k] * T[tb, n, k] + beta
//
// A (regions) and T (text, rows already unit-norm) are bf16, K-major.  A persistent,
// warp-specialised kernel: one CTA per SM, each CTA owns an M tile (128 anchors) and walks all N
// tiles of the vocabulary for it, so that the row max/argmax completes inside one CTA.
//
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) -> 4-stage smem ring
//   warp 1      MMA issuer     tcgen05.mma.kind::f16 128 x N x 16, accumulators in TMEM
//                              (2 stages x 256 columns), tcgen05.commit -> mbarriers
//   warps 2..5  epilogue       tcgen05.ld -> scale / affine -> running max/argmax and/or
//                              smem-staged coalesced logit stores
//
// "split" mode runs three bf16 passes over hi/lo operand halves (A_hi*T_hi + A_hi*T_lo +
// A_lo*T_hi) into the same accumulator: fp32-class accuracy on the bf16 tensor pipe.
#include "common.cuh"
#include "ptx.cuh"

namespace ovdet {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;          // maximum N tile (TMA box rows / UMMA N)
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KiB
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;        // 512
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARPS = 4;
constexpr int STAGE_PITCH = 33;                        // floats, conflict-free transpose

struct SmemLayout {
  // offsets from the 1024-byte aligned base
  static constexpr int a_off = 0;
  static constexpr int b_off = STAGES * A_STAGE_BYTES;
  static constexpr int epi_off = b_off + STAGES * B_STAGE_BYTES;
  static constexpr int epi_bytes = EPI_WARPS * 32 * STAGE_PITCH * 4;
  static constexpr int bar_off = epi_off + epi_bytes;
  static constexpr int num_bars = 2 * STAGES + 2 * ACC_STAGES;
  static constexpr int tmem_ptr_off = bar_off + num_bars * 8;
  static constexpr int total = tmem_ptr_off + 16;
};
constexpr int SMEM_BYTES = SmemLayout::total + 1024;    // slack for manual 1024 B alignment

struct GemmParams {
  int batch;              // independent problems
  int rows;               // M per problem
  int classes;            // N
  int kb_per_pass;        // dim / 64
  int passes;             // 1 (bf16) or 3 (hi/lo split)
  int dim;
  int text_batched;
  int box_n;              // TMA box rows for the text operand = N tile step
  int m_tiles;            // per problem
  int n_tiles;
  float alpha, beta;
  const float* inv_norm_r;
  void* logits;
  int logits_bf16;
  long long ldc;
  float* row_max;
  int* row_arg;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
sim_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = ptx::smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - raw);

  const uint32_t smem_a = base + SmemLayout::a_off;
  const uint32_t smem_b = base + SmemLayout::b_off;
  float* epi_stage = reinterpret_cast<float*>(base_ptr + SmemLayout::epi_off);
  const uint32_t bars = base + SmemLayout::bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tmem_empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + ACC_STAGES + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + SmemLayout::tmem_ptr_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC_STAGES; ++s) {
      ptx::mbar_init(tmem_full_bar(s), 1);
      ptx::mbar_init(tmem_empty_bar(s), EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(base + SmemLayout::tmem_ptr_off, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_tiles = p.batch * p.m_tiles;
  const int num_kb = p.kb_per_pass * p.passes;
  const uint32_t stage_tx = A_STAGE_BYTES + (uint32_t)p.box_n * BLOCK_K * 2;

  if (warp == 0) {
    // ================================ TMA producer ===========================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / p.m_tiles;
        const int m0 = (tile % p.m_tiles) * BLOCK_M;
        const int tb = p.text_batched ? b : 0;
        for (int nt = 0; nt < p.n_tiles; ++nt) {
          const int n0 = nt * p.box_n;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            const int pass = kb / p.kb_per_pass;
            const int kk = (kb % p.kb_per_pass) * BLOCK_K;
            const int ka = kk + (pass == 2 ? p.dim : 0);
            const int kt = kk + (pass == 1 ? p.dim : 0);
            ptx::mbar_wait(empty_bar(s), ph ^ 1u);
            ptx::mbar_arrive_expect_tx(full_bar(s), stage_tx);
            ptx::tma_load_3d(smem_a + s * A_STAGE_BYTES, &tmap_a, full_bar(s), ka, m0, b);
            ptx::tma_load_3d(smem_b + s * B_STAGE_BYTES, &tmap_b, full_bar(s), kt, n0, tb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =============================================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
          const int n0 = nt * p.box_n;
          int n_size = p.classes - n0;
          n_size = n_size >= p.box_n ? p.box_n : ((n_size + 15) & ~15);
          const uint32_t idesc = ptx::umma_idesc_bf16_f32(BLOCK_M, (uint32_t)n_size);
          const int as = acc_it & 1;
          const uint32_t aph = (acc_it >> 1) & 1u;
          ptx::mbar_wait(tmem_empty_bar(as), aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1u;
            ptx::mbar_wait(full_bar(s), ph);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::umma_desc_k_sw128(smem_a + s * A_STAGE_BYTES);
            const uint64_t b_desc = ptx::umma_desc_k_sw128(smem_b + s * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // +32 B per UMMA_K step inside the 128 B swizzle row (descriptor units of 16 B)
              ptx::umma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
            }
            ptx::umma_commit(empty_bar(s));          // smem slot reusable once these MMAs retire
          }
          ptx::umma_commit(tmem_full_bar(as));       // accumulator complete
        }
      }
    }
  } else {
    // ================================ epilogue ===============================================
    const int lg = warp & 3;                         // TMEM lane group this warp may access
    float* stage = epi_stage + (warp - 2) * 32 * STAGE_PITCH;
    const bool want_max = p.row_max != nullptr;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = tile / p.m_tiles;
      const int m0 = (tile % p.m_tiles) * BLOCK_M;
      const int row = m0 + lg * 32 + lane;           // row inside the problem
      const bool row_ok = row < p.rows;
      const long long grow = (long long)b * p.rows + row;
      const float inv = (row_ok && p.inv_norm_r) ? p.inv_norm_r[grow] : (row_ok ? 1.f : 0.f);
      const float scale = p.alpha * inv;
      const float beta = p.beta;
      // four independent (max, argmax) chains (column % 4) keep the compare/select dependency
      // chain short; they are merged, lowest index winning ties, once per M tile
      float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int bi[4] = {0, 0, 0, 0};
      for (int nt = 0; nt < p.n_tiles; ++nt, ++acc_it) {
        const int n0 = nt * p.box_n;
        const int n_valid = min(p.box_n, p.classes - n0);
        const int nchunks = (n_valid + 31) >> 5;
        const int as = acc_it & 1;
        const uint32_t aph = (acc_it >> 1) & 1u;
        ptx::mbar_wait(tmem_full_bar(as), aph);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BLOCK_N);

        auto consume = [&](uint32_t (&r)[32], int c) {
          const int c0 = c << 5;
          const int valid = n_valid - c0;              // > 0
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(scale, __uint_as_float(r[j]), beta));
          if (want_max) {
            const int col = n0 + c0;
            if (valid >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                if (v > bv[j & 3]) { bv[j & 3] = v; bi[j & 3] = col + j; }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                if (j < valid && v > bv[j & 3]) { bv[j & 3] = v; bi[j & 3] = col + j; }
              }
            }
          }
          if (p.logits != nullptr) {
            // registers (thread = row) -> smem transpose -> one coalesced row segment per
            // warp store instruction
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * STAGE_PITCH + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int col = n0 + c0 + lane;
            const bool col_ok = lane < valid;
            const int rows_here = min(32, p.rows - (m0 + lg * 32));
            const long long out_row0 = (long long)b * p.rows + m0 + lg * 32;
            if (p.logits_bf16) {
              __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.logits);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = __float2bfloat16_rn(stage[i * STAGE_PITCH + lane]);
            } else {
              float* out = static_cast<float*>(p.logits);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = stage[i * STAGE_PITCH + lane];
            }
            __syncwarp();
          }
        };

        // software pipeline over 32-column chunks: the TMEM load of chunk c+1 is in flight while
        // chunk c is consumed
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32(t_row, ra);
        for (int c = 0; c < nchunks; c += 2) {
          ptx::tmem_ld_wait();
          if (c + 1 < nchunks) ptx::tmem_ld_32x32(t_row + (uint32_t)((c + 1) << 5), rb);
          consume(ra, c);
          if (c + 1 < nchunks) {
            ptx::tmem_ld_wait();
            if (c + 2 < nchunks) ptx::tmem_ld_32x32(t_row + (uint32_t)((c + 2) << 5), ra);
            consume(rb, c + 1);
          }
        }
        // all TMEM reads of this accumulator stage are complete (wait::ld above)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(as));
      }
      if (want_max && row_ok) {
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (bv[q] > best || (bv[q] == best && bi[q] < best_idx)) { best = bv[q]; best_idx = bi[q]; }
        p.row_max[grow] = best;
        if (p.row_arg != nullptr) p.row_arg[grow] = best_idx;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side: tensor maps -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 [batch, rows, kop] row-major -> 3-D map (kop, rows, batch), box (64, box_rows, 1), SW128.
int make_operand_map(CUtensorMap* map, const void* ptr, int64_t batch, int64_t rows, int64_t kop,
                     int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return OVDET_ERR_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)kop, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)kop * 2, (cuuint64_t)rows * (cuuint64_t)kop * 2};
  cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OVDET_OK : OVDET_ERR_DRIVER;
}

}  // namespace
}  // namespace ovdet

extern "C" int ovdet_similarity(const void* regions_op, const void* text_op, const float* inv_norm_r,
                                int64_t batch, int64_t rows, int64_t classes, int64_t dim,
                                int split, int text_batched, float alpha, float beta,
                                void* logits, int logits_dtype, int64_t ldc,
                                float* row_max, int32_t* row_arg, void* stream) {
  using namespace ovdet;
  if (!regions_op || !text_op || batch < 0 || rows < 0 || classes <= 0 || dim <= 0)
    return OVDET_ERR_INVALID_ARG;
  if (!logits && !row_max) return OVDET_ERR_INVALID_ARG;
  if (row_arg && !row_max) return OVDET_ERR_INVALID_ARG;
  if (logits && (ldc < classes || (logits_dtype != OVDET_F32 && logits_dtype != OVDET_BF16)))
    return OVDET_ERR_INVALID_ARG;
  if (dim % BLOCK_K != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)regions_op & 15) || ((uintptr_t)text_op & 15)) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (batch == 0 || rows == 0) return OVDET_OK;

  // A shared vocabulary turns the batch of GEMMs into one tall GEMM over all anchors.
  int64_t g_batch = batch, g_rows = rows;
  if (!text_batched) { g_rows = batch * rows; g_batch = 1; }
  if (g_rows >= (1ll << 31) || classes >= (1 << 30)) return OVDET_ERR_UNSUPPORTED_SHAPE;

  const int64_t kop = dim * (split ? 2 : 1);
  GemmParams p{};
  p.batch = (int)g_batch;
  p.rows = (int)g_rows;
  p.classes = (int)classes;
  p.kb_per_pass = (int)(dim / BLOCK_K);
  p.passes = split ? 3 : 1;
  p.dim = (int)dim;
  p.text_batched = text_batched ? 1 : 0;
  p.box_n = (int)(classes >= BLOCK_N ? BLOCK_N : ((classes + 15) & ~15ll));
  p.m_tiles = (int)ceil_div<int64_t>(g_rows, BLOCK_M);
  p.n_tiles = (int)ceil_div<int64_t>(classes, p.box_n);
  p.alpha = alpha;
  p.beta = beta;
  p.inv_norm_r = inv_norm_r;
  p.logits = logits;
  p.logits_bf16 = logits_dtype == OVDET_BF16;
  p.ldc = ldc;
  p.row_max = row_max;
  p.row_arg = row_arg;
  if ((int64_t)p.batch * p.m_tiles >= (1ll << 31)) return OVDET_ERR_UNSUPPORTED_SHAPE;

  CUtensorMap map_a, map_b;
  if (int rc = make_operand_map(&map_a, regions_op, g_batch, g_rows, kop, BLOCK_M)) return rc;
  if (int rc = make_operand_map(&map_b, text_op, text_batched ? batch : 1, classes, kop, p.box_n)) return rc;

  if (int rc = once_per_device(0, []() -> int {
        OVDET_CUDA_TRY(cudaFuncSetAttribute(sim_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        return OVDET_OK;
      })) return rc;
  const int64_t total_tiles = (int64_t)p.batch * p.m_tiles;
  const int grid = (int)(total_tiles < sm_count() ? total_tiles : sm_count());
  sim_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, as_stream(stream)>>>(map_a, map_b, p);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
