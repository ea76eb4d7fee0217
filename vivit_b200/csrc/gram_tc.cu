// fp32 Gram / NT-GEMM entry points on the tcgen05 + TMA kernel of gemm_tc.cuh (StdStore epilogue, split-K).
#include <cstdlib>

#include "gemm_tc.cuh"

namespace vvt {

bool gram_tc_eligible(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                      int64_t batch) {
  static const bool disabled = getenv("VVT_NO_TCGEN05") != nullptr;
  if (disabled || batch != 1) return false;
  // tiny products: the SIMT-fed path has less overhead (two k-blocks are enough when the output is large)
  if (K < 2 * tc::BK || M < 32 || N < 32 || (K < 4 * tc::BK && M * N < 512 * 512)) return false;
  if (!tc::operands_ok(A, B, lda, ldb, 0, 0)) return false;  // TMA alignment rules
  if (M >= (int64_t(1) << 31) || N >= (int64_t(1) << 31) || K >= (int64_t(1) << 31)) return false;
  return true;
}

// Split-K factor: one CTA per SM is resident, so the kernel runs in waves of num_sms CTAs; pick the
// split count (<= max_splits, >= 8 k-blocks per CTA) that minimises waves x k-blocks per CTA.
static int64_t plan_splits(int64_t tiles, int64_t kblocks, int64_t max_splits) {
  const int64_t sms = num_sms();
  int64_t best = 1, best_cost = INT64_MAX;
  const int64_t hi = vmax<int64_t>(1, vmin<int64_t>(vmin<int64_t>(64, max_splits), kblocks / 8));
  for (int64_t sp = 1; sp <= hi; ++sp) {
    const int64_t per = ceil_div(kblocks, sp);
    const int64_t ctas = tiles * ceil_div(kblocks, per);
    const int64_t cost = ceil_div(ctas, sms) * (per + 6);  // + prologue/epilogue, in k-block units
    if (cost < best_cost) best_cost = cost, best = sp;
  }
  return best;
}

int64_t gram_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, bool symmetric) {
  const int tiles_m = int(ceil_div(M, tc::BM)), tiles_n = int(ceil_div(N, tc::BN));
  const int tiles = symmetric ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
  const int64_t splits = plan_splits(tiles, ceil_div(K, tc::BK), 64);
  return splits > 1 ? splits * M * N * 4 : 0;
}

int launch_gram_tc(const float* A, const float* B, StdStore<float> st, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, bool symmetric, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                   const char* what) {
  using namespace tc;
  // A operand from tensor memory unless VVT_GRAM_SS is set (both operands from shared memory: the round-1 kernel)
  static const bool ts = getenv("VVT_GRAM_SS") == nullptr;
  auto kern = ts ? gram_tc_kernel<StdStore<float>, true, true> : gram_tc_kernel<StdStore<float>, true, false>;
  static SmemOptIn opt_in[2];
  VVT_TRY(opt_in[ts].ensure(kern, SMEM_BYTES, what));
  CUtensorMap mapA, mapB;
  if (!make_map(&mapA, A, M, K, lda, 1, 0) || !make_map(&mapB, B, N, K, ldb, 1, 0))
    return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", what);
  const int tiles_m = int(ceil_div(M, BM)), tiles_n = int(ceil_div(N, BN));
  const int tiles = symmetric ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
  const int64_t kblocks = ceil_div(K, BK);
  const int64_t fit = workspace ? workspace_bytes / vmax<int64_t>(1, M * N * 4) : 0;
  int64_t splits = plan_splits(tiles, kblocks, vmax<int64_t>(1, fit));
  const int64_t per = ceil_div(kblocks, splits);
  splits = ceil_div(kblocks, per);
  st.N = N;
  st.slab = M * N;
  StdStore<float> kst = st;
  kst.partial = splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
  dim3 grid(unsigned(tiles), unsigned(splits), 1);
  kern<<<grid, ts ? THREADS_TS : THREADS, SMEM_BYTES, stream>>>(mapA, mapB, kst, M, N, tiles_n, symmetric ? 1 : 0, int(kblocks),
                                                               int(per), 1);
  VVT_TRY(launched(what));
  if (splits > 1) {
    const int64_t total = M * N;
    const int blocks = int(vmin<int64_t>(ceil_div(total, 256), 8 * num_sms()));
    splitk_reduce_kernel<float><<<blocks, 256, 0, stream>>>(kst, M, N, int(splits));
    VVT_TRY(launched(what));
  }
  return VVT_OK;
}

}  // namespace vvt
