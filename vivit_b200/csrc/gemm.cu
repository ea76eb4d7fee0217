// GEMM-shaped entry points: Gram assembly (dense, structured Linear, cross terms),
// factor back-propagation through Linear, structured back-transform, Gram-space epilogue.
#include "gemm_core.cuh"

namespace vvt {

// tcgen05 + TMA path for fp32 products with K-contiguous operands (gram_tc.cu)
bool gram_tc_eligible(const float* A, const float* B, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb,
                      int64_t batch);
int64_t gram_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, bool symmetric);
int launch_gram_tc(const float* A, const float* B, StdStore<float> st, int64_t M, int64_t N, int64_t K, int64_t lda,
                   int64_t ldb, bool symmetric, void* workspace, int64_t workspace_bytes, cudaStream_t stream,
                   const char* what);

template <typename T>
static bool tc_path(int& status, const T*, const T*, StdStore<T>, int64_t, int64_t, int64_t, int64_t, int64_t, bool,
                    int64_t, void*, int64_t, cudaStream_t, const char*) {
  (void)status;
  return false;
}
template <>
bool tc_path<float>(int& status, const float* A, const float* B, StdStore<float> st, int64_t M, int64_t N, int64_t K,
                    int64_t lda, int64_t ldb, bool symmetric, int64_t batch, void* ws, int64_t wsb, cudaStream_t s,
                    const char* what) {
  if (!gram_tc_eligible(A, B, M, N, K, lda, ldb, batch)) return false;
  status = launch_gram_tc(A, B, st, M, N, K, lda, ldb, symmetric, ws, wsb, s, what);
  return true;
}

template <typename T>
static StdStore<T> plain_store(T* C, int64_t ldc, int64_t batch_stride, double alpha, double beta) {
  StdStore<T> st{};
  st.C = C;
  st.ldc = ldc;
  st.batch_stride = batch_stride;
  st.alpha = T(alpha);
  st.beta = T(beta);
  st.P = nullptr;
  st.pr = st.pc = 1;
  st.ldp = 0;
  st.padd = T(0);
  st.partial = nullptr;
  return st;
}

// C[M,N] = alpha * opA(A) opB(B)^T + beta*C with runtime transposes
template <typename T>
static int gemm_any(T* C, const T* A, const T* B, int64_t M, int64_t N, int64_t K, bool ta, bool tb,
                    int64_t lda, int64_t ldb, StdStore<T> st, bool symmetric, int64_t batch,
                    int64_t sa, int64_t sb, void* ws, int64_t wsb, cudaStream_t s, const char* what) {
  (void)C;
  if (!ta && !tb) {
    int status = VVT_OK;
    if (tc_path<T>(status, A, B, st, M, N, K, lda, ldb, symmetric, batch, ws, wsb, s, what)) return status;
    StridedLoader<T, true> la{A, lda, sa, M, K}, lb{B, ldb, sb, N, K};
    return launch_gemm_std<T>(la, lb, st, M, N, K, symmetric, batch, ws, wsb, s, what);
  } else if (!ta && tb) {
    StridedLoader<T, true> la{A, lda, sa, M, K};
    StridedLoader<T, false> lb{B, ldb, sb, N, K};
    return launch_gemm_std<T>(la, lb, st, M, N, K, symmetric, batch, ws, wsb, s, what);
  } else if (ta && !tb) {
    StridedLoader<T, false> la{A, lda, sa, M, K};
    StridedLoader<T, true> lb{B, ldb, sb, N, K};
    return launch_gemm_std<T>(la, lb, st, M, N, K, symmetric, batch, ws, wsb, s, what);
  }
  StridedLoader<T, false> la{A, lda, sa, M, K}, lb{B, ldb, sb, N, K};
  return launch_gemm_std<T>(la, lb, st, M, N, K, symmetric, batch, ws, wsb, s, what);
}

// ---- small helper kernels -------------------------------------------------

// norm2[k] += sum_d E[k,d]^2   (double accumulation, one atomic per block)
template <typename T>
__global__ void row_sumsq_kernel(double* norm2, const T* E, int64_t D) {
  const int k = blockIdx.y;
  const T* row = E + int64_t(k) * D;
  double s = 0.0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < D;
       i += int64_t(gridDim.x) * blockDim.x) {
    const double v = double(row[i]);
    s += v * v;
  }
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(norm2 + k, s);
  }
}

template <typename T>
int row_sumsq_accum(double* norm2, const T* E, int64_t K, int64_t D, cudaStream_t s) {
  if (!norm2 || K == 0 || D == 0) return VVT_OK;
  const int bx = int(vmax<int64_t>(1, vmin<int64_t>(ceil_div(D, 256 * 8), 4 * num_sms() / vmax<int64_t>(1, K) + 1)));
  row_sumsq_kernel<T><<<dim3(bx, unsigned(K)), 256, 0, s>>>(norm2, E, D);
  return launched("row_sumsq");
}
template int row_sumsq_accum<float>(double*, const float*, int64_t, int64_t, cudaStream_t);
template int row_sumsq_accum<double>(double*, const double*, int64_t, int64_t, cudaStream_t);

// T[k, n, o] = sum_c U[k, c*N + n] * S[(c*N + n), o]
template <typename T>
__global__ void contract_classes_kernel(T* Tout, const T* U, const T* S, int64_t K, int64_t C,
                                        int64_t N, int64_t n_out) {
  const int64_t total = K * N * n_out;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t o = i % n_out, n = (i / n_out) % N, k = i / (n_out * N);
    T acc = 0;
    for (int64_t c = 0; c < C; ++c) acc += ldg(U + k * C * N + c * N + n) * ldg(S + (c * N + n) * n_out + o);
    Tout[i] = acc;
  }
}

// out[f, c, n] = sum_o S[(c*N+n), o] * A[f, n, o]    (one warp per output)
template <typename T>
__global__ void rowdot_kernel(T* out, const T* S, const T* A, int64_t F, int64_t C, int64_t N,
                              int64_t n_out) {
  const int64_t total = F * C * N;
  const int lane = threadIdx.x & 31;
  for (int64_t w = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5; w < total;
       w += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int64_t n = w % N, c = (w / N) % C, f = w / (N * C);
    const T* s = S + (c * N + n) * n_out;
    const T* a = A + (f * N + n) * n_out;
    T acc = 0;
    for (int64_t o = lane; o < n_out; o += 32) acc += ldg(s + o) * ldg(a + o);
    acc = warp_sum(acc);
    if (lane == 0) out[w] = acc;
  }
}

// gammas[m,k] *= 1/sqrt(evals[k]);  lambdas[n,k] = N_ggn * corr^4 * sum_c W[(c*N+n),k]^2 / evals[k]
template <typename T>
__global__ void dirderiv_finalize_kernel(T* gammas, T* lambdas, const T* W, const T* evals,
                                         int64_t C, int64_t N_ggn, int64_t n_g, int64_t K, T lam_scale) {
  const int64_t ng = n_g * K, nl = N_ggn * K;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < ng + nl;
       i += int64_t(gridDim.x) * blockDim.x) {
    if (i < ng) {
      const int64_t k = i % K;
      gammas[i] = gammas[i] / sqrt(evals[k]);
    } else {
      const int64_t j = i - ng, k = j % K, n = j / K;
      T acc = 0;
      for (int64_t c = 0; c < C; ++c) {
        const T w = W[(c * N_ggn + n) * K + k];
        acc += w * w;
      }
      lambdas[j] = lam_scale * acc / evals[k];
    }
  }
}

static inline int ew_blocks(int64_t total) {
  return int(vmax<int64_t>(1, vmin<int64_t>(ceil_div(total, 256), 16 * num_sms())));
}

}  // namespace vvt

using namespace vvt;

extern "C" {

int vvt_gemm(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, int transA,
             int transB, int64_t lda, int64_t ldb, int64_t ldc, double alpha, double beta,
             int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC, void* workspace,
             int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative size");
  if (M == 0 || N == 0 || batch == 0) return VVT_OK;
  VVT_REQUIRE(C && A && B, "null pointer");
  VVT_DISPATCH(dtype, {
    auto st = plain_store<T>((T*)C, ldc, strideC, alpha, beta);
    return gemm_any<T>((T*)C, (const T*)A, (const T*)B, M, N, K, transA != 0, transB != 0, lda, ldb,
                       st, false, batch, strideA, strideB, workspace, workspace_bytes,
                       as_stream(stream), "vvt_gemm");
  });
}

int64_t vvt_gram_workspace_bytes(int64_t rows, int64_t cols, int64_t depth, int dtype) {
  if (rows <= 0 || cols <= 0) return 0;
  if (dtype == VVT_F32) {
    int64_t b = vmax(gemm_workspace_bytes<float>(rows, cols, depth, false),
                     gram_tc_workspace_bytes(rows, cols, depth, false));
    if (rows == cols)
      b = vmax(b, vmax(gemm_workspace_bytes<float>(rows, cols, depth, true),
                       gram_tc_workspace_bytes(rows, cols, depth, true)));
    return b;
  }
  int64_t b = gemm_workspace_bytes<double>(rows, cols, depth, false);
  if (rows == cols) b = vmax(b, gemm_workspace_bytes<double>(rows, cols, depth, true));
  return b;
}

int vvt_gram_dense_accum(void* G, const void* V, int64_t R, int64_t D, void* workspace,
                         int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0 && D >= 0, "negative size");
  if (R == 0 || D == 0) return VVT_OK;
  VVT_REQUIRE(G && V, "null pointer");
  VVT_DISPATCH(dtype, {
    auto st = plain_store<T>((T*)G, R, 0, 1.0, 1.0);
    return gemm_any<T>((T*)G, (const T*)V, (const T*)V, R, R, D, false, false, D, D, st, true, 1, 0,
                       0, workspace, workspace_bytes, as_stream(stream), "vvt_gram_dense_accum");
  });
}

int vvt_gram_cross_accum(void* X, const void* V, const void* g, int64_t R, int64_t n_g, int64_t D,
                         void* workspace, int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(R >= 0 && D >= 0 && n_g >= 0, "negative size");
  if (R == 0 || D == 0 || n_g == 0) return VVT_OK;
  VVT_REQUIRE(X && V && g, "null pointer");
  VVT_DISPATCH(dtype, {
    auto st = plain_store<T>((T*)X, n_g, 0, 1.0, 1.0);
    return gemm_any<T>((T*)X, (const T*)V, (const T*)g, R, n_g, D, false, false, D, D, st, false, 1,
                       0, 0, workspace, workspace_bytes, as_stream(stream), "vvt_gram_cross_accum");
  });
}

static int64_t p_region_bytes(int64_t N, int64_t n_g, int dtype) {
  const int64_t es = dtype == VVT_F32 ? 4 : 8;
  return align_up(N * (n_g > N ? n_g : N) * es, 256);
}

int64_t vvt_gram_linear_workspace_bytes(int64_t C, int64_t N, int64_t n_out, int64_t n_in,
                                        int64_t n_g, int dtype) {
  const int64_t R = C * N;
  const int64_t cols = n_g > 0 ? n_g : R;
  int64_t a = vvt_gram_workspace_bytes(N, n_g > 0 ? n_g : N, n_in, dtype);
  int64_t b = vvt_gram_workspace_bytes(R, cols, n_out, dtype);
  return p_region_bytes(N, n_g, dtype) + (a > b ? a : b);
}

int vvt_gram_linear_accum(void* G, const void* S, const void* Z, int64_t C, int64_t N,
                          int64_t n_out, int64_t n_in, int with_bias, void* workspace,
                          int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(C >= 0 && N >= 0 && n_out >= 0 && n_in >= 0, "negative size");
  const int64_t R = C * N;
  if (R == 0 || n_out == 0) return VVT_OK;
  VVT_REQUIRE(G && S && Z && workspace, "null pointer");
  const int64_t pbytes = p_region_bytes(N, 0, dtype);
  if (workspace_bytes < pbytes) return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  void* ws2 = (char*)workspace + pbytes;
  const int64_t ws2b = workspace_bytes - pbytes;
  VVT_DISPATCH(dtype, {
    T* P = (T*)workspace;
    // P = Z Z^T  [N, N]
    auto stp = plain_store<T>(P, N, 0, 1.0, 0.0);
    if (n_in > 0) {
      VVT_TRY(gemm_any<T>(P, (const T*)Z, (const T*)Z, N, N, n_in, false, false, n_in, n_in, stp, true,
                          1, 0, 0, ws2, ws2b, as_stream(stream), "vvt_gram_linear_accum(ZZt)"));
    } else {
      VVT_TRY(check_cuda(cudaMemsetAsync(P, 0, N * N * sizeof(T), as_stream(stream)), __func__));
    }
    // G += (P + bias) (.) S S^T
    auto st = plain_store<T>((T*)G, R, 0, 1.0, 1.0);
    st.P = P;
    st.pr = N;
    st.pc = N;
    st.ldp = N;
    st.padd = T(with_bias ? 1 : 0);
    return gemm_any<T>((T*)G, (const T*)S, (const T*)S, R, R, n_out, false, false, n_out, n_out, st,
                       true, 1, 0, 0, ws2, ws2b, as_stream(stream), "vvt_gram_linear_accum");
  });
}

int vvt_gram_cross_linear_accum(void* X, const void* S, const void* Z, const void* Dl,
                                const void* Zg, int64_t C, int64_t N, int64_t n_g, int64_t n_out,
                                int64_t n_in, int with_bias, void* workspace,
                                int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(C >= 0 && N >= 0 && n_out >= 0 && n_in >= 0 && n_g >= 0, "negative size");
  const int64_t R = C * N;
  if (R == 0 || n_out == 0 || n_g == 0) return VVT_OK;
  VVT_REQUIRE(X && S && Z && Dl && Zg && workspace, "null pointer");
  const int64_t pbytes = p_region_bytes(N, n_g, dtype);
  if (workspace_bytes < pbytes) return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  void* ws2 = (char*)workspace + pbytes;
  const int64_t ws2b = workspace_bytes - pbytes;
  VVT_DISPATCH(dtype, {
    T* P = (T*)workspace;  // Z Zg^T  [N, n_g]
    auto stp = plain_store<T>(P, n_g, 0, 1.0, 0.0);
    if (n_in > 0) {
      VVT_TRY(gemm_any<T>(P, (const T*)Z, (const T*)Zg, N, n_g, n_in, false, false, n_in, n_in, stp,
                          false, 1, 0, 0, ws2, ws2b, as_stream(stream),
                          "vvt_gram_cross_linear_accum(ZZg)"));
    } else {
      VVT_TRY(check_cuda(cudaMemsetAsync(P, 0, N * n_g * sizeof(T), as_stream(stream)), __func__));
    }
    auto st = plain_store<T>((T*)X, n_g, 0, 1.0, 1.0);
    st.P = P;
    st.pr = N;
    st.pc = n_g;
    st.ldp = n_g;
    st.padd = T(with_bias ? 1 : 0);
    return gemm_any<T>((T*)X, (const T*)S, (const T*)Dl, R, n_g, n_out, false, false, n_out, n_out,
                       st, false, 1, 0, 0, ws2, ws2b, as_stream(stream),
                       "vvt_gram_cross_linear_accum");
  });
}

int vvt_sqrt_backprop_linear(void* out, const void* S, const void* W, int64_t rows, int64_t n_out,
                             int64_t n_in, int dtype, void* stream) {
  VVT_REQUIRE(rows >= 0 && n_out >= 0 && n_in >= 0, "negative size");
  if (rows == 0 || n_in == 0) return VVT_OK;
  VVT_REQUIRE(out && S && W, "null pointer");
  VVT_DISPATCH(dtype, {
    auto st = plain_store<T>((T*)out, n_in, 0, 1.0, 0.0);
    // out[r,i] = sum_o S[r,o] W[o,i]: B given as [K=n_out, N=n_in]
    return gemm_any<T>((T*)out, (const T*)S, (const T*)W, rows, n_in, n_out, false, true, n_out,
                       n_in, st, false, 1, 0, 0, nullptr, 0, as_stream(stream),
                       "vvt_sqrt_backprop_linear");
  });
}

int vvt_backtransform_linear(void* E, void* Eb, void* norm2, const void* U, const void* S,
                             const void* Z, int64_t K, int64_t C, int64_t N, int64_t n_out,
                             int64_t n_in, void* workspace, int64_t workspace_bytes, int dtype,
                             void* stream) {
  VVT_REQUIRE(K >= 0 && C >= 0 && N >= 0 && n_out >= 0 && n_in >= 0, "negative size");
  if (K == 0 || n_out == 0) return VVT_OK;
  VVT_REQUIRE(E && U && S && Z && workspace, "null pointer");
  VVT_REQUIRE(Eb == nullptr, "fused bias output not implemented; pass NULL");
  const int64_t es = dtype == VVT_F32 ? 4 : 8;
  if (workspace_bytes < K * N * n_out * es)
    return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  cudaStream_t s = as_stream(stream);
  VVT_DISPATCH(dtype, {
    T* Tm = (T*)workspace;  // [K, N, n_out]
    contract_classes_kernel<T><<<ew_blocks(K * N * n_out), 256, 0, s>>>(Tm, (const T*)U, (const T*)S,
                                                                     K, C, N, n_out);
    VVT_TRY(launched("vvt_backtransform_linear(contract)"));
    if (n_in > 0) {
      // E[k] [n_out, n_in] = Tm[k]^T [n_out x N] * Z [N x n_in]
      auto st = plain_store<T>((T*)E, n_in, n_out * n_in, 1.0, 0.0);
      VVT_TRY(gemm_any<T>((T*)E, Tm, (const T*)Z, n_out, n_in, N, true, true, n_out, n_in, st, false,
                          K, N * n_out, 0, nullptr, 0, s, "vvt_backtransform_linear"));
      VVT_TRY(row_sumsq_accum<T>((double*)norm2, (const T*)E, K, n_out * n_in, s));
    }
    return VVT_OK;
  });
}

int vvt_v_apply_linear(void* step, void* stepb, const void* v, const void* S, const void* Z,
                       int64_t C, int64_t N, int64_t n_out, int64_t n_in, void* workspace,
                       int64_t workspace_bytes, int dtype, void* stream) {
  return vvt_backtransform_linear(step, stepb, nullptr, v, S, Z, 1, C, N, n_out, n_in, workspace,
                                  workspace_bytes, dtype, stream);
}

int vvt_vt_mat_prod_linear(void* out, const void* S, const void* Z, const void* Mat, int64_t F,
                           int64_t C, int64_t N, int64_t n_out, int64_t n_in, void* workspace,
                           int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(F >= 0 && C >= 0 && N >= 0 && n_out >= 0 && n_in >= 0, "negative size");
  if (F == 0 || C * N == 0) return VVT_OK;
  VVT_REQUIRE(out && S && Z && Mat && workspace, "null pointer");
  const int64_t es = dtype == VVT_F32 ? 4 : 8;
  if (workspace_bytes < F * N * n_out * es)
    return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  cudaStream_t s = as_stream(stream);
  VVT_DISPATCH(dtype, {
    T* A = (T*)workspace;  // A[f, n, o] = sum_i Z[n,i] Mat[f,o,i]
    auto st = plain_store<T>(A, n_out, N * n_out, 1.0, 0.0);
    VVT_TRY(gemm_any<T>(A, (const T*)Z, (const T*)Mat, N, n_out, n_in, false, false, n_in, n_in, st,
                        false, F, 0, n_out * n_in, nullptr, 0, s, "vvt_vt_mat_prod_linear"));
    const int64_t warps = F * C * N;
    rowdot_kernel<T><<<ew_blocks(warps * 32), 256, 0, s>>>((T*)out, (const T*)S, A, F, C, N, n_out);
    return launched("vvt_vt_mat_prod_linear(rowdot)");
  });
}

int vvt_dirderiv_epilogue(void* gammas, void* lambdas, const void* G, const void* X, const void* U,
                          const void* evals, int64_t C, int64_t N_ggn, int64_t n_g, int64_t K,
                          int64_t N, void* workspace, int64_t workspace_bytes, int dtype,
                          void* stream) {
  VVT_REQUIRE(C >= 0 && N_ggn > 0 && n_g >= 0 && K >= 0 && N > 0, "bad size");
  if (K == 0) return VVT_OK;
  const int64_t R = C * N_ggn;
  VVT_REQUIRE(gammas && lambdas && G && X && U && evals && workspace, "null pointer");
  const int64_t es = dtype == VVT_F32 ? 4 : 8;
  const int64_t wbytes = align_up(R * K * es, 256);
  if (workspace_bytes < wbytes) return fail(VVT_ERR_WORKSPACE, "%s: workspace too small", __func__);
  void* ws2 = (char*)workspace + wbytes;
  const int64_t ws2b = workspace_bytes - wbytes;
  cudaStream_t s = as_stream(stream);
  const double corr2 = double(N) / double(N_ggn);
  const double corr = sqrt(corr2);
  VVT_DISPATCH(dtype, {
    T* W = (T*)workspace;  // W = G U   [R, K]   (un-rescaled G)
    auto stw = plain_store<T>(W, K, 0, 1.0, 0.0);
    VVT_TRY(gemm_any<T>(W, (const T*)G, (const T*)U, R, K, R, false, true, R, K, stw, false, 1, 0, 0,
                        ws2, ws2b, s, "vvt_dirderiv_epilogue(GU)"));
    if (n_g > 0) {
      // gammas = corr*N * X^T U   [n_g, K]
      auto stg = plain_store<T>((T*)gammas, K, 0, corr * double(N), 0.0);
      VVT_TRY(gemm_any<T>((T*)gammas, (const T*)X, (const T*)U, n_g, K, R, true, true, n_g, K, stg,
                          false, 1, 0, 0, ws2, ws2b, s, "vvt_dirderiv_epilogue(XtU)"));
    }
    const T lam_scale = T(double(N_ggn) * corr2 * corr2);
    dirderiv_finalize_kernel<T><<<ew_blocks((n_g + N_ggn) * K), 256, 0, s>>>(
        (T*)gammas, (T*)lambdas, W, (const T*)evals, C, N_ggn, n_g, K, lam_scale);
    return launched("vvt_dirderiv_epilogue(finalize)");
  });
}

}  // extern "C"
