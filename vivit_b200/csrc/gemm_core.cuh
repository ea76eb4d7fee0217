// Tiled GEMM main loop with pluggable operand loaders and result stores.
//
//   acc[m, n] = sum_k A(m, k) * B(n, k)
//
// float : 3xTF32 split on mma.sync.m16n8k8 (hi*hi + hi*lo + lo*hi, fp32 accumulate) -> fp32-grade accuracy
// double: mma.sync.m8n8k4 DMMA
//
// Operands are fetched through functors so the same main loop serves plain
// strided matrices, the per-sample implicit GEMM of the conv V-emit and the
// conv data-gradient.  Loaders return 0 outside their bounds; stores are only
// invoked for in-range (row, col).
#pragma once
#include "common.cuh"

namespace vvt {

template <typename T>
struct MmaCfg;

template <>
struct MmaCfg<float> {
  static constexpr int BM = 128, BN = 128, BK = 32, LDS = 36;
  static constexpr int WARPS_M = 2, WARPS_N = 4, WM = 64, WN = 32;
  static constexpr int MT = 4, NT = 4;  // 16x8 mma tiles per warp
};
template <>
struct MmaCfg<double> {
  static constexpr int BM = 64, BN = 64, BK = 16, LDS = 20;
  static constexpr int WARPS_M = 2, WARPS_N = 4, WM = 32, WN = 16;
  static constexpr int MT = 4, NT = 2;  // 8x8 mma tiles per warp
};

constexpr int kGemmThreads = 256;

template <typename T>
constexpr size_t gemm_smem_bytes() {
  return size_t(2) * (MmaCfg<T>::BM + MmaCfg<T>::BN) * MmaCfg<T>::LDS * sizeof(T);
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// 3xTF32 split  x ~= hi + lo  without conversion instructions.  The tensor core reads a tf32 operand by
// ignoring the 13 low mantissa bits of the fp32 word, so the word itself serves as "hi"; lo = x - trunc(x)
// is exact in fp32, and adding half a tf32 ulp to its bit pattern turns the hardware's truncation of lo into
// round-to-nearest (ties away, as cvt.rna).  3 ALU ops per element instead of ~10 for two cvt.rna.tf32.f32
// (which ptxas expands to compare / add / select / mask on sm_100).  VVT_TF32_TRUNC=0 restores the cvt form.
#ifndef VVT_TF32_TRUNC
#define VVT_TF32_TRUNC 1
#endif
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#if VVT_TF32_TRUNC
  const uint32_t b = __float_as_uint(x);
  hi = b;
  lo = __float_as_uint(x - __uint_as_float(b & 0xFFFFE000u)) + 0x1000u;
#else
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
#endif
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_f64(double (&d)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d[0]), "+d"(d[1])
      : "d"(a), "d"(b));
}

// ---- per-type warp-level compute on one smem k-tile ----------------------
template <typename T>
struct WarpMma;

template <>
struct WarpMma<float> {
  using Cfg = MmaCfg<float>;
  float acc[Cfg::MT][Cfg::NT][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < Cfg::MT; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
  }
  __device__ __forceinline__ void tile(const float* As, const float* Bs, int wm, int wn, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kk = 0; kk < Cfg::BK / 8; ++kk) {
      uint32_t ahi[Cfg::MT][4], alo[Cfg::MT][4], bhi[Cfg::NT][2], blo[Cfg::NT][2];
#pragma unroll
      for (int mi = 0; mi < Cfg::MT; ++mi) {
        const float* p = As + (wm * Cfg::WM + mi * 16 + g) * Cfg::LDS + kk * 8 + t;
        float a[4] = {p[0], p[8 * Cfg::LDS], p[4], p[8 * Cfg::LDS + 4]};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          split_tf32(a[r], ahi[mi][r], alo[mi][r]);
        }
      }
#pragma unroll
      for (int ni = 0; ni < Cfg::NT; ++ni) {
        const float* p = Bs + (wn * Cfg::WN + ni * 8 + g) * Cfg::LDS + kk * 8 + t;
        float b[2] = {p[0], p[4]};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          split_tf32(b[r], bhi[ni][r], blo[ni][r]);
        }
      }
#pragma unroll
      for (int mi = 0; mi < Cfg::MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < Cfg::NT; ++ni) {
          mma_tf32(acc[mi][ni], alo[mi], bhi[ni]);
          mma_tf32(acc[mi][ni], ahi[mi], blo[ni]);
          mma_tf32(acc[mi][ni], ahi[mi], bhi[ni]);
        }
    }
  }
  // visit every accumulator with its (row, col) inside the CTA tile
  template <typename F>
  __device__ __forceinline__ void for_each(int wm, int wn, int lane, F f) const {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < Cfg::MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < Cfg::NT; ++ni) {
        const int r = wm * Cfg::WM + mi * 16 + g, c = wn * Cfg::WN + ni * 8 + 2 * t;
        f(r, c, acc[mi][ni][0]);
        f(r, c + 1, acc[mi][ni][1]);
        f(r + 8, c, acc[mi][ni][2]);
        f(r + 8, c + 1, acc[mi][ni][3]);
      }
  }
};

template <>
struct WarpMma<double> {
  using Cfg = MmaCfg<double>;
  double acc[Cfg::MT][Cfg::NT][2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < Cfg::MT; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }
  __device__ __forceinline__ void tile(const double* As, const double* Bs, int wm, int wn, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int kk = 0; kk < Cfg::BK / 4; ++kk) {
      double a[Cfg::MT], b[Cfg::NT];
#pragma unroll
      for (int mi = 0; mi < Cfg::MT; ++mi) a[mi] = As[(wm * Cfg::WM + mi * 8 + g) * Cfg::LDS + kk * 4 + t];
#pragma unroll
      for (int ni = 0; ni < Cfg::NT; ++ni) b[ni] = Bs[(wn * Cfg::WN + ni * 8 + g) * Cfg::LDS + kk * 4 + t];
#pragma unroll
      for (int mi = 0; mi < Cfg::MT; ++mi)
#pragma unroll
        for (int ni = 0; ni < Cfg::NT; ++ni) mma_f64(acc[mi][ni], a[mi], b[ni]);
    }
  }
  template <typename F>
  __device__ __forceinline__ void for_each(int wm, int wn, int lane, F f) const {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mi = 0; mi < Cfg::MT; ++mi)
#pragma unroll
      for (int ni = 0; ni < Cfg::NT; ++ni) {
        const int r = wm * Cfg::WM + mi * 8 + g, c = wn * Cfg::WN + ni * 8 + 2 * t;
        f(r, c, acc[mi][ni][0]);
        f(r, c + 1, acc[mi][ni][1]);
      }
  }
};

// ---- generic kernel -------------------------------------------------------
// LA/LB: T operator()(int batch, int64_t row, int64_t k) const, static constexpr bool kContigK
// ST   : void operator()(int batch, int64_t row, int64_t col, T v, int split) const
template <typename T, typename LA, typename LB, typename ST>
__global__ void __launch_bounds__(kGemmThreads)
gemm_kernel(LA la, LB lb, ST st, int64_t M, int64_t N, int64_t K, int tiles_n, int symmetric,
            int64_t k_per_split) {
  using Cfg = MmaCfg<T>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, LDS = Cfg::LDS;
  constexpr int EA = BM * BK / kGemmThreads, EB = BN * BK / kGemmThreads;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* As = reinterpret_cast<T*>(smem_raw);
  T* Bs = As + 2 * BM * LDS;

  int tm, tn;
  if (symmetric) {  // upper-triangular tile pairs, row by row
    int rem = blockIdx.x, len = tiles_n;
    tm = 0;
    while (rem >= len) {
      rem -= len;
      ++tm;
      --len;
    }
    tn = tm + rem;
  } else {
    tm = blockIdx.x / tiles_n;
    tn = blockIdx.x % tiles_n;
  }
  const int batch = blockIdx.z, split = blockIdx.y;
  const int64_t m0 = int64_t(tm) * BM, n0 = int64_t(tn) * BN;
  const int64_t k_begin = int64_t(split) * k_per_split;
  const int64_t k_end = min(K, k_begin + k_per_split);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;

  T ra[EA], rb[EB];
  auto gload = [&](int64_t kt) {
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = tid + e * kGemmThreads;
      const int r = LA::kContigK ? idx / BK : idx % BM;
      const int k = LA::kContigK ? idx % BK : idx / BM;
      ra[e] = (kt + k < k_end) ? la(batch, m0 + r, kt + k) : T(0);
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int idx = tid + e * kGemmThreads;
      const int r = LB::kContigK ? idx / BK : idx % BN;
      const int k = LB::kContigK ? idx % BK : idx / BN;
      rb[e] = (kt + k < k_end) ? lb(batch, n0 + r, kt + k) : T(0);
    }
  };
  auto sstore = [&](int buf) {
    T* a = As + buf * BM * LDS;
    T* b = Bs + buf * BN * LDS;
#pragma unroll
    for (int e = 0; e < EA; ++e) {
      const int idx = tid + e * kGemmThreads;
      const int r = LA::kContigK ? idx / BK : idx % BM;
      const int k = LA::kContigK ? idx % BK : idx / BM;
      a[r * LDS + k] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int idx = tid + e * kGemmThreads;
      const int r = LB::kContigK ? idx / BK : idx % BN;
      const int k = LB::kContigK ? idx % BK : idx / BN;
      b[r * LDS + k] = rb[e];
    }
  };

  WarpMma<T> mma;
  mma.clear();
  if (k_begin < k_end) {
    gload(k_begin);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int64_t kt = k_begin; kt < k_end; kt += BK, buf ^= 1) {
      const bool more = kt + BK < k_end;
      if (more) gload(kt + BK);
      mma.tile(As + buf * BM * LDS, Bs + buf * BN * LDS, wm, wn, lane);
      if (more) sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const bool mirror = symmetric && tm != tn;
  mma.for_each(wm, wn, lane, [&](int r, int c, T v) {
    const int64_t row = m0 + r, col = n0 + c;
    if (row < M && col < N) {
      st(batch, row, col, v, split);
      if (mirror) st(batch, col, row, v, split);
    }
  });
}

// ---- standard loaders / stores -------------------------------------------
template <typename T, bool CONTIG_K>
struct StridedLoader {
  static constexpr bool kContigK = CONTIG_K;
  const T* p;
  int64_t ld, batch_stride, rows, depth;
  __device__ __forceinline__ T operator()(int b, int64_t r, int64_t k) const {
    if (r >= rows || k >= depth) return T(0);
    const T* q = p + b * batch_stride;
    return CONTIG_K ? ldg(q + r * ld + k) : ldg(q + k * ld + r);
  }
};

// C = beta*C + alpha * acc * (P ? P[(row % pr), (col % pc)] + padd : 1);  or raw partial for split-K
template <typename T>
struct StdStore {
  T* C;
  int64_t ldc, batch_stride;
  T alpha, beta;
  const T* P;
  int64_t pr, pc, ldp;
  T padd;
  T* partial;    // non-null => split-K partial slabs [splits][M*N]
  int64_t slab, N;
  // index of the Hadamard factor: r % pr, c % pc -- in 32-bit arithmetic whenever the numbers allow it (a 64-bit
  // remainder is ~60 instructions; two of them per output element were most of the epilogue of the structured
  // Linear Gram at R = 10240)
  __device__ __forceinline__ static int64_t wrap(int64_t x, int64_t m) {
    return ((x | m) >> 32) == 0 ? int64_t(uint32_t(x) % uint32_t(m)) : x % m;
  }
  __device__ __forceinline__ T weight(int64_t r, int64_t c) const {
    return P ? ldg(P + wrap(r, pr) * ldp + wrap(c, pc)) + padd : T(1);
  }
  // forms used by the tcgen05 epilogue: row_offset (the row of the Hadamard factor, computed once per output row),
  // and the accumulate split into fetch (the old value of C, if beta needs it) and commit, so that a batch of
  // independent loads can be in flight before the first store (one dependent global load per element made small-K
  // products -- the Cholesky trailing updates -- latency bound: 75 us per launch)
  __device__ __forceinline__ int64_t row_offset(int, int64_t r) const { return P ? wrap(r, pr) * ldp : 0; }
  __device__ __forceinline__ T fetch(int64_t, int64_t r, int64_t c) const {
    return (partial || beta == T(0)) ? T(0) : C[r * ldc + c];
  }
  __device__ __forceinline__ void commit(int64_t prow, int64_t r, int64_t c, T v, T old, int split) const {
    if (partial) {
      partial[int64_t(split) * slab + r * N + c] = v;
      return;
    }
    const T w = P ? ldg(P + prow + wrap(c, pc)) + padd : T(1);
    C[r * ldc + c] = beta * old + alpha * v * w;  // old = 0 when beta = 0
  }
  __device__ __forceinline__ void store(int64_t, int64_t r, int64_t c, T v, int split) const {
    (*this)(0, r, c, v, split);
  }
  __device__ __forceinline__ void operator()(int b, int64_t r, int64_t c, T v, int split) const {
    if (partial) {
      partial[int64_t(split) * slab + r * N + c] = v;
      return;
    }
    T* dst = C + b * batch_stride + r * ldc + c;
    const T x = alpha * v * weight(r, c);
    *dst = (beta == T(0)) ? x : beta * (*dst) + x;
  }
};

template <typename T>
__global__ void splitk_reduce_kernel(StdStore<T> st, int64_t M, int64_t N, int splits) {
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    T s = 0;
    for (int k = 0; k < splits; ++k) s += st.partial[int64_t(k) * st.slab + i];
    const int64_t r = i / N, c = i % N;
    T* dst = st.C + r * st.ldc + c;
    const T x = st.alpha * s * st.weight(r, c);
    *dst = (st.beta == T(0)) ? x : st.beta * (*dst) + x;
  }
}

// ---- host-side launch -------------------------------------------------------
struct GemmPlan {
  int tiles_m, tiles_n, tiles, splits;
  int64_t k_per_split;
};

template <typename T>
inline GemmPlan plan_gemm(int64_t M, int64_t N, int64_t K, bool symmetric, int64_t batch,
                          int64_t workspace_bytes) {
  using Cfg = MmaCfg<T>;
  GemmPlan p;
  p.tiles_m = int(ceil_div(M, Cfg::BM));
  p.tiles_n = int(ceil_div(N, Cfg::BN));
  p.tiles = symmetric ? p.tiles_n * (p.tiles_n + 1) / 2 : p.tiles_m * p.tiles_n;
  const int64_t ksteps = ceil_div(K, Cfg::BK);
  int64_t splits = 1;
  if (batch == 1) {
    const int64_t want = ceil_div(2 * int64_t(num_sms()), p.tiles);
    splits = min(want, vmax<int64_t>(1, ksteps / 4));
    // fp32: an mma.sync accumulator is updated with truncation, a bias that grows with the length of the
    // contraction (3.5e-5 of the largest eigenvalue through the refinement GEMMs of an R = 4607 solve); slabs of at
    // most 2048 are summed with round-to-nearest adds by the reduction kernel, as the tcgen05 kernel promotes its
    // accumulator groups
    if (sizeof(T) == 4) splits = vmax<int64_t>(splits, ceil_div(K, 2048));
    splits = vmin<int64_t>(splits, 64);
    const int64_t fit = workspace_bytes / vmax<int64_t>(1, M * N * int64_t(sizeof(T)));
    splits = vmax<int64_t>(1, min(splits, fit));
  }
  int64_t steps_per = ceil_div(ksteps, splits);
  p.k_per_split = steps_per * Cfg::BK;
  p.splits = int(ceil_div(ksteps, steps_per));
  if (p.splits < 1) p.splits = 1;
  return p;
}

template <typename T>
inline int64_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, bool symmetric) {
  GemmPlan p = plan_gemm<T>(M, N, K, symmetric, 1, INT64_MAX / 4);
  return p.splits > 1 ? int64_t(p.splits) * M * N * int64_t(sizeof(T)) : 0;
}

// Launch with arbitrary loaders and an arbitrary store functor (no split-K).
template <typename T, typename LA, typename LB, typename ST>
inline int launch_gemm_custom(LA la, LB lb, ST st, int64_t M, int64_t N, int64_t K, int64_t batch,
                              cudaStream_t stream, const char* what) {
  using Cfg = MmaCfg<T>;
  if (M <= 0 || N <= 0 || batch <= 0) return VVT_OK;
  auto kern = gemm_kernel<T, LA, LB, ST>;
  static SmemOptIn opt_in;  // per template instantiation
  VVT_TRY(opt_in.ensure(kern, gemm_smem_bytes<T>(), what));
  const int tiles_m = int(ceil_div(M, Cfg::BM)), tiles_n = int(ceil_div(N, Cfg::BN));
  const int64_t k_all = align_up(vmax<int64_t>(K, 1), Cfg::BK);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {  // gridDim.z limit
    const unsigned nb = unsigned(vmin<int64_t>(65535, batch - b0));
    if (b0 != 0) return fail(VVT_ERR_UNSUPPORTED, "%s: batch > 65535", what);
    dim3 grid(tiles_m * tiles_n, 1, nb);
    kern<<<grid, kGemmThreads, gemm_smem_bytes<T>(), stream>>>(la, lb, st, M, N, K, tiles_n, 0, k_all);
    VVT_TRY(launched(what));
  }
  return VVT_OK;
}

// Launch with arbitrary loaders and a StdStore epilogue (split-K handled here).
template <typename T, typename LA, typename LB>
inline int launch_gemm_std(LA la, LB lb, StdStore<T> st, int64_t M, int64_t N, int64_t K,
                           bool symmetric, int64_t batch, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream, const char* what) {
  if (M <= 0 || N <= 0 || batch <= 0) return VVT_OK;
  GemmPlan p = plan_gemm<T>(M, N, K, symmetric, batch, workspace ? workspace_bytes : 0);
  auto kern = gemm_kernel<T, LA, LB, StdStore<T>>;
  static SmemOptIn opt_in;  // per template instantiation
  VVT_TRY(opt_in.ensure(kern, gemm_smem_bytes<T>(), what));
  st.N = N;
  st.slab = M * N;
  StdStore<T> kst = st;
  kst.partial = p.splits > 1 ? reinterpret_cast<T*>(workspace) : nullptr;
  dim3 grid(p.tiles, p.splits, unsigned(batch));
  kern<<<grid, kGemmThreads, gemm_smem_bytes<T>(), stream>>>(la, lb, kst, M, N, K, p.tiles_n,
                                                           symmetric ? 1 : 0, p.k_per_split);
  VVT_TRY(launched(what));
  if (p.splits > 1) {
    const int64_t total = M * N;
    const int blocks = int(vmin<int64_t>(ceil_div(total, 256), 8 * num_sms()));
    splitk_reduce_kernel<T><<<blocks, 256, 0, stream>>>(kst, M, N, p.splits);
    VVT_TRY(launched(what));
  }
  return VVT_OK;
}

}  // namespace vvt
