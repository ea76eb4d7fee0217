// Conv2d pieces of the sqrt-GGN factor as implicit GEMMs on the shared main loop:
//   * V emit: per-sample  Vt[(v,n), (o, j)] = sum_x S[v,n,o,x] * patch[n, j, x]
//   * data gradient of the factor ("virtual batch" of V*N maps)
#include <cstdlib>

#include "gemm_smallk.cuh"

namespace vvt {

struct ConvGeom {
  int c_out, h_out, w_out, c_in, h_in, w_in, kh, kw, sh, sw, ph, pw, dh, dw;
};

// ---- V emit -----------------------------------------------------------------
// batch index = sample n;  row m = (v, o);  col j = (ci, ky, kx);  k = x = (oy, ox)
template <typename T>
struct EmitLoaderS {
  static constexpr bool kContigK = true;
  const T* S;
  int64_t N, c_out, X, M;
  __device__ __forceinline__ T operator()(int n, int64_t m, int64_t x) const {
    if (m >= M || x >= X) return T(0);
    const int64_t v = m / c_out, o = m % c_out;
    return ldg(S + ((v * N + n) * c_out + o) * X + x);
  }
};

template <typename T>
struct EmitLoaderPatch {
  static constexpr bool kContigK = true;
  const T* Xin;
  ConvGeom g;
  int64_t J, X;
  __device__ __forceinline__ T operator()(int n, int64_t j, int64_t x) const {
    if (j >= J || x >= X) return T(0);
    const int kx = int(j % g.kw), ky = int((j / g.kw) % g.kh), ci = int(j / (g.kw * g.kh));
    const int ox = int(x % g.w_out), oy = int(x / g.w_out);
    const int iy = oy * g.sh + ky * g.dh - g.ph, ix = ox * g.sw + kx * g.dw - g.pw;
    if (iy < 0 || iy >= g.h_in || ix < 0 || ix >= g.w_in) return T(0);
    return ldg(Xin + ((int64_t(n) * g.c_in + ci) * g.h_in + iy) * g.w_in + ix);
  }
};

template <typename T>
struct EmitStore {
  T* Vt;
  int64_t N, c_out, J;
  __device__ __forceinline__ void operator()(int n, int64_t m, int64_t j, T val, int) const {
    const int64_t v = m / c_out, o = m % c_out;
    Vt[((v * N + n) * c_out + o) * J + j] = val;
  }
};

// ---- data gradient -------------------------------------------------------------
// row m = (r, y, x) input position;  col = ci;  k = (o, ky, kx)
template <typename T>
struct DgradLoaderS {
  static constexpr bool kContigK = false;  // consecutive rows (x) are contiguous in S for stride 1
  const T* S;
  ConvGeom g;
  int64_t M, Kd;
  __device__ __forceinline__ T operator()(int, int64_t m, int64_t k) const {
    if (m >= M || k >= Kd) return T(0);
    const int hw = g.h_in * g.w_in;
    const int64_t r = m / hw;
    const int p = int(m % hw), y = p / g.w_in, x = p % g.w_in;
    const int kx = int(k % g.kw), ky = int((k / g.kw) % g.kh), o = int(k / (g.kw * g.kh));
    const int ty = y + g.ph - ky * g.dh, tx = x + g.pw - kx * g.dw;
    if (ty < 0 || tx < 0 || ty % g.sh || tx % g.sw) return T(0);
    const int oy = ty / g.sh, ox = tx / g.sw;
    if (oy >= g.h_out || ox >= g.w_out) return T(0);
    return ldg(S + ((r * g.c_out + o) * g.h_out + oy) * g.w_out + ox);
  }
};

template <typename T>
struct DgradLoaderW {
  static constexpr bool kContigK = true;
  const T* W;
  ConvGeom g;
  int64_t Kd;
  __device__ __forceinline__ T operator()(int, int64_t ci, int64_t k) const {
    if (ci >= g.c_in || k >= Kd) return T(0);
    const int khw = g.kh * g.kw;
    const int64_t o = k / khw, kk = k % khw;
    return ldg(W + (o * g.c_in + ci) * khw + kk);
  }
};

template <typename T>
struct DgradStore {
  T* out;
  int64_t c_in, hw;
  __device__ __forceinline__ void operator()(int, int64_t m, int64_t ci, T val, int) const {
    const int64_t r = m / hw, p = m % hw;
    out[(r * c_in + ci) * hw + p] = val;
  }
};


// ---- fp32 fast path: bandwidth-bound re-layout kernels around the batched tcgen05 GEMM ---------------
// HBM is plentiful on B200 (180 GB, ~6.5 TB/s), so the convolution is made explicit: patches are
// unfolded, the factor is re-laid out so that the contraction index is contiguous (what TMA + UMMA
// want), and every flop runs in gemm_tc.cuh.  Each helper reads and writes its operands once.

#define VVT_GRID_STRIDE(i, total)                                              \
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < (total); \
       i += int64_t(gridDim.x) * blockDim.x)

static inline int ew_blocks(int64_t total) {
  return int(vmax<int64_t>(1, vmin<int64_t>(ceil_div(total, 256), 32 * num_sms())));
}

// Sn[n][(v, o)][x] (row pitch Xp) <- S[v][n][o][x]
__global__ void emit_permute_kernel(float* Sn, const float* S, int V, int N, int CoX_rows, int X, int Xp) {
  const int64_t total = int64_t(N) * V * CoX_rows * Xp;
  VVT_GRID_STRIDE(i, total) {
    const int x = int(i % Xp);
    int64_t t = i / Xp;
    const int o = int(t % CoX_rows);
    t /= CoX_rows;
    const int v = int(t % V), n = int(t / V);
    Sn[i] = x < X ? ldg(S + ((int64_t(v) * N + n) * CoX_rows + o) * X + x) : 0.f;
  }
}

// Un[n][j = (ci, ky, kx)][x = (oy, ox)] (row pitch Xp) <- unfolded input patches
// One block row per (n, j) patch row, threads over x: the decomposition of j is done once per row and the index
// arithmetic of an element is two 32-bit divisions (the flat 64-bit form spent more time in divisions than in
// memory instructions: 0.51 ms per c2 step for 0.9 GB of traffic).
__global__ void __launch_bounds__(256) im2col_kernel(float* Un, const float* Xin, ConvGeom g, int N, int J, int X, int Xp) {
  const int64_t rows = int64_t(N) * J;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int j = int(row % J), n = int(row / J);
    const int kx = j % g.kw, ky = (j / g.kw) % g.kh, ci = j / (g.kw * g.kh);
    const float* src = Xin + (int64_t(n) * g.c_in + ci) * g.h_in * g.w_in;
    float* dst = Un + row * Xp;
    const int y0 = ky * g.dh - g.ph, x0 = kx * g.dw - g.pw;
    for (int x = threadIdx.x; x < Xp; x += blockDim.x) {
      float val = 0.f;
      if (x < X) {
        const int oy = x / g.w_out, ox = x - oy * g.w_out;
        const int iy = oy * g.sh + y0, ix = ox * g.sw + x0;
        if (unsigned(iy) < unsigned(g.h_in) && unsigned(ix) < unsigned(g.w_in)) val = ldg(src + iy * g.w_in + ix);
      }
      dst[x] = val;
    }
  }
}

struct EmitStoreTc {
  float* Vt;
  int64_t N, c_out, J, batch0;
  __device__ __forceinline__ int64_t row_offset(int n, int64_t m) const {
    const int64_t v = m / c_out, o = m - v * c_out;
    return ((v * N + (batch0 + n)) * c_out + o) * J;
  }
  __device__ __forceinline__ void store(int64_t off, int64_t, int64_t j, float val, int) const { Vt[off + j] = val; }
  __device__ __forceinline__ float fetch(int64_t, int64_t, int64_t) const { return 0.f; }
  __device__ __forceinline__ void commit(int64_t off, int64_t, int64_t j, float val, float, int) const { Vt[off + j] = val; }
  __device__ __forceinline__ void operator()(int n, int64_t m, int64_t j, float val, int) const {
    Vt[row_offset(n, m) + j] = val;
  }
};

// The same store for the 4-d operand form (no re-layout of the factor): the GEMM's row index is padded,
// m' = tile * 128 + r with tile = v_chunk * o_chunks + o_chunk and r = o_local + 32 v_local (the order in which a
// TMA box {32 x, 32 o, 1 n, 4 v} of S[v][n][o][x] lands in shared memory); rows outside (V, Co) are skipped.
struct EmitStoreTc4 {
  float* Vt;
  int64_t N, c_out, J, batch0, V;
  int o_chunks;
  __device__ __forceinline__ int64_t row_offset(int n, int64_t m) const {
    const int64_t tile = m >> 7;
    const int r = int(m & 127);
    const int64_t v = (tile / o_chunks) * 4 + (r >> 5), o = (tile % o_chunks) * 32 + (r & 31);
    return (v < V && o < c_out) ? ((v * N + (batch0 + n)) * c_out + o) * J : int64_t(-1);
  }
  __device__ __forceinline__ void store(int64_t off, int64_t, int64_t j, float val, int) const {
    if (off >= 0) Vt[off + j] = val;
  }
  __device__ __forceinline__ float fetch(int64_t, int64_t, int64_t) const { return 0.f; }
  __device__ __forceinline__ void commit(int64_t off, int64_t, int64_t j, float val, float, int) const {
    if (off >= 0) Vt[off + j] = val;
  }
  __device__ __forceinline__ void operator()(int n, int64_t m, int64_t j, float val, int) const {
    const int64_t off = row_offset(n, m);
    if (off >= 0) Vt[off + j] = val;
  }
  // persistent small-K kernel: 32 (or fewer) consecutive columns of one row, from registers
  __device__ __forceinline__ void store_chunk(int64_t off, int64_t, int64_t col0, int ncols, const uint32_t (&r)[32]) const {
    float* dst = Vt + off + col0;
    if (ncols == 32 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q)  // 256-bit stores: whole 32-byte sectors
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * q), "r"(r[8 * q]),
                     "r"(r[8 * q + 1]), "r"(r[8 * q + 2]), "r"(r[8 * q + 3]), "r"(r[8 * q + 4]), "r"(r[8 * q + 5]),
                     "r"(r[8 * q + 6]), "r"(r[8 * q + 7])
                     : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) dst[i] = __uint_as_float(r[i]);
    }
  }
};

// A[b] = rows (v, o) of sample b of S[v][n][o][x], read in place through a 4-d tensor map; B[b] = Un[b] [J][Xp]
static int launch_emit_tc_4d(const float* S, const float* Un, float* Vt, int64_t V, int64_t N, int64_t c_out, int64_t J,
                             int64_t X, int64_t Xp, cudaStream_t stream, const char* what) {
  using namespace tc;
  auto kern = gram_tc_kernel<EmitStoreTc4, true, false, true>;
  static SmemOptIn opt_in;
  VVT_TRY(opt_in.ensure(kern, SMEM_BYTES, what));
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled unavailable", what);
  const int o_chunks = int(ceil_div(c_out, 32)), v_chunks = int(ceil_div(V, 4));
  const int64_t Mp = int64_t(o_chunks) * v_chunks * BM;
  const int tiles_n = int(ceil_div(J, BN));
  const int64_t kblocks = vmax<int64_t>(1, ceil_div(X, BK));
  for (int64_t b0 = 0; b0 < N; b0 += 65535) {
    const int64_t nb = vmin<int64_t>(65535, N - b0);
    CUtensorMap mapA, mapB;
    const cuuint64_t dims[4] = {cuuint64_t(X), cuuint64_t(c_out), cuuint64_t(nb), cuuint64_t(V)};
    const cuuint64_t strides[3] = {cuuint64_t(X) * 4, cuuint64_t(c_out) * X * 4, cuuint64_t(N) * c_out * X * 4};
    const cuuint32_t box[4] = {BK, 32, 1, 4};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (fn(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(S + b0 * c_out * X), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        !make_map(&mapB, Un + b0 * J * Xp, J, X, Xp, nb, J * Xp))
      return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", what);
    EmitStoreTc4 st{Vt, N, c_out, J, b0, V, o_chunks};
    if (smallk::worthwhile(Mp, J, X, nb)) {  // a few k-blocks per tile: the persistent form (gemm_smallk.cuh)
      VVT_TRY((smallk::launch<EmitStoreTc4, true>(mapA, mapB, st, Mp, J, X, nb, o_chunks * v_chunks, o_chunks, stream, what)));
      continue;
    }
    dim3 grid(unsigned(o_chunks * v_chunks * tiles_n), 1, unsigned(nb));
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(mapA, mapB, st, Mp, J, tiles_n, 0, int(kblocks), int(kblocks), o_chunks);
    VVT_TRY(launched(what));
  }
  return VVT_OK;
}

// St[(r, x)][o] (row pitch Cop) <- S[r][o][x]: one 32x32 tile per block, transposed through shared memory
__global__ void __launch_bounds__(256) dgrad_transpose_kernel(float* St, const float* S, int Co, int X, int Cop) {
  __shared__ float tile[32][33];
  const int64_t r = blockIdx.x;
  const int x0 = blockIdx.y * 32, o0 = blockIdx.z * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int o = o0 + k, x = x0 + tx;
    tile[k][tx] = (o < Co && x < X) ? ldg(S + (r * Co + o) * X + x) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int x = x0 + k, o = o0 + tx;
    if (x < X && o < Cop) St[(r * X + x) * Cop + o] = tile[tx][k];  // o in [Co, Cop) is zero
  }
}

// Wt[j][o] (row pitch Cop) <- W[o][j]
__global__ void weight_transpose_kernel(float* Wt, const float* W, int Co, int J, int Cop) {
  const int64_t total = int64_t(J) * Cop;
  VVT_GRID_STRIDE(i, total) {
    const int o = int(i % Cop), j = int(i / Cop);
    Wt[i] = o < Co ? ldg(W + int64_t(o) * J + j) : 0.f;
  }
}

// T2[r][j][x] <- C[(r, x), j]
struct DgradStoreTc {
  float* T2;
  int64_t X, J, batch0;
  __device__ __forceinline__ int64_t row_offset(int, int64_t m) const {
    const int64_t r = m / X, x = m - r * X;
    return r * J * X + x;
  }
  __device__ __forceinline__ void store(int64_t off, int64_t, int64_t j, float val, int) const { T2[off + j * X] = val; }
  __device__ __forceinline__ float fetch(int64_t, int64_t, int64_t) const { return 0.f; }
  __device__ __forceinline__ void commit(int64_t off, int64_t, int64_t j, float val, float, int) const { T2[off + j * X] = val; }
  __device__ __forceinline__ void operator()(int b, int64_t m, int64_t j, float val, int) const {
    T2[row_offset(b, m) + j * X] = val;
  }
  // persistent small-K kernel: lanes are consecutive rows m = (r, x), i.e. consecutive x: coalesced per column
  __device__ __forceinline__ void store_chunk(int64_t off, int64_t, int64_t col0, int ncols, const uint32_t (&r)[32]) const {
    float* dst = T2 + off + col0 * X;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < ncols) dst[int64_t(i) * X] = __uint_as_float(r[i]);
  }
};

// out[r][ci][iy][ix] = sum over (ky, kx) of T2[r][(ci, ky, kx)][(oy, ox)] with oy*sh + ky*dh - ph = iy, ...
// Flat over (plane = (r, ci), position): a block owns 256 consecutive outputs, splits its first index once and the
// threads go on in 32-bit arithmetic.  UNIT: stride 1 in both directions (no divisibility test, no division per
// tap); K3: a 3 x 3 filter, unrolled so that the nine loads of an output are in flight together.
template <bool UNIT, bool K3>
__global__ void __launch_bounds__(256) col2im_kernel(float* out, const float* T2, ConvGeom g, int64_t rows, int J, int X) {
  const int hw = g.h_in * g.w_in;
  const int64_t total = rows * g.c_in * hw;
  const int kh = K3 ? 3 : g.kh, kw = K3 ? 3 : g.kw;
  for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x; i0 < total; i0 += int64_t(gridDim.x) * blockDim.x) {
    if (i0 + threadIdx.x >= total) continue;
    const int64_t plane0 = i0 / hw;
    const int off = int(i0 - plane0 * hw) + int(threadIdx.x);  // < hw + 256
    const int dplane = off / hw, p = off - dplane * hw;
    int64_t r = plane0 / g.c_in;
    int ci = int(plane0 - r * g.c_in) + dplane;
    while (ci >= g.c_in) ci -= g.c_in, ++r;
    const float* base = T2 + (r * J + int64_t(ci) * kh * kw) * X;
    const int iy = p / g.w_in, ix = p - iy * g.w_in;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < kh; ++ky) {
      int oy = iy + g.ph - ky * g.dh;
      bool ok_y = true;
      if (!UNIT) {
        ok_y = oy >= 0 && oy % g.sh == 0;
        oy /= g.sh;
      }
      ok_y = ok_y && unsigned(oy) < unsigned(g.h_out);
#pragma unroll
      for (int kx = 0; kx < kw; ++kx) {
        int ox = ix + g.pw - kx * g.dw;
        bool ok = ok_y;
        if (!UNIT) {
          ok = ok && ox >= 0 && ox % g.sw == 0;
          ox /= g.sw;
        }
        ok = ok && unsigned(ox) < unsigned(g.w_out);
        if (ok) acc += ldg(base + int64_t(ky * kw + kx) * X + oy * g.w_out + ox);
      }
    }
    out[i0 + threadIdx.x] = acc;
  }
}

struct ConvWorkspace {
  int64_t a, b, c, total;  // byte offsets of up to three regions
};
static inline int64_t pad4(int64_t v) { return align_up(v, 4); }
// emit: Sn [N][V*Co][Xp] | Un [N][J][Xp]
static ConvWorkspace emit_workspace(int64_t V, int64_t N, int64_t Co, int64_t J, int64_t X) {
  ConvWorkspace w;
  const int64_t Xp = pad4(X);
  w.a = 0;
  w.b = align_up(N * V * Co * Xp * 4, 256);
  w.c = w.b + align_up(N * J * Xp * 4, 256);
  w.total = w.c;
  return w;
}
// dgrad: St [rows*X][Cop] | Wt [J][Cop] | T2 [rows][J][X]
static ConvWorkspace dgrad_workspace(int64_t rows, int64_t Co, int64_t Jd, int64_t X) {
  ConvWorkspace w;
  const int64_t Cop = pad4(Co);
  w.a = 0;
  w.b = align_up(rows * X * Cop * 4, 256);
  w.c = w.b + align_up(Jd * Cop * 4, 256);
  w.total = w.c + align_up(rows * Jd * X * 4, 256);
  return w;
}

static bool tc_conv_enabled() {
  static const bool off = getenv("VVT_CONV_IMPLICIT") != nullptr || getenv("VVT_NO_TCGEN05") != nullptr;
  return !off && tc::encode_fn() != nullptr;
}

}  // namespace vvt

using namespace vvt;

static int make_geom(ConvGeom& g, int64_t c_out, int64_t h_out, int64_t w_out, int64_t c_in,
                     int64_t h_in, int64_t w_in, int64_t kh, int64_t kw, int64_t sh, int64_t sw,
                     int64_t ph, int64_t pw, int64_t dh, int64_t dw) {
  if (c_out < 0 || h_out < 0 || w_out < 0 || c_in < 0 || h_in < 0 || w_in < 0 || kh <= 0 || kw <= 0 ||
      sh <= 0 || sw <= 0 || ph < 0 || pw < 0 || dh <= 0 || dw <= 0)
    return fail(VVT_ERR_INVALID, "%s", "conv2d: bad geometry");
  const int64_t lim = int64_t(1) << 30;
  if (c_out * h_out * w_out > lim || c_in * h_in * w_in > lim || c_in * kh * kw > lim || c_out * kh * kw > lim)
    return fail(VVT_ERR_UNSUPPORTED, "%s", "conv2d: per-sample extent exceeds 2^30");
  g = ConvGeom{int(c_out), int(h_out), int(w_out), int(c_in), int(h_in), int(w_in), int(kh),
               int(kw),    int(sh),    int(sw),    int(ph),   int(pw),   int(dh),   int(dw)};
  return VVT_OK;
}

extern "C" {

int64_t vvt_conv2d_workspace_bytes(int op, int64_t V, int64_t N, int64_t c_out, int64_t h_out, int64_t w_out,
                                   int64_t c_in, int64_t kh, int64_t kw, int dtype) {
  if (dtype != VVT_F32 || V <= 0 || N <= 0 || c_out <= 0 || h_out * w_out <= 0 || c_in * kh * kw <= 0) return 0;
  const int64_t X = h_out * w_out, J = c_in * kh * kw;
  return op == 0 ? emit_workspace(V, N, c_out, J, X).total : dgrad_workspace(V * N, c_out, J, X).total;
}

int vvt_v_emit_conv2d(void* Vt, const void* S, const void* X, int64_t V, int64_t N, int64_t c_out,
                      int64_t h_out, int64_t w_out, int64_t c_in, int64_t h_in, int64_t w_in,
                      int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w, int64_t pad_h,
                      int64_t pad_w, int64_t dil_h, int64_t dil_w, void* workspace, int64_t workspace_bytes,
                      int dtype, void* stream) {
  VVT_REQUIRE(V >= 0 && N >= 0, "negative size");
  ConvGeom g;
  VVT_TRY(make_geom(g, c_out, h_out, w_out, c_in, h_in, w_in, kh, kw, stride_h, stride_w, pad_h, pad_w,
                    dil_h, dil_w));
  const int64_t M = V * c_out, J = c_in * kh * kw, Xn = h_out * w_out;
  if (M == 0 || J == 0 || N == 0) return VVT_OK;
  VVT_REQUIRE(Vt && S && X, "null pointer");
  if (dtype == VVT_F32 && Xn > 0 && workspace && tc_conv_enabled() &&
      workspace_bytes >= emit_workspace(V, N, c_out, J, Xn).total && M < (int64_t(1) << 31)) {
    // unfold + re-layout + batched tcgen05 GEMM (one batch entry per sample)
    const ConvWorkspace w = emit_workspace(V, N, c_out, J, Xn);
    const int64_t Xp = pad4(Xn);
    float* Sn = (float*)((char*)workspace + w.a);
    float* Un = (float*)((char*)workspace + w.b);
    cudaStream_t s = as_stream(stream);
    im2col_kernel<<<int(vmin<int64_t>(N * J, 32 * int64_t(num_sms()))), int(vmin<int64_t>(256, align_up(Xp, 32))), 0, s>>>(
        Un, (const float*)X, g, int(N), int(J), int(Xn), int(Xp));
    VVT_TRY(launched("vvt_v_emit_conv2d(im2col)"));
    // The factor is read in place when TMA can address it: rows (v, o) of one sample through a 4-d tensor map
    // (V >= 4: a tile is 32 channels x 4 rows v), or as it is when there is a single row v per sample.  Only
    // spatial extents that are not a multiple of 4 floats (16-byte TMA strides) go through the re-layout copy.
    static const bool relayout = getenv("VVT_EMIT_PERMUTE") != nullptr;  // experiments: always copy
    const bool aligned = Xp == Xn && (reinterpret_cast<uintptr_t>(S) & 15) == 0;
    if (aligned && V >= 4 && !relayout)
      return launch_emit_tc_4d((const float*)S, Un, (float*)Vt, V, N, c_out, J, Xn, Xp, s, "vvt_v_emit_conv2d");
    EmitStoreTc st{(float*)Vt, N, c_out, J, 0};
    if (aligned && V == 1 && !relayout)  // S [1][N][Co][X] is already [n][(v, o)][x]
      return tc::launch_gemm_tc_batched<EmitStoreTc, true>((const float*)S, Un, st, M, J, Xn, Xp, Xp, N, M * Xp, J * Xp, s,
                                                           "vvt_v_emit_conv2d");
    emit_permute_kernel<<<ew_blocks(N * M * Xp), 256, 0, s>>>(Sn, (const float*)S, int(V), int(N), int(c_out),
                                                             int(Xn), int(Xp));
    VVT_TRY(launched("vvt_v_emit_conv2d(permute)"));
    return tc::launch_gemm_tc_batched<EmitStoreTc, true>(Sn, Un, st, M, J, Xn, Xp, Xp, N, M * Xp, J * Xp, s,
                                                         "vvt_v_emit_conv2d");
  }
  VVT_DISPATCH(dtype, {
    if (Xn == 0) return check_cuda(cudaMemsetAsync(Vt, 0, V * N * c_out * J * sizeof(T), as_stream(stream)), __func__);
    EmitLoaderS<T> la{(const T*)S, N, c_out, Xn, M};
    EmitLoaderPatch<T> lb{(const T*)X, g, J, Xn};
    EmitStore<T> st{(T*)Vt, N, c_out, J};
    return launch_gemm_custom<T>(la, lb, st, M, J, Xn, N, as_stream(stream), "vvt_v_emit_conv2d");
  });
}

int vvt_sqrt_backprop_conv2d(void* out, const void* S, const void* W, int64_t rows, int64_t c_out,
                             int64_t h_out, int64_t w_out, int64_t c_in, int64_t h_in, int64_t w_in,
                             int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w,
                             int64_t pad_h, int64_t pad_w, int64_t dil_h, int64_t dil_w, void* workspace,
                             int64_t workspace_bytes, int dtype, void* stream) {
  VVT_REQUIRE(rows >= 0, "negative size");
  ConvGeom g;
  VVT_TRY(make_geom(g, c_out, h_out, w_out, c_in, h_in, w_in, kh, kw, stride_h, stride_w, pad_h, pad_w,
                    dil_h, dil_w));
  const int64_t hw = h_in * w_in, M = rows * hw, Kd = c_out * kh * kw;
  if (M == 0 || c_in == 0) return VVT_OK;
  VVT_REQUIRE(out && S && W, "null pointer");
  const int64_t Xo = h_out * w_out, Jd = c_in * kh * kw;
  if (dtype == VVT_F32 && Xo > 0 && c_out > 0 && workspace && tc_conv_enabled() &&
      workspace_bytes >= dgrad_workspace(rows, c_out, Jd, Xo).total && rows * Xo < (int64_t(1) << 31)) {
    // transpose (contraction index c_out contiguous) + one tcgen05 GEMM + col2im gather
    const ConvWorkspace w = dgrad_workspace(rows, c_out, Jd, Xo);
    const int64_t Cop = pad4(c_out);
    float* St = (float*)((char*)workspace + w.a);
    float* Wt = (float*)((char*)workspace + w.b);
    float* T2 = (float*)((char*)workspace + w.c);
    cudaStream_t s = as_stream(stream);
    const int64_t ybl = ceil_div(Xo, 32), zbl = ceil_div(Cop, 32);
    if (ybl > 65535 || zbl > 65535) return fail(VVT_ERR_UNSUPPORTED, "%s", "conv2d: map too large");
    dgrad_transpose_kernel<<<dim3(unsigned(rows), unsigned(ybl), unsigned(zbl)), 256, 0, s>>>(
        St, (const float*)S, int(c_out), int(Xo), int(Cop));
    VVT_TRY(launched("vvt_sqrt_backprop_conv2d(transpose)"));
    weight_transpose_kernel<<<ew_blocks(Jd * Cop), 256, 0, s>>>(Wt, (const float*)W, int(c_out), int(Jd), int(Cop));
    VVT_TRY(launched("vvt_sqrt_backprop_conv2d(weights)"));
    DgradStoreTc st{T2, Xo, Jd, 0};
    if (smallk::worthwhile(rows * Xo, Jd, c_out, 1)) {  // K = c_out: a few k-blocks per tile (gemm_smallk.cuh)
      CUtensorMap mapA, mapB;
      if (!tc::make_map(&mapA, St, rows * Xo, c_out, Cop, 1, 0) || !tc::make_map(&mapB, Wt, Jd, c_out, Cop, 1, 0))
        return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", "vvt_sqrt_backprop_conv2d");
      VVT_TRY((smallk::launch<DgradStoreTc, false>(mapA, mapB, st, rows * Xo, Jd, c_out, 1, int(ceil_div(rows * Xo, tc::BM)), 1,
                                                   s, "vvt_sqrt_backprop_conv2d")));
    } else {
      VVT_TRY((tc::launch_gemm_tc_batched<DgradStoreTc, false>(St, Wt, st, rows * Xo, Jd, c_out, Cop, Cop, 1, 0, 0, s,
                                                               "vvt_sqrt_backprop_conv2d")));
    }
    {
      const int blocks = ew_blocks(M * c_in);
      const bool unit = stride_h == 1 && stride_w == 1, k3 = kh == 3 && kw == 3;
      auto kern = unit ? (k3 ? col2im_kernel<true, true> : col2im_kernel<true, false>)
                       : (k3 ? col2im_kernel<false, true> : col2im_kernel<false, false>);
      kern<<<blocks, 256, 0, s>>>((float*)out, T2, g, rows, int(Jd), int(Xo));
    }
    return launched("vvt_sqrt_backprop_conv2d(col2im)");
  }
  VVT_DISPATCH(dtype, {
    DgradLoaderS<T> la{(const T*)S, g, M, Kd};
    DgradLoaderW<T> lb{(const T*)W, g, Kd};
    DgradStore<T> st{(T*)out, c_in, hw};
    return launch_gemm_custom<T>(la, lb, st, M, c_in, Kd, 1, as_stream(stream), "vvt_sqrt_backprop_conv2d");
  });
}

}  // extern "C"
