// Conv2d pieces of the sqrt-GGN factor as implicit GEMMs on the shared main loop:
//   * V emit: per-sample  Vt[(v,n), (o, j)] = sum_x S[v,n,o,x] * patch[n, j, x]
//   * data gradient of the factor ("virtual batch" of V*N maps)
#include "gemm_core.cuh"

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

int vvt_v_emit_conv2d(void* Vt, const void* S, const void* X, int64_t V, int64_t N, int64_t c_out,
                      int64_t h_out, int64_t w_out, int64_t c_in, int64_t h_in, int64_t w_in,
                      int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w, int64_t pad_h,
                      int64_t pad_w, int64_t dil_h, int64_t dil_w, int dtype, void* stream) {
  VVT_REQUIRE(V >= 0 && N >= 0, "negative size");
  ConvGeom g;
  VVT_TRY(make_geom(g, c_out, h_out, w_out, c_in, h_in, w_in, kh, kw, stride_h, stride_w, pad_h, pad_w,
                    dil_h, dil_w));
  const int64_t M = V * c_out, J = c_in * kh * kw, Xn = h_out * w_out;
  if (M == 0 || J == 0 || N == 0) return VVT_OK;
  VVT_REQUIRE(Vt && S && X, "null pointer");
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
                             int64_t pad_h, int64_t pad_w, int64_t dil_h, int64_t dil_w, int dtype,
                             void* stream) {
  VVT_REQUIRE(rows >= 0, "negative size");
  ConvGeom g;
  VVT_TRY(make_geom(g, c_out, h_out, w_out, c_in, h_in, w_in, kh, kw, stride_h, stride_w, pad_h, pad_w,
                    dil_h, dil_w));
  const int64_t hw = h_in * w_in, M = rows * hw, Kd = c_out * kh * kw;
  if (M == 0 || c_in == 0) return VVT_OK;
  VVT_REQUIRE(out && S && W, "null pointer");
  VVT_DISPATCH(dtype, {
    DgradLoaderS<T> la{(const T*)S, g, M, Kd};
    DgradLoaderW<T> lb{(const T*)W, g, Kd};
    DgradStore<T> st{(T*)out, c_in, hw};
    return launch_gemm_custom<T>(la, lb, st, M, c_in, Kd, 1, as_stream(stream), "vvt_sqrt_backprop_conv2d");
  });
}

}  // extern "C"
