// (header) fp32 Gram / NT-GEMM on the 5th-generation tensor cores (tcgen05), fed by TMA.
//
//   acc[m, n] = sum_k A[m, k] * B[n, k]          A: [M, K], B: [N, K], both K-contiguous (fp32)
//
// fp32-grade accuracy comes from the 3xTF32 split  a*b ~= ah*bh + ah*bl + al*bh, where ah is the
// tf32 the tensor core sees when it reads an fp32 word (the 13 low mantissa bits are ignored) and
// al = rna_tf32(a - ah) is produced in shared memory by four converter warps.  Nothing is split in
// global memory: one TMA load per operand tile and k-block, the raw tile doubles as the "hi" operand.
//
// CTA = one 128x128 output tile x one K range (split-K), 6 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor.2d, 128B swizzle, 128 rows x 32 floats per tile)
//   warp 1      MMA issuer     (tcgen05.mma.cta_group::1.kind::tf32, M = N = 128, K = 8; accumulators
//                               in 128 TMEM columns); hi*hi is issued as soon as the tile lands,
//                               hi*lo + lo*hi after the converters are done
//   warps 2..5  converters     (raw -> lo tiles), then the epilogue (tcgen05.ld -> registers -> store)
// Pipeline: 3 stages of {A raw, B raw, A lo, B lo} = 64 KB each; mbarriers tma_full / conv_full /
// empty per stage.  Diagonal tiles of a symmetric product load A only.
//
// Two-level accumulation: the tensor core adds into its fp32 accumulator with truncation, a bias of
// about -2^-24 of the running sum per MMA (measured: -1e-4 on Gram diagonals after 2300 steps).  The
// MMA warp therefore accumulates only PROMOTE k-blocks (48 MMAs) into one of two TMEM buffers; the
// converter warps drain the finished buffer into fp32 registers (round-to-nearest adds) while the
// next group runs in the other buffer.  Residual bias on sums of same-sign products: about -3e-6.
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "gemm_core.cuh"

namespace vvt {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;  // BK floats = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;     // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES; // A raw | B raw | A lo | B lo
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;  // two accumulator buffers of 128 columns
constexpr int PROMOTE = 4;      // k-blocks accumulated in the tensor core before promotion to registers
constexpr size_t SMEM_BYTES = size_t(STAGES) * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait (~2 s): a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (spins & 63) == 63) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row of the tile, one 32-bit column per tf32 element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile [rows][32 floats], 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= uint64_t(1) << 16;                     // leading byte offset (ignored for swizzled K-major)
  d |= uint64_t(1024 >> 4) << 32;             // stride byte offset
  d |= uint64_t(1) << 46;                     // descriptor version (Blackwell)
  d |= uint64_t(2) << 61;                     // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(BM >> 4) << 24);

// COL_LANES = false: every epilogue thread stores its own output row (lanes = consecutive rows);
// COL_LANES = true : the tile is staged through shared memory and stored with lanes = consecutive
//                    columns (coalesced for row-major outputs).  blockIdx.z = batch index.
// TS = true: the A operand (hi and lo) lives in TENSOR MEMORY instead of shared memory.  With both operands in
// shared memory a 128 x 128 x 8 tf32 MMA reads 8 KB per 64 cycles -- the SM's whole shared-memory bandwidth --
// while the converter warps and the TMA writes need the same port: measured 44-46 % tensor-pipe activity.  The
// converter threads already hold a row of the A tile in registers when they form lo; they store hi and lo to
// TMEM (tcgen05.st, lane = row, one column per element) and the MMAs fetch only B from shared memory.
// A4D = true: the rows of an A tile are a 32 x 4 box of TWO row indices (a 4-d tensor map {k, i, batch, j}, box
// {32, 32, 1, 4}: tile row r = i_local + 32 j_local); row tile tm covers i-chunk tm % a4_chunks and j-chunk
// tm / a4_chunks.  Used by the conv factor emit, whose rows (v, o) of a sample are not equidistant in memory: the
// store functor decodes the padded row index (see EmitStoreTc4 in conv.cu).
// TS kernels run EIGHT converter / epilogue warps (THREADS_TS): two per quarter of the tensor-memory lanes, the
// second set takes columns 16..31 of every k-block of A, the second half of the lo tile of B and columns 64..127 of the
// accumulators, so that every scheduler has two warps to hide the shared-memory, ALU and tcgen05.wait latencies of
// the conversion behind, and the epilogue has twice the threads.  Measured: nothing at R = 1280, D = 110592
// (178.5 TFLOP/s before and after: the main loop is not conversion-bound), 4.5 % over the twelve Gram calls of a c2
// step (3.51 -> 3.35 ms: the short ones are prologue / epilogue).
constexpr int THREADS_TS = 320;
template <typename ST, bool COL_LANES, bool TS = false, bool A4D = false>
__global__ void __launch_bounds__(TS ? THREADS_TS : THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, ST st,
               int64_t M, int64_t N, int tiles_n, int symmetric, int kblocks_total, int kblocks_per_split,
               int a4_chunks = 1) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // tiles must be 1024-byte aligned
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  // barrier slots (8 bytes each): tma_full[S] | conv_full[S] | empty[S] | acc_full[2] | acc_empty[2]; then the
  // TMEM address
  auto bar_tma = [&](int s) { return bars + 8u * s; };
  auto bar_conv = [&](int s) { return bars + 8u * (STAGES + s); };
  auto bar_empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto bar_acc_full = [&](int b) { return bars + 8u * (3 * STAGES + b); };
  auto bar_acc_empty = [&](int b) { return bars + 8u * (3 * STAGES + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 4));

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
  const bool diag = symmetric && tm == tn;
  const int split = blockIdx.y, batch = blockIdx.z;
  const int kb0 = split * kblocks_per_split, kb1 = min(kblocks_total, kb0 + kblocks_per_split);
  const int nkb = kb1 - kb0;  // >= 1 (host guarantees)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NCONV = (TS ? THREADS_TS : THREADS) - 64;  // converter / epilogue threads
  constexpr int NW = NCONV / 32;                            // ... warps (4 or 8)
  constexpr int BNT = BN * 4 / NW;                          // accumulator columns per thread (128 or 64)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_tma(s), 1);
      mbar_init(bar_conv(s), NCONV);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full(b), 1);
      mbar_init(bar_acc_empty(b), NCONV);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int kTmemCols = TS ? 512 : TMEM_COLS;  // TS: + 3 stages x (32 hi + 32 lo) columns of A behind the accumulators
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto a_cols = [&](int s) { return tmem_base + uint32_t(TMEM_COLS + 64 * s); };  // hi at +0, lo at +32

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, use = i / STAGES;
        if (use > 0) mbar_wait(bar_empty(s), (use - 1) & 1);
        const uint32_t stage = base + s * STAGE_BYTES;
        mbar_expect_tx(bar_tma(s), diag ? TILE_BYTES : 2 * TILE_BYTES);
        const int k0 = (kb0 + i) * BK;
        if constexpr (A4D) tma_load_4d(stage, &mapA, bar_tma(s), k0, (tm % a4_chunks) * 32, batch, (tm / a4_chunks) * 4);
        else tma_load_3d(stage, &mapA, bar_tma(s), k0, tm * BM, batch);
        if (!diag) tma_load_3d(stage + TILE_BYTES, &mapB, bar_tma(s), k0, tn * BN, batch);
      }
    }
  } else if (warp == 1 && TS) {
    // ===== MMA issuer, A from tensor memory: every product waits for the converters =====
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE, first = (i % PROMOTE) == 0;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t stage = base + s * STAGE_BYTES;
        const uint32_t b_hi = diag ? stage : stage + TILE_BYTES;
        const uint32_t b_lo = diag ? stage + 2 * TILE_BYTES : stage + 3 * TILE_BYTES;
        if (first && grp >= 2) {
          mbar_wait(bar_acc_empty(grp & 1), ((grp >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        mbar_wait(bar_conv(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint32_t a_hi = a_cols(s) + 8 * k, a_lo = a_hi + 32;
          umma_tf32_ts(acc, a_hi, make_desc(b_hi + 32 * k), kIdesc, !(first && k == 0));
          umma_tf32_ts(acc, a_hi, make_desc(b_lo + 32 * k), kIdesc, 1);
          umma_tf32_ts(acc, a_lo, make_desc(b_hi + 32 * k), kIdesc, 1);
        }
        umma_commit(bar_empty(s));  // frees the stage (shared memory and the A columns in TMEM)
        if ((i % PROMOTE) == PROMOTE - 1 || i == nkb - 1) umma_commit(bar_acc_full(grp & 1));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // Software-pipelined by one k-block: hi*hi of block i + 1 is issued BEFORE the lo terms of block i, so the
    // tensor pipe has work while the converter warps are still writing the lo tile of block i (the raw tile
    // doubles as the hi operand and is usable the moment the TMA lands).
    if (lane == 0) {
      auto issue_hi = [&](int i) {
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE, first = (i % PROMOTE) == 0;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t stage = base + s * STAGE_BYTES;
        const uint32_t a_hi = stage, b_hi = diag ? stage : stage + TILE_BYTES;
        if (first && grp >= 2) {  // the buffer was drained two groups ago?
          mbar_wait(bar_acc_empty(grp & 1), ((grp >> 1) - 1) & 1);
          tcgen05_fence_after();
        }
        mbar_wait(bar_tma(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)  // 8 tf32 = 32 bytes per MMA along K
          umma_tf32(acc, make_desc(a_hi + 32 * k), make_desc(b_hi + 32 * k), kIdesc, !(first && k == 0));
      };
      issue_hi(0);
      for (int i = 0; i < nkb; ++i) {
        if (i + 1 < nkb) issue_hi(i + 1);
        const int s = i % STAGES, use = i / STAGES;
        const int grp = i / PROMOTE;
        const uint32_t acc = tmem_base + uint32_t((grp & 1) * BN);
        const uint32_t stage = base + s * STAGE_BYTES;
        const uint32_t a_hi = stage, b_hi = diag ? stage : stage + TILE_BYTES;
        const uint32_t a_lo = stage + 2 * TILE_BYTES, b_lo = diag ? a_lo : stage + 3 * TILE_BYTES;
        mbar_wait(bar_conv(s), use & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          umma_tf32(acc, make_desc(a_hi + 32 * k), make_desc(b_lo + 32 * k), kIdesc, 1);
          umma_tf32(acc, make_desc(a_lo + 32 * k), make_desc(b_hi + 32 * k), kIdesc, 1);
        }
        umma_commit(bar_empty(s));  // frees the stage once these MMAs have read it
        if ((i % PROMOTE) == PROMOTE - 1 || i == nkb - 1) umma_commit(bar_acc_full(grp & 1));
      }
    }
  } else {
    // ===== converters: lo = rna_tf32(x - trunc_tf32(x)); promotion of finished accumulator groups =====
    const int ct = threadIdx.x - 64;  // 0..NCONV-1
    const int n_vec = (diag ? 1 : 2) * (TILE_BYTES / 16);
    const int lane_grp = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
    const int half = (warp - 2) >> 2;  // 0, or 1 for the second set of converter warps (TS)
    const int cbase = half * BNT;      // first accumulator column of this thread
    const int n_groups = (nkb + PROMOTE - 1) / PROMOTE;
    float total[BNT];  // this thread's part of its row of the output tile, summed with round-to-nearest adds
#pragma unroll
    for (int j = 0; j < BNT; ++j) total[j] = 0.f;
    auto drain = [&](int grp) {
      mbar_wait(bar_acc_full(grp & 1), (grp >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BNT; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(lane_grp * 32) << 16) + uint32_t((grp & 1) * BN + cbase + c0), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) total[c0 + j] += __uint_as_float(r[j]);
      }
      tcgen05_fence_before();
      mbar_arrive(bar_acc_empty(grp & 1));
    };
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES, use = i / STAGES;
      unsigned char* stage = base_ptr + s * STAGE_BYTES;
      mbar_wait(bar_tma(s), use & 1);
      // lo buffers of this stage are free: the MMAs that read them committed to `empty` before the
      // producer refilled the stage, and that refill is what we just waited for
      if constexpr (TS) {
        // A role: this thread's row of the raw tile (128-byte swizzle: 16-byte chunk c of row m sits at chunk
        // c ^ (m & 7); the eight lanes of a quarter warp hit eight different chunks) -> hi and lo in TMEM
        // (the two threads of a row take 16 of its 32 columns each)
        const int m = lane_grp * 32 + lane;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cc = 4 * half + c;
          const float4 x = *reinterpret_cast<const float4*>(stage + size_t(m) * 128 + size_t((cc ^ (m & 7)) * 16));
          const float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t bits = __float_as_uint(e[j]);
            hi[4 * c + j] = bits;
            lo[4 * c + j] = __float_as_uint(e[j] - __uint_as_float(bits & 0xFFFFE000u)) + 0x1000u;
          }
        }
        const uint32_t ta = a_cols(s) + (uint32_t(lane_grp * 32) << 16) + uint32_t(16 * half);
        tmem_st16(ta, hi);
        tmem_st16(ta + 32, lo);
        tmem_st_wait();
      }
      // B role (TS: only the B tile, which is the raw A tile itself on diagonal tiles): lo tiles in shared memory
#pragma unroll 4
      for (int v = (TS && !diag) ? ct + TILE_BYTES / 16 : ct; v < n_vec; v += NCONV) {
        const float4 x = *reinterpret_cast<const float4*>(stage + size_t(v) * 16);
        float4 lo;
        {
          const float e[4] = {x.x, x.y, x.z, x.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // exact residual; + half a tf32 ulp on its bit pattern: the tensor core's truncation of the low
            // 13 bits then rounds to nearest (see split_tf32 in gemm_core.cuh) -- 3 ALU ops per element
            const float hi = __uint_as_float(__float_as_uint(e[j]) & 0xFFFFE000u);
            o[j] = __uint_as_float(__float_as_uint(e[j] - hi) + 0x1000u);
          }
          lo = make_float4(o[0], o[1], o[2], o[3]);
        }
        *reinterpret_cast<float4*>(stage + 2 * TILE_BYTES + size_t(v) * 16) = lo;
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      if constexpr (TS) tcgen05_fence_before();
      mbar_arrive(bar_conv(s));
      // one group behind the conversions, so that the wait is (almost) never a stall
      if ((i % PROMOTE) == PROMOTE - 1 && i / PROMOTE >= 1) drain(i / PROMOTE - 1);
    }
    // groups not drained yet: the last one, and the one before it if the last group was complete
    for (int g = (nkb % PROMOTE == 0) ? n_groups - 1 : vmax(0, n_groups - 2); g < n_groups; ++g) drain(g);
    // ===== epilogue: registers -> shared memory -> global =====
    // ST provides  row_offset(batch, row) -> int64 (may divide: called a few times per thread),
    //              store(row_offset, row, col, value, split) (cheap) and operator() for mirrored entries.
    // Every MMA of this CTA has completed (the last group was drained) and so has every TMA load: the
    // pipeline stages are free and hold the staged tile [BM][BN + 1].  The store loops are NOT unrolled
    // over the tile (a fully unrolled 128-store epilogue thrashes the instruction cache: 50% no_inst stalls).
    {
      float* tile = reinterpret_cast<float*>(base_ptr);
      const int trow = lane_grp * 32 + lane;
#pragma unroll
      for (int j = 0; j < BNT; ++j) tile[trow * (BN + 1) + cbase + j] = total[j];
      asm volatile("bar.sync 1, %0;" ::"n"(NCONV) : "memory");
      const int cw = warp - 2;  // 0..NW-1
      const int64_t row0 = int64_t(tm) * BM, col0 = int64_t(tn) * BN;
      if constexpr (COL_LANES) {
        // lanes = consecutive columns (row-major outputs): this warp stores rows cw, cw + NW, ...;
        // lane l prepares the offset of row cw + NW l, the others fetch it by shuffle.  Rows go in batches of
        // 4: the 16 old values an accumulating store needs are fetched before the first store.
        const int64_t my_row = row0 + cw + NW * lane;
        const int64_t my_off = (NW * lane < BM && my_row < M) ? st.row_offset(batch, my_row) : 0;
        for (int i0 = 0; i0 < BM / NW; i0 += 4) {
          if (row0 + cw + NW * i0 >= M) break;  // warp-uniform
          int64_t off[4];
          float old[4][BN / 32];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            off[u] = __shfl_sync(0xffffffffu, my_off, i0 + u);
            const int64_t row = row0 + cw + NW * (i0 + u);
#pragma unroll
            for (int q = 0; q < BN / 32; ++q) {
              const int64_t col = col0 + lane + 32 * q;
              old[u][q] = (row < M && col < N) ? st.fetch(off[u], row, col) : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = cw + NW * (i0 + u);
            const int64_t row = row0 + r;
#pragma unroll
            for (int q = 0; q < BN / 32; ++q) {
              const int c = lane + 32 * q;
              const int64_t col = col0 + c;
              if (row < M && col < N) st.commit(off[u], row, col, tile[r * (BN + 1) + c], old[u][q], split);
            }
          }
        }
      } else {
        // lanes = consecutive rows (outputs contiguous along the row index): this warp stores columns
        // cw, cw + 4, ...
        int64_t offs[BM / 32];
#pragma unroll
        for (int c = 0; c < BM / 32; ++c) {
          const int64_t row = row0 + lane + 32 * c;
          offs[c] = row < M ? st.row_offset(batch, row) : 0;
        }
        for (int j = cw; j < BN; j += NW) {
          const int64_t col = col0 + j;
          if (col >= N) break;  // warp-uniform
#pragma unroll
          for (int c = 0; c < BM / 32; ++c) {
            const int r = lane + 32 * c;
            if (row0 + r < M) st.store(offs[c], row0 + r, col, tile[r * (BN + 1) + j], split);
          }
        }
      }
      if (symmetric && !diag) {  // mirrored entries (col, row): lanes = consecutive rows -> coalesced
        for (int j0 = cw; j0 < BN; j0 += 4 * NW) {  // columns j0, j0 + NW, j0 + 2 NW, j0 + 3 NW per batch
          if (col0 + j0 >= N) break;             // warp-uniform
          float old[4][BM / 32];
          int64_t moff[4];  // the mirrored entry's row is this tile's column: one offset per column
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t col = col0 + j0 + NW * u;
            moff[u] = col < N ? st.row_offset(batch, col) : 0;
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) {
              const int64_t row = row0 + lane + 32 * c;
              old[u][c] = (col < N && row < M) ? st.fetch(moff[u], col, row) : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + NW * u;
            const int64_t col = col0 + j;
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) {
              const int r = lane + 32 * c;
              if (col < N && row0 + r < M) st.commit(moff[u], col, row0 + r, tile[r * (BN + 1) + j], old[u][c], split);
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-d map over `batch` row-major [rows, cols] fp32 matrices (leading dimension ld, `batch_stride`
// elements apart); box = 128 rows x 32 floats of one matrix; out-of-range rows / columns read as zero
static inline bool make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int64_t batch,
                            int64_t batch_stride) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if (batch <= 1) batch = 1, batch_stride = rows * ld;
  const cuuint64_t dims[3] = {cuuint64_t(cols), cuuint64_t(rows), cuuint64_t(batch)};
  const cuuint64_t strides[2] = {cuuint64_t(ld) * 4, cuuint64_t(batch_stride) * 4};
  const cuuint32_t box[3] = {BK, BM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// operands usable by TMA: 16-byte aligned base, row pitch and batch stride multiples of 16 bytes
static inline bool operands_ok(const float* A, const float* B, int64_t lda, int64_t ldb, int64_t sa, int64_t sb) {
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) return false;
  if ((lda | ldb | sa | sb) & 3) return false;
  return encode_fn() != nullptr;
}

// Batched C[b] = A[b] B[b]^T through an arbitrary store functor (no split-K):
//   A[b]: [M, K] (lda, batch stride sa), B[b]: [N, K] (ldb, sb); st(b, row, col, value, 0) for in-range entries.
template <typename ST, bool COL_LANES>
inline int launch_gemm_tc_batched(const float* A, const float* B, ST st, int64_t M, int64_t N, int64_t K, int64_t lda,
                                  int64_t ldb, int64_t batch, int64_t sa, int64_t sb, cudaStream_t stream,
                                  const char* what) {
  if (M <= 0 || N <= 0 || batch <= 0) return VVT_OK;
  auto kern = gram_tc_kernel<ST, COL_LANES>;
  static SmemOptIn opt_in;  // per instantiation
  VVT_TRY(opt_in.ensure(kern, SMEM_BYTES, what));
  const int tiles_m = int(ceil_div(M, BM)), tiles_n = int(ceil_div(N, BN));
  const int64_t kblocks = vmax<int64_t>(1, ceil_div(K, BK));
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {  // gridDim.z limit
    const int64_t nb = vmin<int64_t>(65535, batch - b0);
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A + b0 * sa, M, K, lda, nb, sa) || !make_map(&mapB, B + b0 * sb, N, K, ldb, nb, sb))
      return fail(VVT_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", what);
    ST stb = st;
    stb.batch0 = b0;
    dim3 grid(unsigned(tiles_m * tiles_n), 1, unsigned(nb));
    kern<<<grid, THREADS, SMEM_BYTES, stream>>>(mapA, mapB, stb, M, N, tiles_n, 0, int(kblocks), int(kblocks), 1);
    VVT_TRY(launched(what));
  }
  return VVT_OK;
}

}  // namespace tc
}  // namespace vvt
