/*
 * vivit_b200.h -- C ABI of libvivit_b200.so (hand-written sm_100a CUDA kernels
 * for ViViT's low-rank GGN hot path).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - tensors are dense, row-major, in the layout written next to them;
 *   - `dtype` is VVT_F32 or VVT_F64 and applies to every floating-point buffer
 *     of the call; index buffers are int64;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the caller owns every buffer, including `workspace`; the matching
 *     *_workspace_bytes() function says how large it must be;
 *   - every entry point returns 0 on success and a non-zero vvt_status
 *     otherwise and never throws; vvt_last_error() (host string, thread-local)
 *     describes the last failure.
 *
 * R = C * N_ggn is the Gram dimension ("rows" of V^T), D_p the number of
 * entries of one parameter, K the number of kept directions.
 *
 * Each entry point names the reference call site it replaces; paths are
 * relative to the f-dangel/vivit v1.0.0 source tree. "[BackPACK]" marks
 * behaviour of the un-vendored dependency backpack-for-pytorch>=1.5,<2
 * (setup.cfg:36) that the reference reaches through the cited line.
 */
#ifndef VIVIT_B200_H
#define VIVIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum { VVT_F32 = 0, VVT_F64 = 1 } vvt_dtype;

typedef enum {
  VVT_OK = 0,
  VVT_ERR_INVALID = 1,   /* bad argument (size, dtype, null pointer)     */
  VVT_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed  */
  VVT_ERR_WORKSPACE = 3, /* workspace too small                          */
  VVT_ERR_NOCONV = 4,    /* eigensolver hit the sweep limit              */
  VVT_ERR_UNSUPPORTED = 5
} vvt_status;

/* element-wise Jacobian kinds for vvt_sqrt_backprop_elementwise */
typedef enum {
  VVT_ACT_RELU = 0,    /* ref = layer input : S * (ref > 0)            */
  VVT_ACT_SIGMOID = 1, /* ref = layer output: S * ref * (1 - ref)      */
  VVT_ACT_TANH = 2,    /* ref = layer output: S * (1 - ref^2)          */
  VVT_ACT_DROPOUT = 3, /* ref = layer output: S * (ref != 0) * scale   */
  VVT_ACT_MUL = 4,     /* ref = derivative  : S * ref                  */
  VVT_ACT_LEAKY_RELU = 5, /* ref = layer input : S * (ref > 0 ? 1 : scale)          (scale = negative_slope) */
  VVT_ACT_ELU = 6,        /* ref = layer input : S * (ref > 0 ? 1 : scale * exp(ref))        (scale = alpha) */
  VVT_ACT_SELU = 7,       /* ref = layer input : S * lambda * (ref > 0 ? 1 : alpha * exp(ref))               */
  VVT_ACT_LOGSIGMOID = 8  /* ref = layer input : S / (1 + exp(ref))                                          */
} vvt_act;

int vvt_abi_version(void);
const char* vvt_last_error(void); /* host string */
/* number of kernel launches issued by this library since process start (host counter) */
int64_t vvt_launch_count(void);

/* ------------------------------------------------------------------------ *
 * (1) symmetric factor of the loss Hessian                                  *
 * ------------------------------------------------------------------------ */

/* S[v,n,c] = tau[n,c] (delta_vc - tau[n,v] tau[n,c]) * scale,  tau = sqrt(softmax(logits[sub[n]]))
 * S: [C, n_sub, C]; logits: [n_total, C]; sub: [n_sub] or NULL (= identity).
 * scale = 1/sqrt(n_total) for reduction='mean', 1 for 'sum'.
 * Replaces [BackPACK] CrossEntropyLossDerivatives.sqrt_hessian, reached via
 * vivit/extensions/secondorder/vivit/__init__.py:86. */
int vvt_loss_sqrt_hessian_ce(void* S, const void* logits, const int64_t* sub, int64_t n_total,
                             int64_t n_sub, int64_t C, double scale, int dtype, void* stream);

/* S[m,n,c] = (softmax(logits[sub[n]])[c] - 1[class_ids[m,n]==c]) * scale
 * S: [M, n_sub, C]; class_ids: [M, n_sub]. scale = 1/sqrt(M) (/sqrt(n_total) for 'mean').
 * Replaces [BackPACK] CrossEntropyLossDerivatives.sqrt_hessian_sampled (same line,
 * strategy switch at __init__.py:122-128,173). */
int vvt_loss_sqrt_hessian_ce_mc(void* S, const void* logits, const int64_t* sub,
                                const int64_t* class_ids, int64_t n_total, int64_t n_sub, int64_t C,
                                int64_t M, double scale, int dtype, void* stream);

/* S[v,n,c] = scale * delta_vc;  S: [C, n_sub, C].
 * Replaces [BackPACK] MSELossDerivatives.sqrt_hessian (__init__.py:85). */
int vvt_loss_sqrt_hessian_mse(void* S, int64_t n_sub, int64_t C, double scale, int dtype,
                              void* stream);

/* ------------------------------------------------------------------------ *
 * (1) back-propagation of the factor through layers ([BackPACK] jac_t_mat_prod,*
 *     called by MatToJacMat.backpropagate, base class at base.py:8,19)       *
 * ------------------------------------------------------------------------ */

/* out[r, i] = sum_o S[r, o] W[o, i];  S: [rows, n_out], W: [n_out, n_in], out: [rows, n_in] */
int vvt_sqrt_backprop_linear(void* out, const void* S, const void* W, int64_t rows, int64_t n_out,
                             int64_t n_in, int dtype, void* stream);

/* Scratch for the fp32 tensor-core path of the two Conv2d entry points below: op = 0 for
 * vvt_v_emit_conv2d (V, N as given there), op = 1 for vvt_sqrt_backprop_conv2d (V * N = rows).
 * 0 for fp64.  Without (enough) workspace both fall back to the generic implicit-GEMM kernel. */
int64_t vvt_conv2d_workspace_bytes(int op, int64_t V, int64_t N, int64_t c_out, int64_t h_out,
                                   int64_t w_out, int64_t c_in, int64_t kh, int64_t kw, int dtype);

/* data gradient of a 2d cross-correlation, applied to `rows` = V*N stacked maps.
 * S: [rows, c_out, h_out, w_out], W: [c_out, c_in, kh, kw], out: [rows, c_in, h_in, w_in].
 * groups = 1.  fp32 with workspace: transpose (c_out contiguous) -> tcgen05 GEMM -> col2im gather. */
int vvt_sqrt_backprop_conv2d(void* out, const void* S, const void* W, int64_t rows, int64_t c_out,
                             int64_t h_out, int64_t w_out, int64_t c_in, int64_t h_in, int64_t w_in,
                             int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w,
                             int64_t pad_h, int64_t pad_w, int64_t dil_h, int64_t dil_w,
                             void* workspace, int64_t workspace_bytes, int dtype, void* stream);

/* out[v, n, f] = S[v, n, f] * J(ref[n, f]); S/out: [V, n*feat], ref: [n, feat] (see vvt_act). */
int vvt_sqrt_backprop_elementwise(void* out, const void* S, const void* ref, int64_t V,
                                  int64_t n_feat, int act, double scale, int dtype, void* stream);

/* arg-max positions of a max-pool forward pass over x [planes, h_in, w_in] -> argmax [planes, h_out, w_out] (flat
 * index into h_in * w_in), the index plumbing of [BackPACK]'s max-pool Jacobian (it re-runs
 * max_pool2d(return_indices=True) on the layer input; reached via base.py:8,19).  torch's selection rule: windows
 * scanned row by row, the first of equal values wins, NaN wins.  h_out / w_out are the caller's (floor or ceil mode). */
int vvt_maxpool2d_argmax(int64_t* argmax, const void* x, int64_t planes, int64_t h_out, int64_t w_out, int64_t h_in,
                         int64_t w_in, int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w, int64_t pad_h,
                         int64_t pad_w, int64_t dil_h, int64_t dil_w, int dtype, void* stream);

/* max-pool: out[r, ch, p] = sum over output positions q whose arg-max is p of S[r, ch, q].
 * S: [V*N, ch, h_out*w_out]; argmax: [N, ch, h_out*w_out] flat index into h_in*w_in;
 * out: [V*N, ch, h_in*w_in]; row r belongs to sample r % N. */
int vvt_sqrt_backprop_maxpool2d(void* out, const void* S, const int64_t* argmax, int64_t V,
                                int64_t N, int64_t ch, int64_t h_out, int64_t w_out, int64_t h_in,
                                int64_t w_in, int64_t kh, int64_t kw, int64_t stride_h,
                                int64_t stride_w, int64_t pad_h, int64_t pad_w, int64_t dil_h,
                                int64_t dil_w, int dtype, void* stream);

/* average pool (count_include_pad = true): out[r,ch,p] = sum_{q: p in window q} S[r,ch,q]/(kh kw) */
int vvt_sqrt_backprop_avgpool2d(void* out, const void* S, int64_t rows, int64_t ch, int64_t h_out,
                                int64_t w_out, int64_t h_in, int64_t w_in, int64_t kh, int64_t kw,
                                int64_t stride_h, int64_t stride_w, int64_t pad_h, int64_t pad_w,
                                int dtype, void* stream);

/* ------------------------------------------------------------------------ *
 * (1) emitting V^T of a parameter in the coalesced [R, D_p] layout          *
 *     ([BackPACK] param_mjp(sum_batch=False), called at base.py:84-92)       *
 * ------------------------------------------------------------------------ */

/* Vt[(v,n), (o, ci, ky, kx)] = sum_{oy,ox} S[v,n,o,oy,ox] X[n, ci, oy*sh+ky*dh-ph, ox*sw+kx*dw-pw]
 * S: [V, N, c_out, h_out, w_out], X: [N, c_in, h_in, w_in] (already sub-sampled),
 * Vt: [V*N, c_out*c_in*kh*kw]
 * fp32 with workspace: im2col + re-layout (spatial index contiguous) -> batched tcgen05 GEMM per sample. */
int vvt_v_emit_conv2d(void* Vt, const void* S, const void* X, int64_t V, int64_t N, int64_t c_out,
                      int64_t h_out, int64_t w_out, int64_t c_in, int64_t h_in, int64_t w_in,
                      int64_t kh, int64_t kw, int64_t stride_h, int64_t stride_w, int64_t pad_h,
                      int64_t pad_w, int64_t dil_h, int64_t dil_w, void* workspace,
                      int64_t workspace_bytes, int dtype, void* stream);

/* Vt[r, o] = sum_x S[r, o, x];  S: [rows, c_out, spatial] */
int vvt_v_emit_bias(void* Vt, const void* S, int64_t rows, int64_t c_out, int64_t spatial,
                    int dtype, void* stream);

/* materialised Linear weight factor Vt[(v,n), o, i] = S[(v,n), o] Z[n, i]
 * (only for users who ask for the tensor; the Computations never call it). */
int vvt_v_emit_linear(void* Vt, const void* S, const void* Z, int64_t V, int64_t N, int64_t n_out,
                      int64_t n_in, int dtype, void* stream);

/* ------------------------------------------------------------------------ *
 * (2) Gram assembly                                                          *
 * ------------------------------------------------------------------------ */

/* split-K scratch for a [rows, depth] x [cols, depth]^T product (0 if none is needed) */
int64_t vvt_gram_workspace_bytes(int64_t rows, int64_t cols, int64_t depth, int dtype);

/* General strided (optionally batched) product on the same tensor-core main loop:
 *   C[b] = alpha * opA(A[b]) opB(B[b])^T + beta * C[b],   C: [M, N] with leading dimension ldc
 *   transA == 0: A is [M, K] (lda >= K)   transA != 0: A is [K, M] (lda >= M)
 *   transB == 0: B is [N, K] (ldb >= K)   transB != 0: B is [K, N] (ldb >= N)
 * fp32 uses the 3xTF32 split (fp32-grade accuracy), fp64 uses DMMA.  Stands in for the
 * torch.einsum calls of vivit/utils/gram.py and vivit/utils/ggn.py that are not named below. */
int vvt_gemm(void* C, const void* A, const void* B, int64_t M, int64_t N, int64_t K, int transA,
             int transB, int64_t lda, int64_t ldb, int64_t ldc, double alpha, double beta,
             int64_t batch, int64_t strideA, int64_t strideB, int64_t strideC, void* workspace,
             int64_t workspace_bytes, int dtype, void* stream);

/* G[R,R] += V V^T, V: [R, D].  Replaces pairwise_dot(V_t, 2) at base.py:118-124 and
 * partial_contract(V, V, (2,2)) at directional_derivatives.py:245 / directional_damped_newton.py:254
 * (vivit/utils/gram.py:206-232). */
int vvt_gram_dense_accum(void* G, const void* V, int64_t R, int64_t D, void* workspace,
                         int64_t workspace_bytes, int dtype, void* stream);

/* X[R, n_g] += V g^T, V: [R, D], g: [n_g, D].  Replaces partial_contract(V, g, (2,1)) at
 * directional_derivatives.py:246; also V_t_mat_prod (vivit/utils/gram.py:182-203). */
int vvt_gram_cross_accum(void* X, const void* V, const void* g, int64_t R, int64_t n_g, int64_t D,
                         void* workspace, int64_t workspace_bytes, int dtype, void* stream);

/* G[(c,n),(d,m)] += (sum_i Z[n,i] Z[m,i] + with_bias) * sum_o S[(c,n),o] S[(d,m),o]
 * S: [C*N, n_out], Z: [N, n_in].  Replaces ViViTGGNLinear.weight::gram_mat (linear.py:66-75);
 * with_bias=1 also folds in the bias Gram S S^T (base.py:118-124 on the bias), so S S^T is
 * formed once.  workspace >= vvt_gram_linear_workspace_bytes. */
int64_t vvt_gram_linear_workspace_bytes(int64_t C, int64_t N, int64_t n_out, int64_t n_in,
                                        int64_t n_g, int dtype);
int vvt_gram_linear_accum(void* G, const void* S, const void* Z, int64_t C, int64_t N,
                          int64_t n_out, int64_t n_in, int with_bias, void* workspace,
                          int64_t workspace_bytes, int dtype, void* stream);

/* X[(c,n), m] += (sum_i Z[n,i] Zg[m,i] + with_bias) * sum_o S[(c,n),o] Dl[m,o]
 * structured form of partial_contract(V, grad_batch, (2,1)) for a Linear layer whose
 * per-sample gradient is Dl[m,:] (x) Zg[m,:]; Dl: [n_g, n_out], Zg: [n_g, n_in]. */
int vvt_gram_cross_linear_accum(void* X, const void* S, const void* Z, const void* Dl,
                                const void* Zg, int64_t C, int64_t N, int64_t n_g, int64_t n_out,
                                int64_t n_in, int with_bias, void* workspace,
                                int64_t workspace_bytes, int dtype, void* stream);

/* out[n, d] = g[n, d] - (1/N) sum_m g[m, d]; g: [N, D] row-major, out may alias g.
 * Centering of per-sample gradients: CenteredBatchGrad.param_hook
 * (vivit/extensions/firstorder/batch_grad/gram_batch_grad.py:25-38) and the in-place
 * `grad_batch -= grad_batch.mean(0)` of _GramBatchGradBase.param_hook (:88-89). */
int vvt_center_rows(void* out, const void* g, int64_t N, int64_t D, int dtype, void* stream);

/* T[i] *= alpha  (the N/len(subsampling) rescale at eigh.py:245-246, eigvalsh.py:218-219) */
int vvt_scale(void* T, int64_t numel, double alpha, int dtype, void* stream);

/* The exchange step of the parameter-sharded path: G [numel_G] (the partial Gram matrix of this
 * rank's parameter shard) and X [numel_X] (its partial cross term V^T g; may be NULL / 0) are replaced by
 * alpha * (sum over the ranks of `comm`), in place, with ONE NCCL call group on `stream`.  alpha is the
 * reference's N / len(subsampling) rescale (eigh.py:245-246, eigvalsh.py:218-219) applied inside the reduction
 * (pre-multiplied sum), 1.0 otherwise.  comm: an ncclComm_t of the NCCL instance already loaded in the process
 * (torch.distributed: ProcessGroupNCCL); VVT_ERR_UNSUPPORTED if there is none.  The sums
 * G = sum_p V_p^T V_p of eigh.py:239-242, eigvalsh.py:170-183, directional_derivatives.py:245-246,348-351
 * extended over the ranks. */
int vvt_nccl_available(void);
int vvt_nccl_allreduce_gram(void* comm, void* G, int64_t numel_G, void* X, int64_t numel_X, double alpha,
                            int dtype, void* stream);

/* Y[i] += alpha * X[i]: the sum of the factors that reach a tensor used by several branches of a model,
 * ViViTGGN.accumulate_backpropagated_quantities (vivit/extensions/secondorder/vivit/__init__.py:130-133) */
int vvt_axpy(void* Y, const void* X, int64_t numel, double alpha, int dtype, void* stream);

/* ------------------------------------------------------------------------ *
 * (3) symmetric eigensolver: parallel-order block Jacobi                    *
 * ------------------------------------------------------------------------ */

int64_t vvt_syevj_workspace_bytes(int64_t R, int jobz, int dtype);
int64_t vvt_syevj_batched_workspace_bytes(int64_t R, int64_t batch, int jobz, int dtype);

/* Eigendecomposition of the symmetric POSITIVE SEMI-DEFINITE matrix G [R, R] (a Gram matrix; the upper triangle
 * is read, as symeig(upper=True)).  The solver is one-sided: it works on a Cholesky factor of G + eps I, so an
 * indefinite input is outside its contract (vivit_b200/utils/eig.py shifts such a matrix by a Gershgorin bound
 * first).  evals: [R] ascending; evecs: [R, R] with eigenvectors as COLUMNS (ignored if jobz == 0).
 * G is not modified.  info_host (host int[2], may be NULL): {sweeps used, converged flag}; converged == 0 means
 * the sweep limit was reached (the reference's symeig raises in that case, and so do the Python callers).
 * Replaces Tensor.symeig at eigh.py:248, eigvalsh.py:221, directional_derivatives.py:291,
 * directional_damped_newton.py:315.  Convergence is decided on the device; the host reads it through pinned
 * memory one sweep late, so the stream is never drained between sweeps. */
int vvt_syevj(void* evals, void* evecs, const void* G, int64_t R, int jobz, void* workspace,
              int64_t workspace_bytes, int* info_host, int dtype, void* stream);

/* The same for `batch` independent matrices of one size, G [batch, R, R] -> evals [batch, R],
 * evecs [batch, R, R]: the per-group Grams of block-diagonal param_groups (groups are independent,
 * vivit/utils/hooks.py:214-219, and share R = C * N).  The problem index is a grid dimension of every kernel,
 * each problem stops on its own convergence flag.  info_host: int[2 * batch]. */
int vvt_syevj_batched(void* evals, void* evecs, const void* G, int64_t R, int64_t batch, int jobz,
                      void* workspace, int64_t workspace_bytes, int* info_host, int dtype, void* stream);

/* The same solve with the ROUNDS of the two-level path distributed over the ranks of an NCCL communicator (the
 * GPUs of one box): every rank passes the same G (the all-reduced Gram of the parameter-sharded path, SURVEY 8e:
 * the full-network Gram of one group, R = C * N = 10240 at config 5) and receives the same evals / evecs.  The
 * block pairs of a round are independent, so rank g works on a contiguous slice of the pair index; between two
 * rounds of the round-robin tournament a rank hands one 64-column block to each neighbour (grouped ncclSend /
 * ncclRecv between the block slots of the factor, on `stream`), the sweep's rotation count is summed with an
 * 8-byte all-reduce so that every rank takes the same convergence decision, and the rotated factor is put
 * together again by one all-reduce before the (replicated) Rayleigh quotients and refinement step.  Problems off
 * the two-level path (fp64, R < 4096) are solved by every rank for itself -- same results, no exchange.
 * comm: an ncclComm_t of the NCCL instance already loaded in the process (as vvt_nccl_allreduce_gram); rank and
 * size are read from it.  p2p != 0: the peer-memory arenas below are mapped for exactly the ranks of comm (same
 * rank order) -- blocks then change owner by direct stores.  Collective: every rank of the communicator must call
 * it, with the same R / jobz / p2p / dtype.
 * Not in the reference (single process); it replaces the same Tensor.symeig call sites as vvt_syevj. */
int64_t vvt_syevj_dist_workspace_bytes(int64_t R, int jobz, int dtype, int world);

/* Peer-memory arena of vvt_syevj_dist (optional; without it the blocks travel by ncclSend / ncclRecv).  One per
 * process and device: vvt_dist_arena_alloc allocates `bytes` (vvt_dist_arena_bytes_for(R) holds the factor of an
 * R-column problem) and writes the 64-byte CUDA IPC handle of the allocation to handle_out; the caller gathers the
 * handles of all ranks (any transport) and passes them, in rank order, to vvt_dist_arena_open, which maps the peers'
 * arenas.  With the arenas mapped a rank WRITES the blocks that change owner into the next owner's factor over
 * NVLink from its own kernel and raises a flag there; the receiver's stream waits on the flag -- no NCCL call per
 * round.  To grow or drop an arena: vvt_dist_arena_close_peers on every rank, a barrier, vvt_dist_arena_free. */
int64_t vvt_dist_arena_bytes_for(int64_t R);
int64_t vvt_dist_arena_bytes(void);
int vvt_dist_arena_alloc(int64_t bytes, void* handle_out);
int vvt_dist_arena_open(const void* handles, int world, int rank);
int vvt_dist_arena_close_peers(void);
int vvt_dist_arena_free(void);
int vvt_syevj_dist(void* comm, void* evals, void* evecs, const void* G, int64_t R, int jobz, void* workspace,
                   int64_t workspace_bytes, int* info_host, int p2p, int dtype, void* stream);

/* TEST HOOK, not a reference interface (host logic only, no GPU needed): the block hand-over plan of
 * vvt_syevj_dist before round `round` (-1: the intra round, 0 .. nbw - 2: the tournament) for nbw wide blocks over
 * `world` ranks.  owner [nbw] (in / out): rank holding each block, -1 = every rank; moves_out [3 * max_moves]:
 * {block, from, to} triples in the order every rank issues them; pair_rank_out [nbw / 2] (may be NULL): the rank
 * that works on each pair of the round. */
int vvt_dbg_dist_plan(int nbw, int world, int round, int* owner, int* moves_out, int max_moves, int* n_moves_out,
                      int* pair_rank_out);

/* TEST HOOK, not a reference interface: ONE round of the two-level solver used for fp32 problems of 4096
 * columns and more, on a caller-provided row-major factor L [Np, Np] (Np a multiple of 128), so that the three
 * kernels of a round (tcgen05 pair Grams with MN-major operands, 128 x 128 rotation kernel, tcgen05 apply) can
 * be checked one by one: H_out [Np/128, 128, 128] += pair Grams (zero it first), Qt_out [Np/128, 128, 128] =
 * Q^T of every rotated pair, flag_out [Np/128] = pair rotated, L <- L Q in place.  round: -1 (rotations inside
 * the wide blocks) or 0 .. Np/64 - 2 (round-robin pairing).  Allocates scratch and synchronises: tests only. */
int vvt_dbg_wide_round(float* L, float* H_out, float* Qt_out, int* flag_out, int64_t Np, int round, void* stream);

/* mask[i] = !isclose(evals[i], 0, rtol, atol) = |evals[i]| > atol  (vivit/utils/eig.py:111-134);
 * mask: uint8 [R]; count_host (host, may be NULL) receives the number kept (synchronises). */
int vvt_filter_nonzero(uint8_t* mask, const void* evals, int64_t R, double atol, double rtol,
                       int64_t* count_host, int dtype, void* stream);

/* ------------------------------------------------------------------------ *
 * (4) back-transform, normalisation, directional derivatives, Newton step   *
 * ------------------------------------------------------------------------ */

/* E[K, D] = U[K, R] V[R, D];  norm2[k] += sum_d E[k,d]^2  (norm2: float64 [K] for either dtype,
 * initialised by the caller; may be NULL).
 * Replaces Vmp (vivit/utils/ggn.py:94-115) at base.py:96-105 / eigh.py:268-269 and the squared-norm
 * pass of normalize (vivit/linalg/utils.py:73). */
int vvt_backtransform_dense(void* E, void* norm2, const void* U, const void* V, int64_t K,
                            int64_t R, int64_t D, int dtype, void* stream);

/* E[k, o, i] = sum_{c,n} U[k,(c,n)] S[(c,n),o] Z[n,i]  (linear.py:44-53), norm2 as above.
 * Eb [K, n_out] (may be NULL) additionally receives the bias part sum_{c,n} U S.
 * workspace: K*N*n_out elements. */
int vvt_backtransform_linear(void* E, void* Eb, void* norm2, const void* U, const void* S,
                             const void* Z, int64_t K, int64_t C, int64_t N, int64_t n_out,
                             int64_t n_in, void* workspace, int64_t workspace_bytes, int dtype,
                             void* stream);

/* out[f,(c,n)] = sum_{o,i} S[(c,n),o] Mat[f,o,i] Z[n,i]  -- V^T applied to F stacked vectors
 * (ViViTGGNLinear.weight::V_t_mat_prod, linear.py:55-64).  workspace: F*N*n_out elements. */
int vvt_vt_mat_prod_linear(void* out, const void* S, const void* Z, const void* Mat, int64_t F,
                           int64_t C, int64_t N, int64_t n_out, int64_t n_in, void* workspace,
                           int64_t workspace_bytes, int dtype, void* stream);

/* inv[k] = 1/sqrt(norm2[k]) applied: E[k, :] *= inv[k]  (vivit/linalg/utils.py:75-76) */
int vvt_scale_rows_rsqrt(void* E, const void* norm2 /* float64 [K] */, int64_t K, int64_t D, int dtype, void* stream);

/* gammas[m,k]  = corr*N * sum_r X[r,m] U[r,k] / sqrt(evals[k])                  (directional_derivatives.py:302-317)
 * lambdas[n,k] = N_ggn * sum_c ( sum_j corr^2 G[(c,n),j] U[j,k] )^2 / evals[k]  (directional_derivatives.py:322-325)
 * G: [R,R] un-rescaled accumulation V^T V, X: [R, n_g] un-rescaled V^T g, U: [R, K] kept Gram eigenvectors
 * (columns), evals: [K]; corr = sqrt(N / N_ggn). gammas: [n_g, K], lambdas: [N_ggn, K]. */
int vvt_dirderiv_epilogue(void* gammas, void* lambdas, const void* G, const void* X, const void* U,
                          const void* evals, int64_t C, int64_t N_ggn, int64_t n_g, int64_t K,
                          int64_t N, void* workspace, int64_t workspace_bytes, int dtype,
                          void* stream);

/* coef[k] = -mean_m gammas[m,k] / (mean_n lambdas[n,k] + deltas[k]) / sqrt(evals[k])
 * v[r]    = corr * sum_k U[r,k] coef[k]            (directional_damped_newton.py:353-366) */
int vvt_newton_coeff(void* v, const void* U, const void* gammas, const void* lambdas,
                     const void* deltas, const void* evals, int64_t R, int64_t K, int64_t n_g,
                     int64_t N_ggn, double corr, int dtype, void* stream);

/* step[d] = sum_r v[r] V[r, d]  (directional_damped_newton.py:370-373) */
int vvt_v_apply_dense(void* step, const void* v, const void* V, int64_t R, int64_t D, int dtype,
                      void* stream);

/* step[o,i] = sum_{c,n} v[(c,n)] S[(c,n),o] Z[n,i]; stepb[o] (may be NULL) = sum v S */
int vvt_v_apply_linear(void* step, void* stepb, const void* v, const void* S, const void* Z,
                       int64_t C, int64_t N, int64_t n_out, int64_t n_in, void* workspace,
                       int64_t workspace_bytes, int dtype, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VIVIT_B200_H */
