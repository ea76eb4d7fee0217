// The one exchange step of the path: partial Gram matrices (and partial cross terms V^T g) of the
// parameter shards are summed over the ranks with ONE NCCL call group over NVLink, with the reference's
// sub-sampling rescale N / len(subsampling) (vivit/linalg/eigh.py:245-246, eigvalsh.py:218-219) folded into the
// reduction itself (ncclRedOpCreatePreMulSum: every rank's contribution is multiplied on its way in) -- no
// separate scale launch, no packing copy.
//
// NCCL is not linked: the communicator belongs to the caller (torch.distributed's ProcessGroupNCCL hands out
// its ncclComm_t), so the calls must go to the NCCL instance that is already loaded in the process.  It is
// looked up with dlopen(RTLD_NOLOAD); without it the entry point reports VVT_ERR_UNSUPPORTED.
#include "nccl_dyn.cuh"

using namespace vvt;

extern "C" {

int vvt_nccl_available(void) { return nccl().ok ? 1 : 0; }

int vvt_nccl_allreduce_gram(void* comm, void* G, int64_t numel_G, void* X, int64_t numel_X, double alpha, int dtype,
                            void* stream) {
  VVT_REQUIRE(numel_G >= 0 && numel_X >= 0, "negative size");
  VVT_REQUIRE(comm != nullptr, "null communicator");
  VVT_REQUIRE(dtype == VVT_F32 || dtype == VVT_F64, "unknown dtype");
  VVT_REQUIRE((numel_G == 0 || G) && (numel_X == 0 || X), "null pointer");
  const Nccl& n = nccl();
  if (!n.ok) return fail(VVT_ERR_UNSUPPORTED, "%s: no NCCL library is loaded in this process", __func__);
  if (numel_G == 0 && numel_X == 0) return VVT_OK;
  const int nccl_dtype = dtype == VVT_F32 ? kNcclFloat32 : kNcclFloat64;
  int op = kNcclSum;
  float a32 = float(alpha);
  double a64 = alpha;
  const bool premul = alpha != 1.0;
  if (premul)
    VVT_TRY(check_nccl(n.premul(&op, dtype == VVT_F32 ? (void*)&a32 : (void*)&a64, nccl_dtype, kNcclScalarHostImmediate, comm),
                       __func__));
  cudaStream_t s = as_stream(stream);
  int st = check_nccl(n.group_start(), __func__);
  if (st == VVT_OK && numel_G > 0) st = check_nccl(n.all_reduce(G, G, size_t(numel_G), nccl_dtype, op, comm, s), __func__);
  if (st == VVT_OK && numel_X > 0) st = check_nccl(n.all_reduce(X, X, size_t(numel_X), nccl_dtype, op, comm, s), __func__);
  const int st_end = check_nccl(n.group_end(), __func__);
  if (premul) n.op_destroy(op, comm);
  return st != VVT_OK ? st : st_end;
}

}  // extern "C"
