// bn_fast.cu -- fused kernel plan.  (Round-1 stage 1: no fused pattern registered yet; the
// engine runs the generic one-kernel-per-op plan.)
#include "bn_fast.cuh"

namespace bn {

void fast_plan_build(FastPlan& fp, const bn_blob_header* hdr, const bn_blob_tensor* tensors, const bn_blob_op* ops,
                     uint8_t* d_blob) {
  fp.ok = false;
  fp.hdr = hdr; fp.tensors = tensors; fp.ops = ops; fp.d_blob = d_blob;
}
int fast_plan_alloc_workspace(FastPlan&, int, size_t* total) { *total = 0; return 0; }
void fast_plan_free_workspace(FastPlan& fp) {
  for (void* p : fp.bufs) if (p) cudaFree(p);
  fp.bufs.clear(); fp.tap_ids.clear(); fp.tap_bytes.clear(); fp.wave = 0;
}
int fast_run_pcm(FastPlan&, const int16_t*, const float*, int, float*, int, int, cudaStream_t, int64_t*) { return BN_ERR_UNSUPPORTED; }
int fast_run_spec(FastPlan&, const float*, int, float*, int, int, cudaStream_t, int64_t*) { return BN_ERR_UNSUPPORTED; }
int fast_dump_tensor(FastPlan&, int, int, void*, size_t) { return BN_ERR_UNSUPPORTED; }

}  // namespace bn
