// bn_generic_tc.cu -- tensor-core pointwise convolutions for the GENERIC (one kernel per op) plan.
//
// Graphs outside the fused plan of the shipped checkpoint (BASELINE configs 3 / 4: raw filterbank frontend, SE gates, inverted
// residuals, attention pooling -- reference models/blocks.py:27-175, models/frontend.py:347-358) run one kernel per lowered op.
// Their 1x1 convolutions hold 85 - 95 % of the MACs; every one whose shape the tcgen05 GEMM of bn_pw_tc.cu supports and whose
// requantisation lies in the proven closed-form domain is routed through that kernel instead of the one-thread-per-output
// reference kernel (same integer results; covered by the bit-exact tests of tests/test_ptq.py).
#include "bn_generic_tc.cuh"

#include <cstdlib>
#include <vector>

#include "bn_common.cuh"

namespace bn {

static void* up(GenAccel* a, const void* src, size_t n) {
  void* d = nullptr;
  if (cudaMalloc(&d, n ? n : 4) != cudaSuccess) return nullptr;
  cudaMemcpy(d, src, n, cudaMemcpyHostToDevice);
  a->owned.push_back(d);
  return d;
}

GenAccel* gen_accel_build(const uint8_t* h_blob, const bn_blob_header* hdr, const bn_blob_tensor* T, const bn_blob_op* ops) {
  GenAccel* a = new GenAccel();
  a->ops.resize(hdr->n_ops);
  if (getenv("BN_GENERIC_TC") && atoi(getenv("BN_GENERIC_TC")) == 0) return a;
  for (uint32_t i = 0; i < hdr->n_ops; i++) {
    const bn_blob_op& op = ops[i];
    if (op.kind != BN_OP_CONV2D) continue;
    const int32_t* p = op.p;
    if (p[BN_CONV_KH] != 1 || p[BN_CONV_KW] != 1 || p[BN_CONV_SH] != 1 || p[BN_CONV_SW] != 1 || p[BN_CONV_PAD_T] != 0 || p[BN_CONV_PAD_L] != 0) continue;
    const int K = p[BN_CONV_CIN], N = p[BN_CONV_COUT];
    if (!pw_tc_supported(K, N)) continue;
    const bn_blob_tensor& ti = T[op.in[0]];
    const bn_blob_tensor& to = T[op.out];
    if (ti.dims[2] != K || to.dims[2] != N || ti.dims[0] != to.dims[0] || ti.dims[1] != to.dims[1]) continue;
    const int8_t* w = (const int8_t*)(h_blob + op.off[0]);            // [N][K]
    const int32_t* bias = (const int32_t*)(h_blob + op.off[1]);
    const int32_t* mult = (const int32_t*)(h_blob + op.off[2]);
    const int32_t* shift = (const int32_t*)(h_blob + op.off[3]);
    const int zp = p[BN_CONV_IN_ZP];
    const long xmax = (127 - zp) > (zp + 128) ? (127 - zp) : (zp + 128);
    std::vector<int> m(mult, mult + N), sh(shift, shift + N), bf(N);
    bool ok = true;
    for (int n = 0; n < N && ok; n++) {
      // the closed-form requantisation of the GEMM epilogue needs a right shift in [1, 31] and |SRDHM(acc)| + 2^(n-1) < 2^31
      if (m[n] == 0) sh[n] = -1;
      if (sh[n] > -1 || sh[n] < -31) { ok = false; break; }
      long wsum = 0, ws = 0;
      for (int k = 0; k < K; k++) { wsum += labs((long)w[(long)n * K + k]); ws += w[(long)n * K + k]; }
      const long amax = labs((long)bias[n]) + wsum * xmax;
      const long vmax = (long)(((__int128)amax * m[n] + (1ll << 30)) >> 31) + 1;
      if (vmax + (1l << (-sh[n] - 1)) >= (1l << 31)) ok = false;
      bf[n] = (int)((long)bias[n] - (long)zp * ws);
    }
    if (!ok) continue;
    GenAccelOp& g = a->ops[i];
    PwTcParams& Tc = g.tc;
    std::vector<uint8_t> img;
    pw_tc_weight_image(w, K, N, img, &Tc.KP, &Tc.RW);
    Tc.w_img = (const uint8_t*)up(a, img.data(), img.size());
    Tc.bias = (const int*)up(a, bf.data(), bf.size() * 4);
    Tc.mult = (const int*)up(a, m.data(), m.size() * 4);
    Tc.shift = (const int*)up(a, sh.data(), sh.size() * 4);
    Tc.K = K; Tc.N = N;
    int l = 0; while ((16 << l) < K) l++;
    Tc.cpr_log = l;
    int cols = 32; while (cols < 2 * N) cols <<= 1;
    Tc.tmem_cols = cols;
    Tc.out_zp = p[BN_CONV_OUT_ZP]; Tc.act_min = p[BN_CONV_ACT_MIN]; Tc.act_max = p[BN_CONV_ACT_MAX];
    Tc.has_add = 0;
    g.pw = Tc.w_img && Tc.bias && Tc.mult && Tc.shift;
    if (g.pw) a->n_pw++;
  }
  return a;
}

void gen_accel_destroy(GenAccel* a) {
  if (!a) return;
  for (void* p : a->owned) cudaFree(p);
  delete a;
}

}  // namespace bn
