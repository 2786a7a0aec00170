// bn_generic_tc.cu -- tensor-core pointwise convolutions for the GENERIC (one kernel per op) plan.
//
// Graphs outside the fused plan of the shipped checkpoint (BASELINE configs 3 / 4: raw filterbank frontend, SE gates, inverted
// residuals, attention pooling -- reference models/blocks.py:27-175, models/frontend.py:347-358) run one kernel per lowered op.
// Their 1x1 convolutions hold 85 - 95 % of the MACs; every one whose shape the tcgen05 GEMM of bn_pw_tc.cu supports and whose
// requantisation lies in the proven closed-form domain is routed through that kernel instead of the one-thread-per-output
// reference kernel (same integer results; covered by the bit-exact tests of tests/test_ptq.py).  The 1x1 convolutions the GEMM
// has no build for (K or N = 512 in the inverted-residual expansions) take the tiled dp4a GEMM of the layer path, and every
// depthwise 3x3 (stride 1 | 2) the register-window kernel of the layer path (k_pw / k_dw3x3, bn_fast.cu).
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

// closed-form requantisation domain (right shift in [1, 31], |SRDHM(acc)| + 2^(n-1) < 2^31) of channel n with sum |w| = wsum
static bool rq_domain_ok(long bias, long wsum, long xmax, int mult, int& shift) {
  if (mult == 0) shift = -1;
  if (shift > -1 || shift < -31) return false;
  const long amax = labs(bias) + wsum * xmax;
  const long vmax = (long)(((__int128)amax * mult + (1ll << 30)) >> 31) + 1;
  return vmax + (1l << (-shift - 1)) < (1l << 31);
}

static void build_dw(GenAccel* a, GenAccelOp& g, const uint8_t* h_blob, const bn_blob_tensor* T, const bn_blob_op& op) {
  const int32_t* p = op.p;
  const int C = p[BN_CONV_CIN];
  if (p[BN_CONV_KH] != 3 || p[BN_CONV_KW] != 3 || p[BN_CONV_SH] != p[BN_CONV_SW] || (p[BN_CONV_SH] != 1 && p[BN_CONV_SH] != 2)) return;
  if (C % 4 || p[BN_CONV_COUT] != C) return;
  const bn_blob_tensor& ti = T[op.in[0]];
  const bn_blob_tensor& to = T[op.out];
  if (ti.dims[2] != C || to.dims[2] != C) return;
  const int8_t* w = (const int8_t*)(h_blob + op.off[0]);             // [3][3][C]
  const int32_t* bias = (const int32_t*)(h_blob + op.off[1]);
  const int32_t* mult = (const int32_t*)(h_blob + op.off[2]);
  const int32_t* shift = (const int32_t*)(h_blob + op.off[3]);
  const int zp = p[BN_CONV_IN_ZP];
  const long xmax = (127 - zp) > (zp + 128) ? (127 - zp) : (zp + 128);
  std::vector<int> wm((size_t)9 * C), bf(C), m(mult, mult + C), sh(shift, shift + C);
  int fast = 1;
  for (int c = 0; c < C; c++) {
    long ws = 0, wsum = 0;
    for (int t = 0; t < 9; t++) {
      const int8_t v = w[t * C + c];
      ws += v; wsum += labs((long)v);
      wm[((size_t)t * (C / 4) + c / 4) * 4 + (c & 3)] = (int)((unsigned)(uint8_t)v << (8 * (c & 3)));
    }
    bf[c] = (int)((long)bias[c] - (long)zp * ws);
    if (!rq_domain_ok(bias[c], wsum, xmax, m[c], sh[c])) fast = 0;
  }
  DwParams& D = g.dwp;
  D.wm = (const int*)up(a, wm.data(), wm.size() * 4);
  D.bias = (const int*)up(a, bf.data(), bf.size() * 4);
  D.mult = (const int*)up(a, m.data(), m.size() * 4);
  D.shift = (const int*)up(a, sh.data(), sh.size() * 4);
  D.C = C; D.ih = ti.dims[0]; D.iw = ti.dims[1]; D.oh = to.dims[0]; D.ow = to.dims[1];
  D.sh = p[BN_CONV_SH]; D.sw = p[BN_CONV_SW]; D.pt = p[BN_CONV_PAD_T]; D.pl = p[BN_CONV_PAD_L];
  D.in_zp = zp; D.out_zp = p[BN_CONV_OUT_ZP]; D.act_min = p[BN_CONV_ACT_MIN]; D.act_max = p[BN_CONV_ACT_MAX];
  D.fast = fast;
  g.dw = D.wm && D.bias && D.mult && D.shift;
  if (g.dw) a->n_dw++;
}

// 1x1 convolution on the tiled dp4a GEMM: weights as [K/4][N] words (4 consecutive k of channel n), bias folded with the input zero point
static void build_pwc(GenAccel* a, GenAccelOp& g, const uint8_t* h_blob, const bn_blob_op& op, int K, int N) {
  const int32_t* p = op.p;
  if (K % 4 || N % 4 || pw_cuda_core_smem(K, N) > 225 * 1024) return;
  const int8_t* w = (const int8_t*)(h_blob + op.off[0]);             // [N][K]
  const int32_t* bias = (const int32_t*)(h_blob + op.off[1]);
  const int32_t* mult = (const int32_t*)(h_blob + op.off[2]);
  const int32_t* shift = (const int32_t*)(h_blob + op.off[3]);
  const int zp = p[BN_CONV_IN_ZP];
  const long xmax = (127 - zp) > (zp + 128) ? (127 - zp) : (zp + 128);
  const int KW = K / 4;
  std::vector<int> wt((size_t)KW * N), bf(N), m(mult, mult + N), sh(shift, shift + N);
  int fast = 1;
  for (int n = 0; n < N; n++) {
    long ws = 0, wsum = 0;
    for (int k = 0; k < K; k++) {
      const int8_t v = w[(size_t)n * K + k];
      ws += v; wsum += labs((long)v);
      wt[(size_t)(k / 4) * N + n] |= (int)((unsigned)(uint8_t)v << (8 * (k & 3)));
    }
    bf[n] = (int)((long)bias[n] - (long)zp * ws);
    if (!rq_domain_ok(bias[n], wsum, xmax, m[n], sh[n])) fast = 0;
  }
  PwParams& P = g.pwp;
  P.wt = (const int*)up(a, wt.data(), wt.size() * 4);
  P.bias = (const int*)up(a, bf.data(), bf.size() * 4);
  P.mult = (const int*)up(a, m.data(), m.size() * 4);
  P.shift = (const int*)up(a, sh.data(), sh.size() * 4);
  P.K = K; P.N = N;
  P.out_zp = p[BN_CONV_OUT_ZP]; P.act_min = p[BN_CONV_ACT_MIN]; P.act_max = p[BN_CONV_ACT_MAX];
  P.fast = fast; P.has_add = 0; P.lut_res = nullptr; P.lut_conv = nullptr;
  P.add_mo = 0; P.add_so = 0; P.add_out_zp = 0; P.add_act_min = 0; P.add_act_max = 0;
  g.pwc = P.wt && P.bias && P.mult && P.shift;
  if (g.pwc) a->n_pwc++;
}

// MEAN -> FULLY_CONNECTED -> FULLY_CONNECTED -> LOGISTIC chained through their outputs: the squeeze-and-excitation gate
static void build_se(GenAccel* a, uint32_t i, const uint8_t* d_blob, const bn_blob_header* hdr, const bn_blob_tensor* T, const bn_blob_op* ops) {
  if (i + 3 >= hdr->n_ops) return;
  const bn_blob_op& mo = ops[i]; const bn_blob_op& f1 = ops[i + 1]; const bn_blob_op& f2 = ops[i + 2]; const bn_blob_op& lg = ops[i + 3];
  if (f1.kind != BN_OP_FC || f2.kind != BN_OP_FC || lg.kind != BN_OP_LOGISTIC) return;
  if (f1.in[0] != mo.out || f2.in[0] != f1.out || lg.in[0] != f2.out) return;
  const bn_blob_tensor& ti = T[mo.in[0]];
  SeParams P{};
  P.npix = mo.p[BN_MEAN_COUNT]; P.C = ti.dims[2]; P.C1 = f1.p[BN_CONV_COUT];
  if (ti.dims[0] * ti.dims[1] != P.npix || f1.p[BN_CONV_CIN] != P.C || f2.p[BN_CONV_CIN] != P.C1 || f2.p[BN_CONV_COUT] != P.C) return;
  P.mean_in_zp = mo.p[BN_MEAN_IN_ZP]; P.mean_out_zp = mo.p[BN_MEAN_OUT_ZP];
  P.mean_mult = mo.p[BN_MEAN_MULT]; P.mean_shift = mo.p[BN_MEAN_SHIFT]; P.mean_mult_n = mo.p[BN_MEAN_MULT_N]; P.mean_shift_n = mo.p[BN_MEAN_SHIFT_N];
  P.in_scale = ti.scale; P.out_scale = T[mo.out].scale;
  P.w1 = (const int8_t*)(d_blob + f1.off[0]); P.b1 = (const int32_t*)(d_blob + f1.off[1]); P.m1 = (const int32_t*)(d_blob + f1.off[2]); P.s1 = (const int32_t*)(d_blob + f1.off[3]);
  P.fc1_in_zp = f1.p[BN_CONV_IN_ZP]; P.fc1_out_zp = f1.p[BN_CONV_OUT_ZP]; P.fc1_act_min = f1.p[BN_CONV_ACT_MIN]; P.fc1_act_max = f1.p[BN_CONV_ACT_MAX];
  P.w2 = (const int8_t*)(d_blob + f2.off[0]); P.b2 = (const int32_t*)(d_blob + f2.off[1]); P.m2 = (const int32_t*)(d_blob + f2.off[2]); P.s2 = (const int32_t*)(d_blob + f2.off[3]);
  P.fc2_in_zp = f2.p[BN_CONV_IN_ZP]; P.fc2_out_zp = f2.p[BN_CONV_OUT_ZP]; P.fc2_act_min = f2.p[BN_CONV_ACT_MIN]; P.fc2_act_max = f2.p[BN_CONV_ACT_MAX];
  P.lut = (const int8_t*)(d_blob + lg.off[0]);
  if (!se_supported(P) || (f1.off[0] & 3) || (f2.off[0] & 3)) return;   // dp4a reads the weight rows as 32-bit words
  a->ops[i].sep = P;
  a->ops[i].se = true;
  a->n_se++;
}

GenAccel* gen_accel_build(const uint8_t* h_blob, const uint8_t* d_blob, const bn_blob_header* hdr, const bn_blob_tensor* T, const bn_blob_op* ops) {
  GenAccel* a = new GenAccel();
  a->ops.resize(hdr->n_ops);
  if (getenv("BN_GENERIC_TC") && atoi(getenv("BN_GENERIC_TC")) == 0) return a;
  for (uint32_t i = 0; i < hdr->n_ops; i++) {
    const bn_blob_op& op = ops[i];
    if (op.kind == BN_OP_DWCONV2D) {
      build_dw(a, a->ops[i], h_blob, T, op);
      // whole DS block: the 1x1 convolution behind it (sole user of the depthwise output) and, if present, the residual ADD of
      // the block input (sole user of the convolution output)
      auto users = [&](int slot) { int u = 0; for (uint32_t j = 0; j < hdr->n_ops; j++) for (int k = 0; k < (int)ops[j].n_in; k++) u += ops[j].in[k] == slot; return u; };
      if (i + 1 < hdr->n_ops && op.p[BN_CONV_KH] == 3 && op.p[BN_CONV_KW] == 3) {
        const bn_blob_op& pw = ops[i + 1];
        const bool pw_ok = pw.kind == BN_OP_CONV2D && pw.p[BN_CONV_KH] == 1 && pw.p[BN_CONV_KW] == 1 && pw.p[BN_CONV_SH] == 1 && pw.p[BN_CONV_SW] == 1 &&
                           pw.in[0] == op.out && users(op.out) == 1 && op.out != (int)hdr->output_tensor;
        if (pw_ok) {
          int add_op = -1;
          if (i + 2 < hdr->n_ops && ops[i + 2].kind == BN_OP_ADD && ops[i + 2].p[BN_ADD_BCAST] == 0 && ops[i + 2].in[0] == op.in[0] &&
              ops[i + 2].in[1] == pw.out && users(pw.out) == 1 && pw.p[BN_CONV_CIN] == pw.p[BN_CONV_COUT] && op.p[BN_CONV_SH] == 1)
            add_op = (int)i + 2;
          GenAccelOp& g = a->ops[i];
          if (ds_block_build(h_blob, T, ops, (int)i, (int)i + 1, add_op, a->owned, g.ds, g.dsl)) {
            g.dsb = true;
            g.ds_skip = add_op >= 0 ? 2 : 1;
            g.ds_out_slot = add_op >= 0 ? ops[add_op].out : pw.out;
            a->n_dsb++;
          }
        }
      }
      continue;
    }
    if (op.kind == BN_OP_MEAN) { build_se(a, i, d_blob, hdr, T, ops); continue; }
    if (op.kind != BN_OP_CONV2D) continue;
    const int32_t* p = op.p;
    if (p[BN_CONV_KH] == 3 && p[BN_CONV_CIN] == 1) {
      a->ops[i].stem = stem_build(h_blob, T, op, a->owned, a->ops[i].stp);
      if (a->ops[i].stem) a->n_stem++;
      continue;
    }
    if (p[BN_CONV_KH] != 1 || p[BN_CONV_KW] != 1 || p[BN_CONV_SH] != 1 || p[BN_CONV_SW] != 1 || p[BN_CONV_PAD_T] != 0 || p[BN_CONV_PAD_L] != 0) continue;
    const int K = p[BN_CONV_CIN], N = p[BN_CONV_COUT];
    const bn_blob_tensor& ti = T[op.in[0]];
    const bn_blob_tensor& to = T[op.out];
    if (ti.dims[2] != K || to.dims[2] != N || ti.dims[0] != to.dims[0] || ti.dims[1] != to.dims[1]) continue;
    if (!pw_tc_supported(K, N)) { build_pwc(a, a->ops[i], h_blob, op, K, N); continue; }
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
    if (!ok) { build_pwc(a, a->ops[i], h_blob, op, K, N); continue; }
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
    // residual ADD right behind the convolution: same-shape operands, the convolution's output used by nothing else
    if (g.pw && i + 1 < hdr->n_ops && ops[i + 1].kind == BN_OP_ADD && ops[i + 1].p[BN_ADD_BCAST] == 0) {
      const bn_blob_op& ad = ops[i + 1];
      const int32_t* q = ad.p;
      const int conv_side = ad.in[0] == op.out ? 0 : (ad.in[1] == op.out ? 1 : -1);
      int users = 0;
      for (uint32_t j = 0; j < hdr->n_ops; j++)
        for (int k = 0; k < (int)ops[j].n_in; k++) users += ops[j].in[k] == op.out;
      if (conv_side >= 0 && users == 1 && ad.in[0] != ad.in[1] && op.out != (int)hdr->output_tensor) {
        const int res_slot = ad.in[1 - conv_side];
        const bn_blob_tensor& tr = T[res_slot];
        PwTcParams A = Tc;
        A.has_add = 1;
        // the kernel's ADD input 1 is the residual, input 2 the convolution
        A.add_in1_zp = conv_side == 1 ? q[BN_ADD_IN1_ZP] : q[BN_ADD_IN2_ZP];
        A.add_in2_zp = conv_side == 1 ? q[BN_ADD_IN2_ZP] : q[BN_ADD_IN1_ZP];
        A.add_m1 = conv_side == 1 ? q[BN_ADD_M1] : q[BN_ADD_M2]; A.add_n1 = -(conv_side == 1 ? q[BN_ADD_S1] : q[BN_ADD_S2]);
        A.add_m2 = conv_side == 1 ? q[BN_ADD_M2] : q[BN_ADD_M1]; A.add_n2 = -(conv_side == 1 ? q[BN_ADD_S2] : q[BN_ADD_S1]);
        A.add_out_zp = q[BN_ADD_OUT_ZP]; A.add_mo = q[BN_ADD_MO]; A.add_no = -q[BN_ADD_SO];
        A.add_act_min = q[BN_ADD_ACT_MIN]; A.add_act_max = q[BN_ADD_ACT_MAX];
        const bool dom = q[BN_ADD_LEFT_SHIFT] == 20 && A.add_n1 >= 0 && A.add_n1 <= 30 && A.add_n2 >= 0 && A.add_n2 <= 30 && A.add_no >= 0 && A.add_no <= 30;
        if (dom && !tr.is_const && tr.dims[0] == to.dims[0] && tr.dims[1] == to.dims[1] && tr.dims[2] == N) {
          g.tc_add = A; g.add_res_slot = res_slot; g.add_out_slot = ad.out; g.add_fused = true;
          a->n_add_fused++;
        }
      }
    }
  }
  return a;
}

void gen_accel_destroy(GenAccel* a) {
  if (!a) return;
  for (void* p : a->owned) cudaFree(p);
  delete a;
}

}  // namespace bn
