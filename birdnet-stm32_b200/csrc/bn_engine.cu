// bn_engine.cu -- host side of libbn_b200.so: blob loading, kernel plan, wave pipeline, C ABI.
//
// Implements include/bn_engine.h.  The reference counterpart is the TFLiteRunner
// (birdnet_stm32/models/runners.py:48-95) plus the per-file loop of evaluate()
// (birdnet_stm32/evaluation/metrics.py:117-147); here chunks of many files are processed in
// waves, host buffers are moved with double-buffered async copies, and per-file pooling runs on
// the device.  There is no CPU execution path in this library.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <ctype.h>
#include <sched.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bn_common.cuh"
#include "bn_fast.cuh"
#include "bn_generic_tc.cuh"
#include "bn_kernels.cuh"

using namespace bn;

static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

namespace bn {
// shared with the other translation units of the library (bn_features.cu): same thread-local message
int set_error(int code, const char* msg) { return set_err(code, "%s", msg); }
}  // namespace bn

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) return set_err(BN_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

struct bn_engine {
  int device = 0;
  std::vector<uint8_t> blob;
  uint8_t* d_blob = nullptr;
  const bn_blob_header* hdr = nullptr;
  const bn_blob_tensor* tensors = nullptr;
  const bn_blob_op* ops = nullptr;
  // workspace
  int wave = 0;                       // chunks the workspace is sized for
  int wave_opt = 23680;                // requested wave size = 80 x 296 resident CTAs (whole tile rounds in every kernel).  Measured on
                                       // 21.7 k chunks: wave 2368 -> 4736 -> 9472 -> whole job: 1.222 -> 1.260 -> 1.274 -> 1.291 M chunks/s
                                       // (fewer launches and ragged tails); the workspace is allocated for min(B, wave) chunks, about
                                       // 0.9 MB per chunk, i.e. up to 21 GB of the 180 GB when a call brings that many chunks
  int host_wave = 592;                 // wave size when the input lives in host memory: the upload of wave i+1 hides under the
                                       // compute of wave i, so the exposed part of a call is one wave's upload plus one wave's
                                       // compute -- small waves keep a PCIe-bound call close to the link rate (BN_OPT_HOST_WAVE)
  std::vector<void*> buf;             // per tensor slot: device buffer for one wave (const -> into d_blob)
  std::vector<void*> last_ptr;        // pointers used by the last wave (taps)
  size_t workspace_bytes = 0;
  int last_wave_B = 0;
  // io staging
  int16_t* d_pcm[2] = {nullptr, nullptr};
  float* d_peak[2] = {nullptr, nullptr};
  size_t d_pcm_cap = 0;               // chunks
  float* d_scores = nullptr;
  size_t d_scores_cap = 0;            // floats
  float* d_file_scores = nullptr;
  size_t d_file_scores_cap = 0;
  int* d_offs = nullptr;
  size_t d_offs_cap = 0;
  unsigned* d_mnmx = nullptr;
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr};
  // options
  int rounding = 0, mean_variant = 0, force_generic = 0;
  FastPlan fast;
  GenAccel* accel = nullptr;          // tensor-core pointwise convolutions of the generic plan
  Profiler prof;
  int64_t launches = 0;
};

static bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static long t_elems(const bn_blob_tensor& t) { return (long)t.dims[0] * t.dims[1] * t.dims[2]; }

// ---------------------------------------------------------------------------------------------
// creation
// ---------------------------------------------------------------------------------------------
static int validate_blob(const bn_engine* e) {
  const bn_blob_header* h = e->hdr;
  for (uint32_t i = 0; i < h->n_ops; i++) {
    const bn_blob_op& op = e->ops[i];
    switch (op.kind) {
      case BN_OP_QUANTIZE: case BN_OP_DEQUANTIZE: case BN_OP_REQUANT: case BN_OP_TRANSPOSE: case BN_OP_SLICE:
      case BN_OP_FILL: case BN_OP_CONCAT: case BN_OP_CONV2D: case BN_OP_DWCONV2D: case BN_OP_FC: case BN_OP_ADD:
      case BN_OP_MUL: case BN_OP_MEAN: case BN_OP_LOGISTIC: case BN_OP_RESHAPE: case BN_OP_PAD: case BN_OP_SOFTMAX: case BN_OP_SUM:
        break;
      default:
        return set_err(BN_ERR_UNSUPPORTED, "op %u (tflite op %d): kind %d has no CUDA kernel", i, op.tfl_index, op.kind);
    }
    if (op.out < 0 || (uint32_t)op.out >= h->n_tensors) return set_err(BN_ERR_BLOB, "op %u: bad output slot", i);
    for (int k = 0; k < op.n_in; k++)
      if (op.in[k] < 0 || (uint32_t)op.in[k] >= h->n_tensors) return set_err(BN_ERR_BLOB, "op %u: bad input slot", i);
  }
  return 0;
}

extern "C" int bn_create(const void* blob, size_t nbytes, int device, bn_engine** out) {
  if (!blob || !out) return set_err(BN_ERR_ARG, "bn_create: NULL argument");
  *out = nullptr;
  if (nbytes < sizeof(bn_blob_header)) return set_err(BN_ERR_BLOB, "blob too small (%zu bytes)", nbytes);
  const bn_blob_header* h = (const bn_blob_header*)blob;
  if (memcmp(h->magic, "BNB200\0\0", 8) != 0) return set_err(BN_ERR_BLOB, "bad blob magic");
  if (h->version != BN_BLOB_VERSION) return set_err(BN_ERR_BLOB, "blob version %u, engine expects %u", h->version, BN_BLOB_VERSION);
  if (h->total_bytes != nbytes) return set_err(BN_ERR_BLOB, "blob size mismatch (%llu vs %zu)", (unsigned long long)h->total_bytes, nbytes);
  if (h->tensors_off + (uint64_t)h->n_tensors * sizeof(bn_blob_tensor) > nbytes ||
      h->ops_off + (uint64_t)h->n_ops * sizeof(bn_blob_op) > nbytes)
    return set_err(BN_ERR_BLOB, "blob tables out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(BN_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return set_err(BN_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  bn_engine* e = new bn_engine();
  e->device = device;
  e->blob.assign((const uint8_t*)blob, (const uint8_t*)blob + nbytes);
  e->hdr = (const bn_blob_header*)e->blob.data();
  e->tensors = (const bn_blob_tensor*)(e->blob.data() + e->hdr->tensors_off);
  e->ops = (const bn_blob_op*)(e->blob.data() + e->hdr->ops_off);
  int rc = validate_blob(e);
  if (rc) { delete e; return rc; }
  if (e->hdr->frontend_kind == BN_FE_HYBRID) {
    const bn_blob_tensor& ti = e->tensors[e->hdr->input_tensor];
    if (e->hdr->n_fft != 512) { delete e; return set_err(BN_ERR_UNSUPPORTED, "hybrid frontend: n_fft %u (only 512)", e->hdr->n_fft); }
    if (t_elems(ti) != (long)(e->hdr->n_fft / 2 + 1) * e->hdr->spec_width) {
      delete e;
      return set_err(BN_ERR_BLOB, "graph input has %ld elements, frontend produces %u x %u", t_elems(ti), e->hdr->n_fft / 2 + 1, e->hdr->spec_width);
    }
  }
  cudaError_t ce = cudaMalloc(&e->d_blob, nbytes);
  if (ce == cudaSuccess) ce = cudaMemcpy(e->d_blob, blob, nbytes, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && ce == cudaSuccess; i++) {
    ce = cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming);
  }
  if (ce != cudaSuccess) { bn_destroy(e); return set_err(BN_ERR_CUDA, "bn_create: %s", cudaGetErrorString(ce)); }
  e->buf.assign(e->hdr->n_tensors, nullptr);
  e->last_ptr.assign(e->hdr->n_tensors, nullptr);
  fast_plan_build(e->fast, e->blob.data(), e->hdr, e->tensors, e->ops, e->d_blob);
  // (the generic plan's accelerated-op records are built on its first use: run_generic)
  { int sms = 148; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) e->fast.num_sms = sms; }
  *out = e;
  return BN_OK;
}

static void free_workspace(bn_engine* e) {
  for (uint32_t i = 0; i < e->hdr->n_tensors && i < e->buf.size(); i++)
    if (!e->tensors[i].is_const && e->buf[i]) { cudaFree(e->buf[i]); e->buf[i] = nullptr; }
  fast_plan_free_workspace(e->fast);
  if (e->d_mnmx) { cudaFree(e->d_mnmx); e->d_mnmx = nullptr; }
  e->wave = 0;
  e->workspace_bytes = 0;
}

extern "C" void bn_destroy(bn_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->hdr) free_workspace(e);
  fast_plan_destroy(e->fast);
  gen_accel_destroy(e->accel);
  e->accel = nullptr;
  for (int i = 0; i < 2; i++) {
    if (e->d_pcm[i]) cudaFree(e->d_pcm[i]);
    if (e->d_peak[i]) cudaFree(e->d_peak[i]);
    if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
  }
  if (e->d_scores) cudaFree(e->d_scores);
  if (e->d_file_scores) cudaFree(e->d_file_scores);
  if (e->d_offs) cudaFree(e->d_offs);
  if (e->d_blob) cudaFree(e->d_blob);
  if (e->s_copy) cudaStreamDestroy(e->s_copy);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  delete e;
}

static bool use_fast(const bn_engine* e) { return e->fast.ok && !e->force_generic; }

// (re)allocate the activation workspace for `wave` chunks
static int ensure_workspace(bn_engine* e, int wave) {
  if (e->wave >= wave && e->wave > 0) return 0;
  free_workspace(e);
  size_t total = 0;
  if (use_fast(e)) {
    int rc = fast_plan_alloc_workspace(e->fast, wave, &total);
    if (rc) return set_err(BN_ERR_CUDA, "fast plan workspace allocation failed for wave %d", wave);
    // graph input / output tensors are still owned here
    for (int slot : {e->hdr->input_tensor, e->hdr->output_tensor}) {
      size_t nb = (size_t)e->tensors[slot].nbytes * wave;
      CU(cudaMalloc(&e->buf[slot], nb));
      total += nb;
    }
  } else {
    for (uint32_t i = 0; i < e->hdr->n_tensors; i++) {
      const bn_blob_tensor& t = e->tensors[i];
      if (t.is_const) { e->buf[i] = e->d_blob + t.data_off; continue; }
      size_t nb = ((size_t)t.nbytes * wave + 255) & ~(size_t)255;
      CU(cudaMalloc(&e->buf[i], nb));
      total += nb;
    }
  }
  CU(cudaMalloc(&e->d_mnmx, sizeof(unsigned) * 2 * wave));
  e->wave = wave;
  e->workspace_bytes = total;
  return 0;
}

static int ensure_io(bn_engine* e, int wave, size_t n_scores) {
  if (e->d_pcm_cap < (size_t)wave && e->hdr->chunk_len) {
    for (int i = 0; i < 2; i++) {
      if (e->d_pcm[i]) cudaFree(e->d_pcm[i]);
      if (e->d_peak[i]) cudaFree(e->d_peak[i]);
      CU(cudaMalloc(&e->d_pcm[i], sizeof(int16_t) * (size_t)e->hdr->chunk_len * wave + 16));
      CU(cudaMalloc(&e->d_peak[i], sizeof(float) * wave));
    }
    e->d_pcm_cap = wave;
  }
  if (e->d_scores_cap < n_scores) {
    if (e->d_scores) cudaFree(e->d_scores);
    CU(cudaMalloc(&e->d_scores, sizeof(float) * n_scores));
    e->d_scores_cap = n_scores;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// generic plan: one kernel per op
// ---------------------------------------------------------------------------------------------
static void fill_conv(const bn_engine* e, const bn_blob_op& op, ConvParams& P) {
  const int32_t* p = op.p;
  const bn_blob_tensor& ti = e->tensors[op.in[0]];
  const bn_blob_tensor& to = e->tensors[op.out];
  P.w = (const int8_t*)(e->d_blob + op.off[0]);
  P.bias = (const int32_t*)(e->d_blob + op.off[1]);
  P.mult = (const int32_t*)(e->d_blob + op.off[2]);
  P.shift = (const int32_t*)(e->d_blob + op.off[3]);
  P.kh = p[BN_CONV_KH]; P.kw = p[BN_CONV_KW]; P.sh = p[BN_CONV_SH]; P.sw = p[BN_CONV_SW];
  P.pt = p[BN_CONV_PAD_T]; P.pl = p[BN_CONV_PAD_L];
  P.in_zp = p[BN_CONV_IN_ZP]; P.out_zp = p[BN_CONV_OUT_ZP];
  P.act_min = p[BN_CONV_ACT_MIN]; P.act_max = p[BN_CONV_ACT_MAX];
  P.ih = ti.dims[0]; P.iw = ti.dims[1]; P.ic = p[BN_CONV_CIN];
  P.oh = to.dims[0]; P.ow = to.dims[1]; P.oc = p[BN_CONV_COUT];
  P.rounding = e->rounding;
}

// Runs ops [0, n_ops) on Bw chunks.  ptr[] = per-slot device pointers for this wave.
static int run_generic(bn_engine* e, std::vector<void*>& ptr, int Bw, cudaStream_t st) {
  const bn_blob_header* h = e->hdr;
  // first use of the one-kernel-per-op plan: weight images / folded constants of the ops that have a fast kernel (not needed,
  // and not paid for at bn_create, while the fused plan runs)
  if (!e->accel) e->accel = gen_accel_build(e->blob.data(), e->d_blob, e->hdr, e->tensors, e->ops);
  const int R = e->rounding;
  for (uint32_t oi = 0; oi < h->n_ops; oi++) {
    const bn_blob_op& op = e->ops[oi];
    const bn_blob_tensor& to = e->tensors[op.out];
    const bn_blob_tensor* ti = op.n_in > 0 ? &e->tensors[op.in[0]] : nullptr;
    const long n_out = t_elems(to) * Bw;
    void* y = ptr[op.out];
    const void* x = op.n_in > 0 ? ptr[op.in[0]] : nullptr;
    uint32_t skip = 0;
    if (e->prof.on) {
      char nm[40];
      snprintf(nm, sizeof nm, "G%02u_kind%d", oi, op.kind);
      e->prof.begin(nm, st);
    }
    switch (op.kind) {
      case BN_OP_QUANTIZE: launch_quantize((const float*)x, (int8_t*)y, n_out, op.f[0], op.p[0], st); break;
      case BN_OP_DEQUANTIZE: launch_dequantize((const int8_t*)x, (float*)y, n_out, op.f[0], op.p[0], st); break;
      case BN_OP_REQUANT: launch_requant((const int8_t*)x, (int8_t*)y, n_out, op.p[0], op.p[1], op.p[2], op.p[3], R, st); break;
      case BN_OP_TRANSPOSE: launch_transpose((const int8_t*)x, (int8_t*)y, n_out, ti->dims, to.dims, op.p, st); break;
      case BN_OP_SLICE: launch_slice((const int8_t*)x, (int8_t*)y, n_out, ti->dims, to.dims, op.p, st); break;
      case BN_OP_RESHAPE: {
        cudaError_t ce = cudaMemcpyAsync(y, x, (size_t)to.nbytes * Bw, cudaMemcpyDeviceToDevice, st);
        if (ce != cudaSuccess) return set_err(BN_ERR_CUDA, "reshape copy: %s", cudaGetErrorString(ce));
      } break;
      case BN_OP_FILL: launch_fill((int8_t*)y, n_out, op.p[0], st); break;
      case BN_OP_CONCAT: {
        const bn_blob_tensor& t1 = e->tensors[op.in[1]];
        int axis = op.p[0];
        long outer = 1, in0 = 1, in1 = 1;
        for (int d = 0; d < axis; d++) outer *= to.dims[d];
        for (int d = axis; d < 3; d++) { in0 *= ti->dims[d]; in1 *= t1.dims[d]; }
        launch_concat((const int8_t*)x, (const int8_t*)ptr[op.in[1]], (int8_t*)y, outer * Bw, (int)in0, (int)in1, st);
      } break;
      case BN_OP_CONV2D: {
        if (e->accel && e->accel->ops[oi].pw && R == 0 && e->fast.use_tc) {
          const GenAccelOp& g = e->accel->ops[oi];
          const long M = (long)to.dims[0] * to.dims[1] * Bw;
          int rc;
          if (g.add_fused && (e->fast.fusion & 1)) {       // convolution + the residual ADD behind it in one launch
            rc = launch_pw_tc((const int8_t*)x, (const int8_t*)ptr[g.add_res_slot], (int8_t*)ptr[g.add_out_slot], M, g.tc_add, e->fast.num_sms, st);
            skip = 1;
          } else {
            rc = launch_pw_tc((const int8_t*)x, nullptr, (int8_t*)y, M, g.tc, e->fast.num_sms, st);
          }
          if (rc) return rc;
        } else if (e->accel && e->accel->ops[oi].stem) {
          int rc = launch_stem((const int8_t*)x, (int8_t*)y, Bw, e->accel->ops[oi].stp, R, st);
          if (rc) return rc;
        } else if (e->accel && e->accel->ops[oi].pwc) {
          int rc = launch_pw((const int8_t*)x, nullptr, (int8_t*)y, (long)to.dims[0] * to.dims[1] * Bw, e->accel->ops[oi].pwp, R, st);
          if (rc) return rc;
        } else {
          ConvParams P; fill_conv(e, op, P); launch_conv2d((const int8_t*)x, (int8_t*)y, n_out, P, st);
        }
      } break;
      case BN_OP_DWCONV2D: {
        if (e->accel && e->accel->ops[oi].dsb && R == 0 && e->fast.use_tc && (e->fast.fusion & 1)) {   // the whole DS block in one launch
          const GenAccelOp& g = e->accel->ops[oi];
          int rc = launch_ds((const int8_t*)x, (int8_t*)ptr[g.ds_out_slot], Bw, g.ds, g.dsl, e->fast.num_sms, st);
          if (rc) return rc;
          skip = (uint32_t)g.ds_skip;
        } else if (e->accel && e->accel->ops[oi].dw) {
          int rc = launch_dw3x3((const int8_t*)x, (int8_t*)y, Bw, e->accel->ops[oi].dwp, R, st);
          if (rc) return rc;
        } else {
          ConvParams P; fill_conv(e, op, P); launch_dwconv2d((const int8_t*)x, (int8_t*)y, n_out, P, st);
        }
      } break;
      case BN_OP_FC: {
        ConvParams P; fill_conv(e, op, P);
        launch_fc((const int8_t*)x, (int8_t*)y, n_out, P, st);
      } break;
      case BN_OP_ADD: {
        AddParams P;
        const int32_t* p = op.p;
        P.in1_zp = p[BN_ADD_IN1_ZP]; P.in2_zp = p[BN_ADD_IN2_ZP]; P.out_zp = p[BN_ADD_OUT_ZP];
        P.left_shift = p[BN_ADD_LEFT_SHIFT];
        P.m1 = p[BN_ADD_M1]; P.s1 = p[BN_ADD_S1]; P.m2 = p[BN_ADD_M2]; P.s2 = p[BN_ADD_S2];
        P.mo = p[BN_ADD_MO]; P.so = p[BN_ADD_SO];
        P.act_min = p[BN_ADD_ACT_MIN]; P.act_max = p[BN_ADD_ACT_MAX];
        P.bcast = p[BN_ADD_BCAST]; P.C = to.dims[2]; P.rounding = R;
        launch_add((const int8_t*)x, (const int8_t*)ptr[op.in[1]], (int8_t*)y, n_out, t_elems(to), P, st);
      } break;
      case BN_OP_MUL:
        launch_mul((const int8_t*)x, (const int8_t*)ptr[op.in[1]], (int8_t*)y, n_out, t_elems(to), op.p, to.dims[2], R, st);
        break;
      case BN_OP_MEAN: {
        int variant = e->mean_variant ? e->mean_variant : (op.p[BN_MEAN_KEEP_DIMS] ? 3 : 2);
        if (e->accel && e->accel->ops[oi].se && (e->fast.fusion & 1)) {   // the whole SE gate (this op and the next three)
          int rc = launch_se_gate((const int8_t*)x, (int8_t*)y, (int8_t*)ptr[e->ops[oi + 1].out], (int8_t*)ptr[e->ops[oi + 2].out],
                                  (int8_t*)ptr[e->ops[oi + 3].out], Bw, e->accel->ops[oi].sep, variant, R, st);
          if (rc) return rc;
          skip = 3;
          break;
        }
        launch_mean((const int8_t*)x, (int8_t*)y, n_out, op.p[BN_MEAN_COUNT], ti->dims[2], op.p, ti->scale, to.scale, variant, R, st);
      } break;
      case BN_OP_LOGISTIC: launch_logistic((const int8_t*)x, (int8_t*)y, n_out, (const int8_t*)(e->d_blob + op.off[0]), st); break;
      case BN_OP_PAD: launch_pad((const int8_t*)x, (int8_t*)y, n_out, ti->dims, to.dims, op.p, st); break;
      case BN_OP_SOFTMAX:
        launch_softmax((const int8_t*)x, (int8_t*)y, n_out / to.dims[2], to.dims[2], (const float*)(e->d_blob + op.off[0]), op.f[1], op.p[0], st);
        break;
      case BN_OP_SUM: {
        const int axis = op.p[0];
        int outer = 1, inner = 1;
        for (int d = 0; d < axis; d++) outer *= ti->dims[d];
        for (int d = axis + 1; d < 3; d++) inner *= ti->dims[d];
        launch_sum((const int8_t*)x, (int8_t*)y, n_out, outer, ti->dims[axis], inner, op.f[0], op.f[1], op.p[3], st);
      } break;
      default: return set_err(BN_ERR_UNSUPPORTED, "op kind %d", op.kind);
    }
    if (e->prof.on) e->prof.end(st);
    e->launches++;
    oi += skip;                                        // ops covered by a fused launch
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(BN_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(ce));
  return 0;
}

// frontend for one wave: pcm (device) -> float32 graph input (device)
static int run_frontend(bn_engine* e, const void* d_pcm, int f32, const float* d_peak, int Bw, float* d_spec, cudaStream_t st) {
  const bn_blob_header* h = e->hdr;
  if (h->frontend_kind != BN_FE_HYBRID)
    return set_err(BN_ERR_UNSUPPORTED, "frontend kind %u has no CUDA kernel yet", h->frontend_kind);
  if (e->prof.on) e->prof.begin("K1_stft_binmajor", st);
  int rc = launch_stft_mag(d_pcm, f32, d_peak, d_spec, e->d_mnmx, Bw, (int)h->chunk_len, (int)h->n_fft, (int)h->hop, (int)h->spec_width, st);
  if (e->prof.on) e->prof.end(st);
  if (rc) return set_err(rc, "stft launch rejected (n_fft %u hop %u)", h->n_fft, h->hop);
  const long per = (long)(h->n_fft / 2 + 1) * h->spec_width;
  if (e->prof.on) e->prof.begin("K1b_normalize", st);
  launch_minmax_normalize(d_spec, per, per * Bw, e->d_mnmx, st);
  if (e->prof.on) e->prof.end(st);
  e->launches += 3;
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(BN_ERR_CUDA, "frontend launch failed: %s", cudaGetErrorString(ce));
  return 0;
}

// One wave of the whole path on device data.  d_pcm may be NULL (then d_spec is the input).
static int run_wave(bn_engine* e, const void* d_pcm, int f32, const float* d_peak, const float* d_spec_in, int Bw,
                    float* d_scores_out, cudaStream_t st) {
  const bn_blob_header* h = e->hdr;
  if (use_fast(e)) {
    int rc;
    Profiler* pr = e->prof.on ? &e->prof : nullptr;
    if (d_pcm) rc = fast_run_pcm(e->fast, d_pcm, f32, d_peak, Bw, d_scores_out, e->rounding, e->mean_variant, st, &e->launches, pr);
    else rc = fast_run_spec(e->fast, d_spec_in, Bw, d_scores_out, e->rounding, e->mean_variant, st, &e->launches, pr);
    if (rc) return set_err(rc, "fused plan failed: %s", cudaGetErrorString(cudaGetLastError()));
    e->last_wave_B = Bw;
    return 0;
  }
  std::vector<void*> ptr = e->buf;
  if (d_pcm) {
    int rc = run_frontend(e, d_pcm, f32, d_peak, Bw, (float*)ptr[h->input_tensor], st);
    if (rc) return rc;
  } else {
    ptr[h->input_tensor] = (void*)d_spec_in;
  }
  ptr[h->output_tensor] = d_scores_out;
  int rc = run_generic(e, ptr, Bw, st);
  e->last_ptr = ptr;
  e->last_wave_B = Bw;
  return rc;
}

// ---------------------------------------------------------------------------------------------
// public inference entry points
// ---------------------------------------------------------------------------------------------
static int check_engine(bn_engine* e) {
  if (!e) return set_err(BN_ERR_ARG, "NULL engine");
  cudaError_t ce = cudaSetDevice(e->device);
  if (ce != cudaSuccess) return set_err(BN_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce));
  return 0;
}

static int pick_wave(const bn_engine* e, int B) { return B < e->wave_opt ? B : e->wave_opt; }

// Shared implementation.  Exactly one of (pcm, spec) is non-NULL.  If offs != NULL pooled file
// scores are produced instead of chunk scores.
// `pcm` holds int16 samples (sbytes = 2) or float32 waveform samples (sbytes = 4).
static int infer_impl(bn_engine* e, const void* pcm_v, int sbytes, const float* peak, const float* spec, int B,
                      const int32_t* offs, int F, int pooling, float beta, float* out, cudaStream_t user_stream) {
  const char* pcm = (const char*)pcm_v;
  const int f32 = sbytes == 4;
  int rc = check_engine(e);
  if (rc) return rc;
  if (B < 0 || (!pcm && !spec) || !out) return set_err(BN_ERR_ARG, "bad arguments");
  const bn_blob_header* h = e->hdr;
  const int C = (int)h->num_classes;
  const long in_elems = t_elems(e->tensors[h->input_tensor]);
  const void* in_ptr = pcm ? (const void*)pcm : (const void*)spec;
  const bool dev_in = is_device_ptr(in_ptr);
  const bool dev_out = is_device_ptr(out);
  if (B > 0 && dev_in != dev_out) return set_err(BN_ERR_ARG, "input and output must both be host or both be device pointers");
  if (peak && B > 0 && is_device_ptr(peak) != dev_in) return set_err(BN_ERR_ARG, "peak must live where pcm lives");
  if (pcm && ((uintptr_t)pcm & 3)) return set_err(BN_ERR_ARG, "pcm must be 4-byte aligned");
  if (pcm && h->frontend_kind == BN_FE_NONE) return set_err(BN_ERR_UNSUPPORTED, "blob has no frontend description (export with the model config)");
  if (offs) {
    if (F < 0) return set_err(BN_ERR_ARG, "F < 0");
    if (pooling < BN_POOL_AVG || pooling > BN_POOL_LME) return set_err(BN_ERR_ARG, "Unsupported pooling method: %d", pooling);
  }
  if (B == 0 && !offs) return BN_OK;

  int wave = B > 0 ? pick_wave(e, B) : 1;
  if (!dev_in && wave > e->host_wave) wave = e->host_wave;
  rc = ensure_workspace(e, wave);
  if (rc) return rc;
  const bool pooled = offs != nullptr;
  // chunk scores live on the device when pooling or when the caller's buffers are host memory
  const bool scores_internal = pooled || !dev_out;
  rc = ensure_io(e, dev_in ? 0 : wave * (sbytes / 2), scores_internal ? (size_t)B * C + 1 : 0);
  if (rc) return rc;
  float* d_scores = scores_internal ? e->d_scores : out;

  cudaStream_t st = dev_in ? user_stream : e->s_comp;
  const int nw = B > 0 ? (B + wave - 1) / wave : 0;
  for (int w = 0; w < nw; w++) {
    const int b0 = w * wave;
    const int Bw = (B - b0) < wave ? (B - b0) : wave;
    const void* d_pcm = nullptr;
    const float* d_peak = nullptr;
    const float* d_spec = nullptr;
    if (dev_in) {
      if (pcm) { d_pcm = pcm + (size_t)b0 * h->chunk_len * sbytes; d_peak = peak ? peak + b0 : nullptr; }
      else d_spec = spec + (long)b0 * in_elems;
    } else {
      const int slot = w & 1;
      // the compute of wave w-2 must have consumed this slot before it is overwritten
      CU(cudaStreamWaitEvent(e->s_copy, e->ev_comp[slot], 0));
      if (pcm) {
        CU(cudaMemcpyAsync(e->d_pcm[slot], pcm + (size_t)b0 * h->chunk_len * sbytes, (size_t)sbytes * h->chunk_len * Bw, cudaMemcpyHostToDevice, e->s_copy));
        if (peak) CU(cudaMemcpyAsync(e->d_peak[slot], peak + b0, sizeof(float) * Bw, cudaMemcpyHostToDevice, e->s_copy));
        d_pcm = e->d_pcm[slot];
        d_peak = peak ? e->d_peak[slot] : nullptr;
      } else {
        // spectrogram input goes straight into the graph-input buffer (single slot: serialise)
        CU(cudaStreamWaitEvent(e->s_copy, e->ev_comp[slot ^ 1], 0));
        CU(cudaMemcpyAsync(e->buf[h->input_tensor], spec + (long)b0 * in_elems, sizeof(float) * (size_t)in_elems * Bw, cudaMemcpyHostToDevice, e->s_copy));
        d_spec = (const float*)e->buf[h->input_tensor];
      }
      CU(cudaEventRecord(e->ev_h2d[slot], e->s_copy));
      CU(cudaStreamWaitEvent(st, e->ev_h2d[slot], 0));
    }
    rc = run_wave(e, d_pcm, f32, d_peak, d_spec, Bw, d_scores + (long)b0 * C, st);
    if (rc) return rc;
    if (!dev_in) CU(cudaEventRecord(e->ev_comp[w & 1], st));
  }

  if (pooled) {
    // file offsets -> device, pool, return [F, C]
    if (e->d_offs_cap < (size_t)F + 1) {
      if (e->d_offs) cudaFree(e->d_offs);
      CU(cudaMalloc(&e->d_offs, sizeof(int) * ((size_t)F + 1)));
      e->d_offs_cap = (size_t)F + 1;
    }
    const bool dev_offs = is_device_ptr(offs);
    const int* d_offs = offs;
    if (!dev_offs) {
      CU(cudaMemcpyAsync(e->d_offs, offs, sizeof(int) * ((size_t)F + 1), cudaMemcpyHostToDevice, st));
      d_offs = e->d_offs;
    }
    float* d_fs = out;
    if (!dev_out) {
      if (e->d_file_scores_cap < (size_t)F * C + 1) {
        if (e->d_file_scores) cudaFree(e->d_file_scores);
        CU(cudaMalloc(&e->d_file_scores, sizeof(float) * ((size_t)F * C + 1)));
        e->d_file_scores_cap = (size_t)F * C + 1;
      }
      d_fs = e->d_file_scores;
    }
    if (F > 0) {
      launch_pool(d_scores, d_offs, d_fs, F, C, pooling, beta, st);
      e->launches++;
      if (!dev_out) CU(cudaMemcpyAsync(out, d_fs, sizeof(float) * (size_t)F * C, cudaMemcpyDeviceToHost, st));
    }
  } else if (!dev_out) {
    CU(cudaMemcpyAsync(out, d_scores, sizeof(float) * (size_t)B * C, cudaMemcpyDeviceToHost, st));
  }
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return set_err(BN_ERR_CUDA, "launch failed: %s", cudaGetErrorString(ce));
  if (!dev_in) CU(cudaStreamSynchronize(st));
  return BN_OK;
}

extern "C" int bn_infer_spec_f32(bn_engine* e, const float* spec, int B, float* scores, void* stream) {
  return infer_impl(e, nullptr, 2, nullptr, spec, B, nullptr, 0, 0, 0.f, scores, (cudaStream_t)stream);
}

extern "C" int bn_infer_pcm16(bn_engine* e, const int16_t* pcm, const float* peak, int B, float* scores, void* stream) {
  if (!pcm) return set_err(BN_ERR_ARG, "pcm is NULL");
  return infer_impl(e, pcm, 2, peak, nullptr, B, nullptr, 0, 0, 0.f, scores, (cudaStream_t)stream);
}

extern "C" int bn_infer_wave_f32(bn_engine* e, const float* wave, const float* peak, int B, float* scores, void* stream) {
  if (!wave) return set_err(BN_ERR_ARG, "wave is NULL");
  return infer_impl(e, wave, 4, peak, nullptr, B, nullptr, 0, 0, 0.f, scores, (cudaStream_t)stream);
}

static int infer_pool_any(bn_engine* e, const void* pcm, int sbytes, const float* peak, const int32_t* file_offsets, int F,
                          int pooling, float beta, float* file_scores, void* stream) {
  if (!pcm && F > 0) return set_err(BN_ERR_ARG, "pcm is NULL");
  if (!file_offsets) return set_err(BN_ERR_ARG, "file_offsets is NULL");
  int B = 0;
  if (is_device_ptr(file_offsets)) {
    if (cudaMemcpy(&B, file_offsets + F, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
      return set_err(BN_ERR_CUDA, "reading file_offsets[F] failed");
  } else {
    B = file_offsets[F];
    for (int f = 0; f < F; f++)
      if (file_offsets[f] > file_offsets[f + 1] || file_offsets[f] < 0) return set_err(BN_ERR_ARG, "file_offsets must be non-decreasing");
  }
  static const float dummy[1] = {0.f};
  return infer_impl(e, pcm ? pcm : (const void*)dummy, sbytes, peak, nullptr, B, file_offsets, F, pooling, beta, file_scores, (cudaStream_t)stream);
}

extern "C" int bn_infer_pool(bn_engine* e, const int16_t* pcm, const float* peak, const int32_t* file_offsets, int F,
                             int pooling, float beta, float* file_scores, void* stream) {
  return infer_pool_any(e, pcm, 2, peak, file_offsets, F, pooling, beta, file_scores, stream);
}

extern "C" int bn_infer_pool_wave_f32(bn_engine* e, const float* wave, const float* peak, const int32_t* file_offsets, int F,
                                      int pooling, float beta, float* file_scores, void* stream) {
  return infer_pool_any(e, wave, 4, peak, file_offsets, F, pooling, beta, file_scores, stream);
}

static int frontend_any(bn_engine* e, const void* pcm_v, int sbytes, const float* peak, int B, float* spec_out, void* stream) {
  const char* pcm = (const char*)pcm_v;
  int rc = check_engine(e);
  if (rc) return rc;
  if (!pcm || !spec_out || B < 0) return set_err(BN_ERR_ARG, "bad arguments");
  if ((uintptr_t)pcm & 3) return set_err(BN_ERR_ARG, "pcm must be 4-byte aligned");
  if (B == 0) return BN_OK;
  const bn_blob_header* h = e->hdr;
  const long in_elems = t_elems(e->tensors[h->input_tensor]);
  const bool dev_in = is_device_ptr(pcm), dev_out = is_device_ptr(spec_out);
  if (dev_in != dev_out) return set_err(BN_ERR_ARG, "input and output must both be host or both be device pointers");
  const int wave = pick_wave(e, B);
  rc = ensure_workspace(e, wave);
  if (rc) return rc;
  rc = ensure_io(e, dev_in ? 0 : wave * (sbytes / 2), 0);
  if (rc) return rc;
  cudaStream_t st = dev_in ? (cudaStream_t)stream : e->s_comp;
  for (int b0 = 0; b0 < B; b0 += wave) {
    const int Bw = (B - b0) < wave ? (B - b0) : wave;
    const void* d_pcm = pcm + (size_t)b0 * h->chunk_len * sbytes;
    const float* d_peak = peak ? peak + b0 : nullptr;
    float* d_spec = spec_out + (long)b0 * in_elems;
    if (!dev_in) {
      CU(cudaMemcpyAsync(e->d_pcm[0], d_pcm, (size_t)sbytes * h->chunk_len * Bw, cudaMemcpyHostToDevice, st));
      if (peak) CU(cudaMemcpyAsync(e->d_peak[0], d_peak, sizeof(float) * Bw, cudaMemcpyHostToDevice, st));
      d_pcm = e->d_pcm[0];
      d_peak = peak ? e->d_peak[0] : nullptr;
      d_spec = (float*)e->buf[h->input_tensor];
    }
    rc = run_frontend(e, d_pcm, sbytes == 4, d_peak, Bw, d_spec, st);
    if (rc) return rc;
    if (!dev_in) CU(cudaMemcpyAsync(spec_out + (long)b0 * in_elems, d_spec, sizeof(float) * (size_t)in_elems * Bw, cudaMemcpyDeviceToHost, st));
  }
  if (!dev_in) CU(cudaStreamSynchronize(st));
  return BN_OK;
}

extern "C" int bn_frontend_pcm16(bn_engine* e, const int16_t* pcm, const float* peak, int B, float* spec_out, void* stream) {
  return frontend_any(e, pcm, 2, peak, B, spec_out, stream);
}

extern "C" int bn_frontend_wave_f32(bn_engine* e, const float* wave, const float* peak, int B, float* spec_out, void* stream) {
  return frontend_any(e, wave, 4, peak, B, spec_out, stream);
}

extern "C" int bn_pool_scores(bn_engine* e, const float* chunk_scores, const int32_t* file_offsets, int F, int C,
                              int pooling, float beta, float* file_scores, void* stream) {
  int rc = check_engine(e);
  if (rc) return rc;
  if (!file_offsets || !file_scores || F < 0 || C <= 0) return set_err(BN_ERR_ARG, "bad arguments");
  if (pooling < BN_POOL_AVG || pooling > BN_POOL_LME) return set_err(BN_ERR_ARG, "Unsupported pooling method: %d", pooling);
  if (F == 0) return BN_OK;
  const bool dev = is_device_ptr(file_scores);
  if (dev) {
    if (!is_device_ptr(file_offsets) || (chunk_scores && !is_device_ptr(chunk_scores)))
      return set_err(BN_ERR_ARG, "all pointers must be device pointers");
    launch_pool(chunk_scores, file_offsets, file_scores, F, C, pooling, beta, (cudaStream_t)stream);
    e->launches++;
    return BN_OK;
  }
  const int N = file_offsets[F];
  float* d_s = nullptr; int* d_o = nullptr; float* d_f = nullptr;
  CU(cudaMalloc(&d_s, sizeof(float) * ((size_t)N * C + 1)));
  CU(cudaMalloc(&d_o, sizeof(int) * (F + 1)));
  CU(cudaMalloc(&d_f, sizeof(float) * (size_t)F * C));
  if (N > 0) CU(cudaMemcpy(d_s, chunk_scores, sizeof(float) * (size_t)N * C, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_o, file_offsets, sizeof(int) * (F + 1), cudaMemcpyHostToDevice));
  launch_pool(d_s, d_o, d_f, F, C, pooling, beta, e->s_comp);
  e->launches++;
  CU(cudaMemcpyAsync(file_scores, d_f, sizeof(float) * (size_t)F * C, cudaMemcpyDeviceToHost, e->s_comp));
  CU(cudaStreamSynchronize(e->s_comp));
  cudaFree(d_s); cudaFree(d_o); cudaFree(d_f);
  return BN_OK;
}

extern "C" int bn_dump_tensor(bn_engine* e, int tfl_tensor_id, void* out, size_t nbytes) {
  int rc = check_engine(e);
  if (rc) return rc;
  if (!out) return set_err(BN_ERR_ARG, "out is NULL");
  if (e->last_wave_B <= 0) return set_err(BN_ERR_STATE, "no inference has run yet");
  CU(cudaDeviceSynchronize());
  if (use_fast(e)) {
    int r = fast_dump_tensor(e->fast, tfl_tensor_id, e->last_wave_B, out, nbytes);
    if (r) return set_err(r, "tensor %d is not materialised by the fused plan (set BN_OPT_FORCE_GENERIC)", tfl_tensor_id);
    return BN_OK;
  }
  for (uint32_t i = 0; i < e->hdr->n_tensors; i++) {
    const bn_blob_tensor& t = e->tensors[i];
    if (t.id != tfl_tensor_id) continue;
    size_t want = t.is_const ? (size_t)t.nbytes : (size_t)t.nbytes * e->last_wave_B;
    if (nbytes != want) return set_err(BN_ERR_ARG, "tensor %d: expected %zu bytes, got %zu", tfl_tensor_id, want, nbytes);
    if (!e->last_ptr[i]) return set_err(BN_ERR_STATE, "tensor %d has no buffer", tfl_tensor_id);
    CU(cudaMemcpy(out, e->last_ptr[i], want, cudaMemcpyDeviceToHost));
    return BN_OK;
  }
  return set_err(BN_ERR_ARG, "no tensor with id %d", tfl_tensor_id);
}

extern "C" int bn_query(const bn_engine* e, bn_info* out) {
  if (!e || !out) return set_err(BN_ERR_ARG, "NULL argument");
  memset(out, 0, sizeof *out);
  const bn_blob_header* h = e->hdr;
  out->frontend_kind = (int32_t)h->frontend_kind;
  out->sample_rate = (int32_t)h->sample_rate;
  out->chunk_len = (int32_t)h->chunk_len;
  out->n_fft = (int32_t)h->n_fft;
  out->hop = (int32_t)h->hop;
  out->spec_width = (int32_t)h->spec_width;
  out->fft_bins = (int32_t)(h->n_fft / 2 + 1);
  out->num_classes = (int32_t)h->num_classes;
  out->input_elems = t_elems(e->tensors[h->input_tensor]);
  out->n_ops = (int32_t)h->n_ops;
  out->n_tensors = (int32_t)h->n_tensors;
  out->device = e->device;
  out->wave = e->wave ? e->wave : e->wave_opt;
  out->workspace_bytes = (int64_t)e->workspace_bytes;
  out->fast_path = use_fast(e) ? 1 : 0;
  return BN_OK;
}

extern "C" int bn_set_option(bn_engine* e, int key, int value) {
  int rc = check_engine(e);
  if (rc) return rc;
  switch (key) {
    case BN_OPT_ROUNDING:
      if (value != 0 && value != 1) return set_err(BN_ERR_ARG, "rounding must be 0 or 1");
      e->rounding = value; break;
    case BN_OPT_MEAN_VARIANT:
      if (value < 0 || value > 3) return set_err(BN_ERR_ARG, "mean variant must be 0..3");
      e->mean_variant = value; break;
    case BN_OPT_FORCE_GENERIC:
      if ((value != 0) != (e->force_generic != 0)) { cudaDeviceSynchronize(); free_workspace(e); }
      e->force_generic = value ? 1 : 0; break;
    case BN_OPT_WAVE:
      if (value < 1 || value > 65535) return set_err(BN_ERR_ARG, "wave must be in [1, 65535]");
      if (value != e->wave_opt) { cudaDeviceSynchronize(); free_workspace(e); }
      e->wave_opt = value; break;
    case BN_OPT_HOST_WAVE:
      if (value < 1) return set_err(BN_ERR_ARG, "host wave must be >= 1");
      e->host_wave = value; break;
    case BN_OPT_TENSOR_CORE:
      e->fast.use_tc = value ? 1 : 0; break;
    case BN_OPT_FUSION:
      e->fast.fusion = value & 511; break;
    case BN_OPT_PROFILE:
      cudaDeviceSynchronize();
      e->prof.collect();
      e->prof.on = value != 0;
      if (value == 2) e->prof.reset();
      break;
    default: return set_err(BN_ERR_ARG, "unknown option %d", key);
  }
  return BN_OK;
}

extern "C" int bn_profile_read(bn_engine* e, int index, char* name, size_t name_cap, double* ms, int64_t* count) {
  int rc = check_engine(e);
  if (rc) return rc;
  cudaDeviceSynchronize();
  e->prof.collect();
  if (index < 0 || (size_t)index >= e->prof.names.size()) return BN_ERR_ARG;
  if (name && name_cap) { strncpy(name, e->prof.names[index].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ms) *ms = e->prof.ms[index];
  if (count) *count = e->prof.count[index];
  return BN_OK;
}

extern "C" int64_t bn_launch_count(const bn_engine* e) { return e ? e->launches : 0; }

// NUMA placement of pinned buffers.  On a two-socket host the pages of a pinned buffer land on the socket of the thread that
// calls cudaMallocHost; when that is not the socket the GPU hangs off, every upload crosses the inter-socket link.  So the
// calling thread is bound to the CPUs of the current device's NUMA node for the duration of the allocation.  (On the round-1
// measurement boxes this is a no-op: they are VMs that expose one NUMA node and numa_node = -1 for every PCI device; their
// aggregate host-to-device rate saturates at about 115 GB/s with 4 and 186 GB/s with 8 GPUs uploading at once, which is what
// bounds the multi-GPU end-to-end numbers in profiles/r1/bench_N4 / bench_N8.)
static int device_numa_node(int dev) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* c = bus; *c; c++) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

static bool numa_node_cpus(int node, cpu_set_t* set) {
  char path[128];
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char buf[4096] = {0};
  const bool ok = fgets(buf, sizeof buf, f) != nullptr;
  fclose(f);
  if (!ok) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int got = sscanf(tok, "%d-%d", &a, &b);
    if (got == 1) b = a;
    if (got < 1) continue;
    for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); n++; }
  }
  return n > 0;
}

extern "C" void* bn_host_alloc(size_t nbytes) {
  void* p = nullptr;
  cpu_set_t old_set, want;
  bool bound = false;
  if (!getenv("BN_NO_NUMA_BIND")) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
      const int node = device_numa_node(dev);
      if (node >= 0 && numa_node_cpus(node, &want) && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
        cpu_set_t both;
        CPU_AND(&both, &want, &old_set);                // stay inside the CPUs this process is allowed to use
        if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof both, &both) == 0) bound = true;
      }
    } else {
      cudaGetLastError();
    }
  }
  const cudaError_t ce = cudaMallocHost(&p, nbytes);
  if (ce == cudaSuccess && bound && nbytes) memset(p, 0, nbytes < (1u << 20) ? nbytes : (1u << 20));   // harmless; pages are already resident
  if (bound) sched_setaffinity(0, sizeof old_set, &old_set);
  if (ce != cudaSuccess) { cudaGetLastError(); set_err(BN_ERR_CUDA, "cudaMallocHost(%zu) failed", nbytes); return nullptr; }
  return p;
}
extern "C" void bn_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" const char* bn_last_error(void) { return g_err; }
extern "C" const char* bn_version(void) { return "birdnet-b200 0.1 (sm_100a)"; }
