// bn_metrics.cu -- ROC-AUC / average precision / cmAP / precision-recall-F1 of an [F, C] score matrix on the GPU.
//
// Reference: the metric tail of evaluate(), birdnet_stm32/evaluation/metrics.py:152-190, which calls scikit-learn
// (roc_auc_score micro, average_precision_score per class and micro).  sklearn sorts every column (and the raveled
// matrix) on one host core; here:
//
//   k_transpose      [F, C] -> class-major [C, F] keys (scores) and values (labels)
//   CUB segmented radix sort (descending) of the C columns, CUB radix sort of the raveled matrix   (library primitives:
//                    the sort is not part of the classification hot path; everything around it is written here)
//   CUB inclusive sum of the sorted labels = running true-positive count (global; per-segment by subtracting the base)
//   k_mark_ends      threshold ends (last element of a run of equal scores, or of a segment), packed (index + 1, tp)
//   CUB exclusive max-scan of the packed ends = "previous threshold" for every position (both fields are monotone)
//   k_accumulate     per end: AP term (tp - tp_prev) / P * tp / (tp + fp), AUC term (fp - fp_prev) / N * (tp + tp_prev) / 2P,
//                    float64, warp-reduced, one atomicAdd per warp and segment
//   k_prf            tp / fp / fn at threshold 0.5
#include "../../include/bn_metrics.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include <vector>

#include "bn_common.cuh"
#include "bn_kernels.cuh"

namespace bn {

constexpr int MT_THREADS = 256;

__global__ void __launch_bounds__(MT_THREADS)
k_transpose(const float* __restrict__ y_true, const float* __restrict__ y_score, float* __restrict__ keys, int* __restrict__ vals,
            int F, int C) {
  __shared__ float ts[32][33], tt[32][33];
  const int f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r, c = c0 + tx;
    if (f < F && c < C) { ts[r][tx] = y_score[(size_t)f * C + c]; tt[r][tx] = y_true[(size_t)f * C + c]; }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, f = f0 + tx;
    if (f < F && c < C) { keys[(size_t)c * F + f] = ts[tx][r]; vals[(size_t)c * F + f] = tt[tx][r] != 0.0f ? 1 : 0; }
  }
}

__global__ void __launch_bounds__(MT_THREADS)
k_flatten(const float* __restrict__ y_true, int* __restrict__ vals, long n) {
  for (long i = blockIdx.x * (long)MT_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * MT_THREADS) vals[i] = y_true[i] != 0.0f ? 1 : 0;
}

// segments all have length L; packed = (k + 1) << 32 | tp_global(k) at threshold ends, 0 elsewhere
__global__ void __launch_bounds__(MT_THREADS)
k_mark_ends(const float* __restrict__ keys, const int* __restrict__ tpg, unsigned long long* __restrict__ packed, long n, long L) {
  for (long k = blockIdx.x * (long)MT_THREADS + threadIdx.x; k < n; k += (long)gridDim.x * MT_THREADS) {
    const bool end = ((k + 1) % L == 0) || keys[k] != keys[k + 1];
    packed[k] = end ? (((unsigned long long)(k + 1) << 32) | (unsigned)tpg[k]) : 0ull;
  }
}

struct MaxU64 {
  __host__ __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; }
};

// out[2 * s] += AP terms of segment s, out[2 * s + 1] += AUC terms
__global__ void __launch_bounds__(MT_THREADS)
k_accumulate(const float* __restrict__ keys, const int* __restrict__ tpg, const unsigned long long* __restrict__ prev,
             double* __restrict__ out, long n, long L) {
  const long stride = (long)gridDim.x * MT_THREADS;
  for (long k0 = blockIdx.x * (long)MT_THREADS; k0 < n; k0 += stride) {
    const long k = k0 + threadIdx.x;
    double ap = 0.0, auc = 0.0;
    long seg = -1;
    if (k < n) {
      seg = k / L;
      const bool end = ((k + 1) % L == 0) || keys[k] != keys[k + 1];
      if (end) {
        const long s0 = seg * L;
        const long base = s0 > 0 ? tpg[s0 - 1] : 0;
        const long P = tpg[s0 + L - 1] - base, N = L - P;
        const unsigned long long pv = prev[k];
        const long pidx = (long)(pv >> 32), ptpg = (long)(pv & 0xffffffffull);
        const long tp = tpg[k] - base, cnt = k - s0 + 1, fp = cnt - tp;
        const long ptp = ptpg - base, pcnt = pidx - s0, pfp = pcnt - ptp;
        if (P > 0) ap = ((double)tp / (double)P - (double)ptp / (double)P) * ((double)tp / (double)(tp + fp));
        if (P > 0 && N > 0) auc = ((double)fp / (double)N - (double)pfp / (double)N) * ((double)tp / (double)P + (double)ptp / (double)P) * 0.5;
      }
    }
    // warp reduction when the whole warp sits in one segment, per-lane atomics across a boundary
    const long seg0 = __shfl_sync(0xffffffffu, seg, 0);
    const bool uniform = __all_sync(0xffffffffu, seg == seg0 || seg < 0);
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { ap += __shfl_xor_sync(0xffffffffu, ap, o); auc += __shfl_xor_sync(0xffffffffu, auc, o); }
      if ((threadIdx.x & 31) == 0 && seg0 >= 0 && (ap != 0.0 || auc != 0.0)) { atomicAdd(out + 2 * seg0, ap); atomicAdd(out + 2 * seg0 + 1, auc); }
    } else if (seg >= 0 && (ap != 0.0 || auc != 0.0)) {
      atomicAdd(out + 2 * seg, ap);
      atomicAdd(out + 2 * seg + 1, auc);
    }
  }
}

// counts[0..2] = tp, fp, fn at threshold 0.5 ; counts[3] = positives
__global__ void __launch_bounds__(MT_THREADS)
k_prf(const float* __restrict__ y_true, const float* __restrict__ y_score, unsigned long long* __restrict__ counts, long n) {
  unsigned tp = 0, fp = 0, fn = 0, pos = 0;
  for (long i = blockIdx.x * (long)MT_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * MT_THREADS) {
    const bool t = y_true[i] != 0.0f, h = y_score[i] >= 0.5f;
    tp += t && h; fp += !t && h; fn += t && !h; pos += t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
    fn += __shfl_xor_sync(0xffffffffu, fn, o); pos += __shfl_xor_sync(0xffffffffu, pos, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counts + 0, (unsigned long long)tp); atomicAdd(counts + 1, (unsigned long long)fp);
    atomicAdd(counts + 2, (unsigned long long)fn); atomicAdd(counts + 3, (unsigned long long)pos);
  }
}

}  // namespace bn

using namespace bn;

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t n) { return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? 0 : -1; }
  template <typename T> T* as() { return (T*)p; }
};
bool mt_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
}  // namespace

#define MT_CU(call)                                                                                  \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      char _m[256];                                                                                  \
      snprintf(_m, sizeof _m, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return set_error(BN_ERR_CUDA, _m);                                                             \
    }                                                                                                \
  } while (0)

// sorted-run statistics of `S` segments of length L held in keys / vals (unsorted on entry); sums[2 s] = AP, sums[2 s + 1] = AUC
static int segment_stats(float* keys_in, int* vals_in, float* keys, int* vals, int* tpg, unsigned long long* packed,
                         unsigned long long* prev, double* d_sums, long L, int S, int grid, int* launches, cudaStream_t st) {
  const long n = L * S;
  if (n >= (1L << 31)) return set_error(BN_ERR_ARG, "bn_metrics_compute: more than 2^31 cells");
  size_t t1 = 0, t2 = 0, t3 = 0;
  const int nn = (int)n;
  if (S == 1) {
    MT_CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, t1, keys_in, keys, vals_in, vals, nn, 0, 32, st));
  }
  std::vector<int> h_offs;
  DevBuf d_offs;
  if (S > 1) {
    h_offs.resize(S + 1);
    for (int s = 0; s <= S; s++) h_offs[s] = (int)(s * L);
    if (d_offs.alloc(sizeof(int) * (S + 1))) return set_error(BN_ERR_CUDA, "cudaMalloc failed (metrics offsets)");
    MT_CU(cudaMemcpyAsync(d_offs.p, h_offs.data(), sizeof(int) * (S + 1), cudaMemcpyHostToDevice, st));
    MT_CU(cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, t1, keys_in, keys, vals_in, vals, nn, S, d_offs.as<int>(), d_offs.as<int>() + 1, 0, 32, st));
  }
  MT_CU(cub::DeviceScan::InclusiveSum(nullptr, t2, vals, tpg, nn, st));
  MT_CU(cub::DeviceScan::ExclusiveScan(nullptr, t3, packed, prev, MaxU64(), 0ull, nn, st));
  size_t tmax = t1 > t2 ? t1 : t2;
  if (t3 > tmax) tmax = t3;
  DevBuf temp;
  if (temp.alloc(tmax)) return set_error(BN_ERR_CUDA, "cudaMalloc failed (metrics sort workspace)");
  if (S == 1) MT_CU(cub::DeviceRadixSort::SortPairsDescending(temp.p, t1, keys_in, keys, vals_in, vals, nn, 0, 32, st));
  else MT_CU(cub::DeviceSegmentedRadixSort::SortPairsDescending(temp.p, t1, keys_in, keys, vals_in, vals, nn, S, d_offs.as<int>(), d_offs.as<int>() + 1, 0, 32, st));
  MT_CU(cub::DeviceScan::InclusiveSum(temp.p, t2, vals, tpg, nn, st));
  k_mark_ends<<<grid, MT_THREADS, 0, st>>>(keys, tpg, packed, n, L);
  MT_CU(cub::DeviceScan::ExclusiveScan(temp.p, t3, packed, prev, MaxU64(), 0ull, nn, st));
  MT_CU(cudaMemsetAsync(d_sums, 0, sizeof(double) * 2 * S, st));
  k_accumulate<<<grid, MT_THREADS, 0, st>>>(keys, tpg, prev, d_sums, n, L);
  *launches += 5;
  MT_CU(cudaStreamSynchronize(st));                   // temp / d_offs go out of scope
  return 0;
}

extern "C" int bn_metrics_compute(const float* y_true, const float* y_score, int F, int C, int device, bn_metrics_result* out,
                                  double* ap_per_class) {
  if (!y_true || !y_score || !out || F <= 0 || C <= 0) return set_error(BN_ERR_ARG, "bn_metrics_compute: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return set_error(BN_ERR_CUDA, "no CUDA device for the metric kernels (there is no CPU fallback)");
  }
  MT_CU(cudaSetDevice(device));
  const long n = (long)F * C;
  const bool dev_in = mt_is_device_ptr(y_score);
  if (dev_in != mt_is_device_ptr(y_true)) return set_error(BN_ERR_ARG, "y_true and y_score must both be host or both be device pointers");
  cudaStream_t st = nullptr;
  DevBuf b_true, b_score, b_keys_in, b_vals_in, b_keys, b_vals, b_tpg, b_packed, b_prev, b_sums, b_counts;
  const float* d_true = y_true;
  const float* d_score = y_score;
  if (!dev_in) {
    if (b_true.alloc(sizeof(float) * n) || b_score.alloc(sizeof(float) * n)) return set_error(BN_ERR_CUDA, "cudaMalloc failed (metrics inputs)");
    MT_CU(cudaMemcpyAsync(b_true.p, y_true, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    MT_CU(cudaMemcpyAsync(b_score.p, y_score, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    d_true = b_true.as<float>();
    d_score = b_score.as<float>();
  }
  if (b_keys_in.alloc(sizeof(float) * n) || b_vals_in.alloc(sizeof(int) * n) || b_keys.alloc(sizeof(float) * (n + 1)) ||
      b_vals.alloc(sizeof(int) * n) || b_tpg.alloc(sizeof(int) * n) || b_packed.alloc(8 * n) || b_prev.alloc(8 * n) ||
      b_sums.alloc(sizeof(double) * 2 * (C > 1 ? C : 1)) || b_counts.alloc(8 * 4))
    return set_error(BN_ERR_CUDA, "cudaMalloc failed (metrics workspace)");
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  long gb = (n + MT_THREADS - 1) / MT_THREADS;
  const int grid = (int)(gb < (long)sms * 8 ? gb : (long)sms * 8);
  int launches = 0;
  memset(out, 0, sizeof *out);

  // ---- precision / recall / F1 at 0.5 -------------------------------------------------------------------
  MT_CU(cudaMemsetAsync(b_counts.p, 0, 32, st));
  k_prf<<<grid, MT_THREADS, 0, st>>>(d_true, d_score, b_counts.as<unsigned long long>(), n);
  launches++;
  unsigned long long cnt[4];
  MT_CU(cudaMemcpyAsync(cnt, b_counts.p, 32, cudaMemcpyDeviceToHost, st));
  MT_CU(cudaStreamSynchronize(st));
  {
    // the reference computes these in float32: np.sum of float32 indicator products (exact integers below 2^24), then
    // tp / (tp + fp + 1e-12) with the Python float as a weak scalar (evaluation/metrics.py:165-174)
    const float tp = (float)cnt[0], fp = (float)cnt[1], fn = (float)cnt[2];
    const float precision = tp / (tp + fp + 1e-12f), recall = tp / (tp + fn + 1e-12f);
    out->precision = (double)precision;
    out->recall = (double)recall;
    out->f1 = precision + recall > 0 ? (double)(2.0f * (precision * recall) / (precision + recall)) : 0.0;
    out->n_positive = (int64_t)cnt[3];
    out->n_cells = n;
  }

  // ---- per-class AP -> cmAP -----------------------------------------------------------------------------
  {
    dim3 tg((F + 31) / 32, (C + 31) / 32);
    k_transpose<<<tg, MT_THREADS, 0, st>>>(d_true, d_score, b_keys_in.as<float>(), b_vals_in.as<int>(), F, C);
    launches++;
    int rc = segment_stats(b_keys_in.as<float>(), b_vals_in.as<int>(), b_keys.as<float>(), b_vals.as<int>(), b_tpg.as<int>(),
                           b_packed.as<unsigned long long>(), b_prev.as<unsigned long long>(), b_sums.as<double>(), F, C, grid, &launches, st);
    if (rc) return rc;
    std::vector<double> sums(2 * (size_t)C);
    std::vector<int> tpg_end(C);
    MT_CU(cudaMemcpy(sums.data(), b_sums.p, sizeof(double) * 2 * C, cudaMemcpyDeviceToHost));
    // positives per class = tp count at each segment end
    MT_CU(cudaMemcpy2D(tpg_end.data(), sizeof(int), b_tpg.as<int>() + (F - 1), sizeof(int) * (size_t)F, sizeof(int), C, cudaMemcpyDeviceToHost));
    double acc = 0.0;
    int none = 0, prev_end = 0;
    for (int c = 0; c < C; c++) {
      const int P = tpg_end[c] - prev_end;
      prev_end = tpg_end[c];
      const double ap = P > 0 ? sums[2 * c] : 0.0;
      if (P == 0) none++;
      if (ap_per_class) ap_per_class[c] = ap;
      acc += ap;
    }
    out->cmap = acc / C;
    out->classes_without_positives = none;
  }

  // ---- micro AP and micro ROC-AUC over the raveled matrix -------------------------------------------------
  {
    MT_CU(cudaMemcpyAsync(b_keys_in.p, d_score, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    k_flatten<<<grid, MT_THREADS, 0, st>>>(d_true, b_vals_in.as<int>(), n);
    launches++;
    int rc = segment_stats(b_keys_in.as<float>(), b_vals_in.as<int>(), b_keys.as<float>(), b_vals.as<int>(), b_tpg.as<int>(),
                           b_packed.as<unsigned long long>(), b_prev.as<unsigned long long>(), b_sums.as<double>(), n, 1, grid, &launches, st);
    if (rc) return rc;
    double sums[2];
    MT_CU(cudaMemcpy(sums, b_sums.p, sizeof sums, cudaMemcpyDeviceToHost));
    const long P = out->n_positive, N = n - P;
    out->map_micro = P > 0 ? sums[0] : 0.0;
    out->roc_auc_micro = (P > 0 && N > 0) ? sums[1] : NAN;
  }
  out->n_launches = launches;
  return BN_OK;
}

// =================================================================================================================
// Bootstrap confidence intervals of the per-class average precision (reference evaluation/metrics.py:240-322)
// =================================================================================================================
namespace bn {

// Philox-4x32-10 (counter-based: resample r, draw i are the counter, the seed is the key)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int round = 0; round < 10; round++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

// mult[r][i] = how often file i is drawn in resample r0 + r (F draws with replacement)
__global__ void __launch_bounds__(MT_THREADS)
k_boot_draw(int* __restrict__ mult, int F, int R, int r0, unsigned long long seed) {
  const long n4 = ((long)F + 3) / 4;
  for (long t = blockIdx.x * (long)MT_THREADS + threadIdx.x; t < n4 * R; t += (long)gridDim.x * MT_THREADS) {
    const int r = (int)(t / n4);
    const long q = t - (long)r * n4;
    const uint4 u = philox4x32(make_uint4((unsigned)q, (unsigned)(q >> 32), (unsigned)(r0 + r), 0x424e4231u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const unsigned v[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (4 * q + j < F) atomicAdd(mult + (size_t)r * F + __umulhi(v[j], (unsigned)F), 1);
  }
}

// class-major sorted order: lab[c][p] = label of the file at sorted position p, end[c][p] = last position of a run of equal scores
__global__ void __launch_bounds__(MT_THREADS)
k_boot_prepare(const float* __restrict__ keys, const int* __restrict__ perm, const float* __restrict__ y_true, unsigned char* __restrict__ lab,
               unsigned char* __restrict__ end, int F, int C) {
  for (long k = blockIdx.x * (long)MT_THREADS + threadIdx.x; k < (long)F * C; k += (long)gridDim.x * MT_THREADS) {
    const int c = (int)(k / F);
    const long p = k - (long)c * F;
    lab[k] = y_true[(size_t)perm[k] * C + c] != 0.0f ? 1 : 0;
    end[k] = (p == F - 1 || keys[k] != keys[k + 1]) ? 1 : 0;
  }
}

// One CTA per (resample, class): weighted AP = sum over thresholds of (tp_k - tp_{k-1}) / P * tp_k / cnt_k with tp / cnt the
// multiplicity-weighted running counts in descending-score order (sklearn's average_precision_score on the materialised
// resample).  NaN when the resample holds a single class (the reference skips those).
__global__ void __launch_bounds__(MT_THREADS)
k_boot_ap(const int* __restrict__ mult, const int* __restrict__ perm, const unsigned char* __restrict__ lab, const unsigned char* __restrict__ end,
          double* __restrict__ ap, int F, int R, int ld_ap) {
  typedef cub::BlockScan<int2, MT_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ double wsum[MT_THREADS / 32];
  __shared__ int2 carry;
  const int r = blockIdx.x, c = blockIdx.y;
  const int* m = mult + (size_t)r * F;
  const int* pm = perm + (size_t)c * F;
  const unsigned char* lb = lab + (size_t)c * F;
  const unsigned char* en = end + (size_t)c * F;
  struct Add2 { __device__ int2 operator()(int2 a, int2 b) const { return make_int2(a.x + b.x, a.y + b.y); } };
  if (threadIdx.x == 0) carry = make_int2(0, 0);
  __syncthreads();
  double acc = 0.0;
  int prev_tp_thread = 0;      // tp at the previous threshold is needed: recomputed through a second scan of "tp at ends"
  for (int base = 0; base < F; base += MT_THREADS) {
    const int p = base + threadIdx.x;
    int w = 0, l = 0, e = 0;
    if (p < F) { w = m[pm[p]]; l = lb[p]; e = en[p]; }
    int2 run;
    Scan(tmp).InclusiveScan(make_int2(w * l, w), run, Add2());
    const int2 c0 = carry;
    __syncthreads();
    const int tp = run.x + c0.x, cnt = run.y + c0.y;
    // previous-threshold tp: the largest tp among earlier threshold ends = max-scan of (e && cnt > 0 ? tp : 0) shifted by one
    // (tp is non-decreasing along p, so "tp at the last earlier end" = max over earlier ends)
    const int mark = (e && w >= 0) ? tp : 0;
    int prev_end_tp;
    {
      typedef cub::BlockScan<int, MT_THREADS> ScanI;
      __shared__ typename ScanI::TempStorage tmp2;
      __shared__ int carry2;
      if (base == 0 && threadIdx.x == 0) carry2 = 0;
      __syncthreads();
      int ex;
      ScanI(tmp2).ExclusiveScan(mark, ex, 0, cub::Max());
      prev_end_tp = max(ex, carry2);
      __syncthreads();
      if (threadIdx.x == MT_THREADS - 1) carry2 = max(prev_end_tp, mark);
    }
    if (p < F && e && cnt > 0) acc += (double)(tp - prev_end_tp) * ((double)tp / (double)cnt);
    if (threadIdx.x == MT_THREADS - 1) carry = make_int2(tp, cnt);
    __syncthreads();
  }
  (void)prev_tp_thread;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < MT_THREADS / 32; i++) t += wsum[i];
    const int P = carry.x, N = carry.y;
    ap[(size_t)c * ld_ap + r] = (P > 0 && P < N) ? t / (double)P : NAN;
  }
}

}  // namespace bn

extern "C" int bn_metrics_bootstrap_ap(const float* y_true, const float* y_score, int F, int C, int n_boot, unsigned long long seed,
                                       const int32_t* multiplicities, double* ap_samples, int device) {
  if (!y_true || !y_score || !ap_samples || F <= 0 || C <= 0 || n_boot <= 0) return set_error(BN_ERR_ARG, "bn_metrics_bootstrap_ap: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return set_error(BN_ERR_CUDA, "no CUDA device for the metric kernels (there is no CPU fallback)");
  }
  MT_CU(cudaSetDevice(device));
  const long n = (long)F * C;
  if (n >= (1L << 31)) return set_error(BN_ERR_ARG, "bn_metrics_bootstrap_ap: more than 2^31 cells");
  const bool dev_in = mt_is_device_ptr(y_score);
  if (dev_in != mt_is_device_ptr(y_true)) return set_error(BN_ERR_ARG, "y_true and y_score must both be host or both be device pointers");
  cudaStream_t st = nullptr;
  DevBuf b_true, b_score, b_keys_in, b_idx_in, b_keys, b_perm, b_lab, b_end, b_mult, b_ap, b_offs, b_tmp;
  const float* d_true = y_true;
  const float* d_score = y_score;
  if (!dev_in) {
    if (b_true.alloc(sizeof(float) * n) || b_score.alloc(sizeof(float) * n)) return set_error(BN_ERR_CUDA, "cudaMalloc failed (bootstrap inputs)");
    MT_CU(cudaMemcpyAsync(b_true.p, y_true, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    MT_CU(cudaMemcpyAsync(b_score.p, y_score, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    d_true = b_true.as<float>();
    d_score = b_score.as<float>();
  }
  // resamples per pass: multiplicity matrix of at most 256 MB
  int rblk = (int)((256L << 20) / ((long)F * 4));
  if (rblk < 1) rblk = 1;
  if (rblk > n_boot) rblk = n_boot;
  if (b_keys_in.alloc(sizeof(float) * n) || b_idx_in.alloc(sizeof(int) * n) || b_keys.alloc(sizeof(float) * (n + 1)) || b_perm.alloc(sizeof(int) * n) ||
      b_lab.alloc(n) || b_end.alloc(n) || b_mult.alloc(sizeof(int) * (size_t)rblk * F) || b_ap.alloc(sizeof(double) * (size_t)C * n_boot) ||
      b_offs.alloc(sizeof(int) * (C + 1)))
    return set_error(BN_ERR_CUDA, "cudaMalloc failed (bootstrap workspace)");
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  long gb = (n + MT_THREADS - 1) / MT_THREADS;
  const int grid = (int)(gb < (long)sms * 8 ? gb : (long)sms * 8);
  // class-major scores; the sort's values are the file indices (k_transpose writes labels: reuse it for the keys only)
  {
    dim3 tg((F + 31) / 32, (C + 31) / 32);
    k_transpose<<<tg, MT_THREADS, 0, st>>>(d_true, d_score, b_keys_in.as<float>(), b_idx_in.as<int>(), F, C);
    std::vector<int> idx((size_t)n);
    for (int c = 0; c < C; c++) for (int f = 0; f < F; f++) idx[(size_t)c * F + f] = f;
    MT_CU(cudaMemcpyAsync(b_idx_in.p, idx.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    std::vector<int> offs(C + 1);
    for (int c = 0; c <= C; c++) offs[c] = (int)((long)c * F);
    MT_CU(cudaMemcpyAsync(b_offs.p, offs.data(), sizeof(int) * (C + 1), cudaMemcpyHostToDevice, st));
    size_t t1 = 0;
    MT_CU(cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, t1, b_keys_in.as<float>(), b_keys.as<float>(), b_idx_in.as<int>(), b_perm.as<int>(),
                                                             (int)n, C, b_offs.as<int>(), b_offs.as<int>() + 1, 0, 32, st));
    if (b_tmp.alloc(t1)) return set_error(BN_ERR_CUDA, "cudaMalloc failed (bootstrap sort workspace)");
    MT_CU(cub::DeviceSegmentedRadixSort::SortPairsDescending(b_tmp.p, t1, b_keys_in.as<float>(), b_keys.as<float>(), b_idx_in.as<int>(), b_perm.as<int>(),
                                                             (int)n, C, b_offs.as<int>(), b_offs.as<int>() + 1, 0, 32, st));
    MT_CU(cudaStreamSynchronize(st));                 // idx / offs are host vectors
    k_boot_prepare<<<grid, MT_THREADS, 0, st>>>(b_keys.as<float>(), b_perm.as<int>(), d_true, b_lab.as<unsigned char>(), b_end.as<unsigned char>(), F, C);
  }
  for (int r0 = 0; r0 < n_boot; r0 += rblk) {
    const int nr = n_boot - r0 < rblk ? n_boot - r0 : rblk;
    if (multiplicities) {
      MT_CU(cudaMemcpyAsync(b_mult.p, multiplicities + (size_t)r0 * F, sizeof(int) * (size_t)nr * F,
                            mt_is_device_ptr(multiplicities) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else {
      MT_CU(cudaMemsetAsync(b_mult.p, 0, sizeof(int) * (size_t)nr * F, st));
      const long work = (((long)F + 3) / 4) * nr;
      long g2 = (work + MT_THREADS - 1) / MT_THREADS;
      k_boot_draw<<<(int)(g2 < (long)sms * 16 ? g2 : (long)sms * 16), MT_THREADS, 0, st>>>(b_mult.as<int>(), F, nr, r0, seed);
    }
    k_boot_ap<<<dim3(nr, C), MT_THREADS, 0, st>>>(b_mult.as<int>(), b_perm.as<int>(), b_lab.as<unsigned char>(), b_end.as<unsigned char>(),
                                                  b_ap.as<double>() + r0, F, nr, n_boot);
  }
  MT_CU(cudaMemcpyAsync(ap_samples, b_ap.p, sizeof(double) * (size_t)C * n_boot, cudaMemcpyDeviceToHost, st));
  MT_CU(cudaStreamSynchronize(st));
  MT_CU(cudaGetLastError());
  return BN_OK;
}
