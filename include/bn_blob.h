/* bn_blob.h -- on-disk / in-memory layout of the flattened model blob.
 *
 * The blob is what `birdnet_stm32.conversion.export_blob` writes from a parsed
 * `.tflite` (reference: the file `tf.lite.Interpreter` loads in
 * birdnet_stm32/models/runners.py:57) and what both the CUDA engine
 * (birdnet-stm32_b200/csrc) and the CPU oracle (oracle/) load.  All
 * quantisation parameters are already resolved to the integers the kernels
 * need (TFLite `QuantizeMultiplier` results computed in double precision at
 * export time, activation clamp ranges, LOGISTIC lookup table), so neither
 * side re-derives them.
 *
 * Little-endian, naturally aligned, every data section 16-byte aligned.
 * Tensor shapes exclude the batch dimension (per audio chunk).
 */
#ifndef BN_BLOB_H
#define BN_BLOB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_BLOB_MAGIC "BNB200\0"   /* 8 bytes incl. the two trailing NULs */
#define BN_BLOB_VERSION 2u

/* tensor element types */
enum { BN_F32 = 0, BN_I8 = 1, BN_I32 = 2 };

/* host frontend kind: what turns PCM into the graph's float input
 * (reference: birdnet_stm32/evaluation/metrics.py:49-69) */
enum {
  BN_FE_NONE = 0,    /* caller feeds the graph input directly                      */
  BN_FE_HYBRID = 1,  /* |STFT| linear spectrogram + per-chunk min-max normalise     */
  BN_FE_MEL = 2,     /* librosa-style Slaney mel + magnitude scaling + normalise    */
  BN_FE_RAW = 3      /* raw waveform x / (max|x| + 1e-6)                            */
};
enum { BN_MAG_NONE = 0, BN_MAG_PWL = 1, BN_MAG_PCEN = 2, BN_MAG_DB = 3 };

/* op kinds (lowered from TFLite builtins; see export_blob.py) */
enum {
  BN_OP_QUANTIZE = 1,    /* f32 -> i8          f[0]=scale  p[0]=zp                              */
  BN_OP_DEQUANTIZE = 2,  /* i8  -> f32         f[0]=scale  p[0]=zp                              */
  BN_OP_TRANSPOSE = 3,   /* p[0..2] = perm of the 3 non-batch dims                               */
  BN_OP_SLICE = 4,       /* p[0..2] = begin, out dims give the size (stride 1)                   */
  BN_OP_FILL = 5,        /* p[0] = value                                                         */
  BN_OP_CONCAT = 6,      /* p[0] = axis (0..2 over non-batch dims), 2 inputs                      */
  BN_OP_CONV2D = 7,      /* see BN_CONV_* param slots                                            */
  BN_OP_DWCONV2D = 8,    /* same slots, weights [kh,kw,C]                                        */
  BN_OP_FC = 9,          /* same slots with kh=kw=1, in [K] out [N]                              */
  BN_OP_ADD = 10,        /* see BN_ADD_* slots; in[1] may be a const broadcast over last dim     */
  BN_OP_MUL = 11,        /* p: in1_zp,in2_zp,out_zp,mult,shift,act_min,act_max,bcast (0 none, 1 const [C],
                            2 per-chunk [1,1,C] over H,W, 3 per-position [H,W,1] over C)                   */
  BN_OP_MEAN = 12,       /* mean over H,W : see BN_MEAN_* slots                                  */
  BN_OP_LOGISTIC = 13,   /* off[0] = 256-byte LUT indexed by (uint8)(q+128)                      */
  BN_OP_RESHAPE = 14,    /* pure re-interpretation (copy)                                        */
  BN_OP_SOFTMAX = 15,    /* over last dim; p[0]=out_zp f[0]=in_scale*beta f[1]=out_scale off[0]=float table[256],
                            table[255-v] = expf(-in_scale*beta*v)                                   */
  BN_OP_PAD = 16,        /* p[0..5] = before/after per non-batch dim, p[6] = pad value           */
  BN_OP_SUM = 17,        /* sum over axis p[0] (non-batch): p[1]=count p[2]=in_zp p[3]=out_zp f[0]=in/out scale
                            f[1]=bias = -in_zp*scale*count                                         */
  BN_OP_REDUCE_MAX = 18, /* max over axes bitmask p[0]                                           */
  BN_OP_REQUANT = 19     /* i8 -> i8 QUANTIZE: p[0]=in_zp p[1]=out_zp p[2]=mult p[3]=shift       */
};

/* param slots for CONV2D / DWCONV2D / FC */
enum {
  BN_CONV_KH = 0, BN_CONV_KW = 1, BN_CONV_SH = 2, BN_CONV_SW = 3,
  BN_CONV_PAD_T = 4, BN_CONV_PAD_L = 5,
  BN_CONV_IN_ZP = 6, BN_CONV_OUT_ZP = 7, BN_CONV_ACT_MIN = 8, BN_CONV_ACT_MAX = 9,
  BN_CONV_CIN = 10, BN_CONV_COUT = 11
  /* off[0]=weights i8  off[1]=bias i32[cout]  off[2]=mult i32[cout]  off[3]=shift i32[cout] */
};
enum {
  BN_ADD_IN1_ZP = 0, BN_ADD_IN2_ZP = 1, BN_ADD_OUT_ZP = 2, BN_ADD_LEFT_SHIFT = 3,
  BN_ADD_M1 = 4, BN_ADD_S1 = 5, BN_ADD_M2 = 6, BN_ADD_S2 = 7, BN_ADD_MO = 8, BN_ADD_SO = 9,
  BN_ADD_ACT_MIN = 10, BN_ADD_ACT_MAX = 11,
  BN_ADD_BCAST = 12  /* 1 when in[1] is a const [C] vector broadcast over the last dim */
};
enum {
  BN_MEAN_COUNT = 0, BN_MEAN_IN_ZP = 1, BN_MEAN_OUT_ZP = 2,
  BN_MEAN_MULT = 3, BN_MEAN_SHIFT = 4,     /* QuantizeMultiplier(s_in / s_out)                 */
  BN_MEAN_MULT_N = 5, BN_MEAN_SHIFT_N = 6, /* variant (ii): 1/N folded into the multiplier     */
  BN_MEAN_KEEP_DIMS = 7
};

typedef struct bn_blob_header {
  char     magic[8];
  uint32_t version;
  uint32_t header_bytes;
  uint32_t n_tensors;
  uint32_t n_ops;
  uint64_t tensors_off;
  uint64_t ops_off;
  uint64_t data_off;
  uint64_t total_bytes;
  /* audio / frontend (reference: _model_config.json keys, training/config.py:43-51) */
  uint32_t frontend_kind;
  uint32_t mag_scale;
  uint32_t sample_rate;
  uint32_t chunk_len;    /* T = int(sample_rate * chunk_duration)      */
  uint32_t n_fft;
  uint32_t hop;          /* chunk_len // spec_width (spectrogram.py:61) */
  uint32_t spec_width;
  uint32_t num_mels;
  uint32_t num_classes;
  int32_t  input_tensor;
  int32_t  output_tensor;
  uint32_t reserved[7];
} bn_blob_header;       /* 128 bytes */

typedef struct bn_blob_tensor {
  int32_t  id;           /* TFLite tensor index (debug taps use it) */
  int32_t  dtype;
  int32_t  rank;         /* non-batch rank, 1..3                     */
  int32_t  dims[3];      /* non-batch dims, leading-padded with 1    */
  float    scale;
  int32_t  zero_point;
  int32_t  is_const;
  int32_t  reserved;
  uint64_t data_off;     /* const payload (absolute) or 0            */
  uint64_t nbytes;       /* bytes per batch item                     */
} bn_blob_tensor;       /* 56 bytes */

typedef struct bn_blob_op {
  int32_t  kind;
  int32_t  tfl_index;    /* index of the TFLite operator it came from */
  int32_t  n_in;
  int32_t  in[3];        /* tensor table slots (not TFLite ids)        */
  int32_t  out;
  int32_t  reserved;
  int32_t  p[24];
  float    f[4];
  uint64_t off[4];       /* absolute blob offsets                      */
} bn_blob_op;           /* 176 bytes */

#ifdef __cplusplus
}
#endif
#endif /* BN_BLOB_H */
