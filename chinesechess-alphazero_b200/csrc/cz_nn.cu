// cz_nn.cu — policy + value network forward (agent/model.py:32-83) on B200.
//
//   packed boards --k_conv_first--> strip activations (5x5 input conv over one-hot planes is a gather-sum
//                                   of <= 25 weight rows per pixel; plane encoding never materialises)
//   2 x blocks of  igemm::k_igemm   3x3 conv as implicit GEMM on tcgen05 (BN folded, +skip, ReLU fused)
//   k_heads                         1x1 policy/value convs + BN + ReLU, value MLP + tanh
//   igemm::k_igemm (GEMM mode)      policy_out Dense 360 -> 2086 on tcgen05
//   k_softmax                       2086-way softmax
//
// BatchNormalization is inference-mode (moving statistics, eps = 1e-3, data/model/model_best_config.json)
// and folded into fp16 weights + fp32 shift.  Weights arrive in Keras layout, caller-owned.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "cz_err.h"
#include "cz_igemm.cuh"
#include "cz_igemm3.cuh"
#include "cz_nn.cuh"

namespace cznn {

#define CZ_CUDA(x)                                                                           \
  do {                                                                                       \
    cudaError_t e__ = (x);                                                                   \
    if (e__ != cudaSuccess) return cz_fail(CZ_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e__)); \
  } while (0)

constexpr int kLabels = CZ_N_LABELS;
// Head widths are a property of the weight file: agent/model.py:47-61 builds 4 policy / 2 value channels, the older configs shipped
// under data/model/ (model_128f.json, model_256f.json: 2 / 4; model_128_l1_config.json: 32 / 4) are served too.
//   pol_k1 = policy features (policy channels x 90) padded to whole 64-column k-blocks, pol_k = 3 * pol_k1:
//            the policy Dense runs as a split-precision GEMM on the tensor cores:
                                   //   x = x_hi + x_lo, w = w_hi + w_lo (fp16 each); logits = x_hi.w_hi + x_lo.w_hi + x_hi.w_lo
                                   // laid out along K as A' = [x_hi | x_lo | x_hi], W' = [w_hi | w_hi | w_lo] -> one GEMM, ~fp32 accuracy
                                   // (fp16 operands alone cost 1.1e-3 of policy probability on the reference's trained 192x10 net)
constexpr int kPolN = 2304;     // 2086 labels padded to 9 N tiles of 256
constexpr float kBnEps = 1e-3f;

// ------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) return cz_fail(CZ_ERR_CUDA, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
  g_encode = (EncodeTiledFn)fn;
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && fn) g_encode_im2col = (EncodeIm2colFn)fn;
  return 0;
}

// fp16 NHWC activations [n][10][9][c] read in im2col mode for a 3x3 "same" convolution: the bounding box of base pixels is
// [-1, dim-2] in w and h (lower corner = -pad, upper corner = pad - (filter-1)), 64 channels x 128 output pixels per load;
// taps outside the image are zero-filled by the TMA unit, and the 128-pixel column walks across rows and images.
static int make_map_im2col(CUtensorMap* m, const void* base, int c, long long n_images) {
  if (load_encode()) return CZ_ERR_CUDA;
  if (!g_encode_im2col) return cz_fail(CZ_ERR_UNSUPPORTED, "cuTensorMapEncodeIm2col not available");
  cuuint64_t dims[4] = {(cuuint64_t)c, 9, 10, (cuuint64_t)n_images};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * 9, (cuuint64_t)c * 2 * 90};
  int lower[2] = {-1, -1}, upper[2] = {-1, -1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, 64, 128,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cz_fail(CZ_ERR_CUDA, "cuTensorMapEncodeIm2col failed: %d", (int)r);
  return 0;
}

// fp16 tensor [rows][w][c] (c contiguous), box {64, box_w, box_r}, 128B swizzle, zero OOB fill
static int make_map_3d(CUtensorMap* m, const void* base, int c, int w, long long rows, int box_w, int box_r) {
  if (load_encode()) return CZ_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * w};
  cuuint32_t box[3] = {64, (cuuint32_t)box_w, (cuuint32_t)box_r};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cz_fail(CZ_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return 0;
}
// fp16 matrix [rows][k] (k contiguous), box {64, box_rows}
static int make_map_2d(CUtensorMap* m, const void* base, int k, long long rows, int box_rows) {
  if (load_encode()) return CZ_ERR_CUDA;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cz_fail(CZ_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return 0;
}

// activation matrix [rows][c] (fp16 or fp32), box {16 columns, 32 rows}: the tiles the conv epilogue loads (skip stream) and
// stores (outputs) with TMA.  16 fp32 = 64-byte rows -> SWIZZLE_64B, 16 fp16 = 32-byte rows -> SWIZZLE_32B.
static int make_map_tile32(CUtensorMap* m, const void* base, int c, long long rows, bool f32) {
  if (load_encode()) return CZ_ERR_CUDA;
  const cuuint64_t es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)c, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)c * es};
  cuuint32_t box[2] = {igemm::kChunkCols3, 32};
  cuuint32_t est[2] = {1, 1};
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                        box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cz_fail(CZ_ERR_CUDA, "cuTensorMapEncodeTiled(tile32) failed: %d", (int)r);
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int N_TILE>
static int launch_igemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const igemm::Args& a, cudaStream_t st) {
  using C = igemm::Cfg<N_TILE>;
  static bool attr_set = false;
  if (!attr_set) {
    CZ_CUDA(cudaFuncSetAttribute(igemm::k_igemm<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  if (tiles <= 0) return 0;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  igemm::k_igemm<N_TILE><<<grid, igemm::kThreads, C::kSmemBytes, st>>>(tmA, tmB, a);
  CZ_CUDA(cudaGetLastError());
  return 0;
}

// CTA-pair conv: B tensor map must have box rows = N_TILE / 2
template <int N_TILE>
static int launch_igemm2_t(const CUtensorMap& tmA, const CUtensorMap& tmB_half, const igemm::Args& a, cudaStream_t st) {
  using C = igemm::Cfg2<N_TILE>;
  static bool attr_set = false;
  if (!attr_set) {
    CZ_CUDA(cudaFuncSetAttribute(igemm::k_igemm2<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int pairs = (a.m_tiles + 1) / 2;
  if (pairs <= 0) return 0;
  const int clusters = pairs < num_sms() / 2 ? pairs : num_sms() / 2;
  igemm::k_igemm2<N_TILE><<<2 * clusters, igemm::kThreads2, C::kSmemBytes, st>>>(tmA, tmB_half, a);
  CZ_CUDA(cudaGetLastError());
  return 0;
}
static int launch_igemm2(int n_tile, const CUtensorMap& tmA, const CUtensorMap& tmB_half, const igemm::Args& a, cudaStream_t st) {
  switch (n_tile) {
    case 64: return launch_igemm2_t<64>(tmA, tmB_half, a, st);
    case 128: return launch_igemm2_t<128>(tmA, tmB_half, a, st);
    case 192: return launch_igemm2_t<192>(tmA, tmB_half, a, st);
    case 256: return launch_igemm2_t<256>(tmA, tmB_half, a, st);
  }
  return cz_fail(CZ_ERR_UNSUPPORTED, "igemm2: unsupported N tile %d", n_tile);
}
// k_igemm3: same mainloop, all-TMA epilogue (cz_igemm3.cuh).  Measured (profiles/r02c_*): +15 % (C=128, no skip) to +41 % (C=128,
// fp16 skip) and +35 % (C=192) over k_igemm2, equal at C=256 without the fp32 skip stream (tensor pipe 90 %); with the fp32
// stream (4-deep operand ring beside 96 KB of epilogue tiles, TMA round trips serialised per chunk) it is ~4 % slower in the
// power-capped 256x20 forward, so that one launch shape keeps k_igemm2.  CZ_EPI=2 / 3 force the old / new epilogue everywhere.
static int epilogue_choice() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CZ_EPI"); v = (e && e[0] == '2') ? 2 : (e && e[0] == '3') ? 3 : 0; }
  return v;
}
static bool use_tma_epilogue() { return epilogue_choice() != 2; }
static bool use_tma_epilogue_for(int c, bool fp32_stream) {
  if (epilogue_choice() == 2) return false;
  if (epilogue_choice() == 3) return true;
  return !(c == 256 && fp32_stream);
}
static int g_cluster4 = 0;                 // set by nn_create from CZ_CLUSTER4 (launch_igemm3, C = 256)
template <int N_TILE, int MT, int PAIRS = 1>
static int launch_igemm3_t(const CUtensorMap& tmA, const CUtensorMap& tmB_half, const CUtensorMap& tmOut16, const CUtensorMap& tmSkip,
                           const CUtensorMap& tmOut32, const igemm::Args& a, int skip_mode, bool out32, cudaStream_t st, int n_split = 1) {
  using C = igemm::Cfg3<N_TILE, MT>;
  static bool attr_set = false;
  static int max_clusters = 0;
  if (!attr_set) {
    CZ_CUDA(cudaFuncSetAttribute(igemm::k_igemm3<N_TILE, MT, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, igemm::kSmemLimit3));
    attr_set = true;
    max_clusters = num_sms() / (2 * PAIRS);
    if (PAIRS > 1) {                       // 4-CTA clusters must fit inside a GPC: ask how many can be resident at once
      cudaLaunchConfig_t oc;
      memset(&oc, 0, sizeof(oc));
      oc.gridDim = dim3(2 * PAIRS * max_clusters); oc.blockDim = dim3(igemm::kThreads2); oc.dynamicSmemBytes = igemm::kSmemLimit3;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, igemm::k_igemm3<N_TILE, MT, PAIRS>, &oc) == cudaSuccess && nc > 0 && nc < max_clusters) max_clusters = nc;
      else (void)cudaGetLastError();
      fprintf(stderr, "[cczero] 4-CTA clusters of k_igemm3<%d, %d>: %d resident at once (%d of %d SMs)\n", N_TILE, MT, max_clusters,
              4 * max_clusters, num_sms());
    }
  }
  igemm::Args3 p;
  p.a = a;
  p.skip_mode = skip_mode; p.out32 = out32 ? 1 : 0; p.n_split = n_split;
  p.fbytes = (skip_mode == 2 || out32) ? 2048 : (skip_mode == 1 ? 1024 : 0);
  { static int nf = -1; if (nf < 0) { const char* e = getenv("CZ_NF"); nf = e ? atoi(e) : 3; if (nf < 3) nf = 3; if (nf > igemm::kMaxNF3) nf = igemm::kMaxNF3; } p.nf = nf; }
  { static int sp = -1; if (sp < 0) { const char* e = getenv("CZ_SPLIT_PROD"); sp = (e && e[0] == '1') ? 1 : 0; } p.split_producer = sp; }
  p.stages = C::max_stages(p.fbytes, p.nf);
  { static int cap = -1; if (cap < 0) { const char* e = getenv("CZ_STAGES"); cap = e ? atoi(e) : 0; } if (cap > 1 && cap < p.stages) p.stages = cap; }
  if (p.stages < 2) return cz_fail(CZ_ERR_UNSUPPORTED, "igemm3: no room for the operand ring");
  const int pairs = ((a.n_dev ? (a.rows + igemm::kTileM - 1) / igemm::kTileM : a.m_tiles) + 2 * MT - 1) / (2 * MT);
  if (pairs <= 0) return 0;
  const int items = (pairs * (n_split > 1 ? n_split : 1) + PAIRS - 1) / PAIRS;     // cluster-level steps
  const int clusters = items < max_clusters ? items : max_clusters;
  // Programmatic dependent launch: this conv's CTAs may become resident and run their prologue (barriers, TMEM, tensor map
  // prefetch) while the previous kernel of the stream is still running; griddepcontrol.wait in the kernel orders the data.
  // Measured on one box, interleaved (profiles/r02o_*): UCI go depth 8 56.7 -> 54.3 ms, c2 1.521 -> 1.540 M sims/s, c3 equal.
  // CZ_PDL=0 launches the plain way.
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("CZ_PDL"); pdl = (e && e[0] == '0') ? 0 : 1; }
  if (pdl) {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(2 * PAIRS * clusters); lc.blockDim = dim3(igemm::kThreads2);
    lc.dynamicSmemBytes = C::smem_bytes(p.stages, p.fbytes, p.nf); lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CZ_CUDA(cudaLaunchKernelEx(&lc, igemm::k_igemm3<N_TILE, MT, PAIRS>, tmA, tmB_half, tmOut16, tmSkip, tmOut32, p));
  } else {
    igemm::k_igemm3<N_TILE, MT, PAIRS><<<2 * PAIRS * clusters, igemm::kThreads2, C::smem_bytes(p.stages, p.fbytes, p.nf), st>>>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, p);
  }
  CZ_CUDA(cudaGetLastError());
  return 0;
}
static int launch_igemm3(int n_tile, const CUtensorMap& tmA, const CUtensorMap& tmB_half, const CUtensorMap& tmOut16, const CUtensorMap& tmSkip,
                         const CUtensorMap& tmOut32, const igemm::Args& a, int skip_mode, bool out32, cudaStream_t st) {
  // two M-tiles per CTA against each weight stage wherever the accumulators fit TMEM (C <= 128); CZ_MT=1 forces the single-tile form
  static int mt2 = -1;
  if (mt2 < 0) { const char* e = getenv("CZ_MT"); mt2 = (e && e[0] == '1') ? 0 : 1; }
  switch (n_tile) {
    case 64: return mt2 ? launch_igemm3_t<64, 2>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st)
                        : launch_igemm3_t<64, 1>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st);
    case 128: return mt2 ? launch_igemm3_t<128, 2>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st)
                         : launch_igemm3_t<128, 1>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st);
    case 192: return launch_igemm3_t<192, 1>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st);
    case 256: {
      // CZ_CLUSTER4=1 when a network runtime is created (experiment, off by default): two CTA pairs per cluster share every
      // weight stage by TMA multicast.  Bit-identical results (test_cluster4_weight_multicast); on the pool's B200 it is 1.6 %
      // SLOWER on the 256x20 forward (profiles/r02y_ab_nn.log) — see DESIGN.md section 3.1.
      if (g_cluster4) return launch_igemm3_t<256, 1, 2>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st);
      return launch_igemm3_t<256, 1>(tmA, tmB_half, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st);
    }
  }
  return cz_fail(CZ_ERR_UNSUPPORTED, "igemm3: unsupported N tile %d", n_tile);
}
// Small batches (one game's leaves: the UCI / play_games latency path): 64-column tiles, pairs x C/64 work items, as long as
// every item gets its own CTA pair in one wave.  tmB_32 = the weight map with 32-row boxes.  CZ_NSPLIT=0 turns it off.
static bool use_n_split(int n_boards, int c) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CZ_NSPLIT"); on = (e && e[0] == '0') ? 0 : 1; }
  if (!on || c <= 64) return false;
  const int pairs = ((n_boards * 90 + igemm::kTileM - 1) / igemm::kTileM + 1) / 2;
  return pairs * (c / 64) <= num_sms() / 2;
}
static int launch_igemm3_split(int c, const CUtensorMap& tmA, const CUtensorMap& tmB_32, const CUtensorMap& tmOut16, const CUtensorMap& tmSkip,
                               const CUtensorMap& tmOut32, const igemm::Args& a, int skip_mode, bool out32, cudaStream_t st) {
  return launch_igemm3_t<64, 1>(tmA, tmB_32, tmOut16, tmSkip, tmOut32, a, skip_mode, out32, st, c / 64);
}
static bool use_im2col() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CZ_CONV_STRIP"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}
static bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CZ_IGEMM_1CTA"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

static int launch_igemm(int n_tile, const CUtensorMap& tmA, const CUtensorMap& tmB, const igemm::Args& a, cudaStream_t st) {
  switch (n_tile) {
    case 64: return launch_igemm_t<64>(tmA, tmB, a, st);
    case 128: return launch_igemm_t<128>(tmA, tmB, a, st);
    case 192: return launch_igemm_t<192>(tmA, tmB, a, st);
    case 256: return launch_igemm_t<256>(tmA, tmB, a, st);
  }
  return cz_fail(CZ_ERR_UNSUPPORTED, "igemm: unsupported N tile %d (filters must be 64/128/192/256)", n_tile);
}

static igemm::Args conv_args(int n_boards, int c, const float* bias, const __half* residual, __half* out, int relu) {
  igemm::Args a;
  memset(&a, 0, sizeof(a));
  a.n_taps = 9; a.k_chunks = c / 64; a.box_w = 9; a.box_r = 14;
  a.rows = n_boards * 11; a.m_tiles = (a.rows + 13) / 14; a.n_tiles = 1;
  a.n_total = c; a.n_valid = c; a.ldo = c; a.conv = 1; a.relu = relu; a.out_f32 = 0;
  a.bias = bias; a.residual = residual; a.out = out; a.a_bytes = 64 * 9 * 14 * 2;
  return a;
}

// dense pixel layout [n_boards*90][c] + im2col TMA (CTA-pair kernel only)
static igemm::Args conv_args_dense(int n_boards, int c, const float* bias, const __half* residual, __half* out, int relu) {
  igemm::Args a = conv_args(n_boards, c, bias, residual, out, relu);
  a.conv = 2; a.rows = n_boards * 90; a.m_tiles = (a.rows + 127) / 128; a.a_bytes = 128 * 128;
  return a;
}

static igemm::Args dense_args(int m, int n_valid, int n_pad, int k_pad, int n_tile, const float* bias, float* out, int ldo) {
  igemm::Args a;
  memset(&a, 0, sizeof(a));
  a.n_taps = 1; a.k_chunks = k_pad / 64; a.box_w = 1; a.box_r = 128;
  a.rows = m; a.m_tiles = (m + 127) / 128; a.n_tiles = n_pad / n_tile;
  a.n_total = n_pad; a.n_valid = n_valid; a.ldo = ldo; a.conv = 0; a.relu = 0; a.out_f32 = 1;
  a.bias = bias; a.residual = nullptr; a.out = out; a.a_bytes = 64 * 128 * 2;
  return a;
}

// ------------------------------------------------------------------------------------------------
// small kernels
// packed board -> plane index per network pixel (pix = r*9 + col, r = 9 - y), -1 = empty
__device__ __forceinline__ int plane_of(uint8_t c) { return c == 0 ? -1 : ((c & 8) ? c - 2 : c - 1); }

// 5x5 "same" input convolution + BN + ReLU from packed boards (model.py:34-41, static_env.py:137-156 fused).
// The 14 input planes are one-hot, so an output pixel is the sum of <= 25 weight rows w[tap][plane(piece on the
// tapped square)][:].  Phase 1: 90 threads list the occupied taps of their pixel (row index = tap*14 + plane);
// phase 2: every thread owns two adjacent output channels of a subset of the pixels and walks their lists (half2 loads, fp32
// accumulate).  grid = batch, block = (C/2) * n_groups >= 96 threads (conv_first_threads).  w: HWIO [5][5][in_planes][C] fp16 (BN scale folded).
// in_planes = 28 (use_history, static_env.py:158-194): every board record is followed by the history board whose pieces
// select planes 14-27; board_stride = bytes between records.
__global__ void k_conv_first(const uint8_t* __restrict__ boards, const __half* __restrict__ w,
                             const float* __restrict__ shift, __half* __restrict__ out, float* __restrict__ out32, int c_out,
                             int board_pixels, int in_planes, int board_stride, const int* __restrict__ n_dev) {
  if ((int)blockIdx.x >= __ldg(n_dev)) return;              // fixed-shape launch: the batch size lives on the device
  __shared__ int8_t pl[2][90];
  __shared__ uint16_t rows[90][52];
  __shared__ uint8_t cnt[90];
  const int b = blockIdx.x, t = threadIdx.x;
  const int n_boards = in_planes / 14;
  if (t < 90) {
    const int r = t / 9, col = t % 9;
    for (int h = 0; h < n_boards; ++h)
      pl[h][t] = (int8_t)plane_of(boards[(size_t)b * board_stride + h * CZ_BOARD_STRIDE + (9 - r) * 9 + col]);
  }
  __syncthreads();
  // gridDim.y slices of the 90 pixels (small batches: one CTA per position would leave the GPU to a handful of CTAs that each
  // walk 90 x <= 50 dependent-latency loads)
  const int per = (90 + (int)gridDim.y - 1) / (int)gridDim.y, p0 = (int)blockIdx.y * per, p1 = p0 + per < 90 ? p0 + per : 90;
  if (t < p1 - p0) {
    const int r = (p0 + t) / 9, col = (p0 + t) % 9;
    int n = 0;
    for (int kh = 0; kh < 5; ++kh) {
      const int rr = r + kh - 2;
      if (rr < 0 || rr > 9) continue;
      for (int kw = 0; kw < 5; ++kw) {
        const int cc = col + kw - 2;
        if (cc < 0 || cc > 8) continue;
        for (int h = 0; h < n_boards; ++h) {
          const int p = pl[h][rr * 9 + cc];
          if (p >= 0) rows[t][n++] = (uint16_t)((kh * 5 + kw) * in_planes + h * 14 + p);
        }
      }
    }
    cnt[t] = (uint8_t)n;
  }
  __syncthreads();
  // phase 2: thread = (channel pair, pixel group): blockDim.x = (c_out / 2) * n_groups, group g takes pixels g, g + n_groups, ...
  const int pairs = c_out / 2;
  const int c = 2 * (t % pairs), grp = t / pairs, n_groups = blockDim.x / pairs;
  if (grp >= n_groups) return;
  const float2 sh = *reinterpret_cast<const float2*>(shift + c);
  __half* o = out + (size_t)b * board_pixels * c_out;
  const __half2* w2 = reinterpret_cast<const __half2*>(w + c);
  const int stride2 = c_out / 2;
  for (int pix = p0 + grp; pix < p1; pix += n_groups) {
    float a0 = sh.x, a1 = sh.y;
    const int n = cnt[pix - p0];
    for (int k = 0; k < n; ++k) {
      const float2 v = __half22float2(__ldg(w2 + (size_t)rows[pix - p0][k] * stride2));
      a0 += v.x; a1 += v.y;
    }
    a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f);
    *reinterpret_cast<__half2*>(o + (size_t)pix * c_out + c) = __floats2half2_rn(a0, a1);
    if (out32) *reinterpret_cast<float2*>(out32 + ((size_t)b * board_pixels + pix) * c_out + c) = make_float2(a0, a1);
  }
  if (blockIdx.y == 0)
  for (int col = 90 + grp; col < board_pixels; col += n_groups)
    *reinterpret_cast<__half2*>(o + (size_t)col * c_out + c) = __floats2half2_rn(0.f, 0.f);          // separator row (strip layout)
}

// one-hot planes [B][in_planes][10][9] f32 -> packed boards (inverse of state_to_planes / state_history_to_planes):
// in_planes / 14 consecutive board records per position
__global__ void k_planes_to_boards(const float* __restrict__ planes, uint8_t* __restrict__ boards, int n, int in_planes) {
  const int b = blockIdx.x;
  if (b >= n) return;
  const int t = threadIdx.x;
  if (t < 96) {
    for (int h = 0; h < in_planes / 14; ++h) {
      uint8_t code = 0;
      if (t < 90) {
        const int y = t / 9, x = t % 9, r = 9 - y;
        for (int p = 0; p < 14; ++p)
          if (planes[((size_t)b * in_planes + h * 14 + p) * 90 + r * 9 + x] > 0.5f) code = (uint8_t)(p < 7 ? p + 1 : p + 2);
      }
      boards[((size_t)b * (in_planes / 14) + h) * CZ_BOARD_STRIDE + t] = code;
    }
  }
}

// Heads (model.py:47-63): 1x1 conv to 4 policy + 2 value channels, BN, ReLU; policy features to the GEMM
// operand [B][384] (index c*90 + pix, Keras Flatten of channels_first); value: Dense(H)+ReLU, Dense(1)+tanh.
// A block handles kHeadPos positions so the 180 x H value weights are read once per group.
// Phase 1: one thread per (position, pixel): it streams that pixel's C activations (16-byte loads along its own row; the rows
// of a warp's 32 pixels are adjacent in memory) against the 6 x C folded weights held in shared memory (broadcast reads) —
// no cross-lane reduction.  (The round-1 version reduced six sums with 30 shuffles per pixel: 110 us per 2048 positions.)
constexpr int kHeadPos = 4;
constexpr int kMaxHeadOut = 36;                                   // 32 policy + 4 value channels
static size_t heads_smem_bytes(int c_in, int n_out) { return ((size_t)kHeadPos * n_out * 90 + (size_t)(c_in / 8) * n_out * 8 + kHeadPos * 8) * sizeof(float); }
__global__ void __launch_bounds__(256) k_heads(const __half* __restrict__ act, const float* __restrict__ act32, int c_in,
                                                const int* __restrict__ n_dev, int board_pixels, int pol_c, int val_c, int pol_k1,
                                                const float* __restrict__ wh,      // [pol_c + val_c][c_in], BN scale folded
                                                const float* __restrict__ shifth,  // [pol_c + val_c]
                                                const float* __restrict__ wv1,     // [val_c * 90][H]
                                                const float* __restrict__ bv1,     // [H]
                                                const float* __restrict__ wv2,     // [H]
                                                const float* __restrict__ bv2,     // [1]
                                                int hidden, __half* __restrict__ pol_feat, float* __restrict__ value,
                                                int hp) {          // positions per block: kHeadPos, or 1 for small batches
  extern __shared__ __align__(16) float hsm[];
  const int n_out = pol_c + val_c;
  float* feat = hsm;                                         // [kHeadPos][n_out][90]
  float* wsm = feat + kHeadPos * n_out * 90;                 // [c_in / 8][n_out][8]
  float* red = wsm + (c_in / 8) * n_out * 8;                 // [kHeadPos][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * hp;
  const int n_pos = __ldg(n_dev);
  if (b0 >= n_pos) return;
  const int npos = n_pos - b0 < hp ? n_pos - b0 : hp;
  for (int i = tid; i < n_out * c_in; i += 256) { const int o = i / c_in, c = i % c_in; wsm[((c >> 3) * n_out + o) * 8 + (c & 7)] = __ldg(wh + i); }
  __syncthreads();
  for (int item = tid; item < npos * 90; item += 256) {
    const int p = item / 90, pix = item % 90;
    const size_t row = ((size_t)(b0 + p) * board_pixels + pix) * c_in;
    for (int o0 = 0; o0 < n_out; o0 += 6) {                  // six outputs per pass over the pixel's row (the row stays in L1)
      float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int g = 0; g < c_in / 8; ++g) {
        float x[8];
        if (act32) {
          const float4* a4 = reinterpret_cast<const float4*>(act32 + row + g * 8);
          const float4 u = __ldg(a4), w4 = __ldg(a4 + 1);
          x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = w4.x; x[5] = w4.y; x[6] = w4.z; x[7] = w4.w;
        } else {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(act + row + g * 8));
          const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); x[2 * j] = f.x; x[2 * j + 1] = f.y; }
        }
#pragma unroll
        for (int o = 0; o < 6; ++o) {
          if (o0 + o < n_out) {
            const float* wp = wsm + (g * n_out + o0 + o) * 8;
            const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
            s[o] += x[0] * wa.x + x[1] * wa.y + x[2] * wa.z + x[3] * wa.w + x[4] * wb.x + x[5] * wb.y + x[6] * wb.z + x[7] * wb.w;
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 6; ++o)
        if (o0 + o < n_out) feat[(p * n_out + o0 + o) * 90 + pix] = fmaxf(s[o] + __ldg(shifth + o0 + o), 0.f);
    }
  }
  __syncthreads();
  const int pol_in = pol_c * 90, pol_k = 3 * pol_k1;
  for (int i = tid; i < npos * pol_k1; i += 256) {
    const int p = i / pol_k1, k = i % pol_k1;
    const float f = k < pol_in ? feat[(p * n_out + k / 90) * 90 + k % 90] : 0.f;      // Keras Flatten of channels_first: c*90 + pix
    const __half hi = __float2half_rn(f);
    const __half lo = __float2half_rn(f - __half2float(hi));
    __half* row = pol_feat + (size_t)(b0 + p) * pol_k;
    row[k] = hi; row[pol_k1 + k] = lo; row[2 * pol_k1 + k] = hi;
  }
  float h[kHeadPos];
#pragma unroll
  for (int p = 0; p < kHeadPos; ++p) h[p] = 0.f;
  if (tid < hidden) {
    float acc[kHeadPos];
    const float bb = bv1[tid];
#pragma unroll
    for (int p = 0; p < kHeadPos; ++p) acc[p] = bb;
#pragma unroll 10
    for (int i = 0; i < val_c * 90; ++i) {
      const float wv = __ldg(wv1 + (size_t)i * hidden + tid);
#pragma unroll
      for (int p = 0; p < kHeadPos; ++p) acc[p] += feat[(p * n_out + pol_c + i / 90) * 90 + i % 90] * wv;
    }
    const float w2 = wv2[tid];
#pragma unroll
    for (int p = 0; p < kHeadPos; ++p) h[p] = fmaxf(acc[p], 0.f) * w2;
  }
#pragma unroll
  for (int p = 0; p < kHeadPos; ++p) {
    float v = h[p];
    for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if (lane == 0) red[p * 8 + warp] = v;
  }
  __syncthreads();
  if (tid < npos) {
    float sum = bv2[0];
    for (int i = 0; i < 8; ++i) sum += red[tid * 8 + i];
    value[b0 + tid] = tanhf(sum);
  }
}

// Softmax over the 2086 labels, finished from the per-N-tile statistics the policy GEMM epilogue wrote:
//   stats[row][t] = {m_t = max_j x_j, s_t = sum_j exp(x_j - m_t)} over the valid columns of tile t
//   m = max_t m_t,  S = sum_t s_t * exp(m_t - m),  p_j = exp(x_j - m) * (1 / S)
// `policy_prob` is the ONE definition of a policy probability in this library: k_softmax (the [B][2086] vector the
// reference-facing API returns) and k_legal_priors (only the legal moves of a search leaf) both evaluate it, so the
// integrated search sees bit for bit the numbers an external caller of cz_nn_forward would feed back.
struct RowStat { float mx, inv; };
__device__ __forceinline__ RowStat combine_stats(const float2* __restrict__ st, int n_tiles) {
  float mx = -INFINITY;
  for (int t = 0; t < n_tiles; ++t) mx = fmaxf(mx, __ldg(&st[t].x));
  float s = 0.f;
  for (int t = 0; t < n_tiles; ++t) { const float2 v = __ldg(st + t); s = __fmaf_rn(v.y, expf(v.x - mx), s); }
  RowStat r; r.mx = mx; r.inv = __frcp_rn(s);
  return r;
}
__device__ __forceinline__ float policy_prob(float logit, const RowStat& r) { return __fmul_rn(expf(logit - r.mx), r.inv); }

// logits [B][ldl] f32 -> policy [B][2086] f32. grid = batch, block = 256.
__global__ void __launch_bounds__(256) k_softmax(const float* __restrict__ logits, int ldl, const float2* __restrict__ stats, int n_tiles,
                                                  float* __restrict__ policy) {
  const int b = blockIdx.x;
  const RowStat rs = combine_stats(stats + (size_t)b * n_tiles, n_tiles);
  const float* l = logits + (size_t)b * ldl;
  for (int i = threadIdx.x; i < kLabels; i += 256) policy[(size_t)b * kLabels + i] = policy_prob(l[i], rs);
}

// Integrated search: softmax probabilities of the LEGAL moves of every leaf only (player.py:272-284 reads nothing else of the
// policy vector).  labels [n][CZ_MAX_MOVES] int16 (-1 = the move has no label), counts [n]; out [n][CZ_MAX_MOVES] f32.  Warp per leaf.
__global__ void __launch_bounds__(128) k_legal_priors(const float* __restrict__ logits, int ldl, const float2* __restrict__ stats, int n_tiles,
                                                       const int16_t* __restrict__ labels, const int32_t* __restrict__ counts,
                                                       const int* __restrict__ n_dev, float* __restrict__ out) {
  const int leaf = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (leaf >= __ldg(n_dev)) return;
  const RowStat rs = combine_stats(stats + (size_t)leaf * n_tiles, n_tiles);
  const float* l = logits + (size_t)leaf * ldl;
  const int L = counts[leaf];
  for (int i = lane; i < L; i += 32) {
    const int lab = labels[(size_t)leaf * CZ_MAX_MOVES + i];
    out[(size_t)leaf * CZ_MAX_MOVES + i] = lab >= 0 ? policy_prob(l[lab], rs) : 0.f;
  }
}
__global__ void k_set_int(int* p, int v) { *p = v; }

// ---- weight preparation (Keras layout f32 -> folded operands)
__global__ void k_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float* scale,
                          float* shift, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    const float s = gamma[i] / sqrtf(var[i] + kBnEps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s;
  }
}
// HWIO [kh][kw][ci][co] -> same layout fp16 with scale[co] folded (first conv)
__global__ void k_prep_hwio(const float* w, const float* scale, __half* out, long long n, int co) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(w[i] * scale[i % co]);
}
// HWIO [3][3][ci][co] -> [tap][co][ci] fp16, scale[co] folded (B operand, K-major)
__global__ void k_prep_conv3(const float* w, const float* scale, __half* out, int c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = 9LL * c * c;
  if (i < n) {
    const int ci = (int)(i % c), co = (int)((i / c) % c), tap = (int)(i / ((long long)c * c));
    out[i] = __float2half_rn(w[((long long)tap * c + ci) * c + co] * scale[co]);
  }
}
// 1x1 conv HWIO [1][1][ci][co] -> [co][ci] f32 rows at out_row0.., scale folded
__global__ void k_prep_1x1(const float* w, const float* scale, float* out, int ci_n, int co_n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ci_n * co_n) {
    const int ci = i % ci_n, co = i / ci_n;
    out[(size_t)co * ci_n + ci] = w[(size_t)ci * co_n + co] * scale[co];
  }
}
// Dense (in,out) [pol_in][2086] -> [kPolN][3 * pol_k1] fp16 (zero padded), K-major, split as [w_hi | w_hi | w_lo]
__global__ void k_prep_policy(const float* w, __half* out, int pol_in, int pol_k1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)kPolN * pol_k1) {
    const int k = (int)(i % pol_k1), n = (int)(i / pol_k1);
    const float f = (k < pol_in && n < kLabels) ? w[(size_t)k * kLabels + n] : 0.f;
    const __half hi = __float2half_rn(f);
    const __half lo = __float2half_rn(f - __half2float(hi));
    __half* row = out + (size_t)n * 3 * pol_k1;
    row[k] = hi; row[pol_k1 + k] = hi; row[2 * pol_k1 + k] = lo;
  }
}
__global__ void k_copy_pad(const float* src, float* dst, int n_src, int n_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_dst) dst[i] = i < n_src ? src[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------
struct Carver {
  uint8_t* base; size_t off, cap;
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// One network's folded weights (the arena holds two: best vs next generation, worker/evaluator.py:28-82).
struct NetWeights {
  __half* w_first; float* shift_first;
  __half* w_conv;  float* shift_conv;
  float *wh, *shifth, *wv1, *bv1, *wv2, *bv2;
  __half* w_pol; float* b_pol;
  CUtensorMap map_wpol;
  std::vector<CUtensorMap> map_w, map_w_half, map_w_32;
  bool ready;
};

struct NnRuntime {
  int filters, blocks, value_fc, max_batch;
  int pol_c, val_c, pol_k1;               // head widths (policy / value conv channels) and the padded policy feature count
  NetWeights nets[2]; int n_nets, cur;     // the weight fields below alias nets[cur] (select_net / store_net)
  cudaStream_t stream;
  bool ready;
  uint64_t launches;
  // activations
  __half *x, *t, *y, *pol_feat;
  float *x32, *y32;                      // fp32 skip stream (dense layout only)
  float* logits;
  float2* stats;                         // [max_batch][kPolN / 256] softmax statistics of the policy GEMM's N tiles
  int* n_scalar;                         // device copy of a host-known batch size (reference-facing forward)
  bool heads_attr;                       // k_heads was granted > 48 KB of dynamic shared memory (wide legacy heads)
  bool capturing;                        // inside cudaStreamBeginCapture: no event records / synchronisation
  size_t prof_open;                      // event pair opened by nn_prof_begin
  uint8_t* boards_tmp;
  int in_planes;                         // 14, or 28 with use_history (board + history board per position)
  // weights
  __half* w_first; float* shift_first;
  __half* w_conv;  float* shift_conv;      // [2*blocks][9*C*C], [2*blocks][C]
  float *wh, *shifth, *wv1, *bv1, *wv2, *bv2;
  __half* w_pol; float* b_pol;
  float* scratch;                            // 2*C floats for BN folding
  // tensor maps
  CUtensorMap map_x, map_t, map_y, map_pf, map_wpol;
  std::vector<CUtensorMap> map_w;
  std::vector<CUtensorMap> map_w_half;   // box rows = C/2 for the CTA-pair kernel
  std::vector<CUtensorMap> map_w_32;     // box rows = 32: 64-column tiles of the small-batch launches (use_n_split)
  bool fp32_skip;                        // keep the residual (skip) stream in fp32: halves the value error of deep nets, ~+30 % time
  int board_pixels;                      // 99 = strip layout (separator row per board), 90 = dense + im2col TMA
  CUtensorMap imap_x, imap_t, imap_y;    // im2col maps of the three activation buffers (dense layout)
  CUtensorMap omap_x, omap_t, omap_y;    // 32x32 fp16 tile maps of the same buffers (conv epilogue: TMA stores / fp16 skip loads)
  CUtensorMap fmap_x32, fmap_y32;        // 32x32 fp32 tile maps of the fp32 skip stream
  // optional CUDA-event timing of the residual-tower launches (bench.py roofline)
  bool profile;
  std::vector<cudaEvent_t> ev;        // pairs, recycled
  size_t ev_used;
  std::vector<double> ev_flops;       // algorithmic flops bracketed by pair i
  double prof_ms, prof_flops; uint64_t prof_launches;
};

static void store_net(NnRuntime* r, int k) {
  NetWeights& n = r->nets[k];
  n.w_first = r->w_first; n.shift_first = r->shift_first; n.w_conv = r->w_conv; n.shift_conv = r->shift_conv;
  n.wh = r->wh; n.shifth = r->shifth; n.wv1 = r->wv1; n.bv1 = r->bv1; n.wv2 = r->wv2; n.bv2 = r->bv2;
  n.w_pol = r->w_pol; n.b_pol = r->b_pol; n.map_wpol = r->map_wpol; n.map_w = r->map_w; n.map_w_half = r->map_w_half; n.map_w_32 = r->map_w_32;
  n.ready = r->ready;
}
static void select_net(NnRuntime* r, int k) {
  if (r->cur == k) return;
  store_net(r, r->cur);
  const NetWeights& n = r->nets[k];
  r->w_first = n.w_first; r->shift_first = n.shift_first; r->w_conv = n.w_conv; r->shift_conv = n.shift_conv;
  r->wh = n.wh; r->shifth = n.shifth; r->wv1 = n.wv1; r->bv1 = n.bv1; r->wv2 = n.wv2; r->bv2 = n.bv2;
  r->w_pol = n.w_pol; r->b_pol = n.b_pol; r->map_wpol = n.map_wpol; r->map_w = n.map_w; r->map_w_half = n.map_w_half; r->map_w_32 = n.map_w_32;
  r->ready = n.ready;
  r->cur = k;
}

static void prof_collect(NnRuntime* r) {
  for (size_t i = 0; i + 1 < r->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r->ev[i], r->ev[i + 1]) == cudaSuccess) {
      r->prof_ms += ms; r->prof_flops += r->ev_flops[i / 2]; r->prof_launches += (uint64_t)(2 * r->blocks);
    }
  }
  r->ev_used = 0;
}

static void layout(NnRuntime* r, Carver& cv) {
  const int c = r->filters;
  const size_t act = (size_t)r->max_batch * 11 * 9 * c * sizeof(__half);
  r->x = (__half*)cv.take(act);
  r->t = (__half*)cv.take(act);
  r->y = (__half*)cv.take(act);
  r->x32 = (float*)cv.take(act * 2);
  r->y32 = (float*)cv.take(act * 2);
  r->pol_feat = (__half*)cv.take(((size_t)r->max_batch + 128) * 3 * r->pol_k1 * sizeof(__half));
  r->logits = (float*)cv.take((size_t)r->max_batch * kPolN * sizeof(float));
  r->stats = (float2*)cv.take((size_t)r->max_batch * (kPolN / 256) * sizeof(float2));
  r->n_scalar = (int*)cv.take(64);
  r->boards_tmp = (uint8_t*)cv.take((size_t)r->max_batch * 2 * CZ_BOARD_STRIDE);
  for (int net = 0; net < r->n_nets; ++net) {
  r->w_first = (__half*)cv.take((size_t)25 * 28 * c * sizeof(__half));
  r->shift_first = (float*)cv.take(c * sizeof(float));
  r->w_conv = (__half*)cv.take((size_t)2 * r->blocks * 9 * c * c * sizeof(__half));
  r->shift_conv = (float*)cv.take((size_t)2 * r->blocks * c * sizeof(float));
  r->wh = (float*)cv.take((size_t)(r->pol_c + r->val_c) * c * sizeof(float));
  r->shifth = (float*)cv.take((size_t)(r->pol_c + r->val_c + 4) * sizeof(float));
  r->wv1 = (float*)cv.take((size_t)r->val_c * 90 * r->value_fc * sizeof(float));
  r->bv1 = (float*)cv.take(r->value_fc * sizeof(float));
  r->wv2 = (float*)cv.take(r->value_fc * sizeof(float));
  r->bv2 = (float*)cv.take(4 * sizeof(float));
  r->w_pol = (__half*)cv.take((size_t)kPolN * 3 * r->pol_k1 * sizeof(__half));
  r->b_pol = (float*)cv.take(kPolN * sizeof(float));
  r->ready = false;
  r->cur = net;
  store_net(r, net);
  }
  if (r->n_nets > 1) { r->cur = r->n_nets - 1; select_net(r, 0); }
  r->cur = 0;
  r->scratch = (float*)cv.take((size_t)(2 * 256 + 2 * kMaxHeadOut) * sizeof(float));
}

static void set_heads(NnRuntime* r, int pol_c, int val_c) {
  r->pol_c = pol_c > 0 ? pol_c : 4; r->val_c = val_c > 0 ? val_c : 2;     // agent/model.py:47-61 defaults
  r->pol_k1 = (r->pol_c * 90 + 63) / 64 * 64;
}
size_t nn_workspace_bytes(int filters, int blocks, int value_fc, int max_batch, int n_nets, int pol_c, int val_c) {
  NnRuntime tmp;
  tmp.filters = filters; tmp.blocks = blocks; tmp.value_fc = value_fc; tmp.max_batch = max_batch; tmp.n_nets = n_nets; tmp.cur = 0;
  set_heads(&tmp, pol_c, val_c);
  Carver cv{nullptr, 0, 0};
  layout(&tmp, cv);
  return cv.off + 4096;
}

NnRuntime* nn_create(int device, int filters, int blocks, int value_fc, int max_batch, void* workspace, size_t bytes,
                     void* stream, int fp32_skip_mode, int n_nets, int in_planes, int pol_c, int val_c) {
  (void)device;
  if (filters % 64 != 0 || filters < 64 || filters > 256) { cz_fail(CZ_ERR_UNSUPPORTED, "nn: filters must be 64..256 step 64"); return nullptr; }
  if (value_fc > 256 || value_fc < 1) { cz_fail(CZ_ERR_UNSUPPORTED, "nn: value_fc_size must be <= 256"); return nullptr; }
  if (n_nets < 1 || n_nets > 2) { cz_fail(CZ_ERR_ARG, "nn: 1 or 2 networks"); return nullptr; }
  if (pol_c < 0 || val_c < 0 || pol_c > 32 || val_c > 4 || (pol_c > 0 ? pol_c : 4) + (val_c > 0 ? val_c : 2) > kMaxHeadOut) {
    cz_fail(CZ_ERR_UNSUPPORTED, "nn: head widths up to 32 policy / 4 value channels"); return nullptr;
  }
  if (bytes < nn_workspace_bytes(filters, blocks, value_fc, max_batch, n_nets, pol_c, val_c)) { cz_fail(CZ_ERR_ARG, "nn: workspace too small"); return nullptr; }
  NnRuntime* r = new NnRuntime();
  set_heads(r, pol_c, val_c);
  r->heads_attr = false;
  r->filters = filters; r->blocks = blocks; r->value_fc = value_fc; r->max_batch = max_batch; r->n_nets = n_nets; r->cur = 0;
  r->stream = (cudaStream_t)stream; r->ready = false; r->launches = 0;
  r->in_planes = in_planes == 28 ? 28 : 14;
  // 0 = auto (fp32 skip stream for towers of 10 blocks and more, where fp16 rounding of the skip stream pushes the outputs
  // past 1e-3: value 1.1e-3 .. 1.5e-3 at 20 random-init blocks vs <= 6e-4 with fp32; policy 1.6e-3 vs 9.8e-4 on the
  // reference's trained 192x10 net), 1 = always, 2 = never
  r->fp32_skip = fp32_skip_mode == 1 || (fp32_skip_mode == 0 && blocks >= 10);
  { const char* e = getenv("CZ_FP32_SKIP"); if (e && e[0] == '1') r->fp32_skip = true; if (e && e[0] == '0') r->fp32_skip = false; }
  { const char* e = getenv("CZ_CLUSTER4"); g_cluster4 = (e && e[0] == '1') ? 1 : 0; }
  r->profile = false; r->ev_used = 0; r->prof_ms = 0; r->prof_flops = 0; r->prof_launches = 0; r->capturing = false;
  Carver cv{(uint8_t*)workspace, 0, bytes};
  layout(r, cv);
  const int c = filters;
  const long long rows = (long long)max_batch * 11;
  int rc = 0;
  r->board_pixels = (use_im2col() && use_pair_kernel()) ? 90 : 99;
  if (r->board_pixels == 90) {
    rc |= make_map_im2col(&r->imap_x, r->x, c, max_batch);
    rc |= make_map_im2col(&r->imap_t, r->t, c, max_batch);
    rc |= make_map_im2col(&r->imap_y, r->y, c, max_batch);
    rc |= make_map_tile32(&r->omap_x, r->x, c, (long long)max_batch * 90, false);
    rc |= make_map_tile32(&r->omap_t, r->t, c, (long long)max_batch * 90, false);
    rc |= make_map_tile32(&r->omap_y, r->y, c, (long long)max_batch * 90, false);
    rc |= make_map_tile32(&r->fmap_x32, r->x32, c, (long long)max_batch * 90, true);
    rc |= make_map_tile32(&r->fmap_y32, r->y32, c, (long long)max_batch * 90, true);
  }
  rc |= make_map_3d(&r->map_x, r->x, c, 9, rows, 9, 14);
  rc |= make_map_3d(&r->map_t, r->t, c, 9, rows, 9, 14);
  rc |= make_map_3d(&r->map_y, r->y, c, 9, rows, 9, 14);
  rc |= make_map_3d(&r->map_pf, r->pol_feat, 3 * r->pol_k1, 1, (long long)max_batch + 128, 1, 128);
  for (int net = n_nets - 1; net >= 0; --net) {
    select_net(r, net);
    rc |= make_map_2d(&r->map_wpol, r->w_pol, 3 * r->pol_k1, kPolN, 256);
    r->map_w.resize(2 * blocks);
    r->map_w_half.resize(2 * blocks);
    r->map_w_32.resize(2 * blocks);
    for (int i = 0; i < 2 * blocks; ++i) {
      rc |= make_map_2d(&r->map_w[i], r->w_conv + (size_t)i * 9 * c * c, c, 9LL * c, c);
      rc |= make_map_2d(&r->map_w_half[i], r->w_conv + (size_t)i * 9 * c * c, c, 9LL * c, c / 2);
      rc |= make_map_2d(&r->map_w_32[i], r->w_conv + (size_t)i * 9 * c * c, c, 9LL * c, 32);
    }
    store_net(r, net);
  }
  if (rc) { delete r; return nullptr; }
  // separator rows and padding must start as zeros
  cudaMemsetAsync(r->x, 0, (size_t)max_batch * 11 * 9 * c * 2, r->stream);
  cudaMemsetAsync(r->t, 0, (size_t)max_batch * 11 * 9 * c * 2, r->stream);
  cudaMemsetAsync(r->y, 0, (size_t)max_batch * 11 * 9 * c * 2, r->stream);
  cudaMemsetAsync(r->pol_feat, 0, ((size_t)max_batch + 128) * 3 * r->pol_k1 * 2, r->stream);
  return r;
}

void nn_destroy(NnRuntime* r) {
  if (!r) return;
  for (cudaEvent_t e : r->ev) cudaEventDestroy(e);
  delete r;
}
void nn_profile(NnRuntime* r, bool on) { if (r) { r->profile = on; } }
// Synchronises the stream. ms = device time spent in the residual-tower igemm launches since the last read.
int nn_profile_read(NnRuntime* r, double* ms, uint64_t* launches, double* flops) {
  if (!r) return cz_fail(CZ_ERR_STATE, "no network");
  CZ_CUDA(cudaStreamSynchronize(r->stream));
  prof_collect(r);
  if (ms) *ms = r->prof_ms;
  if (launches) *launches = r->prof_launches;
  if (flops) *flops = r->prof_flops;
  r->prof_ms = 0; r->prof_flops = 0; r->prof_launches = 0;
  return 0;
}
bool nn_ready(const NnRuntime* r) {
  if (!r) return false;
  for (int k = 0; k < r->n_nets; ++k)
    if (!(k == r->cur ? r->ready : r->nets[k].ready)) return false;
  return true;
}
uint64_t nn_launches(const NnRuntime* r) { return r ? r->launches : 0; }

// ---- weights -----------------------------------------------------------------------------------
struct WeightSet {
  const cz_tensor_desc* d; int n;
  // find "<layer prefix>...<'/'><weight>" ; Keras appends "-<k>-<f>" to conv layer names and ":0" to weights
  const cz_tensor_desc* find(const std::string& layer, const std::string& weight) const {
    for (int i = 0; i < n; ++i) {
      std::string nm = d[i].name ? d[i].name : "";
      const size_t slash = nm.find('/');
      if (slash == std::string::npos) continue;
      std::string l = nm.substr(0, slash), w = nm.substr(slash + 1);
      const size_t colon = w.find(':');
      if (colon != std::string::npos) w = w.substr(0, colon);
      const size_t s2 = w.find('/');            // "layer/layer/kernel" style
      if (s2 != std::string::npos) w = w.substr(s2 + 1);
      if (w != weight) continue;
      if (l == layer || (l.size() > layer.size() && l.compare(0, layer.size(), layer) == 0 && l[layer.size()] == '-')) return &d[i];
    }
    return nullptr;
  }
};

#define NEED(var, layer, weight, count)                                                                      \
  const cz_tensor_desc* var = ws.find(layer, weight);                                                        \
  if (!var || var->numel != (long long)(count))                                                              \
    return cz_fail(CZ_ERR_ARG, "cz_nn_set_weights: missing or mis-sized tensor %s/%s (want %lld)", std::string(layer).c_str(), weight, (long long)(count));

static int fold_bn(NnRuntime* r, const WeightSet& ws, const std::string& layer, int c, float* scale, float* shift) {
  NEED(g, layer, "gamma", c);
  NEED(b, layer, "beta", c);
  NEED(m, layer, "moving_mean", c);
  NEED(v, layer, "moving_variance", c);
  k_bn_fold<<<(c + 127) / 128, 128, 0, r->stream>>>((const float*)g->dev, (const float*)b->dev, (const float*)m->dev,
                                                     (const float*)v->dev, scale, shift, c);
  r->launches++;
  return 0;
}

int nn_set_weights(NnRuntime* r, int net, const cz_tensor_desc* descs, int n) {
  if (!r) return cz_fail(CZ_ERR_STATE, "cz_nn_set_weights: engine was created without a network (nn_filters = 0)");
  if (net < 0 || net >= r->n_nets) return cz_fail(CZ_ERR_ARG, "cz_nn_set_weights: network %d of %d", net, r->n_nets);
  select_net(r, net);
  WeightSet ws{descs, n};
  const int c = r->filters;
  cudaStream_t st = r->stream;
  float* scale = r->scratch;
  {
    NEED(k, "input_conv", "kernel", 25LL * r->in_planes * c);
    if (fold_bn(r, ws, "input_batchnorm", c, scale, r->shift_first)) return CZ_ERR_ARG;
    const long long nn = 25LL * r->in_planes * c;
    k_prep_hwio<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>((const float*)k->dev, scale, r->w_first, nn, c);
  }
  for (int i = 0; i < r->blocks; ++i) {
    for (int j = 0; j < 2; ++j) {
      const std::string conv = "res" + std::to_string(i + 1) + "_conv" + std::to_string(j + 1);
      const std::string bn = "res" + std::to_string(i + 1) + "_batchnorm" + std::to_string(j + 1);
      NEED(k, conv, "kernel", 9LL * c * c);
      const int li = 2 * i + j;
      if (fold_bn(r, ws, bn, c, scale, r->shift_conv + (size_t)li * c)) return CZ_ERR_ARG;
      const long long nn = 9LL * c * c;
      k_prep_conv3<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>((const float*)k->dev, scale, r->w_conv + (size_t)li * nn, c);
    }
  }
  {
    const int pc = r->pol_c, vc = r->val_c;
    NEED(kp, "policy_conv", "kernel", (long long)pc * c);
    if (fold_bn(r, ws, "policy_batchnorm", pc, scale, r->shifth)) return CZ_ERR_ARG;
    k_prep_1x1<<<(pc * c + 255) / 256, 256, 0, st>>>((const float*)kp->dev, scale, r->wh, c, pc);
    NEED(kv, "value_conv", "kernel", (long long)vc * c);
    float* scale_v = r->scratch + 2 * 256 + kMaxHeadOut;     // k_bn_fold of the policy head above may still be reading `scale`
    if (fold_bn(r, ws, "value_batchnorm", vc, scale_v, r->shifth + pc)) return CZ_ERR_ARG;
    k_prep_1x1<<<(vc * c + 255) / 256, 256, 0, st>>>((const float*)kv->dev, scale_v, r->wh + (size_t)pc * c, c, vc);
  }
  {
    NEED(k, "policy_out", "kernel", (long long)r->pol_c * 90 * kLabels);
    NEED(b, "policy_out", "bias", kLabels);
    const long long nn = (long long)kPolN * r->pol_k1;
    k_prep_policy<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>((const float*)k->dev, r->w_pol, r->pol_c * 90, r->pol_k1);
    k_copy_pad<<<(kPolN + 255) / 256, 256, 0, st>>>((const float*)b->dev, r->b_pol, kLabels, kPolN);
  }
  {
    const int h = r->value_fc;
    NEED(k1, "value_dense", "kernel", (long long)r->val_c * 90 * h);
    NEED(b1, "value_dense", "bias", h);
    NEED(k2, "value_out", "kernel", h);
    NEED(b2, "value_out", "bias", 1);
    CZ_CUDA(cudaMemcpyAsync(r->wv1, k1->dev, (size_t)r->val_c * 90 * h * 4, cudaMemcpyDeviceToDevice, st));
    CZ_CUDA(cudaMemcpyAsync(r->bv1, b1->dev, (size_t)h * 4, cudaMemcpyDeviceToDevice, st));
    CZ_CUDA(cudaMemcpyAsync(r->wv2, k2->dev, (size_t)h * 4, cudaMemcpyDeviceToDevice, st));
    CZ_CUDA(cudaMemcpyAsync(r->bv2, b2->dev, 4, cudaMemcpyDeviceToDevice, st));
  }
  r->launches += 6 + 2 * r->blocks;
  CZ_CUDA(cudaGetLastError());
  CZ_CUDA(cudaStreamSynchronize(st));
  r->ready = true;
  return 0;
}

// ---- forward -----------------------------------------------------------------------------------
// One pass over at most `n_max` positions; the ACTUAL batch size is the device integer *n_dev (every launch has a fixed
// shape sized for n_max, kernels read *n_dev and leave the rest untouched), so a search never has to tell the host how
// many leaves a wave produced.  Leaves logits [n][kPolN] + per-tile softmax statistics in r->logits / r->stats.
// The three parts are separate so that the search can capture them into three CUDA graphs and bracket the tower with events.
static int conv_first_threads(int c) {                    // (c/2) channel pairs x as many pixel groups as fit 256 threads
  const int pairs = c / 2;
  int g = 256 / pairs;
  if (g < 1) g = 1;
  while (pairs * g < 96) ++g;                             // phase 1 needs 90 threads
  return pairs * g;
}
static int fw_first(NnRuntime* r, const uint8_t* boards, int n, const int* n_dev) {
  const int c = r->filters;
  const bool s32 = r->board_pixels == 90 && r->fp32_skip;
  int slices = (2 * num_sms() + n - 1) / n;                // >= 2 CTAs per SM in flight; big batches: one CTA per position
  if (slices > 15) slices = 15;
  k_conv_first<<<dim3(n, slices), conv_first_threads(c), 0, r->stream>>>(boards, r->w_first, r->shift_first, r->x, s32 ? r->x32 : nullptr, c, r->board_pixels,
                                                            r->in_planes, (r->in_planes / 14) * CZ_BOARD_STRIDE, n_dev);
  r->launches++;
  CZ_CUDA(cudaGetLastError());
  return 0;
}
static int fw_tower(NnRuntime* r, int n, const int* n_dev) {
  const int c = r->filters;
  cudaStream_t st = r->stream;
  const bool dense = r->board_pixels == 90;
  const bool s32 = dense && r->fp32_skip;
  float *x32 = s32 ? r->x32 : nullptr, *y32 = s32 ? r->y32 : nullptr;
  CUtensorMap *ix = &r->imap_x, *iy = &r->imap_y;
  CUtensorMap *ox = &r->omap_x, *oy = &r->omap_y;           // fp16 tile maps of x / y
  CUtensorMap *fx = &r->fmap_x32, *fy = &r->fmap_y32;       // fp32 tile maps of x32 / y32
  __half *x = r->x, *y = r->y;
  CUtensorMap *mx = &r->map_x, *my = &r->map_y;
  const bool epi3 = dense && use_tma_epilogue();
  for (int i = 0; i < r->blocks; ++i) {
    const size_t wsz = (size_t)c;
    igemm::Args a1 = conv_args(n, c, r->shift_conv + (size_t)(2 * i) * wsz, nullptr, r->t, 1);
    igemm::Args a2 = conv_args(n, c, r->shift_conv + (size_t)(2 * i + 1) * wsz, x, y, 1);
    if (dense) {
      igemm::Args d1 = conv_args_dense(n, c, a1.bias, nullptr, r->t, 1);
      igemm::Args d2 = conv_args_dense(n, c, a2.bias, x, y, 1);
      d1.n_dev = n_dev; d1.rows_per_unit = 90; d2.n_dev = n_dev; d2.rows_per_unit = 90;
      d2.residual32 = x32; d2.out32 = y32;
      { float* t32 = x32; x32 = y32; y32 = t32; }
      // conv1: x -> t (no skip);  conv2: t (+ skip x or x32) -> y (+ y32)
      if (epi3 && use_n_split(n, c)) {
        if (launch_igemm3_split(c, *ix, r->map_w_32[2 * i], r->omap_t, r->omap_t, r->omap_t, d1, 0, false, st)) return CZ_ERR_CUDA;
        if (launch_igemm3_split(c, r->imap_t, r->map_w_32[2 * i + 1], *oy, s32 ? *fx : *ox, *fy, d2, s32 ? 2 : 1, s32, st)) return CZ_ERR_CUDA;
      } else {
      if (epi3 && use_tma_epilogue_for(c, false)) {
        if (launch_igemm3(c, *ix, r->map_w_half[2 * i], r->omap_t, r->omap_t, r->omap_t, d1, 0, false, st)) return CZ_ERR_CUDA;
      } else if (launch_igemm2(c, *ix, r->map_w_half[2 * i], d1, st)) return CZ_ERR_CUDA;
      if (epi3 && use_tma_epilogue_for(c, s32)) {
        if (launch_igemm3(c, r->imap_t, r->map_w_half[2 * i + 1], *oy, s32 ? *fx : *ox, *fy, d2, s32 ? 2 : 1, s32, st)) return CZ_ERR_CUDA;
      } else if (launch_igemm2(c, r->imap_t, r->map_w_half[2 * i + 1], d2, st)) return CZ_ERR_CUDA;
      }
      CUtensorMap* ti = ix; ix = iy; iy = ti;
      ti = ox; ox = oy; oy = ti;
      ti = fx; fx = fy; fy = ti;
    } else {
      // strip layout (CZ_CONV_STRIP=1 / CZ_IGEMM_1CTA=1 A-B baselines): host-known batch only
      if (launch_igemm(c, *mx, r->map_w[2 * i], a1, st)) return CZ_ERR_CUDA;
      if (launch_igemm(c, r->map_t, r->map_w[2 * i + 1], a2, st)) return CZ_ERR_CUDA;
    }
    r->launches += 2;
    __half* tx = x; x = y; y = tx;
    CUtensorMap* tm = mx; mx = my; my = tm;
  }
  return 0;
}
static int fw_heads(NnRuntime* r, int n, const int* n_dev, float* value) {
  const int c = r->filters;
  const bool s32 = r->board_pixels == 90 && r->fp32_skip;
  const bool odd = (r->blocks & 1) != 0;                   // the tower ping-pongs x <-> y once per block
  const __half* x = odd ? r->y : r->x;
  const float* x32 = s32 ? (odd ? r->y32 : r->x32) : nullptr;
  const size_t hsm = heads_smem_bytes(c, r->pol_c + r->val_c);
  if (hsm > 48 * 1024 && !r->heads_attr) {
    CZ_CUDA(cudaFuncSetAttribute(k_heads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
    r->heads_attr = true;
  }
  const int hp = n <= 2 * num_sms() ? 1 : kHeadPos;       // small batches: a block per position (same arithmetic per position)
  k_heads<<<(n + hp - 1) / hp, 256, hsm, r->stream>>>(x, x32, c, n_dev, r->board_pixels, r->pol_c, r->val_c, r->pol_k1, r->wh, r->shifth,
                                                      r->wv1, r->bv1, r->wv2, r->bv2, r->value_fc, r->pol_feat, value, hp);
  igemm::Args ap = dense_args(n, kLabels, kPolN, 3 * r->pol_k1, 256, r->b_pol, r->logits, kPolN);
  ap.n_dev = n_dev; ap.rows_per_unit = 1; ap.row_stats = r->stats;
  if (launch_igemm(256, r->map_pf, r->map_wpol, ap, r->stream)) return CZ_ERR_CUDA;
  r->launches += 2;
  CZ_CUDA(cudaGetLastError());
  return 0;
}
// events around the tower of one forward (bench.py roofline); flops = algorithmic flops of the bracketed launches, or < 0 when
// only the device knows the batch size (the reader then takes the positions from the search's device counter)
void nn_prof_begin(NnRuntime* r, double flops) {
  r->prof_open = (size_t)-1;
  if (!r->profile) return;
  if (r->ev_used + 2 > 4096) { cudaStreamSynchronize(r->stream); prof_collect(r); }
  while (r->ev.size() < r->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); r->ev.push_back(e); }
  r->prof_open = r->ev_used; r->ev_used += 2;
  if (r->ev_flops.size() < r->ev_used / 2) r->ev_flops.resize(r->ev_used / 2);
  r->ev_flops[r->prof_open / 2] = flops > 0 ? flops : 0.0;
  cudaEventRecord(r->ev[r->prof_open], r->stream);
}
void nn_prof_end(NnRuntime* r) {
  if (r->prof_open != (size_t)-1) cudaEventRecord(r->ev[r->prof_open + 1], r->stream);
  r->prof_open = (size_t)-1;
}
static int forward_tower(NnRuntime* r, const uint8_t* boards, int n_max, const int* n_dev, float* value) {
  int rc = fw_first(r, boards, n_max, n_dev);
  if (rc) return rc;
  nn_prof_begin(r, 2.0 * 90.0 * 9.0 * r->filters * r->filters * (double)n_max * 2.0 * r->blocks);
  rc = fw_tower(r, n_max, n_dev);
  nn_prof_end(r);
  if (rc) return rc;
  return fw_heads(r, n_max, n_dev, value);
}

// host-known batch: the reference-facing predict_on_batch (api.py:62-64) -> the full softmax vector
static int forward_chunk(NnRuntime* r, const uint8_t* boards, int n, float* policy, float* value) {
  k_set_int<<<1, 1, 0, r->stream>>>(r->n_scalar, n);
  if (r->board_pixels != 90) {                      // strip-layout baselines launch exact shapes
    // (the device-side batch size is still honoured by the first conv, the heads and the policy GEMM)
  }
  const int rc = forward_tower(r, boards, n, r->n_scalar, value);
  if (rc) return rc;
  k_softmax<<<n, 256, 0, r->stream>>>(r->logits, kPolN, r->stats, kPolN / 256, policy);
  r->launches += 2;
  CZ_CUDA(cudaGetLastError());
  return 0;
}

int nn_forward_boards(NnRuntime* r, int net, const uint8_t* boards, int batch, float* policy, float* value) {
  if (!r || net < 0 || net >= r->n_nets) return cz_fail(CZ_ERR_STATE, "no such network");
  select_net(r, net);
  if (!r->ready) return cz_fail(CZ_ERR_STATE, "network weights not set (cz_nn_set_weights)");
  for (int off = 0; off < batch; off += r->max_batch) {
    const int n = batch - off < r->max_batch ? batch - off : r->max_batch;
    const int rc = forward_chunk(r, boards + (size_t)off * (r->in_planes / 14) * CZ_BOARD_STRIDE, n, policy + (size_t)off * kLabels, value + off);
    if (rc) return rc;
  }
  return 0;
}

int nn_forward_planes(NnRuntime* r, int net, const float* planes, int batch, float* policy, float* value) {
  if (!r || net < 0 || net >= r->n_nets) return cz_fail(CZ_ERR_STATE, "no such network");
  select_net(r, net);
  if (!r->ready) return cz_fail(CZ_ERR_STATE, "network weights not set (cz_nn_set_weights)");
  for (int off = 0; off < batch; off += r->max_batch) {
    const int n = batch - off < r->max_batch ? batch - off : r->max_batch;
    k_planes_to_boards<<<n, 96, 0, r->stream>>>(planes + (size_t)off * r->in_planes * 90, r->boards_tmp, n, r->in_planes);
    r->launches++;
    const int rc = forward_chunk(r, r->boards_tmp, n, policy + (size_t)off * kLabels, value + off);
    if (rc) return rc;
  }
  return 0;
}

// The search's evaluation step: up to n_max leaves (actual count *n_dev), boards + legal-move labels in, value [n] and the
// softmax probabilities of the legal moves [n][CZ_MAX_MOVES] out.  Fixed launch shapes: safe to capture into a CUDA graph.
int nn_forward_leaves(NnRuntime* r, int net, int part, const uint8_t* boards, int n_max, const int* n_dev, const int16_t* labels,
                      const int32_t* label_counts, float* legal_p, float* value) {
  if (!r || net < 0 || net >= r->n_nets) return cz_fail(CZ_ERR_STATE, "no such network");
  if (n_max > r->max_batch) return cz_fail(CZ_ERR_ARG, "nn_forward_leaves: %d leaves > max batch %d", n_max, r->max_batch);
  if (r->board_pixels != 90) return cz_fail(CZ_ERR_UNSUPPORTED, "nn_forward_leaves needs the dense (im2col) layout");
  select_net(r, net);
  if (!r->ready) return cz_fail(CZ_ERR_STATE, "network weights not set (cz_nn_set_weights)");
  int rc = 0;
  if (part & 1) rc = fw_first(r, boards, n_max, n_dev);
  if (!rc && (part & 2)) rc = fw_tower(r, n_max, n_dev);
  if (!rc && (part & 4)) rc = fw_heads(r, n_max, n_dev, value);
  if (rc) return rc;
  if (part & 4) k_legal_priors<<<(n_max + 3) / 4, 128, 0, r->stream>>>(r->logits, kPolN, r->stats, kPolN / 256, labels, label_counts, n_dev, legal_p);
  if (part & 4) r->launches++;
  CZ_CUDA(cudaGetLastError());
  return 0;
}
void nn_set_capturing(NnRuntime* r, bool on) { if (r) r->capturing = on; }
bool nn_profiling(const NnRuntime* r) { return r && r->profile; }
void nn_set_stream(NnRuntime* r, void* stream) { if (r) r->stream = (cudaStream_t)stream; }
int nn_launches_per_forward(const NnRuntime* r) { return r ? 1 + 2 * r->blocks + 3 : 0; }
double nn_tower_flops_per_position(const NnRuntime* r) { return r ? 2.0 * 90.0 * 9.0 * r->filters * r->filters * 2.0 * r->blocks : 0.0; }

}  // namespace cznn

// ------------------------------------------------------------------------------------------------
// Building blocks exported for parity tests and profiling (not part of the reference-facing surface).
extern "C" {

// 3x3 "same" convolution on strip-layout activations: out = relu?(conv(in, w) + bias (+ residual)).
//   act_in/out/residual: fp16 [n_boards*11][9][c] (separator rows of act_in must be zero)
//   w: fp16 [9][c_out = c][c_in = c] ; bias f32 [c]
int cz_igemm_conv3x3(const void* act_in, const void* w, const float* bias, const void* residual, void* act_out,
                     int n_boards, int c, int relu, void* stream) {
  using namespace cznn;
  if (c % 64 || c < 64 || c > 256 || n_boards <= 0) return cz_fail(CZ_ERR_ARG, "cz_igemm_conv3x3: bad shape");
  CUtensorMap ma, mb;
  if (make_map_3d(&ma, act_in, c, 9, (long long)n_boards * 11, 9, 14)) return CZ_ERR_CUDA;
  igemm::Args a = conv_args(n_boards, c, bias, (const __half*)residual, (__half*)act_out, relu);
  if (make_map_2d(&mb, w, c, 9LL * c, c)) return CZ_ERR_CUDA;
  return launch_igemm(c, ma, mb, a, (cudaStream_t)stream);
}

// Same convolution on DENSE activations fp16 [n_boards][10][9][c] through the im2col TMA path (CTA-pair kernel).
int cz_igemm_conv3x3_dense(const void* act_in, const void* w, const float* bias, const void* residual, void* act_out,
                           int n_boards, int c, int relu, void* stream) {
  using namespace cznn;
  if (c % 64 || c < 64 || c > 256 || n_boards <= 0) return cz_fail(CZ_ERR_ARG, "cz_igemm_conv3x3_dense: bad shape");
  CUtensorMap ma, mb;
  if (make_map_im2col(&ma, act_in, c, n_boards)) return CZ_ERR_CUDA;
  if (make_map_2d(&mb, w, c, 9LL * c, c / 2)) return CZ_ERR_CUDA;
  igemm::Args a = conv_args_dense(n_boards, c, bias, (const __half*)residual, (__half*)act_out, relu);
  if (use_tma_epilogue()) {
    CUtensorMap mo, ms;
    if (make_map_tile32(&mo, act_out, c, (long long)n_boards * 90, false)) return CZ_ERR_CUDA;
    if (make_map_tile32(&ms, residual ? residual : act_out, c, (long long)n_boards * 90, false)) return CZ_ERR_CUDA;
    return launch_igemm3(c, ma, mb, mo, ms, mo, a, residual ? 1 : 0, false, (cudaStream_t)stream);
  }
  return launch_igemm2(c, ma, mb, a, (cudaStream_t)stream);
}

// out[m][n] = sum_k a[m][k] * w[n][k] + bias[n]; a fp16 [m_alloc >= ceil128(m)][k], w fp16 [n_pad][k], k % 64 == 0,
// n_pad % n_tile == 0, out f32 [m][ldo].
int cz_igemm_dense(const void* a_dev, const void* w_dev, const float* bias, float* out, int m, int n_valid, int n_pad,
                   int k, int n_tile, int ldo, void* stream) {
  using namespace cznn;
  if (k % 64 || n_pad % n_tile || m <= 0) return cz_fail(CZ_ERR_ARG, "cz_igemm_dense: bad shape");
  CUtensorMap ma, mb;
  if (make_map_3d(&ma, a_dev, k, 1, (long long)m, 1, 128)) return CZ_ERR_CUDA;
  if (make_map_2d(&mb, w_dev, k, n_pad, n_tile)) return CZ_ERR_CUDA;
  igemm::Args a = dense_args(m, n_valid, n_pad, k, n_tile, bias, out, ldo);
  return launch_igemm(n_tile, ma, mb, a, (cudaStream_t)stream);
}

}  // extern "C"
