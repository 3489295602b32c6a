// pgbart_predict.cu — C ABI of the history-based posterior prediction (rows N1 / N4); kernels in pgbart_predict.cuh.
#include "pgbart_predict.cuh"

#include <stdio.h>

extern "C" void bk_set_error_message(const char* msg);   // pgbart_b200.cu: thread-local message behind bk_last_error()

namespace {
struct DevGuard {
  int prev = -1;
  cudaError_t err;
  explicit DevGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
int fail(const char* what, cudaError_t e) {
  char buf[256];
  snprintf(buf, sizeof(buf), "CUDA error %s in %s", cudaGetErrorString(e), what);
  bk_set_error_message(buf);
  return BK_ERR_CUDA;
}
}  // namespace

extern "C" {

int bk_predict_history(int device, void* stream, const bk_node* nodes_dev, const int32_t* ver_off_dev, const int32_t* ver_tbl_dev,
                       int n_trees, int max_forest_nodes, const float* X_dev, int n, int n_cols, const int32_t* sel_dev, int n_sel,
                       int sel_per_mask, const uint8_t* excluded_masks_dev, int n_masks, const int32_t* split_rules_dev,
                       const float* leaf_values_dev, int n_values, float* out_dev, int32_t* err_dev) {
  if (!nodes_dev || !ver_off_dev || !ver_tbl_dev || !X_dev || !sel_dev || !out_dev || !err_dev || n < 0 || n_sel < 0 || n_trees < 1 ||
      n_cols < 1 || n_masks < 0 || (n_masks > 0 && !excluded_masks_dev)) {
    bk_set_error_message("bk_predict_history: bad argument");
    return BK_ERR_ARG;
  }
  if (n == 0 || n_sel == 0) return BK_OK;
  if (n_sel > 65535 || n_masks > 65535) { bk_set_error_message("bk_predict_history: at most 65535 forests / masks per call"); return BK_ERR_ARG; }
  DevGuard g(device);
  if (g.err != cudaSuccess) return fail("cudaSetDevice", g.err);
  PredictArgs A;
  A.nodes = nodes_dev; A.ver_off = ver_off_dev; A.ver_tbl = ver_tbl_dev; A.m = n_trees; A.X = X_dev; A.n = n; A.p = n_cols;
  A.sel = sel_dev; A.n_sel = n_sel; A.sel_stride = (sel_per_mask && n_masks > 0) ? n_sel : 0; A.excl = excluded_masks_dev; A.n_masks = n_masks; A.rules = split_rules_dev; A.out = out_dev;
  A.err = err_dev;
  A.vals = leaf_values_dev; A.K = leaf_values_dev ? n_values : 1;
  if (leaf_values_dev && (n_values < 1 || n_values > BK_MAX_OUTPUTS)) { bk_set_error_message("bk_predict_history: n_values out of range"); return BK_ERR_ARG; }
  int cap = max_forest_nodes < 0 ? 0 : max_forest_nodes;
  if (cap > BKP_SMEM_NODES) cap = BKP_SMEM_NODES;
  A.smem_nodes = cap;
  const int m_s = n_trees <= BKP_MAX_TREES_SMEM ? n_trees : 0;
  const size_t smem = (((size_t)(m_s + 1) * 4 + 15) & ~(size_t)15) + (size_t)cap * sizeof(bk_node);
  const dim3 grid((unsigned)((n + BKP_TILE - 1) / BKP_TILE), (unsigned)n_sel, (unsigned)(n_masks > 0 ? n_masks : 1));
  cudaError_t e;
  if (n_masks > 0) {
    e = cudaFuncSetAttribute(pgbart_predict_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute", e);
    pgbart_predict_hist_kernel<true><<<grid, BKP_THREADS, smem, (cudaStream_t)stream>>>(A);
  } else {
    e = cudaFuncSetAttribute(pgbart_predict_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute", e);
    pgbart_predict_hist_kernel<false><<<grid, BKP_THREADS, smem, (cudaStream_t)stream>>>(A);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail("pgbart_predict_hist_kernel launch", e);
  return BK_OK;
}

int bk_pearson_r2(int device, void* stream, const float* a_dev, const float* b_dev, int len, int n_samples, int n_subsets,
                  double* out_dev) {
  if (!a_dev || !b_dev || !out_dev || len < 1 || n_samples < 1 || n_subsets < 1 || n_subsets > 65535) {
    bk_set_error_message("bk_pearson_r2: bad argument");
    return BK_ERR_ARG;
  }
  DevGuard g(device);
  if (g.err != cudaSuccess) return fail("cudaSetDevice", g.err);
  pgbart_pearson_r2_kernel<<<dim3((unsigned)n_samples, (unsigned)n_subsets), 256, 0, (cudaStream_t)stream>>>(a_dev, b_dev, len, n_samples, out_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("pgbart_pearson_r2_kernel launch", e);
  return BK_OK;
}

}  // extern "C"
