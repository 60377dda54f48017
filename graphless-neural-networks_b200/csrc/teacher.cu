// Teacher forward (path A): layer planning for SAGE("gcn") / GCN eval-mode inference on top of the
// aggregation and projection kernels, plus the host-buffer entry point.
//
// Per layer the cheaper order is chosen (rows are gathered at min(d_in, d_out) width):
//   aggregate-first : T = agg(H)            ; Y = epi(T W + b)         (epilogue in the GEMM)
//   project-first   : Z = H W               ; Y = epi(agg(Z) + b)      (epilogue in the gather)
// Narrow outputs (47 classes, 7 classes) are padded to a multiple of 4 columns inside the
// workspace so that every gathered row is 16-byte aligned; the pad is stripped at the boundary.
#include <vector>

#include <algorithm>

#include "common.cuh"

namespace glnn {

static inline int pad4(int d) { return (d + 3) / 4 * 4; }
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Plan {
  int64_t n;
  int dmax;             // widest padded activation
  int64_t buf_floats;   // per activation buffer
  int64_t pad_floats;   // padded weight / epilogue scratch
  int64_t total_bytes;
};

static Plan make_plan(int64_t n, const glnn_gnn_layer* layers, int L) {
  Plan p{};
  p.n = n;
  p.dmax = 4;
  p.pad_floats = 0;
  for (int l = 0; l < L; ++l) {
    p.dmax = max(p.dmax, max(pad4(layers[l].d_in), pad4(layers[l].d_out)));
    const int dpad = pad4(layers[l].d_out);
    p.pad_floats += align_up(static_cast<int64_t>(dpad) * layers[l].d_in, 64) + 3 * align_up(dpad, 64);
  }
  p.buf_floats = align_up(n * p.dmax, 64);
  p.total_bytes = (3 * p.buf_floats + p.pad_floats) * static_cast<int64_t>(sizeof(float));
  return p;
}

struct Padded {
  const float* w;
  const float* bias;
  const float* scale;
  const float* shift;
};

// Pads the weight and the epilogue vectors of one layer to dpad output columns (zero fill).
// w_rows_are_out: SAGE weight [d_out, d_in] (pad rows); else GCN weight [d_in, d_out] (pad columns).
static int pad_layer(const glnn_gnn_layer& ly, bool w_rows_are_out, float*& scratch, Padded* out,
                     cudaStream_t st) {
  const int dpad = pad4(ly.d_out);
  out->w = ly.weight;
  out->bias = ly.bias;
  out->scale = ly.bn_scale;
  out->shift = ly.bn_shift;
  if (dpad == ly.d_out) return 0;
  float* w = scratch;
  scratch += align_up(static_cast<int64_t>(dpad) * ly.d_in, 64);
  GLNN_CUDA_OK(cudaMemsetAsync(w, 0, sizeof(float) * dpad * ly.d_in, st));
  if (w_rows_are_out) {
    GLNN_CUDA_OK(cudaMemcpyAsync(w, ly.weight, sizeof(float) * ly.d_out * ly.d_in,
                                 cudaMemcpyDeviceToDevice, st));
  } else {
    GLNN_CUDA_OK(cudaMemcpy2DAsync(w, sizeof(float) * dpad, ly.weight, sizeof(float) * ly.d_out,
                                   sizeof(float) * ly.d_out, ly.d_in, cudaMemcpyDeviceToDevice, st));
  }
  out->w = w;
  const float* src[3] = {ly.bias, ly.bn_scale, ly.bn_shift};
  const float** dst[3] = {&out->bias, &out->scale, &out->shift};
  for (int i = 0; i < 3; ++i) {
    float* v = scratch;
    scratch += align_up(dpad, 64);
    if (!src[i]) continue;
    GLNN_CUDA_OK(cudaMemsetAsync(v, 0, sizeof(float) * dpad, st));
    GLNN_CUDA_OK(cudaMemcpyAsync(v, src[i], sizeof(float) * ly.d_out, cudaMemcpyDeviceToDevice, st));
    *dst[i] = v;
  }
  return 0;
}

static int check_common(const void* indptr, const int32_t* indices, int64_t n, const float* X,
                        int64_t ldx, const glnn_gnn_layer* layers, int L, float* out, int64_t ldo,
                        void* ws, int64_t ws_bytes) {
  GLNN_REQUIRE(n >= 0 && L >= 1, GLNN_ERR_ARG, "gnn_forward: bad n / num_layers");
  GLNN_REQUIRE(indptr && X && layers && out && ws, GLNN_ERR_ARG, "gnn_forward: null pointer");
  GLNN_REQUIRE(ldx >= layers[0].d_in, GLNN_ERR_SHAPE, "gnn_forward: ldx < feat_dim");
  GLNN_REQUIRE(ldo >= layers[L - 1].d_out, GLNN_ERR_SHAPE, "gnn_forward: ldo < label_dim");
  for (int l = 0; l < L; ++l) {
    GLNN_REQUIRE(layers[l].weight && layers[l].bias, GLNN_ERR_ARG, "gnn_forward: layer %d null", l);
    GLNN_REQUIRE(layers[l].d_in > 0 && layers[l].d_out > 0, GLNN_ERR_SHAPE, "layer %d dims", l);
    GLNN_REQUIRE(l == 0 || layers[l].d_in == layers[l - 1].d_out, GLNN_ERR_SHAPE,
                 "gnn_forward: layer %d d_in does not chain", l);
    GLNN_REQUIRE((layers[l].bn_scale == nullptr) == (layers[l].bn_shift == nullptr), GLNN_ERR_ARG,
                 "gnn_forward: bn_scale/bn_shift must be given together");
  }
  const Plan p = make_plan(n, layers, L);
  GLNN_REQUIRE(ws_bytes >= p.total_bytes, GLNN_ERR_WORKSPACE,
               "gnn_forward: workspace %lld < required %lld bytes", (long long)ws_bytes,
               (long long)p.total_bytes);
  GLNN_REQUIRE(aligned16(ws), GLNN_ERR_ALIGN, "gnn_forward: workspace must be 16-byte aligned");
  (void)indices;
  return 0;
}

static int finish(const float* H, int64_t ldh, int64_t n, int c, float* out, int64_t ldo,
                  int log_softmax, cudaStream_t st) {
  if (log_softmax) return glnn_log_softmax_f32(H, ldh, out, ldo, n, c, st);
  GLNN_CUDA_OK(cudaMemcpy2DAsync(out, sizeof(float) * ldo, H, sizeof(float) * ldh, sizeof(float) * c,
                                 n, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // namespace glnn

extern "C" int64_t glnn_gnn_forward_workspace_bytes(int64_t n, const glnn_gnn_layer* layers,
                                                    int num_layers) {
  if (n < 0 || !layers || num_layers < 1) return -1;
  return glnn::make_plan(n, layers, num_layers).total_bytes;
}

extern "C" int glnn_sage_forward(const void* indptr, int indptr64, const int32_t* indices, int64_t n,
                                 const float* X, int64_t ldx, const glnn_gnn_layer* layers,
                                 int num_layers, float* out, int64_t ldo, int log_softmax,
                                 void* workspace, int64_t workspace_bytes, glnn_stream_t stream) {
  using namespace glnn;
  const int L = num_layers;
  int rc = check_common(indptr, indices, n, X, ldx, layers, L, out, ldo, workspace, workspace_bytes);
  if (rc != 0) return rc;
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(n, layers, L);
  float* buf[3];
  for (int i = 0; i < 3; ++i) buf[i] = static_cast<float*>(workspace) + i * p.buf_floats;
  float* scratch = static_cast<float*>(workspace) + 3 * p.buf_floats;

  const float* H = X;
  int64_t ldh = ldx;
  int hb = -1;  // buffer index holding H (-1 = caller's X)
  for (int l = 0; l < L; ++l) {
    const glnn_gnn_layer& ly = layers[l];
    const bool last = (l == L - 1);
    const int dpad = pad4(ly.d_out);
    const int relu = last ? 0 : 1;
    int ib = 0;
    while (ib == hb) ++ib;
    int ob = 0;
    while (ob == hb || ob == ib) ++ob;
    if (dpad < ly.d_in) {  // project first
      Padded pd;
      rc = pad_layer(ly, true, scratch, &pd, st);
      if (rc != 0) return rc;
      rc = glnn_gemm_f32(H, ldh, 0, pd.w, ly.d_in, 1, buf[ib], dpad, n, dpad, ly.d_in, nullptr,
                         nullptr, nullptr, nullptr, 0, 0, st);
      if (rc != 0) return rc;
      rc = glnn_spmm_csr_f32(indptr, indptr64, indices, buf[ib], dpad, buf[ob], dpad, n, n, dpad, 1,
                             1, nullptr, nullptr, pd.bias, pd.scale, pd.shift, relu, st);
      if (rc != 0) return rc;
    } else {  // aggregate first (DGL 0.6.1 order)
      const int ldt = pad4(ly.d_in);
      rc = glnn_spmm_csr_f32(indptr, indptr64, indices, H, ldh, buf[ib], ldt, n, n, ly.d_in, 1, 1,
                             nullptr, nullptr, nullptr, nullptr, nullptr, 0, st);
      if (rc != 0) return rc;
      rc = glnn_gemm_f32(buf[ib], ldt, 0, ly.weight, ly.d_in, 1, buf[ob], dpad, n, ly.d_out, ly.d_in,
                         nullptr, ly.bias, ly.bn_scale, ly.bn_shift, relu, 0, st);
      if (rc != 0) return rc;
    }
    H = buf[ob];
    ldh = dpad;
    hb = ob;
  }
  return finish(H, ldh, n, layers[L - 1].d_out, out, ldo, log_softmax, st);
}

extern "C" int glnn_gcn_forward(const void* indptr, int indptr64, const int32_t* indices, int64_t n,
                                const float* src_norm, const float* dst_norm, const float* X,
                                int64_t ldx, const glnn_gnn_layer* layers, int num_layers, float* out,
                                int64_t ldo, int log_softmax, void* workspace,
                                int64_t workspace_bytes, glnn_stream_t stream) {
  using namespace glnn;
  const int L = num_layers;
  int rc = check_common(indptr, indices, n, X, ldx, layers, L, out, ldo, workspace, workspace_bytes);
  if (rc != 0) return rc;
  GLNN_REQUIRE(src_norm && dst_norm, GLNN_ERR_ARG, "gcn_forward: null degree-norm vector");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(n, layers, L);
  float* buf[3];
  for (int i = 0; i < 3; ++i) buf[i] = static_cast<float*>(workspace) + i * p.buf_floats;
  float* scratch = static_cast<float*>(workspace) + 3 * p.buf_floats;

  const float* H = X;
  int64_t ldh = ldx;
  int hb = -1;
  for (int l = 0; l < L; ++l) {
    const glnn_gnn_layer& ly = layers[l];
    const bool last = (l == L - 1);
    const int dpad = pad4(ly.d_out);
    const int relu = last ? 0 : 2;  // GraphConv applies the activation before the norm layer
    int ib = 0;
    while (ib == hb) ++ib;
    int ob = 0;
    while (ob == hb || ob == ib) ++ob;
    if (ly.d_in > ly.d_out) {  // DGL: multiply by W first when it shrinks the rows
      Padded pd;
      rc = pad_layer(ly, false, scratch, &pd, st);
      if (rc != 0) return rc;
      rc = glnn_gemm_f32(H, ldh, 0, pd.w, dpad, 0, buf[ib], dpad, n, dpad, ly.d_in, src_norm, nullptr,
                         nullptr, nullptr, 0, 0, st);
      if (rc != 0) return rc;
      rc = glnn_spmm_csr_f32(indptr, indptr64, indices, buf[ib], dpad, buf[ob], dpad, n, n, dpad, 0,
                             0, nullptr, dst_norm, pd.bias, pd.scale, pd.shift, relu, st);
      if (rc != 0) return rc;
    } else {
      const int ldt = pad4(ly.d_in);
      rc = glnn_spmm_csr_f32(indptr, indptr64, indices, H, ldh, buf[ib], ldt, n, n, ly.d_in, 0, 0,
                             src_norm, nullptr, nullptr, nullptr, nullptr, 0, st);
      if (rc != 0) return rc;
      rc = glnn_gemm_f32(buf[ib], ldt, 0, ly.weight, ly.d_out, 0, buf[ob], dpad, n, ly.d_out, ly.d_in,
                         dst_norm, ly.bias, ly.bn_scale, ly.bn_shift, relu, 0, st);
      if (rc != 0) return rc;
    }
    H = buf[ob];
    ldh = dpad;
    hb = ob;
  }
  return finish(H, ldh, n, layers[L - 1].d_out, out, ldo, log_softmax, st);
}

// ------------------------------------------------------------------------------------------------
// host-buffer entry point
// ------------------------------------------------------------------------------------------------
namespace glnn {
struct DevFree {
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  ~DevFree() {
    for (void* p : ptrs) cudaFree(p);
    if (st) cudaStreamDestroy(st);
  }
  template <typename T>
  int alloc(T** p, int64_t count) {
    void* q = nullptr;
    GLNN_CUDA_OK(cudaMalloc(&q, static_cast<size_t>(std::max<int64_t>(count, 1)) * sizeof(T)));
    ptrs.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
  }
};
}  // namespace glnn

extern "C" int glnn_sage_inference_host(const int64_t* indptr_host, const int32_t* indices_host,
                                        int64_t n, const float* feats_host,
                                        const glnn_sage_layer_host* layers, int num_layers,
                                        float bn_eps, float* out_logprob_host) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0 && num_layers >= 1 && num_layers <= 64, GLNN_ERR_ARG, "sage_host: bad sizes");
  GLNN_REQUIRE(indptr_host && feats_host && layers && out_logprob_host, GLNN_ERR_ARG,
               "sage_host: null pointer");
  if (n == 0) return 0;
  int rc = glnn_device_info(nullptr, nullptr, nullptr);
  if (rc != 0) return rc;
  const int64_t nnz = indptr_host[n];
  GLNN_REQUIRE(nnz >= 0 && (nnz == 0 || indices_host), GLNN_ERR_ARG, "sage_host: bad indptr/indices");
  DevFree g;
  GLNN_CUDA_OK(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
  int64_t* d_indptr;
  int32_t* d_indices;
  float *d_x, *d_out;
  const int f = layers[0].d_in, c = layers[num_layers - 1].d_out;
  if ((rc = g.alloc(&d_indptr, n + 1))) return rc;
  if ((rc = g.alloc(&d_indices, nnz))) return rc;
  if ((rc = g.alloc(&d_x, n * f))) return rc;
  if ((rc = g.alloc(&d_out, n * c))) return rc;
  GLNN_CUDA_OK(cudaMemcpyAsync(d_indptr, indptr_host, sizeof(int64_t) * (n + 1),
                               cudaMemcpyHostToDevice, g.st));
  GLNN_CUDA_OK(cudaMemcpyAsync(d_indices, indices_host, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice,
                               g.st));
  GLNN_CUDA_OK(cudaMemcpyAsync(d_x, feats_host, sizeof(float) * n * f, cudaMemcpyHostToDevice, g.st));
  std::vector<glnn_gnn_layer> dl(num_layers);
  for (int l = 0; l < num_layers; ++l) {
    const glnn_sage_layer_host& h = layers[l];
    GLNN_REQUIRE(h.weight && h.bias && h.d_in > 0 && h.d_out > 0, GLNN_ERR_ARG, "sage_host: layer %d",
                 l);
    float *w, *b;
    if ((rc = g.alloc(&w, static_cast<int64_t>(h.d_in) * h.d_out))) return rc;
    if ((rc = g.alloc(&b, h.d_out))) return rc;
    GLNN_CUDA_OK(cudaMemcpyAsync(w, h.weight, sizeof(float) * h.d_in * h.d_out,
                                 cudaMemcpyHostToDevice, g.st));
    GLNN_CUDA_OK(cudaMemcpyAsync(b, h.bias, sizeof(float) * h.d_out, cudaMemcpyHostToDevice, g.st));
    dl[l] = glnn_gnn_layer{w, b, nullptr, nullptr, h.d_in, h.d_out};
    if (h.bn_gamma) {
      GLNN_REQUIRE(h.bn_beta && h.bn_mean && h.bn_var, GLNN_ERR_ARG, "sage_host: layer %d BN", l);
      float* v;
      if ((rc = g.alloc(&v, 6LL * h.d_out))) return rc;
      const float* src[4] = {h.bn_gamma, h.bn_beta, h.bn_mean, h.bn_var};
      for (int i = 0; i < 4; ++i)
        GLNN_CUDA_OK(cudaMemcpyAsync(v + i * h.d_out, src[i], sizeof(float) * h.d_out,
                                     cudaMemcpyHostToDevice, g.st));
      rc = glnn_bn_fold_f32(v, v + h.d_out, v + 2 * h.d_out, v + 3 * h.d_out, bn_eps, v + 4 * h.d_out,
                            v + 5 * h.d_out, h.d_out, g.st);
      if (rc != 0) return rc;
      dl[l].bn_scale = v + 4 * h.d_out;
      dl[l].bn_shift = v + 5 * h.d_out;
    }
  }
  const int64_t ws_bytes = glnn_gnn_forward_workspace_bytes(n, dl.data(), num_layers);
  char* ws;
  if ((rc = g.alloc(&ws, ws_bytes))) return rc;
  rc = glnn_sage_forward(d_indptr, 1, d_indices, n, d_x, f, dl.data(), num_layers, d_out, c, 1, ws,
                         ws_bytes, g.st);
  if (rc != 0) return rc;
  GLNN_CUDA_OK(cudaMemcpyAsync(out_logprob_host, d_out, sizeof(float) * n * c, cudaMemcpyDeviceToHost,
                               g.st));
  GLNN_CUDA_OK(cudaStreamSynchronize(g.st));
  return 0;
}
