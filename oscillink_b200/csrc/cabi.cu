// extern "C" surface declared in include/oscillink_b200.h.  No exceptions cross this boundary:
// every entry point returns a status code and records the failure text per host thread.
#include <cstdio>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "pcg.cuh"

namespace osc {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return OSC_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// launchers implemented in the kernel translation units
int launch_normalize(const float*, int64_t, int, float*, float*, float*, cudaStream_t, void* h16 = nullptr);
int launch_knn_simt(const float*, const float*, int64_t, int64_t, int64_t, int64_t, int, int, int32_t*,
                    float*, cudaStream_t);
int knn_tc_supported(int64_t N, int D, int kc);
int knn_tc1_supported(int64_t n_rows, int64_t N, int D, int kc);
int knn_tch_supported(int64_t n_rows, int64_t N, int D, int kc);
int launch_knn_tc(const float* q_hi, const float* q_lo, const float* all_hi, const float* all_lo,
                  int64_t batch, int64_t n_rows, int64_t row0, int64_t N, int D, int kc,
                  int32_t* cand_idx, float* cand_sim, bool onepass, cudaStream_t st, bool f16 = false);
int launch_rescore(const float*, const float*, int64_t, int64_t, int64_t, int64_t, int, const int32_t*,
                   const float*, int, int, float, int32_t*, float*, float*, int64_t*, int*, cudaStream_t,
                   int64_t exhaustive_limit = -1, float* dedup_S = nullptr);
int launch_assemble(const int32_t*, const float*, int64_t, int64_t, int, float, int32_t*, float*, float*,
                    int32_t*, float*, int64_t*, float*, cudaStream_t);
int launch_receipt_full(const osc_graph_t*, const osc_params_t*, const float*, const float*,
                        const float*, const float*, int, float, float*, float*, float*, int32_t*,
                        float*, float*, float*, float*, cudaStream_t);
int launch_row_align(const float*, const float*, int64_t, int, float*, cudaStream_t);
int launch_pair_d2(const float*, const float*, const int32_t*, int64_t, int, float*, cudaStream_t);
size_t mmr_workspace(int64_t N);
int launch_mmr(const float*, const float*, int64_t, int, int, int32_t*, void*, size_t, cudaStream_t);
int dist_nccl_version(int*);
int dist_unique_id(unsigned char*);
int dist_comm_init(const unsigned char*, int, int, void**);
int dist_comm_destroy(void*);
int dist_halo_plan_workspace(int64_t, size_t*);
int dist_halo_plan(const int32_t*, int64_t, int, const int32_t*, int64_t, int64_t, int64_t, int64_t, int32_t*,
                   int64_t, int32_t*, int32_t*, int64_t*, void*, size_t, cudaStream_t);
int dist_pcg_workspace(const osc_dist_t*, int64_t, int, size_t*);
int dist_halo_exchange(const osc_dist_t*, int, float*, cudaStream_t);
int dist_pcg_solve(const osc_dist_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*, int, float,
                   int, float, int, double, int, const float*, const float*, const float*, const float*, int,
                   float*, int*, float*, void*, size_t, cudaStream_t);
int dist_delta_h(const osc_dist_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*, const float*,
                 const float*, const float*, int, double*, void*, size_t, cudaStream_t);
int batched_supported(int64_t, int, int);
int batched_workspace(int64_t, int64_t, int, size_t*);
int batched_settle(const osc_graph_t*, const osc_params_t*, const osc_batched_args_t*, void*, size_t,
                   cudaStream_t);

// Candidate-list width per engine.
//  3xTF32 / fp32 FMA engines (error ~1e-6): k + 4 spare candidates.
//  single-product TF32 engine (error bound 2^-10): the list must reach ~1e-3 below the k-th
//  score for the completeness check to pass without the exhaustive path -- about as many spare
//  candidates as neighbours; the register top-k exists for 16 / 24 / 32 entries, so the whole
//  template width is used (k <= 8: 16, k <= 12: 24, else 32).
static int candidate_width(int64_t N, int k, int eng) {
  int64_t kc = (int64_t)k + 4;
  if (eng == OSC_KNN_TC1 || eng == OSC_KNN_TCH) {
    const int want = 2 * k < 16 ? 16 : 2 * k;
    const int64_t wide = want <= 16 ? 16 : (want <= 24 ? 24 : 32);
    if (wide > kc) kc = wide;
  }
  if (kc > N - 1) kc = N - 1;
  if (kc < 1) kc = 1;
  return (int)kc;
}

static float engine_eps(int eng) {
  return (eng == OSC_KNN_TC1 || eng == OSC_KNN_TCH) ? OSC_KNN_EPS_TC1 : OSC_KNN_EPS;
}

// AUTO resolves to a single-product engine wherever one covers the shape -- fp16 operands first, tf32
// where D % 8 != 0: measured on B200 they build the same graphs as the 3xTF32 engine (0 rows need the
// exhaustive path on Gaussian anchors), tf32 1.33x faster at N=1200 D=384 and 3.4x faster at N=1M
// D=768 (7.36 s -> 2.19 s).  OSC_KNN_AUTO=tc / tc1 / tch forces one engine (A/B switch).
static int auto_choice() {
  const char* e = getenv("OSC_KNN_AUTO");
  if (e == nullptr) return OSC_KNN_AUTO;
  if (strcmp(e, "tc") == 0 || strcmp(e, "TC") == 0) return OSC_KNN_TC;
  if (strcmp(e, "tc1") == 0 || strcmp(e, "TC1") == 0) return OSC_KNN_TC1;
  if (strcmp(e, "tch") == 0 || strcmp(e, "TCH") == 0) return OSC_KNN_TCH;
  return OSC_KNN_AUTO;
}

// engine for `n_rows` query rows against N columns; -1: the requested engine does not cover the shape
static int pick_engine(int flags, int64_t n_rows, int64_t N, int D, int k) {
  const int want = flags & 7;
  if (want == OSC_KNN_SIMT) return OSC_KNN_SIMT;
  const int ok3 = knn_tc_supported(N, D, candidate_width(N, k, OSC_KNN_TC));
  const int ok1 = knn_tc1_supported(n_rows, N, D, candidate_width(N, k, OSC_KNN_TC1));
  const int okh = knn_tch_supported(n_rows, N, D, candidate_width(N, k, OSC_KNN_TCH));
  if (want == OSC_KNN_TC) return ok3 ? OSC_KNN_TC : -1;
  if (want == OSC_KNN_TC1) return ok1 ? OSC_KNN_TC1 : -1;
  if (want == OSC_KNN_TCH) return okh ? OSC_KNN_TCH : -1;
  // AUTO: 128x256 tensor tiles only pay off once a lattice fills a few of them
  if (N < 256) return OSC_KNN_SIMT;
  const int pref = auto_choice();
  if (pref == OSC_KNN_TC && ok3) return OSC_KNN_TC;
  if (pref == OSC_KNN_TC1 && ok1) return OSC_KNN_TC1;
  if (k <= 16) {
    if (okh && pref != OSC_KNN_TC1) return OSC_KNN_TCH;
    if (ok1) return OSC_KNN_TC1;
  }
  return ok3 ? OSC_KNN_TC : OSC_KNN_SIMT;
}

}  // namespace osc

using namespace osc;

extern "C" {

int osc_abi_version(void) { return OSC_ABI_VERSION; }
const char* osc_last_error(void) { return g_last_error.c_str(); }

int osc_device_info(int device, int* h_sm_count, int* h_smem_optin, int* h_cc) {
  int n = 0, s = 0, mj = 0, mn = 0;
  OSC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  OSC_CUDA(cudaDeviceGetAttribute(&s, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  OSC_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, device));
  OSC_CUDA(cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, device));
  if (h_sm_count) *h_sm_count = n;
  if (h_smem_optin) *h_smem_optin = s;
  if (h_cc) *h_cc = mj * 10 + mn;
  return OSC_OK;
}

int osc_normalize_rows(const float* Y, int64_t rows, int32_t D, float* Yn, float* Yn_hi, float* Yn_lo,
                       void* stream) {
  OSC_REQUIRE(Y != nullptr && Yn != nullptr && rows >= 0 && D >= 1, "normalize_rows: bad argument");
  OSC_REQUIRE(Yn_lo == nullptr || Yn_hi != nullptr, "normalize_rows: lo without hi");
  return launch_normalize(Y, rows, D, Yn, Yn_hi, Yn_lo, (cudaStream_t)stream);
}

int osc_normalize_rows_f16(const float* Y, int64_t rows, int32_t D, float* Yn, void* Yn_h16, void* stream) {
  OSC_REQUIRE(Y != nullptr && Yn != nullptr && Yn_h16 != nullptr && rows >= 0 && D >= 1,
              "normalize_rows_f16: bad argument");
  return launch_normalize(Y, rows, D, Yn, nullptr, nullptr, (cudaStream_t)stream, Yn_h16);
}

int osc_knn_tc_supported(int64_t N, int32_t D, int32_t kc) { return knn_tc_supported(N, D, kc); }

int osc_knn_plan(int64_t n_rows, int64_t N, int32_t D, int32_t k, int32_t flags, int32_t* h_engine,
                 int32_t* h_kc, float* h_eps) {
  OSC_REQUIRE(n_rows >= 1 && N >= 2 && D >= 1 && k >= 1 && k <= N - 1, "knn_plan: bad shape");
  const int eng = pick_engine(flags, n_rows, N, D, k);
  if (eng < 0) return fail(OSC_ERR_UNSUPPORTED, "knn_plan: tensor-core engine does not cover this shape");
  if (h_engine) *h_engine = eng;
  if (h_kc) *h_kc = candidate_width(N, k, eng);
  if (h_eps) *h_eps = engine_eps(eng);
  return OSC_OK;
}

int osc_knn_candidates_workspace(int64_t, int64_t, int64_t, int32_t, int32_t, int32_t, size_t* h_bytes) {
  if (h_bytes) *h_bytes = 256;
  return OSC_OK;
}

int osc_knn_candidates(const float* Yn_q, const float* Yn_all, const float* q_hi, const float* q_lo,
                       const float* all_hi, const float* all_lo, int64_t batch, int64_t n_rows,
                       int64_t row0, int64_t N, int32_t D, int32_t kc, int32_t flags,
                       int32_t* cand_idx, float* cand_sim, void*, size_t, void* stream) {
  OSC_REQUIRE(batch >= 0 && n_rows >= 0 && N >= 2 && D >= 1, "knn_candidates: bad shape");
  OSC_REQUIRE(kc >= 1 && kc <= N - 1 && kc <= 132, "knn_candidates: kc must be in [1, min(N-1,132)]");
  OSC_REQUIRE(batch <= 65535, "knn_candidates: batch > 65535 (split the batch)");
  OSC_REQUIRE(cand_idx != nullptr && cand_sim != nullptr, "knn_candidates: NULL output");
  if (batch == 0 || n_rows == 0) return OSC_OK;
  // The caller fixed kc (and the eps it will hand to the checked re-scoring), so AUTO never picks
  // the single-product engine here: that one is entered on explicit request only (osc_knn_plan).
  int eng = flags & 7;
  if (eng == OSC_KNN_AUTO) eng = (N >= 256 && knn_tc_supported(N, D, kc)) ? OSC_KNN_TC : OSC_KNN_SIMT;
  if ((eng == OSC_KNN_TC && !knn_tc_supported(N, D, kc)) ||
      (eng == OSC_KNN_TC1 && !knn_tc1_supported(n_rows, N, D, kc)) ||
      (eng == OSC_KNN_TCH && !knn_tch_supported(n_rows, N, D, kc)))
    return fail(OSC_ERR_UNSUPPORTED, "knn_candidates: tensor-core engine does not cover this shape");
  if (eng == OSC_KNN_TCH) {
    OSC_REQUIRE(q_hi && all_hi, "knn_candidates: TCH engine needs the fp16 rows (q_hi / all_hi)");
    return launch_knn_tc(q_hi, nullptr, all_hi, nullptr, batch, n_rows, row0, N, D, kc, cand_idx, cand_sim,
                         true, (cudaStream_t)stream, true);
  }
  if (eng == OSC_KNN_TC1) {
    OSC_REQUIRE(q_hi && all_hi, "knn_candidates: TC1 engine needs the tf32-rounded rows (hi)");
    return launch_knn_tc(q_hi, nullptr, all_hi, nullptr, batch, n_rows, row0, N, D, kc, cand_idx, cand_sim,
                         true, (cudaStream_t)stream);
  }
  if (eng == OSC_KNN_TC) {
    OSC_REQUIRE(q_hi && q_lo && all_hi && all_lo, "knn_candidates: TC engine needs the hi/lo split");
    return launch_knn_tc(q_hi, q_lo, all_hi, all_lo, batch, n_rows, row0, N, D, kc, cand_idx, cand_sim,
                         false, (cudaStream_t)stream);
  }
  OSC_REQUIRE(Yn_q != nullptr && Yn_all != nullptr, "knn_candidates: NULL input");
  return launch_knn_simt(Yn_q, Yn_all, batch, n_rows, row0, N, D, kc, cand_idx, cand_sim,
                         (cudaStream_t)stream);
}

int osc_knn_rescore(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows, int64_t N,
                    int32_t D, const int32_t* cand_idx, int32_t kc, int32_t k, int32_t* top_idx,
                    float* top_sim, float* gap, void* stream) {
  OSC_REQUIRE(Yn_q && Yn_all && cand_idx && top_idx && top_sim, "knn_rescore: NULL argument");
  OSC_REQUIRE(k >= 1 && kc >= k && batch <= 65535, "knn_rescore: need 1 <= k <= kc");
  if (batch == 0 || n_rows == 0) return OSC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(cudaMemsetAsync(top_idx, 0xFF, sizeof(int32_t) * batch * n_rows * k, st));
  OSC_CUDA(cudaMemsetAsync(top_sim, 0, sizeof(float) * batch * n_rows * k, st));
  return launch_rescore(Yn_q, Yn_all, batch, n_rows, 0, N, D, cand_idx, nullptr, kc, k, 0.f, top_idx, top_sim,
                        gap, nullptr, nullptr, st);
}

int osc_knn_rescore_workspace(int64_t batch, int64_t n_rows, size_t* h_bytes) {
  OSC_REQUIRE(h_bytes != nullptr && batch >= 0 && n_rows >= 0, "knn_rescore_workspace: bad argument");
  // flagged-row list + the [rows][<= 32] score table of the duplicate-free re-scoring (knn_rescore_owned_kernel)
  *h_bytes = align_up((size_t)batch * n_rows * sizeof(int64_t)) + 256 +
             align_up((size_t)batch * n_rows * 32 * sizeof(float));
  return OSC_OK;
}

int osc_knn_rescore_guarded(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows,
                            int64_t row0, int64_t N, int32_t D, const int32_t* cand_idx,
                            const float* cand_sim, int32_t kc, int32_t k, float eps, int64_t exhaustive_limit,
                            int32_t* top_idx, float* top_sim, float* gap, int32_t* d_n_flagged,
                            void* workspace, size_t ws_bytes, void* stream) {
  OSC_REQUIRE(Yn_q && Yn_all && cand_idx && cand_sim && top_idx && top_sim && d_n_flagged,
              "knn_rescore_checked: NULL argument");
  OSC_REQUIRE(k >= 1 && kc >= k && batch <= 65535 && eps >= 0.f, "knn_rescore_checked: need 1 <= k <= kc");
  if (batch == 0 || n_rows == 0) return OSC_OK;
  size_t need = 0;
  osc_knn_rescore_workspace(batch, n_rows, &need);
  if (ws_bytes < need || workspace == nullptr) return fail(OSC_ERR_WORKSPACE, "knn_rescore_checked: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  OSC_CUDA(cudaMemsetAsync(top_idx, 0xFF, sizeof(int32_t) * batch * n_rows * k, st));
  OSC_CUDA(cudaMemsetAsync(top_sim, 0, sizeof(float) * batch * n_rows * k, st));
  float* S = reinterpret_cast<float*>(static_cast<char*>(workspace) +
                                      align_up((size_t)batch * n_rows * sizeof(int64_t)) + 256);
  return launch_rescore(Yn_q, Yn_all, batch, n_rows, row0, N, D, cand_idx, cand_sim, kc, k, eps, top_idx,
                        top_sim, gap, static_cast<int64_t*>(workspace), d_n_flagged, st, exhaustive_limit, S);
}

int osc_knn_rescore_checked(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows,
                            int64_t row0, int64_t N, int32_t D, const int32_t* cand_idx,
                            const float* cand_sim, int32_t kc, int32_t k, float eps, int32_t* top_idx,
                            float* top_sim, float* gap, int32_t* d_n_flagged, void* workspace,
                            size_t ws_bytes, void* stream) {
  return osc_knn_rescore_guarded(Yn_q, Yn_all, batch, n_rows, row0, N, D, cand_idx, cand_sim, kc, k, eps, -1,
                                 top_idx, top_sim, gap, d_n_flagged, workspace, ws_bytes, stream);
}

int64_t osc_knn_exhaustive_limit(int64_t rows) {
  const int64_t one_percent = rows / 100;
  return one_percent > 64 ? one_percent : 64;
}

int osc_graph_assemble(const int32_t* top_idx, const float* top_sim, int64_t batch, int64_t N, int32_t k,
                       float row_cap, int32_t* nbr, float* A, float* W, int32_t* deg, float* sqrt_deg,
                       int64_t* nnz, float* scratch, void* stream) {
  OSC_REQUIRE(top_idx && top_sim && nbr && A && W && deg && sqrt_deg && scratch,
              "graph_assemble: NULL argument");
  OSC_REQUIRE(k >= 1 && batch <= 65535, "graph_assemble: bad k/batch");
  return launch_assemble(top_idx, top_sim, batch, N, k, row_cap, nbr, A, W, deg, sqrt_deg, nnz, scratch,
                         (cudaStream_t)stream);
}

// ------------------------------------------------------------------ chain prior (a6), host side
// graph.py:101-111 / lattice.py:129-149 in sparse form.  A chain is O(len) data that the caller holds on
// the host, so this is plain C++: max-merged symmetric path weights, CSR over the distinct chain nodes
// (rows ascending, columns ascending inside a row), sdp_u = sqrt(max(rowsum_u, 1e-12)) accumulated in
// fp32 in column order, Wp_uv = (Ap_uv * (1/sdp_u)) * (1/sdp_v) -- the rounding sequence of the
// reference's normalized_laplacian (graph.py:87-92).
static int chain_collect(const int32_t* h_chain, int32_t len, const float* h_weights, int64_t N,
                         std::map<std::pair<int32_t, int32_t>, float>& ap) {
  OSC_REQUIRE(h_chain != nullptr && len >= 2, "chain_build: chain must contain at least two indices");
  for (int32_t t = 0; t < len; ++t)
    OSC_REQUIRE(h_chain[t] >= 0 && (int64_t)h_chain[t] < N, "chain_build: chain indices out of bounds");
  for (int32_t t = 0; t + 1 < len; ++t) {
    const int32_t u = h_chain[t], v = h_chain[t + 1];
    const float w = h_weights ? h_weights[t] : 1.0f;
    float& a = ap[{u, v}];  // value-initialised to 0
    if (w > a) a = w;
    float& b = ap[{v, u}];
    if (w > b) b = w;
  }
  return OSC_OK;
}

int osc_chain_build_size(const int32_t* h_chain, int32_t len, int64_t N, int32_t* h_n_rows, int32_t* h_nnz) {
  std::map<std::pair<int32_t, int32_t>, float> ap;
  const int rc = chain_collect(h_chain, len, nullptr, N, ap);
  if (rc) return rc;
  int32_t rows = 0, last = -1;
  for (const auto& e : ap)
    if (e.first.first != last) {
      last = e.first.first;
      ++rows;
    }
  if (h_n_rows) *h_n_rows = rows;
  if (h_nnz) *h_nnz = (int32_t)ap.size();
  return OSC_OK;
}

int osc_chain_build(const int32_t* h_chain, int32_t len, const float* h_weights, int64_t N, int32_t* h_rows,
                    int32_t* h_rowptr, int32_t* h_col, float* h_Wp, float* h_Ap, int32_t* h_slot) {
  OSC_REQUIRE(h_rows && h_rowptr && h_col && h_Wp && h_Ap && h_slot, "chain_build: NULL output");
  std::map<std::pair<int32_t, int32_t>, float> ap;
  const int rc = chain_collect(h_chain, len, h_weights, N, ap);
  if (rc) return rc;
  for (int64_t i = 0; i < N; ++i) h_slot[i] = -1;
  std::vector<float> sdp;
  int32_t n_rows = 0, nnz = 0, last = -1;
  float d = 0.f;
  h_rowptr[0] = 0;
  for (const auto& e : ap) {  // ordered by (u, v)
    const int32_t u = e.first.first;
    if (u != last) {
      if (last >= 0) {
        sdp.push_back(sqrtf(fmaxf(d, 1e-12f)));
        h_rowptr[n_rows] = nnz;
      }
      h_rows[n_rows] = u;
      h_slot[u] = n_rows;
      ++n_rows;
      last = u;
      d = 0.f;
    }
    d = d + e.second;  // fp32, column order
    h_col[nnz] = e.first.second;
    h_Ap[nnz] = e.second;
    ++nnz;
  }
  sdp.push_back(sqrtf(fmaxf(d, 1e-12f)));
  h_rowptr[n_rows] = nnz;
  for (int32_t r = 0; r < n_rows; ++r) {
    const float inv_u = 1.0f / sdp[r];
    for (int32_t e = h_rowptr[r]; e < h_rowptr[r + 1]; ++e) {
      // a neighbour that only appears as a column always has its own row (symmetry)
      const float inv_v = 1.0f / sdp[h_slot[h_col[e]]];
      const float t = h_Ap[e] * inv_u;
      h_Wp[e] = t * inv_v;
    }
  }
  return OSC_OK;
}

int osc_knn_build_workspace(int64_t batch, int64_t N, int32_t D, int32_t k, int32_t flags,
                            size_t* h_bytes) {
  OSC_REQUIRE(h_bytes != nullptr && batch >= 0 && N >= 0 && D >= 1 && k >= 1, "knn_build_workspace: bad argument");
  if (N < 2) {
    *h_bytes = 256;
    return OSC_OK;
  }
  int eng = pick_engine(flags, N, N, D, k);
  if (eng < 0) eng = OSC_KNN_TC;  // osc_knn_build reports the error; size for the larger layout
  const size_t rows = (size_t)batch * N;
  auto layout = [&](int e) {
    const int kc = candidate_width(N, k, e);
    size_t b = align_up(rows * D * sizeof(float));                     // Yn
    if (e == OSC_KNN_TC) b += 2 * align_up(rows * D * sizeof(float));  // hi, lo
    if (e == OSC_KNN_TC1) b += align_up(rows * D * sizeof(float));     // hi
    if (e == OSC_KNN_TCH) b += align_up(rows * D * 2);                 // fp16 rows
    b += align_up(rows * kc * sizeof(int32_t)) + align_up(rows * kc * sizeof(float));  // candidates
    b += align_up(rows * k * sizeof(int32_t)) + align_up(rows * k * sizeof(float));    // top-k
    b += align_up(rows * sizeof(float));                                               // cap scale
    b += align_up(rows * sizeof(int64_t)) + 512;                                       // flagged rows + counter
    b += align_up(rows * 32 * sizeof(float));                                          // duplicate-free score table
    return b + 1024;
  };
  size_t b = layout(eng);
  // a single-product engine may hand the build over to the 3xTF32 engine (see osc_knn_build)
  if ((eng == OSC_KNN_TC1 || eng == OSC_KNN_TCH) && knn_tc_supported(N, D, candidate_width(N, k, OSC_KNN_TC))) {
    const size_t b3 = layout(OSC_KNN_TC);
    if (b3 > b) b = b3;
  }
  *h_bytes = b;
  return OSC_OK;
}

int osc_knn_build(const float* Y, int64_t batch, int64_t N, int32_t D, int32_t k, float row_cap,
                  int32_t flags, int32_t* nbr, float* A, float* W, int32_t* deg, float* sqrt_deg,
                  int64_t* nnz, float* gap, void* workspace, size_t ws_bytes, void* stream) {
  OSC_REQUIRE(Y && nbr && A && W && deg && sqrt_deg, "knn_build: NULL argument");
  OSC_REQUIRE(batch >= 0 && N >= 0 && D >= 1 && k >= 1, "knn_build: bad shape");
  OSC_REQUIRE(batch <= 65535, "knn_build: batch > 65535 (split the batch)");
  OSC_REQUIRE(k <= 128, "knn_build: kneighbors > 128 is not supported");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows = (size_t)batch * N;
  if (rows == 0) return OSC_OK;
  if (N < 2) {
    // graph.py:30-32 -- no neighbours possible; sqrt_deg = sqrt(1e-12)
    OSC_CUDA(cudaMemsetAsync(nbr, 0xFF, rows * k * sizeof(int32_t), st));
    OSC_CUDA(cudaMemsetAsync(A, 0, rows * k * sizeof(float), st));
    OSC_CUDA(cudaMemsetAsync(W, 0, rows * k * sizeof(float), st));
    OSC_CUDA(cudaMemsetAsync(deg, 0, rows * sizeof(int32_t), st));
    if (nnz) OSC_CUDA(cudaMemsetAsync(nnz, 0, batch * sizeof(int64_t), st));
    std::vector<float> sd(rows, 1e-6f);
    OSC_CUDA(cudaMemcpyAsync(sqrt_deg, sd.data(), rows * sizeof(float), cudaMemcpyHostToDevice, st));
    OSC_CUDA(cudaStreamSynchronize(st));
    return OSC_OK;
  }
  OSC_REQUIRE(k <= N - 1, "knn_build: k must be clamped to N-1 by the caller (lattice.py:60)");
  size_t need = 0;
  int rc = osc_knn_build_workspace(batch, N, D, k, flags, &need);
  if (rc) return rc;
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "knn_build: workspace too small");
  int eng = pick_engine(flags, N, N, D, k);
  if (eng < 0) return fail(OSC_ERR_UNSUPPORTED, "knn_build: tensor-core engine does not cover this shape");
  // Engine robustness: the single-product engines prove a row's candidate list complete only if the k-th
  // exact score clears the list's tail by eps = 1.25e-3; on clustered / near-duplicate anchors many rows
  // fail that test and would each cost an exhaustive N*D fp64 scan.  When more than max(64, 1 %) of the
  // rows are flagged, the exhaustive pass is skipped and the candidate pass is re-run ONCE with the 3xTF32
  // engine (eps = 1e-5), whose flagged rows (true near-ties only) then do take the exhaustive path.
  const bool may_fall_back = (eng == OSC_KNN_TC1 || eng == OSC_KNN_TCH) &&
                             knn_tc_supported(N, D, candidate_width(N, k, OSC_KNN_TC));
  for (int attempt = 0; attempt < 2; ++attempt) {
  const int kc = candidate_width(N, k, eng);
  Arena ar(workspace, ws_bytes);
  float* Yn = ar.take<float>(rows * D);
  float *hi = nullptr, *lo = nullptr;
  if (eng == OSC_KNN_TC || eng == OSC_KNN_TC1) hi = ar.take<float>(rows * D);
  if (eng == OSC_KNN_TC) lo = ar.take<float>(rows * D);
  uint16_t* h16 = nullptr;
  if (eng == OSC_KNN_TCH) {
    h16 = ar.take<uint16_t>(rows * D);
    hi = reinterpret_cast<float*>(h16);  // handed to osc_knn_candidates as q_hi / all_hi
  }
  int32_t* cand_idx = ar.take<int32_t>(rows * kc);
  float* cand_sim = ar.take<float>(rows * kc);
  int32_t* top_idx = ar.take<int32_t>(rows * k);
  float* top_sim = ar.take<float>(rows * k);
  float* cscale = ar.take<float>(rows);
  int64_t* flagged = ar.take<int64_t>(rows);
  int* n_flagged = ar.take<int>(1);
  float* dedup_S = ar.take<float>(rows * 32);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "knn_build: workspace too small");
  if ((rc = launch_normalize(Y, (int64_t)rows, D, Yn, h16 ? nullptr : hi, lo, st, h16))) return rc;
  if ((rc = osc_knn_candidates(Yn, Yn, hi, lo, hi, lo, batch, N, 0, N, D, kc, eng, cand_idx, cand_sim,
                               nullptr, 0, stream)))
    return rc;
  OSC_CUDA(cudaMemsetAsync(top_idx, 0xFF, sizeof(int32_t) * rows * k, st));
  OSC_CUDA(cudaMemsetAsync(top_sim, 0, sizeof(float) * rows * k, st));
  const bool guarded = may_fall_back && attempt == 0;
  const int64_t limit = guarded ? osc_knn_exhaustive_limit((int64_t)rows) : -1;
  if ((rc = launch_rescore(Yn, Yn, batch, N, 0, N, D, cand_idx, cand_sim, kc, k, engine_eps(eng), top_idx,
                           top_sim, gap, flagged, n_flagged, st, limit, dedup_S)))
    return rc;
  if (guarded) {
    int h_flagged = 0;
    OSC_CUDA(cudaMemcpyAsync(&h_flagged, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, st));
    OSC_CUDA(cudaStreamSynchronize(st));
    if ((int64_t)h_flagged > limit) {
      eng = OSC_KNN_TC;
      continue;
    }
  }
  return launch_assemble(top_idx, top_sim, batch, N, k, row_cap, nbr, A, W, deg, sqrt_deg, nnz, cscale, st);
  }
  return fail(OSC_ERR_CUDA, "knn_build: unreachable");
}

// ------------------------------------------------------------------ PCG
int osc_pcg_plan(osc_pcg_dims_t* dims, size_t* h_ws_bytes) { return pcg_plan(dims, h_ws_bytes); }
int osc_pcg_max_ell_width(int32_t D) { return D >= 1 ? pcg_max_ell_width(D) : 0; }

int osc_pcg_setup(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                  int32_t warm_start, float inertia, const float* Y_loc, const float* U_loc,
                  const float* psi, const float* gates_loc, float* X_loc, float* Bv_loc, void* stream) {
  OSC_REQUIRE(dims && prm && Y_loc && U_loc && psi && X_loc && Bv_loc, "pcg_setup: NULL argument");
  return pcg_setup(dims, prm, mode, dt, warm_start, inertia, Y_loc, U_loc, psi, gates_loc, X_loc, Bv_loc,
                   (cudaStream_t)stream);
}

int osc_pcg_residual0(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                      const osc_params_t* prm, int32_t mode, float dt, int32_t jacobi,
                      const float* gates_loc, const float* X_all, float* RBv_loc, float* P_loc,
                      double* part_rz, void* stream) {
  OSC_REQUIRE(dims && g && prm && X_all && RBv_loc && P_loc && part_rz, "pcg_residual0: NULL argument");
  return pcg_residual0(dims, g, chain, prm, mode, dt, jacobi, gates_loc, X_all, RBv_loc, P_loc, part_rz,
                       (cudaStream_t)stream);
}

int osc_pcg_spmm_dot(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                     const osc_params_t* prm, int32_t mode, float dt, const float* gates_loc,
                     const float* P_all, float* AP_loc, double* part_pap, void* stream) {
  OSC_REQUIRE(dims && g && prm && P_all && AP_loc && part_pap, "pcg_spmm_dot: NULL argument");
  return pcg_spmm_dot(dims, g, chain, prm, mode, dt, gates_loc, P_all, AP_loc, part_pap,
                      (cudaStream_t)stream);
}

int osc_enable_peer_access(int32_t peer_device) {
  int dev = 0;
  OSC_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return OSC_OK;
  int can = 0;
  OSC_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) return fail(OSC_ERR_UNSUPPORTED, "enable_peer_access: no P2P path between the two devices");
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-free error state
    return OSC_OK;
  }
  OSC_CUDA(e);
  return OSC_OK;
}

// Peer-mapped buffers: a plain cudaMalloc allocation (so the IPC handle covers exactly this buffer,
// offset 0) exported with cudaIpcGetMemHandle; a peer process maps it from ITS device with
// cudaIpcOpenMemHandle(..., cudaIpcMemLazyEnablePeerAccess), which also enables NVLink peer access.
int osc_peer_alloc(size_t bytes, void** d_ptr, unsigned char* handle64) {
  OSC_REQUIRE(d_ptr != nullptr && handle64 != nullptr && bytes > 0, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  OSC_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "peer_alloc");
  }
  memcpy(handle64, &h, 64);
  *d_ptr = p;
  return OSC_OK;
}

int osc_peer_open(const unsigned char* handle64, void** d_ptr) {
  OSC_REQUIRE(d_ptr != nullptr && handle64 != nullptr, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  OSC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return OSC_OK;
}

int osc_peer_close(void* d_ptr) {
  if (d_ptr != nullptr) OSC_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return OSC_OK;
}

int osc_peer_free(void* d_ptr) {
  if (d_ptr != nullptr) OSC_CUDA(cudaFree(d_ptr));
  return OSC_OK;
}

int osc_pcg_residual0_p2p(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                          const osc_params_t* prm, int32_t mode, float dt, int32_t jacobi,
                          const float* gates_loc, const float* const* d_peer_X, int32_t world, int64_t shard,
                          float* RBv_loc, float* P_loc, double* part_rz, void* stream) {
  OSC_REQUIRE(dims && g && prm && d_peer_X && RBv_loc && P_loc && part_rz, "pcg_residual0_p2p: NULL argument");
  OSC_REQUIRE(world >= 1 && shard >= 1 && (int64_t)world * shard >= dims->N, "pcg_residual0_p2p: bad partition");
  return pcg_residual0_p2p(dims, g, chain, prm, mode, dt, jacobi, gates_loc, d_peer_X, shard, RBv_loc, P_loc,
                           part_rz, (cudaStream_t)stream);
}

int osc_pcg_spmm_dot_p2p(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                         const osc_params_t* prm, int32_t mode, float dt, const float* gates_loc,
                         const float* const* d_peer_P, int32_t world, int64_t shard, float* AP_loc,
                         double* part_pap, void* stream) {
  OSC_REQUIRE(dims && g && prm && d_peer_P && AP_loc && part_pap, "pcg_spmm_dot_p2p: NULL argument");
  OSC_REQUIRE(world >= 1 && shard >= 1 && (int64_t)world * shard >= dims->N, "pcg_spmm_dot_p2p: bad partition");
  return pcg_spmm_dot_p2p(dims, g, chain, prm, mode, dt, gates_loc, d_peer_P, shard, AP_loc, part_pap,
                          (cudaStream_t)stream);
}

int osc_pcg_reduce(const double* part, int32_t n_blocks, int32_t D, float* out, float* d_max,
                   void* stream) {
  OSC_REQUIRE(part && out && n_blocks >= 1 && D >= 1, "pcg_reduce: bad argument");
  return pcg_reduce(part, n_blocks, D, out, d_max, nullptr, (cudaStream_t)stream);
}

int osc_pcg_update(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                   int32_t jacobi, const float* gates_loc, const float* rz, const float* pap,
                   const float* P_loc, const float* AP_loc, float* X_loc, float* R_loc,
                   double* part_rr, double* part_rz, void* stream) {
  // X_loc == NULL: r only -- the x update then belongs to osc_pcg_pupdate_x
  OSC_REQUIRE(dims && prm && rz && pap && P_loc && AP_loc && R_loc && part_rr && part_rz,
              "pcg_update: NULL argument");
  return pcg_update(dims, prm, mode, dt, jacobi, gates_loc, rz, pap, P_loc, AP_loc, X_loc, R_loc,
                    part_rr, part_rz, (cudaStream_t)stream);
}

int osc_pcg_pupdate(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                    int32_t jacobi, const float* gates_loc, const float* rz_new, const float* rz_old,
                    const float* R_loc, float* P_loc, void* stream) {
  OSC_REQUIRE(dims && prm && rz_new && rz_old && R_loc && P_loc, "pcg_pupdate: NULL argument");
  return pcg_pupdate(dims, prm, mode, dt, jacobi, gates_loc, rz_new, rz_old, R_loc, P_loc,
                     (cudaStream_t)stream);
}

int osc_pcg_pupdate_x(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                      int32_t jacobi, const float* gates_loc, const float* rz_new, const float* rz_old,
                      const float* pap, const float* R_loc, float* P_loc, float* X_loc, int32_t last,
                      void* stream) {
  OSC_REQUIRE(dims && prm && rz_new && rz_old && pap && R_loc && P_loc && X_loc, "pcg_pupdate_x: NULL argument");
  return pcg_pupdate_x(dims, prm, mode, dt, jacobi, gates_loc, rz_new, rz_old, pap, R_loc, P_loc, X_loc, nullptr,
                       0, last, (cudaStream_t)stream);
}

int osc_pcg_solve(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm, int32_t mode,
                  float dt, int32_t warm_start, float inertia, int32_t jacobi, double tol,
                  int32_t max_iters, const float* Y, const float* U, const float* psi,
                  const float* gates, int32_t D, float* X, int32_t* h_iters, float* h_res,
                  void* workspace, size_t ws_bytes, void* stream) {
  return pcg_solve(g, chain, prm, mode, dt, warm_start, inertia, jacobi, tol, max_iters, Y, U, psi, gates,
                   D, X, h_iters, h_res, workspace, ws_bytes, (cudaStream_t)stream);
}

int osc_pcg_solve_system(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                         int32_t mode, float dt, int32_t jacobi, double tol, int32_t max_iters,
                         const float* gates, int32_t D, float* X, float* B, int32_t* h_iters,
                         float* h_res, void* workspace, size_t ws_bytes, void* stream) {
  return pcg_solve_system(g, chain, prm, mode, dt, jacobi, tol, max_iters, gates, D, X, B, h_iters, h_res,
                          workspace, ws_bytes, (cudaStream_t)stream);
}

int osc_delta_h(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm, const float* U,
                const float* Ustar, const float* gates, int32_t D, double* h_deltaH, void* workspace,
                size_t ws_bytes, void* stream) {
  return delta_h(g, chain, prm, U, Ustar, gates, D, h_deltaH, workspace, ws_bytes, (cudaStream_t)stream);
}

int osc_receipt_full(const osc_graph_t* g, const osc_params_t* prm, const float* Y, const float* Ustar,
                     const float* psi, const float* gates, int32_t D, float z_th, float* coh,
                     float* anchor, float* query, int32_t* null_j, float* null_z, float* null_R,
                     float* row_mu, float* row_sigma, void* stream) {
  OSC_REQUIRE(coh && anchor && query && null_j && null_z && null_R, "receipt_full: NULL output");
  return launch_receipt_full(g, prm, Y, Ustar, psi, gates, D, z_th, coh, anchor, query, null_j, null_z,
                             null_R, row_mu, row_sigma, (cudaStream_t)stream);
}

int osc_row_align(const float* Ustar, const float* psi, int64_t N, int32_t D, float* align, void* stream) {
  OSC_REQUIRE(Ustar && psi && align && D >= 1, "row_align: bad argument");
  return launch_row_align(Ustar, psi, N, D, align, (cudaStream_t)stream);
}

int osc_pair_d2(const float* V, const float* sqrt_deg, const int32_t* pairs, int64_t M, int32_t D,
                float* out, void* stream) {
  OSC_REQUIRE(V && sqrt_deg && (M == 0 || (pairs && out)) && D >= 1, "pair_d2: bad argument");
  return launch_pair_d2(V, sqrt_deg, pairs, M, D, out, (cudaStream_t)stream);
}

int osc_mmr_workspace(int64_t N, size_t* h_bytes) {
  OSC_REQUIRE(h_bytes != nullptr && N >= 0, "mmr_workspace: bad argument");
  *h_bytes = mmr_workspace(N);
  return OSC_OK;
}

int osc_mmr_select(const float* Yn, const float* score, int64_t N, int32_t D, int32_t k, int32_t* chosen,
                   void* workspace, size_t ws_bytes, void* stream) {
  OSC_REQUIRE(Yn && score && chosen && D >= 1, "mmr_select: bad argument");
  return launch_mmr(Yn, score, N, D, k, chosen, workspace, ws_bytes, (cudaStream_t)stream);
}

int osc_dist_nccl_version(int32_t* h_version) {
  OSC_REQUIRE(h_version != nullptr, "osc_dist_nccl_version: NULL argument");
  int v = 0;
  const int rc = dist_nccl_version(&v);
  *h_version = v;
  return rc;
}
int osc_dist_unique_id(unsigned char* h_id128) {
  OSC_REQUIRE(h_id128 != nullptr, "osc_dist_unique_id: NULL argument");
  return dist_unique_id(h_id128);
}
int osc_dist_comm_init(const unsigned char* h_id128, int32_t world, int32_t rank, void** h_comm) {
  OSC_REQUIRE(h_id128 != nullptr && h_comm != nullptr && world >= 1 && rank >= 0 && rank < world,
              "osc_dist_comm_init: bad argument");
  return dist_comm_init(h_id128, world, rank, h_comm);
}
int osc_dist_comm_destroy(void* comm) { return dist_comm_destroy(comm); }
int osc_dist_halo_plan_workspace(int64_t N, size_t* h_bytes) {
  OSC_REQUIRE(h_bytes != nullptr && N >= 0, "osc_dist_halo_plan_workspace: bad argument");
  return dist_halo_plan_workspace(N, h_bytes);
}
int osc_dist_halo_plan(const int32_t* nbr_loc, int64_t n_local, int32_t k, const int32_t* extra_ids,
                       int64_t n_extra, int64_t N, int64_t row0, int64_t shard, int32_t* halo_rows,
                       int64_t halo_cap, int32_t* nbr_out, int32_t* extra_out, int64_t* h_n_halo,
                       void* workspace, size_t ws_bytes, void* stream) {
  return dist_halo_plan(nbr_loc, n_local, k, extra_ids, n_extra, N, row0, shard, halo_rows, halo_cap, nbr_out,
                        extra_out, h_n_halo, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}
int osc_dist_halo_exchange(const osc_dist_t* dist, int32_t D, float* d_flag, void* stream) {
  return dist_halo_exchange(dist, D, d_flag, static_cast<cudaStream_t>(stream));
}
int osc_dist_pcg_workspace(const osc_dist_t* dist, int64_t n_loc, int32_t D, size_t* h_bytes) {
  return dist_pcg_workspace(dist, n_loc, D, h_bytes);
}
int osc_dist_pcg_solve(const osc_dist_t* dist, const osc_graph_t* g, const osc_chain_t* chain,
                       const osc_params_t* prm, int32_t mode, float dt, int32_t warm_start, float inertia,
                       int32_t jacobi, double tol, int32_t max_iters, const float* Y, const float* U,
                       const float* psi, const float* gates, int32_t D, float* X, int32_t* h_iters,
                       float* h_res, void* workspace, size_t ws_bytes, void* stream) {
  return dist_pcg_solve(dist, g, chain, prm, mode, dt, warm_start, inertia, jacobi, tol, max_iters, Y, U, psi,
                        gates, D, X, h_iters, h_res, workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}
int osc_dist_delta_h(const osc_dist_t* dist, const osc_graph_t* g, const osc_chain_t* chain,
                     const osc_params_t* prm, const float* U, const float* Ustar, const float* gates,
                     int32_t D, double* h_deltaH, void* workspace, size_t ws_bytes, void* stream) {
  return dist_delta_h(dist, g, chain, prm, U, Ustar, gates, D, h_deltaH, workspace, ws_bytes,
                      static_cast<cudaStream_t>(stream));
}

int osc_batched_supported(int64_t N, int32_t D, int32_t k) { return batched_supported(N, D, k); }
int osc_batched_workspace(int64_t batch, int64_t N, int32_t D, size_t* h_bytes) {
  OSC_REQUIRE(h_bytes != nullptr, "batched_workspace: NULL");
  return batched_workspace(batch, N, D, h_bytes);
}
int osc_batched_settle(const osc_graph_t* g, const osc_params_t* prm, const osc_batched_args_t* a,
                       void* workspace, size_t ws_bytes, void* stream) {
  return batched_settle(g, prm, a, workspace, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
