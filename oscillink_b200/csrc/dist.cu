// Multi-GPU PCG driver behind the C ABI (osc_dist_*): one process per GPU, NCCL called from here.
//
//   solver.py:19-37      per-column recurrences; every reduction is over rows (axis 0)
//   lattice.py:173-182   the operator whose SpMM needs neighbour rows owned by other ranks
//
// Partitions of the same recurrences:
//   OSC_PART_ROWS     rank g owns rows [g*shard, (g+1)*shard).  Per iteration: the search direction's
//                     remote rows reach the SpMM (halo), the 3*D column dots are all-reduced (SUM).
//       OSC_HALO_ALLGATHER  ncclAllGather of p in front of every SpMM (every rank holds all N rows).
//       OSC_HALO_PULL       each rank keeps [own rows | halo rows] in one block; before an SpMM it PULLS the
//                           unique remote rows its graph references -- sorted by owner and row, so the
//                           reads are near-sequential -- from the peers' blocks over NVLink peer memory
//                           (CUDA IPC mappings), one fetch per row instead of one per reference, ~64 % of
//                           what the all-gather moves on kNN graphs of random anchors.
//   OSC_PART_COLUMNS  rank g owns all N rows of D/world columns: no halo, no dot all-reduce (every
//                     reduction is per column); one 1-float MAX all-reduce per iteration for the stop test.
//
// The stop test is evaluated on the device (pcg_decide) after the all-reduce, so every rank takes the
// same decision; the host polls the verdict (lag 1 for column slabs, lag 0 where an iteration carries an
// all-gather worth milliseconds).  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already
// loaded in the process, e.g. torch's, is preferred): the library has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include <cub/device/device_scan.cuh>
#include <mutex>

#include "pcg.cuh"

namespace osc {

// ---------------------------------------------------------------- NCCL binding
struct NcclApi {
  bool ok = false;
  std::string why;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void nccl_load() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    g_nccl.why = "libnccl.so.2 not found";
    return;
  }
#define OSC_SYM(field, name)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  if (g_nccl.field == nullptr) {                                          \
    g_nccl.why = std::string("libnccl: missing symbol ") + name;           \
    return;                                                               \
  }
  OSC_SYM(GetVersion, "ncclGetVersion")
  OSC_SYM(GetUniqueId, "ncclGetUniqueId")
  OSC_SYM(CommInitRank, "ncclCommInitRank")
  OSC_SYM(CommDestroy, "ncclCommDestroy")
  OSC_SYM(CommCount, "ncclCommCount")
  OSC_SYM(CommUserRank, "ncclCommUserRank")
  OSC_SYM(AllReduce, "ncclAllReduce")
  OSC_SYM(AllGather, "ncclAllGather")
  OSC_SYM(GetErrorString, "ncclGetErrorString")
#undef OSC_SYM
  g_nccl.ok = true;
}

static NcclApi* nccl_api() {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl.ok) {
    set_error("NCCL unavailable: " + g_nccl.why);
    return nullptr;
  }
  return &g_nccl;
}

#define OSC_NCCL(api, expr)                                                                       \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess && _r != ncclInProgress)                                                \
      return ::osc::fail(OSC_ERR_CUDA, std::string("NCCL ") + #expr + ": " + (api)->GetErrorString(_r)); \
  } while (0)

int dist_nccl_version(int* h_version) {
  NcclApi* api = nccl_api();
  if (api == nullptr) return OSC_ERR_UNSUPPORTED;
  OSC_NCCL(api, api->GetVersion(h_version));
  return OSC_OK;
}

int dist_unique_id(unsigned char* h_id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi* api = nccl_api();
  if (api == nullptr) return OSC_ERR_UNSUPPORTED;
  ncclUniqueId id;
  OSC_NCCL(api, api->GetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return OSC_OK;
}

int dist_comm_init(const unsigned char* h_id128, int world, int rank, void** h_comm) {
  NcclApi* api = nccl_api();
  if (api == nullptr) return OSC_ERR_UNSUPPORTED;
  ncclUniqueId id;
  memcpy(&id, h_id128, sizeof(id));
  ncclComm_t comm = nullptr;
  OSC_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  *h_comm = comm;
  return OSC_OK;
}

int dist_comm_destroy(void* comm) {
  NcclApi* api = nccl_api();
  if (api == nullptr) return OSC_ERR_UNSUPPORTED;
  if (comm != nullptr) OSC_NCCL(api, api->CommDestroy(static_cast<ncclComm_t>(comm)));
  return OSC_OK;
}

// ---------------------------------------------------------------- halo plan (OSC_HALO_PULL)
// flags[j] = 1 for every remote row j referenced by a local neighbour list (or an extra column list)
__global__ void halo_mark_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t row0, int64_t n_local,
                                 int32_t* __restrict__ flags) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = ids[e];
    if (j >= 0 && (j < row0 || j >= row0 + n_local)) flags[j] = 1;
  }
}
// halo_rows[pos[j]] = j for the marked rows (pos = exclusive scan of flags): ascending, i.e. grouped by owner
__global__ void halo_list_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ pos, int64_t N,
                                 int32_t* __restrict__ halo_rows, int64_t cap) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (int64_t)gridDim.x * blockDim.x)
    if (flags[j] && pos[j] < cap) halo_rows[pos[j]] = (int32_t)j;
}
// ids -> rows of the block [own rows (shard slots) | halo rows]
__global__ void halo_remap_kernel(const int32_t* __restrict__ ids, int64_t n, int64_t row0, int64_t n_local,
                                  int64_t shard, const int32_t* __restrict__ pos, int32_t* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = ids[e];
    int32_t r = j;
    if (j >= 0) r = (j >= row0 && j < row0 + n_local) ? (int32_t)(j - row0) : (int32_t)(shard + pos[j]);
    out[e] = r;
  }
}

static int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

int dist_halo_plan_workspace(int64_t N, size_t* bytes) {
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(N + 1));
  *bytes = 2 * align_up((size_t)(N + 1) * sizeof(int32_t)) + align_up(scan) + 1024;
  return OSC_OK;
}

int dist_halo_plan(const int32_t* nbr_loc, int64_t n_local, int k, const int32_t* extra_ids, int64_t n_extra,
                   int64_t N, int64_t row0, int64_t shard, int32_t* halo_rows, int64_t halo_cap,
                   int32_t* nbr_out, int32_t* extra_out, int64_t* h_n_halo, void* workspace, size_t ws_bytes,
                   cudaStream_t st) {
  OSC_REQUIRE(N >= 1 && N < 2147483647LL && n_local >= 0 && k >= 1 && h_n_halo != nullptr, "halo_plan: bad shape");
  size_t need = 0;
  dist_halo_plan_workspace(N, &need);
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "halo_plan: workspace too small");
  Arena ar(workspace, ws_bytes);
  int32_t* flags = ar.take<int32_t>((size_t)N + 1);
  int32_t* pos = ar.take<int32_t>((size_t)N + 1);
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, flags, pos, (int)(N + 1));
  void* scan_ws = ar.take<char>(scan);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "halo_plan: workspace too small");
  OSC_CUDA(cudaMemsetAsync(flags, 0, (size_t)(N + 1) * sizeof(int32_t), st));
  const int64_t n_ids = n_local * k;
  if (n_ids > 0) halo_mark_kernel<<<ew_blocks(n_ids), 256, 0, st>>>(nbr_loc, n_ids, row0, n_local, flags);
  if (n_extra > 0) halo_mark_kernel<<<ew_blocks(n_extra), 256, 0, st>>>(extra_ids, n_extra, row0, n_local, flags);
  OSC_LAUNCH_CHECK("halo_mark_kernel");
  OSC_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan, flags, pos, (int)(N + 1), st));  // pos[N] = count
  int32_t n_halo32 = 0;
  OSC_CUDA(cudaMemcpyAsync(&n_halo32, pos + N, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OSC_CUDA(cudaStreamSynchronize(st));
  *h_n_halo = n_halo32;
  if (halo_rows == nullptr || n_halo32 > halo_cap) return OSC_OK;  // size query (or retry with more room)
  halo_list_kernel<<<ew_blocks(N), 256, 0, st>>>(flags, pos, N, halo_rows, halo_cap);
  if (n_ids > 0 && nbr_out != nullptr)
    halo_remap_kernel<<<ew_blocks(n_ids), 256, 0, st>>>(nbr_loc, n_ids, row0, n_local, shard, pos, nbr_out);
  if (n_extra > 0 && extra_out != nullptr)
    halo_remap_kernel<<<ew_blocks(n_extra), 256, 0, st>>>(extra_ids, n_extra, row0, n_local, shard, pos, extra_out);
  OSC_LAUNCH_CHECK("halo_remap_kernel");
  OSC_CUDA(cudaStreamSynchronize(st));  // flags/pos live in the caller's workspace
  return OSC_OK;
}

// One warp per halo row: copy row halo_rows[h] from its owner's peer-mapped block into slot shard + h of
// the local block.  The list is ascending, so consecutive warps read consecutive (or nearby) rows of one
// peer: the NVLink traffic is long runs, not random 4*D-byte fetches.
template <int VEC>
__global__ void __launch_bounds__(256)
halo_pull_kernel(const float* const* __restrict__ peers, const int32_t* __restrict__ halo_rows, int64_t n_halo,
                 int64_t rot, int64_t shard, int D, float* __restrict__ block, const int* __restrict__ done) {
  if (done != nullptr && *done != 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t v = w0; v < n_halo; v += nwarps) {
    // the list is ordered by owner; rank r walks it from its first row above the own block and wraps:
    // owners r+1, ..., G-1, 0, ..., r-1 -- a different source for every rank at every moment
    int64_t h = v + rot;
    if (h >= n_halo) h -= n_halo;
    const int64_t j = halo_rows[h];
    const int64_t g = j / shard;
    const float* src = peers[g] + (j - g * shard) * D;
    float* dst = block + (shard + h) * D;
    if constexpr (VEC == 4) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int c = lane; c < D / 4; c += 32) d4[c] = s4[c];
    } else {
      for (int c = lane; c < D; c += 32) dst[c] = src[c];
    }
  }
}

static int halo_pull(const osc_dist_t* ds, int D, const int* done, cudaStream_t st) {
  if (ds->n_halo == 0) return OSC_OK;
  const int64_t warps_wanted = ds->n_halo;
  int64_t blocks = (warps_wanted + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const bool v4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(ds->P_block) & 15) == 0);
  int64_t rot = ds->halo_below;
  if (rot < 0 || rot >= ds->n_halo) rot = 0;
  if (v4)
    halo_pull_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(ds->d_peer_P, ds->halo_rows, ds->n_halo, rot, ds->shard, D,
                                                          ds->P_block, done);
  else
    halo_pull_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(ds->d_peer_P, ds->halo_rows, ds->n_halo, rot, ds->shard, D,
                                                          ds->P_block, done);
  OSC_LAUNCH_CHECK("halo_pull_kernel");
  return OSC_OK;
}

// ---------------------------------------------------------------- the distributed solve
struct DistBufs {
  float *R, *P, *AP, *P_all, *rz0, *pair[2], *pap, *d_res, *flag;
  double *part_a, *part_b, *col64, *total;
  PcgCtl* ctl;
};

static bool rows_mode(const osc_dist_t* ds) { return ds->partition == OSC_PART_ROWS; }
static bool pull_mode(const osc_dist_t* ds) { return rows_mode(ds) && ds->world > 1 && ds->halo == OSC_HALO_PULL; }
static bool gather_mode(const osc_dist_t* ds) { return rows_mode(ds) && ds->world > 1 && ds->halo != OSC_HALO_PULL; }

static int dist_check(const osc_dist_t* ds, const osc_graph_t* g, int D) {
  OSC_REQUIRE(ds != nullptr && g != nullptr, "dist: NULL argument");
  OSC_REQUIRE(ds->world >= 1 && ds->rank >= 0 && ds->rank < ds->world, "dist: bad world/rank");
  OSC_REQUIRE(ds->partition == OSC_PART_ROWS || ds->partition == OSC_PART_COLUMNS, "dist: bad partition");
  OSC_REQUIRE(ds->world == 1 || ds->nccl_comm != nullptr, "dist: NULL communicator");
  OSC_REQUIRE(g->batch == 1 && D >= 1, "dist: one lattice, D >= 1");
  if (rows_mode(ds)) {
    OSC_REQUIRE(ds->shard >= 1 && ds->shard * ds->world >= ds->N, "dist: shard * world < N");
    OSC_REQUIRE(g->N <= ds->shard, "dist: local rows exceed the shard size");
  } else {
    OSC_REQUIRE(g->N == ds->N, "dist: the column-slab partition holds all rows on every rank");
  }
  if (pull_mode(ds))
    OSC_REQUIRE(ds->d_peer_P != nullptr && ds->P_block != nullptr && (ds->n_halo == 0 || ds->halo_rows != nullptr) &&
                    ds->halo_nbr != nullptr,
                "dist: OSC_HALO_PULL needs the peer table, the block and the halo plan");
  return OSC_OK;
}

static size_t dist_ws_bytes(const osc_dist_t* ds, int64_t n_loc, int D, int n_blocks) {
  const size_t vec = align_up((size_t)n_loc * D * sizeof(float));
  const size_t vecp = pull_mode(ds) ? 0 : align_up((size_t)(rows_mode(ds) ? ds->shard : n_loc) * D * sizeof(float));
  const size_t part = align_up((size_t)n_blocks * D * sizeof(double));
  const size_t col = align_up((size_t)D * sizeof(float));
  size_t b = 2 * vec + vecp + 2 * part + 6 * col + 2 * align_up((size_t)D * sizeof(double)) + 4096;
  if (gather_mode(ds)) b += align_up((size_t)ds->world * ds->shard * D * sizeof(float));
  return b;
}

int dist_pcg_workspace(const osc_dist_t* ds, int64_t n_loc, int D, size_t* bytes) {
  OSC_REQUIRE(ds != nullptr && bytes != nullptr, "dist: NULL argument");
  osc_pcg_dims_t d{ds->N, 0, n_loc, D, 0};
  int rc = pcg_plan(&d, nullptr);
  if (rc) return rc;
  *bytes = dist_ws_bytes(ds, n_loc, D, d.n_blocks);
  return OSC_OK;
}

static int dist_take(const osc_dist_t* ds, int64_t n_loc, int D, int n_blocks, Arena& ar, DistBufs& b) {
  const size_t nd = (size_t)n_loc * D;
  b.R = ar.take<float>(nd);
  b.AP = ar.take<float>(nd);
  // the search direction: padded to `shard` rows (equal all-gather counts); OSC_HALO_PULL keeps it in the
  // caller's peer-mapped block instead
  b.P = pull_mode(ds) ? ds->P_block : ar.take<float>((size_t)(rows_mode(ds) ? ds->shard : n_loc) * D);
  b.P_all = gather_mode(ds) ? ar.take<float>((size_t)ds->world * ds->shard * D) : nullptr;
  b.part_a = ar.take<double>((size_t)n_blocks * D);
  b.part_b = ar.take<double>((size_t)n_blocks * D);
  b.rz0 = nullptr;
  b.pair[0] = ar.take<float>(2 * (size_t)D);  // [rr | rz] of even iterations (iteration 0: rz of the start)
  b.pair[1] = ar.take<float>(2 * (size_t)D);
  b.pap = ar.take<float>(D);
  b.col64 = ar.take<double>(D);
  b.total = ar.take<double>(8);
  b.d_res = ar.take<float>(32);
  b.flag = ar.take<float>(32);
  b.ctl = ar.take<PcgCtl>(1);
  return ar.ok ? OSC_OK : fail(OSC_ERR_WORKSPACE, "dist: workspace too small");
}

// make the rows of `src_loc` (this rank's block of the gathered vector, already in b.P) reachable by the
// SpMM of every rank; returns the view to gather from
static int dist_expose(const osc_dist_t* ds, NcclApi* api, const DistBufs& b, int64_t n_loc, int D,
                       const int* done, cudaStream_t st, VecView* vv) {
  ncclComm_t comm = static_cast<ncclComm_t>(ds->nccl_comm);
  if (gather_mode(ds)) {
    OSC_NCCL(api, api->AllGather(b.P, b.P_all, (size_t)ds->shard * D, ncclFloat, comm, st));
    *vv = VecView{b.P_all, nullptr, 0, 0};
  } else if (pull_mode(ds)) {
    // every rank's block is written before any peer reads it: a stream-ordered 1-float all-reduce
    OSC_NCCL(api, api->AllReduce(b.flag, b.flag, 1, ncclFloat, ncclMax, comm, st));
    int rc = halo_pull(ds, D, done, st);
    if (rc) return rc;
    *vv = VecView{b.P, nullptr, 0, 1};
  } else {
    *vv = VecView{b.P, nullptr, 0, 0};  // one GPU, or column slabs: all rows are local
  }
  (void)n_loc;
  return OSC_OK;
}

int dist_pcg_solve(const osc_dist_t* ds, const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                   int mode, float dt, int warm, float inertia, int jacobi, double tol, int max_iters,
                   const float* Y, const float* U, const float* psi, const float* gates, int D, float* X,
                   int* h_iters, float* h_res, void* workspace, size_t ws_bytes, cudaStream_t st) {
  int rc = dist_check(ds, g, D);
  if (rc) return rc;
  OSC_REQUIRE(prm != nullptr && Y != nullptr && X != nullptr && psi != nullptr, "dist_pcg_solve: NULL argument");
  NcclApi* api = nullptr;
  if (ds->world > 1 && (api = nccl_api()) == nullptr) return OSC_ERR_UNSUPPORTED;
  ncclComm_t comm = static_cast<ncclComm_t>(ds->nccl_comm);
  const bool rows = rows_mode(ds), multi = ds->world > 1;
  const int64_t n_loc = g->N;
  if (U == nullptr) U = Y;
  osc_pcg_dims_t d{ds->N, rows ? (int64_t)ds->rank * ds->shard : 0, n_loc, D, 0};
  if ((rc = pcg_plan(&d, nullptr))) return rc;
  if (ws_bytes < dist_ws_bytes(ds, n_loc, D, d.n_blocks)) return fail(OSC_ERR_WORKSPACE, "dist_pcg_solve: workspace too small");
  if (h_iters) *h_iters = 0;
  if (h_res) *h_res = __builtin_nanf("");
  Arena ar(workspace, ws_bytes);
  DistBufs b;
  if ((rc = dist_take(ds, n_loc, D, d.n_blocks, ar, b))) return rc;
  CtlPoll* poll = ctl_poll();
  if (poll == nullptr) return OSC_ERR_CUDA;
  const size_t nd = (size_t)n_loc * D;
  osc_graph_t gl = *g;
  if (pull_mode(ds)) gl.nbr = ds->halo_nbr;  // neighbour ids as rows of the block
  const int* done = &b.ctl->done;
  OSC_CUDA(cudaMemsetAsync(b.ctl, 0, sizeof(PcgCtl), st));
  OSC_CUDA(cudaMemsetAsync(b.flag, 0, sizeof(float), st));
  if (gather_mode(ds) && n_loc < ds->shard)  // the pad rows of the last rank travel with the all-gather
    OSC_CUDA(cudaMemsetAsync(b.P + nd, 0, (size_t)(ds->shard - n_loc) * D * sizeof(float), st));

  // x0 and right-hand side (lattice.py:171,184 / :245, :751-758).  All rows local (one GPU, column slabs) and a
  // start vector that is Y or U itself: no setup pass -- the first residual forms the right-hand side in place
  // and gathers the start vector where it lies (pcg_solve does the same).
  const bool settle = mode == OSC_MODE_SETTLE;
  const float* x0 = (!settle || !warm) ? Y : U;
  const bool fused = !(multi && rows) && max_iters >= 1 && ds->N > 0 && !(settle && warm && inertia > 0.f) &&
                     pcg_fuse_x() && x0 != X && pcg_fused_init_ok(&d, &gl);
  const InitSrc init{Y, U, psi, x0 == Y ? 1 : 0, x0 == U ? 1 : 0};
  if (!fused && (rc = pcg_setup(&d, prm, mode, dt, warm, inertia, Y, U, psi, gates, X, b.R, st))) return rc;
  if (max_iters < 1 || ds->N == 0) return OSC_OK;

  // ---- r0 = b - A x0 ; p0 = z0 ; rz
  VecView vv;
  float* rz = b.pair[0] + D;
  if (multi && rows) {
    // x0's remote rows travel like p's: stage the local block in P, expose it, then P receives z0
    if (nd) OSC_CUDA(cudaMemcpyAsync(b.P, X, nd * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = dist_expose(ds, api, b, n_loc, D, nullptr, st, &vv))) return rc;
    float* z0 = pull_mode(ds) ? b.AP : b.P;  // pull: peers may still be reading this rank's block
    if ((rc = spmm_launch(true, &d, &gl, chain, prm, mode, dt, jacobi, gates, vv, b.R, z0, b.part_a, st))) return rc;
    if ((rc = pcg_reduce(b.part_a, d.n_blocks, D, rz, nullptr, nullptr, st))) return rc;
    OSC_NCCL(api, api->AllReduce(rz, rz, (size_t)D, ncclFloat, ncclSum, comm, st));
    // (the all-reduce completes only after every rank has issued its own, i.e. finished pulling)
    if (pull_mode(ds) && nd) OSC_CUDA(cudaMemcpyAsync(b.P, b.AP, nd * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    vv = VecView{fused ? x0 : X, nullptr, 0, 0};
    if ((rc = spmm_launch(true, &d, &gl, chain, prm, mode, dt, jacobi, gates, vv, b.R, b.P, b.part_a, st, nullptr,
                          fused ? &init : nullptr)))
      return rc;
    if ((rc = pcg_reduce(b.part_a, d.n_blocks, D, rz, nullptr, nullptr, st))) return rc;
  }

  // ---- iterations.  lag: how many iterations the host runs ahead of the device-side verdict
  const int lag = gather_mode(ds) ? 0 : 1;
  const bool fuse_x = pcg_fuse_x();
  PcgCtl h{0, 0, __builtin_nanf(""), 0};
  int it = 0;
  for (it = 1; it <= max_iters; ++it) {
    float* rr = b.pair[it & 1];
    float* rz_new = rr + D;
    if ((rc = dist_expose(ds, api, b, n_loc, D, done, st, &vv))) return rc;
    if ((rc = spmm_launch(false, &d, &gl, chain, prm, mode, dt, 1, gates, vv, b.AP, nullptr, b.part_a, st, done)))
      return rc;
    if ((rc = pcg_reduce(b.part_a, d.n_blocks, D, b.pap, nullptr, nullptr, st, done))) return rc;
    if (multi && rows) OSC_NCCL(api, api->AllReduce(b.pap, b.pap, (size_t)D, ncclFloat, ncclSum, comm, st));
    if ((rc = pcg_update(&d, prm, mode, dt, jacobi, gates, rz, b.pap, b.P, b.AP, fuse_x ? nullptr : X, b.R, b.part_a,
                         b.part_b, st, done)))
      return rc;
    if ((rc = pcg_reduce(b.part_a, d.n_blocks, D, rr, b.d_res, nullptr, st, done))) return rc;
    if ((rc = pcg_reduce(b.part_b, d.n_blocks, D, rz_new, nullptr, nullptr, st, done))) return rc;
    if (multi && rows) {
      OSC_NCCL(api, api->AllReduce(rr, rr, 2 * (size_t)D, ncclFloat, ncclSum, comm, st));  // rr | rz_new
      if ((rc = pcg_decide(b.ctl, rr, nullptr, D, tol, it, max_iters, st))) return rc;
    } else {
      if (multi) OSC_NCCL(api, api->AllReduce(b.d_res, b.d_res, 1, ncclFloat, ncclMax, comm, st));
      if ((rc = pcg_decide(b.ctl, nullptr, b.d_res, D, tol, it, max_iters, st))) return rc;
    }
    if ((rc = poll->record(it, b.ctl, st))) return rc;
    // x += alpha p (always part of this iteration) rides with p = z + beta p (skipped by the last one)
    if (fuse_x &&
        (rc = pcg_pupdate_x(&d, prm, mode, dt, jacobi, gates, rz_new, rz, b.pap, b.R, b.P, X, b.ctl, it, 0, st,
                            (fused && it == 1) ? x0 : nullptr)))
      return rc;
    if (it > lag) {
      if ((rc = poll->wait(it - lag, &h))) return rc;
      if (h.done) break;
    }
    if (it == max_iters) break;
    if (!fuse_x && (rc = pcg_pupdate(&d, prm, mode, dt, jacobi, gates, rz_new, rz, b.R, b.P, st, done))) return rc;
    rz = rz_new;
  }
  if (it > max_iters) it = max_iters;
  if (!h.done && (rc = poll->wait(it, &h))) return rc;
  if (h_iters) *h_iters = h.iters;
  if (h_res) *h_res = h.res;
  return OSC_OK;
}

// the halo exchange of OSC_HALO_PULL alone (cross-rank ordering point + pull), for measurement
int dist_halo_exchange(const osc_dist_t* ds, int D, float* d_flag, cudaStream_t st) {
  OSC_REQUIRE(ds != nullptr && pull_mode(ds) && ds->nccl_comm != nullptr && d_flag != nullptr,
              "dist_halo_exchange: needs a rows partition with OSC_HALO_PULL and a communicator");
  NcclApi* api = nccl_api();
  if (api == nullptr) return OSC_ERR_UNSUPPORTED;
  OSC_NCCL(api, api->AllReduce(d_flag, d_flag, 1, ncclFloat, ncclMax, static_cast<ncclComm_t>(ds->nccl_comm), st));
  return halo_pull(ds, D, nullptr, st);
}

// deltaH = <U - U*, M (U - U*)> summed over the ranks (receipts.py:21-25)
int dist_delta_h(const osc_dist_t* ds, const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                 const float* U, const float* Ustar, const float* gates, int D, double* h_out, void* workspace,
                 size_t ws_bytes, cudaStream_t st) {
  int rc = dist_check(ds, g, D);
  if (rc) return rc;
  OSC_REQUIRE(prm != nullptr && U != nullptr && Ustar != nullptr && h_out != nullptr, "dist_delta_h: NULL argument");
  NcclApi* api = nullptr;
  if (ds->world > 1 && (api = nccl_api()) == nullptr) return OSC_ERR_UNSUPPORTED;
  ncclComm_t comm = static_cast<ncclComm_t>(ds->nccl_comm);
  const bool rows = rows_mode(ds), multi = ds->world > 1;
  const int64_t n_loc = g->N;
  osc_pcg_dims_t d{ds->N, rows ? (int64_t)ds->rank * ds->shard : 0, n_loc, D, 0};
  if ((rc = pcg_plan(&d, nullptr))) return rc;
  if (ws_bytes < dist_ws_bytes(ds, n_loc, D, d.n_blocks)) return fail(OSC_ERR_WORKSPACE, "dist_delta_h: workspace too small");
  *h_out = 0.0;
  Arena ar(workspace, ws_bytes);
  DistBufs b;
  if ((rc = dist_take(ds, n_loc, D, d.n_blocks, ar, b))) return rc;
  const size_t nd = (size_t)n_loc * D;
  osc_graph_t gl = *g;
  if (pull_mode(ds)) gl.nbr = ds->halo_nbr;
  OSC_CUDA(cudaMemsetAsync(b.flag, 0, sizeof(float), st));
  if (gather_mode(ds) && n_loc < ds->shard)
    OSC_CUDA(cudaMemsetAsync(b.P + nd, 0, (size_t)(ds->shard - n_loc) * D * sizeof(float), st));
  if ((rc = launch_diff(U, Ustar, b.P, (int64_t)nd, st))) return rc;
  VecView vv;
  if ((rc = dist_expose(ds, api, b, n_loc, D, nullptr, st, &vv))) return rc;
  if ((rc = spmm_launch(false, &d, &gl, chain, prm, OSC_MODE_STATIONARY, 0.f, 1, gates, vv, b.AP, nullptr,
                        b.part_a, st)))
    return rc;
  if ((rc = pcg_reduce(b.part_a, d.n_blocks, D, b.pap, nullptr, b.col64, st))) return rc;
  if ((rc = launch_sum_doubles(b.col64, D, b.total, st))) return rc;
  if (multi) OSC_NCCL(api, api->AllReduce(b.total, b.total, 1, ncclDouble, ncclSum, comm, st));
  OSC_CUDA(cudaMemcpyAsync(h_out, b.total, sizeof(double), cudaMemcpyDeviceToHost, st));
  OSC_CUDA(cudaStreamSynchronize(st));
  return OSC_OK;
}

}  // namespace osc
