/*
 * oscillink_b200.h -- C ABI of the B200-native lattice-settle hot path.
 *
 * The reference (Maverick0351a/Oscillink) has NO FFI: its boundary is the Python class
 * `OscillinkLattice` (oscillink/core/lattice.py:23) calling NumPy.  Each entry point
 * below replaces the NumPy expression(s) cited next to it; the Python host mirror
 * (oscillink_b200/lattice.py) keeps the class surface unchanged and binds these
 * symbols with ctypes (see INTEGRATION.md for the stub a reference maintainer adds).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - Every call returns an int status (OSC_OK == 0); osc_last_error() gives the text of
 *     the last failure on the calling thread.  No exceptions cross the boundary.
 *   - Pointers are DEVICE pointers unless the name starts with h_ (host).
 *   - The caller owns every buffer, including the workspace (query the size first).
 *   - Every launch goes to the cudaStream_t passed in (as void*); calls are re-entrant
 *     across host threads (no global mutable state).
 *   - All arrays are row-major fp32 / int32.  `batch` independent lattices are stored
 *     back to back ([batch][N][D], [batch][N][k], ...); a single lattice is batch == 1.
 *
 * Graph layout (ELL, width k): nbr[N][k] neighbour column ascending, -1 padded on the
 * right; A[N][k] capped adjacency weights (graph.py:77-83); W[N][k] normalised weights
 * A_ij/(sd_i*sd_j) (graph.py:87-90); deg[N]; sqrt_deg[N].
 */
#ifndef OSCILLINK_B200_H
#define OSCILLINK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSC_OK 0
#define OSC_ERR_INVALID 1     /* bad argument (maps to Python ValueError) */
#define OSC_ERR_CUDA 2        /* CUDA runtime / launch failure */
#define OSC_ERR_WORKSPACE 3   /* workspace too small */
#define OSC_ERR_UNSUPPORTED 4 /* shape outside what the kernels cover */

#define OSC_ABI_VERSION 1

/* kNN engine selector for osc_knn_candidates (flags bits 0-2) */
#define OSC_KNN_AUTO 0
#define OSC_KNN_SIMT 1 /* fp32 CUDA-core tile kernel (any shape) */
#define OSC_KNN_TC 2   /* tcgen05 3xTF32 tensor-core kernel (sm_100a) */
#define OSC_KNN_TC1 3  /* tcgen05 single-product TF32 kernel: a third of the tensor work, scores only \
                          pre-select candidates (error bound OSC_KNN_EPS_TC1, wider candidate lists) */
#define OSC_KNN_TCH 4  /* tcgen05 single-product fp16 kernel: same 11-bit operands and error bound as TC1, \
                          half the operand bytes and twice the MMA rate; needs D % 8 == 0.  The q_hi /  \
                          all_hi arguments then point at fp16 arrays (osc_normalize_rows_f16) */

/* solve modes for osc_pcg_* / osc_batched_* */
#define OSC_MODE_SETTLE 0     /* (I + dt M) X = U + dt RHS   lattice.py:170-192 */
#define OSC_MODE_STATIONARY 1 /* M X = RHS, x0 = Y           lattice.py:245-259 */

typedef struct osc_graph {
  int64_t batch;
  int64_t N;
  int32_t k;
  int32_t _pad;
  const int32_t* nbr;
  const float* A;
  const float* W;
  const int32_t* deg;
  const float* sqrt_deg;
} osc_graph_t;

/* Chain prior (graph.py:96-111, lattice.py:129-149), CSR over the distinct chain nodes.
 * n_rows == 0 or a NULL osc_chain_t* means "no chain".  slot[N] maps a lattice row to its
 * index in rows[] (or -1).  Only valid for batch == 1. */
typedef struct osc_chain {
  int32_t n_rows;
  int32_t nnz;
  const int32_t* rows;
  const int32_t* rowptr;
  const int32_t* col;
  const float* Wp; /* Ap_uv/(sdp_u*sdp_v) */
  const float* Ap; /* raw max-merged path weights */
  const int32_t* slot;
} osc_chain_t;

/* lambda parameters; chain_present mirrors `L_path is not None` (lattice.py:180,190):
 * the operator gets the lamP term only if chain_present && lamP > 0, the Jacobi diagonal
 * gets +lamP whenever chain_present. */
typedef struct osc_params {
  float lamG, lamC, lamQ, lamP;
  int32_t chain_present;
  int32_t _pad;
} osc_params_t;

/* ------------------------------------------------------------------ misc */
int osc_abi_version(void);
const char* osc_last_error(void);
/* sm count, shared memory per block (opt-in), compute capability major*10+minor */
int osc_device_info(int device, int* h_sm_count, int* h_smem_optin, int* h_cc);

/* ------------------------------------------------------------------ K1: kNN build
 * Replaces oscillink/core/graph.py:29-62 (normalise, S = Yn Yn^T, diag = -inf, top-k). */

/* Yn = Y / (||Y||_2 + 1e-12) row-wise (graph.py:35).  Yn_hi/Yn_lo (optional, may be NULL)
 * receive the TF32 split used by the tensor-core kernels: hi = tf32(Yn), lo = tf32(Yn - hi);
 * OSC_KNN_TC1 reads hi only (pass Yn_lo = NULL). */
int osc_normalize_rows(const float* Y, int64_t rows, int32_t D, float* Yn, float* Yn_hi,
                       float* Yn_lo, void* stream);
/* Same normalisation; Yn_h16 ([rows][D] IEEE fp16, round-to-nearest of Yn) feeds OSC_KNN_TCH. */
int osc_normalize_rows_f16(const float* Y, int64_t rows, int32_t D, float* Yn, void* Yn_h16,
                           void* stream);

/* Candidate pass.  For each of the `n_rows` query rows (global ids row0..row0+n_rows-1 of
 * every lattice in the batch) keep the kc best columns of Yn_q . Yn_all^T by (approximate
 * similarity desc, column asc), self column excluded.  Yn_q may alias Yn_all + row0*D.
 * cand_idx/cand_sim: [batch][n_rows][kc].  flags: OSC_KNN_*.  The TC engine needs the
 * hi/lo split arrays (same shapes as Yn_q / Yn_all); the SIMT engine ignores them. */
int osc_knn_candidates(const float* Yn_q, const float* Yn_all, const float* q_hi, const float* q_lo,
                       const float* all_hi, const float* all_lo, int64_t batch, int64_t n_rows,
                       int64_t row0, int64_t N, int32_t D, int32_t kc, int32_t flags,
                       int32_t* cand_idx, float* cand_sim, void* workspace, size_t ws_bytes,
                       void* stream);
/* 1 if the tcgen05 engine covers (N, D, kc) on the current device */
int osc_knn_tc_supported(int64_t N, int32_t D, int32_t kc);
/* Resolve `flags` (OSC_KNN_*, AUTO included) for n_rows query rows against N columns into the
 * engine that osc_knn_build would run, the candidate-list width kc that engine needs and the
 * bound eps on |approximate - exact| to hand to osc_knn_rescore_checked.  Any output may be NULL. */
int osc_knn_plan(int64_t n_rows, int64_t N, int32_t D, int32_t k, int32_t flags, int32_t* h_engine,
                 int32_t* h_kc, float* h_eps);
int osc_knn_candidates_workspace(int64_t batch, int64_t n_rows, int64_t N, int32_t D, int32_t kc,
                                 int32_t flags, size_t* h_bytes);

/* Canonical ranking (graph.py:46-52): re-score every candidate with an fp64-accumulated dot
 * of the fp32 rows, round once to fp32, order by (similarity desc, column asc), keep k.
 * top_idx/top_sim: [batch][n_rows][k].  gap (optional): k-th minus (k+1)-th similarity. */
int osc_knn_rescore(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows, int64_t N,
                    int32_t D, const int32_t* cand_idx, int32_t kc, int32_t k, int32_t* top_idx,
                    float* top_sim, float* gap, void* stream);

/* Checked re-scoring: as osc_knn_rescore, plus a completeness test of every row's candidate list.
 * A column outside the list scored <= a_min (smallest approximate score kept) in the approximate
 * pass, so it can only be a true top-k column if a_min + eps >= the exact k-th score; such rows are
 * recomputed exhaustively (all N columns, same fp64-accumulated dot, graph.py:46-52 ordering).
 * The same bound prunes the work: cand_sim is sorted descending, k candidates score >= cand_sim[k-1]
 * approximately, hence >= cand_sim[k-1] - eps exactly, so a candidate below cand_sim[k] - 2 eps can
 * be neither a top-k column nor the (k+1)-th (the `gap` output) and its row is never fetched.
 * eps bounds |approximate - exact| of the candidate engine (OSC_KNN_EPS covers 3xTF32 with
 * truncating accumulation and the fp32 FMA engine).  *d_n_flagged (device int32) receives the
 * number of rows that took the exhaustive path.  row0 = global id of the first query row. */
#define OSC_KNN_EPS 1e-5f
/* single-product engines: both operands rounded to 11 significant bits => |error| <= 2^-10 * sum|a_i b_i|
 * <= 2^-10 for unit rows (Cauchy-Schwarz); the packed-key top-k lists used for N <= 2048 keep 21 bits
 * of a score (<= 2^-12 more); plus the accumulation noise covered by OSC_KNN_EPS */
#define OSC_KNN_EPS_TC1 1.25e-3f
int osc_knn_rescore_workspace(int64_t batch, int64_t n_rows, size_t* h_bytes);
int osc_knn_rescore_checked(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows,
                            int64_t row0, int64_t N, int32_t D, const int32_t* cand_idx,
                            const float* cand_sim, int32_t kc, int32_t k, float eps, int32_t* top_idx,
                            float* top_sim, float* gap, int32_t* d_n_flagged, void* workspace,
                            size_t ws_bytes, void* stream);

/* As osc_knn_rescore_checked with a bound on the exhaustive path: if MORE than exhaustive_limit rows are
 * flagged (exhaustive_limit < 0: no bound) the exhaustive scans are skipped, the flagged rows keep the
 * result of their (possibly incomplete) candidate list, and the caller -- who reads *d_n_flagged -- re-runs
 * the candidate pass with a tighter engine (OSC_KNN_TC, eps = OSC_KNN_EPS).  This is what keeps the
 * single-product engines safe on clustered / near-duplicate anchors, where many rows have neighbours
 * closer than their error bound.  osc_knn_exhaustive_limit(rows) = max(64, rows / 100) is the bound the
 * library's own build flows use. */
int osc_knn_rescore_guarded(const float* Yn_q, const float* Yn_all, int64_t batch, int64_t n_rows,
                            int64_t row0, int64_t N, int32_t D, const int32_t* cand_idx,
                            const float* cand_sim, int32_t kc, int32_t k, float eps, int64_t exhaustive_limit,
                            int32_t* top_idx, float* top_sim, float* gap, int32_t* d_n_flagged,
                            void* workspace, size_t ws_bytes, void* stream);
int64_t osc_knn_exhaustive_limit(int64_t rows);

/* K1b -- graph.py:50-52 (S>0 filter), :64-65 (mutual), :77-83 (row cap), :87-92 (degree,
 * normalised weights).  top_idx/top_sim are the directed lists of ALL N rows of each lattice.
 * scratch: N*batch floats.  nnz: [batch] int64 (device). */
int osc_graph_assemble(const int32_t* top_idx, const float* top_sim, int64_t batch, int64_t N,
                       int32_t k, float row_cap, int32_t* nbr, float* A, float* W, int32_t* deg,
                       float* sqrt_deg, int64_t* nnz, float* scratch, void* stream);

/* Whole build for `batch` lattices resident on one GPU (normalise + candidates + rescore +
 * assemble).  flags: OSC_KNN_*.  h_near_ties (optional): rows whose k/(k+1) gap < 4e-7. */
int osc_knn_build_workspace(int64_t batch, int64_t N, int32_t D, int32_t k, int32_t flags,
                            size_t* h_bytes);
int osc_knn_build(const float* Y, int64_t batch, int64_t N, int32_t D, int32_t k, float row_cap,
                  int32_t flags, int32_t* nbr, float* A, float* W, int32_t* deg, float* sqrt_deg,
                  int64_t* nnz, float* gap, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ K2: PCG (single lattice)
 * Replaces lattice.py:170-207 / :245-265 driving solver.py:15-37.
 * Phase entry points are exported so a multi-GPU driver can put NCCL all-reduces between
 * them; osc_pcg_solve composes them on one GPU.
 *
 * Sharded use: the local rows are [row0, row0+n_local); vectors passed as `*_all` are
 * indexed by GLOBAL row (N rows), vectors passed as `*_loc` by local row.  On one GPU
 * row0 = 0, n_local = N and both views coincide.  graph arrays are LOCAL rows with
 * GLOBAL neighbour ids. */
typedef struct osc_pcg_dims {
  int64_t N;       /* global rows */
  int64_t row0;    /* first local row */
  int64_t n_local; /* local rows */
  int32_t D;
  int32_t n_blocks; /* row blocks used for the partial sums (from osc_pcg_plan) */
} osc_pcg_dims_t;

/* fills dims->n_blocks and reports the workspace bytes osc_pcg_solve needs */
int osc_pcg_plan(osc_pcg_dims_t* dims, size_t* h_ws_bytes);
/* Largest ELL width k the SpMM kernels can launch for D columns (the graph chunk of a row block is staged
 * in shared memory).  Graph loaders that adopt a caller-supplied adjacency (lattice.py:709-713) check it. */
int osc_pcg_max_ell_width(int32_t D);

/* x0 (lattice.py:751-758) and right-hand side (lattice.py:171,184 / :245):
 * X_loc = Y | U | (1-w)Y + wU ; Bv_loc = U + dt*(lamG*Y + lamQ*b*psi^T)  (settle)
 *                               Bv_loc = lamG*Y + lamQ*b*psi^T            (stationary) */
int osc_pcg_setup(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                  int32_t warm_start, float inertia, const float* Y_loc, const float* U_loc,
                  const float* psi, const float* gates_loc, float* X_loc, float* Bv_loc,
                  void* stream);

/* ------------------------------------------------------------------ chain prior (a6)
 * Replaces build_path_laplacian (oscillink/core/graph.py:96-111) as called by add_chain
 * (oscillink/core/lattice.py:129-149).  HOST function (every pointer is a host pointer): a chain is
 * O(len) data, the N x N path Laplacian of the reference never exists.  Fills the arrays an osc_chain_t
 * points at (upload them to the device): rows[n_rows] ascending, rowptr[n_rows+1], col/Wp/Ap[nnz]
 * (columns ascending inside a row), slot[N].  weights: len-1 entries or NULL (all 1, lattice.py:129).
 * OSC_ERR_INVALID for len < 2 or an index outside [0, N) (lattice.py:135-142). */
int osc_chain_build_size(const int32_t* h_chain, int32_t len, int64_t N, int32_t* h_n_rows, int32_t* h_nnz);
int osc_chain_build(const int32_t* h_chain, int32_t len, const float* h_weights, int64_t N, int32_t* h_rows,
                    int32_t* h_rowptr, int32_t* h_col, float* h_Wp, float* h_Ap, int32_t* h_slot);

/* R = Bv - Aop(X); P = R/(Mdiag+1e-12) (or R); partial rz.  X_all is the gathered x0. */
int osc_pcg_residual0(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                      const osc_params_t* prm, int32_t mode, float dt, int32_t jacobi,
                      const float* gates_loc, const float* X_all, float* RBv_loc, float* P_loc,
                      double* part_rz, void* stream);

/* AP = Aop(P), partial pAp (solver.py:24-25).  P_all indexed by global row. */
int osc_pcg_spmm_dot(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                     const osc_params_t* prm, int32_t mode, float dt, const float* gates_loc,
                     const float* P_all, float* AP_loc, double* part_pap, void* stream);

/* Row-sharded SpMM with the halo exchange FUSED into the kernel (north_star "U halo rows are
 * exchanged per SpMM"): d_peer_* is a device array of `world` pointers, entry r = base of rank r's
 * block of `shard` rows ([shard][D] fp32) in rank r's HBM, mapped into this process with CUDA IPC.
 * Neighbour rows owned by another rank are read by plain loads over NVLink while the tile is being
 * computed -- no all-gather in front of the SpMM, no N x D staging buffer.  The caller orders the
 * phases across ranks (every rank's writes to its block complete before any peer's SpMM starts).
 * osc_enable_peer_access(peer) must have been called once per peer device in this process. */
int osc_enable_peer_access(int32_t peer_device);
/* Peer-mapped row blocks: osc_peer_alloc = cudaMalloc (zero-filled) + cudaIpcGetMemHandle (64-byte
 * handle, to be sent to the peer processes); osc_peer_open = cudaIpcOpenMemHandle with lazy peer
 * access from the CURRENT device of the calling process; osc_peer_close / osc_peer_free undo them. */
int osc_peer_alloc(size_t bytes, void** d_ptr, unsigned char* handle64);
int osc_peer_open(const unsigned char* handle64, void** d_ptr);
int osc_peer_close(void* d_ptr);
int osc_peer_free(void* d_ptr);
int osc_pcg_residual0_p2p(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                          const osc_params_t* prm, int32_t mode, float dt, int32_t jacobi,
                          const float* gates_loc, const float* const* d_peer_X, int32_t world, int64_t shard,
                          float* RBv_loc, float* P_loc, double* part_rz, void* stream);
int osc_pcg_spmm_dot_p2p(const osc_pcg_dims_t* dims, const osc_graph_t* g, const osc_chain_t* chain,
                         const osc_params_t* prm, int32_t mode, float dt, const float* gates_loc,
                         const float* const* d_peer_P, int32_t world, int64_t shard, float* AP_loc,
                         double* part_pap, void* stream);

/* column reduction of the row-block partials: out[D] (fp32) = sum_blocks part[b][D];
 * if h_or_d_max != NULL also writes max_c sqrt(out[c]) to *d_max (device float). */
int osc_pcg_reduce(const double* part, int32_t n_blocks, int32_t D, float* out, float* d_max,
                   void* stream);

/* x += alpha p; r -= alpha Ap; partial rr, partial rz' (solver.py:26-28,32-33);
 * alpha_c = rz_c/(pap_c + 1e-18).  X_loc == NULL: r and the partials only -- x is then updated by
 * osc_pcg_pupdate_x (what osc_pcg_solve / osc_dist_pcg_solve do: one vector stream less per iteration). */
int osc_pcg_update(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                   int32_t jacobi, const float* gates_loc, const float* rz, const float* pap,
                   const float* P_loc, const float* AP_loc, float* X_loc, float* R_loc,
                   double* part_rr, double* part_rz, void* stream);

/* p = z + beta p, beta_c = rz_new_c/(rz_old_c + 1e-18) (solver.py:34-35) */
int osc_pcg_pupdate(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                    int32_t jacobi, const float* gates_loc, const float* rz_new, const float* rz_old,
                    const float* R_loc, float* P_loc, void* stream);

/* x += alpha p (solver.py:25, alpha from rz_old / pap of the iteration just tested) and, unless `last`,
 * p = z + beta p (solver.py:33-35) in one pass over p.  Same operations on the same operands as
 * osc_pcg_update(X) + osc_pcg_pupdate: the iterates are bit-identical. */
int osc_pcg_pupdate_x(const osc_pcg_dims_t* dims, const osc_params_t* prm, int32_t mode, float dt,
                      int32_t jacobi, const float* gates_loc, const float* rz_new, const float* rz_old,
                      const float* pap, const float* R_loc, float* P_loc, float* X_loc, int32_t last,
                      void* stream);

/* Full single-GPU solve.  X receives the solution ([N][D]); *h_iters / *h_res the
 * iteration count and last max-column residual (never fails on non-convergence). */
int osc_pcg_solve(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                  int32_t mode, float dt, int32_t warm_start, float inertia, int32_t jacobi,
                  double tol, int32_t max_iters, const float* Y, const float* U, const float* psi,
                  const float* gates, int32_t D, float* X, int32_t* h_iters, float* h_res,
                  void* workspace, size_t ws_bytes, void* stream);

/* The same recurrences for a caller-supplied start vector and right-hand side: on entry X = x0
 * ([N][D]) and B = right-hand side (clobbered with the residual).  This is the solve behind
 * compute_diffusion_gates (oscillink/preprocess/diffusion.py:132-150): with lamG = gamma,
 * lamC = 1, lamQ = 0 the stationary operator is exactly L_sym + gamma*I.  Workspace from
 * osc_pcg_plan. */
int osc_pcg_solve_system(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                         int32_t mode, float dt, int32_t jacobi, double tol, int32_t max_iters,
                         const float* gates, int32_t D, float* X, float* B, int32_t* h_iters,
                         float* h_res, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ multi-GPU solve (configs #4/#5)
 * One process per GPU; these entry points take the rank's share of the lattice and an NCCL communicator
 * and run the WHOLE solve (solver.py:19-37 across ranks): NCCL is called from inside the library on the
 * caller's stream, the stop test (solver.py:29-31) is evaluated on the device after the all-reduce, and the
 * host only polls its verdict.  NCCL is bound at run time (dlopen libnccl.so.2; the copy already loaded in
 * the process is used), so the library itself has no NCCL link dependency.
 *
 * Partitions:
 *   OSC_PART_ROWS     rank r owns rows [r*shard, min(N,(r+1)*shard)); graph arrays are the LOCAL rows with
 *                     GLOBAL neighbour ids; Y/U/X/gates are local rows; D is the full width.  Per iteration:
 *                     halo exchange of the search direction + all-reduce (SUM) of the 3*D column dots.
 *       OSC_HALO_ALLGATHER  ncclAllGather of p in front of every SpMM.
 *       OSC_HALO_PULL       every rank pulls the unique remote rows its graph references from the peers'
 *                           blocks over NVLink peer memory (osc_peer_alloc / osc_peer_open mappings):
 *                           P_block = [shard own rows | n_halo pulled rows][D]; halo_rows / halo_nbr come from
 *                           osc_dist_halo_plan; a chain's `col` must then hold rows of the block as well
 *                           (pass them through osc_dist_halo_plan's extra ids).
 *   OSC_PART_COLUMNS  rank r owns ALL N rows of D columns (a slab of the lattice's width); the graph is the
 *                     full graph; psi is the slab of psi.  No halo, no dot all-reduce: every reduction of
 *                     solver.py:22-36 is per column; one 1-float MAX all-reduce per iteration. */
#define OSC_PART_ROWS 0
#define OSC_PART_COLUMNS 1
#define OSC_HALO_ALLGATHER 0
#define OSC_HALO_PULL 1

typedef struct osc_dist {
  void* nccl_comm; /* ncclComm_t (from osc_dist_comm_init, or any communicator of the same NCCL library) */
  int32_t world, rank;
  int32_t partition; /* OSC_PART_* */
  int32_t halo;      /* OSC_HALO_* (rows partition) */
  int64_t N;         /* global rows */
  int64_t shard;     /* rows per rank = ceil(N / world) (rows partition) */
  /* OSC_HALO_PULL only */
  const float* const* d_peer_P; /* device table [world]: base of every rank's P_block (own entry included) */
  float* P_block;               /* this rank's block [shard + n_halo][D] */
  const int32_t* halo_rows;     /* [n_halo] ascending global ids of the remote rows this rank gathers */
  const int32_t* halo_nbr;      /* [n_local][k] neighbour ids as rows of P_block (-1 padded) */
  int64_t n_halo;
  int64_t halo_below;           /* number of halo rows owned by lower ranks (ids < rank*shard): the pull starts
                                   with the rows of rank+1 and wraps around, so that at any moment every
                                   rank reads from a DIFFERENT peer (all ranks starting at peer 0 shared one
                                   GPU's NVLink egress: 212 GB/s per GPU instead of 660 on 8 GPUs) */
} osc_dist_t;

/* NCCL version of the bound library (OSC_ERR_UNSUPPORTED if NCCL cannot be loaded) */
int osc_dist_nccl_version(int32_t* h_version);
/* ncclGetUniqueId on rank 0 (128 bytes, to be sent to the other ranks by any means), ncclCommInitRank on the
 * CURRENT device of every rank, ncclCommDestroy. */
int osc_dist_unique_id(unsigned char* h_id128);
int osc_dist_comm_init(const unsigned char* h_id128, int32_t world, int32_t rank, void** h_comm);
int osc_dist_comm_destroy(void* comm);

/* Halo plan of the rows partition for OSC_HALO_PULL.  nbr_loc: the local rows' neighbour ids [n_local][k]
 * (global, -1 padded); extra_ids (optional): further global row ids the operator touches (a chain's `col`).
 * Outputs: *h_n_halo = number of distinct remote rows; halo_rows[n_halo] ascending; nbr_out / extra_out =
 * the same ids as rows of the block [shard own rows | halo rows] (local j -> j - row0, remote j -> shard +
 * position in halo_rows).  Call with halo_rows == NULL (or too small a halo_cap) to query *h_n_halo only. */
int osc_dist_halo_plan_workspace(int64_t N, size_t* h_bytes);
int osc_dist_halo_plan(const int32_t* nbr_loc, int64_t n_local, int32_t k, const int32_t* extra_ids,
                       int64_t n_extra, int64_t N, int64_t row0, int64_t shard, int32_t* halo_rows,
                       int64_t halo_cap, int32_t* nbr_out, int32_t* extra_out, int64_t* h_n_halo,
                       void* workspace, size_t ws_bytes, void* stream);

/* The halo exchange of OSC_HALO_PULL on its own (what osc_dist_pcg_solve runs in front of every SpMM): a
 * stream-ordered cross-rank ordering point (1-float all-reduce on d_flag) + the pull of the halo rows into
 * P_block.  Exported for measurement. */
int osc_dist_halo_exchange(const osc_dist_t* dist, int32_t D, float* d_flag, void* stream);

/* Whole distributed solve (replaces solver.py:19-36 + lattice.py:173-182 across ranks).  n_loc = g->N local
 * rows, D = local columns.  Arguments as osc_pcg_solve; every rank receives the same *h_iters / *h_res. */
int osc_dist_pcg_workspace(const osc_dist_t* dist, int64_t n_loc, int32_t D, size_t* h_bytes);
int osc_dist_pcg_solve(const osc_dist_t* dist, const osc_graph_t* g, const osc_chain_t* chain,
                       const osc_params_t* prm, int32_t mode, float dt, int32_t warm_start, float inertia,
                       int32_t jacobi, double tol, int32_t max_iters, const float* Y, const float* U,
                       const float* psi, const float* gates, int32_t D, float* X, int32_t* h_iters,
                       float* h_res, void* workspace, size_t ws_bytes, void* stream);
/* deltaH over the sharded state (receipts.py:21-25), summed over the ranks; workspace as above */
int osc_dist_delta_h(const osc_dist_t* dist, const osc_graph_t* g, const osc_chain_t* chain,
                     const osc_params_t* prm, const float* U, const float* Ustar, const float* gates,
                     int32_t D, double* h_deltaH, void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ K4: receipts
 * deltaH = <U-U*, M (U-U*)>  (receipts.py:21-25).  workspace from osc_pcg_plan. */
int osc_delta_h(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm,
                const float* U, const float* Ustar, const float* gates, int32_t D,
                double* h_deltaH, void* workspace, size_t ws_bytes, void* stream);

/* Per-node terms (receipts.py:40-59) and null points (receipts.py:70-82), one pass.
 * coh/anchor/query: [N]; null_j[N] (-1 = none), null_z[N], null_R[N];
 * row_mu/row_sigma (optional): the per-row mean / std+1e-12 of R over all N columns. */
int osc_receipt_full(const osc_graph_t* g, const osc_params_t* prm, const float* Y,
                     const float* Ustar, const float* psi, const float* gates, int32_t D, float z_th,
                     float* coh, float* anchor, float* query, int32_t* null_j, float* null_z,
                     float* null_R, float* row_mu, float* row_sigma, void* stream);

/* ------------------------------------------------------------------ f1/f2: bundle, chain receipt
 * align_i = <U*_i/(||U*_i||+1e-12), psi/(||psi||+1e-12)>   (lattice.py:559-561) */
int osc_row_align(const float* Ustar, const float* psi, int64_t N, int32_t D, float* align, void* stream);
/* d2[p] = || V_i/(sd_i+1e-12) - V_j/(sd_j+1e-12) ||^2 for pairs[p] = (i,j)   (lattice.py:468-471) */
int osc_pair_d2(const float* V, const float* sqrt_deg, const int32_t* pairs, int64_t M, int32_t D,
                float* out, void* stream);
/* greedy MMR over cosine-normalised anchors Yn (graph.py:114-133, lambda_div = 0.5):
 * chosen[s] = argmax_i 0.5*score_i - 0.5*max_{j chosen} <Yn_i,Yn_j>, lowest index on ties */
int osc_mmr_workspace(int64_t N, size_t* h_bytes);
int osc_mmr_select(const float* Yn, const float* score, int64_t N, int32_t D, int32_t k, int32_t* chosen,
                   void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ K3: batched serving
 * One persistent kernel settles `batch` independent lattices of equal (N, D, k): each CTA
 * owns a 4-column slab of one lattice with the CG vectors in registers/shared memory and
 * iterates with no cross-CTA synchronisation; the lattice-wide stop test (solver.py:29-31)
 * is resolved afterwards (slabs that stopped early are re-run for the lattice's count).
 * do_settle: U_out = settle(U_in or Y).  do_ustar: Ustar_out = stationary solve.
 * do_deltaH: deltaH[b] = <U_out - Ustar, M (U_out - Ustar)> (needs both).
 * psi: [batch][D]; gates: [batch][N] or NULL (= ones).  stats: [batch][4] floats
 * {settle_iters, settle_res, ustar_iters, ustar_res}.  unresolved (optional): [batch] int32
 * flags -- bit 0: the per-slab residuals were not monotone around the stop and the lattice
 * must be settled with osc_pcg_solve instead (never seen on real inputs; the host mirror
 * handles it); bit 1 (informational): some slab of the lattice was re-run in pass 2. */
typedef struct osc_batched_args {
  const float* Y;
  const float* U_in; /* NULL -> Y */
  const float* psi;
  const float* gates; /* NULL -> ones */
  float* U_out;
  float* Ustar_out; /* may be NULL unless do_ustar output is wanted */
  float* stats;
  double* deltaH; /* [batch] */
  int32_t D;
  int32_t do_settle, do_ustar, do_deltaH;
  float dt;
  double tol_settle, tol_ustar;
  int32_t max_iters_settle, max_iters_ustar;
  int32_t* unresolved; /* [batch] or NULL */
} osc_batched_args_t;

int osc_batched_supported(int64_t N, int32_t D, int32_t k);
int osc_batched_workspace(int64_t batch, int64_t N, int32_t D, size_t* h_bytes);
int osc_batched_settle(const osc_graph_t* g, const osc_params_t* prm, const osc_batched_args_t* a,
                       void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSCILLINK_B200_H */
