// Host-side interface of the PCG phase kernels (pcg.cu), shared with the C ABI (cabi.cu) and the
// multi-GPU driver (dist.cu).
#pragma once
#include "common.cuh"

namespace osc {

// Where the gathered vector lives.
//   all != nullptr, local_ids == 0: one buffer of N rows indexed by GLOBAL row id (one GPU, column slabs,
//                                   or the all-gathered search direction of the rows partition);
//   all != nullptr, local_ids == 1: a rank-local block [own rows | pulled halo rows]: the graph's neighbour
//                                   ids (and the chain's column ids) index its rows directly and the own
//                                   row i sits at row i (rows partition, OSC_HALO_PULL);
//   peers != nullptr              : rank g's block of `shard` rows sits in peers[g], a buffer in GPU g's HBM
//                                   mapped into this process (CUDA IPC); remote rows are fetched by plain
//                                   loads over NVLink inside the SpMM (osc_pcg_*_p2p).
struct VecView {
  const float* all;
  const float* const* peers;
  int64_t shard;
  int local_ids;
};

// Device-resident loop control of a solve: the stop test of solver.py:29-31 is evaluated ON the device
// (pcg_decide); every phase kernel returns at once when `done` is set, so the host may enqueue an
// iteration ahead of the test without changing the result.
struct PcgCtl {
  int done;
  int iters;
  float res;
  int _pad;
};

int pcg_plan(osc_pcg_dims_t*, size_t*);
int pcg_max_ell_width(int D);
int pcg_setup(const osc_pcg_dims_t*, const osc_params_t*, int mode, float dt, int warm, float inertia,
              const float* Y, const float* U, const float* psi, const float* gates, float* X, float* Bv,
              cudaStream_t);
// Fused first residual: the right-hand side is formed inside the SpMM from Y, U and psi (rows of the local
// block), and the gathered vector is the start vector itself -- Y (y_is_x0) or U (u_is_x0).
struct InitSrc {
  const float* Y;
  const float* U;
  const float* psi;
  int y_is_x0, u_is_x0;
};
bool pcg_fused_init_ok(const osc_pcg_dims_t* d, const osc_graph_t* g);
int spmm_launch(bool res0, const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                const osc_params_t* prm, int mode, float dt, int jacobi, const float* gates, VecView vv,
                float* out, float* Pout, double* part, cudaStream_t st, const int* done = nullptr,
                const InitSrc* init = nullptr);
int pcg_residual0(const osc_pcg_dims_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*,
                  int mode, float dt, int jacobi, const float* gates, const float* Xall, float* RBv,
                  float* P, double* part_rz, cudaStream_t);
int pcg_spmm_dot(const osc_pcg_dims_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*,
                 int mode, float dt, const float* gates, const float* Pall, float* AP, double* part_pap,
                 cudaStream_t, const int* done = nullptr);
int pcg_residual0_p2p(const osc_pcg_dims_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*,
                      int mode, float dt, int jacobi, const float* gates, const float* const* peers,
                      int64_t shard, float* RBv, float* P, double* part_rz, cudaStream_t);
int pcg_spmm_dot_p2p(const osc_pcg_dims_t*, const osc_graph_t*, const osc_chain_t*, const osc_params_t*,
                     int mode, float dt, const float* gates, const float* const* peers, int64_t shard,
                     float* AP, double* part_pap, cudaStream_t);
int pcg_reduce(const double* part, int n_blocks, int D, float* out, float* d_max, double* out64,
               cudaStream_t, const int* done = nullptr);
int pcg_update(const osc_pcg_dims_t*, const osc_params_t*, int mode, float dt, int jacobi,
               const float* gates, const float* rz, const float* pap, const float* P, const float* AP,
               float* X, float* R, double* part_rr, double* part_rz, cudaStream_t,
               const int* done = nullptr);
int pcg_pupdate(const osc_pcg_dims_t*, const osc_params_t*, int mode, float dt, int jacobi,
                const float* gates, const float* rz_new, const float* rz_old, const float* R, float* P,
                cudaStream_t, const int* done = nullptr);
bool pcg_fuse_x();
// x += alpha p fused with p = z + beta p (after the iteration's verdict; see pcg_pupdate_x_kernel)
int pcg_pupdate_x(const osc_pcg_dims_t*, const osc_params_t*, int mode, float dt, int jacobi, const float* gates,
                  const float* rz_new, const float* rz_old, const float* pap, const float* R, float* P, float* X,
                  const PcgCtl* ctl, int it, int x_only, cudaStream_t st, const float* Xsrc = nullptr);
// stop test on the device: res = max_c sqrt(rr_c) from the column sums rr[D] (or, if rr == nullptr, the
// already reduced *d_res); records {iters = it, res}; sets done when res <= tol or it >= max_iters
int pcg_decide(PcgCtl* ctl, const float* rr, const float* d_res, int D, double tol, int it, int max_iters,
               cudaStream_t);
int pcg_solve(const osc_graph_t*, const osc_chain_t*, const osc_params_t*, int mode, float dt, int warm,
              float inertia, int jacobi, double tol, int max_iters, const float* Y, const float* U,
              const float* psi, const float* gates, int D, float* X, int* h_iters, float* h_res,
              void* workspace, size_t ws_bytes, cudaStream_t);
int pcg_solve_system(const osc_graph_t*, const osc_chain_t*, const osc_params_t*, int mode, float dt,
                     int jacobi, double tol, int max_iters, const float* gates, int D, float* X, float* B,
                     int* h_iters, float* h_res, void* workspace, size_t ws_bytes, cudaStream_t);
int delta_h(const osc_graph_t*, const osc_chain_t*, const osc_params_t*, const float* U,
            const float* Ustar, const float* gates, int D, double* h_out, void* workspace, size_t ws_bytes,
            cudaStream_t);
int launch_diff(const float* a, const float* b, float* out, int64_t n, cudaStream_t);
int launch_sum_doubles(const double* v, int D, double* total, cudaStream_t);

// Host-side poll of a device PcgCtl with a lag: record() enqueues an async copy of the control block into
// a pinned slot + an event; done(it) waits for THAT iteration's copy only.  Slots and events are cached per
// host thread and device (no per-solve allocation).
struct CtlPoll {
  PcgCtl* h = nullptr;      // two pinned slots
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init();
  int record(int it, const PcgCtl* d_ctl, cudaStream_t st);
  int wait(int it, PcgCtl* out);
};
CtlPoll* ctl_poll();  // nullptr on allocation failure (error text set)

}  // namespace osc
