// K4: per-node receipt terms and null points, one warp per lattice row.
//
//   receipts.py:40-59  coh_drop_i = sum_{j in nbr(i)} 0.5 lamC A_ij (||Yd_i-Yd_j||^2 - ||Ud_i-Ud_j||^2)
//                      anchor_i = lamG ||U*_i - Y_i||^2 ; query_i = lamQ b_i ||U*_i - psi||^2
//                      (Yd, Ud are rows divided by sqrt_deg + 1e-12 -- degree, not L2, normalised)
//   receipts.py:70-82  R_ij = lamC A_ij ||Ud_i-Ud_j||^2 ; z-score of the largest residual of the
//                      row against mean / population std over ALL N columns (zeros included)
//
// The reference evaluates an N x N x D broadcast (2.2 GB at N=1200); here every edge is visited
// once: HBM traffic is (deg+1) rows of U* and Y per lattice row.
#include "common.cuh"

namespace osc {

__global__ void __launch_bounds__(256)
receipt_full_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ A,
                    const int32_t* __restrict__ deg, const float* __restrict__ sqrt_deg, int64_t N,
                    int k, int D, float lamG, float lamC, float lamQ, const float* __restrict__ Y,
                    const float* __restrict__ Us, const float* __restrict__ psi,
                    const float* __restrict__ gates, float z_th, float* __restrict__ coh,
                    float* __restrict__ anchor, float* __restrict__ query,
                    int32_t* __restrict__ null_j, float* __restrict__ null_z,
                    float* __restrict__ null_R, float* __restrict__ row_mu,
                    float* __restrict__ row_sigma) {
  const int lane = threadIdx.x & 31;
  const int64_t b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int64_t row = b * N + i;
  const float* yi = Y + row * D;
  const float* ui = Us + row * D;
  const float* q = psi + b * D;
  const float den_i = sqrt_deg[row] + 1e-12f;

  float sa = 0.f, sq = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float u = ui[d];
    const float da = u - yi[d];
    const float dq = u - q[d];
    sa = fmaf(da, da, sa);
    sq = fmaf(dq, dq, sq);
  }
  sa = warp_sum(sa);
  sq = warp_sum(sq);

  const int n = deg[row];
  float cacc = 0.f;           // fp32 accumulator updated through a double add (receipts.py:59)
  double rsum = 0.0;
  float rbest = -1.f;
  int jbest = -1;
  float rvals_local = 0.f;    // lane t keeps R_t for t < 32; larger k handled by recompute below
  for (int t = 0; t < n; ++t) {
    const int j = nbr[row * k + t];
    const float a = A[row * k + t];
    const int64_t rj = b * N + j;
    const float den_j = sqrt_deg[rj] + 1e-12f;
    const float* yj = Y + rj * D;
    const float* uj = Us + rj * D;
    float d2y = 0.f, d2u = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float ey = __fdiv_rn(yi[d], den_i) - __fdiv_rn(yj[d], den_j);
      const float eu = __fdiv_rn(ui[d], den_i) - __fdiv_rn(uj[d], den_j);
      d2y = fmaf(ey, ey, d2y);
      d2u = fmaf(eu, eu, d2u);
    }
    d2y = warp_sum(d2y);
    d2u = warp_sum(d2u);
    if (a > 0.f) {
      const double term = 0.5 * (double)lamC * (double)a * ((double)d2y - (double)d2u);
      cacc = (float)((double)cacc + term);
    }
    const float R = __fmul_rn(__fmul_rn(lamC, a), d2u);
    rsum += (double)R;
    if (R > rbest) {  // strict: lowest column wins ties (columns ascend with t)
      rbest = R;
      jbest = j;
    }
    if ((t & 31) == lane) rvals_local = (t < 32) ? R : rvals_local;
  }
  // mean / population std over all N columns (numpy two-pass form)
  const float mu = (float)(rsum / (double)N);
  double dev = 0.0;
  if (n <= 32) {
    double mine = 0.0;
    if (lane < n) {
      const double e = (double)rvals_local - (double)mu;
      mine = e * e;
    }
    dev = warp_sum(mine);
  } else {
    // rare wide rows: recompute the residuals (same arithmetic) for the second pass
    for (int t = 0; t < n; ++t) {
      const int j = nbr[row * k + t];
      const float a = A[row * k + t];
      const int64_t rj = b * N + j;
      const float den_j = sqrt_deg[rj] + 1e-12f;
      const float* uj = Us + rj * D;
      float d2u = 0.f;
      for (int d = lane; d < D; d += 32) {
        const float eu = __fdiv_rn(ui[d], den_i) - __fdiv_rn(uj[d], den_j);
        d2u = fmaf(eu, eu, d2u);
      }
      d2u = warp_sum(d2u);
      const double e = (double)__fmul_rn(__fmul_rn(lamC, a), d2u) - (double)mu;
      dev += e * e;
    }
  }
  dev += (double)(N - n) * (double)mu * (double)mu;
  const float sigma = (float)sqrt(dev / (double)N) + 1e-12f;
  if (lane == 0) {
    const float bq = gates ? gates[row] : 1.0f;
    coh[row] = cacc;
    anchor[row] = __fmul_rn(lamG, sa);
    query[row] = __fmul_rn(__fmul_rn(lamQ, bq), sq);
    int32_t oj = -1;
    float oz = 0.f, oR = 0.f;
    if (n > 0 && jbest >= 0) {
      const float z = __fdiv_rn(rbest - mu, sigma);
      if (rbest > 0.f && z > z_th) {
        oj = jbest;
        oz = z;
        oR = rbest;
      }
    }
    null_j[row] = oj;
    null_z[row] = oz;
    null_R[row] = oR;
    if (row_mu != nullptr) row_mu[row] = mu;
    if (row_sigma != nullptr) row_sigma[row] = sigma;
  }
}

int launch_receipt_full(const osc_graph_t* g, const osc_params_t* prm, const float* Y,
                        const float* Us, const float* psi, const float* gates, int D, float z_th,
                        float* coh, float* anchor, float* query, int32_t* null_j, float* null_z,
                        float* null_R, float* row_mu, float* row_sigma, cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && Y != nullptr && Us != nullptr && psi != nullptr,
              "receipt_full: NULL argument");
  if (g->N == 0 || g->batch == 0) return OSC_OK;
  const int warps = 8;
  dim3 grid((unsigned)((g->N + warps - 1) / warps), (unsigned)g->batch);
  receipt_full_kernel<<<grid, warps * 32, 0, st>>>(g->nbr, g->A, g->deg, g->sqrt_deg, g->N, g->k, D,
                                                   prm->lamG, prm->lamC, prm->lamQ, Y, Us, psi, gates,
                                                   z_th, coh, anchor, query, null_j, null_z, null_R, row_mu,
                                                   row_sigma);
  OSC_LAUNCH_CHECK("receipt_full_kernel");
  return OSC_OK;
}

}  // namespace osc
