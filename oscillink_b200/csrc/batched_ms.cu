// K3 fast path: settle + stationary solve of one lattice slab as ONE multi-shift CG.
//
// First settle after the constructor (U = Y, lattice.py:55), uniform gates (b = 1), no chain:
//   settle     (I + dt M) U+ = Y + dt RHS , x0 = Y      (lattice.py:170-207)
//   stationary        M  U* =        RHS , x0 = Y      (lattice.py:245-265)
// have initial residuals dt*r0 and r0 with r0 = RHS - M Y, and I + dt M = dt (M + sigma I) with
// sigma = 1/dt: the two systems are SHIFTS of one another with the same right-hand side, and with
// uniform gates the Jacobi preconditioner of either (solver.py:18, lattice.py:187-192,257-259) is a
// scalar, i.e. PCG produces the iterates of plain CG.  The Krylov spaces coincide, so ONE sequence
// of SpMMs drives both recurrences (multi-shift CG, Frommer/Jegerlehner): the residual of the shifted
// system is collinear with the base one, r^s_k = zeta_k r_k, and per column
//   zeta_{k+1} = zeta_k zeta_{k-1} a_{k-1} / (a_k b_{k-1} (zeta_{k-1} - zeta_k) + zeta_{k-1} a_{k-1} (1 + sigma a_k))
//   a^s_k = a_k zeta_{k+1}/zeta_k ,  b^s_k = b_k (zeta_{k+1}/zeta_k)^2
//   x^s += a^s_k p^s ,  p^s = zeta_{k+1} r_{k+1} + b^s_k p^s
// (a_k, b_k: the alpha/beta of solver.py:23,34 for the stationary system in plain-CG scaling).
// Iterates, iteration counts and residuals are those of the two separate solves up to fp32 rounding
// (tools/dev_multishift.py: U, U* within 3e-7, deltaH within 5e-8 of the oracle's two PCG runs).
// Gather passes per lattice: 1 (over Y) + T_u instead of 1 + T_s + T_u  (6 instead of 10 at config #2).
//
// Second change against batched.cu: the vector that is gathered is the RESIDUAL, not the search
// direction.  p_{k+1} = im r_{k+1} + beta_k p_k is linear, so sum_j W_ij p_j = im * gather(r_{k+1}) +
// beta_k * G_k with G_k the previous gather sums (one more register vector).  r can be published right
// after the r update, the next gather pass starts behind the SAME barrier that delivers r.r, and p never
// visits shared memory: two barriers and one STS per row and iteration instead of three and LDS + STS.
//
// State per row (registers): X_u, X_s, P, P_s, R, G -- six float4.  Two slabs in flight do not fit the
// register file, so the kernel runs ONE CTA per SM with T threads x 2 rows.
#include <cstdlib>

#include "batched_common.cuh"

namespace osc {

template <int TPT, int KQ, int T>
__global__ void __launch_bounds__(T, 1) batched_ms_kernel(BatchedK P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int Np = T * TPT;
  constexpr int nw = T >> 5;
  // both gather sources are STATIC shared memory: a gather is LDS.128 [u16 offset + constant]
  __shared__ __align__(16) float4 r_static[Np];  // r_k (the gathered vector)
  __shared__ __align__(16) float4 y_static[Np];  // the slab's 4 columns of Y, prefetched with cp.async
  const int N = P.N;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float4* w_s = reinterpret_cast<float4*>(smem_raw);              // [KQ][Np]
  float4* redA = w_s + (size_t)Np * KQ;                           // p.Ap / deltaH partials
  float4* redB = redA + RED_F4;                                   // r.r partials
  ushort4* nbr_s = reinterpret_cast<ushort4*>(redB + RED_F4);     // [KQ][Np]
  for (int e = tid; e < Np * KQ; e += T) {
    w_s[e] = f4_zero();
    nbr_s[e] = make_ushort4(0, 0, 0, 0);
  }
  for (int e = tid; e < Np; e += T) {
    r_static[e] = f4_zero();
    y_static[e] = f4_zero();
  }
  bool act[TPT];
#pragma unroll
  for (int m = 0; m < TPT; ++m) act[m] = (tid + T * m) < N;

  const bool list_mode = P.fix_list != nullptr;
  const int64_t n_work = list_mode ? (int64_t)(*P.fix_count) : P.n_work;
  float4* scr = P.scratch + (size_t)blockIdx.x * 2 * N;  // r_{T_u} when the base system stops first (rare)
  bool y_ahead = false;
  auto y_fetch = [&](int64_t fb, int fs) {
    const float* src = P.Y + fb * (int64_t)N * P.D + fs * SC;
#pragma unroll
    for (int m = 0; m < TPT; ++m)
      if (act[m]) cp_async16(y_static + tid + T * m, src + (int64_t)(tid + T * m) * P.D);
    cp_async_commit();
  };

  // stationary operator M (b = 1) and the scalar Jacobi factor of solver.py:18
  const float diag = (P.lamG + P.lamC) + P.lamQ;
  const float noffc = -P.lamC;
  const float im = __fdiv_rn(1.0f, __fadd_rn(P.lamG, P.lamQ) + 1e-12f);
  const float sigma = __fdiv_rn(1.0f, P.dt);
  const V4 IM = v4_bc(im);

  for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
    int64_t b;
    int s0, s1, Fs = 0, Fu = 0;
    if (list_mode) {
      const int4 e = P.fix_list[wk];
      b = e.x; s0 = e.y; s1 = s0 + 1; Fs = e.z; Fu = e.w;
    } else {
      b = wk / P.cpl;
      s0 = (int)(wk - b * P.cpl) * P.CH;
      s1 = min(s0 + P.CH, P.G);
    }
    __syncthreads();  // readers of the previous graph image are done
    {
      const uint4* src_w = reinterpret_cast<const uint4*>(P.pk_w) + b * N * KQ;
      uint4* dst_w = reinterpret_cast<uint4*>(w_s);
      const uint2* src_n = reinterpret_cast<const uint2*>(P.pk_nbr) + b * N * KQ;
      uint2* dst_n = reinterpret_cast<uint2*>(nbr_s);
#pragma unroll
      for (int c = 0; c < KQ; ++c) {
        for (int e = tid; e < N; e += T) {
          dst_w[c * Np + e] = __ldg(src_w + c * N + e);
          dst_n[c * Np + e] = __ldg(src_n + c * N + e);
        }
      }
    }
    float* Uo = P.U_out + b * (int64_t)N * P.D;
    float* So = P.Ustar_out ? P.Ustar_out + b * (int64_t)N * P.D : nullptr;

    for (int s = s0; s < s1; ++s) {
      const int col = s * SC;
      const float4 psi4 = *reinterpret_cast<const float4*>(P.psi + b * P.D + col);
      if (!y_ahead) y_fetch(b, s);
      cp_async_wait_all();
      y_ahead = false;
      __syncthreads();  // Y slab + graph image visible; the previous slab's readers of r_static / red are done
      auto y_next = [&]() {  // issued once this slab has read Y for the last time
        if (s + 1 < s1) {
          y_fetch(b, s + 1);
          y_ahead = true;
        } else if (!list_mode && wk + gridDim.x < n_work) {
          const int64_t wn = wk + gridDim.x;
          const int64_t bn = wn / P.cpl;
          y_fetch(bn, (int)(wn - bn * P.cpl) * P.CH);
          y_ahead = true;
        }
      };

      // ---- r0 = RHS - M Y ; p0 = im r0 ; p^s_0 = r0 ; x_u = x_s = Y
      V4 Xu[TPT], Xs[TPT], Pv[TPT], Ps[TPT], R[TPT], G[TPT];
      V4 part = v4_zero();
#pragma unroll
      for (int m = 0; m < TPT; ++m) {
        const int row = tid + T * m;
        const float4 y = y_static[row];
        const V4 g0 = gather_row<KQ>(y_static, nbr_s, w_s, row, Np, KQ);
        // lattice.py:184,256 (same rounding order); pad rows carry zeros
        const float4 rhs = act[m] ? make_float4(__fadd_rn(__fmul_rn(P.lamG, y.x), __fmul_rn(P.lamQ, psi4.x)),
                                                __fadd_rn(__fmul_rn(P.lamG, y.y), __fmul_rn(P.lamQ, psi4.y)),
                                                __fadd_rn(__fmul_rn(P.lamG, y.z), __fmul_rn(P.lamQ, psi4.z)),
                                                __fadd_rn(__fmul_rn(P.lamG, y.w), __fmul_rn(P.lamQ, psi4.w)))
                                  : f4_zero();
        const V4 yv = to_v4(y);
        Xu[m] = yv;
        Xs[m] = yv;
        R[m] = v4_sub(to_v4(rhs), combine_row(yv, g0, diag, noffc));
        Pv[m] = v4_mul(IM, R[m]);
        Ps[m] = R[m];
        G[m] = v4_zero();
        part = v4_fma(R[m], R[m], part);
        sts_v4(r_static + row, R[m]);
      }
      warp_reduce4(to_f4(part), redB + warp, lane);
      __syncthreads();  // r0 visible, r0.r0 partials visible
      float rr = block_total_c(redB, nw, lane);  // column lane & 3
      float rz = rr * im;
      // per-column shift state (lane's column)
      float zeta = 1.f, zeta_p = 1.f, a_prev = 1.f, b_prev = 0.f, beta = 0.f;
      bool fs = false, fu = false, ru_in_scr = false;
      int Ts = 0, Tu = 0;
      float rrs_rec = 0.f, rru_rec = 0.f;
      int k = 0;
      while (true) {
        ++k;
        // ---- gather r_k ; G = im * gather + beta G ; p.Ap
        const V4 BETA = to_v4(bcast4(beta));
        part = v4_zero();
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          const int row = tid + T * m;
          const V4 g = gather_row<KQ>(r_static, nbr_s, w_s, row, Np, KQ);
          G[m] = v4_fma(IM, g, v4_mul(BETA, G[m]));
          part = v4_fma(Pv[m], combine_row(Pv[m], G[m], diag, noffc), part);
        }
        warp_reduce4(to_f4(part), redA + warp, lane);
        __syncthreads();  // B1: every gather of r_k is done
        const float pap = block_total_c(redA, nw, lane);
        const float alpha = __fdiv_rn(rz, pap + 1e-18f);  // solver.py:23
        const float a = alpha * im;                        // plain-CG step length
        const float den = a * b_prev * (zeta_p - zeta) + zeta_p * a_prev * (1.0f + sigma * a);
        const float zn = den != 0.f ? __fdiv_rn(zeta * zeta_p * a_prev, den) : zeta;
        const float ratio = zeta != 0.f ? __fdiv_rn(zn, zeta) : 0.f;
        const float4 al4 = bcast4(alpha);
        const V4 AL = to_v4(al4);
        const V4 NAL = to_v4(make_float4(-al4.x, -al4.y, -al4.z, -al4.w));
        const V4 AS = to_v4(bcast4(fs ? 0.f : a * ratio));
        const V4 AU = fu ? v4_zero() : AL;
        // ---- x_u += alpha p ; x_s += a^s p^s ; r -= alpha A p ; publish r ; r.r
        part = v4_zero();
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          const V4 ap = combine_row(Pv[m], G[m], diag, noffc);
          Xu[m] = v4_fma(Pv[m], AU, Xu[m]);
          Xs[m] = v4_fma(Ps[m], AS, Xs[m]);
          R[m] = v4_fma(ap, NAL, R[m]);
          part = v4_fma(R[m], R[m], part);
          sts_v4(r_static + tid + T * m, R[m]);
        }
        warp_reduce4(to_f4(part), redB + warp, lane);
        __syncthreads();  // B2: r_{k+1} visible, r.r partials visible
        const float rr_new = block_total_c(redB, nw, lane);
        const float rzn = rr_new * im;
        beta = __fdiv_rn(rzn, rz + 1e-18f);  // solver.py:34
        const float bs = beta * ratio * ratio;
        // stop tests (solver.py:29-31): stationary on ||r||, settle on ||dt zeta r||, max over the slab's columns
        const float zs = P.dt * zn;
        float mu = rr_new, ms = zs * zs * rr_new;
        mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, 1));
        mu = fmaxf(mu, __shfl_xor_sync(0xffffffffu, mu, 2));
        ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, 1));
        ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, 2));
        // identical in every thread (same summation order) -> uniform branches
        const bool stop_s = !fs && (Fs > 0 ? (k >= Fs)
                                           : ((double)__fsqrt_rn(ms) <= P.tol_settle || k >= P.max_iters_settle));
        const bool stop_u = !fu && (Fu > 0 ? (k >= Fu)
                                           : ((double)__fsqrt_rn(mu) <= P.tol_ustar || k >= P.max_iters_ustar));
        if (stop_s) {
          // U+ out; M U+ = RHS + (Y - U+)/dt - zeta r  (the settle system's own residual is dt zeta r):
          // the dead p^s registers take t1 = (Y - U+)/dt - zeta r for the deltaH identity
          const V4 ZN = to_v4(bcast4(zn));
          const V4 SG = v4_bc(sigma);
#pragma unroll
          for (int m = 0; m < TPT; ++m) {
            const int row = tid + T * m;
            if (act[m]) *reinterpret_cast<float4*>(Uo + (int64_t)row * P.D + col) = to_f4(Xs[m]);
            const V4 dy = v4_sub(lds_v4(y_static + row), Xs[m]);
            Ps[m] = v4_sub(v4_mul(SG, dy), v4_mul(ZN, R[m]));
          }
          fs = true;
          Ts = k;
          rrs_rec = ms;
          y_next();
        }
        if (stop_u) {
          if (So != nullptr) {
#pragma unroll
            for (int m = 0; m < TPT; ++m)
              if (act[m]) *reinterpret_cast<float4*>(So + (int64_t)(tid + T * m) * P.D + col) = to_f4(Xu[m]);
          }
          if (!fs) {  // the settle system needs more iterations: park r_{T_u} (thread-private rows)
#pragma unroll
            for (int m = 0; m < TPT; ++m)
              if (act[m]) scr[tid + T * m] = to_f4(R[m]);
            ru_in_scr = true;
          }
          fu = true;
          Tu = k;
          rru_rec = mu;
        }
        if (fs && fu) break;
        // ---- p = im r + beta p ; p^s = zeta r + b^s p^s
        const V4 BT = to_v4(bcast4(beta));
        const V4 ZN = to_v4(bcast4(zn));
        const V4 BS = to_v4(bcast4(bs));
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          Pv[m] = v4_fma(Pv[m], BT, v4_mul(IM, R[m]));
          if (!fs) Ps[m] = v4_fma(Ps[m], BS, v4_mul(ZN, R[m]));
        }
        zeta_p = zeta;
        zeta = zn;
        a_prev = a;
        b_prev = beta;
        rz = rzn;
      }
      if (tid == 0) {
        P.rec[(b * 2 + 0) * P.G + s] = make_int2(Ts, __float_as_int(rrs_rec));
        P.rec[(b * 2 + 1) * P.G + s] = make_int2(Tu, __float_as_int(rru_rec));
      }
      if (P.do_dh) {
        // deltaH = <U - U*, M(U - U*)> with M U* = RHS - r_{T_u}: M(U - U*) = t1 + r_{T_u}  (receipts.py:21-25)
        float4 dpart = f4_zero();
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          float4 ru = to_f4(R[m]);
          if (ru_in_scr) ru = act[m] ? scr[tid + T * m] : f4_zero();
          const float4 xs = to_f4(Xs[m]), xu = to_f4(Xu[m]), t1 = to_f4(Ps[m]);
          const float4 d = make_float4(__fsub_rn(xs.x, xu.x), __fsub_rn(xs.y, xu.y), __fsub_rn(xs.z, xu.z),
                                       __fsub_rn(xs.w, xu.w));
          dpart = f4_add(dpart, f4_mul(d, f4_add(t1, ru)));
        }
        warp_reduce4(dpart, redA + warp, lane);
        __syncthreads();
        if (warp == 0) {
          const float4 tot = block_total(redA, nw, lane);
          if (lane == 0) P.dh_part[b * P.G + s] = (double)((tot.x + tot.y) + (tot.z + tot.w));
        }
      }
    }
  }
}

// ================================================================= host side
// T threads x 2 rows; the smallest block that covers N keeps the pad rows (gathered like real ones) few
static int ms_threads(int64_t N) {
  static const int ts[] = {128, 256, 384, 512, 608, 640};
  for (int t : ts)
    if (2 * (int64_t)t >= N) return t;
  return 0;
}

template <int T>
static BatchedFn ms_pick_kq(int kq) {
  switch (kq) {
    case 1: return batched_ms_kernel<2, 1, T>;
    case 2: return batched_ms_kernel<2, 2, T>;
    case 3: return batched_ms_kernel<2, 3, T>;
    default: return batched_ms_kernel<2, 4, T>;
  }
}

// The multi-shift kernel that serves (N, kq), its block size and dynamic shared memory; nullptr if none.
BatchedFn batched_ms_pick(int64_t N, int kq, int* threads, size_t* smem_dyn) {
  const int t = ms_threads(N);
  if (t == 0 || kq < 1 || kq > 4) return nullptr;
  const size_t Np = (size_t)t * 2;
  *threads = t;
  *smem_dyn = Np * kq * (16 + 8) + 2 * RED_F4 * 16;
  switch (t) {
    case 128: return ms_pick_kq<128>(kq);
    case 256: return ms_pick_kq<256>(kq);
    case 384: return ms_pick_kq<384>(kq);
    case 512: return ms_pick_kq<512>(kq);
    case 608: return ms_pick_kq<608>(kq);
    default: return ms_pick_kq<640>(kq);
  }
}

}  // namespace osc
