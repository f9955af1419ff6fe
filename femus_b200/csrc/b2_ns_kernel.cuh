// Device code of the steady Navier-Stokes assembly (b2_stokes.cu, b2_ns_* entry points): the element loop of the
// reference's library routine src/08_equations/assemble/03_navier_stokes.hpp:305-413 -- residual of the Galerkin form
// and its exact Newton Jacobian, which the reference obtains by recording the loop with adept -- for three velocity
// components of one Lagrange family and a pressure of another, scattered into the system rows [rank][variable][dof].
// Free of host / runtime calls so that the same source compiles for the CPU thread emulator (tests/cpp/cuda_emu.hpp).
//
// Per Gauss point, with u, grad u, p evaluated from the current solution:
//     aResV[k][i] += ( nu grad phi_i . grad u_k + phi_i u . grad u_k - p dphi_i/dx_k ) w          RES = -aRes
//     aResP[i]    += - div u psi_i w
//     D[i][j]       += ( nu grad phi_i . grad phi_j + phi_i u . grad phi_j ) w      d aResV[k] / d u_k  (every k)
//     N[k][l][i][j] += phi_i phi_j du_k/dx_l w                                      d aResV[k] / d u_l  (Newton term)
//     G_k[i][j]     -= dphi_i/dx_k psi_j w                                          d aResV[k] / d p, transposed: d aResP / d u_k
//
// One CTA per element, thread t owning the (i, j) pairs t, t + T, t + 2T of the nv x nv blocks and keeping the TEN
// accumulators of a pair (D and the nine Newton blocks N_kl share phi_i phi_j w) in registers, plus up to three entries of
// G_0..2; T = ns_threads(nv, np) threads, so that three pairs per thread cover the block (256 for Q2-Q1, 64 for P2-P1).
// The solution at the Gauss points (u, grad u, p: 13 values per point) is evaluated ONCE per element in a pre-pass over
// (point, variable) tasks; in the point loop the physical gradients, phi and psi are staged in a double-buffered tile:
// one barrier per point, no shared-memory accumulators.  Scatter: fp64 atomicAdd through a precomputed element -> CSR
// slot map (ns_slot_kernel, once per plan), consecutive threads along a row of a block.
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kNsMaxThreads = 256;
constexpr int kNsPairs = 3;      // (i, j) pairs per thread
constexpr int kNsGacc = 3;       // entries of G_0..2 per thread
/* X[3][32], Geo[10][ng], G[2][3][32], phi[2][32], psi[2][8], U[3][32], P[8], Q[ng][16], row starts [4][32] (int64), dofs [4][32] (int32) */
#define B2_NS_CTA_DOUBLES(nv, np, ng) (96 + 10 * (ng) + 192 + 64 + 16 + 96 + 8 + 16 * (ng) + 128 + 64)
inline int ns_cta_doubles_host(int nv, int np, int ng) { (void)nv; (void)np; return B2_NS_CTA_DOUBLES(nv, np, ng); }
/* threads per element: three pairs per thread cover nv^2, three entries per thread cover 3 nv np, and the first
 * 3 nv + np threads own the residual entries */
__host__ __device__ __forceinline__ int ns_threads(int nv, int np) {
  int t = (nv * nv + kNsPairs - 1) / kNsPairs;
  if (t < nv * np) t = nv * np;
  if (t < 3 * nv + np) t = 3 * nv + np;
  t = (t + 31) & ~31;
  return t < 64 ? 64 : t;
}
/* slots per element: the nine velocity blocks [9][nv nv] ((k, l) row-major), then G_k in the velocity rows and its
 * transpose in the pressure rows [2][3 nv np] */
__host__ __device__ __forceinline__ int ns_slots_per_element(int nv, int np) { return 9 * nv * nv + 6 * nv * np; }

// position of every element coupling inside its CSR row, once per plan (err as in stokes_slot_kernel)
__global__ void ns_slot_kernel(int64_t nel, int nv, int np, const int32_t* __restrict__ edof, const int64_t* __restrict__ rowptr,
                               const int32_t* __restrict__ col, unsigned short* __restrict__ slot, int* err) {
  const int nK = nv * nv, nG = nv * np, S = 9 * nK + 6 * nG;
  const int64_t total = nel * (int64_t)S;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t el = t / S;
    const int r = (int)(t - el * S);
    const int32_t* ed = edof + el * 108;
    int32_t row, c;
    if (r < 9 * nK) {
      const int kl = r / nK, q = r - kl * nK, i = q / nv, j = q - i * nv, k = kl / 3, l = kl - 3 * k;
      row = ed[27 * k + i];
      c = ed[27 * l + j];
    } else {
      const int tr = (r - 9 * nK) / (3 * nG), q = (r - 9 * nK) - tr * 3 * nG, k = q / nG, rr = q - k * nG, i = rr / np, j = rr - i * np;
      const int32_t du = ed[27 * k + i], dp = ed[81 + j];
      row = tr ? dp : du;
      c = tr ? du : dp;
    }
    const int64_t s0 = rowptr[row];
    int64_t lo = s0, hi = rowptr[row + 1];
    const int64_t end = hi;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (col[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (lo >= end || col[lo] != c) *err = 1;
    else if (lo - s0 > 65535) *err = 2;
    slot[t] = (unsigned short)(lo - s0);
  }
}

// tabv: phi, dxi, deta, dzeta [ng][nv], w[ng] of the velocity element; tabp: psi [ng][np]; edof [nel][4][27];
// slot [nel][ns_slots_per_element]; launched with ns_threads(nv, np) threads
// RP, RG: pairs / G entries per thread known at compile time (the accumulation loops unroll without guards: threads
// beyond the block compute on entry 0 and scatter nothing), or 0, 0 = the guarded instantiation for any nv, np
template <int RP, int RG>
__global__ void __launch_bounds__(kNsMaxThreads, 2)
ns_kernel(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
          const int32_t* __restrict__ edof, const double* __restrict__ tabv, const double* __restrict__ tabp,
          const int64_t* __restrict__ rowptr, const unsigned short* __restrict__ slot, double* Aval, const double* __restrict__ sol, double* rhs,
          double nu) {
  B2_DYN_SHARED(double, smem);
  const int tid = threadIdx.x, T = blockDim.x;
  const double* t_phi = tabv;
  const double* t_dx = t_phi + ng * nv;
  const double* t_dy = t_dx + ng * nv;
  const double* t_dz = t_dy + ng * nv;
  const double* t_w = t_dz + ng * nv;
  double* sX = smem;                 // [3][32]
  double* sGeo = sX + 96;            // [10][ng]
  double* sG = sGeo + 10 * ng;       // [2][3][32] physical gradients at the current point
  double* sPhi = sG + 192;           // [2][32]
  double* sPsi = sPhi + 64;          // [2][8]
  double* sU = sPsi + 16;            // [3][32]
  double* sP = sU + 96;              // [8]
  double* sQ = sP + 8;               // [ng][16]: u[3], grad u [3][3] (du_k/dx_l at 3 + 3 k + l), p
  int64_t* sRow = reinterpret_cast<int64_t*>(sQ + 16 * ng);     // [4][32]
  int32_t* sDof = reinterpret_cast<int32_t*>(sRow + 128);       // [4][32]
  const int nK = nv * nv, nG = nv * np;

  constexpr int NPAIR = RP ? RP : kNsPairs, NGACC = RP ? RG : kNsGacc;
  // the pairs / G entries this thread owns, decoded once
  int pi[NPAIR], pj[NPAIR], gk[NGACC], gi[NGACC], gj[NGACC];
#pragma unroll
  for (int r = 0; r < NPAIR; r++) {
    const int p = tid + r * T;
    pi[r] = p < nK ? p / nv : 0;
    pj[r] = p < nK ? p - pi[r] * nv : 0;
  }
#pragma unroll
  for (int r = 0; r < NGACC; r++) {
    const int e = tid + r * T;
    const int k = e < 3 * nG ? e / nG : 0, rr = e < 3 * nG ? e - k * nG : 0;
    gk[r] = k;
    gi[r] = rr / np;
    gj[r] = rr - gi[r] * np;
  }
  // the residual entry this thread owns: velocity (k, i) for tid < 3 nv, pressure i for the next np threads
  const int rk = tid < 3 * nv ? tid / nv : 3, ri = tid < 3 * nv ? tid - rk * nv : tid - 3 * nv;

  for (int64_t el = blockIdx.x; el < nel; el += gridDim.x) {
    const int32_t* ed = edof + el * 108;
    if (tid < nv) {
      const int64_t nd = conn[el * 27 + tid];
      sX[tid] = xyz[nd];
      sX[32 + tid] = xyz[nnode + nd];
      sX[64 + tid] = xyz[2 * nnode + nd];
      for (int k = 0; k < 3; k++) {
        const int32_t d = ed[27 * k + tid];
        sDof[32 * k + tid] = d;
        sRow[32 * k + tid] = rowptr[d];
        sU[32 * k + tid] = sol ? sol[d] : 0.0;
      }
    } else if (tid >= 32 && tid < 32 + np) {
      const int32_t d = ed[81 + tid - 32];
      sDof[96 + tid - 32] = d;
      sRow[96 + tid - 32] = rowptr[d];
      sP[tid - 32] = sol ? sol[d] : 0.0;
    }
    __syncthreads();

    // ---- A. geometry at the Gauss points (Jacobian_type, ElemType.hpp:1438-1537)
    for (int g = tid; g < ng; g += T) {
      double J00 = 0, J01 = 0, J02 = 0, J10 = 0, J11 = 0, J12 = 0, J20 = 0, J21 = 0, J22 = 0;
      for (int n = 0; n < nv; n++) {
        const double x0 = sX[n], x1 = sX[32 + n], x2 = sX[64 + n];
        const double a = t_dx[g * nv + n], b = t_dy[g * nv + n], c = t_dz[g * nv + n];
        J00 = fma(a, x0, J00); J01 = fma(a, x1, J01); J02 = fma(a, x2, J02);
        J10 = fma(b, x0, J10); J11 = fma(b, x1, J11); J12 = fma(b, x2, J12);
        J20 = fma(c, x0, J20); J21 = fma(c, x1, J21); J22 = fma(c, x2, J22);
      }
      const double det = J00 * (J11 * J22 - J12 * J21) + J01 * (J12 * J20 - J10 * J22) + J02 * (J10 * J21 - J11 * J20);
      const double id = 1.0 / det;
      sGeo[0 * ng + g] = (-J12 * J21 + J11 * J22) * id;
      sGeo[1 * ng + g] = (J02 * J21 - J01 * J22) * id;
      sGeo[2 * ng + g] = (-J02 * J11 + J01 * J12) * id;
      sGeo[3 * ng + g] = (J12 * J20 - J10 * J22) * id;
      sGeo[4 * ng + g] = (-J02 * J20 + J00 * J22) * id;
      sGeo[5 * ng + g] = (J02 * J10 - J00 * J12) * id;
      sGeo[6 * ng + g] = (-J11 * J20 + J10 * J21) * id;
      sGeo[7 * ng + g] = (J01 * J20 - J00 * J21) * id;
      sGeo[8 * ng + g] = (-J01 * J10 + J00 * J11) * id;
      sGeo[9 * ng + g] = det * t_w[g];
    }
    __syncthreads();

    // ---- A2. the solution at the Gauss points: task (g, v), v < 3: u_v and its physical gradient; v = 3: p
    for (int t = tid; t < 4 * ng; t += T) {
      const int g = t >> 2, v = t & 3;
      if (v < 3) {
        double u = 0.0, a = 0.0, b = 0.0, c = 0.0;       // value and reference gradient of component v
        for (int i = 0; i < nv; i++) {
          const double ui = sU[32 * v + i];
          u = fma(ui, t_phi[g * nv + i], u);
          a = fma(ui, t_dx[g * nv + i], a);
          b = fma(ui, t_dy[g * nv + i], b);
          c = fma(ui, t_dz[g * nv + i], c);
        }
        sQ[16 * g + v] = u;
        for (int l = 0; l < 3; l++)
          sQ[16 * g + 3 + 3 * v + l] = fma(c, sGeo[(3 * l + 2) * ng + g], fma(b, sGeo[(3 * l + 1) * ng + g], a * sGeo[(3 * l) * ng + g]));
      } else {
        double s = 0.0;
        for (int j = 0; j < np; j++) s = fma(sP[j], tabp[g * np + j], s);
        sQ[16 * g + 12] = s;
      }
    }
    // stage point 0
    if (tid < nv) {
      const double a = t_dx[tid], b = t_dy[tid], c = t_dz[tid];
      sG[tid] = fma(c, sGeo[2 * ng], fma(b, sGeo[1 * ng], a * sGeo[0]));
      sG[32 + tid] = fma(c, sGeo[5 * ng], fma(b, sGeo[4 * ng], a * sGeo[3 * ng]));
      sG[64 + tid] = fma(c, sGeo[8 * ng], fma(b, sGeo[7 * ng], a * sGeo[6 * ng]));
      sPhi[tid] = t_phi[tid];
    } else if (tid >= 32 && tid < 32 + np) {
      sPsi[tid - 32] = tabp[tid - 32];
    }
    __syncthreads();

    // ---- B. Gauss point loop: accumulators in registers
    double aD[NPAIR], aN[NPAIR][9], aG[NGACC], aRes = 0.0;
#pragma unroll
    for (int r = 0; r < NPAIR; r++) {
      aD[r] = 0.0;
#pragma unroll
      for (int kl = 0; kl < 9; kl++) aN[r][kl] = 0.0;
    }
#pragma unroll
    for (int r = 0; r < NGACC; r++) aG[r] = 0.0;
    for (int g = 0; g < ng; g++) {
      const double* G = sG + (g & 1) * 96;
      const double* Phi = sPhi + (g & 1) * 32;
      const double* Psi = sPsi + (g & 1) * 8;
      if (g + 1 < ng) {              // stage the next point in the other buffer
        const int h = g + 1;
        double* Gn = sG + (h & 1) * 96;
        if (tid < nv) {
          const double a = t_dx[h * nv + tid], b = t_dy[h * nv + tid], c = t_dz[h * nv + tid];
          Gn[tid] = fma(c, sGeo[2 * ng + h], fma(b, sGeo[1 * ng + h], a * sGeo[0 * ng + h]));
          Gn[32 + tid] = fma(c, sGeo[5 * ng + h], fma(b, sGeo[4 * ng + h], a * sGeo[3 * ng + h]));
          Gn[64 + tid] = fma(c, sGeo[8 * ng + h], fma(b, sGeo[7 * ng + h], a * sGeo[6 * ng + h]));
          sPhi[(h & 1) * 32 + tid] = t_phi[h * nv + tid];
        } else if (tid >= 32 && tid < 32 + np) {
          sPsi[(h & 1) * 8 + tid - 32] = tabp[h * np + tid - 32];
        }
      }
      const double* Q = sQ + 16 * g;
      const double wg = sGeo[9 * ng + g];
      const double u0 = Q[0], u1 = Q[1], u2 = Q[2];
#pragma unroll
      for (int r = 0; r < NPAIR; r++) {
        if (RP || r * T < nK) {      // guarded instantiation only (uniform over the CTA)
          const int i = pi[r], j = pj[r];
          const double lap = fma(G[64 + i], G[64 + j], fma(G[32 + i], G[32 + j], G[i] * G[j]));
          const double adv = fma(u2, G[64 + j], fma(u1, G[32 + j], u0 * G[j]));
          aD[r] = fma(fma(Phi[i], adv, nu * lap), wg, aD[r]);
          const double pp = Phi[i] * Phi[j] * wg;
#pragma unroll
          for (int kl = 0; kl < 9; kl++) aN[r][kl] = fma(pp, Q[3 + kl], aN[r][kl]);
        }
      }
#pragma unroll
      for (int r = 0; r < NGACC; r++)
        if (RP || r * T < 3 * nG) aG[r] = fma(-G[32 * gk[r] + gi[r]] * Psi[gj[r]], wg, aG[r]);
      if (tid < 3 * nv) {
        const double* gu = Q + 3 + 3 * rk;
        const double conv = fma(u2, gu[2], fma(u1, gu[1], u0 * gu[0]));
        const double visc = fma(G[64 + ri], gu[2], fma(G[32 + ri], gu[1], G[ri] * gu[0]));
        aRes = fma(fma(Phi[ri], conv, fma(nu, visc, -Q[12] * G[32 * rk + ri])), wg, aRes);
      } else if (tid < 3 * nv + np) {
        aRes = fma(-(Q[3] + Q[7] + Q[11]) * Psi[ri], wg, aRes);
      }
      __syncthreads();
    }

    // ---- C. scatter: RES = -aRes, KK += Jacobian
    if (rhs && tid < 3 * nv + np) atomicAdd(&rhs[sDof[32 * rk + ri]], -aRes);
    const unsigned short* sl = slot + (size_t)el * ns_slots_per_element(nv, np);
#pragma unroll
    for (int r = 0; r < NPAIR; r++) {
      const int p = tid + r * T;
      if (p < nK) {
#pragma unroll
        for (int kl = 0; kl < 9; kl++) {
          const int k = kl / 3, l = kl - 3 * k;
          const double v = aN[r][kl] + (k == l ? aD[r] : 0.0);
          atomicAdd(&Aval[sRow[32 * k + pi[r]] + (int64_t)sl[kl * nK + p]], v);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NGACC; r++) {
      const int e = tid + r * T;
      if (e < 3 * nG) {
        atomicAdd(&Aval[sRow[32 * gk[r] + gi[r]] + (int64_t)sl[9 * nK + e]], aG[r]);
        atomicAdd(&Aval[sRow[96 + gj[r]] + (int64_t)sl[9 * nK + 3 * nG + e]], aG[r]);
      }
    }
    __syncthreads();      // the next element overwrites X, U, P, dofs, Geo, Q
  }
}

// per-thread counts of the Taylor-Hood pairs of the reference's element families with ns_threads() threads: (3, 3) for
// 27 + 8 / 20 + 8 / 21 + 6 / 15 + 6, (3, 2) for 15 + 4, (2, 2) for 10 + 4; anything else runs the guarded instantiation
typedef void (*ns_kernel_t)(int64_t, int64_t, int, int, int, const double*, const int32_t*, const int32_t*, const double*, const double*,
                            const int64_t*, const unsigned short*, double*, const double*, double*, double);
inline ns_kernel_t ns_kernel_for(int nv, int np) {
  const int T = ns_threads(nv, np), rp = (nv * nv + T - 1) / T, rg = (3 * nv * np + T - 1) / T;
  if (rp == 3 && rg == 3) return ns_kernel<3, 3>;
  if (rp == 3 && rg == 2) return ns_kernel<3, 2>;
  if (rp == 2 && rg == 2) return ns_kernel<2, 2>;
  return ns_kernel<0, 0>;
}
