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
// One CTA (128 threads) per element: the 10 nv^2 + 3 nv np accumulators (7 938 doubles = 62 KB for Q2-Q1) live in
// shared memory, thread t owns the entries t, t + 128, ...; per Gauss point the physical gradients, phi, psi and the
// 13 solution values (u, grad u, p) are staged once.  Scatter: fp64 atomicAdd, column positions by bisection in the
// CSR row (first correct path).
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kNsThreads = 128;
#define B2_NS_FIXED_DOUBLES(ng) (96 + 10 * (ng) + 96 + 32 + 8 + 96 + 8 + 16 + 104)   /* X, Geo, G, phi, psi, U, P, Q, Res */
#define B2_NS_CTA_DOUBLES(nv, np, ng) (B2_NS_FIXED_DOUBLES(ng) + 10 * (nv) * (nv) + 3 * (nv) * (np))
inline int ns_cta_doubles_host(int nv, int np, int ng) { return B2_NS_CTA_DOUBLES(nv, np, ng); }

__device__ __forceinline__ int64_t ns_find(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int32_t row, int32_t c) {
  int64_t lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;          // the pattern holds every element coupling: col[lo] == c
}

// tabv: phi, dxi, deta, dzeta [ng][nv], w[ng] of the velocity element; tabp: psi [ng][np]; edof [nel][4][27]
__global__ void __launch_bounds__(kNsThreads)
ns_kernel(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
          const int32_t* __restrict__ edof, const double* __restrict__ tabv, const double* __restrict__ tabp,
          const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, double* Aval, const double* __restrict__ sol, double* rhs,
          double nu) {
  B2_DYN_SHARED(double, smem);
  const int tid = threadIdx.x;
  const double* t_phi = tabv;
  const double* t_dx = t_phi + ng * nv;
  const double* t_dy = t_dx + ng * nv;
  const double* t_dz = t_dy + ng * nv;
  const double* t_w = t_dz + ng * nv;
  double* sX = smem;                 // [3][32]
  double* sGeo = sX + 96;            // [10][ng]
  double* sG = sGeo + 10 * ng;       // [3][32] physical gradients at the current point
  double* sPhi = sG + 96;            // [32]
  double* sPsi = sPhi + 32;          // [8]
  double* sU = sPsi + 8;             // [3][32]
  double* sP = sU + 96;              // [8]
  double* sQ = sP + 8;               // u[3], grad u [3][3], p
  double* sRes = sQ + 16;            // aResV [3][32], aResP [8]
  double* sD = sRes + 104;           // [nv][nv]
  const int nK = nv * nv, nG = nv * np, nacc = 10 * nK + 3 * nG;
  double* sN = sD + nK;              // [9][nv][nv]
  double* sGk = sN + 9 * nK;         // [3][nv][np]

  for (int64_t el = blockIdx.x; el < nel; el += gridDim.x) {
    const int32_t* ed = edof + el * 108;
    if (tid < nv) {
      const int64_t nd = conn[el * 27 + tid];
      sX[tid] = xyz[nd];
      sX[32 + tid] = xyz[nnode + nd];
      sX[64 + tid] = xyz[2 * nnode + nd];
      for (int k = 0; k < 3; k++) sU[32 * k + tid] = sol ? sol[ed[27 * k + tid]] : 0.0;
    }
    if (tid < np) sP[tid] = sol ? sol[ed[81 + tid]] : 0.0;
    for (int e = tid; e < nacc; e += kNsThreads) sD[e] = 0.0;          // sD, sN, sGk are contiguous
    if (tid < 104) sRes[tid] = 0.0;
    __syncthreads();

    // ---- A. geometry at the Gauss points (Jacobian_type, ElemType.hpp:1438-1537)
    for (int g = tid; g < ng; g += kNsThreads) {
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

    // ---- B. Gauss point loop
    for (int g = 0; g < ng; g++) {
      if (tid < nv) {
        const double a = t_dx[g * nv + tid], b = t_dy[g * nv + tid], c = t_dz[g * nv + tid];
        sG[tid] = fma(c, sGeo[2 * ng + g], fma(b, sGeo[1 * ng + g], a * sGeo[0 * ng + g]));
        sG[32 + tid] = fma(c, sGeo[5 * ng + g], fma(b, sGeo[4 * ng + g], a * sGeo[3 * ng + g]));
        sG[64 + tid] = fma(c, sGeo[8 * ng + g], fma(b, sGeo[7 * ng + g], a * sGeo[6 * ng + g]));
        sPhi[tid] = t_phi[g * nv + tid];
      } else if (tid >= 32 && tid < 32 + np) {
        sPsi[tid - 32] = tabp[g * np + tid - 32];
      }
      __syncthreads();
      if (tid < 13) {            // u_k, du_k/dx_l, p at the point
        double s = 0.0;
        if (tid < 3) {
          for (int i = 0; i < nv; i++) s = fma(sU[32 * tid + i], sPhi[i], s);
        } else if (tid < 12) {
          const int k = (tid - 3) / 3, l = (tid - 3) - 3 * k;
          for (int i = 0; i < nv; i++) s = fma(sU[32 * k + i], sG[32 * l + i], s);
        } else {
          for (int j = 0; j < np; j++) s = fma(sP[j], sPsi[j], s);
        }
        sQ[tid] = s;
      }
      __syncthreads();
      const double wg = sGeo[9 * ng + g];
      const double u0 = sQ[0], u1 = sQ[1], u2 = sQ[2];
      for (int e = tid; e < nacc; e += kNsThreads) {
        if (e < nK) {
          const int i = e / nv, j = e - i * nv;
          const double lap = fma(sG[64 + i], sG[64 + j], fma(sG[32 + i], sG[32 + j], sG[i] * sG[j]));
          const double adv = fma(u2, sG[64 + j], fma(u1, sG[32 + j], u0 * sG[j]));
          sD[e] = fma(fma(sPhi[i], adv, nu * lap), wg, sD[e]);
        } else if (e < 10 * nK) {
          const int q = e - nK, kl = q / nK, r = q - kl * nK, i = r / nv, j = r - i * nv;
          sD[e] = fma(sPhi[i] * sPhi[j] * sQ[3 + kl], wg, sD[e]);
        } else {
          const int q = e - 10 * nK, k = q / nG, r = q - k * nG, i = r / np, j = r - i * np;
          sD[e] = fma(-sG[32 * k + i] * sPsi[j], wg, sD[e]);
        }
      }
      if (tid < 3 * nv) {
        const int k = tid / nv, i = tid - k * nv;
        const double* gu = sQ + 3 + 3 * k;
        const double conv = fma(u2, gu[2], fma(u1, gu[1], u0 * gu[0]));
        const double visc = fma(sG[64 + i], gu[2], fma(sG[32 + i], gu[1], sG[i] * gu[0]));
        sRes[32 * k + i] = fma(fma(sPhi[i], conv, fma(nu, visc, -sQ[12] * sG[32 * k + i])), wg, sRes[32 * k + i]);
      } else if (tid < 3 * nv + np) {
        const int i = tid - 3 * nv;
        sRes[96 + i] = fma(-(sQ[3] + sQ[7] + sQ[11]) * sPsi[i], wg, sRes[96 + i]);
      }
      __syncthreads();
    }

    // ---- C. scatter: RES = -aRes, KK += Jacobian
    if (rhs) {
      if (tid < 3 * nv) {
        const int k = tid / nv, i = tid - k * nv;
        atomicAdd(&rhs[ed[27 * k + i]], -sRes[32 * k + i]);
      } else if (tid < 3 * nv + np) {
        atomicAdd(&rhs[ed[81 + tid - 3 * nv]], -sRes[96 + tid - 3 * nv]);
      }
    }
    for (int e = tid; e < 9 * nK; e += kNsThreads) {
      const int kl = e / nK, r = e - kl * nK, i = r / nv, j = r - i * nv, k = kl / 3, l = kl - 3 * k;
      const double v = sN[e] + (k == l ? sD[r] : 0.0);
      atomicAdd(&Aval[ns_find(rowptr, col, ed[27 * k + i], ed[27 * l + j])], v);
    }
    for (int e = tid; e < 3 * nG; e += kNsThreads) {
      const int k = e / nG, r = e - k * nG, i = r / np, j = r - i * np;
      const int32_t ru = ed[27 * k + i], rp = ed[81 + j];
      atomicAdd(&Aval[ns_find(rowptr, col, ru, rp)], sGk[e]);
      atomicAdd(&Aval[ns_find(rowptr, col, rp, ru)], sGk[e]);
    }
    __syncthreads();
  }
}
