// Device code of the steady Stokes assembly (b2_stokes.cu): the element loop of the reference's
// applications/003_NavierStokes/SteadyStokes/main.cpp:290-598 (AssembleMatrixResNS) for three velocity components of
// one Lagrange family and a pressure of another (Taylor-Hood pairs; the equal-order stabilisation, alpha != 0, is
// not implemented), scattered into the system rows [rank][variable][dof].  Free of host / runtime calls so that the
// same source compiles for the CPU thread emulator (tests/cpp/cuda_emu.hpp).
//
// Per Gauss point g of the VELOCITY element (geometry = the velocity family's nodes, main.cpp:448-450):
//     K[i][j]    += sum_d dphi2_i/dx_d dphi2_j/dx_d w        B[k][k] = IRe K  (the same block for every component)
//     G_k[i][j]  -= dphi2_i/dx_k phi1_j w                    B[k][p] = G_k,  B[p][k] = G_k^T
// and, the problem being linear, the residual is  F = -B sol  (F_u[k] = -IRe K U_k - G_k P,  F_p = -sum_k G_k^T U_k),
// which is what the reference accumulates term by term (:486-494, :521-528).  B[p][p] holds zeros (alpha = 0) that
// belong to the pattern; nothing is added there.
//
// One warp per element, table-driven like assemble_general_kernel (any element type: nv <= 27 velocity nodes,
// np <= 8 pressure nodes, ng <= 64 points).  Phase A: lanes = Gauss points (J, det, J^-1).  Phase B: per Gauss point
// the nv physical gradients go to shared memory and lane l accumulates the entries l, l + 32, ... of K and G_0..2 in
// the warp's shared accumulator (nv^2 + 3 nv np doubles; owned entries, no conflicts).  Phase C: residual from the
// element blocks, then fp64 atomicAdd scatter; the position of a column inside its CSR row is found by bisection
// (first correct path: a slot map as in the Poisson kernels is the obvious next step).
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kStokesWarps = 4;

#define B2_STOKES_WARP_DOUBLES(nv, np, ng) (3 * 32 + 3 * 32 + 10 * (ng) + 3 * 32 + 8 + (nv) * (nv) + 3 * (nv) * (np))   /* X, G, Geo, U, P, K, G_k */
__device__ __forceinline__ int stokes_warp_doubles(int nv, int np, int ng) { return B2_STOKES_WARP_DOUBLES(nv, np, ng); }
inline int stokes_warp_doubles_host(int nv, int np, int ng) { return B2_STOKES_WARP_DOUBLES(nv, np, ng); }

__device__ __forceinline__ int64_t stokes_find(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int32_t row, int32_t c) {
  int64_t lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;          // the pattern holds every element coupling: col[lo] == c
}

// tabv: dxi, deta, dzeta [ng][nv], w[ng] of the velocity element; tabp: phi [ng][np] of the pressure element;
// edof: [nel][4][27] system dofs (U, V, W, P), -1 padded
__global__ void __launch_bounds__(kStokesWarps * 32)
stokes_kernel(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
              const int32_t* __restrict__ edof, const double* __restrict__ tabv, const double* __restrict__ tabp,
              const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, double* Aval, const double* __restrict__ sol,
              double* rhs, double IRe) {
  B2_DYN_SHARED(double, smem);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const double* t_dx = tabv;
  const double* t_dy = t_dx + ng * nv;
  const double* t_dz = t_dy + ng * nv;
  const double* t_w = t_dz + ng * nv;
  double* sX = smem + wib * stokes_warp_doubles(nv, np, ng);      // [3][32]
  double* sG = sX + 96;                                           // [3][32] physical gradients at the current point
  double* sGeo = sG + 96;                                         // [10][ng]
  double* sU = sGeo + 10 * ng;                                    // [3][32]
  double* sP = sU + 96;                                           // [8]
  double* sK = sP + 8;                                            // [nv][nv]
  double* sGk = sK + nv * nv;                                     // [3][nv][np]
  const int nK = nv * nv, nG = nv * np, nacc = nK + 3 * nG;

  for (int64_t el = (int64_t)blockIdx.x * kStokesWarps + wib; el < nel; el += (int64_t)gridDim.x * kStokesWarps) {
    const int32_t* ed = edof + el * 108;
    if (lane < nv) {
      const int64_t nd = conn[el * 27 + lane];
      sX[lane] = xyz[nd];
      sX[32 + lane] = xyz[nnode + nd];
      sX[64 + lane] = xyz[2 * nnode + nd];
      for (int k = 0; k < 3; k++) sU[32 * k + lane] = sol ? sol[ed[27 * k + lane]] : 0.0;
    }
    if (lane < np) sP[lane] = sol ? sol[ed[81 + lane]] : 0.0;
    for (int e = lane; e < nacc; e += 32) sK[e] = 0.0;             // sK and sGk are contiguous
    __syncwarp();

    // ---- A. geometry at the Gauss points owned by this lane (Jacobian_type, ElemType.hpp:1438-1537)
    for (int g = lane; g < ng; g += 32) {
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
    __syncwarp();

    // ---- B. element blocks K and G_0..2
    for (int g = 0; g < ng; g++) {
      if (lane < nv) {
        const double a = t_dx[g * nv + lane], b = t_dy[g * nv + lane], c = t_dz[g * nv + lane];
        sG[lane] = fma(c, sGeo[2 * ng + g], fma(b, sGeo[1 * ng + g], a * sGeo[0 * ng + g]));
        sG[32 + lane] = fma(c, sGeo[5 * ng + g], fma(b, sGeo[4 * ng + g], a * sGeo[3 * ng + g]));
        sG[64 + lane] = fma(c, sGeo[8 * ng + g], fma(b, sGeo[7 * ng + g], a * sGeo[6 * ng + g]));
      }
      __syncwarp();
      const double wg = sGeo[9 * ng + g];
      for (int e = lane; e < nacc; e += 32) {
        if (e < nK) {
          const int i = e / nv, j = e - i * nv;
          const double d = fma(sG[64 + i], sG[64 + j], fma(sG[32 + i], sG[32 + j], sG[i] * sG[j]));
          sK[e] = fma(d, wg, sK[e]);
        } else {
          const int q = e - nK, k = q / nG, r = q - k * nG, i = r / np, j = r - i * np;
          sK[e] = fma(-sG[32 * k + i] * tabp[g * np + j], wg, sK[e]);
        }
      }
      __syncwarp();
    }

    // ---- C. residual F = -B sol, then the scatter
    if (rhs) {
      for (int q = lane; q < 3 * nv + np; q += 32) {
        double f = 0.0;
        if (q < 3 * nv) {
          const int k = q / nv, i = q - k * nv;
          double s = 0.0;
          for (int j = 0; j < nv; j++) s = fma(sK[i * nv + j], sU[32 * k + j], s);
          f = -IRe * s;
          for (int j = 0; j < np; j++) f = fma(-sGk[(k * nv + i) * np + j], sP[j], f);
          atomicAdd(&rhs[ed[27 * k + i]], f);
        } else {
          const int i = q - 3 * nv;
          for (int k = 0; k < 3; k++)
            for (int j = 0; j < nv; j++) f = fma(-sGk[(k * nv + j) * np + i], sU[32 * k + j], f);
          atomicAdd(&rhs[ed[81 + i]], f);
        }
      }
    }
    for (int e = lane; e < nacc; e += 32) {
      if (e < nK) {
        const int i = e / nv, j = e - i * nv;
        const double v = IRe * sK[e];
        for (int k = 0; k < 3; k++) {
          const int32_t r = ed[27 * k + i];
          atomicAdd(&Aval[stokes_find(rowptr, col, r, ed[27 * k + j])], v);
        }
      } else {
        const int q = e - nK, k = q / nG, rr = q - k * nG, i = rr / np, j = rr - i * np;
        const int32_t ru = ed[27 * k + i], rp = ed[81 + j];
        atomicAdd(&Aval[stokes_find(rowptr, col, ru, rp)], sK[e]);
        atomicAdd(&Aval[stokes_find(rowptr, col, rp, ru)], sK[e]);
      }
    }
    __syncwarp();
  }
}
