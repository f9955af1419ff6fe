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
// np <= 8 pressure nodes, ng <= 64 points).  Phase A: lanes = Gauss points (J, det, J^-1).  Phase B: lane j computes the
// physical gradient of ITS velocity function at the point, keeps it in registers and publishes it in a double-buffered
// shared tile; it owns COLUMN j of K (row i needs the three broadcast loads of function i's gradient and three fused
// multiply-adds) and ROW j of G_0..2 (one broadcast load of psi per entry): nv + 3 np accumulators in registers, no index
// decoding, no shared-memory accumulators, every row independent of the others.  The row counts are template parameters
// (one instantiation per Taylor-Hood pair of the reference's element families, a guarded one for anything else), so the
// loops unroll without branches.  Phase C: blocks to shared memory, residual from the element blocks, fp64 atomicAdd
// scatter through a precomputed element -> CSR slot map (stokes_slot_kernel, once per plan): lanes run along a row of
// the element block, so the atomics of one instruction fall into the same CSR row.
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

constexpr int kStokesWarps = 4;

/* per warp: X[3][32], G[2][3][32], Psi[2][8], Geo[10][ng], U[3][32], P[8], row starts [4][32] (int64), dofs [4][32] (int32),
 * blocks K | G_k [nv nv + 3 nv np] */
#define B2_STOKES_WARP_DOUBLES(nv, np, ng) (96 + 192 + 16 + 10 * (ng) + 96 + 8 + 128 + 64 + (nv) * (nv) + 3 * (nv) * (np))
__device__ __forceinline__ int stokes_warp_doubles(int nv, int np, int ng) { return B2_STOKES_WARP_DOUBLES(nv, np, ng); }
inline int stokes_warp_doubles_host(int nv, int np, int ng) { return B2_STOKES_WARP_DOUBLES(nv, np, ng); }
/* slots per element: K on the three velocity diagonals [3][nv nv], then G_k in the velocity rows and its transpose in
 * the pressure rows [2][3 nv np] */
__host__ __device__ __forceinline__ int stokes_slots_per_element(int nv, int np) { return 3 * nv * nv + 6 * nv * np; }

// position of every element coupling inside its CSR row, once per plan.  t < 3 nK: (k, i, j) of K on diagonal k;
// then (k, i, j) of G_k in row (U_k, i), column (P, j); then the transposed entry.  err: 1 = coupling not in the pattern,
// 2 = position does not fit 16 bits
__global__ void stokes_slot_kernel(int64_t nel, int nv, int np, const int32_t* __restrict__ edof, const int64_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ col, unsigned short* __restrict__ slot, int* err) {
  const int nK = nv * nv, nG = nv * np, S = 3 * nK + 6 * nG;
  const int64_t total = nel * (int64_t)S;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t el = t / S;
    const int r = (int)(t - el * S);
    const int32_t* ed = edof + el * 108;
    int32_t row, c;
    if (r < 3 * nK) {
      const int k = r / nK, q = r - k * nK, i = q / nv, j = q - i * nv;
      row = ed[27 * k + i];
      c = ed[27 * k + j];
    } else {
      const int tr = (r - 3 * nK) / (3 * nG), q = (r - 3 * nK) - tr * 3 * nG, k = q / nG, rr = q - k * nG, i = rr / np, j = rr - i * np;
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

// tabv: dxi, deta, dzeta [ng][nv], w[ng] of the velocity element; tabp: phi [ng][np] of the pressure element;
// edof: [nel][4][27] system dofs (U, V, W, P), -1 padded; slot: [nel][stokes_slots_per_element].
// NV, NP: compile-time nv, np (GUARD = false), or their upper bounds with the rows beyond nv / np skipped (GUARD = true)
template <int NV, int NP, bool GUARD>
__global__ void __launch_bounds__(kStokesWarps * 32)
stokes_kernel(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
              const int32_t* __restrict__ edof, const double* __restrict__ tabv, const double* __restrict__ tabp,
              const int64_t* __restrict__ rowptr, const unsigned short* __restrict__ slot, double* Aval, const double* __restrict__ sol,
              double* rhs, double IRe) {
  B2_DYN_SHARED(double, smem);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const double* t_dx = tabv;
  const double* t_dy = t_dx + ng * nv;
  const double* t_dz = t_dy + ng * nv;
  const double* t_w = t_dz + ng * nv;
  double* sX = smem + wib * stokes_warp_doubles(nv, np, ng);      // [3][32]
  double* sG = sX + 96;                                           // [2][3][32] physical gradients at the current point
  double* sPsi = sG + 192;                                        // [2][8]
  double* sGeo = sPsi + 16;                                       // [10][ng]
  double* sU = sGeo + 10 * ng;                                    // [3][32]
  double* sP = sU + 96;                                           // [8]
  int64_t* sRow = reinterpret_cast<int64_t*>(sP + 8);             // [4][32] first slot of the row of (variable, node)
  int32_t* sDof = reinterpret_cast<int32_t*>(sRow + 128);         // [4][32]
  double* sK = reinterpret_cast<double*>(sDof + 128);             // [nv][nv]
  double* sGk = sK + nv * nv;                                     // [3][nv][np]
  const int nK = nv * nv, nG = nv * np;
  const bool mine = lane < nv;                                    // this lane owns a velocity function

  for (int64_t el = (int64_t)blockIdx.x * kStokesWarps + wib; el < nel; el += (int64_t)gridDim.x * kStokesWarps) {
    const int32_t* ed = edof + el * 108;
    if (mine) {
      const int64_t nd = conn[el * 27 + lane];
      sX[lane] = xyz[nd];
      sX[32 + lane] = xyz[nnode + nd];
      sX[64 + lane] = xyz[2 * nnode + nd];
      for (int k = 0; k < 3; k++) {
        const int32_t d = ed[27 * k + lane];
        sDof[32 * k + lane] = d;
        sRow[32 * k + lane] = rowptr[d];
        sU[32 * k + lane] = sol ? sol[d] : 0.0;
      }
    }
    if (lane < np) {
      const int32_t d = ed[81 + lane];
      sDof[96 + lane] = d;
      sRow[96 + lane] = rowptr[d];
      sP[lane] = sol ? sol[d] : 0.0;
    }
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

    // ---- B. lane j: column j of K (aK[i] = K[i][j]) and row j of G_0..2 (aG[k][q] = G_k[j][q]) in registers
    double aK[NV], aG[3][NP];
#pragma unroll
    for (int i = 0; i < NV; i++) aK[i] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int q = 0; q < NP; q++) aG[k][q] = 0.0;
    for (int g = 0; g < ng; g++) {
      double* G = sG + (g & 1) * 96;
      double* Psi = sPsi + (g & 1) * 8;
      double gx = 0.0, gy = 0.0, gz = 0.0;
      if (mine) {
        const double a = t_dx[g * nv + lane], b = t_dy[g * nv + lane], c = t_dz[g * nv + lane];
        gx = fma(c, sGeo[2 * ng + g], fma(b, sGeo[1 * ng + g], a * sGeo[0 * ng + g]));
        gy = fma(c, sGeo[5 * ng + g], fma(b, sGeo[4 * ng + g], a * sGeo[3 * ng + g]));
        gz = fma(c, sGeo[8 * ng + g], fma(b, sGeo[7 * ng + g], a * sGeo[6 * ng + g]));
        G[lane] = gx;
        G[32 + lane] = gy;
        G[64 + lane] = gz;
      }
      if (lane < np) Psi[lane] = tabp[g * np + lane];
      __syncwarp();
      const double wg = sGeo[9 * ng + g];
      const double wx = gx * wg, wy = gy * wg, wz = gz * wg;      // this lane's gradient times the weight
#pragma unroll
      for (int i = 0; i < NV; i++)
        if (!GUARD || i < nv) aK[i] = fma(G[64 + i], wz, fma(G[32 + i], wy, fma(G[i], wx, aK[i])));
#pragma unroll
      for (int q = 0; q < NP; q++)
        if (!GUARD || q < np) {
          const double ps = Psi[q];
          aG[0][q] = fma(-wx, ps, aG[0][q]);
          aG[1][q] = fma(-wy, ps, aG[1][q]);
          aG[2][q] = fma(-wz, ps, aG[2][q]);
        }
    }

    // ---- C. blocks to shared memory (K is symmetric in exact arithmetic; stored as computed: K[i][j] by lane j),
    // residual F = -B sol, scatter through the slot map
    if (mine) {
#pragma unroll
      for (int i = 0; i < NV; i++)
        if (!GUARD || i < nv) sK[i * nv + lane] = aK[i];
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int q = 0; q < NP; q++)
          if (!GUARD || q < np) sGk[(k * nv + lane) * np + q] = aG[k][q];
    }
    __syncwarp();
    if (rhs) {
      for (int q = lane; q < 3 * nv + np; q += 32) {
        double f = 0.0;
        if (q < 3 * nv) {
          const int k = q / nv, i = q - k * nv;
          double s = 0.0;
          for (int j = 0; j < nv; j++) s = fma(sK[i * nv + j], sU[32 * k + j], s);
          f = -IRe * s;
          for (int j = 0; j < np; j++) f = fma(-sGk[(k * nv + i) * np + j], sP[j], f);
          atomicAdd(&rhs[sDof[32 * k + i]], f);
        } else {
          const int i = q - 3 * nv;
          for (int k = 0; k < 3; k++)
            for (int j = 0; j < nv; j++) f = fma(-sGk[(k * nv + j) * np + i], sU[32 * k + j], f);
          atomicAdd(&rhs[sDof[96 + i]], f);
        }
      }
    }
    const unsigned short* sl = slot + (size_t)el * stokes_slots_per_element(nv, np);
    const int dv = GUARD ? nv : NV, dp = GUARD ? np : NP;         // compile-time divisors where the pair is known
    for (int e = lane; e < nK; e += 32) {
      const int i = e / dv;
      const double val = IRe * sK[e];
      for (int k = 0; k < 3; k++) atomicAdd(&Aval[sRow[32 * k + i] + (int64_t)sl[k * nK + e]], val);
    }
    for (int r = lane; r < 3 * nG; r += 32) {
      const int ki = r / dp, j = r - ki * dp, k = ki / dv, i = ki - k * dv;      // r = (k nv + i) np + j
      const double val = sGk[r];
      atomicAdd(&Aval[sRow[32 * k + i] + (int64_t)sl[3 * nK + r]], val);              // row (U_k, i), column (P, j)
      atomicAdd(&Aval[sRow[96 + j] + (int64_t)sl[3 * nK + 3 * nG + r]], val);         // row (P, j), column (U_k, i)
    }
    __syncwarp();      // the next element overwrites X, U, P, dofs, Geo, the blocks
  }
}

// the Taylor-Hood pairs of the reference's element families (hexahedra 20 / 27 + 8, tetrahedra 10 / 15 + 4, wedges
// 15 / 21 + 6) get their own instantiation; anything else within nv <= 27, np <= 8 runs the guarded one
#define B2_STOKES_PAIRS(X) X(27, 8) X(20, 8) X(10, 4) X(15, 4) X(21, 6) X(15, 6)
typedef void (*stokes_kernel_t)(int64_t, int64_t, int, int, int, const double*, const int32_t*, const int32_t*, const double*, const double*,
                                const int64_t*, const unsigned short*, double*, const double*, double*, double);
inline stokes_kernel_t stokes_kernel_for(int nv, int np) {
#define B2_STOKES_CASE(V, P) if (nv == V && np == P) return stokes_kernel<V, P, false>;
  B2_STOKES_PAIRS(B2_STOKES_CASE)
#undef B2_STOKES_CASE
  return stokes_kernel<27, 8, true>;
}
