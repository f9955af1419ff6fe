// Device code of the Neumann boundary integrals (b2_assemble.cu), free of host / runtime calls so that the same
// source compiles for the CPU thread emulator of tests/cpp/cuda_emu.hpp (tests/test_kernel_emulation.py).
#pragma once

// Neumann boundary integrals (applications/001_Poisson/main.cpp:495-548 with elem_type_2D::JacobianSur,
// ElemType.hpp:1330-1379): one warp per boundary face, lanes = the Gauss points of the face rule (16 on a
// quadrilateral, 13 on a triangle) for the surface Jacobian, then lanes = face dofs for
// F_i += sum_g phi_i(g) value weight_g.  The face element is the tables: nvf dofs, ngf points.
__global__ void neumann_kernel(int64_t nfaces, const int32_t* __restrict__ felem, const int32_t* __restrict__ flocal,
                               const double* __restrict__ fvalue, int nvf, int ngf, int nve, const double* __restrict__ ftab,
                               const int32_t* __restrict__ fnodes, int64_t nnode, const double* __restrict__ xyz,
                               const int32_t* __restrict__ conn, const int32_t* __restrict__ dof, double* __restrict__ rhs) {
  constexpr int NG2 = 16;                            // most points any face rule has
  __shared__ double sW[8][NG2];
  __shared__ double sX[8][3][9];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const double* phi = ftab;
  const double* dxi = phi + ngf * nvf;
  const double* deta = dxi + ngf * nvf;
  const double* w = deta + ngf * nvf;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < nfaces; k += nwarps) {
    const int64_t e = felem[k];
    const int f = flocal[k];
    int loc = 0;
    if (lane < nvf) {
      loc = fnodes[f * 9 + lane];                    // element-local node of face dof `lane`
      const int64_t nd = conn[e * 27 + loc];
      sX[wib][0][lane] = xyz[nd];
      sX[wib][1][lane] = xyz[nnode + nd];
      sX[wib][2][lane] = xyz[2 * nnode + nd];
    }
    __syncwarp();
    if (lane < ngf) {
      double J00 = 0, J10 = 0, J20 = 0, J01 = 0, J11 = 0, J21 = 0;
      for (int i = 0; i < nvf; i++) {
        const double a = dxi[lane * nvf + i], b = deta[lane * nvf + i];
        J00 = fma(a, sX[wib][0][i], J00); J10 = fma(a, sX[wib][1][i], J10); J20 = fma(a, sX[wib][2][i], J20);
        J01 = fma(b, sX[wib][0][i], J01); J11 = fma(b, sX[wib][1][i], J11); J21 = fma(b, sX[wib][2][i], J21);
      }
      const double nx = J10 * J21 - J11 * J20, ny = J01 * J20 - J21 * J00, nz = J00 * J11 - J01 * J10;
      const double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
      const double n0 = nx * inv, n1 = ny * inv, n2 = nz * inv;
      // the reference takes the determinant of [t1 t2 n] as the area element
      const double det = J00 * (J11 * n2 - n1 * J21) + J01 * (n1 * J20 - J10 * n2) + n0 * (J10 * J21 - J11 * J20);
      sW[wib][lane] = det * w[lane];
    }
    __syncwarp();
    if (lane < nvf) {
      double s = 0.0;
      for (int g = 0; g < ngf; g++) s = fma(phi[g * nvf + lane] * fvalue[k], sW[wib][g], s);
      atomicAdd(&rhs[dof[e * nve + loc]], s);
    }
    __syncwarp();
  }
}


// Boundary pressure term of the Navier-Stokes residual (src/08_equations/assemble/03_navier_stokes.hpp:196-300):
//     aResV[k][i] += phi_i tau n_k weight_g      over the listed boundary faces, tau = the prescribed boundary pressure
// (a constant per face here; the reference evaluates a callback at the Gauss point), n = the unit normal of
// JacobianSur at the Gauss point, and RES = -aRes.  Same work distribution as neumann_kernel; edof [nel][4][27] are
// the system dofs of U, V, W, P.
__global__ void pressure_face_kernel(int64_t nfaces, const int32_t* __restrict__ felem, const int32_t* __restrict__ flocal,
                                     const double* __restrict__ fvalue, int nvf, int ngf, const double* __restrict__ ftab,
                                     const int32_t* __restrict__ fnodes, int64_t nnode, const double* __restrict__ xyz,
                                     const int32_t* __restrict__ conn, const int32_t* __restrict__ edof, double* __restrict__ rhs) {
  constexpr int NG2 = 16;
  __shared__ double sWn[8][3][NG2];
  __shared__ double sX[8][3][9];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const double* phi = ftab;
  const double* dxi = phi + ngf * nvf;
  const double* deta = dxi + ngf * nvf;
  const double* w = deta + ngf * nvf;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp; k < nfaces; k += nwarps) {
    const int64_t e = felem[k];
    const int f = flocal[k];
    int loc = 0;
    if (lane < nvf) {
      loc = fnodes[f * 9 + lane];
      const int64_t nd = conn[e * 27 + loc];
      sX[wib][0][lane] = xyz[nd];
      sX[wib][1][lane] = xyz[nnode + nd];
      sX[wib][2][lane] = xyz[2 * nnode + nd];
    }
    __syncwarp();
    if (lane < ngf) {
      double J00 = 0, J10 = 0, J20 = 0, J01 = 0, J11 = 0, J21 = 0;
      for (int i = 0; i < nvf; i++) {
        const double a = dxi[lane * nvf + i], b = deta[lane * nvf + i];
        J00 = fma(a, sX[wib][0][i], J00); J10 = fma(a, sX[wib][1][i], J10); J20 = fma(a, sX[wib][2][i], J20);
        J01 = fma(b, sX[wib][0][i], J01); J11 = fma(b, sX[wib][1][i], J11); J21 = fma(b, sX[wib][2][i], J21);
      }
      const double nx = J10 * J21 - J11 * J20, ny = J01 * J20 - J21 * J00, nz = J00 * J11 - J01 * J10;
      const double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
      const double n0 = nx * inv, n1 = ny * inv, n2 = nz * inv;
      const double det = J00 * (J11 * n2 - n1 * J21) + J01 * (n1 * J20 - J10 * n2) + n0 * (J10 * J21 - J11 * J20);
      const double wg = det * w[lane];
      sWn[wib][0][lane] = wg * n0;
      sWn[wib][1][lane] = wg * n1;
      sWn[wib][2][lane] = wg * n2;
    }
    __syncwarp();
    if (lane < nvf) {
      for (int d = 0; d < 3; d++) {
        double s = 0.0;
        for (int g = 0; g < ngf; g++) s = fma(phi[g * nvf + lane] * fvalue[k], sWn[wib][d][g], s);
        atomicAdd(&rhs[edof[e * 108 + 27 * d + loc]], -s);
      }
    }
    __syncwarp();
  }
}
