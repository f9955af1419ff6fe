// Element assembly of the 3-D Poisson stiffness matrix / residual on HEX27 meshes, one warp per
// element.  Replaces the element loop of applications/001_Poisson/main.cpp:350-602 together with
// elem_type_3D::Jacobian_type (src/02_reference_geom_elements/03_fe_evaluations_at_quadrature/
// ElemType.hpp:1438-1537), MatSetValuesBlocked (PetscMatrix.cpp:699-729) and VecSetValues
// (PetscVector.cpp:132-141).
//
// Per element (warp):
//   A. lanes = Gauss points (2 each for the 64-point rule): J = sum_n dphi[g][n] x[n] with the
//      shape-derivative tables staged in shared memory once per CTA, then det, J^-1 and
//      weight = det * w_g, kept in shared memory (10 doubles per point).
//   B. for every Gauss point: lanes < nve form grad phi_n = J^-1 dphi_n (3 doubles each, shared
//      memory, double buffered), then every lane updates its TI x TJ register tile of
//      B_ij += (grad phi_i . grad phi_j) weight.
//   C. residual F_i = fsrc * sum_g phi_i weight - (B u)_i (row sums reduced with shuffles),
//      scatter: fp64 atomicAdd into the CSR through a precomputed element->slot map, and into rhs.
// The geometry map uses the unknown's own family and its first nve nodes, like the reference
// (ElemType.hpp:1462).  The kernel is FP64-FMA bound (~1.7e5 FMA per triquadratic element); the
// only HBM traffic is 27 node ids + 81 coordinates in and 729 atomics out.
#include "b2_common.cuh"

struct b2_mesh {
  b2_ctx* ctx;
  int64_t nnode, nel;
  double* xyz;     // [3][nnode]
  int32_t* conn;   // [nel][27]
};

struct b2_asm {
  b2_mesh* mesh;
  b2_csr* A;
  int nve, ngauss;
  int32_t* dof;    // [nel][nve]
  double* tab;     // phi, dxi, deta, dzeta [ng][nve] each, then w[ng]
  void* slot;      // [nel][TI*TJ][32] uint8 or uint16: position of (i,j) inside row dof_i
  int slot_bytes;  // 1 or 2
  size_t slot_count;
  double last_ms;
};

namespace {

constexpr int NG = 64;          // Gauss points per element ("seventh" hex rule)
constexpr int kWarps = 16;      // warps (= elements in flight) per CTA, one CTA per SM
constexpr int GP = 32;          // padded node stride of the gradient buffers

template <int NVE> struct Tile;
template <> struct Tile<27> { static constexpr int TI = 7, TJ = 4; };   // 4 x 7 lane grid, 28 lanes busy
template <> struct Tile<8> { static constexpr int TI = 2, TJ = 1; };    // 4 x 8 lane grid

template <int NVE>
struct SmemLayout {
  static constexpr int tab_doubles = 4 * NG * NVE + NG;
  // per warp: X[3][GP], U[GP], geo[10][NG], G[2][3][GP]
  static constexpr int warp_doubles = 3 * GP + GP + 10 * NG + 2 * 3 * GP;
  static constexpr size_t bytes = (size_t)(tab_doubles + kWarps * warp_doubles) * sizeof(double);
};

template <int NVE, typename SlotT>
__global__ void __launch_bounds__(kWarps * 32, 1)
assemble_poisson_kernel(int64_t nel, int64_t nnode, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
                        const int32_t* __restrict__ dof, const double* __restrict__ tab,
                        const SlotT* __restrict__ slot, const int64_t* __restrict__ rowptr, double* __restrict__ Aval,
                        const double* __restrict__ u, double* __restrict__ rhs, double nu, double fsrc) {
  constexpr int TI = Tile<NVE>::TI, TJ = Tile<NVE>::TJ;
  extern __shared__ double smem[];
  double* s_phi = smem;
  double* s_dx = s_phi + NG * NVE;
  double* s_dy = s_dx + NG * NVE;
  double* s_dz = s_dy + NG * NVE;
  double* s_w = s_dz + NG * NVE;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = s_w + NG + wib * SmemLayout<NVE>::warp_doubles;
  double* sX = wbase;                 // [3][GP]
  double* sU = sX + 3 * GP;           // [GP]
  double* sGeo = sU + GP;             // [10][NG]: J^-1 (9, row-major), weight
  double* sG = sGeo + 10 * NG;        // [2][3][GP]

  for (int t = threadIdx.x; t < SmemLayout<NVE>::tab_doubles; t += blockDim.x) smem[t] = tab[t];
  for (int t = lane; t < 2 * 3 * GP; t += 32) sG[t] = 0.0;     // padding nodes stay zero forever
  __syncthreads();

  const int rg = lane & 3, cg = lane >> 2;                      // row group / column group of the tile
  const int i0 = rg * TI, j0 = cg * TJ;

  for (int64_t e = (int64_t)blockIdx.x * kWarps + wib; e < nel; e += (int64_t)gridDim.x * kWarps) {
    // ---- gather: node ids (coalesced), coordinates, dofs, current solution
    int mydof = 0;
    if (lane < NVE) {
      const int64_t nd = conn[e * 27 + lane];
      sX[0 * GP + lane] = xyz[nd];
      sX[1 * GP + lane] = xyz[nnode + nd];
      sX[2 * GP + lane] = xyz[2 * nnode + nd];
      mydof = dof[e * NVE + lane];
      sU[lane] = u ? u[mydof] : 0.0;
    }
    __syncwarp();

    // ---- A. geometry at the Gauss points owned by this lane
#pragma unroll 1
    for (int g = lane; g < NG; g += 32) {
      double J00 = 0, J01 = 0, J02 = 0, J10 = 0, J11 = 0, J12 = 0, J20 = 0, J21 = 0, J22 = 0;
      const double* dx = s_dx + g * NVE;
      const double* dy = s_dy + g * NVE;
      const double* dz = s_dz + g * NVE;
#pragma unroll 9
      for (int n = 0; n < NVE; n++) {
        const double x0 = sX[n], x1 = sX[GP + n], x2 = sX[2 * GP + n];
        const double a = dx[n], b = dy[n], c = dz[n];
        J00 = fma(a, x0, J00); J01 = fma(a, x1, J01); J02 = fma(a, x2, J02);
        J10 = fma(b, x0, J10); J11 = fma(b, x1, J11); J12 = fma(b, x2, J12);
        J20 = fma(c, x0, J20); J21 = fma(c, x1, J21); J22 = fma(c, x2, J22);
      }
      const double det = J00 * (J11 * J22 - J12 * J21) + J01 * (J12 * J20 - J10 * J22) + J02 * (J10 * J21 - J11 * J20);
      const double id = 1.0 / det;
      sGeo[0 * NG + g] = (-J12 * J21 + J11 * J22) * id;
      sGeo[1 * NG + g] = (J02 * J21 - J01 * J22) * id;
      sGeo[2 * NG + g] = (-J02 * J11 + J01 * J12) * id;
      sGeo[3 * NG + g] = (J12 * J20 - J10 * J22) * id;
      sGeo[4 * NG + g] = (-J02 * J20 + J00 * J22) * id;
      sGeo[5 * NG + g] = (J02 * J10 - J00 * J12) * id;
      sGeo[6 * NG + g] = (-J11 * J20 + J10 * J21) * id;
      sGeo[7 * NG + g] = (J01 * J20 - J00 * J21) * id;
      sGeo[8 * NG + g] = (-J01 * J10 + J00 * J11) * id;
      sGeo[9 * NG + g] = det * s_w[g];
    }
    __syncwarp();

    // ---- B. stiffness tile
    double B[TI][TJ];
#pragma unroll
    for (int a = 0; a < TI; a++)
#pragma unroll
      for (int b = 0; b < TJ; b++) B[a][b] = 0.0;

#pragma unroll 2
    for (int g = 0; g < NG; g++) {
      double* G = sG + (g & 1) * 3 * GP;
      if (lane < NVE) {
        const double a = s_dx[g * NVE + lane], b = s_dy[g * NVE + lane], c = s_dz[g * NVE + lane];
        G[0 * GP + lane] = fma(c, sGeo[2 * NG + g], fma(b, sGeo[1 * NG + g], a * sGeo[0 * NG + g]));
        G[1 * GP + lane] = fma(c, sGeo[5 * NG + g], fma(b, sGeo[4 * NG + g], a * sGeo[3 * NG + g]));
        G[2 * GP + lane] = fma(c, sGeo[8 * NG + g], fma(b, sGeo[7 * NG + g], a * sGeo[6 * NG + g]));
      }
      __syncwarp();
      const double wg = sGeo[9 * NG + g];
      double gj[TJ][3];
#pragma unroll
      for (int b = 0; b < TJ; b++) {
        gj[b][0] = G[0 * GP + j0 + b] * wg;
        gj[b][1] = G[1 * GP + j0 + b] * wg;
        gj[b][2] = G[2 * GP + j0 + b] * wg;
      }
#pragma unroll
      for (int a = 0; a < TI; a++) {
        const double g0 = G[0 * GP + i0 + a], g1 = G[1 * GP + i0 + a], g2 = G[2 * GP + i0 + a];
#pragma unroll
        for (int b = 0; b < TJ; b++) B[a][b] = fma(g2, gj[b][2], fma(g1, gj[b][1], fma(g0, gj[b][0], B[a][b])));
      }
    }

    // ---- C. residual: F_i = fsrc * sum_g phi_i[g] weight_g - nu * (B u)_i
    if (rhs) {
      double rs[TI];
#pragma unroll
      for (int a = 0; a < TI; a++) {
        double s = 0.0;
#pragma unroll
        for (int b = 0; b < TJ; b++) s = fma(B[a][b], (j0 + b < NVE) ? sU[j0 + b] : 0.0, s);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        rs[a] = s;
      }
      double src = 0.0;
      if (lane < NVE) {
#pragma unroll 8
        for (int g = 0; g < NG; g++) src = fma(s_phi[g * NVE + lane], sGeo[9 * NG + g], src);
      }
      // row sums live (replicated) in every column group; lane (rg, cg=0) adds rows i0..i0+TI-1
#pragma unroll
      for (int a = 0; a < TI; a++) {
        const int i = i0 + a;
        const double srci = __shfl_sync(0xffffffffu, src, i < NVE ? i : 0);
        const int di = __shfl_sync(0xffffffffu, mydof, i < NVE ? i : 0);
        if (cg == 0 && i < NVE) atomicAdd(&rhs[di], fsrc * srci - nu * rs[a]);
      }
    }

    // ---- scatter the tile
    const SlotT* sl = slot + (size_t)e * (TI * TJ * 32) + lane;
#pragma unroll
    for (int a = 0; a < TI; a++) {
      const int i = i0 + a;
      const int di = __shfl_sync(0xffffffffu, mydof, i < NVE ? i : 0);
      if (i < NVE) {
        const int64_t base = rowptr[di];
#pragma unroll
        for (int b = 0; b < TJ; b++) {
          if (j0 + b < NVE) atomicAdd(&Aval[base + (int64_t)sl[(a * TJ + b) * 32]], nu * B[a][b]);
        }
      }
    }
    __syncwarp();
  }
}

// element -> CSR slot map in the tile layout of the assembly kernel
template <int NVE, typename SlotT>
__global__ void slot_map_kernel(int64_t nel, const int32_t* __restrict__ dof, const int64_t* __restrict__ rowptr,
                                const int32_t* __restrict__ col, SlotT* __restrict__ slot, int* err) {
  constexpr int TI = Tile<NVE>::TI, TJ = Tile<NVE>::TJ;
  const int lane = threadIdx.x & 31;
  const int rg = lane & 3, cg = lane >> 2;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e = w; e < nel; e += nw) {
    for (int a = 0; a < TI; a++)
      for (int b = 0; b < TJ; b++) {
        const int i = rg * TI + a, j = cg * TJ + b;
        SlotT out = 0;
        if (i < NVE && j < NVE) {
          const int32_t r = dof[e * NVE + i], c = dof[e * NVE + j];
          const int64_t s = rowptr[r], en = rowptr[r + 1];
          int64_t lo = s, hi = en;
          while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (col[mid] < c) lo = mid + 1;
            else hi = mid;
          }
          if (lo >= en || col[lo] != c) atomicExch(err, 1);
          out = (SlotT)(lo - s);
        }
        slot[(size_t)e * (TI * TJ * 32) + (a * TJ + b) * 32 + lane] = out;
      }
  }
}

template <int NVE, typename SlotT>
int build_slots(b2_asm* p) {
  b2_ctx* c = p->mesh->ctx;
  constexpr int per = Tile<NVE>::TI * Tile<NVE>::TJ * 32;
  p->slot_count = (size_t)p->mesh->nel * per;
  SlotT* s = nullptr;
  B2_TRY(b2_malloc(c, &s, p->slot_count));
  p->slot = s;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  const int grid = b2_grid_for(c, p->mesh->nel * 32, 256, 8);
  B2_LAUNCH(c, (slot_map_kernel<NVE, SlotT>), grid, 256, 0, p->mesh->nel, p->dof, p->A->rowptr, p->A->col, s, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_asm_create: an element couples dofs outside the matrix pattern");
  return 0;
}

template <int NVE, typename SlotT>
int launch_assemble(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_poisson_kernel<NVE, SlotT>;
  const size_t smem = SmemLayout<NVE>::bytes;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((p->mesh->nel + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, p->mesh->nel, p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dof,
            p->tab, (const SlotT*)p->slot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu, fsrc);
  return 0;
}

}  // namespace

extern "C" {

int b2_mesh_create(b2_ctx* c, int64_t nnode, int64_t nel, const double* xyz, const int32_t* conn, b2_mesh** out) {
  *out = nullptr;
  B2_CHECK(c && nnode > 0 && nel > 0 && xyz && conn, "b2_mesh_create: bad arguments");
  b2_mesh* m = new b2_mesh{c, nnode, nel, nullptr, nullptr};
  B2_TRY(b2_malloc(c, &m->xyz, (size_t)3 * nnode));
  B2_TRY(b2_malloc(c, &m->conn, (size_t)27 * nel));
  B2_TRY(b2_upload(c, m->xyz, xyz, (size_t)3 * nnode));
  B2_TRY(b2_upload(c, m->conn, conn, (size_t)27 * nel));
  *out = m;
  return 0;
}
int b2_mesh_update(b2_mesh* m, const double* xyz, const int32_t* conn) {
  // asynchronous re-upload (pinned host memory makes it a true async copy)
  if (xyz) B2_CUDA(cudaMemcpyAsync(m->xyz, xyz, (size_t)3 * m->nnode * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  if (conn) B2_CUDA(cudaMemcpyAsync(m->conn, conn, (size_t)27 * m->nel * sizeof(int32_t), cudaMemcpyHostToDevice, m->ctx->stream));
  return 0;
}
int b2_mesh_destroy(b2_mesh* m) {
  if (!m) return 0;
  cudaStreamSynchronize(m->ctx->stream);
  b2_free(m->ctx, m->xyz, (size_t)3 * m->nnode);
  b2_free(m->ctx, m->conn, (size_t)27 * m->nel);
  delete m;
  return 0;
}

int b2_asm_create(b2_mesh* m, b2_csr* A, int nve, const int32_t* dof, int ngauss, const double* phi,
                  const double* dxi, const double* deta, const double* dzeta, const double* weights, b2_asm** out) {
  *out = nullptr;
  B2_CHECK(m && A && dof && phi && dxi && deta && dzeta && weights, "b2_asm_create: null argument");
  B2_CHECK(nve == 8 || nve == 27, "b2_asm_create: nve=%d (supported: 8 trilinear, 27 triquadratic)", nve);
  B2_CHECK(ngauss == NG, "b2_asm_create: ngauss=%d (supported: 64, the 'seventh' hex rule)", ngauss);
  B2_CHECK(A->max_row <= 65536, "b2_asm_create: rows longer than 65536 entries");
  b2_ctx* c = m->ctx;
  b2_asm* p = new b2_asm();
  p->mesh = m;
  p->A = A;
  p->nve = nve;
  p->ngauss = ngauss;
  p->last_ms = 0.;
  B2_TRY(b2_malloc(c, &p->dof, (size_t)m->nel * nve));
  B2_TRY(b2_upload(c, p->dof, dof, (size_t)m->nel * nve));
  const size_t tn = (size_t)ngauss * nve;
  B2_TRY(b2_malloc(c, &p->tab, 4 * tn + ngauss));
  B2_TRY(b2_upload(c, p->tab, phi, tn));
  B2_TRY(b2_upload(c, p->tab + tn, dxi, tn));
  B2_TRY(b2_upload(c, p->tab + 2 * tn, deta, tn));
  B2_TRY(b2_upload(c, p->tab + 3 * tn, dzeta, tn));
  B2_TRY(b2_upload(c, p->tab + 4 * tn, weights, (size_t)ngauss));
  p->slot_bytes = A->max_row <= 256 ? 1 : 2;
  if (nve == 27) {
    if (p->slot_bytes == 1) B2_TRY((build_slots<27, uint8_t>(p)));
    else B2_TRY((build_slots<27, uint16_t>(p)));
  } else {
    if (p->slot_bytes == 1) B2_TRY((build_slots<8, uint8_t>(p)));
    else B2_TRY((build_slots<8, uint16_t>(p)));
  }
  *out = p;
  return 0;
}

int b2_asm_destroy(b2_asm* p) {
  if (!p) return 0;
  b2_ctx* c = p->mesh->ctx;
  cudaStreamSynchronize(c->stream);
  b2_free(c, p->dof, (size_t)p->mesh->nel * p->nve);
  b2_free(c, p->tab, (size_t)4 * p->ngauss * p->nve + p->ngauss);
  if (p->slot_bytes == 1) b2_free(c, (uint8_t*)p->slot, p->slot_count);
  else b2_free(c, (uint16_t*)p->slot, p->slot_count);
  delete p;
  return 0;
}

int b2_asm_poisson(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  B2_CHECK(!u || u->n >= p->A->nrows, "b2_asm_poisson: solution vector too short");
  B2_CHECK(!rhs || rhs->n >= p->A->nrows, "b2_asm_poisson: rhs vector too short");
  if (p->nve == 27) {
    if (p->slot_bytes == 1) return launch_assemble<27, uint8_t>(p, u, rhs, nu, fsrc);
    return launch_assemble<27, uint16_t>(p, u, rhs, nu, fsrc);
  }
  if (p->slot_bytes == 1) return launch_assemble<8, uint8_t>(p, u, rhs, nu, fsrc);
  return launch_assemble<8, uint16_t>(p, u, rhs, nu, fsrc);
}
double b2_asm_last_kernel_ms(const b2_asm* p) { return p->last_ms; }

}  // extern "C"
