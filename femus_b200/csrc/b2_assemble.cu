// Element assembly of the 3-D Poisson stiffness matrix / residual on HEX27 meshes, one warp per
// element.  Replaces the element loop of applications/001_Poisson/main.cpp:350-602 together with
// elem_type_3D::Jacobian_type (src/02_reference_geom_elements/03_fe_evaluations_at_quadrature/
// ElemType.hpp:1438-1537), MatSetValuesBlocked (PetscMatrix.cpp:699-729), VecSetValues
// (PetscVector.cpp:132-141), the Neumann face integrals (main.cpp:495-548, elem_type_2D::JacobianSur)
// and -- fused -- the first Galerkin product of LinearImplicitSystem.cpp:347-370.
//
// Kernels in this file
//   assemble_q2_mma_kernel      triquadratic elements, element matrix on the FP64 tensor cores (default)
//   assemble_poisson_kernel     CUDA-core register tiles: trilinear elements, and triquadratic on request
//   galerkin_from_elements_kernel   coarser Galerkin products from recorded element matrices
//   neumann_kernel              boundary faces
//
// Common structure per element (warp):
//   A. lanes = Gauss points (2 each for the 64-point rule): J = sum_n dphi[g][n] x[n] with the
//      shape-derivative tables staged in shared memory once per CTA, then det, J^-1, weight.
//   B. element matrix B_ij = sum_g w_g grad phi_i . grad phi_j (tensor-core GEMM or register tiles).
//   C. residual F_i = fsrc * sum_g phi_i weight - (B u)_i, scatter: fp64 atomicAdd into the CSR
//      through a precomputed element->slot map (no searches), and into rhs.
//   D. (fused Galerkin) D_e = Pc^T B Pc, summed over the 8 children of a coarse element, scattered
//      to the coarse matrix and recorded for the element-matrix chain.
// The geometry map uses the unknown's own family and its first nve nodes, like the reference
// (ElemType.hpp:1462).  HBM traffic: 27 node ids + 81 coordinates in, 729 atomics out per element.
#include "b2_common.cuh"

struct b2_mesh {
  b2_ctx* ctx;
  int64_t nnode, nel;
  double* xyz;     // [3][nnode]
  int32_t* conn;   // [nel][27]
  double* xyz_shadow;    // second copy of both, allocated by the first b2_mesh_prefetch: the next step's
  int32_t* conn_shadow;  // mesh is uploaded here while the current step still reads xyz / conn
};

void b2_mesh_view(const b2_mesh* m, b2_ctx** ctx, int64_t* nnode, int64_t* nel, const double** xyz, const int32_t** conn) {
  *ctx = m->ctx;
  *nnode = m->nnode;
  *nel = m->nel;
  *xyz = m->xyz;
  *conn = m->conn;
}

struct b2_asm {
  b2_mesh* mesh;   // borrowed
  b2_csr* A;       // borrowed
  b2_ctx* ctx = nullptr;   // the borrowed objects may be gone when the plan is destroyed: what b2_asm_destroy needs is kept here
  int64_t nel = 0;
  int nve, ngauss;
  bool general;    // not a hexahedron with the 64-point rule: table-driven kernel
  int32_t* dof;    // [nel][nve]
  double* tab;     // phi, dxi, deta, dzeta [ng][nve] each, then w[ng]
  void* slot;      // [nel][TI*TJ][32] uint8 or uint16: position of (i,j) inside row dof_i (tile kernel)
  void* nslot;     // [nel][nve*nve] the same in natural (i,j) order (tensor-core kernel, nve = 27)
  int slot_bytes;  // 1 or 2
  size_t slot_count, nslot_count;
  double last_ms;
  const char* last_kernel = "";      // name of the kernel the last b2_asm_poisson* call launched
  // fused Galerkin plan (b2_asm_poisson_galerkin): per-child element prolongators of the plan `gal`
  const b2_galerkin* gal;
  void* gal_tab;          // device: GalTables<nve>
  // sum-factorised kernel (b2_assemble_sumfac.cuh): 1-D tables, dofs and slot map in lattice order, Kronecker factors
  // of the child prolongators; sf_tab == null: the tables are not a 3 x 3 x 3 / 4 x 4 x 4 tensor product (kernel not used)
  void* sf_tab;           // device: SfTables
  int32_t* dofL;          // [nel][27]
  void* lslot;            // [nel][736] uint8 or uint16 (729 entries in lattice order + padding)
  void* sf_gal;           // device: SfGalTables of the plan `gal`, or null (the child prolongators are not Kronecker products)
};

// what the fused kernel needs from a Galerkin plan (b2_galerkin.cu)
struct b2_galerkin_view {
  b2_csr *Af, *Ac;
  int64_t nelc;
  int nf, nc;
  const int32_t *fd, *cd;
  const double* ploc;
  const uint8_t *fmask, *cmask;
  const void* slot;
  int slot_bytes;
  double** emat;        // storage slots owned by the plan (element-matrix chain)
  void** chain_tab;
  int* chain_tab_nve;
  void** chain_sf;      // device SfGalTables of the chain (sum-factorised chain kernel), or null
  int* chain_sf_tried;
};
int b2_galerkin_get_view(const b2_galerkin* g, b2_galerkin_view* v);

namespace {

constexpr int NG = 64;          // Gauss points per element ("seventh" hex rule)
constexpr int kWarps = 16;      // warps (= elements in flight) per CTA, one CTA per SM
constexpr int GP = 32;          // padded node stride of the gradient buffers

template <int NVE> struct Tile;
template <> struct Tile<27> { static constexpr int TI = 7, TJ = 4; };   // 4 x 7 lane grid, 28 lanes busy
template <> struct Tile<8> { static constexpr int TI = 2, TJ = 1; };    // 4 x 8 lane grid

// Element prolongator of each of the 8 children of a refined hexahedron, restricted to the child's
// own NVE nodes: Pc[j][n][J] = ploc[lattice(j, n)][J], stored by columns as compressed lists
// (both products of the fused kernel walk columns); exact zeros dropped (Q2: 125 entries per child).
template <int NVE>
struct GalTables {
  static constexpr int MAXNNZ = NVE == 27 ? 128 : 64;
  int colptr[8][NVE + 1];
  unsigned char crow[8][MAXNNZ];   // fine node n of every entry of column J
  double cval[8][MAXNNZ];
};

struct GalArgs {
  const void* tab;            // GalTables<NVE> (device)
  const int32_t* cd;          // [nelc][NVE] coarse dofs
  const void* cslot;          // [nelc][NVE*NVE] slot of (I,J) inside row cd_I of the coarse matrix
  const uint8_t* fmask;       // fine Dirichlet rows of P (may be null)
  const uint8_t* cmask;       // coarse Dirichlet columns of P (may be null)
  const int64_t* Cp;          // coarse rowptr
  double* Cv;                 // coarse values
  double* emat;               // [nelc][NVE*NVE] record of every coarse element's Galerkin matrix, or null
};

#include "b2_assemble_sumfac.cuh"
#include "b2_assemble_sumfac_host.hpp"

template <int NVE>
struct SmemLayout {
  static constexpr int tab_doubles = 4 * NG * NVE + NG;
  // per warp: X[3][GP], U[GP], geo[10][NG], G[2][3][GP]
  static constexpr int warp_doubles = 3 * GP + GP + 10 * NG + 2 * 3 * GP;
  static constexpr size_t bytes = (size_t)(tab_doubles + kWarps * warp_doubles) * sizeof(double);
  static constexpr size_t bytes_gal = bytes + sizeof(GalTables<NVE>);
  // the per-warp geometry + gradient buffers (10*NG + 6*GP doubles) are reused for the element matrix
  static_assert(10 * NG + 2 * 3 * GP >= NVE * NVE, "element matrix does not fit the reused buffers");
};

// GAL: additionally forms the Galerkin coarse operator C = P^T A P of the next-coarser level from
// the element matrices while they are still on chip (see the header of this file).
template <int NVE, typename SlotT, bool GAL, typename CSlotT>
__global__ void __launch_bounds__(kWarps * 32, 1)
assemble_poisson_kernel(int64_t nel, int64_t nnode, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
                        const int32_t* __restrict__ dof, const double* __restrict__ tab,
                        const SlotT* __restrict__ slot, const int64_t* __restrict__ rowptr, double* __restrict__ Aval,
                        const double* __restrict__ u, double* __restrict__ rhs, double nu, double fsrc, const GalArgs ga) {
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
  const GalTables<NVE>* gt = nullptr;
  if (GAL) {
    int* dst = reinterpret_cast<int*>(smem + SmemLayout<NVE>::tab_doubles + kWarps * SmemLayout<NVE>::warp_doubles);
    const int* src = reinterpret_cast<const int*>(ga.tab);
    for (int t = threadIdx.x; t < (int)(sizeof(GalTables<NVE>) / 4); t += blockDim.x) dst[t] = src[t];
    gt = reinterpret_cast<const GalTables<NVE>*>(dst);
  }
  __syncthreads();

  const int rg = lane & 3, cg = lane >> 2;                      // row group / column group of the tile
  const int i0 = rg * TI, j0 = cg * TJ;

  // In one trip the 8 warps of a group hold the 8 children of ONE coarse element (kWarps
  // consecutive elements per CTA and trip; GAL requires nel to be a multiple of 8, so a group is
  // either complete or absent and its named barrier below is always reached by all 8 warps).
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

    if (GAL) {
      // ---- D = Pc^T (nu B) Pc for this child; the element matrix goes to shared memory (geometry
      //      and gradient buffers are dead now), rows/columns of Dirichlet fine dofs dropped
      double* Bs = sGeo;                       // [NVE][NVE]
      const int child = (int)(e & 7);
      const int fm = (lane < NVE && ga.fmask) ? (int)ga.fmask[mydof] : 0;
#pragma unroll
      for (int a = 0; a < TI; a++) {
        const int i = i0 + a;
        const int fi = __shfl_sync(0xffffffffu, fm, i < NVE ? i : 0);
#pragma unroll
        for (int b = 0; b < TJ; b++) {
          const int j = j0 + b;
          const int fj = __shfl_sync(0xffffffffu, fm, j < NVE ? j : 0);
          if (i < NVE && j < NVE) Bs[i * NVE + j] = (fi | fj) ? 0.0 : nu * B[a][b];
        }
      }
      __syncwarp();
      // T = Bs Pc.  Lane J walks the compressed column J of Pc (<= 8 entries); for every entry
      // (n, v) it updates its whole column T[0..NVE)[J] += Bs[.][n] v: NVE independent FMAs per
      // step, registers with static indices.
      double R[NVE];
#pragma unroll
      for (int i = 0; i < NVE; i++) R[i] = 0.0;
      if (lane < NVE) {
#pragma unroll 1
        for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
          const int n = gt->crow[child][q];
          const double v = gt->cval[child][q];
#pragma unroll
          for (int i = 0; i < NVE; i++) R[i] = fma(Bs[i * NVE + n], v, R[i]);
        }
      }
      __syncwarp();
      if (lane < NVE) {
#pragma unroll
        for (int i = 0; i < NVE; i++) Bs[i * NVE + lane] = R[i];      // Bs now holds T
      }
      __syncwarp();
      // D = Pc^T T.  Lane I walks the compressed column I of Pc; for every entry (i, v) it updates
      // its whole row D[I][0..NVE) += v T[i][.]
#pragma unroll
      for (int j = 0; j < NVE; j++) R[j] = 0.0;
      if (lane < NVE) {
#pragma unroll 1
        for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
          const int i = gt->crow[child][q];
          const double v = gt->cval[child][q];
#pragma unroll
          for (int j = 0; j < NVE; j++) R[j] = fma(v, Bs[i * NVE + j], R[j]);
        }
      }
      __syncwarp();
      if (lane < NVE) {
#pragma unroll
        for (int j = 0; j < NVE; j++) Bs[lane * NVE + j] = R[j];      // Bs now holds D
      }
      // ---- the 8 children of a coarse element sit in the 8 warps of a group: sum their D and
      //      scatter once per coarse element (one atomic per entry instead of eight).  Named
      //      barrier per group: the other group keeps running.
      const int grp = wib >> 3;
      asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
      {
        const int64_t E = e >> 3;
        const CSlotT* cslot = reinterpret_cast<const CSlotT*>(ga.cslot) + (size_t)E * (NVE * NVE);
        const double* D0 = s_w + NG + (size_t)(8 * grp) * SmemLayout<NVE>::warp_doubles + (3 * GP + GP);
        for (int idx = threadIdx.x & 255; idx < NVE * NVE; idx += 256) {
          double v = 0.0;
#pragma unroll
          for (int w = 0; w < 8; w++) v += D0[(size_t)w * SmemLayout<NVE>::warp_doubles + idx];
          if (ga.emat) ga.emat[(size_t)E * (NVE * NVE) + idx] = v;
          if (v == 0.0) continue;
          const int I = idx / NVE, J = idx - I * NVE;
          const int32_t dI = ga.cd[E * NVE + I];
          if (ga.cmask) {
            const int32_t dJ = ga.cd[E * NVE + J];
            if (ga.cmask[dI] || ga.cmask[dJ]) continue;
          }
          atomicAdd(&ga.Cv[ga.Cp[dI] + (int64_t)cslot[idx]], v);
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------
// Triquadratic elements on the FP64 tensor cores.  With D_g the 27 x 3 REFERENCE gradients at Gauss
// point g (the same table for every element) and K_g = w_g det_g J_g^-1 J_g^-T (3 x 3, symmetric),
// B = sum_g D_g K_g D_g^T = D (K D^T) is a 27 x 27 x 192 GEMM per element whose left operand never
// changes: its fragments are read straight from the table in shared memory, only the right operand
// D_g K_g (9 FMA per node and point) is formed and staged per element.  The GEMM runs as 48 k-steps of mma.sync.m8n8k4.f64 (DMMA.8x8x4
// is the only fp64 tensor shape of sm_100a: m16n8k8 compiles to four of them) on the 4 x 4 grid of
// 8 x 8 output tiles, of which only the 10 upper ones are formed (B is symmetric).  Against the CUDA-core tile kernel above this needs 8
// shared-memory fragment loads per lane and Gauss point instead of 33 (that kernel is bound by
// shared-memory wavefronts, ncu: l1tex data pipe 94 %, fp64 pipe 57 %).  The element matrix is then
// written to shared memory once; residual, scatter and the fused Galerkin product read it from there,
// so the scatter uses a slot map in natural (i, j) order.
constexpr int MS = 36;            // column stride of the fragment buffers: conflict-free 8-byte fragment loads
constexpr int KT = 3 * NG;        // K dimension of the element GEMM: (Gauss point, reference direction)
struct MmaSmem {
  // reference gradients twice: D[k = 3g + a][node] with row stride MS (fragment loads of the GEMM, rows of
  // D_g K_g) and [a][g][node] with row stride 27 (geometry phase, lanes = Gauss points: conflict-free), weights
  static constexpr int tab_doubles = KT * MS + 3 * NG * 27 + NG;
  // per warp: X[3][GP], U[GP], rowbase[GP] (int64), geo[7][NG], M[2 halves][4][MS]
  static constexpr int warp_doubles = 3 * GP + GP + GP + 7 * NG + 2 * 4 * MS;
  static constexpr size_t bytes = (size_t)(tab_doubles + kWarps * warp_doubles) * sizeof(double);
  static constexpr size_t bytes_gal = bytes + sizeof(GalTables<27>);
  static_assert(7 * NG + 2 * 4 * MS >= 27 * 27, "element matrix does not fit the reused buffers");
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename SlotT, bool GAL, typename CSlotT>
__global__ void __launch_bounds__(kWarps * 32, 1)
assemble_q2_mma_kernel(int64_t nel, int64_t nnode, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
                       const int32_t* __restrict__ dof, const double* __restrict__ tab, const SlotT* __restrict__ nslot,
                       const int64_t* __restrict__ rowptr, double* __restrict__ Aval, const double* __restrict__ u,
                       double* __restrict__ rhs, double nu, double fsrc, const GalArgs ga) {
  constexpr int NVE = 27;
  extern __shared__ double smem[];
  double* sT = smem;                                    // [KT][MS] reference gradients, row k = 3 g + a
  double* s_dx = sT + KT * MS;                          // [3][NG][27] the same, geometry-phase layout
  double* s_dy = s_dx + NG * NVE;
  double* s_dz = s_dy + NG * NVE;
  double* s_w = s_dz + NG * NVE;                        // [NG] Gauss weights
  const double* g_phi = tab;          // the shape-value table is only needed for the source term: read through L1
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = s_w + NG + wib * MmaSmem::warp_doubles;
  double* sX = wbase;                                   // [3][GP]
  double* sU = sX + 3 * GP;                             // [GP]
  long long* sRow = reinterpret_cast<long long*>(sU + GP);   // [GP] rowptr[dof_i]
  double* sGeo = reinterpret_cast<double*>(sRow + GP);  // [7][NG]: K00 K01 K02 K11 K12 K22, weight
  double* sM = sGeo + 7 * NG;                           // [2][4][MS]: two k-steps of D K, (k, node)
  double* Bs = sGeo;                                    // [27][27] element matrix, reuses geo + M

  for (int t = threadIdx.x; t < KT * MS; t += blockDim.x) {
    const int k = t / MS, n = t - k * MS, g = k / 3, a = k - 3 * g;
    sT[t] = n < NVE ? tab[(1 + a) * NG * NVE + g * NVE + n] : 0.0;
  }
  for (int t = threadIdx.x; t < 3 * NG * NVE + NG; t += blockDim.x) s_dx[t] = tab[NG * NVE + t];
  const GalTables<NVE>* gt = nullptr;
  if (GAL) {
    int* dst = reinterpret_cast<int*>(smem + MmaSmem::tab_doubles + kWarps * MmaSmem::warp_doubles);
    const int* src = reinterpret_cast<const int*>(ga.tab);
    for (int t = threadIdx.x; t < (int)(sizeof(GalTables<NVE>) / 4); t += blockDim.x) dst[t] = src[t];
    gt = reinterpret_cast<const GalTables<NVE>*>(dst);
  }
  __syncthreads();

  // fragment coordinates of this lane: row l/4 of an 8-row tile, k-column l%4; outputs (l/4, 2(l%4)+{0,1})
  const int fr = lane >> 2, fk = lane & 3;

  for (int64_t e = (int64_t)blockIdx.x * kWarps + wib; e < nel; e += (int64_t)gridDim.x * kWarps) {
    // ---- gather: node ids (coalesced), coordinates, dofs, current solution, row starts
    int mydof = 0;
    if (lane < NVE) {
      const int64_t nd = conn[e * 27 + lane];
      sX[0 * GP + lane] = xyz[nd];
      sX[1 * GP + lane] = xyz[nnode + nd];
      sX[2 * GP + lane] = xyz[2 * nnode + nd];
      mydof = dof[e * NVE + lane];
      sU[lane] = u ? u[mydof] : 0.0;
      sRow[lane] = rowptr[mydof];
    }
    for (int t = lane; t < 2 * 4 * MS; t += 32) sM[t] = 0.0;      // padding rows 27..31 stay zero
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
      // JacI[d][a] as ElemType.hpp:1478-1486; grad phi_n [d] = sum_a dphi_n/dxi_a JacI[d][a]
      const double I00 = (-J12 * J21 + J11 * J22) * id, I01 = (J02 * J21 - J01 * J22) * id, I02 = (-J02 * J11 + J01 * J12) * id;
      const double I10 = (J12 * J20 - J10 * J22) * id, I11 = (-J02 * J20 + J00 * J22) * id, I12 = (J02 * J10 - J00 * J12) * id;
      const double I20 = (-J11 * J20 + J10 * J21) * id, I21 = (J01 * J20 - J00 * J21) * id, I22 = (-J01 * J10 + J00 * J11) * id;
      const double wd = det * s_w[g];
      // K_ab = weight * sum_d JacI[d][a] JacI[d][b]
      sGeo[0 * NG + g] = wd * fma(I20, I20, fma(I10, I10, I00 * I00));
      sGeo[1 * NG + g] = wd * fma(I20, I21, fma(I10, I11, I00 * I01));
      sGeo[2 * NG + g] = wd * fma(I20, I22, fma(I10, I12, I00 * I02));
      sGeo[3 * NG + g] = wd * fma(I21, I21, fma(I11, I11, I01 * I01));
      sGeo[4 * NG + g] = wd * fma(I21, I22, fma(I11, I12, I01 * I02));
      sGeo[5 * NG + g] = wd * fma(I22, I22, fma(I12, I12, I02 * I02));
      sGeo[6 * NG + g] = wd;
    }
    __syncwarp();

    // ---- B. stiffness on the tensor cores: 10 upper 8x8 tiles, 2 accumulators per lane and tile
    double C[10][2];
#pragma unroll
    for (int t = 0; t < 10; t++) C[t][0] = C[t][1] = 0.0;
    double src = 0.0;
    // The GEMM's K dimension is (Gauss point, component) flattened, 192 = 48 k-steps of 4 with no
    // padding: a Gauss point contributes 3 consecutive columns, so 4 points make 3 k-steps.  The two
    // halves of the fragment buffer hold alternate k-steps; a half is rewritten only after a
    // __syncwarp that follows the fragment loads of the k-step it held before (see the order below).
    auto grad = [&](int g, double& h0, double& h1, double& h2, double& wg) {     // row `lane` of D_g K_g
      wg = sGeo[6 * NG + g];
      h0 = h1 = h2 = 0.0;
      if (lane < NVE) {
        const double a = sT[(3 * g + 0) * MS + lane], b = sT[(3 * g + 1) * MS + lane], c = sT[(3 * g + 2) * MS + lane];
        const double K00 = sGeo[0 * NG + g], K01 = sGeo[1 * NG + g], K02 = sGeo[2 * NG + g];
        const double K11 = sGeo[3 * NG + g], K12 = sGeo[4 * NG + g], K22 = sGeo[5 * NG + g];
        h0 = fma(c, K02, fma(b, K01, a * K00));
        h1 = fma(c, K12, fma(b, K11, a * K01));
        h2 = fma(c, K22, fma(b, K12, a * K02));
        if (rhs) src = fma(__ldg(g_phi + g * NVE + lane), wg, src);
      }
    };
    auto put = [&](int kstep, int col, double v, double) {       // column `col` of k-step `kstep` of D K
      if (lane < NVE) sM[(kstep & 1) * (4 * MS) + col * MS + lane] = v;
    };
    auto mma_step = [&](int kstep) {
      const double* M = sM + (kstep & 1) * (4 * MS);
      const double* D = sT + (4 * kstep) * MS;          // rows k = 4 kstep .. 4 kstep + 3 of the reference-gradient table
      double fa[4], fb[4];
#pragma unroll
      for (int T = 0; T < 4; T++) {
        fa[T] = D[fk * MS + 8 * T + fr];                // D[8T + l/4][k]
        fb[T] = M[fk * MS + 8 * T + fr];                // (D K)[8T + l/4][k]
      }
      int t = 0;
#pragma unroll
      for (int I = 0; I < 4; I++)
#pragma unroll
        for (int Jt = I; Jt < 4; Jt++) {
          dmma(C[t][0], C[t][1], fa[I], fb[Jt]);
          t++;
        }
    };
#pragma unroll 2
    for (int q = 0; q < NG / 4; q++) {
      const int s0 = 3 * q;
      double g0, g1, g2, wg;
      grad(4 * q + 0, g0, g1, g2, wg);
      put(s0, 0, g0, wg); put(s0, 1, g1, wg); put(s0, 2, g2, wg);
      __syncwarp();
      grad(4 * q + 1, g0, g1, g2, wg);
      put(s0, 3, g0, wg); put(s0 + 1, 0, g1, wg); put(s0 + 1, 1, g2, wg);
      __syncwarp();
      mma_step(s0);
      grad(4 * q + 2, g0, g1, g2, wg);
      put(s0 + 1, 2, g0, wg); put(s0 + 1, 3, g1, wg);
      __syncwarp();
      mma_step(s0 + 1);
      put(s0 + 2, 0, g2, wg);          // the half of k-step s0: its fragments were loaded before the last __syncwarp
      grad(4 * q + 3, g0, g1, g2, wg);
      put(s0 + 2, 1, g0, wg); put(s0 + 2, 2, g1, wg); put(s0 + 2, 3, g2, wg);
      __syncwarp();
      mma_step(s0 + 2);
    }
    __syncwarp();

    // ---- C. element matrix -> shared memory (both triangles), scaled by nu
    {
      int t = 0;
#pragma unroll
      for (int I = 0; I < 4; I++)
#pragma unroll
        for (int Jt = I; Jt < 4; Jt++) {
          const int i = 8 * I + fr;
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const int j = 8 * Jt + 2 * fk + r;
            if (i < NVE && j < NVE) {
              const double v = nu * C[t][r];
              Bs[i * NVE + j] = v;
              if (I != Jt) Bs[j * NVE + i] = v;
            }
          }
          t++;
        }
    }
    __syncwarp();

    // ---- residual F_i = fsrc * sum_g phi_i w_g - (B u)_i
    if (rhs && lane < NVE) {
      double s = 0.0;
#pragma unroll 9
      for (int j = 0; j < NVE; j++) s = fma(Bs[lane * NVE + j], sU[j], s);
      atomicAdd(&rhs[mydof], fsrc * src - s);
    }
    // ---- scatter: natural (i, j) order through the slot map
    {
      const SlotT* sl = nslot + (size_t)e * (NVE * NVE);
      for (int idx = lane; idx < NVE * NVE; idx += 32) {
        const int i = idx / NVE;
        atomicAdd(&Aval[sRow[i] + (long long)sl[idx]], Bs[idx]);
      }
    }
    __syncwarp();

    if (GAL) {
      const int child = (int)(e & 7);
      // rows/columns of Dirichlet fine dofs do not take part in the Galerkin product
      if (ga.fmask) {
        const int fm = lane < NVE ? (int)ga.fmask[mydof] : 0;
        const unsigned mask = __ballot_sync(0xffffffffu, fm != 0);
        if (mask) {
          for (int idx = lane; idx < NVE * NVE; idx += 32) {
            const int i = idx / NVE, j = idx - i * NVE;
            if (((mask >> i) | (mask >> j)) & 1u) Bs[idx] = 0.0;
          }
          __syncwarp();
        }
      }
      double R[NVE];
#pragma unroll
      for (int i = 0; i < NVE; i++) R[i] = 0.0;
      if (lane < NVE) {
#pragma unroll 1
        for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
          const int n = gt->crow[child][q];
          const double v = gt->cval[child][q];
#pragma unroll
          for (int i = 0; i < NVE; i++) R[i] = fma(Bs[i * NVE + n], v, R[i]);
        }
      }
      __syncwarp();
      if (lane < NVE) {
#pragma unroll
        for (int i = 0; i < NVE; i++) Bs[i * NVE + lane] = R[i];      // Bs now holds T = B Pc
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NVE; j++) R[j] = 0.0;
      if (lane < NVE) {
#pragma unroll 1
        for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
          const int i = gt->crow[child][q];
          const double v = gt->cval[child][q];
#pragma unroll
          for (int j = 0; j < NVE; j++) R[j] = fma(v, Bs[i * NVE + j], R[j]);
        }
      }
      __syncwarp();
      if (lane < NVE) {
#pragma unroll
        for (int j = 0; j < NVE; j++) Bs[lane * NVE + j] = R[j];      // Bs now holds D = Pc^T B Pc
      }
      const int grp = wib >> 3;
      asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
      {
        const int64_t E = e >> 3;
        const CSlotT* cslot = reinterpret_cast<const CSlotT*>(ga.cslot) + (size_t)E * (NVE * NVE);
        const double* D0 = s_w + NG + (size_t)(8 * grp) * MmaSmem::warp_doubles + (3 * GP + GP + GP);
        for (int idx = threadIdx.x & 255; idx < NVE * NVE; idx += 256) {
          double v = 0.0;
#pragma unroll
          for (int w = 0; w < 8; w++) v += D0[(size_t)w * MmaSmem::warp_doubles + idx];
          if (ga.emat) ga.emat[(size_t)E * (NVE * NVE) + idx] = v;
          if (v == 0.0) continue;
          const int I = idx / NVE, J = idx - I * NVE;
          const int32_t dI = ga.cd[E * NVE + I];
          if (ga.cmask) {
            const int32_t dJ = ga.cd[E * NVE + J];
            if (ga.cmask[dI] || ga.cmask[dJ]) continue;
          }
          atomicAdd(&ga.Cv[ga.Cp[dI] + (int64_t)cslot[idx]], v);
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
    }
  }
}

// element -> CSR slot map in natural (i, j) order
template <typename SlotT>
__global__ void natural_slot_kernel(int64_t total, int nve, const int32_t* __restrict__ dof, const int64_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ col, SlotT* __restrict__ slot, int* err, int out_stride = 0) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nn = nve * nve;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t e = t / nn;
    const int idx = (int)(t - e * nn);
    const int i = idx / nve, j = idx - i * nve;
    const int32_t r = dof[e * nve + i], c = dof[e * nve + j];
    const int64_t s = rowptr[r], en = rowptr[r + 1];
    int64_t lo = s, hi = en;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (col[mid] < c) lo = mid + 1;
      else hi = mid;
    }
    if (lo >= en || col[lo] != c) atomicExch(err, 1);
    slot[out_stride ? e * out_stride + idx : t] = (SlotT)(lo - s);
  }
}

// ------------------------------------------------------------------------------------------
// Galerkin product of a COARSER level pair from recorded element matrices.  The fused assembly (and
// this kernel itself) records, for every element E of level l, its Galerkin element matrix
// D_E = sum over the 8 children of Pc^T B Pc.  Since A_l = sum_E D_E (scattered), the next operator is
// A_{l-1} = sum_E Pc(E)^T D_E Pc(E), again 8 children per coarser element: the chain never re-reads
// an assembled matrix (12 B per nonzero) -- it streams 5.8 KB per element once.
template <int NVE, typename CSlotT>
__global__ void __launch_bounds__(kWarps * 32, 1)
galerkin_from_elements_kernel(int64_t nel, const double* __restrict__ emat_in, const int32_t* __restrict__ dof,
                              const GalArgs ga) {
  extern __shared__ double smem[];
  constexpr int WD = NVE * NVE + 3;                                  // odd stride between the warps' matrices
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* Bs = smem + (size_t)wib * WD;
  const GalTables<NVE>* gt;
  {
    int* dst = reinterpret_cast<int*>(smem + (size_t)kWarps * WD);
    const int* src = reinterpret_cast<const int*>(ga.tab);
    for (int t = threadIdx.x; t < (int)(sizeof(GalTables<NVE>) / 4); t += blockDim.x) dst[t] = src[t];
    gt = reinterpret_cast<const GalTables<NVE>*>(dst);
  }
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * kWarps + wib; e < nel; e += (int64_t)gridDim.x * kWarps) {
    const int child = (int)(e & 7);
    const double* De = emat_in + (size_t)e * (NVE * NVE);
    for (int idx = lane; idx < NVE * NVE; idx += 32) Bs[idx] = De[idx];
    // rows/columns of this level's Dirichlet dofs do not take part (rows of P zeroed)
    unsigned mask = 0;
    if (ga.fmask) {
      const int fm = lane < NVE ? (int)ga.fmask[dof[e * NVE + lane]] : 0;
      mask = __ballot_sync(0xffffffffu, fm != 0);
    }
    __syncwarp();
    if (mask) {
      for (int idx = lane; idx < NVE * NVE; idx += 32) {
        const int i = idx / NVE, j = idx - i * NVE;
        if (((mask >> i) | (mask >> j)) & 1u) Bs[idx] = 0.0;
      }
      __syncwarp();
    }
    double R[NVE];
#pragma unroll
    for (int i = 0; i < NVE; i++) R[i] = 0.0;
    if (lane < NVE) {
#pragma unroll 1
      for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
        const int n = gt->crow[child][q];
        const double v = gt->cval[child][q];
#pragma unroll
        for (int i = 0; i < NVE; i++) R[i] = fma(Bs[i * NVE + n], v, R[i]);
      }
    }
    __syncwarp();
    if (lane < NVE) {
#pragma unroll
      for (int i = 0; i < NVE; i++) Bs[i * NVE + lane] = R[i];      // T = D Pc
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NVE; j++) R[j] = 0.0;
    if (lane < NVE) {
#pragma unroll 1
      for (int q = gt->colptr[child][lane]; q < gt->colptr[child][lane + 1]; q++) {
        const int i = gt->crow[child][q];
        const double v = gt->cval[child][q];
#pragma unroll
        for (int j = 0; j < NVE; j++) R[j] = fma(v, Bs[i * NVE + j], R[j]);
      }
    }
    __syncwarp();
    if (lane < NVE) {
#pragma unroll
      for (int j = 0; j < NVE; j++) Bs[lane * NVE + j] = R[j];      // Pc^T D Pc
    }
    const int grp = wib >> 3;
    asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
    {
      const int64_t E = e >> 3;
      const CSlotT* cslot = reinterpret_cast<const CSlotT*>(ga.cslot) + (size_t)E * (NVE * NVE);
      const double* D0 = smem + (size_t)(8 * grp) * WD;
      for (int idx = threadIdx.x & 255; idx < NVE * NVE; idx += 256) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += D0[(size_t)w * WD + idx];
        if (ga.emat) ga.emat[(size_t)E * (NVE * NVE) + idx] = v;
        if (v == 0.0) continue;
        const int I = idx / NVE, J = idx - I * NVE;
        const int32_t dI = ga.cd[E * NVE + I];
        if (ga.cmask) {
          const int32_t dJ = ga.cd[E * NVE + J];
          if (ga.cmask[dI] || ga.cmask[dJ]) continue;
        }
        atomicAdd(&ga.Cv[ga.Cp[dI] + (int64_t)cslot[idx]], v);
      }
    }
    asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
  }
}

// host-side check of the fused plan: fine element 8E+j must be child j of coarse element E with its
// local node n at lattice point lat[j][n] of E
__global__ void gal_check_kernel(int64_t nelc, int nve, int nf, const int32_t* __restrict__ dof,
                                 const int32_t* __restrict__ fd, const unsigned char* __restrict__ lat, int* err) {
  const int64_t total = nelc * 8 * nve;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t E = t / (8 * nve);
    const int jn = (int)(t - E * 8 * nve);
    if (dof[(E * 8) * nve + jn] != fd[E * nf + lat[jn]]) atomicExch(err, 1);
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

template <typename SlotT>
int build_natural_slots(b2_asm* p) {
  b2_ctx* c = p->mesh->ctx;
  const int64_t total = p->mesh->nel * p->nve * p->nve;
  p->nslot_count = (size_t)total;
  SlotT* s = nullptr;
  B2_TRY(b2_malloc(c, &s, p->nslot_count));
  p->nslot = s;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  B2_LAUNCH(c, natural_slot_kernel<SlotT>, b2_grid_for(c, total, 256, 8), 256, 0, total, p->nve, p->dof, p->A->rowptr,
            p->A->col, s, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_asm_create: an element couples dofs outside the matrix pattern");
  return 0;
}


// ------------------------------------------------------------------------------------------
// Host side of the sum-factorised kernel (b2_assemble_sumfac.cuh; table factorisation in b2_assemble_sumfac_host.hpp).
// the stage-3 coefficients live in constant memory: one set per process (every Hex27 / 64-point plan has the same
// tables; a plan whose tables differ from the loaded set stays on the tensor-core kernel)
static bool g_sfM_loaded = false;
static double g_sfM[4][9][4];

template <typename SlotT>
static int sf_build_slots(b2_asm* p) {
  b2_ctx* c = p->mesh->ctx;
  const int64_t total = p->mesh->nel * 729;
  SlotT* s = nullptr;
  B2_TRY(b2_malloc(c, &s, (size_t)p->mesh->nel * kSfSlotStride));      // 729 entries per element, padded to 16-byte multiples
  B2_CUDA(cudaMemsetAsync(s, 0, (size_t)p->mesh->nel * kSfSlotStride * sizeof(SlotT), c->stream));
  p->lslot = s;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  B2_LAUNCH(c, natural_slot_kernel<SlotT>, b2_grid_for(c, total, 256, 8), 256, 0, total, 27, p->dofL, p->A->rowptr, p->A->col, s, d_err, kSfSlotStride);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_asm_create: an element couples dofs outside the matrix pattern");
  return 0;
}

// called by b2_asm_create for 27 dofs / 64 points; leaves p->sf_tab null when the kernel does not apply
static int sf_prepare(b2_asm* p, const int32_t* dof, const double* phi, const double* dxi, const double* deta, const double* dzeta, const double* w) {
  b2_ctx* c = p->mesh->ctx;
  SfTables T;
  if (!sf_factor_tables(phi, dxi, deta, dzeta, w, &T)) return 0;
  if (!g_sfM_loaded) {
    B2_CUDA(cudaMemcpyToSymbol(c_sfM, T.M, sizeof(T.M)));
    B2_CUDA(cudaMemcpyToSymbol(c_sfU, T.L, sizeof(T.L) + sizeof(T.D)));      // L then D, contiguous in SfTables
    memcpy(g_sfM, T.M, sizeof(T.M));
    g_sfM_loaded = true;
  } else if (memcmp(g_sfM, T.M, sizeof(T.M)) != 0) {
    return 0;
  }
  const int64_t nel = p->mesh->nel;
  std::vector<int32_t> dl((size_t)nel * 27);
  for (int64_t e = 0; e < nel; e++)
    for (int m = 0; m < 27; m++) dl[(size_t)e * 27 + m] = dof[(size_t)e * 27 + T.node_of[m]];
  B2_TRY(b2_malloc(c, &p->dofL, (size_t)nel * 27));
  B2_TRY(b2_upload(c, p->dofL, dl.data(), (size_t)nel * 27));
  if (p->slot_bytes == 1) B2_TRY(sf_build_slots<uint8_t>(p));
  else B2_TRY(sf_build_slots<uint16_t>(p));
  SfTables* d_T = nullptr;
  B2_TRY(b2_malloc(c, &d_T, 1));
  B2_TRY(b2_upload(c, d_T, &T, 1));
  p->sf_tab = d_T;
  return 0;
}

// Kronecker factors of the 8 child prolongators of the Galerkin plan; leaves p->sf_gal null if they are not products
static int sf_build_gal(b2_asm* p, const b2_galerkin_view& g) {
  b2_ctx* c = p->mesh->ctx;
  if (p->sf_gal) { b2_free(c, (SfGalTables*)p->sf_gal, 1); p->sf_gal = nullptr; }
  if (!p->sf_tab || g.nc != 27 || g.nf != 125 || p->mesh->nel != 8 * g.nelc) return 0;
  std::vector<int32_t> hdof(8 * 27), hfd(125);
  std::vector<double> hp((size_t)125 * 27), Pc((size_t)8 * 27 * 27);
  B2_TRY(b2_download(c, hdof.data(), p->dof, (size_t)8 * 27));
  B2_TRY(b2_download(c, hfd.data(), g.fd, (size_t)125));
  B2_TRY(b2_download(c, hp.data(), g.ploc, (size_t)125 * 27));
  for (int jn = 0; jn < 8 * 27; jn++) {        // child j, local node n -> its row of the parent's element prolongator
    int a = -1;
    for (int t = 0; t < 125; t++)
      if (hfd[t] == hdof[jn]) { a = t; break; }
    if (a < 0) return 0;
    for (int J = 0; J < 27; J++) Pc[(size_t)jn * 27 + J] = hp[(size_t)a * 27 + J];
  }
  SfGalTables G;
  if (!sf_factor_children(Pc.data(), &G)) return 0;
  SfGalTables* d_G = nullptr;
  B2_TRY(b2_malloc(c, &d_G, 1));
  B2_TRY(b2_upload(c, d_G, &G, 1));
  p->sf_gal = d_G;
  return 0;
}

template <int WARPS, typename SlotT, bool GAL, typename CSlotT>
int launch_assemble_sumfac_w(b2_asm* p, const SfGalArgs& ga, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_q2_sumfac_kernel<WARPS, SlotT, GAL, CSlotT>;
  const size_t smem = GAL ? SfSmem<WARPS>::bytes_gal : SfSmem<WARPS>::bytes;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nunits = GAL ? p->mesh->nel / 8 : p->mesh->nel;
  int grid = (int)((nunits + WARPS - 1) / WARPS);
  if (grid > c->sm_count) grid = c->sm_count;
  B2_LAUNCH(c, kern, grid, WARPS * 32, smem, p->mesh->nel, p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dofL, (const SfTables*)p->sf_tab,
            (const SlotT*)p->lslot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu, fsrc, ga);
  return 0;
}
template <typename SlotT, bool GAL, typename CSlotT>
int launch_assemble_sumfac(b2_asm* p, const SfGalArgs& ga, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  // 16 warps: only where the CTA's shared memory fits the SM (not with the fused Galerkin tile)
  if (p->mesh->ctx->asm_warps == 16 && (GAL ? SfSmem<16>::bytes_gal : SfSmem<16>::bytes) <= (size_t)227 * 1024)
    return launch_assemble_sumfac_w<16, SlotT, GAL, CSlotT>(p, ga, u, rhs, nu, fsrc);
  return launch_assemble_sumfac_w<kSfWarpsDefault, SlotT, GAL, CSlotT>(p, ga, u, rhs, nu, fsrc);
}

template <typename SlotT, bool GAL, typename CSlotT>
int launch_assemble_mma(b2_asm* p, const GalArgs& ga, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_q2_mma_kernel<SlotT, GAL, CSlotT>;
  const size_t smem = GAL ? MmaSmem::bytes_gal : MmaSmem::bytes;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((p->mesh->nel + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, p->mesh->nel, p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dof, p->tab,
            (const SlotT*)p->nslot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu, fsrc, ga);
  return 0;
}

template <int NVE, typename SlotT>
int launch_assemble(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_poisson_kernel<NVE, SlotT, false, uint8_t>;
  const size_t smem = SmemLayout<NVE>::bytes;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((p->mesh->nel + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  GalArgs ga = {};
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, p->mesh->nel, p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dof,
            p->tab, (const SlotT*)p->slot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu, fsrc, ga);
  return 0;
}

template <int NVE, typename SlotT, typename CSlotT>
int launch_assemble_gal(b2_asm* p, const b2_galerkin_view& g, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_poisson_kernel<NVE, SlotT, true, CSlotT>;
  const size_t smem = SmemLayout<NVE>::bytes_gal;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((p->mesh->nel + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  GalArgs ga = {p->gal_tab, g.cd, g.slot, g.fmask, g.cmask, g.Ac->rowptr, g.Ac->val, *g.emat};
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, p->mesh->nel, p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dof,
            p->tab, (const SlotT*)p->slot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu, fsrc, ga);
  return 0;
}

// Per-child prolongator tables of a Galerkin plan and the check that the fine elements are the
// children of the plan's coarse elements in the reference's order (children 8*iel + j,
// MeshRefinement.cpp:188-507; local nodes through fine2CoarseVertexMapping, Hexahedron.cpp:75-83).
template <int NVE>
int make_gal_tables(b2_ctx* c, const int32_t* d_dof, int64_t nel_fine, const b2_galerkin_view& g, GalTables<NVE>** out,
                    std::vector<double>* pc_dense = nullptr) {
  const int nf = g.nf, nc = g.nc;
  B2_CHECK(nc == NVE, "Galerkin from element matrices: coarse and fine unknowns must be of the same family");
  B2_CHECK(nel_fine == 8 * g.nelc, "Galerkin from element matrices: %lld fine elements are not 8 x %lld coarse elements",
           (long long)nel_fine, (long long)g.nelc);
  std::vector<int32_t> hdof(8 * NVE), hfd(nf);
  std::vector<double> hp((size_t)nf * nc);
  B2_TRY(b2_download(c, hdof.data(), d_dof, (size_t)8 * NVE));
  B2_TRY(b2_download(c, hfd.data(), g.fd, (size_t)nf));
  B2_TRY(b2_download(c, hp.data(), g.ploc, (size_t)nf * nc));
  std::vector<unsigned char> lat(8 * NVE);
  for (int jn = 0; jn < 8 * NVE; jn++) {
    int a = -1;
    for (int t = 0; t < nf; t++)
      if (hfd[t] == hdof[jn]) { a = t; break; }
    B2_CHECK(a >= 0, "fused Galerkin: fine element %d is not a child of coarse element 0", jn / NVE);
    lat[jn] = (unsigned char)a;
  }
  unsigned char* d_lat = nullptr;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_lat, lat.size()));
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_TRY(b2_upload(c, d_lat, lat.data(), lat.size()));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  B2_LAUNCH(c, gal_check_kernel, b2_grid_for(c, g.nelc * 8 * NVE, 256, 8), 256, 0, g.nelc, NVE, nf, d_dof, g.fd, d_lat, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_lat, lat.size());
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "fused Galerkin: the fine elements are not ordered as children 8*E+j of the coarse elements");
  if (pc_dense) {      // Pc[child][fine local node][coarse local node]
    pc_dense->assign((size_t)8 * NVE * NVE, 0.0);
    for (int jn = 0; jn < 8 * NVE; jn++)
      for (int J = 0; J < NVE; J++) (*pc_dense)[(size_t)jn * NVE + J] = hp[(size_t)lat[jn] * nc + J];
  }
  auto* T = new GalTables<NVE>();
  memset(T, 0, sizeof(*T));
  for (int j = 0; j < 8; j++) {
    int q = 0;
    for (int J = 0; J < NVE; J++) {
      T->colptr[j][J] = q;
      for (int n = 0; n < NVE; n++) {
        const double v = hp[(size_t)lat[j * NVE + n] * nc + J];
        if (v != 0.0) {
          if (q >= GalTables<NVE>::MAXNNZ) { delete T; B2_CHECK(false, "fused Galerkin: child prolongator too dense"); }
          T->crow[j][q] = (unsigned char)n;
          T->cval[j][q] = v;
          q++;
        }
      }
    }
    T->colptr[j][NVE] = q;
  }
  GalTables<NVE>* d_T = nullptr;
  int st = b2_malloc(c, &d_T, 1);
  if (!st) st = b2_upload(c, d_T, T, 1);
  delete T;
  B2_TRY(st);
  *out = d_T;
  return 0;
}

template <int NVE>
int build_gal_tables(b2_asm* p, const b2_galerkin* gal, const b2_galerkin_view& g) {
  b2_ctx* c = p->mesh->ctx;
  B2_CHECK(p->nve == NVE, "fused Galerkin: coarse and fine unknowns must be of the same family");
  B2_CHECK(g.Af == p->A, "fused Galerkin: the plan's fine matrix is not the assembled matrix");
  GalTables<NVE>* d_T = nullptr;
  B2_TRY(make_gal_tables<NVE>(c, p->dof, p->mesh->nel, g, &d_T));
  p->gal = gal;
  p->gal_tab = d_T;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Table-driven assembly for ANY Lagrange family elem_type_3D can build (tetrahedra 4/10/15, wedges
// 6/15/21, hexahedra with another quadrature rule ...): the kernel only sees nve <= 27 dofs per element,
// ng <= 64 Gauss points and the four tables of ElemType.cpp:637-740, exactly what
// elem_type_3D::Jacobian_type (ElemType.hpp:1438-1537) works from -- the per-element-type dispatch of the
// reference is the choice of tables.  One warp per element:
//   A. lanes = Gauss points: J, det, J^-1, weight (node-order sums as the reference's :1462-1472);
//   B. per Gauss point the nve physical gradients go to a double-buffered shared tile; lane l owns the
//      entries l, l+32, ... of the row-major nve x nve element matrix in registers;
//   C. element matrix to shared memory, residual F_i = fsrc sum_g phi_i w_g - nu (B u)_i by lane i,
//      fp64 atomicAdd scatter through the natural-order slot map (coalesced slot reads, no searches).
// CUDA cores only: these element matrices are 4x4 ... 21x21 with 5-45 Gauss points, too small and too
// ragged for the DMMA tiling of the Hex27 kernel.  HBM traffic per element: nve node ids + 3 nve
// coordinates + nve dofs + nve^2 slots in, nve^2 + nve atomics out.
constexpr int kGenWarps = 8;
constexpr int kGenMaxNg = 64;
constexpr int kGenAcc = (27 * 27 + 31) / 32;      // 23 entries per lane at most

struct GenSmem {
  // per warp: X[3][32], U[32], dofs[32] (as int), G[2][3][32], Geo[10][ng], B[nve*nve rounded up to 8]
  static int warp_doubles(int nve, int ng) { return 3 * 32 + 32 + 16 + 2 * 3 * 32 + 10 * ng + ((nve * nve + 7) & ~7); }
  static size_t bytes(int nve, int ng) { return (size_t)(4 * ng * nve + ng + kGenWarps * warp_doubles(nve, ng)) * sizeof(double); }
};

// NROWS = ceil(nve^2 / 32) register rows known at compile time (the loops over the entries unroll WITHOUT guards: with
// run-time guards, warp-uniform as they are, the compiler keeps every row in its own basic block and every row exposes
// its load -> multiply-add latency), or 0 = any nve <= 27 behind guards
template <typename SlotT, int NROWS>
__global__ void __launch_bounds__(kGenWarps * 32)
assemble_general_kernel(int64_t nel, int64_t nnode, int nve, int ng, const double* __restrict__ xyz,
                        const int32_t* __restrict__ conn, const int32_t* __restrict__ dof, const double* __restrict__ tab,
                        const SlotT* __restrict__ slot, const int64_t* __restrict__ rowptr, double* __restrict__ Aval,
                        const double* __restrict__ u, double* __restrict__ rhs, double nu, double fsrc) {
  extern __shared__ double smem[];
  const int tn = ng * nve;
  double* s_phi = smem;
  double* s_dx = s_phi + tn;
  double* s_dy = s_dx + tn;
  double* s_dz = s_dy + tn;
  double* s_w = s_dz + tn;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = s_w + ng + wib * (336 + 10 * ng + ((nve * nve + 7) & ~7));
  double* sX = wbase;                                  // [3][32]
  double* sU = sX + 96;                                // [32]
  int* sDof = reinterpret_cast<int*>(sU + 32);         // [32] ints in 16 doubles
  double* sG = sU + 32 + 16;                           // [2][3][32]
  double* sGeo = sG + 192;                             // [10][ng]
  double* sB = sGeo + 10 * ng;                         // [nve][nve]
  for (int t = threadIdx.x; t < 4 * tn + ng; t += blockDim.x) smem[t] = tab[t];
  __syncthreads();

  const int nn = nve * nve;
  constexpr int NACC = NROWS ? NROWS : kGenAcc;
  // (i, j) of the entries this lane owns, packed i * 32 + j
  int ij[NACC];
#pragma unroll
  for (int k = 0; k < NACC; k++) {
    const int e = lane + 32 * k;
    const int i = e < nn ? e / nve : 0;
    ij[k] = i * 32 + (e < nn ? e - i * nve : 0);
  }

  for (int64_t el = (int64_t)blockIdx.x * kGenWarps + wib; el < nel; el += (int64_t)gridDim.x * kGenWarps) {
    if (lane < nve) {
      const int64_t nd = conn[el * 27 + lane];
      sX[lane] = xyz[nd];
      sX[32 + lane] = xyz[nnode + nd];
      sX[64 + lane] = xyz[2 * nnode + nd];
      const int d = dof[el * nve + lane];
      sDof[lane] = d;
      sU[lane] = u ? u[d] : 0.0;
    }
    __syncwarp();

    // ---- A. geometry at the Gauss points owned by this lane
    for (int g = lane; g < ng; g += 32) {
      double J00 = 0, J01 = 0, J02 = 0, J10 = 0, J11 = 0, J12 = 0, J20 = 0, J21 = 0, J22 = 0;
      const double* dx = s_dx + g * nve;
      const double* dy = s_dy + g * nve;
      const double* dz = s_dz + g * nve;
      for (int n = 0; n < nve; n++) {
        const double x0 = sX[n], x1 = sX[32 + n], x2 = sX[64 + n];
        const double a = dx[n], b = dy[n], c = dz[n];
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
      sGeo[9 * ng + g] = det * s_w[g];
    }
    __syncwarp();

    // ---- B. element matrix, entries lane + 32 k in registers
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0.0;
    for (int g = 0; g < ng; g++) {
      double* G = sG + (g & 1) * 96;
      if (lane < nve) {
        const double a = s_dx[g * nve + lane], b = s_dy[g * nve + lane], c = s_dz[g * nve + lane];
        G[lane] = fma(c, sGeo[2 * ng + g], fma(b, sGeo[1 * ng + g], a * sGeo[0 * ng + g]));
        G[32 + lane] = fma(c, sGeo[5 * ng + g], fma(b, sGeo[4 * ng + g], a * sGeo[3 * ng + g]));
        G[64 + lane] = fma(c, sGeo[8 * ng + g], fma(b, sGeo[7 * ng + g], a * sGeo[6 * ng + g]));
      }
      __syncwarp();
      const double wg = sGeo[9 * ng + g];
#pragma unroll
      for (int k = 0; k < NACC; k++) {
        if (NROWS || 32 * k < nn) {      // generic instantiation only: skips the unused register rows of small elements
          const int i = ij[k] >> 5, j = ij[k] & 31;
          const double d = fma(G[64 + i], G[64 + j], fma(G[32 + i], G[32 + j], G[i] * G[j]));
          acc[k] = fma(d, wg, acc[k]);
        }
      }
    }

    // ---- C. element matrix to shared memory, residual, scatter
#pragma unroll
    for (int k = 0; k < NACC; k++) {
      const int e = lane + 32 * k;
      if (e < nn) sB[e] = acc[k];
    }
    __syncwarp();
    if (rhs && lane < nve) {
      double src = 0.0, bu = 0.0;
      for (int g = 0; g < ng; g++) src = fma(s_phi[g * nve + lane], sGeo[9 * ng + g], src);
      for (int j = 0; j < nve; j++) bu = fma(sB[lane * nve + j], sU[j], bu);
      atomicAdd(&rhs[sDof[lane]], fsrc * src - nu * bu);
    }
    const SlotT* sl = slot + (size_t)el * nn;
#pragma unroll
    for (int k = 0; k < NACC; k++) {
      const int e = lane + 32 * k;
      if (e < nn) atomicAdd(&Aval[rowptr[sDof[ij[k] >> 5]] + (int64_t)sl[e]], nu * acc[k]);
    }
    __syncwarp();      // the next element overwrites X, U, dofs, Geo, B
  }
}

template <typename SlotT, int NROWS>
int launch_assemble_general_n(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  b2_ctx* c = p->mesh->ctx;
  b2_prof_scope prof(c, p);
  auto kern = assemble_general_kernel<SlotT, NROWS>;
  const size_t smem = GenSmem::bytes(p->nve, p->ngauss);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;      // resident CTAs by registers and shared memory
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGenWarps * 32, smem));
  per_sm = per_sm < 1 ? 1 : per_sm;
  const int grid = b2_grid_for(c, p->mesh->nel, kGenWarps, per_sm);
  B2_LAUNCH(c, kern, grid, kGenWarps * 32, smem, p->mesh->nel, p->mesh->nnode, p->nve, p->ngauss, p->mesh->xyz, p->mesh->conn,
            p->dof, p->tab, (const SlotT*)p->nslot, p->A->rowptr, p->A->val, u ? u->d : nullptr, rhs ? rhs->d : nullptr, nu,
            fsrc);
  return 0;
}
// one instantiation per register-row count of the reference's 3-D Lagrange families (4 / 10 / 15 tetrahedra, 6 / 15 / 21
// wedges, 8 / 20 / 27 hexahedra: 1, 4, 8, 2, 8, 14, 2, 13, 23 rows), the guarded one for anything else
template <typename SlotT>
int launch_assemble_general(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  switch ((p->nve * p->nve + 31) / 32) {
    case 1: return launch_assemble_general_n<SlotT, 1>(p, u, rhs, nu, fsrc);
    case 2: return launch_assemble_general_n<SlotT, 2>(p, u, rhs, nu, fsrc);
    case 4: return launch_assemble_general_n<SlotT, 4>(p, u, rhs, nu, fsrc);
    case 8: return launch_assemble_general_n<SlotT, 8>(p, u, rhs, nu, fsrc);
    case 13: return launch_assemble_general_n<SlotT, 13>(p, u, rhs, nu, fsrc);
    case 14: return launch_assemble_general_n<SlotT, 14>(p, u, rhs, nu, fsrc);
    case 23: return launch_assemble_general_n<SlotT, 23>(p, u, rhs, nu, fsrc);
    default: return launch_assemble_general_n<SlotT, 0>(p, u, rhs, nu, fsrc);
  }
}
}  // namespace

#include "b2_neumann_kernel.cuh"

// fp64 tensor-core issue-rate probe: every warp keeps 8 independent DMMA.8x8x4 chains busy
__global__ void __launch_bounds__(512) dmma_probe_kernel(int iters, double* out) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; t++) c[t][0] = c[t][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 8; t++) dmma(c[t][0], c[t][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; t++) s += c[t][0] + c[t][1];
  if (s == -1.0) out[0] = s;      // never true: keeps the chains alive
}

// fp64 CUDA-core issue-rate probe: every thread keeps 8 independent DFMA chains busy
__global__ void __launch_bounds__(512) dfma_probe_kernel(int iters, double* out) {
  double c[8];
#pragma unroll
  for (int t = 0; t < 8; t++) c[t] = 1e-3 * t;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 8; t++) c[t] = fma(c[t], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; t++) s += c[t];
  if (s == -1.0) out[0] = s;      // never true: keeps the chains alive
}

extern "C" {

/* rhs += Neumann integrals over the listed boundary faces (host arrays: element, local face, flux value) of ONE
 * face kind: the face element is its tables phi, dxi, deta [ngf][nvf] and weights[ngf] (ngf <= 16 Gauss points,
 * nvf <= 9 dofs: quadrilaterals 4 / 8 / 9 with 16 points, triangles 3 / 6 / 7 with 13), face_nodes[6][9] the
 * element-local nodes of the element type's faces (rows of unused faces / entries: -1).  Face dofs must be
 * element dofs: every listed face's first nvf local nodes are < nve. */
int b2_asm_neumann_faces(b2_asm* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value,
                         int nvf, int ngf, const double* phi, const double* dxi, const double* deta, const double* weights,
                         const int32_t* face_nodes, b2_vec* rhs) {
  B2_CHECK(p && rhs && phi && dxi && deta && weights && face_nodes && (nfaces == 0 || (face_elem && face_local && face_value)),
           "b2_asm_neumann_faces: null argument");
  B2_CHECK(nvf >= 1 && nvf <= 9 && ngf >= 1 && ngf <= 16, "b2_asm_neumann_faces: nvf=%d (1..9) or ngf=%d (1..16) out of range", nvf, ngf);
  B2_CHECK(rhs->n >= p->A->nrows, "b2_asm_neumann_faces: rhs vector too short");
  if (nfaces == 0) return 0;
  b2_ctx* c = p->mesh->ctx;
  for (int64_t k = 0; k < nfaces; k++) {
    B2_CHECK(face_elem[k] >= 0 && face_elem[k] < p->mesh->nel && face_local[k] >= 0 && face_local[k] < 6,
             "b2_asm_neumann_faces: face %lld out of range", (long long)k);
    for (int i = 0; i < nvf; i++) {
      const int loc = face_nodes[face_local[k] * 9 + i];
      B2_CHECK(loc >= 0 && loc < p->nve, "b2_asm_neumann_faces: face %lld (local face %d): face dof %d is local node %d, not one of the element's %d dofs",
               (long long)k, (int)face_local[k], i, loc, p->nve);
    }
  }
  int32_t *d_e = nullptr, *d_f = nullptr, *d_fn = nullptr;
  double *d_v = nullptr, *d_t = nullptr;
  const size_t nt = (size_t)3 * ngf * nvf + ngf;
  std::vector<double> tab(nt);
  std::copy(phi, phi + ngf * nvf, tab.begin());
  std::copy(dxi, dxi + ngf * nvf, tab.begin() + ngf * nvf);
  std::copy(deta, deta + ngf * nvf, tab.begin() + 2 * ngf * nvf);
  std::copy(weights, weights + ngf, tab.begin() + 3 * ngf * nvf);
  B2_TRY(b2_malloc(c, &d_e, (size_t)nfaces));
  B2_TRY(b2_malloc(c, &d_f, (size_t)nfaces));
  B2_TRY(b2_malloc(c, &d_v, (size_t)nfaces));
  B2_TRY(b2_malloc(c, &d_t, nt));
  B2_TRY(b2_malloc(c, &d_fn, 54));
  B2_TRY(b2_upload(c, d_e, face_elem, (size_t)nfaces));
  B2_TRY(b2_upload(c, d_f, face_local, (size_t)nfaces));
  B2_TRY(b2_upload(c, d_v, face_value, (size_t)nfaces));
  B2_TRY(b2_upload(c, d_t, tab.data(), nt));
  B2_TRY(b2_upload(c, d_fn, face_nodes, 54));
  B2_LAUNCH(c, neumann_kernel, b2_grid_for(c, nfaces * 32, 256, 8), 256, 0, nfaces, d_e, d_f, d_v, nvf, ngf, p->nve, d_t, d_fn,
            p->mesh->nnode, p->mesh->xyz, p->mesh->conn, p->dof, rhs->d);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_e, (size_t)nfaces);
  b2_free(c, d_f, (size_t)nfaces);
  b2_free(c, d_v, (size_t)nfaces);
  b2_free(c, d_t, nt);
  b2_free(c, d_fn, 54);
  return 0;
}

/* The same for hexahedra with 8 or 27 dofs: ftab = phi, dxi, deta [16][nvf] and weights[16] of the quadrilateral
 * face element (nvf = 4 with nve = 8, 9 with 27), face_nodes[6][9] the local nodes of the hexahedron's faces. */
int b2_asm_neumann(b2_asm* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value,
                   int nvf, const double* phi, const double* dxi, const double* deta, const double* weights,
                   const int32_t* face_nodes, b2_vec* rhs) {
  B2_CHECK(p && rhs && (nfaces == 0 || (face_elem && face_local && face_value)), "b2_asm_neumann: null argument");
  B2_CHECK((nvf == 4 && p->nve == 8) || (nvf == 9 && p->nve == 27), "b2_asm_neumann: nvf=%d does not match nve=%d", nvf, p->nve);
  return b2_asm_neumann_faces(p, nfaces, face_elem, face_local, face_value, nvf, 16, phi, dxi, deta, weights, face_nodes, rhs);
}

/* measured fp64 tensor-core peak (TFLOP/s) of this device: the denominator next to the assembly
 * kernel's achieved rate in bench.py (MEASURED_PEAKS.json carries no fp64 figure) */
int b2_ctx_measure_fp64_tensor(b2_ctx* c, double* tflops) {
  double* d = nullptr;
  B2_TRY(b2_malloc(c, &d, 1));
  const int iters = 20000, grid = c->sm_count * 2;
  dmma_probe_kernel<<<grid, 512, 0, c->stream>>>(2000, d);       // warm-up
  cudaEvent_t e0, e1;
  B2_CUDA(cudaEventCreate(&e0));
  B2_CUDA(cudaEventCreate(&e1));
  B2_CUDA(cudaEventRecord(e0, c->stream));
  dmma_probe_kernel<<<grid, 512, 0, c->stream>>>(iters, d);
  c->launches += 2;
  B2_CUDA(cudaEventRecord(e1, c->stream));
  B2_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  b2_free(c, d, 1);
  const double flop = (double)grid * 16.0 * iters * 8.0 * 512.0;     // warps x iterations x chains x (8x8x4 FMA = 512 flop)
  *tflops = flop / (ms * 1e-3) / 1e12;
  return 0;
}

/* measured issue-rate peak of DFMA on this device, TFLOP/s (2 flop per lane and instruction) */
int b2_ctx_measure_fp64_fma(b2_ctx* c, double* tflops) {
  B2_CHECK(c && tflops, "b2_ctx_measure_fp64_fma: null argument");
  double* d = nullptr;
  B2_TRY(b2_malloc(c, &d, 1));
  const int iters = 20000, grid = c->sm_count * 2;
  dfma_probe_kernel<<<grid, 512, 0, c->stream>>>(2000, d);       // warm-up
  cudaEvent_t e0, e1;
  B2_CUDA(cudaEventCreate(&e0));
  B2_CUDA(cudaEventCreate(&e1));
  B2_CUDA(cudaEventRecord(e0, c->stream));
  dfma_probe_kernel<<<grid, 512, 0, c->stream>>>(iters, d);
  c->launches += 2;
  B2_CUDA(cudaEventRecord(e1, c->stream));
  B2_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  B2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  b2_free(c, d, 1);
  const double flop = (double)grid * 512.0 * iters * 8.0 * 2.0;     // threads x iterations x chains x 2
  *tflops = flop / (ms * 1e-3) / 1e12;
  return 0;
}

int b2_mesh_create(b2_ctx* c, int64_t nnode, int64_t nel, const double* xyz, const int32_t* conn, b2_mesh** out) {
  *out = nullptr;
  B2_CHECK(c && nnode > 0 && nel > 0 && xyz && conn, "b2_mesh_create: bad arguments");
  b2_mesh* m = new b2_mesh{c, nnode, nel, nullptr, nullptr, nullptr, nullptr};
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
// upload into the shadow buffers on the copy stream (call b2_ctx_open_copies first), then
// b2_ctx_join_copies + b2_mesh_swap before the step that uses the new mesh
int b2_mesh_prefetch(b2_mesh* m, const double* xyz, const int32_t* conn) {
  b2_ctx* c = m->ctx;
  if (!m->xyz_shadow) {
    B2_TRY(b2_malloc(c, &m->xyz_shadow, (size_t)3 * m->nnode));
    B2_TRY(b2_malloc(c, &m->conn_shadow, (size_t)27 * m->nel));
    B2_CUDA(cudaMemcpyAsync(m->xyz_shadow, m->xyz, (size_t)3 * m->nnode * sizeof(double), cudaMemcpyDeviceToDevice, c->copy_stream));
    B2_CUDA(cudaMemcpyAsync(m->conn_shadow, m->conn, (size_t)27 * m->nel * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->copy_stream));
  }
  if (xyz) B2_CUDA(cudaMemcpyAsync(m->xyz_shadow, xyz, (size_t)3 * m->nnode * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
  if (conn) B2_CUDA(cudaMemcpyAsync(m->conn_shadow, conn, (size_t)27 * m->nel * sizeof(int32_t), cudaMemcpyHostToDevice, c->copy_stream));
  return 0;
}
int b2_mesh_swap(b2_mesh* m) {
  B2_CHECK(m->xyz_shadow, "b2_mesh_swap: nothing was prefetched");
  std::swap(m->xyz, m->xyz_shadow);
  std::swap(m->conn, m->conn_shadow);
  return 0;
}
int b2_mesh_destroy(b2_mesh* m) {
  if (!m) return 0;
  cudaStreamSynchronize(m->ctx->stream);
  cudaStreamSynchronize(m->ctx->copy_stream);
  if (m->xyz_shadow) b2_free(m->ctx, m->xyz_shadow, (size_t)3 * m->nnode);
  if (m->conn_shadow) b2_free(m->ctx, m->conn_shadow, (size_t)27 * m->nel);
  b2_free(m->ctx, m->xyz, (size_t)3 * m->nnode);
  b2_free(m->ctx, m->conn, (size_t)27 * m->nel);
  delete m;
  return 0;
}

int b2_asm_create(b2_mesh* m, b2_csr* A, int nve, const int32_t* dof, int ngauss, const double* phi,
                  const double* dxi, const double* deta, const double* dzeta, const double* weights, b2_asm** out) {
  *out = nullptr;
  B2_CHECK(m && A && dof && phi && dxi && deta && dzeta && weights, "b2_asm_create: null argument");
  // hexahedra with the 64-point rule run the specialised kernels; every other family / rule of
  // elem_type_3D (tetrahedra, wedges, other quadrature orders) runs the table-driven kernel
  const bool general = !((nve == 8 || nve == 27) && ngauss == NG);
  B2_CHECK(nve >= 1 && nve <= 27, "b2_asm_create: nve=%d (1..27 dofs per element)", nve);
  B2_CHECK(ngauss >= 1 && ngauss <= kGenMaxNg, "b2_asm_create: ngauss=%d (1..%d Gauss points)", ngauss, kGenMaxNg);
  B2_CHECK(GenSmem::bytes(nve, ngauss) <= (size_t)227 * 1024, "b2_asm_create: tables of %d x %d do not fit shared memory", ngauss, nve);
  B2_CHECK(A->max_row <= 65536, "b2_asm_create: rows longer than 65536 entries");
  b2_ctx* c = m->ctx;
  b2_asm* p = new b2_asm();
  p->mesh = m;
  p->A = A;
  p->ctx = c;
  p->nel = m->nel;
  p->nve = nve;
  p->ngauss = ngauss;
  p->last_ms = 0.;
  p->gal = nullptr;
  p->gal_tab = nullptr;
  p->sf_tab = nullptr;
  p->dofL = nullptr;
  p->lslot = nullptr;
  p->sf_gal = nullptr;
  p->general = general;
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
  p->nslot = nullptr;
  p->nslot_count = 0;
  p->slot = nullptr;
  p->slot_count = 0;
  if (nve == 27 || general) {     // natural-order map (tensor-core and table-driven kernels); the tile-order map of the CUDA-core kernel is built on demand
    if (p->slot_bytes == 1) B2_TRY(build_natural_slots<uint8_t>(p));
    else B2_TRY(build_natural_slots<uint16_t>(p));
  } else {
    if (p->slot_bytes == 1) B2_TRY((build_slots<8, uint8_t>(p)));
    else B2_TRY((build_slots<8, uint16_t>(p)));
  }
  if (nve == 27 && !general) B2_TRY(sf_prepare(p, dof, phi, dxi, deta, dzeta, weights));
  *out = p;
  return 0;
}

int b2_asm_destroy(b2_asm* p) {
  if (!p) return 0;
  b2_ctx* c = p->ctx;      // never through the borrowed mesh / matrix (they may have been destroyed first)
  cudaStreamSynchronize(c->stream);
  b2_free(c, p->dof, (size_t)p->nel * p->nve);
  b2_free(c, p->tab, (size_t)4 * p->ngauss * p->nve + p->ngauss);
  if (p->slot) {
    if (p->slot_bytes == 1) b2_free(c, (uint8_t*)p->slot, p->slot_count);
    else b2_free(c, (uint16_t*)p->slot, p->slot_count);
  }
  if (p->nslot) {
    if (p->slot_bytes == 1) b2_free(c, (uint8_t*)p->nslot, p->nslot_count);
    else b2_free(c, (uint16_t*)p->nslot, p->nslot_count);
  }
  if (p->gal_tab) {
    if (p->nve == 27) b2_free(c, (GalTables<27>*)p->gal_tab, 1);
    else b2_free(c, (GalTables<8>*)p->gal_tab, 1);
  }
  b2_free(c, (SfTables*)p->sf_tab, 1);
  b2_free(c, p->dofL, (size_t)p->nel * 27);
  if (p->lslot) {
    if (p->slot_bytes == 1) b2_free(c, (uint8_t*)p->lslot, (size_t)p->nel * kSfSlotStride);
    else b2_free(c, (uint16_t*)p->lslot, (size_t)p->nel * kSfSlotStride);
  }
  b2_free(c, (SfGalTables*)p->sf_gal, 1);
  delete p;
  return 0;
}

int b2_asm_poisson(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  B2_CHECK(!u || u->n >= p->A->nrows, "b2_asm_poisson: solution vector too short");
  B2_CHECK(!rhs || rhs->n >= p->A->nrows, "b2_asm_poisson: rhs vector too short");
  if (p->general || p->mesh->ctx->asm_variant == 2) {       // table-driven kernel (any family)
    p->last_kernel = "assemble_general_kernel";
    if (!p->nslot) {
      if (p->slot_bytes == 1) B2_TRY(build_natural_slots<uint8_t>(p));
      else B2_TRY(build_natural_slots<uint16_t>(p));
    }
    if (p->slot_bytes == 1) return launch_assemble_general<uint8_t>(p, u, rhs, nu, fsrc);
    return launch_assemble_general<uint16_t>(p, u, rhs, nu, fsrc);
  }
  if (p->nve == 27 && p->mesh->ctx->asm_variant == 3 && p->sf_tab) {      // sum-factorised kernel (default)
    SfGalArgs ga = {};
    p->last_kernel = "assemble_q2_sumfac_kernel";
    if (p->slot_bytes == 1) return launch_assemble_sumfac<uint8_t, false, uint8_t>(p, ga, u, rhs, nu, fsrc);
    return launch_assemble_sumfac<uint16_t, false, uint8_t>(p, ga, u, rhs, nu, fsrc);
  }
  if (p->nve == 27 && (p->mesh->ctx->asm_variant == 1 || p->mesh->ctx->asm_variant == 3)) {      // FP64 tensor-core kernel
    GalArgs ga = {};
    p->last_kernel = "assemble_q2_mma_kernel";
    if (p->slot_bytes == 1) return launch_assemble_mma<uint8_t, false, uint8_t>(p, ga, u, rhs, nu, fsrc);
    return launch_assemble_mma<uint16_t, false, uint8_t>(p, ga, u, rhs, nu, fsrc);
  }
  p->last_kernel = "assemble_poisson_kernel";
  if (p->nve == 27) {
    if (!p->slot) {      // first use of the CUDA-core kernel on this plan
      if (p->slot_bytes == 1) B2_TRY((build_slots<27, uint8_t>(p)));
      else B2_TRY((build_slots<27, uint16_t>(p)));
    }
    if (p->slot_bytes == 1) return launch_assemble<27, uint8_t>(p, u, rhs, nu, fsrc);
    return launch_assemble<27, uint16_t>(p, u, rhs, nu, fsrc);
  }
  if (p->slot_bytes == 1) return launch_assemble<8, uint8_t>(p, u, rhs, nu, fsrc);
  return launch_assemble<8, uint16_t>(p, u, rhs, nu, fsrc);
}
int b2_asm_poisson_galerkin(b2_asm* p, b2_galerkin* gal, const b2_vec* u, b2_vec* rhs, double nu, double fsrc) {
  B2_CHECK(p && gal, "b2_asm_poisson_galerkin: null argument");
  B2_CHECK(!u || u->n >= p->A->nrows, "b2_asm_poisson_galerkin: solution vector too short");
  B2_CHECK(!rhs || rhs->n >= p->A->nrows, "b2_asm_poisson_galerkin: rhs vector too short");
  B2_CHECK(!p->general, "b2_asm_poisson_galerkin: the fused Galerkin product is for refined hexahedra (8 or 27 dofs, 64 Gauss points)");
  b2_ctx* c = p->mesh->ctx;
  b2_galerkin_view g;
  B2_TRY(b2_galerkin_get_view(gal, &g));
  if (p->gal != gal) {
    if (p->gal_tab) {
      if (p->nve == 27) b2_free(c, (GalTables<27>*)p->gal_tab, 1);
      else b2_free(c, (GalTables<8>*)p->gal_tab, 1);
      p->gal_tab = nullptr;
      p->gal = nullptr;
    }
    if (p->nve == 27) B2_TRY(build_gal_tables<27>(p, gal, g));
    else B2_TRY(build_gal_tables<8>(p, gal, g));
    if (p->nve == 27) B2_TRY(sf_build_gal(p, g));
  }
  B2_CUDA(cudaMemsetAsync(g.Ac->val, 0, (size_t)g.Ac->nnz * sizeof(double), c->stream));
  const bool s1 = p->slot_bytes == 1, c1 = g.slot_bytes == 1;
  if (p->nve == 27 && c->asm_variant == 3 && p->sf_tab && p->sf_gal) {      // sum-factorised kernel (default)
    SfGalArgs ga = {(const SfGalTables*)p->sf_gal, g.cd, g.slot, g.fmask, g.cmask, g.Ac->rowptr, g.Ac->val, *g.emat};
    p->last_kernel = "assemble_q2_sumfac_kernel<fused Galerkin>";
    if (s1 && c1) return launch_assemble_sumfac<uint8_t, true, uint8_t>(p, ga, u, rhs, nu, fsrc);
    if (s1) return launch_assemble_sumfac<uint8_t, true, uint16_t>(p, ga, u, rhs, nu, fsrc);
    if (c1) return launch_assemble_sumfac<uint16_t, true, uint8_t>(p, ga, u, rhs, nu, fsrc);
    return launch_assemble_sumfac<uint16_t, true, uint16_t>(p, ga, u, rhs, nu, fsrc);
  }
  if (p->nve == 27 && (c->asm_variant == 1 || c->asm_variant == 3)) {      // FP64 tensor-core kernel
    GalArgs ga = {p->gal_tab, g.cd, g.slot, g.fmask, g.cmask, g.Ac->rowptr, g.Ac->val, *g.emat};
    p->last_kernel = "assemble_q2_mma_kernel<fused Galerkin>";
    if (s1 && c1) return launch_assemble_mma<uint8_t, true, uint8_t>(p, ga, u, rhs, nu, fsrc);
    if (s1) return launch_assemble_mma<uint8_t, true, uint16_t>(p, ga, u, rhs, nu, fsrc);
    if (c1) return launch_assemble_mma<uint16_t, true, uint8_t>(p, ga, u, rhs, nu, fsrc);
    return launch_assemble_mma<uint16_t, true, uint16_t>(p, ga, u, rhs, nu, fsrc);
  }
  p->last_kernel = "assemble_poisson_galerkin_kernel";
  if (p->nve == 27) {
    if (!p->slot) {
      if (s1) B2_TRY((build_slots<27, uint8_t>(p)));
      else B2_TRY((build_slots<27, uint16_t>(p)));
    }
    if (s1 && c1) return launch_assemble_gal<27, uint8_t, uint8_t>(p, g, u, rhs, nu, fsrc);
    if (s1) return launch_assemble_gal<27, uint8_t, uint16_t>(p, g, u, rhs, nu, fsrc);
    if (c1) return launch_assemble_gal<27, uint16_t, uint8_t>(p, g, u, rhs, nu, fsrc);
    return launch_assemble_gal<27, uint16_t, uint16_t>(p, g, u, rhs, nu, fsrc);
  }
  if (s1 && c1) return launch_assemble_gal<8, uint8_t, uint8_t>(p, g, u, rhs, nu, fsrc);
  if (s1) return launch_assemble_gal<8, uint8_t, uint16_t>(p, g, u, rhs, nu, fsrc);
  if (c1) return launch_assemble_gal<8, uint16_t, uint8_t>(p, g, u, rhs, nu, fsrc);
  return launch_assemble_gal<8, uint16_t, uint16_t>(p, g, u, rhs, nu, fsrc);
}
/* name of the kernel the last b2_asm_poisson / b2_asm_poisson_galerkin call on this plan launched (bench.py labels
 * its roofline with it) */
const char* b2_asm_kernel_name(const b2_asm* p) { return p ? p->last_kernel : ""; }
/* record the Galerkin element matrices of gal's coarse elements whenever gal is applied by the fused
 * assembly or from element matrices (needed by b2_galerkin_apply_from_elements of the next plan) */
int b2_galerkin_record_elements(b2_galerkin* gal, int on) {
  b2_galerkin_view g;
  B2_TRY(b2_galerkin_get_view(gal, &g));
  b2_ctx* c = g.Af->ctx;
  const size_t n = (size_t)g.nelc * g.nc * g.nc;
  if (on && !*g.emat) B2_TRY(b2_malloc(c, g.emat, n));
  if (!on && *g.emat) { cudaStreamSynchronize(c->stream); b2_free(c, *g.emat, n); *g.emat = nullptr; }
  return 0;
}

}  // extern "C"

template <int NVE, typename CSlotT>
static int launch_from_elements(b2_ctx* c, const b2_galerkin_view& g, const b2_galerkin_view& f) {
  auto kern = galerkin_from_elements_kernel<NVE, CSlotT>;
  const size_t smem = ((size_t)kWarps * (NVE * NVE + 3)) * sizeof(double) + sizeof(GalTables<NVE>);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nel = f.nelc;
  int grid = (int)((nel + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  GalArgs ga = {*g.chain_tab, g.cd, g.slot, g.fmask, g.cmask, g.Ac->rowptr, g.Ac->val, *g.emat};
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, nel, (const double*)*f.emat, f.cd, ga);
  return 0;
}

// the same product through the Kronecker factors of the child prolongators (galerkin_chain_sumfac_kernel)
template <typename CSlotT>
static int launch_from_elements_sumfac(b2_ctx* c, const b2_galerkin_view& g, const b2_galerkin_view& f) {
  constexpr int WARPS = kSfChainWarps;
  auto kern = galerkin_chain_sumfac_kernel<WARPS, CSlotT>;
  const size_t smem = SfChainSmem<WARPS>::bytes;
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((g.nelc + WARPS - 1) / WARPS);
  if (grid > c->sm_count) grid = c->sm_count;
  SfGalArgs ga = {(const SfGalTables*)*g.chain_sf, g.cd, g.slot, g.fmask, g.cmask, g.Ac->rowptr, g.Ac->val, *g.emat};
  B2_LAUNCH(c, kern, grid, WARPS * 32, smem, g.nelc, (const double*)*f.emat, f.cd, ga);
  return 0;
}

extern "C" {

/* gal: Ac = P^T Af P formed from the element matrices recorded by `finer` (the plan whose coarse
 * matrix is gal's fine matrix): A_{l-1} = sum_E Pc(E)^T D_E Pc(E).  Equal to b2_galerkin_apply(gal)
 * up to summation order, without reading Af. */
int b2_galerkin_apply_from_elements(b2_galerkin* gal, b2_galerkin* finer) {
  B2_CHECK(gal && finer, "b2_galerkin_apply_from_elements: null argument");
  b2_galerkin_view g, f;
  B2_TRY(b2_galerkin_get_view(gal, &g));
  B2_TRY(b2_galerkin_get_view(finer, &f));
  b2_ctx* c = g.Af->ctx;
  B2_CHECK(f.Ac == g.Af, "b2_galerkin_apply_from_elements: the finer plan's coarse matrix is not this plan's fine matrix");
  B2_CHECK(*f.emat, "b2_galerkin_apply_from_elements: the finer plan has no recorded element matrices "
                    "(b2_galerkin_record_elements, then apply it through the fused assembly or from elements)");
  B2_CHECK(g.nc == f.nc && (g.nc == 27 || g.nc == 8), "b2_galerkin_apply_from_elements: unsupported element family");
  if (!*g.chain_tab) {
    if (g.nc == 27) {
      GalTables<27>* t = nullptr;
      std::vector<double> Pc;
      B2_TRY(make_gal_tables<27>(c, f.cd, f.nelc, g, &t, &Pc));
      *g.chain_tab = t;
      // Kronecker factors of the child prolongators (checked): the chain then costs 13 x fewer multiply-adds
      SfGalTables G;
      *g.chain_sf_tried = 1;
      if (g.nf == 125 && sf_factor_children(Pc.data(), &G)) {
        SfGalTables* d_G = nullptr;
        B2_TRY(b2_malloc(c, &d_G, 1));
        B2_TRY(b2_upload(c, d_G, &G, 1));
        *g.chain_sf = d_G;
      }
    }
    else { GalTables<8>* t = nullptr; B2_TRY(make_gal_tables<8>(c, f.cd, f.nelc, g, &t)); *g.chain_tab = t; }
    *g.chain_tab_nve = g.nc;
  }
  B2_CUDA(cudaMemsetAsync(g.Ac->val, 0, (size_t)g.Ac->nnz * sizeof(double), c->stream));
  b2_prof_scope prof(c, gal);
  if (g.nc == 27 && *g.chain_sf && c->asm_variant == 3) {
    if (g.slot_bytes == 1) return launch_from_elements_sumfac<uint8_t>(c, g, f);
    return launch_from_elements_sumfac<uint16_t>(c, g, f);
  }
  if (g.nc == 27) {
    if (g.slot_bytes == 1) return launch_from_elements<27, uint8_t>(c, g, f);
    return launch_from_elements<27, uint16_t>(c, g, f);
  }
  if (g.slot_bytes == 1) return launch_from_elements<8, uint8_t>(c, g, f);
  return launch_from_elements<8, uint16_t>(c, g, f);
}
double b2_asm_last_kernel_ms(const b2_asm* p) { return p->last_ms; }

}  // extern "C"
