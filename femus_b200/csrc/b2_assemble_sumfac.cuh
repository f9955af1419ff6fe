// Triquadratic hexahedra (27 dofs, 4 x 4 x 4 Gauss points) by SUM FACTORISATION -- the default assembly kernel of the
// Hex27 path.  Same arithmetic contract as the tensor-core kernel of b2_assemble.cu (the element loop of
// applications/001_Poisson/main.cpp:350-602 with elem_type_3D::Jacobian_type, ElemType.hpp:1438-1537, the blocked
// scatter of PetscMatrix.cpp:699-729 and, fused, the first Galerkin product of LinearImplicitSystem.cpp:347-370), but
// it uses what the reference's tables ARE: phi_n(g) = l_i1(p_a) l_i2(p_b) l_i3(p_c) for node n at lattice position
// (i1, i2, i3) and Gauss point g = 16 a + 4 b + c (Hexahedron.cpp:95-163, quadrature_Hexahedron.cpp), so every
// contraction over nodes or Gauss points splits into three 1-D contractions:
//   A. J(g) = sum_n dphi_n(g) x_n          3 x 99 FMA per lane instead of 486 (lanes = Gauss points (b, c), two a each)
//   B. B_ij = sum_{alpha beta} sum_g K_ab(g) d_alpha phi_i(g) d_beta phi_j(g)  with K = w det J^-1 J^-T:
//        S1[i3 j3][a][b] = sum_c M3[i3 j3][c] K_ab[a][b][c]     lanes = Gauss points (a, b): every K value is read ONCE,
//                                                              the coefficients are constant-bank operands; S1 goes
//                                                              through shared memory (4 terms per round, two rounds)
//        T [j2][a]  += sum_b M2[i2 j2][b] S1[i3 j3][a][b]       (accumulated over the (alpha, beta) of one class)
//        B[i1 j1 ..] += sum_a M1[i1 j1][a] T[j2][a]             M_d[ij][.] = u_i u_j, u = l or l' as alpha / beta = d
//      lane = (i2, i3, j3) owns the 27 entries (i1, j1, j2): ~1300 FMA per lane and element -- 41 k FMA per element
//      against 123 k (480 DMMA) + 15 k of the tensor-core formulation; on B200 the FP64 tensor pipe has the SAME peak
//      as the FP64 CUDA-core pipe (37 TFLOP/s measured), so the flop count, not the pipe, is what matters.  (First
//      version: every lane contracted all 64 points itself, 1440 FMA -- and 9 x 64 broadcast K loads per lane; a
//      shared-memory load costs register-writeback bandwidth per LANE, broadcast or not: ncu showed the l1tex data
//      pipe at 88 %, 1150 of 1480 wavefronts per element from those loads.)
//   D. fused Galerkin: the child's element prolongator is a Kronecker product A1 x A2 x A3 of 3 x 3 matrices, so
//      Pc^T B Pc is six passes of 3-vectors through a 3 x 3 matrix, all in registers (lane = row, then column).
// One warp takes the 8 children of a coarse element one after the other and sums their Galerkin contributions in
// its own shared-memory tile: no barrier between warps, and the coarse element matrix is summed in a fixed order.
// The tables handed to b2_asm_create are CHECKED to have this structure (b2_assemble.cu: sf_prepare); if they do
// not (another rule, another node order) the plan stays on the tensor-core kernel.
//
// Kept free of host / runtime calls and of inline PTX so that the same source runs on the CPU thread emulator
// (tests/cpp/cuda_emu.hpp, tests/test_kernel_emulation.py).  Included inside an anonymous namespace.
#pragma once
#ifndef B2_DYN_SHARED
#define B2_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

// WARPS = elements in flight per CTA (one CTA per SM): 12 (168 registers per thread) or 16 (128, a few spills);
// both are built, "asm_warps" of b2_ctx_set_option / B2_SF_WARPS picks one
constexpr int kSfWarpsDefault = 12;
constexpr int kSfNG = 64, kSfNVE = 27;
// slot map of one element: 729 entries padded to 736 so that every element starts on a 16-byte boundary (vector loads)
constexpr int kSfSlotStride = 736;

struct SfTables {                      // built by sf_prepare from the caller's phi / dphi tables
  double L[3][4];                      // l_i(p_a)
  double D[3][4];                      // l_i'(p_a)
  double M[4][9][4];                   // M[2p+q][3i+j][a] = u^p_i(p_a) u^q_j(p_a), u^0 = l, u^1 = l'
  double w[64];                        // Gauss weights (the reference's 64 truncated literals, NOT a tensor product)
  int node_of[32];                     // lattice position m = 9 i1 + 3 i2 + i3 -> local node of the reference's order
};
struct SfGalTables {
  // 1-D factors of the child prolongators.  Refining a quadratic 1-D element, a child's three nodes take
  //   lower child: (v0, a v0 + b v1 + c v2, v1)      upper child: (v1, c v0 + b v1 + a v2, v2)
  // from the parent's (v0, v1, v2) -- (a, b, c) = (3/8, 3/4, -1/8) for Lagrange polynomials; sf_factor_children
  // checks that every factor has one of these two forms, so a pass costs 3 instead of 9 multiply-adds per triple
  double abc[4];                       // a, b, c, (unused)
  int hi[8][4];                        // [child][dim d]: 1 = upper child in that dimension
  unsigned short nat2lat[736];         // natural entry I * 27 + J -> lattice entry m(I) * 27 + m(J)
  double A[8][3][3][3];                // [child][dim d][fine lattice index n_d][coarse lattice index J_d] (host-side check only)
};
struct SfGalArgs {
  const SfGalTables* tab;              // device
  const int32_t* cd;                   // [nelc][27] coarse dofs (natural order)
  const void* cslot;                   // [nelc][729] slot of (I, J) inside row cd_I of the coarse matrix (natural order)
  const uint8_t* fmask;                // fine Dirichlet rows of P (may be null)
  const uint8_t* cmask;                // coarse Dirichlet columns of P (may be null)
  const int64_t* Cp;                   // coarse rowptr
  double* Cv;                          // coarse values
  double* emat;                        // [nelc][729] record of every coarse element's Galerkin matrix (natural order), or null
};

// coefficients that are uniform over the warp and indexed by compile-time constants -> constant-bank operands:
// c_sfM (stages 1 and 3), c_sfU[p][j][b] = u^p_j(p_b) (stage 2: M2[i2 j2][b] = u^p_i2(b) u^q_j2(b), the i2 factor is a
// per-lane register constant)
__constant__ double c_sfM[4][9][4];
__constant__ double c_sfU[2][3][4];

template <int WARPS>
struct SfSmem {
  // tables: L, D (12 each), M (144), w (64), A (216, fused Galerkin only) | per warp: X[3][28], U[28], row starts[28],
  // KB = K[7][64] + S1[4][9][18] during the quadrature loop, then the element matrix [27][27]; the element's slot map
  // [736]; Dacc [27][27] (fused Galerkin only)
  static constexpr int tab_doubles = 12 + 12 + 144 + 64;
  static constexpr int gal_doubles = 4 + 16 + 736 / 4;      // abc, hi flags, nat2lat
  static constexpr int s1_doubles = 4 * 9 * 18;        // S1 of the four terms of a round: [term][i3 j3][a b], rows padded to 18
                                                       // (stride 16 puts the 9 rows a quarter-warp reads on the same banks)
  static constexpr int kb_doubles = 7 * 64 + s1_doubles;   // K, S1; later the element matrix (729)
  static constexpr int dacc_doubles = 736;
  static constexpr int slot_doubles = kSfSlotStride * 2 / 8;      // this element's slot map, staged early (16-bit entries at most)
  static constexpr int warp_doubles = 3 * 28 + 28 + 28 + kb_doubles + slot_doubles;
  static constexpr int warp_doubles_gal = warp_doubles + dacc_doubles + 28;      // + row starts of the coarse dofs
  static constexpr size_t bytes = (size_t)(tab_doubles + WARPS * warp_doubles) * sizeof(double);
  static constexpr size_t bytes_gal = (size_t)(tab_doubles + gal_doubles + WARPS * warp_doubles_gal) * sizeof(double);
};

// stage 1 of one term, lanes 0..15 = Gauss points (a, b): S1[i3 j3][a b] = sum_c M3[i3 j3][c] K[a][b][c]
template <int PQ3>
__device__ __forceinline__ void sf_stage1(const double* __restrict__ Kc, double* __restrict__ S1, int ab) {
  const double2 k01 = *reinterpret_cast<const double2*>(Kc + ab * 4);
  const double2 k23 = *reinterpret_cast<const double2*>(Kc + ab * 4 + 2);
#pragma unroll
  for (int ij = 0; ij < 9; ij++)
    S1[ij * 18 + ab] = fma(c_sfM[PQ3][ij][3], k23.y, fma(c_sfM[PQ3][ij][2], k23.x, fma(c_sfM[PQ3][ij][1], k01.y, c_sfM[PQ3][ij][0] * k01.x)));
}
// stage 2 of one term, lane = (i2, i3, j3): T[j2][a] += sum_b u^P2_i2(b) u^Q2_j2(b) S1[i3 j3][a][b]
template <int P2, int Q2>
__device__ __forceinline__ void sf_stage2(const double* __restrict__ S1, const double (&ui2)[2][4], double (&T)[3][4]) {
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const double2 s01 = *reinterpret_cast<const double2*>(S1 + a * 4);
    const double2 s23 = *reinterpret_cast<const double2*>(S1 + a * 4 + 2);
    const double s0 = ui2[P2][0] * s01.x, s1 = ui2[P2][1] * s01.y, s2 = ui2[P2][2] * s23.x, s3 = ui2[P2][3] * s23.y;
#pragma unroll
    for (int j2 = 0; j2 < 3; j2++)
      T[j2][a] = fma(c_sfU[Q2][j2][3], s3, fma(c_sfU[Q2][j2][2], s2, fma(c_sfU[Q2][j2][1], s1, fma(c_sfU[Q2][j2][0], s0, T[j2][a]))));
  }
}
// stage 3 of one class: out[(3 i1 + j1) * 3 + j2] += sum_a M1[i1 j1][a] T[j2][a]
template <int PQ1>
__device__ __forceinline__ void sf_stage3(const double (&T)[3][4], double (&out)[27]) {
#pragma unroll
  for (int ij = 0; ij < 9; ij++)
#pragma unroll
    for (int j2 = 0; j2 < 3; j2++)
#pragma unroll
      for (int a = 0; a < 4; a++) out[ij * 3 + j2] = fma(c_sfM[PQ1][ij][a], T[j2][a], out[ij * 3 + j2]);
}
__device__ __forceinline__ void sf_zero(double (&T)[3][4]) {
#pragma unroll
  for (int j2 = 0; j2 < 3; j2++)
#pragma unroll
    for (int a = 0; a < 4; a++) T[j2][a] = 0.0;
}
// one 1-D pass of the Kronecker product on the triples (v0, v1, v2) at stride S: v -> v A with A the transposed
// lower / upper child factor (see SfGalTables)
template <int S>
__device__ __forceinline__ void sf_kron_pass(double (&R)[27], bool hi, double a, double b, double c) {
  if (hi) {
#pragma unroll
    for (int h = 0; h < 27; h += 3 * S)
#pragma unroll
      for (int lo = 0; lo < S; lo++) {
        const double r0 = R[h + lo], r1 = R[h + lo + S], r2 = R[h + lo + 2 * S];
        R[h + lo] = c * r1;
        R[h + lo + S] = fma(b, r1, r0);
        R[h + lo + 2 * S] = fma(a, r1, r2);
      }
  } else {
#pragma unroll
    for (int h = 0; h < 27; h += 3 * S)
#pragma unroll
      for (int lo = 0; lo < S; lo++) {
        const double r0 = R[h + lo], r1 = R[h + lo + S], r2 = R[h + lo + 2 * S];
        R[h + lo] = fma(a, r1, r0);
        R[h + lo + S] = fma(b, r1, r2);
        R[h + lo + 2 * S] = c * r1;
      }
  }
}

template <int WARPS, typename SlotT, bool GAL, typename CSlotT>
__global__ void __launch_bounds__(WARPS * 32, 1)
assemble_q2_sumfac_kernel(int64_t nel, int64_t nnode, const double* __restrict__ xyz, const int32_t* __restrict__ conn,
                          const int32_t* __restrict__ dofL, const SfTables* __restrict__ tabs, const SlotT* __restrict__ lslot,
                          const int64_t* __restrict__ rowptr, double* __restrict__ Aval, const double* __restrict__ u,
                          double* __restrict__ rhs, double nu, double fsrc, const SfGalArgs ga) {
  constexpr int NVE = kSfNVE, NG = kSfNG;
  using Smem = SfSmem<WARPS>;
  B2_DYN_SHARED(double, smem);
  double* sL = smem;                  // [3][4]
  double* sD = sL + 12;               // [3][4]
  double* sM = sD + 12;               // [4][9][4]
  double* sW = sM + 144;              // [64]
  double* sA = sW + NG;               // a, b, c, -; then int hi[8][4]; then nat2lat  (GAL)
  const int* sHi = reinterpret_cast<const int*>(sA + 4);
  const unsigned short* sN2L = reinterpret_cast<const unsigned short*>(sA + 4 + 16);      // [729] (GAL)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = smem + Smem::tab_doubles + (GAL ? Smem::gal_doubles : 0) + wib * (GAL ? Smem::warp_doubles_gal : Smem::warp_doubles);
  double* sX = wbase;                                   // [3][28] coordinates, lattice order
  double* sU = sX + 3 * 28;                             // [28] current solution
  long long* sRow = reinterpret_cast<long long*>(sU + 28);     // [28] rowptr[dof]
  double* sK = reinterpret_cast<double*>(sRow + 28);    // [7][64]: K00 K01 K02 K11 K12 K22, w det
  double* sS1 = sK + 7 * NG;                            // [4][9][18] stage-1 sums of the four terms of a round
  double* Bs = sK;                                      // [27][27] element matrix (lattice order), reuses sK / sS1
  SlotT* sSlot = reinterpret_cast<SlotT*>(sK + Smem::kb_doubles);      // [736] slot map of the current element
  double* Dacc = sK + Smem::kb_doubles + Smem::slot_doubles;      // [27][27] Galerkin matrix of the coarse element (GAL)
  long long* sCRow = reinterpret_cast<long long*>(Dacc + Smem::dacc_doubles);      // [28] Cp[coarse dof] (GAL)

  for (int t = threadIdx.x; t < Smem::tab_doubles; t += blockDim.x) smem[t] = reinterpret_cast<const double*>(tabs)[t];
  if (GAL) {
    const double* src = reinterpret_cast<const double*>(ga.tab);
    for (int t = threadIdx.x; t < Smem::gal_doubles; t += blockDim.x) sA[t] = src[t];
  }
  const int my_node = tabs->node_of[lane < NVE ? lane : 0];
  __syncthreads();

  // phase A: lane = (a half, b, c); phase B: lane = (i2, i3, j3) (lanes 27..31 shadow lane 26 and store nothing)
  const int pa_c = lane & 3, pa_b = (lane >> 2) & 3, pa_ah = lane >> 4;
  const int lb = lane < NVE ? lane : NVE - 1;
  const int pb_i2 = lb / 9, pb_i3j3 = lb % 9;
  const int ri1 = lb / 9, ri2 = (lb / 3) % 3, ri3 = lb % 3;      // lattice position of row `lane` (source term)
  double ui2[2][4];                                              // u^p_i2(p_b) of this lane's i2 (stage 2)
#pragma unroll
  for (int b = 0; b < 4; b++) {
    ui2[0][b] = sL[pb_i2 * 4 + b];
    ui2[1][b] = sD[pb_i2 * 4 + b];
  }

  const int64_t nunits = GAL ? (nel >> 3) : nel;
  const int64_t ustride = (int64_t)gridDim.x * WARPS;
  // node id and dof of the NEXT element are fetched one element ahead (the coordinate / solution / row-start loads
  // that depend on them then start without waiting for this first level of the gather)
  int64_t nd_pf = 0;
  int dof_pf = 0;
  {
    const int64_t u0 = (int64_t)blockIdx.x * WARPS + wib;
    if (u0 < nunits && lane < NVE) {
      const int64_t e0 = GAL ? u0 * 8 : u0;
      nd_pf = conn[e0 * 27 + my_node];
      dof_pf = dofL[e0 * NVE + lane];
    }
  }
  for (int64_t unit = (int64_t)blockIdx.x * WARPS + wib; unit < nunits; unit += ustride) {
    unsigned cmaskbits = 0;            // coarse Dirichlet dofs of this coarse element (natural order), one bit per dof
    if (GAL) {
      for (int t = lane; t < Smem::dacc_doubles; t += 32) Dacc[t] = 0.0;
      int cm = 0;
      if (lane < NVE) {
        const int32_t dI = ga.cd[unit * NVE + lane];
        sCRow[lane] = ga.Cp[dI];
        cm = ga.cmask ? (int)ga.cmask[dI] : 0;
      }
      cmaskbits = __ballot_sync(0xffffffffu, cm != 0);
    }
#pragma unroll 1
    for (int child = 0; child < (GAL ? 8 : 1); child++) {
      const int64_t e = GAL ? unit * 8 + child : unit;
      // ---- gather (lattice order): coordinates, dofs, current solution, row starts
      // the element's slot map feeds the addresses of the scatter at the END of the element: its 16-byte pieces are
      // requested now, sit in registers while the geometry is computed (whose register demand is far below the
      // stiffness stages') and go to shared memory after phase A -- the scatter then never waits for global memory
      constexpr int SLOT_VECS = kSfSlotStride * (int)sizeof(SlotT) / 16, SLOT_LOADS = (SLOT_VECS + 31) / 32;
      double2 sv[SLOT_LOADS];
      {
        const double2* src = reinterpret_cast<const double2*>(lslot + (size_t)e * kSfSlotStride);
#pragma unroll
        for (int q = 0; q < SLOT_LOADS; q++)
          if (lane + 32 * q < SLOT_VECS) sv[q] = src[lane + 32 * q];
      }
      int mydof = 0, fm = 0;
      if (lane < NVE) {
        const int64_t nd = nd_pf;
        mydof = dof_pf;
        const int64_t en = (GAL && child < 7) ? e + 1 : (GAL ? (unit + ustride) * 8 : unit + ustride);
        if (en < nel) {
          nd_pf = conn[en * 27 + my_node];
          dof_pf = dofL[en * NVE + lane];
        }
        sX[lane] = xyz[nd];
        sX[28 + lane] = xyz[nnode + nd];
        sX[56 + lane] = xyz[2 * nnode + nd];
        sU[lane] = u ? u[mydof] : 0.0;
        sRow[lane] = rowptr[mydof];
        if (GAL && ga.fmask) fm = (int)ga.fmask[mydof];      // used after the element matrix is formed
      }
      __syncwarp();

      // ---- A. geometry at this lane's two Gauss points (a = 2 ah, 2 ah + 1; b; c)
      {
        double Lc[3], Dc[3], Lb[3], Db[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          Lc[i] = sL[i * 4 + pa_c]; Dc[i] = sD[i * 4 + pa_c];
          Lb[i] = sL[i * 4 + pa_b]; Db[i] = sD[i * 4 + pa_b];
        }
        double J[2][3][3];              // [a of the pair][reference direction][coordinate]
#pragma unroll
        for (int d = 0; d < 3; d++) {
          double Z00[3], Z10[3], Z01[3];
#pragma unroll
          for (int i1 = 0; i1 < 3; i1++) {
            double y0[3], y1[3];
#pragma unroll
            for (int i2 = 0; i2 < 3; i2++) {
              const double* x = sX + d * 28 + i1 * 9 + i2 * 3;
              const double x0 = x[0], x1 = x[1], x2 = x[2];
              y0[i2] = fma(Lc[2], x2, fma(Lc[1], x1, Lc[0] * x0));
              y1[i2] = fma(Dc[2], x2, fma(Dc[1], x1, Dc[0] * x0));
            }
            Z00[i1] = fma(Lb[2], y0[2], fma(Lb[1], y0[1], Lb[0] * y0[0]));
            Z10[i1] = fma(Db[2], y0[2], fma(Db[1], y0[1], Db[0] * y0[0]));
            Z01[i1] = fma(Lb[2], y1[2], fma(Lb[1], y1[1], Lb[0] * y1[0]));
          }
#pragma unroll
          for (int t = 0; t < 2; t++) {
            const int a = 2 * pa_ah + t;
            const double La0 = sL[a], La1 = sL[4 + a], La2 = sL[8 + a];
            const double Da0 = sD[a], Da1 = sD[4 + a], Da2 = sD[8 + a];
            J[t][0][d] = fma(Da2, Z00[2], fma(Da1, Z00[1], Da0 * Z00[0]));
            J[t][1][d] = fma(La2, Z10[2], fma(La1, Z10[1], La0 * Z10[0]));
            J[t][2][d] = fma(La2, Z01[2], fma(La1, Z01[1], La0 * Z01[0]));
          }
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const int g = (2 * pa_ah + t) * 16 + pa_b * 4 + pa_c;
          const double J00 = J[t][0][0], J01 = J[t][0][1], J02 = J[t][0][2];
          const double J10 = J[t][1][0], J11 = J[t][1][1], J12 = J[t][1][2];
          const double J20 = J[t][2][0], J21 = J[t][2][1], J22 = J[t][2][2];
          const double det = J00 * (J11 * J22 - J12 * J21) + J01 * (J12 * J20 - J10 * J22) + J02 * (J10 * J21 - J11 * J20);
          const double id = 1.0 / det;
          // JacI[d][a] as ElemType.hpp:1478-1486; grad phi_n [d] = sum_a dphi_n/dxi_a JacI[d][a]
          const double I00 = (-J12 * J21 + J11 * J22) * id, I01 = (J02 * J21 - J01 * J22) * id, I02 = (-J02 * J11 + J01 * J12) * id;
          const double I10 = (J12 * J20 - J10 * J22) * id, I11 = (-J02 * J20 + J00 * J22) * id, I12 = (J02 * J10 - J00 * J12) * id;
          const double I20 = (-J11 * J20 + J10 * J21) * id, I21 = (J01 * J20 - J00 * J21) * id, I22 = (-J01 * J10 + J00 * J11) * id;
          const double wd = det * sW[g];
          // K_ab = weight * sum_d JacI[d][a] JacI[d][b]
          sK[0 * NG + g] = wd * fma(I20, I20, fma(I10, I10, I00 * I00));
          sK[1 * NG + g] = wd * fma(I20, I21, fma(I10, I11, I00 * I01));
          sK[2 * NG + g] = wd * fma(I20, I22, fma(I10, I12, I00 * I02));
          sK[3 * NG + g] = wd * fma(I21, I21, fma(I11, I11, I01 * I01));
          sK[4 * NG + g] = wd * fma(I21, I22, fma(I11, I12, I01 * I02));
          sK[5 * NG + g] = wd * fma(I22, I22, fma(I12, I12, I02 * I02));
          sK[6 * NG + g] = wd;
        }
      }
      {
        double2* dst = reinterpret_cast<double2*>(sSlot);
#pragma unroll
        for (int q = 0; q < SLOT_LOADS; q++)
          if (lane + 32 * q < SLOT_VECS) dst[lane + 32 * q] = sv[q];
      }
      __syncwarp();

      // ---- B. stiffness: the nine (alpha, beta) terms grouped by what dimension 1 contributes (class = 2 p1 + q1)
      double out[27];
#pragma unroll
      for (int t = 0; t < 27; t++) out[t] = 0.0;
      {
        double T[3][4];
        const double* S = sS1 + pb_i3j3 * 18;
        // round 1: the terms with xi on either side.  Stage-1 sums: (K00, l l), (K01, l l), (K02, l l'), (K02, l' l)
        if (lane < 16) {
          sf_stage1<0>(sK + 0 * NG, sS1 + 0 * 162, lane);
          sf_stage1<0>(sK + 1 * NG, sS1 + 1 * 162, lane);
          sf_stage1<1>(sK + 2 * NG, sS1 + 2 * 162, lane);
          sf_stage1<2>(sK + 2 * NG, sS1 + 3 * 162, lane);
        }
        __syncwarp();
        sf_zero(T);                                   // class (1,1): (xi, xi)
        sf_stage2<0, 0>(S + 0 * 162, ui2, T);
        sf_stage3<3>(T, out);
        sf_zero(T);                                   // class (1,0): (xi, eta), (xi, zeta)
        sf_stage2<0, 1>(S + 1 * 162, ui2, T);
        sf_stage2<0, 0>(S + 2 * 162, ui2, T);
        sf_stage3<2>(T, out);
        sf_zero(T);                                   // class (0,1): (eta, xi), (zeta, xi)
        sf_stage2<1, 0>(S + 1 * 162, ui2, T);
        sf_stage2<0, 0>(S + 3 * 162, ui2, T);
        sf_stage3<1>(T, out);
        __syncwarp();
        // round 2, class (0,0): (eta, eta), (eta, zeta), (zeta, eta), (zeta, zeta)
        if (lane < 16) {
          sf_stage1<0>(sK + 3 * NG, sS1 + 0 * 162, lane);
          sf_stage1<1>(sK + 4 * NG, sS1 + 1 * 162, lane);
          sf_stage1<2>(sK + 4 * NG, sS1 + 2 * 162, lane);
          sf_stage1<3>(sK + 5 * NG, sS1 + 3 * 162, lane);
        }
        __syncwarp();
        sf_zero(T);
        sf_stage2<1, 1>(S + 0 * 162, ui2, T);
        sf_stage2<1, 0>(S + 1 * 162, ui2, T);
        sf_stage2<0, 1>(S + 2 * 162, ui2, T);
        sf_stage2<0, 0>(S + 3 * 162, ui2, T);
        sf_stage3<0>(T, out);
      }
      // source term of row `lane`: sum_g phi(g) w det = sum_a l_i1(a) sum_b l_i2(b) sum_c l_i3(c) wd[a][b][c]
      double src = 0.0;
      if (rhs) {
        const double* Wd = sK + 6 * NG;
#pragma unroll
        for (int a = 0; a < 4; a++) {
          double ta = 0.0;
#pragma unroll
          for (int b = 0; b < 4; b++) {
            const double2 k01 = *reinterpret_cast<const double2*>(Wd + a * 16 + b * 4);
            const double2 k23 = *reinterpret_cast<const double2*>(Wd + a * 16 + b * 4 + 2);
            const double tb = fma(sL[ri3 * 4 + 3], k23.y, fma(sL[ri3 * 4 + 2], k23.x, fma(sL[ri3 * 4 + 1], k01.y, sL[ri3 * 4] * k01.x)));
            ta = fma(sL[ri2 * 4 + b], tb, ta);
          }
          src = fma(sL[ri1 * 4 + a], ta, src);
        }
      }
      __syncwarp();                    // every lane is done with K: the buffer becomes the element matrix

      // ---- C. element matrix -> shared memory (lattice order), scaled by nu
      if (lane < NVE) {
        const int ri = (lb / 3) * 27 + lb % 3;          // (3 i2 + i3) * 27 + j3
#pragma unroll
        for (int i1 = 0; i1 < 3; i1++)
#pragma unroll
          for (int j1 = 0; j1 < 3; j1++)
#pragma unroll
            for (int j2 = 0; j2 < 3; j2++) Bs[i1 * 243 + j1 * 9 + j2 * 3 + ri] = nu * out[(i1 * 3 + j1) * 3 + j2];
      }
      __syncwarp();

      // ---- residual F_i = fsrc * sum_g phi_i w_g - (B u)_i; with the fused Galerkin product row `lane` stays in registers
      double R[27];
      if (GAL || rhs) {
#pragma unroll
        for (int j = 0; j < NVE; j++) R[j] = Bs[lb * NVE + j];
      }
      if (rhs && lane < NVE) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NVE; j++) s = fma(R[j], sU[j], s);
        atomicAdd(&rhs[mydof], fsrc * src - s);
      }
      // ---- scatter: lattice (i, j) order through the slot map
      {
        for (int idx = lane; idx < NVE * NVE; idx += 32) {
          const int i = idx / NVE;
          atomicAdd(&Aval[sRow[i] + (long long)sSlot[idx]], Bs[idx]);
        }
      }
      __syncwarp();

      if (GAL) {
        const double ka = sA[0], kb = sA[1], kc = sA[2];
        const bool hi1 = sHi[child * 4 + 0] != 0, hi2 = sHi[child * 4 + 1] != 0, hi3 = sHi[child * 4 + 2] != 0;
        // rows / columns of Dirichlet fine dofs do not take part in the Galerkin product
        const unsigned mask = __ballot_sync(0xffffffffu, fm != 0);
        if (mask) {
          const bool rowdead = (mask >> lb) & 1u;
#pragma unroll
          for (int j = 0; j < NVE; j++)
            if (rowdead || ((mask >> j) & 1u)) R[j] = 0.0;
        }
        sf_kron_pass<1>(R, hi3, ka, kb, kc);          // T = B Pc, row `lane`: the column index runs over (j1, j2, j3)
        sf_kron_pass<3>(R, hi2, ka, kb, kc);
        sf_kron_pass<9>(R, hi1, ka, kb, kc);
        if (lane < NVE) {
#pragma unroll
          for (int j = 0; j < NVE; j++) Bs[lane * NVE + j] = R[j];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NVE; i++) R[i] = Bs[i * NVE + lb];       // column `lane` of T
        sf_kron_pass<1>(R, hi3, ka, kb, kc);          // D = Pc^T T, column `lane`
        sf_kron_pass<3>(R, hi2, ka, kb, kc);
        sf_kron_pass<9>(R, hi1, ka, kb, kc);
        if (lane < NVE) {
#pragma unroll
          for (int i = 0; i < NVE; i++) Dacc[i * NVE + lane] += R[i];
        }
        __syncwarp();
      }
    }
    if (GAL) {
      // ---- the coarse element's Galerkin matrix: record (natural order) and scatter into the coarse operator
      const int64_t E = unit;
      const CSlotT* cslot = reinterpret_cast<const CSlotT*>(ga.cslot) + (size_t)E * (NVE * NVE);
      for (int idx = lane; idx < NVE * NVE; idx += 32) {
        const double v = Dacc[sN2L[idx]];
        if (ga.emat) ga.emat[(size_t)E * (NVE * NVE) + idx] = v;
        if (v == 0.0) continue;
        const int I = idx / NVE, Jn = idx - I * NVE;
        if (((cmaskbits >> I) | (cmaskbits >> Jn)) & 1u) continue;
        atomicAdd(&ga.Cv[sCRow[I] + (long long)cslot[idx]], v);
      }
      __syncwarp();
    }
  }
}

// ---- Galerkin product of a COARSER level pair from the recorded element matrices, through the same Kronecker factors:
// A_{l-1} = sum_E Pc(E)^T D_E Pc(E) with D_E the Galerkin element matrix of element E of level l (natural order, as the
// fused kernel above and this kernel record them).  One warp takes the 8 children of a coarser element one after the
// other: load D (5.8 KB, coalesced) into the lattice order, six 1-D passes in registers (2 x 3 x 27 triples per lane
// instead of 2 x 125 x 27 multiply-adds of the column-list form), sum in the warp's own tile, scatter + record once.
constexpr int kSfChainWarps = 12;
template <int WARPS>
struct SfChainSmem {
  static constexpr int gal_doubles = 4 + 16 + 736 / 4;      // abc, hi flags, nat2lat (as SfSmem)
  static constexpr int warp_doubles = 736 + 736 + 28 + 16;  // B, Dacc, row starts of the coarse dofs, lattice -> natural node (ints)
  static constexpr size_t bytes = (size_t)(gal_doubles + WARPS * warp_doubles) * sizeof(double);
};

template <int WARPS, typename CSlotT>
__global__ void __launch_bounds__(WARPS * 32, 1)
galerkin_chain_sumfac_kernel(int64_t nelc, const double* __restrict__ emat_in, const int32_t* __restrict__ dof_in, const SfGalArgs ga) {
  constexpr int NVE = kSfNVE;
  using Smem = SfChainSmem<WARPS>;
  B2_DYN_SHARED(double, smem);
  double* sA = smem;                  // a, b, c, -; then int hi[8][4]; then nat2lat
  const int* sHi = reinterpret_cast<const int*>(sA + 4);
  const unsigned short* sN2L = reinterpret_cast<const unsigned short*>(sA + 4 + 16);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* wbase = smem + Smem::gal_doubles + wib * Smem::warp_doubles;
  double* Bs = wbase;                                   // [27][27] lattice order
  double* Dacc = Bs + 736;                              // [27][27] lattice order
  long long* sCRow = reinterpret_cast<long long*>(Dacc + 736);      // [28]
  int* sNodeAt = reinterpret_cast<int*>(sCRow + 28);    // [27] natural node at lattice position m
  {
    const double* src = reinterpret_cast<const double*>(ga.tab);
    for (int t = threadIdx.x; t < Smem::gal_doubles; t += blockDim.x) sA[t] = src[t];
  }
  __syncthreads();
  if (lane < NVE) sNodeAt[sN2L[lane * NVE] / NVE] = lane;      // nat2lat[I * 27 + J] = lattice(I) * 27 + lattice(J), lattice(J) < 27
  __syncwarp();
  const int lb = lane < NVE ? lane : NVE - 1;
  const double ka = sA[0], kb = sA[1], kc = sA[2];
  for (int64_t unit = (int64_t)blockIdx.x * WARPS + wib; unit < nelc; unit += (int64_t)gridDim.x * WARPS) {
    for (int t = lane; t < 736; t += 32) Dacc[t] = 0.0;
    int cm = 0;
    if (lane < NVE) {
      const int32_t dI = ga.cd[unit * NVE + lane];
      sCRow[lane] = ga.Cp[dI];
      cm = ga.cmask ? (int)ga.cmask[dI] : 0;
    }
    const unsigned cmaskbits = __ballot_sync(0xffffffffu, cm != 0);
#pragma unroll 1
    for (int child = 0; child < 8; child++) {
      const int64_t e = unit * 8 + child;
      const double* De = emat_in + (size_t)e * (NVE * NVE);
      for (int idx = lane; idx < NVE * NVE; idx += 32) Bs[sN2L[idx]] = De[idx];
      int fm = 0;
      if (ga.fmask && lane < NVE) fm = (int)ga.fmask[dof_in[e * NVE + sNodeAt[lane]]];
      const unsigned mask = __ballot_sync(0xffffffffu, fm != 0);      // this level's Dirichlet dofs, lattice order
      __syncwarp();
      const bool hi1 = sHi[child * 4 + 0] != 0, hi2 = sHi[child * 4 + 1] != 0, hi3 = sHi[child * 4 + 2] != 0;
      double R[27];
#pragma unroll
      for (int j = 0; j < NVE; j++) R[j] = Bs[lb * NVE + j];
      if (mask) {
        const bool rowdead = (mask >> lb) & 1u;
#pragma unroll
        for (int j = 0; j < NVE; j++)
          if (rowdead || ((mask >> j) & 1u)) R[j] = 0.0;
      }
      sf_kron_pass<1>(R, hi3, ka, kb, kc);
      sf_kron_pass<3>(R, hi2, ka, kb, kc);
      sf_kron_pass<9>(R, hi1, ka, kb, kc);
      __syncwarp();
      if (lane < NVE) {
#pragma unroll
        for (int j = 0; j < NVE; j++) Bs[lane * NVE + j] = R[j];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < NVE; i++) R[i] = Bs[i * NVE + lb];
      sf_kron_pass<1>(R, hi3, ka, kb, kc);
      sf_kron_pass<3>(R, hi2, ka, kb, kc);
      sf_kron_pass<9>(R, hi1, ka, kb, kc);
      if (lane < NVE) {
#pragma unroll
        for (int i = 0; i < NVE; i++) Dacc[i * NVE + lane] += R[i];
      }
      __syncwarp();
    }
    const CSlotT* cslot = reinterpret_cast<const CSlotT*>(ga.cslot) + (size_t)unit * (NVE * NVE);
    for (int idx = lane; idx < NVE * NVE; idx += 32) {
      const double v = Dacc[sN2L[idx]];
      if (ga.emat) ga.emat[(size_t)unit * (NVE * NVE) + idx] = v;
      if (v == 0.0) continue;
      const int I = idx / NVE, Jn = idx - I * NVE;
      if (((cmaskbits >> I) | (cmaskbits >> Jn)) & 1u) continue;
      atomicAdd(&ga.Cv[sCRow[I] + (long long)cslot[idx]], v);
    }
    __syncwarp();
  }
}
