// Galerkin coarse operator C = P^T A P by COARSE-ELEMENT GATHER.  Fast path of
// SparseMatrix::matrix_PtAP -> MatPtAP (reference src/03_algebra/01_matrices/PetscMatrix.cpp:733-751)
// as LinearImplicitSystem::MGsolve calls it down the hierarchy (src/08_equations/00_stationary/
// LinearImplicitSystem.cpp:347-370) with the geometric prolongators of BuildProlongatorMatrix
// (:826-909) after ZeroInterpolatorDirichletNodes (:1032-1120).
//
// The prolongator of a refined mesh is element-local: every fine dof i of a coarse element E
// interpolates only from the coarse dofs of E, with the reference-element weights
// P_loc[a][J] = phi_J(x_a) (elem_type::set_prolongation_OneElement_All_FE, ElemType.cpp:439-532),
// and an assembled entry A[i][j] != 0 needs i and j in one fine element, hence in one coarse
// element.  So
//     C = sum_E  P_E^T ( W_E o A|_E ) P_E ,   P_E = D_f P_loc D_c ,
// where A|_E is the block of A on the NF fine dofs of E, W_E[a][b] = 1 / #(coarse elements that
// contain both a and b) removes the multiple counting of entries on shared faces/edges/vertices,
// and D_f / D_c zero the fine rows / coarse columns of Dirichlet dofs.  The multiplicity is the
// valence of the smallest sub-entity of E containing both points; in a HEX27 element the 27
// sub-entities (8 vertices, 12 edges, 6 faces, cell) are labelled by a 3-trit lattice code and
// each has one node at its centre, so the host passes valence[E][27] = number of elements at that
// node.  No A*P temporary, no hashing on global columns, no binary searches: one warp owns one
// coarse element, streams the NF rows of A once (coalesced, software-prefetched), maps columns to
// local indices through a 256-slot shared-memory table, and contracts with P_loc in shared
// memory.  HBM traffic ~2.2x the fine matrix; result scattered with fp64 atomics through a
// precomputed slot map (729 per coarse element).
#include "b2_common.cuh"

struct b2_galerkin {
  b2_ctx* ctx;
  b2_csr *Af, *Ac;  // borrowed
  int64_t nrows_f = 0, nrows_c = 0;   // sizes of the masks (the borrowed matrices may be gone at destruction)
  int64_t nelc;
  int nf, nc, pnnz;
  int32_t* fd;      // [nelc][nf] fine dofs in lattice order
  int32_t* cd;      // [nelc][nc] coarse dofs
  double* ploc;     // [nf][nc]
  int32_t* prs;     // [nf+1] sparse rows of ploc
  uint8_t* pi;      // [pnnz]
  double* pv;       // [pnnz]
  uint8_t* fent;    // [nf] entity code of every fine point
  uint8_t* val;     // [nelc][27]
  uint8_t* fmask;   // [Af->nrows] or null
  uint8_t* cmask;   // [Ac->nrows] or null
  void* slot;       // [nelc][nc*nc]
  int slot_bytes;
  // element-matrix chain (b2_assemble.cu): the nc x nc Galerkin matrix of every coarse element of this
  // plan, recorded when the plan is applied from element matrices; feeds the next-coarser plan
  double* emat;     // [nelc][nc*nc] or null
  void* chain_tab;  // per-child prolongator tables for b2_galerkin_apply_from_elements, or null
  int chain_tab_nve;
  void* chain_sf;   // Kronecker factors of the same prolongators (sum-factorised chain kernel), or null
  int chain_sf_tried;
};

namespace {

constexpr int kWarps = 16;
constexpr int kChunks = 4;      // prefetched 32-entry chunks per row (rows up to 128 entries)

template <int NF> struct GTraits;
template <> struct GTraits<125> { static constexpr int HB = 256, NFP = 128; };
template <> struct GTraits<27> { static constexpr int HB = 64, NFP = 32; };

__host__ __device__ constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <int NF, int NC>
struct GSmem {
  static constexpr int HB = GTraits<NF>::HB, NFP = GTraits<NF>::NFP;
  // per-warp doubles: winv[32], lv[NFP], Cw[NC*NC]; int64 rs[NFP]; int32 hkey[HB], fdw[NFP], rl[NFP];
  // bytes hval[HB], fmw[NFP], lb[NFP]
  static constexpr size_t warp_bytes =
      align_up((size_t)(32 + NFP + NC * NC) * 8 + (size_t)NFP * 8 + (size_t)(HB + 2 * NFP) * 4 + (size_t)(HB + 2 * NFP), 16);
  static size_t bytes(int pnnz) {
    return align_up((size_t)NF * NC * 8 + (size_t)pnnz * 8 + (size_t)(NF + 1) * 4 + (size_t)pnnz + NF + 729, 16) +
           kWarps * warp_bytes;
  }
};

__device__ __forceinline__ unsigned hashf(int j, int bits) { return ((unsigned)j * 2654435761u) >> (32 - bits); }

template <int NF, int NC, typename SlotT>
__global__ void __launch_bounds__(kWarps * 32, 1)
galerkin_kernel(int64_t nelc, int pnnz, const int32_t* __restrict__ fd, const int32_t* __restrict__ cd,
                const double* __restrict__ ploc, const int32_t* __restrict__ prs, const uint8_t* __restrict__ pi,
                const double* __restrict__ pv, const uint8_t* __restrict__ fent, const uint8_t* __restrict__ valence,
                const uint8_t* __restrict__ fmask, const uint8_t* __restrict__ cmask, const SlotT* __restrict__ slot,
                const int64_t* __restrict__ Ap, const int32_t* __restrict__ Ac, const double* __restrict__ Av,
                const int64_t* __restrict__ Cp, double* __restrict__ Cv) {
  constexpr int HB = GTraits<NF>::HB, NFP = GTraits<NF>::NFP;
  constexpr int HBITS = HB == 256 ? 8 : 6;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- CTA-shared tables
  double* sP = reinterpret_cast<double*>(smem_raw);          // [NF][NC]
  double* sPv = sP + NF * NC;                                  // [pnnz]
  int32_t* sPrs = reinterpret_cast<int32_t*>(sPv + pnnz);      // [NF+1]
  uint8_t* sPi = reinterpret_cast<uint8_t*>(sPrs + NF + 1);    // [pnnz]
  uint8_t* sEnt = sPi + pnnz;                                  // [NF]
  uint8_t* sJoin = sEnt + NF;                                  // [27][27]
  const size_t shared_bytes = align_up((size_t)NF * NC * 8 + (size_t)pnnz * 8 + (size_t)(NF + 1) * 4 + (size_t)pnnz + NF + 729, 16);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wb = smem_raw + shared_bytes + (size_t)wib * GSmem<NF, NC>::warp_bytes;
  double* winv = reinterpret_cast<double*>(wb);                // [32]
  double* lv = winv + 32;                                      // [NFP]
  double* Cw = lv + NFP;                                       // [NC*NC]
  int64_t* rs = reinterpret_cast<int64_t*>(Cw + NC * NC);      // [NFP]
  int32_t* hkey = reinterpret_cast<int32_t*>(rs + NFP);        // [HB]
  int32_t* fdw = hkey + HB;                                    // [NFP]
  int32_t* rl = fdw + NFP;                                     // [NFP]
  uint8_t* hval = reinterpret_cast<uint8_t*>(rl + NFP);        // [HB]
  uint8_t* fmw = hval + HB;                                    // [NFP]
  uint8_t* lb = fmw + NFP;                                     // [NFP]

  for (int t = threadIdx.x; t < NF * NC; t += blockDim.x) sP[t] = ploc[t];
  for (int t = threadIdx.x; t < pnnz; t += blockDim.x) { sPv[t] = pv[t]; sPi[t] = pi[t]; }
  for (int t = threadIdx.x; t <= NF; t += blockDim.x) sPrs[t] = prs[t];
  for (int t = threadIdx.x; t < NF; t += blockDim.x) sEnt[t] = fent[t];
  for (int t = threadIdx.x; t < 729; t += blockDim.x) {
    // join of two entity codes: per direction keep the side only if both points sit on it
    const int ea = t / 27, eb = t - ea * 27;
    int code = 0, mul = 1;
    int xa = ea, xb = eb;
    for (int d = 0; d < 3; d++) {
      const int ta = xa % 3, tb = xb % 3;
      xa /= 3; xb /= 3;
      code += ((ta == tb && ta != 1) ? ta : 1) * mul;
      mul *= 3;
    }
    sJoin[t] = (uint8_t)code;
  }
  __syncthreads();

  const unsigned lt_mask = (1u << lane) - 1u;

  for (int64_t E = (int64_t)blockIdx.x * kWarps + wib; E < nelc; E += (int64_t)gridDim.x * kWarps) {
    // ---- 1. per-element setup: local index table, row extents, weights
    for (int t = lane; t < HB; t += 32) hkey[t] = -1;
    for (int t = lane; t < NC * NC; t += 32) Cw[t] = 0.0;
    if (lane < 27) winv[lane] = 1.0 / (double)valence[E * 27 + lane];
    __syncwarp();
    for (int a = lane; a < NFP; a += 32) {
      if (a < NF) {
        const int32_t j = fd[E * NF + a];
        fdw[a] = j;
        fmw[a] = fmask ? fmask[j] : (uint8_t)0;
        const int64_t s = Ap[j];
        rs[a] = s;
        rl[a] = (int32_t)(Ap[j + 1] - s);
        unsigned h = hashf(j, HBITS);
        while (true) {
          const int old = atomicCAS(&hkey[h], -1, j);
          if (old == -1 || old == j) break;
          h = (h + 1) & (HB - 1);
        }
        hval[h] = (uint8_t)a;
      } else {
        fmw[a] = 1;
      }
    }
    __syncwarp();

    // ---- 2. rows of A on the fine dofs of E, one row at a time, next row prefetched in registers
    int32_t cj[kChunks], nj[kChunks];
    double cv[kChunks], nv[kChunks];
    int a = 0;
    while (a < NF && fmw[a]) a++;
    if (a < NF) {
      const int64_t s = rs[a];
      const int len = rl[a];
#pragma unroll
      for (int c = 0; c < kChunks; c++) {
        const int k = c * 32 + lane;
        nj[c] = -1;
        nv[c] = 0.0;
        if (k < len) { nj[c] = Ac[s + k]; nv[c] = Av[s + k]; }
      }
    }
    while (a < NF) {
#pragma unroll
      for (int c = 0; c < kChunks; c++) { cj[c] = nj[c]; cv[c] = nv[c]; }
      int an = a + 1;
      while (an < NF && fmw[an]) an++;
      if (an < NF) {
        const int64_t s = rs[an];
        const int len = rl[an];
#pragma unroll
        for (int c = 0; c < kChunks; c++) {
          const int k = c * 32 + lane;
          nj[c] = -1;
          nv[c] = 0.0;
          if (k < len) { nj[c] = Ac[s + k]; nv[c] = Av[s + k]; }
        }
      }
      const int ea27 = (int)sEnt[a] * 27;
      int n = 0;
      // entries of this row that fall inside E -> compact list (local column, weighted value)
      auto take = [&](int32_t j, double v) {
        bool ok = false;
        int b = 0;
        if (j >= 0) {
          unsigned h = hashf(j, HBITS);
          while (true) {
            const int key = hkey[h];
            if (key == j) { b = hval[h]; ok = !fmw[b]; break; }
            if (key == -1) break;
            h = (h + 1) & (HB - 1);
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = n + __popc(m & lt_mask);
          lb[pos] = (uint8_t)b;
          lv[pos] = v * winv[sJoin[ea27 + sEnt[b]]];
        }
        n += __popc(m);
      };
#pragma unroll
      for (int c = 0; c < kChunks; c++) {
        if (c * 32 < rl[a]) take(cj[c], cv[c]);
      }
      for (int k0 = kChunks * 32; k0 < rl[a]; k0 += 32) {      // rows longer than the prefetch window
        const int k = k0 + lane;
        int32_t j = -1;
        double v = 0.0;
        if (k < rl[a]) { j = Ac[rs[a] + k]; v = Av[rs[a] + k]; }
        take(j, v);
      }
      __syncwarp();
      // stage 1: T[J] = sum_b (W o A)[a][b] * P_loc[b][J]      (lane = J)
      // stage 2: C_E[I][J] += P_loc[a][I] * T[J]
      if (lane < NC) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int k = 0;
        for (; k + 3 < n; k += 4) {
          t0 = fma(lv[k], sP[(int)lb[k] * NC + lane], t0);
          t1 = fma(lv[k + 1], sP[(int)lb[k + 1] * NC + lane], t1);
          t2 = fma(lv[k + 2], sP[(int)lb[k + 2] * NC + lane], t2);
          t3 = fma(lv[k + 3], sP[(int)lb[k + 3] * NC + lane], t3);
        }
        for (; k < n; k++) t0 = fma(lv[k], sP[(int)lb[k] * NC + lane], t0);
        const double T = (t0 + t1) + (t2 + t3);
        for (int q = sPrs[a]; q < sPrs[a + 1]; q++) {
          const int I = sPi[q];
          Cw[I * NC + lane] = fma(sPv[q], T, Cw[I * NC + lane]);
        }
      }
      __syncwarp();
      a = an;
    }

    // ---- 3. scatter the coarse element matrix
    const SlotT* sl = slot + (size_t)E * (NC * NC);
    for (int idx = lane; idx < NC * NC; idx += 32) {
      const int I = idx / NC, J = idx - I * NC;
      const double v = Cw[idx];
      if (v == 0.0) continue;
      const int32_t dI = cd[E * NC + I];
      if (cmask) {
        const int32_t dJ = cd[E * NC + J];
        if (cmask[dI] || cmask[dJ]) continue;
      }
      atomicAdd(&Cv[Cp[dI] + (int64_t)sl[idx]], v);
    }
    __syncwarp();
  }
}

template <typename SlotT>
__global__ void cslot_kernel(int64_t total, int nc, const int32_t* __restrict__ cd, const int64_t* __restrict__ Cp,
                             const int32_t* __restrict__ Cc, SlotT* __restrict__ slot, int* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int nc2 = nc * nc;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t E = t / nc2;
    const int idx = (int)(t - E * nc2);
    const int I = idx / nc, J = idx - I * nc;
    const int32_t r = cd[E * nc + I], c = cd[E * nc + J];
    const int64_t s = Cp[r], e = Cp[r + 1];
    int64_t lo = s, hi = e;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (Cc[mid] < c) lo = mid + 1;
      else hi = mid;
    }
    if (lo >= e || Cc[lo] != c) atomicExch(err, 1);
    slot[t] = (SlotT)(lo - s);
  }
}

template <typename SlotT>
int build_cslots(b2_galerkin* g) {
  b2_ctx* c = g->ctx;
  const int64_t total = g->nelc * g->nc * g->nc;
  SlotT* s = nullptr;
  B2_TRY(b2_malloc(c, &s, (size_t)total));
  g->slot = s;
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  B2_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  B2_LAUNCH(c, cslot_kernel<SlotT>, b2_grid_for(c, total, 256, 8), 256, 0, total, g->nc, g->cd, g->Ac->rowptr,
            g->Ac->col, s, d_err);
  int err = 0;
  B2_TRY(b2_download(c, &err, d_err, 1));
  b2_free(c, d_err, 1);
  B2_CHECK(err == 0, "b2_galerkin_create: a coarse element couples dofs outside the pattern of C");
  return 0;
}

template <int NF, int NC, typename SlotT>
int launch_galerkin(b2_galerkin* g) {
  b2_ctx* c = g->ctx;
  b2_prof_scope prof(c, g);
  auto kern = galerkin_kernel<NF, NC, SlotT>;
  const size_t smem = GSmem<NF, NC>::bytes(g->pnnz);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)((g->nelc + kWarps - 1) / kWarps);
  if (grid > c->sm_count) grid = c->sm_count;
  B2_LAUNCH(c, kern, grid, kWarps * 32, smem, g->nelc, g->pnnz, g->fd, g->cd, g->ploc, g->prs, g->pi, g->pv, g->fent,
            g->val, g->fmask, g->cmask, (const SlotT*)g->slot, g->Af->rowptr, g->Af->col, g->Af->val, g->Ac->rowptr,
            g->Ac->val);
  return 0;
}

}  // namespace

// read-only view of a plan for the fused assembly kernel (b2_assemble.cu)
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
  void** chain_sf;
  int* chain_sf_tried;
};
int b2_galerkin_get_view(const b2_galerkin* g, b2_galerkin_view* v) {
  B2_CHECK(g && v, "b2_galerkin_get_view: null argument");
  b2_galerkin* m = const_cast<b2_galerkin*>(g);
  *v = b2_galerkin_view{g->Af, g->Ac, g->nelc, g->nf, g->nc, g->fd, g->cd, g->ploc, g->fmask, g->cmask, g->slot, g->slot_bytes,
                        &m->emat, &m->chain_tab, &m->chain_tab_nve, &m->chain_sf, &m->chain_sf_tried};
  return 0;
}

extern "C" {

int b2_galerkin_create(b2_csr* Af, b2_csr* Ac, int64_t nelc, int nf, int nc, const int32_t* fine_dofs,
                       const int32_t* coarse_dofs, const double* ploc, const uint8_t* fine_entity,
                       const uint8_t* valence, const uint8_t* fine_mask, const uint8_t* coarse_mask,
                       b2_galerkin** out) {
  *out = nullptr;
  B2_CHECK(Af && Ac && nelc > 0 && fine_dofs && coarse_dofs && ploc && fine_entity && valence,
           "b2_galerkin_create: null argument");
  B2_CHECK((nf == 125 && nc == 27) || (nf == 27 && nc == 8),
           "b2_galerkin_create: nf=%d nc=%d (supported: 125/27 triquadratic, 27/8 trilinear hexahedra)", nf, nc);
  B2_CHECK(Af->nrows == Af->ncols && Ac->nrows == Ac->ncols, "b2_galerkin_create: operators must be square");
  B2_CHECK(Ac->max_row <= 65536, "b2_galerkin_create: rows longer than 65536 entries");
  b2_ctx* c = Af->ctx;
  b2_galerkin* g = new b2_galerkin();
  g->ctx = c;
  g->Af = Af;
  g->Ac = Ac;
  g->nelc = nelc;
  g->nf = nf;
  g->nc = nc;
  // sparse rows of the element prolongator (exact zeros dropped)
  std::vector<int32_t> prs(nf + 1, 0);
  std::vector<uint8_t> pi;
  std::vector<double> pv;
  for (int a = 0; a < nf; a++) {
    for (int J = 0; J < nc; J++)
      if (ploc[a * nc + J] != 0.0) { pi.push_back((uint8_t)J); pv.push_back(ploc[a * nc + J]); }
    prs[a + 1] = (int32_t)pi.size();
  }
  g->pnnz = (int)pi.size();
  for (int a = 0; a < nf; a++) B2_CHECK(fine_entity[a] < 27, "b2_galerkin_create: bad entity code");
  B2_TRY(b2_malloc(c, &g->fd, (size_t)nelc * nf));
  B2_TRY(b2_upload(c, g->fd, fine_dofs, (size_t)nelc * nf));
  B2_TRY(b2_malloc(c, &g->cd, (size_t)nelc * nc));
  B2_TRY(b2_upload(c, g->cd, coarse_dofs, (size_t)nelc * nc));
  B2_TRY(b2_malloc(c, &g->ploc, (size_t)nf * nc));
  B2_TRY(b2_upload(c, g->ploc, ploc, (size_t)nf * nc));
  B2_TRY(b2_malloc(c, &g->prs, (size_t)nf + 1));
  B2_TRY(b2_upload(c, g->prs, prs.data(), (size_t)nf + 1));
  B2_TRY(b2_malloc(c, &g->pi, (size_t)g->pnnz));
  B2_TRY(b2_upload(c, g->pi, pi.data(), (size_t)g->pnnz));
  B2_TRY(b2_malloc(c, &g->pv, (size_t)g->pnnz));
  B2_TRY(b2_upload(c, g->pv, pv.data(), (size_t)g->pnnz));
  B2_TRY(b2_malloc(c, &g->fent, (size_t)nf));
  B2_TRY(b2_upload(c, g->fent, fine_entity, (size_t)nf));
  B2_TRY(b2_malloc(c, &g->val, (size_t)nelc * 27));
  B2_TRY(b2_upload(c, g->val, valence, (size_t)nelc * 27));
  g->fmask = g->cmask = nullptr;
  g->emat = nullptr;
  g->chain_tab = nullptr;
  g->chain_tab_nve = 0;
  g->chain_sf = nullptr;
  g->chain_sf_tried = 0;
  if (fine_mask) {
    g->nrows_f = Af->nrows;
    B2_TRY(b2_malloc(c, &g->fmask, (size_t)Af->nrows));
    B2_TRY(b2_upload(c, g->fmask, fine_mask, (size_t)Af->nrows));
  }
  if (coarse_mask) {
    g->nrows_c = Ac->nrows;
    B2_TRY(b2_malloc(c, &g->cmask, (size_t)Ac->nrows));
    B2_TRY(b2_upload(c, g->cmask, coarse_mask, (size_t)Ac->nrows));
  }
  g->slot_bytes = Ac->max_row <= 256 ? 1 : 2;
  if (g->slot_bytes == 1) B2_TRY(build_cslots<uint8_t>(g));
  else B2_TRY(build_cslots<uint16_t>(g));
  *out = g;
  return 0;
}

int b2_galerkin_apply(b2_galerkin* g) {
  b2_ctx* c = g->ctx;
  B2_CUDA(cudaMemsetAsync(g->Ac->val, 0, (size_t)g->Ac->nnz * sizeof(double), c->stream));
  if (g->nf == 125) {
    if (g->slot_bytes == 1) return launch_galerkin<125, 27, uint8_t>(g);
    return launch_galerkin<125, 27, uint16_t>(g);
  }
  if (g->slot_bytes == 1) return launch_galerkin<27, 8, uint8_t>(g);
  return launch_galerkin<27, 8, uint16_t>(g);
}

int b2_galerkin_destroy(b2_galerkin* g) {
  if (!g) return 0;
  b2_ctx* c = g->ctx;
  cudaStreamSynchronize(c->stream);
  b2_free(c, g->fd, (size_t)g->nelc * g->nf);
  b2_free(c, g->cd, (size_t)g->nelc * g->nc);
  b2_free(c, g->ploc, (size_t)g->nf * g->nc);
  b2_free(c, g->prs, (size_t)g->nf + 1);
  b2_free(c, g->pi, (size_t)g->pnnz);
  b2_free(c, g->pv, (size_t)g->pnnz);
  b2_free(c, g->fent, (size_t)g->nf);
  b2_free(c, g->val, (size_t)g->nelc * 27);
  if (g->fmask) b2_free(c, g->fmask, (size_t)g->nrows_f);
  if (g->cmask) b2_free(c, g->cmask, (size_t)g->nrows_c);
  const size_t ns = (size_t)g->nelc * g->nc * g->nc;
  if (g->emat) b2_free(c, g->emat, ns);
  if (g->chain_tab) { cudaFree(g->chain_tab); }
  if (g->chain_sf) { cudaFree(g->chain_sf); }
  if (g->slot_bytes == 1) b2_free(c, (uint8_t*)g->slot, ns);
  else b2_free(c, (uint16_t*)g->slot, ns);
  delete g;
  return 0;
}

}  // extern "C"
