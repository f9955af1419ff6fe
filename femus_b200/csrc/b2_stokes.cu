// Steady Stokes assembly plan: host side of stokes_kernel (b2_stokes_kernel.cuh).  Replaces the assembly callback of
// the reference's applications/003_NavierStokes/SteadyStokes/main.cpp (AssembleMatrixResNS, :290-598) for Taylor-Hood
// pairs on meshes of one element type per plan; the system matrix carries the pattern of
// LinearEquation::GetSparsityPatternSize for the variables U, V, W, P (b2h_system_sparsity_create).
#include "b2_common.cuh"

struct b2_stokes {
  b2_ctx* ctx = nullptr;
  b2_mesh* mesh = nullptr;      // borrowed
  b2_csr* A = nullptr;          // borrowed
  int nv = 0, np = 0, ng = 0;
  int32_t* edof = nullptr;      // [nel][4][27]
  double *tabv = nullptr, *tabp = nullptr;
  double* tabns = nullptr;      // phi, dxi, deta, dzeta, w of the velocity element (Navier-Stokes kernel); null without phi_v
  unsigned short* slot = nullptr;     // [nel][3 nv nv + 6 nv np] position of every element coupling inside its CSR row
  unsigned short* slot_ns = nullptr;  // [nel][9 nv nv + 6 nv np] the same for the ten blocks of the Newton Jacobian
  int64_t nel = 0;
};

namespace {
#include "b2_stokes_kernel.cuh"
#include "b2_ns_kernel.cuh"
#include "b2_neumann_kernel.cuh"
size_t stokes_smem(int nv, int np, int ng) { return (size_t)kStokesWarps * (size_t)stokes_warp_doubles_host(nv, np, ng) * sizeof(double); }

// element -> CSR slot map of a plan (ns: the nine velocity blocks of the Newton Jacobian instead of the three diagonals);
// fails loudly when an element coupling is not an entry of the pattern
int build_slot_map(b2_stokes* p, bool ns) {
  b2_ctx* c = p->ctx;
  const size_t per = ns ? (size_t)ns_slots_per_element(p->nv, p->np) : (size_t)stokes_slots_per_element(p->nv, p->np);
  const size_t total = (size_t)p->nel * per;
  unsigned short** dst = ns ? &p->slot_ns : &p->slot;
  B2_TRY(b2_malloc(c, dst, total));
  int* d_err = nullptr;
  B2_TRY(b2_malloc(c, &d_err, 1));
  int h_err = 0;
  B2_TRY(b2_upload(c, d_err, &h_err, 1));
  const int grid = b2_grid_for(c, (int64_t)total, 256, 8);
  if (ns) B2_LAUNCH(c, ns_slot_kernel, grid, 256, 0, p->nel, p->nv, p->np, p->edof, p->A->rowptr, p->A->col, *dst, d_err);
  else B2_LAUNCH(c, stokes_slot_kernel, grid, 256, 0, p->nel, p->nv, p->np, p->edof, p->A->rowptr, p->A->col, *dst, d_err);
  const int rc = b2_download(c, &h_err, d_err, 1);
  b2_free(c, d_err, 1);
  if (rc) return rc;
  B2_CHECK(h_err != 1, "b2_%s_create: an element coupling is not an entry of the system matrix's pattern", ns ? "ns" : "stokes");
  B2_CHECK(h_err != 2, "b2_%s_create: a row of the system matrix is longer than 65536 entries", ns ? "ns" : "stokes");
  return 0;
}
}  // namespace

extern "C" {

int b2_stokes_create(b2_mesh* mesh, b2_csr* A, const int32_t* elem_dofs, int nve_v, int nve_p, int ngauss, const double* dxi,
                     const double* deta, const double* dzeta, const double* weights, const double* phi_p, b2_stokes** out) {
  B2_CHECK(mesh && A && elem_dofs && dxi && deta && dzeta && weights && phi_p && out, "b2_stokes_create: null argument");
  B2_CHECK(nve_v >= 4 && nve_v <= 27 && nve_p >= 1 && nve_p <= 8 && ngauss >= 1 && ngauss <= 64,
           "b2_stokes_create: nve_v=%d (4..27), nve_p=%d (1..8) or ngauss=%d (1..64) out of range", nve_v, nve_p, ngauss);
  B2_CHECK(A->nrows == A->ncols, "b2_stokes_create: the system matrix must be square");
  b2_ctx* c = nullptr;
  int64_t nnode = 0, nel = 0;
  const double* xyz = nullptr;
  const int32_t* conn = nullptr;
  b2_mesh_view(mesh, &c, &nnode, &nel, &xyz, &conn);
  // every listed dof must be a row of A (checked here on the host) and every element coupling a pattern entry
  // (checked on the device while the slot map is built)
  for (int64_t e = 0; e < nel; e++)
    for (int k = 0; k < 4; k++)
      for (int i = 0; i < (k < 3 ? nve_v : nve_p); i++) {
        const int32_t d = elem_dofs[(e * 4 + k) * 27 + i];
        B2_CHECK(d >= 0 && d < A->nrows, "b2_stokes_create: element %lld variable %d node %d: dof %d outside the system", (long long)e, k, i, (int)d);
      }
  b2_stokes* p = new b2_stokes();
  p->ctx = c;
  p->mesh = mesh;
  p->A = A;
  p->nv = nve_v;
  p->np = nve_p;
  p->ng = ngauss;
  p->nel = nel;
  *out = p;
  const size_t nt = (size_t)3 * ngauss * nve_v + ngauss;
  std::vector<double> tab(nt);
  std::copy(dxi, dxi + ngauss * nve_v, tab.begin());
  std::copy(deta, deta + ngauss * nve_v, tab.begin() + ngauss * nve_v);
  std::copy(dzeta, dzeta + ngauss * nve_v, tab.begin() + 2 * ngauss * nve_v);
  std::copy(weights, weights + ngauss, tab.begin() + 3 * ngauss * nve_v);
  const int rc = [&]() -> int {
    B2_TRY(b2_malloc(c, &p->edof, (size_t)nel * 108));
    B2_TRY(b2_malloc(c, &p->tabv, nt));
    B2_TRY(b2_malloc(c, &p->tabp, (size_t)ngauss * nve_p));
    B2_TRY(b2_upload(c, p->edof, elem_dofs, (size_t)nel * 108));
    B2_TRY(b2_upload(c, p->tabv, tab.data(), nt));
    B2_TRY(b2_upload(c, p->tabp, phi_p, (size_t)ngauss * nve_p));
    B2_CHECK(stokes_smem(nve_v, nve_p, ngauss) <= 227 * 1024, "b2_stokes_create: the element tables exceed the SM's shared memory");
    B2_TRY(build_slot_map(p, false));
    return 0;
  }();
  if (rc) {               // nothing half-built is handed out
    b2_stokes_destroy(p);
    *out = nullptr;
  }
  return rc;
}

/* A += the Stokes element blocks, rhs += the residual F = -B sol at the current solution (sol in system numbering,
 * NULL = zero; rhs NULL = matrix only).  A and rhs are not zeroed here (the callback calls myKK->zero(), :372). */
int b2_stokes_assemble(b2_stokes* p, const b2_vec* sol, b2_vec* rhs, double IRe) {
  B2_CHECK(p, "b2_stokes_assemble: null plan");
  B2_CHECK((!sol || sol->n >= p->A->nrows) && (!rhs || rhs->n >= p->A->nrows), "b2_stokes_assemble: vector shorter than the system");
  b2_ctx* c = nullptr;
  int64_t nnode = 0, nel = 0;
  const double* xyz = nullptr;
  const int32_t* conn = nullptr;
  b2_mesh_view(p->mesh, &c, &nnode, &nel, &xyz, &conn);
  const size_t smem = stokes_smem(p->nv, p->np, p->ng);
  // the attribute belongs to the function, not to the plan: set for THIS launch (several plans of a mixed mesh coexist)
  const stokes_kernel_t kern = stokes_kernel_for(p->nv, p->np);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
  int per_sm = 1;      // resident CTAs by registers and shared memory
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kStokesWarps * 32, smem));
  per_sm = per_sm < 1 ? 1 : per_sm;
  B2_LAUNCH(c, kern, b2_grid_for(c, nel, kStokesWarps, per_sm), kStokesWarps * 32, smem, nel, nnode, p->nv, p->np, p->ng, xyz, conn, p->edof,
            p->tabv, p->tabp, p->A->rowptr, p->slot, p->A->val, sol ? sol->d : nullptr, rhs ? rhs->d : nullptr, IRe);
  return 0;
}

/* the Navier-Stokes twin of b2_stokes_create: additionally the velocity element's phi table [ngauss][nve_v] */
int b2_ns_create(b2_mesh* mesh, b2_csr* A, const int32_t* elem_dofs, int nve_v, int nve_p, int ngauss, const double* phi_v, const double* dxi,
                 const double* deta, const double* dzeta, const double* weights, const double* phi_p, b2_stokes** out) {
  B2_CHECK(phi_v, "b2_ns_create: null argument");
  B2_TRY(b2_stokes_create(mesh, A, elem_dofs, nve_v, nve_p, ngauss, dxi, deta, dzeta, weights, phi_p, out));
  b2_stokes* p = *out;
  const int rc = [&]() -> int {
    const size_t nt = (size_t)4 * ngauss * nve_v + ngauss;
    std::vector<double> tab(nt);
    std::copy(phi_v, phi_v + ngauss * nve_v, tab.begin());
    std::copy(dxi, dxi + ngauss * nve_v, tab.begin() + ngauss * nve_v);
    std::copy(deta, deta + ngauss * nve_v, tab.begin() + 2 * ngauss * nve_v);
    std::copy(dzeta, dzeta + ngauss * nve_v, tab.begin() + 3 * ngauss * nve_v);
    std::copy(weights, weights + ngauss, tab.begin() + 4 * ngauss * nve_v);
    B2_TRY(b2_malloc(p->ctx, &p->tabns, nt));
    B2_TRY(b2_upload(p->ctx, p->tabns, tab.data(), nt));
    const size_t smem = (size_t)ns_cta_doubles_host(nve_v, nve_p, ngauss) * sizeof(double);
    B2_CHECK(smem <= 227 * 1024, "b2_ns_create: %zu bytes of shared memory per element exceed the SM", smem);
    B2_CHECK(ns_threads(nve_v, nve_p) <= kNsMaxThreads, "b2_ns_create: %d threads per element exceed the kernel's bound", ns_threads(nve_v, nve_p));
    B2_TRY(build_slot_map(p, true));
    return 0;
  }();
  if (rc) {
    b2_stokes_destroy(p);
    *out = nullptr;
  }
  return rc;
}

/* A += the exact Newton Jacobian, rhs += RES = -aRes of the steady Navier-Stokes residual at the current solution
 * (03_navier_stokes.hpp:305-413); neither is zeroed */
int b2_ns_assemble(b2_stokes* p, const b2_vec* sol, b2_vec* rhs, double nu) {
  B2_CHECK(p && p->tabns, "b2_ns_assemble: the plan was not created by b2_ns_create");
  B2_CHECK((!sol || sol->n >= p->A->nrows) && (!rhs || rhs->n >= p->A->nrows), "b2_ns_assemble: vector shorter than the system");
  b2_ctx* c = nullptr;
  int64_t nnode = 0, nel = 0;
  const double* xyz = nullptr;
  const int32_t* conn = nullptr;
  b2_mesh_view(p->mesh, &c, &nnode, &nel, &xyz, &conn);
  const size_t smem = (size_t)ns_cta_doubles_host(p->nv, p->np, p->ng) * sizeof(double);
  const int threads = ns_threads(p->nv, p->np);
  const ns_kernel_t kern = ns_kernel_for(p->nv, p->np);
  B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
  int per_sm = 1;      // resident CTAs by registers and shared memory
  B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  per_sm = per_sm < 1 ? 1 : per_sm;
  B2_LAUNCH(c, kern, b2_grid_for(c, nel, 1, per_sm), threads, smem, nel, nnode, p->nv, p->np, p->ng, xyz, conn, p->edof, p->tabns,
            p->tabp, p->A->rowptr, p->slot_ns, p->A->val, sol ? sol->d : nullptr, rhs ? rhs->d : nullptr, nu);
  return 0;
}

/* rhs += the boundary pressure term of the Navier-Stokes residual over the listed faces of ONE face kind (tables as in
 * b2_asm_neumann_faces): RES[U_k dof of face node i] -= int phi_i tau n_k (03_navier_stokes.hpp:196-300) */
int b2_ns_pressure_faces(b2_stokes* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value, int nvf,
                         int ngf, const double* phi, const double* dxi, const double* deta, const double* weights, const int32_t* face_nodes,
                         b2_vec* rhs) {
  B2_CHECK(p && rhs && phi && dxi && deta && weights && face_nodes && (nfaces == 0 || (face_elem && face_local && face_value)),
           "b2_ns_pressure_faces: null argument");
  B2_CHECK(nvf >= 1 && nvf <= 9 && ngf >= 1 && ngf <= 16, "b2_ns_pressure_faces: nvf=%d (1..9) or ngf=%d (1..16) out of range", nvf, ngf);
  B2_CHECK(rhs->n >= p->A->nrows, "b2_ns_pressure_faces: rhs vector too short");
  if (nfaces == 0) return 0;
  b2_ctx* c = nullptr;
  int64_t nnode = 0, nel = 0;
  const double* xyz = nullptr;
  const int32_t* conn = nullptr;
  b2_mesh_view(p->mesh, &c, &nnode, &nel, &xyz, &conn);
  for (int64_t k = 0; k < nfaces; k++) {
    B2_CHECK(face_elem[k] >= 0 && face_elem[k] < nel && face_local[k] >= 0 && face_local[k] < 6, "b2_ns_pressure_faces: face %lld out of range", (long long)k);
    for (int i = 0; i < nvf; i++) {
      const int loc = face_nodes[face_local[k] * 9 + i];
      B2_CHECK(loc >= 0 && loc < p->nv, "b2_ns_pressure_faces: face %lld: face dof %d is local node %d, not one of the %d velocity nodes", (long long)k, i, loc, p->nv);
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
  B2_LAUNCH(c, pressure_face_kernel, b2_grid_for(c, nfaces * 32, 256, 8), 256, 0, nfaces, d_e, d_f, d_v, nvf, ngf, d_t, d_fn, nnode, xyz, conn,
            p->edof, rhs->d);
  B2_CUDA(cudaStreamSynchronize(c->stream));
  b2_free(c, d_e, (size_t)nfaces);
  b2_free(c, d_f, (size_t)nfaces);
  b2_free(c, d_v, (size_t)nfaces);
  b2_free(c, d_t, nt);
  b2_free(c, d_fn, 54);
  return 0;
}

int b2_stokes_destroy(b2_stokes* p) {
  if (!p) return 0;
  b2_free(p->ctx, p->tabns, (size_t)4 * p->ng * p->nv + p->ng);
  b2_free(p->ctx, p->slot, (size_t)p->nel * stokes_slots_per_element(p->nv, p->np));
  b2_free(p->ctx, p->slot_ns, (size_t)p->nel * ns_slots_per_element(p->nv, p->np));
  b2_free(p->ctx, p->edof, (size_t)p->nel * 108);
  b2_free(p->ctx, p->tabv, (size_t)3 * p->ng * p->nv + p->ng);
  b2_free(p->ctx, p->tabp, (size_t)p->ng * p->np);
  delete p;
  return 0;
}

}  // extern "C"
