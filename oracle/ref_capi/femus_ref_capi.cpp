// oracle/_ref: C entry points around the UNMODIFIED reference FE kernel
// (/root/reference/src/02_reference_geom_elements, compiled where it lies).
// TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Never linked by the product.
//
// The element loop in fref_poisson_* restates the body of
// applications/001_Poisson/main.cpp:350-602 (3-D branch: nu=1, V=0, supgTau=0)
// and calls the reference's own elem_type_3D::Jacobian (ElemType.hpp:1438-1537).
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ElemType.hpp"

using femus::elem_type_3D;
using femus::elem_type_2D;

extern "C" {

void* fref_create(const char* geom, const char* order, const char* gauss) {
  return new elem_type_3D(geom, order, gauss);
}
void fref_destroy(void* h) { delete static_cast<elem_type_3D*>(h); }
int fref_ndofs(void* h) { return static_cast<elem_type_3D*>(h)->GetNDofs(); }
int fref_ngauss(void* h) { return (int)static_cast<elem_type_3D*>(h)->GetGaussPointNumber(); }
int fref_ndofs_fine(void* h) { return static_cast<elem_type_3D*>(h)->GetNDofsFine(); }

// quadrature: w[ng], xi[3][ng]
void fref_gauss(void* h, double* w, double* xi) {
  auto* e = static_cast<elem_type_3D*>(h);
  const int ng = (int)e->GetGaussPointNumber();
  const femus::Gauss* g = e->GetGaussRule();
  for (int i = 0; i < ng; i++) {
    w[i] = g->GetGaussWeightsPointer()[i];
    for (int d = 0; d < 3; d++) xi[d * ng + i] = g->GetGaussCoordinatePointer(d)[i];
  }
}
// tables [ng][ndofs] row-major
void fref_tables(void* h, double* phi, double* dxi, double* deta, double* dzeta) {
  auto* e = static_cast<elem_type_3D*>(h);
  const int ng = (int)e->GetGaussPointNumber(), n = e->GetNDofs();
  for (int g = 0; g < ng; g++)
    for (int i = 0; i < n; i++) {
      phi[g * n + i] = e->GetPhi(g)[i];
      dxi[g * n + i] = e->GetDPhiDXi(g)[i];
      deta[g * n + i] = e->GetDPhiDEta(g)[i];
      dzeta[g * n + i] = e->GetDPhiDZeta(g)[i];
    }
}
// coords: [3][n] (only the first ndofs are read by the reference)
void fref_jacobian(void* h, const double* coords, int ncoord, int ig, double* weight,
                   double* phi, double* gradphi, double* nablaphi) {
  auto* e = static_cast<elem_type_3D*>(h);
  std::vector<std::vector<double>> vt(3, std::vector<double>(ncoord));
  for (int d = 0; d < 3; d++) for (int i = 0; i < ncoord; i++) vt[d][i] = coords[d * ncoord + i];
  std::vector<double> p, g, nb;
  double w;
  e->Jacobian(vt, (unsigned)ig, w, p, g, nb);
  *weight = w;
  std::copy(p.begin(), p.end(), phi);
  std::copy(g.begin(), g.end(), gradphi);
  if (nablaphi) std::copy(nb.begin(), nb.end(), nablaphi);
}
// local prolongator row i (fine dof i of a refined element)
int fref_prol_row(void* h, int i, int* idx, double* val, int* child, int* node) {
  auto* e = static_cast<elem_type_3D*>(h);
  const int nc = e->Get_Prolongator_Num_Columns(i);
  for (int k = 0; k < nc; k++) { idx[k] = e->Get_Prolongator_Index(i, k); val[k] = e->Get_Prolongator_Value(i, k); }
  auto kv = e->GetKVERT_IND(i);
  *child = kv.first; *node = kv.second;
  return nc;
}

// One element of the Poisson loop.  xg: [3][nve] element coordinates of the unknown's
// own nodes, sol[nve], source f (constant).  Out: F[nve], B[nve*nve] row-major.
static void poisson_element(const elem_type_3D* e, int nve, std::vector<std::vector<double>>& vt,
                            const double* sol, double fsrc, double* F, double* B,
                            std::vector<double>& phi, std::vector<double>& gradphi,
                            std::vector<double>& nablaphi) {
  const int dim = 3;
  const double nu = 1.;
  const double supgTau = 0.;  // V = 0 in 3-D (main.cpp:406-428)
  const double V[3] = {0., 0., 0.};
  const int ng = (int)e->GetGaussPointNumber();
  std::fill(F, F + nve, 0.);
  std::fill(B, B + nve * nve, 0.);
  double weight;
  for (int ig = 0; ig < ng; ig++) {
    e->Jacobian(vt, (unsigned)ig, weight, phi, gradphi, nablaphi);
    double gradSolT[3] = {0, 0, 0}, NablaSolT[3] = {0, 0, 0};
    for (int i = 0; i < nve; i++)
      for (int d = 0; d < dim; d++) {
        gradSolT[d] += gradphi[i * dim + d] * sol[i];
        NablaSolT[d] += nablaphi[i * 6 + d] * sol[i];
      }
    for (int i = 0; i < nve; i++) {
      double advRhs = 0., lapRhs = 0., resRhs = 0., supgPhi = 0.;
      for (int d = 0; d < dim; d++) {
        lapRhs += nu * gradphi[i * dim + d] * gradSolT[d];
        advRhs += V[d] * gradSolT[d] * phi[i];
        resRhs += -nu * NablaSolT[d] + V[d] * gradSolT[d];
        supgPhi += (V[d] * gradphi[i * dim + d] + nu * nablaphi[i * 6 + d]) * supgTau;
      }
      F[i] += (fsrc * phi[i] - lapRhs - advRhs + (fsrc - resRhs) * supgPhi) * weight;
      for (int j = 0; j < nve; j++) {
        double lap = 0, adv = 0;
        for (int d = 0; d < dim; d++) {
          lap += nu * (gradphi[i * dim + d] * gradphi[j * dim + d] - nablaphi[j * 6 + d] * supgPhi) * weight;
          adv += V[d] * gradphi[j * dim + d] * (phi[i] + supgPhi) * weight;
        }
        B[i * nve + j] += lap + adv;
      }
    }
  }
}

void fref_poisson_element(void* h, const double* xg, const double* sol, double fsrc, double* F, double* B) {
  auto* e = static_cast<elem_type_3D*>(h);
  const int nve = e->GetNDofs();
  std::vector<std::vector<double>> vt(3, std::vector<double>(nve));
  for (int d = 0; d < 3; d++) for (int i = 0; i < nve; i++) vt[d][i] = xg[d * nve + i];
  std::vector<double> phi, gradphi, nablaphi;
  poisson_element(e, nve, vt, sol, fsrc, F, B, phi, gradphi, nablaphi);
}

// Whole-mesh assembly into a preallocated CSR (sorted columns), elements [e0,e1).
//   conn  [nel][27] node ids (coordinates), dof [nel][nve] matrix rows/cols of the unknown,
//   xyz   [3][nnode], sol [ndof].  vals/rhs must be zeroed by the caller.
// Returns seconds spent in the element loop.  nthreads<=1 -> serial.
double fref_poisson_assemble_csr(void* h, long e0, long e1, const int* conn, const int* dof,
                                 const double* xyz, long nnode, const double* sol,
                                 const long long* rowptr, const int* col, double* vals, double* rhs,
                                 double fsrc, int nthreads) {
  auto* e = static_cast<elem_type_3D*>(h);
  const int nve = e->GetNDofs();
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
  double t0 = omp_get_wtime();
#pragma omp parallel num_threads(nthreads)
#else
  timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
#endif
  {
    std::vector<std::vector<double>> vt(3, std::vector<double>(nve));
    std::vector<double> phi, gradphi, nablaphi, F(nve), B(nve * nve), u(nve);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (long iel = e0; iel < e1; iel++) {
      for (int i = 0; i < nve; i++) {
        const long nd = conn[iel * 27 + i];
        for (int d = 0; d < 3; d++) vt[d][i] = xyz[d * nnode + nd];
        u[i] = sol[dof[iel * nve + i]];
      }
      poisson_element(e, nve, vt, u.data(), fsrc, F.data(), B.data(), phi, gradphi, nablaphi);
      for (int i = 0; i < nve; i++) {
        const int r = dof[iel * nve + i];
#ifdef _OPENMP
#pragma omp atomic
#endif
        rhs[r] += F[i];
        const int* cb = col + rowptr[r];
        const int* ce = col + rowptr[r + 1];
        for (int j = 0; j < nve; j++) {
          const int* p = std::lower_bound(cb, ce, dof[iel * nve + j]);
          double* dst = vals + (p - col);
#ifdef _OPENMP
#pragma omp atomic
#endif
          *dst += B[i * nve + j];
        }
      }
    }
  }
#ifdef _OPENMP
  return omp_get_wtime() - t0;
#else
  timespec ts1; clock_gettime(CLOCK_MONOTONIC, &ts1);
  return (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);
#endif
}

// ---- boundary faces: the reference's own elem_type_2D ("quad", order, gauss) and JacobianSur
//      (ElemType.hpp:1330-1379), as applications/001_Poisson/main.cpp:495-594 calls them for the
//      Neumann integrals
void* fref2_create(const char* geom, const char* order, const char* gauss) { return new elem_type_2D(geom, order, gauss); }
void fref2_destroy(void* h) { delete static_cast<elem_type_2D*>(h); }
int fref2_ndofs(void* h) { return static_cast<elem_type_2D*>(h)->GetNDofs(); }
int fref2_ngauss(void* h) { return (int)static_cast<elem_type_2D*>(h)->GetGaussPointNumber(); }
void fref2_gauss(void* h, double* w, double* xi) {
  auto* e = static_cast<elem_type_2D*>(h);
  const int ng = (int)e->GetGaussPointNumber();
  const femus::Gauss* g = e->GetGaussRule();
  for (int i = 0; i < ng; i++) {
    w[i] = g->GetGaussWeightsPointer()[i];
    for (int d = 0; d < 2; d++) xi[d * ng + i] = g->GetGaussCoordinatePointer(d)[i];
  }
}
void fref2_tables(void* h, double* phi, double* dxi, double* deta) {
  auto* e = static_cast<elem_type_2D*>(h);
  const int ng = (int)e->GetGaussPointNumber(), n = e->GetNDofs();
  for (int g = 0; g < ng; g++)
    for (int i = 0; i < n; i++) {
      phi[g * n + i] = e->GetPhi(g)[i];
      dxi[g * n + i] = e->GetDPhiDXi(g)[i];
      deta[g * n + i] = e->GetDPhiDEta(g)[i];
    }
}
// coords: [3][n]; out: weight, phi[ndofs], normal[3]
void fref2_jacobian_sur(void* h, const double* coords, int ncoord, int ig, double* weight, double* phi, double* normal) {
  auto* e = static_cast<elem_type_2D*>(h);
  std::vector<std::vector<double>> vt(3, std::vector<double>(ncoord));
  for (int d = 0; d < 3; d++) for (int i = 0; i < ncoord; i++) vt[d][i] = coords[d * ncoord + i];
  std::vector<double> p, g, nrm;
  double w;
  e->JacobianSur(vt, (unsigned)ig, w, p, g, nrm);
  *weight = w;
  std::copy(p.begin(), p.end(), phi);
  std::copy(nrm.begin(), nrm.end(), normal);
}

}  // extern "C"
