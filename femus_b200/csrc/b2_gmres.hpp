// Host-only: the left-preconditioned GMRES cycle of the level solver (KSPGMRES restated, b2_mg.cu: smooth_gmres),
// written against a small vector-operation interface so that the SAME code runs on device vectors in the library and
// on host vectors in the CPU tests (tests/cpp/emu_kernels.cpp: emu_gmres).  k iterations from the current iterate, no
// restart: modified Gram-Schmidt, Givens rotations, x <- x + V y with y from the triangularised Hessenberg system.
//
// Ops must provide (every call returns 0 on success):
//   int start(double* beta)          v_0 <- M^-1 (b - A x);  *beta = ||v_0||
//   int scale(int j, double a)       v_j <- a v_j
//   int apply(int j)                 w <- M^-1 A v_j
//   int dot_w(int i, double* h)      *h = <w, v_i>
//   int axpy_w(double a, int i)      w <- w + a v_i
//   int norm_w(double* n)            *n = ||w||
//   int store(int j)                 v_j <- w
//   int update_x(double a, int i)    x <- x + a v_i
#pragma once
#include <cmath>
#include <vector>

template <class Ops>
int b2_gmres_cycle(Ops& ops, int k) {
  if (k <= 0) return 0;
  double beta = 0.0;
  if (int s = ops.start(&beta)) return s;
  if (!(beta > 0.0)) return 0;
  if (int s = ops.scale(0, 1.0 / beta)) return s;
  std::vector<double> H((size_t)(k + 1) * k, 0.0), cs(k, 0.0), sn(k, 0.0), g(k + 1, 0.0);
  g[0] = beta;
  int m = 0;
  for (int j = 0; j < k; j++) {
    if (int s = ops.apply(j)) return s;
    for (int i = 0; i <= j; i++) {
      double h = 0.0;
      if (int s = ops.dot_w(i, &h)) return s;
      H[(size_t)i * k + j] = h;
      if (int s = ops.axpy_w(-h, i)) return s;
    }
    double hn = 0.0;
    if (int s = ops.norm_w(&hn)) return s;
    H[(size_t)(j + 1) * k + j] = hn;
    for (int i = 0; i < j; i++) {          // earlier rotations on the new column
      const double a = H[(size_t)i * k + j], b = H[(size_t)(i + 1) * k + j];
      H[(size_t)i * k + j] = cs[i] * a + sn[i] * b;
      H[(size_t)(i + 1) * k + j] = -sn[i] * a + cs[i] * b;
    }
    const double a = H[(size_t)j * k + j], b = H[(size_t)(j + 1) * k + j], d = std::sqrt(a * a + b * b);
    m = j + 1;
    if (!(d > 0.0)) { m = j; break; }
    cs[j] = a / d;
    sn[j] = b / d;
    H[(size_t)j * k + j] = d;
    H[(size_t)(j + 1) * k + j] = 0.0;
    g[j + 1] = -sn[j] * g[j];
    g[j] = cs[j] * g[j];
    if (!(hn > 0.0)) break;                // happy breakdown: the Krylov space is exhausted
    if (j + 1 < k) {                       // the next basis vector (the last one is never used)
      if (int s = ops.store(j + 1)) return s;
      if (int s = ops.scale(j + 1, 1.0 / hn)) return s;
    }
  }
  std::vector<double> y(m, 0.0);
  for (int i = m - 1; i >= 0; i--) {
    double t = g[i];
    for (int j = i + 1; j < m; j++) t -= H[(size_t)i * k + j] * y[j];
    y[i] = t / H[(size_t)i * k + i];
  }
  for (int i = 0; i < m; i++)
    if (int s = ops.update_x(y[i], i)) return s;
  return 0;
}
