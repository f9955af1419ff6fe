// TEST INFRASTRUCTURE: the device sources femus_b200/csrc/b2_schwarz_kernels.cuh and b2_neumann_kernel.cuh compiled
// for the CPU thread emulator (cuda_emu.hpp) behind two C entry points that mirror what b2_schwarz_setup /
// b2_schwarz_apply and b2_asm_neumann_faces launch.  Built and driven by tests/test_kernel_emulation.py.
#include "cuda_emu.hpp"
#include "../../femus_b200/csrc/b2_schwarz_levels.hpp"
#include "../../femus_b200/csrc/b2_gmres.hpp"

namespace {
#include "../../femus_b200/csrc/b2_schwarz_kernels.cuh"
#include "../../femus_b200/csrc/b2_schwarz_walk.cuh"
}
#include "../../femus_b200/csrc/b2_neumann_kernel.cuh"
namespace {
#include "../../femus_b200/csrc/b2_stokes_kernel.cuh"
#include "../../femus_b200/csrc/b2_ns_kernel.cuh"
}

// host-vector operations for b2_gmres_cycle: CSR operator, Jacobi preconditioner (what gmres_level_ops does on device vectors)
struct host_gmres_ops {
  int64_t n;
  const int64_t* rowptr;
  const int32_t* col;
  const double* val;
  const double* dinv;
  const double* b;
  double* x;
  bool zero_guess;
  std::vector<std::vector<double>> v;
  std::vector<double> w, t;
  void spmv(const double* in, double* out) const {
    for (int64_t i = 0; i < n; i++) {
      double s = 0.0;
      for (int64_t q = rowptr[i]; q < rowptr[i + 1]; q++) s += val[q] * in[col[q]];
      out[i] = s;
    }
  }
  int start(double* beta) {
    if (zero_guess) t.assign(b, b + n);
    else {
      spmv(x, t.data());
      for (int64_t i = 0; i < n; i++) t[i] = b[i] - t[i];
    }
    for (int64_t i = 0; i < n; i++) v[0][i] = dinv[i] * t[i];
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += v[0][i] * v[0][i];
    *beta = std::sqrt(s);
    return 0;
  }
  int scale(int j, double a) { for (double& e : v[j]) e *= a; return 0; }
  int apply(int j) {
    spmv(v[j].data(), t.data());
    for (int64_t i = 0; i < n; i++) w[i] = dinv[i] * t[i];
    return 0;
  }
  int dot_w(int i, double* h) { double s = 0.0; for (int64_t q = 0; q < n; q++) s += w[q] * v[i][q]; *h = s; return 0; }
  int axpy_w(double a, int i) { for (int64_t q = 0; q < n; q++) w[q] += a * v[i][q]; return 0; }
  int norm_w(double* nn) { double s = 0.0; for (double e : w) s += e * e; *nn = std::sqrt(s); return 0; }
  int store(int j) { v[j] = w; return 0; }
  int update_x(double a, int i) { for (int64_t q = 0; q < n; q++) x[q] += a * v[i][q]; return 0; }
};

extern "C" {

// the library's GMRES cycle (b2_gmres.hpp) on host vectors: k iterations from x (zero_guess: from 0)
int emu_gmres(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, const double* dinv, const double* b, double* x, int k,
              int zero_guess) {
  if (zero_guess)
    for (int64_t i = 0; i < n; i++) x[i] = 0.0;
  host_gmres_ops ops{n, rowptr, col, val, dinv, b, x, zero_guess != 0, std::vector<std::vector<double>>(k > 0 ? k : 1, std::vector<double>((size_t)n)),
                     std::vector<double>((size_t)n), std::vector<double>((size_t)n)};
  return b2_gmres_cycle(ops, k);
}

// extract + invert + the sweep over the schedule's groups; returns the singular-block flag of the invert kernel
int emu_schwarz(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, int64_t nblocks, const int64_t* blk_ptr,
                const int32_t* blk_dofs, int64_t ngroups, const int64_t* group_ptr, const int32_t* group_blocks, const double* r,
                double* y, double* inv_out, int threads, int grid, int sub) {
  // sub >= 20: the staged row walk (b2_schwarz_walk.cuh: local indices, block vectors in shared memory, map-based ILU)
  if (sub >= 20) {
    const bool ilu = sub % 10 == 2;
    const int64_t nd = blk_ptr[nblocks];
    std::vector<int64_t> frow((size_t)nd + 1);
    int64_t tot = 0;
    int max_m = 0;
    for (int64_t k = 0; k < nd; k++) { frow[k] = tot; tot += rowptr[blk_dofs[k] + 1] - rowptr[blk_dofs[k]]; }
    frow[(size_t)nd] = tot;
    for (int64_t b = 0; b < nblocks; b++) max_m = std::max<int>(max_m, (int)(blk_ptr[b + 1] - blk_ptr[b]));
    std::vector<unsigned short> lidx((size_t)tot, 0x7777);
    std::vector<double> fac(ilu ? (size_t)tot : 1, -7.0);
    emu::launch(schwarz_lidx_kernel, (unsigned)grid, (unsigned)threads, 0, nblocks, blk_ptr, blk_dofs, (const int64_t*)frow.data(), rowptr, col,
                lidx.data());
    const size_t smem = walk_smem_bytes(max_m);
    int err = 0;
    if (ilu)
      for (int64_t g = 0; g < ngroups; g++)
        emu::launch(schwarz_walk_ilu_factor_kernel, (unsigned)grid, (unsigned)threads, smem, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                    blk_dofs, (const int64_t*)frow.data(), (const unsigned short*)lidx.data(), rowptr, val, fac.data(), &err, max_m);
    if (err) return err;
    std::vector<double> y2((size_t)n, 0.0);
    for (int pass = 0; pass < 2; pass++) {          // twice: nothing may survive in the shared arrays between applications
      double* yy = pass ? y2.data() : y;
      for (int64_t i = 0; i < n; i++) yy[i] = 0.0;
      for (int64_t g = 0; g < ngroups; g++) {
        int max_row = 0;
        for (int64_t i = 0; i < n; i++) max_row = std::max<int>(max_row, (int)(rowptr[i + 1] - rowptr[i]));
        const walk_apply_kernel_t kern = ilu ? schwarz_walk_apply_kernel_for<true>(max_row) : schwarz_walk_apply_kernel_for<false>(max_row);
        emu::launch(kern, (unsigned)grid, (unsigned)threads, smem, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr, blk_dofs,
                    (const int64_t*)frow.data(), (const unsigned short*)lidx.data(), rowptr, col, val, ilu ? (const double*)fac.data() : (const double*)nullptr,
                    r, yy, max_m);
      }
    }
    for (int64_t i = 0; i < n; i++)
      if (y2[i] != y[i]) return -1;
    return 0;
  }
  // sub >= 10: the same block solve with the rows of every block sorted into dependency levels (LEV kernels)
  const bool lev = sub >= 10;
  sub %= 10;
  std::vector<int64_t> lptr[2];
  std::vector<int32_t> loff[2], lrows[2];
  if (lev) b2_schwarz_row_level_schedule(nblocks, blk_ptr, blk_dofs, rowptr, col, lptr, loff, lrows);
  const int64_t *pf = lev ? lptr[0].data() : nullptr, *pb = lev ? lptr[1].data() : nullptr;
  const int32_t *of = lev ? loff[0].data() : nullptr, *ob = lev ? loff[1].data() : nullptr;
  const int32_t *rf = lev ? lrows[0].data() : nullptr, *rb = lev ? lrows[1].data() : nullptr;
  if (sub == 1) {       // SSOR block solves: scratch vectors only (b2_schwarz_setup with kind 1)
    std::vector<double> tg((size_t)n, -7.0), dg((size_t)n, -7.0), zg((size_t)n, -7.0);
    std::vector<int32_t> mark((size_t)n, -1);
    std::vector<double> y2((size_t)n, 0.0);
    for (int pass = 0; pass < 2; pass++) {          // twice on the used scratch (stale marks of overlapping blocks)
      double* yy = pass ? y2.data() : y;
      for (int64_t i = 0; i < n; i++) yy[i] = 0.0;
      for (int64_t g = 0; g < ngroups; g++) {
        if (lev)
          emu::launch(schwarz_apply_ssor_kernel<true>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                      blk_dofs, rowptr, col, val, r, yy, tg.data(), dg.data(), zg.data(), mark.data(), pf, of, rf, pb, ob, rb);
        else
          emu::launch(schwarz_apply_ssor_kernel<false>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                      blk_dofs, rowptr, col, val, r, yy, tg.data(), dg.data(), zg.data(), mark.data(), pf, of, rf, pb, ob, rb);
      }
    }
    for (int64_t i = 0; i < n; i++)
      if (y2[i] != y[i]) return -1;
    return 0;
  }
  if (sub == 2) {       // ILU(0) block solves: factor group by group, then the sweep (twice, on the used scratch)
    std::vector<int64_t> frow((size_t)blk_ptr[nblocks]);
    int64_t tot = 0;
    for (int64_t k = 0; k < blk_ptr[nblocks]; k++) { frow[k] = tot; tot += rowptr[blk_dofs[k] + 1] - rowptr[blk_dofs[k]]; }
    std::vector<double> fac((size_t)tot, -7.0), zg((size_t)n, -7.0);
    std::vector<int64_t> foff((size_t)n, -1);
    std::vector<int32_t> mark((size_t)n, -1);
    int err = 0;
    for (int64_t g = 0; g < ngroups; g++) {
      if (lev)
        emu::launch(schwarz_ilu_factor_kernel<true>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                    blk_dofs, (const int64_t*)frow.data(), rowptr, col, val, fac.data(), mark.data(), foff.data(), &err, pf, of, rf);
      else
        emu::launch(schwarz_ilu_factor_kernel<false>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                    blk_dofs, (const int64_t*)frow.data(), rowptr, col, val, fac.data(), mark.data(), foff.data(), &err, pf, of, rf);
    }
    if (err) return err;
    std::vector<double> y2((size_t)n, 0.0);
    for (int pass = 0; pass < 2; pass++) {
      double* yy = pass ? y2.data() : y;
      for (int64_t i = 0; i < n; i++) yy[i] = 0.0;
      for (int64_t g = 0; g < ngroups; g++) {
        if (lev)
          emu::launch(schwarz_apply_ilu_kernel<true>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                      blk_dofs, (const int64_t*)frow.data(), rowptr, col, val, (const double*)fac.data(), r, yy, zg.data(), mark.data(), pf, of, rf, pb,
                      ob, rb);
        else
          emu::launch(schwarz_apply_ilu_kernel<false>, (unsigned)grid, (unsigned)threads, 0, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr,
                      blk_dofs, (const int64_t*)frow.data(), rowptr, col, val, (const double*)fac.data(), r, yy, zg.data(), mark.data(), pf, of, rf, pb,
                      ob, rb);
      }
    }
    for (int64_t i = 0; i < n; i++)
      if (y2[i] != y[i]) return -1;
    return 0;
  }
  std::vector<int64_t> inv_ptr((size_t)nblocks + 1, 0);
  int max_m = 0;
  for (int64_t b = 0; b < nblocks; b++) {
    const int64_t m = blk_ptr[b + 1] - blk_ptr[b];
    inv_ptr[b + 1] = inv_ptr[b] + m * m;
    if (m > max_m) max_m = (int)m;
  }
  std::vector<double> inv((size_t)inv_ptr[nblocks], -7.0);          // poisoned: the extract kernel must write every entry
  int err = 0;
  const size_t smem = 2 * (size_t)max_m * sizeof(double);
  emu::launch(schwarz_extract_kernel, (unsigned)grid, (unsigned)threads, 0, nblocks, blk_ptr, blk_dofs, inv_ptr.data(), rowptr, col, val,
              inv.data());
  emu::launch(schwarz_invert_kernel, (unsigned)grid, (unsigned)threads, smem + (size_t)max_m * sizeof(int), nblocks, blk_ptr,
              (const int64_t*)inv_ptr.data(), inv.data(), max_m, &err);
  for (int64_t i = 0; i < n; i++) y[i] = 0.0;
  for (int64_t g = 0; g < ngroups; g++)
    emu::launch(schwarz_apply_kernel, (unsigned)grid, (unsigned)threads, smem, group_ptr[g], group_ptr[g + 1], group_blocks, blk_ptr, blk_dofs,
                (const int64_t*)inv_ptr.data(), (const double*)inv.data(), rowptr, col, val, r, y, max_m);
  if (inv_out) std::copy(inv.begin(), inv.end(), inv_out);
  return err;
}

void emu_neumann(int64_t nfaces, const int32_t* felem, const int32_t* flocal, const double* fvalue, int nvf, int ngf, int nve,
                 const double* ftab, const int32_t* fnodes, int64_t nnode, const double* xyz, const int32_t* conn, const int32_t* dof,
                 double* rhs, int grid) {
  emu::launch(neumann_kernel, (unsigned)grid, 256u, 0, nfaces, felem, flocal, fvalue, nvf, ngf, nve, ftab, fnodes, nnode, xyz, conn, dof, rhs);
}

// what b2_stokes_assemble launches: tabv = dxi, deta, dzeta [ng][nv], w[ng]; tabp = phi [ng][np]; edof [nel][4][27]
void emu_stokes(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* xyz, const int32_t* conn, const int32_t* edof,
                const double* tabv, const double* tabp, const int64_t* rowptr, const int32_t* col, double* Aval, const double* sol, double* rhs,
                double IRe, int grid) {
  const size_t smem = (size_t)kStokesWarps * (size_t)stokes_warp_doubles_host(nv, np, ng) * sizeof(double);
  std::vector<unsigned short> slot((size_t)nel * stokes_slots_per_element(nv, np));
  int err = 0;
  emu::launch(stokes_slot_kernel, 2u, 64u, 0, nel, nv, np, edof, rowptr, col, slot.data(), &err);
  if (err) std::abort();
  // grid < 0: the guarded instantiation (what any other (nv, np) pair runs) on -grid CTAs
  const stokes_kernel_t kern = grid < 0 ? stokes_kernel<27, 8, true> : stokes_kernel_for(nv, np);
  if (grid < 0) grid = -grid;
  emu::launch(kern, (unsigned)grid, (unsigned)(kStokesWarps * 32), smem, nel, nnode, nv, np, ng, xyz, conn, edof, tabv, tabp, rowptr,
              (const unsigned short*)slot.data(), Aval, sol, rhs, IRe);
}

// what b2_ns_assemble launches: tabv = phi, dxi, deta, dzeta [ng][nv], w[ng]
void emu_ns(int64_t nel, int64_t nnode, int nv, int np, int ng, const double* xyz, const int32_t* conn, const int32_t* edof, const double* tabv,
            const double* tabp, const int64_t* rowptr, const int32_t* col, double* Aval, const double* sol, double* rhs, double nu, int grid) {
  const size_t smem = (size_t)ns_cta_doubles_host(nv, np, ng) * sizeof(double);
  std::vector<unsigned short> slot((size_t)nel * ns_slots_per_element(nv, np));
  int err = 0;
  emu::launch(ns_slot_kernel, 2u, 64u, 0, nel, nv, np, edof, rowptr, col, slot.data(), &err);
  if (err) std::abort();
  const ns_kernel_t kern = grid < 0 ? ns_kernel<0, 0> : ns_kernel_for(nv, np);      // grid < 0: the guarded instantiation
  if (grid < 0) grid = -grid;
  emu::launch(kern, (unsigned)grid, (unsigned)ns_threads(nv, np), smem, nel, nnode, nv, np, ng, xyz, conn, edof, tabv, tabp, rowptr,
              (const unsigned short*)slot.data(), Aval, sol, rhs, nu);
}

void emu_pressure_faces(int64_t nfaces, const int32_t* felem, const int32_t* flocal, const double* fvalue, int nvf, int ngf, const double* ftab,
                        const int32_t* fnodes, int64_t nnode, const double* xyz, const int32_t* conn, const int32_t* edof, double* rhs, int grid) {
  emu::launch(pressure_face_kernel, (unsigned)grid, 256u, 0, nfaces, felem, flocal, fvalue, nvf, ngf, ftab, fnodes, nnode, xyz, conn, edof, rhs);
}

}  // extern "C"
