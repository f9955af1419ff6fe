// TEST INFRASTRUCTURE: the sum-factorised Hex27 assembly kernel (femus_b200/csrc/b2_assemble_sumfac.cuh) compiled for
// the CPU thread emulator, behind one C entry point that does what b2_asm_create (sf_prepare) / sf_build_gal /
// b2_asm_poisson(_galerkin) do around the launch.  Built and driven by tests/test_kernel_emulation.py.
#include "cuda_emu.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace {
#include "../../femus_b200/csrc/b2_assemble_sumfac.cuh"
#include "../../femus_b200/csrc/b2_assemble_sumfac_host.hpp"

// position of column c inside row r of a CSR pattern (what natural_slot_kernel computes on the device)
int slot_of(const int64_t* rowptr, const int32_t* col, int32_t r, int32_t c) {
  const int32_t* b = col + rowptr[r];
  const int32_t* e = col + rowptr[r + 1];
  const int32_t* it = std::lower_bound(b, e, c);
  return (it != e && *it == c) ? (int)(it - b) : -1;
}
}  // namespace

extern "C" {

// returns 0, or 1 if the tables / child prolongators are not tensor products, 2 if an element couples dofs outside a pattern
int emu_sumfac(int64_t nel, int64_t nnode, const double* xyz, const int32_t* conn, const int32_t* dof, const double* phi, const double* dxi,
               const double* deta, const double* dzeta, const double* w, const int64_t* rowptr, const int32_t* col, double* Aval, const double* u,
               double* rhs, double nu, double fsrc, int gal, const int32_t* cd, const double* Pc, const int64_t* Cp, const int32_t* Ccol, double* Cv,
               const uint8_t* fmask, const uint8_t* cmask, double* emat, int grid) {
  SfTables T;
  if (!sf_factor_tables(phi, dxi, deta, dzeta, w, &T)) return 1;
  std::memcpy(c_sfM, T.M, sizeof(T.M));
  std::memcpy(c_sfU, T.L, sizeof(T.L) + sizeof(T.D));
  std::vector<int32_t> dofL((size_t)nel * 27);
  std::vector<uint16_t> lslot((size_t)nel * kSfSlotStride, 0);
  for (int64_t e = 0; e < nel; e++) {
    for (int m = 0; m < 27; m++) dofL[e * 27 + m] = dof[e * 27 + T.node_of[m]];
    for (int i = 0; i < 27; i++)
      for (int j = 0; j < 27; j++) {
        const int s = slot_of(rowptr, col, dofL[e * 27 + i], dofL[e * 27 + j]);
        if (s < 0) return 2;
        lslot[e * kSfSlotStride + i * 27 + j] = (uint16_t)s;
      }
  }
  SfGalArgs ga = {};
  SfGalTables G;
  std::vector<uint16_t> cslot;
  if (gal) {
    if (!sf_factor_children(Pc, &G)) return 1;
    const int64_t nelc = nel / 8;
    cslot.resize((size_t)nelc * 729);
    for (int64_t E = 0; E < nelc; E++)
      for (int I = 0; I < 27; I++)
        for (int J = 0; J < 27; J++) {
          const int s = slot_of(Cp, Ccol, cd[E * 27 + I], cd[E * 27 + J]);
          if (s < 0) return 2;
          cslot[E * 729 + I * 27 + J] = (uint16_t)s;
        }
    ga = SfGalArgs{&G, cd, cslot.data(), fmask, cmask, Cp, Cv, emat};
    emu::launch(assemble_q2_sumfac_kernel<kSfWarpsDefault, uint16_t, true, uint16_t>, (unsigned)grid, (unsigned)(kSfWarpsDefault * 32), SfSmem<kSfWarpsDefault>::bytes_gal, nel, nnode, xyz, conn,
                (const int32_t*)dofL.data(), (const SfTables*)&T, (const uint16_t*)lslot.data(), rowptr, Aval, u, rhs, nu, fsrc, ga);
  } else {
    emu::launch(assemble_q2_sumfac_kernel<kSfWarpsDefault, uint16_t, false, uint16_t>, (unsigned)grid, (unsigned)(kSfWarpsDefault * 32), SfSmem<kSfWarpsDefault>::bytes, nel, nnode, xyz, conn,
                (const int32_t*)dofL.data(), (const SfTables*)&T, (const uint16_t*)lslot.data(), rowptr, Aval, u, rhs, nu, fsrc, ga);
  }
  return 0;
}

}  // extern "C"
