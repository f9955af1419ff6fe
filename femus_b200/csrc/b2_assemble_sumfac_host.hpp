// Host-only helpers of the sum-factorised assembly kernel (b2_assemble_sumfac.cuh): they FACTOR the tables the caller
// hands to b2_asm_create / the Galerkin plan into their 1-D pieces and check the result, so the kernel is used only
// where the tensor-product structure really holds.  Shared by b2_assemble.cu and the CPU emulator harness of the tests.
// Included after b2_assemble_sumfac.cuh, inside the same namespace.
#pragma once
// (no #include here: the file is included inside a namespace; the includer provides <cmath> and <cstring>)

// Lattice position of the 27 local nodes in the reference's order (Hexahedron.cpp:32-37: 8 vertices, 12 edge
// midpoints, 6 face centres, centre), coordinates in {-1, 0, 1}.
static const signed char kHex27Lattice[27][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1}, {1, 0, -1}, {0, 1, -1}, {-1, 0, -1}, {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1},
    {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
    {0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, -1}, {0, 0, 1}, {0, 0, 0}};
static inline int sf_lattice_of(int node) {       // m = 9 i1 + 3 i2 + i3
  return 9 * (kHex27Lattice[node][0] + 1) + 3 * (kHex27Lattice[node][1] + 1) + (kHex27Lattice[node][2] + 1);
}

// 1-D factors of the caller's tables ([64][27] row-major, Gauss point g = 16 a + 4 b + c): with the partition of unity
// of the 1-D bases, l_i(p_a) is the sum of phi over the nodes with first lattice index i at any Gauss point with first
// index a (likewise l' from dphi/dxi).  The tables are then REBUILT from the factors and compared: false if they
// differ (another quadrature rule or node order: the kernel is not used).
static inline bool sf_factor_tables(const double* phi, const double* dxi, const double* deta, const double* dzeta, const double* w, SfTables* T) {
  memset(T, 0, sizeof(*T));
  int node_at[27];
  for (int n = 0; n < 27; n++) node_at[sf_lattice_of(n)] = n;
  for (int m = 0; m < 27; m++) T->node_of[m] = node_at[m];
  for (int i = 0; i < 3; i++)
    for (int a = 0; a < 4; a++) {
      double l = 0.0, d = 0.0;
      for (int r = 0; r < 9; r++) {
        l += phi[(16 * a) * 27 + node_at[9 * i + r]];
        d += dxi[(16 * a) * 27 + node_at[9 * i + r]];
      }
      T->L[i][a] = l;
      T->D[i][a] = d;
    }
  for (int g = 0; g < 64; g++) {
    const int a = g >> 4, b = (g >> 2) & 3, c = g & 3;
    for (int m = 0; m < 27; m++) {
      const int i1 = m / 9, i2 = (m / 3) % 3, i3 = m % 3, n = node_at[m];
      const double want[4] = {T->L[i1][a] * T->L[i2][b] * T->L[i3][c], T->D[i1][a] * T->L[i2][b] * T->L[i3][c],
                              T->L[i1][a] * T->D[i2][b] * T->L[i3][c], T->L[i1][a] * T->L[i2][b] * T->D[i3][c]};
      const double have[4] = {phi[g * 27 + n], dxi[g * 27 + n], deta[g * 27 + n], dzeta[g * 27 + n]};
      for (int k = 0; k < 4; k++)
        if (!(fabs(want[k] - have[k]) <= 1e-14)) return false;
    }
    T->w[g] = w[g];
  }
  for (int p = 0; p < 2; p++)
    for (int q = 0; q < 2; q++)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          for (int a = 0; a < 4; a++) T->M[2 * p + q][3 * i + j][a] = (p ? T->D[i][a] : T->L[i][a]) * (q ? T->D[j][a] : T->L[j][a]);
  return true;
}

// Kronecker factors of the child prolongators Pc[child][fine local node][coarse local node] (both in the reference's
// local order): rows of every 1-D factor sum to one, so A_d[n][J] is the sum of row (n at dimension d, 0 elsewhere)
// over the coarse indices of the other two dimensions; the product is rebuilt and compared.  false: not products.
static inline bool sf_factor_children(const double* Pc, SfGalTables* G) {
  memset(G, 0, sizeof(*G));
  int node_at[27];
  for (int n = 0; n < 27; n++) node_at[sf_lattice_of(n)] = n;
  const int stride[3] = {9, 3, 1};
  for (int j = 0; j < 8; j++) {
    auto P = [&](int m, int M) { return Pc[((size_t)j * 27 + node_at[m]) * 27 + node_at[M]]; };      // lattice order
    for (int d = 0; d < 3; d++)
      for (int n = 0; n < 3; n++)
        for (int J = 0; J < 3; J++) {
          double s = 0.0;
          for (int M = 0; M < 27; M++)
            if ((M / stride[d]) % 3 == J) s += P(n * stride[d], M);
          G->A[j][d][n][J] = s;
        }
    for (int m = 0; m < 27; m++)
      for (int M = 0; M < 27; M++) {
        const double want = G->A[j][0][m / 9][M / 9] * G->A[j][1][(m / 3) % 3][(M / 3) % 3] * G->A[j][2][m % 3][M % 3];
        if (!(fabs(want - P(m, M)) <= 1e-15)) return false;
      }
  }
  // every factor must be the lower- or upper-child form of SfGalTables, with the same (a, b, c)
  bool have = false;
  for (int j = 0; j < 8; j++)
    for (int d = 0; d < 3; d++) {
      const double (*A)[3] = G->A[j][d];
      const bool lo = A[0][0] == 1.0 && A[0][1] == 0.0 && A[0][2] == 0.0 && A[2][0] == 0.0 && A[2][1] == 1.0 && A[2][2] == 0.0;
      const bool up = A[0][0] == 0.0 && A[0][1] == 1.0 && A[0][2] == 0.0 && A[2][0] == 0.0 && A[2][1] == 0.0 && A[2][2] == 1.0;
      if (!lo && !up) return false;
      const double abc[3] = {lo ? A[1][0] : A[1][2], A[1][1], lo ? A[1][2] : A[1][0]};
      if (!have) { G->abc[0] = abc[0]; G->abc[1] = abc[1]; G->abc[2] = abc[2]; have = true; }
      else if (abc[0] != G->abc[0] || abc[1] != G->abc[1] || abc[2] != G->abc[2]) return false;
      G->hi[j][d] = up ? 1 : 0;
    }
  for (int I = 0; I < 27; I++)
    for (int J = 0; J < 27; J++) G->nat2lat[I * 27 + J] = (unsigned short)(sf_lattice_of(I) * 27 + sf_lattice_of(J));
  return true;
}
