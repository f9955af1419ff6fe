// TEST INFRASTRUCTURE (oracle/): a plain single-process HOST backend behind the reference's own algebra interfaces --
// NumericVector (src/03_algebra/00_vectors/NumericVector.hpp), SparseMatrix (01_matrices/SparseMatrix.hpp) and
// LinearEquationSolver (src/08_algebra.../03_solvers_with_preconditioner/LinearEquationSolver.hpp) -- so that the
// reference's UNMODIFIED mesh / solution / system sources (Mesh.cpp, MeshRefinement.cpp, MultiLevelSolution.cpp,
// LinearEquation.cpp, LinearImplicitSystem.cpp ...) and its unmodified application applications/001_Poisson/main.cpp
// compile and run here without PETSc / MPI.  What comes out of such a run (node numbering, dof offsets, sparsity
// counts, prolongators, Dirichlet flags, assembled matrices, residual norms) is REFERENCE OUTPUT: it pins the numpy
// restatement (oracle/mesh_box.py, oracle/mg.py) and the product's host layer (femus_b200/host/*.hpp).
//
// The classes take the NAMES of the reference's PETSc classes (the reference's factories do `new PetscVector` etc.;
// their real headers are skipped through their include guards, oracle/ref_shims/FemusConfig.hpp).  The multigrid
// solver restates what LinearEquationSolverPetsc configures in PETSc's PCMG (LinearEquationSolverPetsc.cpp:185-353):
// multiplicative V-cycle, Richardson(scale) + Jacobi smoothing, restriction with P^T, direct coarse solve.
// Pre-included (-include) into the three factory translation units and the drivers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

#include "NumericVector.hpp"
#include "SparseMatrix.hpp"
#include "DenseMatrix.hpp"
#include "DenseVector.hpp"
#include "DenseSubvector.hpp"
#include "LinearEquationSolver.hpp"
#include "Mesh.hpp"
#include "Solution.hpp"

namespace femus {

#define HOSTBACKEND_ABORT(what)                                                   \
  do {                                                                            \
    std::fprintf(stderr, "oracle host backend: %s is not implemented\n", what);    \
    std::abort();                                                                 \
  } while (0)

class HostMatrix;

class HostVector : public NumericVector {
 public:
  explicit HostVector(const ParallelType type = AUTOMATIC) : NumericVector(type) {}
  explicit HostVector(const int n, const ParallelType type = AUTOMATIC) : NumericVector(type) { this->init(n, n, false, type); }
  HostVector(const int n, const int n_local, const ParallelType type = AUTOMATIC) : NumericVector(type) { this->init(n, n_local, false, type); }
  HostVector(const int N, const int n_local, const std::vector<int>& ghost, const ParallelType type = AUTOMATIC) : NumericVector(type) {
    this->init(N, n_local, ghost, false, type);
  }
  ~HostVector() { this->clear(); }
  std::vector<double>& data() { return _v; }
  const std::vector<double>& data() const { return _v; }
  static const HostVector& cast(const NumericVector& v) { return static_cast<const HostVector&>(v); }

  void clear() override { _v.clear(); _is_closed = false; _is_initialized = false; }
  std::unique_ptr<NumericVector> clone() const override {
    std::unique_ptr<NumericVector> c(new HostVector);
    c->init(*this, true);
    *c = *this;
    return c;
  }
  void close() override { _is_closed = true; }
  void closeWithMinValues() override { _is_closed = true; }
  void init(const int n, const int n_local, const bool = false, const ParallelType type = AUTOMATIC) override {
    if (n != n_local) { std::fprintf(stderr, "oracle host backend: one rank only (N=%d, n_local=%d)\n", n, n_local); std::abort(); }
    _v.assign((size_t)n, 0.0);
    _type = type == AUTOMATIC ? SERIAL : type;
    _is_initialized = true;
    _is_closed = true;
  }
  void init(const int n, const bool fast = false, const ParallelType type = AUTOMATIC) override { this->init(n, n, fast, type); }
  void init(const int N, const int n_local, const std::vector<int>&, const bool fast = false, const ParallelType type = AUTOMATIC) override {
    this->init(N, n_local, fast, type);      // one rank: no ghost entries
  }
  void init(const NumericVector& other, const bool fast = false) override { this->init(other.size(), other.local_size(), fast, other.type()); }

  void set(const int i, const double value) override { _v[i] = value; _is_closed = false; }
  void add(const int i, const double value) override { _v[i] += value; _is_closed = false; }
  void zero() override { std::fill(_v.begin(), _v.end(), 0.0); }
  NumericVector& operator=(const double s) override { std::fill(_v.begin(), _v.end(), s); return *this; }
  NumericVector& operator=(const NumericVector& V) override { _v = cast(V)._v; _is_initialized = true; _is_closed = true; return *this; }
  HostVector& operator=(const HostVector& V) { _v = V._v; _is_initialized = true; _is_closed = true; return *this; }
  NumericVector& operator=(const std::vector<double>& v) override { _v = v; return *this; }
  void insert(const std::vector<double>& v, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] = v[k]; }
  void insert(const NumericVector& V, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] = V((int)k); }
  void insert(const DenseVector& V, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] = V((unsigned)k); }
  void insert(const DenseSubVector& V, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] = V((unsigned)k); }
  double min() const override { return *std::min_element(_v.begin(), _v.end()); }
  double max() const override { return *std::max_element(_v.begin(), _v.end()); }
  double sum() const override { double s = 0; for (double x : _v) s += x; return s; }
  double l1_norm() const override { double s = 0; for (double x : _v) s += std::fabs(x); return s; }
  double l2_norm() const override { double s = 0; for (double x : _v) s += x * x; return std::sqrt(s); }
  double linfty_norm() const override { double s = 0; for (double x : _v) s = std::max(s, std::fabs(x)); return s; }
  int size() const override { return (int)_v.size(); }
  int local_size() const override { return (int)_v.size(); }
  int first_local_index() const override { return 0; }
  int last_local_index() const override { return (int)_v.size(); }
  double operator()(const int i) const override { return _v[i]; }
  void get(const std::vector<int>& index, std::vector<double>& values) const override {
    values.resize(index.size());
    for (size_t k = 0; k < index.size(); k++) values[k] = _v[index[k]];
  }
  NumericVector& operator+=(const NumericVector& V) override { this->add(1.0, V); return *this; }
  NumericVector& operator-=(const NumericVector& V) override { this->add(-1.0, V); return *this; }
  void add(const double s) override { for (double& x : _v) x += s; }
  void add(const NumericVector& V) override { this->add(1.0, V); }
  void add(const double a, const NumericVector& V) override {
    const std::vector<double>& w = cast(V)._v;
    for (size_t i = 0; i < _v.size(); i++) _v[i] += a * w[i];
  }
  void add_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] += v[k]; }
  void add_vector_blocked(const std::vector<double>& v, const std::vector<unsigned>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] += v[k]; }
  void insert_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof) override { this->insert(v, dof); }
  void add_vector(const std::vector<double>& v, const std::vector<int>& dof) override { this->add_vector_blocked(v, dof); }
  void add_vector(const NumericVector& V, const std::vector<int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] += V((int)k); }
  void add_vector(const DenseVector& V, const std::vector<unsigned int>& dof) override { for (size_t k = 0; k < dof.size(); k++) _v[dof[k]] += V((unsigned)k); }
  void add_vector(const NumericVector& x, const SparseMatrix& A) override;      // this += A x
  void resid(const NumericVector& rhs, const NumericVector& x, const SparseMatrix& A) override;      // this = rhs - A x
  void matrix_mult(const NumericVector& x, const SparseMatrix& A) override;     // this = A x
  void matrix_mult_transpose(const NumericVector& x, const SparseMatrix& A) override;      // this = A^T x
  void scale(const double factor) override { for (double& x : _v) x *= factor; }
  void abs() override { for (double& x : _v) x = std::fabs(x); }
  double dot(const NumericVector& V) const override {
    const std::vector<double>& w = cast(V)._v;
    double s = 0;
    for (size_t i = 0; i < _v.size(); i++) s += _v[i] * w[i];
    return s;
  }
  void localize(std::vector<double>& v_local) const override { v_local = _v; }
  void localize(NumericVector& v_local) const override { v_local = *this; }
  void localize(NumericVector& v_local, const std::vector<int>&) const override { v_local = *this; }
  void localize(const int, const int, const std::vector<int>&) override {}
  void localize_to_one(std::vector<double>& v_local, const int = 0) const override { v_local = _v; }
  void localize_to_all(std::vector<double>& v_local) const override { v_local = _v; }
  void pointwise_mult(const NumericVector& a, const NumericVector& b) override {
    for (size_t i = 0; i < _v.size(); i++) _v[i] = cast(a)._v[i] * cast(b)._v[i];
  }
  void swap(NumericVector& v) override { _v.swap(static_cast<HostVector&>(v)._v); }
  void print_personal(std::ostream& os = std::cout) const { for (double x : _v) os << x << "\n"; }
  void print_hdf5(const std::string) const {}
  void BinaryPrint(const char* fileName) override {
    std::ofstream f(fileName, std::ios::binary);
    f.write(reinterpret_cast<const char*>(_v.data()), (std::streamsize)(_v.size() * sizeof(double)));
  }
  void BinaryLoad(const char* fileName) override {
    std::ifstream f(fileName, std::ios::binary);
    f.read(reinterpret_cast<char*>(_v.data()), (std::streamsize)(_v.size() * sizeof(double)));
  }

 private:
  std::vector<double> _v;
};

// Rows as ordered maps: insertion keeps every entry that was ever set or added (explicit zeros included, as MatSetValues
// does), so after an assembly the structure IS the reference's sparsity pattern.
class HostMatrix : public SparseMatrix {
 public:
  HostMatrix() : _ncols(0) {}
  ~HostMatrix() { this->clear(); }
  static const HostMatrix& cast(const SparseMatrix& A) { return static_cast<const HostMatrix&>(A); }
  const std::vector<std::map<int, double>>& rows() const { return _rows; }
  const std::vector<int>& n_nz() const { return _n_nz; }      // what LinearEquation::GetSparsityPatternSize handed in
  const std::vector<int>& n_oz() const { return _n_oz; }

  void clear() override { _rows.clear(); _ncols = 0; _is_initialized = false; }
  void init() override {}
  void init(const int m, const int n, const int m_l, const int n_l, const int = 30, const int = 10) override { this->init_dims(m, n, m_l, n_l); }
  void init(const int m, const int n) override { this->init_dims(m, n, m, n); }
  void init(const int m, const int n, const int m_l, const int n_l, const std::vector<int>& n_nz, const std::vector<int>& n_oz) override {
    this->init_dims(m, n, m_l, n_l);
    _n_nz = n_nz;
    _n_oz = n_oz;
  }
  void init(const int, const int, const std::vector<SparseMatrix*>&) override { HOSTBACKEND_ABORT("SparseMatrix::init(nr, nc, P)"); }
  void set(const int i, const int j, const double value) override { _rows[i][j] = value; }
  void add(const int i, const int j, const double value) override { _rows[i][j] += value; }
  void zero() override {
    for (auto& r : _rows)
      for (auto& e : r) e.second = 0.0;
  }
  void zero_rows(std::vector<int>& rows, double diag_value = 0.0) override { this->mat_zero_rows(rows, diag_value); }
  void close() const override { const_cast<HostMatrix*>(this)->_closed = true; }
  double operator()(const int i, const int j) const override {
    auto it = _rows[i].find(j);
    return it == _rows[i].end() ? 0.0 : it->second;
  }
  int MatGetRowM(const int i, int* cols = NULL, double* vals = NULL) override {
    int k = 0;
    for (const auto& e : _rows[i]) {
      if (cols) cols[k] = e.first;
      if (vals) vals[k] = e.second;
      k++;
    }
    return k;
  }
  void RemoveZeroEntries(double& tolerance) override {
    for (auto& r : _rows)
      for (auto it = r.begin(); it != r.end();)
        if (std::fabs(it->second) < tolerance) it = r.erase(it); else ++it;
  }
  bool closed() const override { return _closed; }
  void update_sparsity_pattern_old(const Graph&) override { HOSTBACKEND_ABORT("update_sparsity_pattern_old"); }
  void update_sparsity_pattern(const Graph&) override { HOSTBACKEND_ABORT("update_sparsity_pattern(Graph)"); }
  void update_sparsity_pattern(int m, int n, int m_l, int n_l, const std::vector<int> n_oz, const std::vector<int> n_nz) override {
    this->init(m, n, m_l, n_l, n_nz, n_oz);
  }
  int m() const override { return (int)_rows.size(); }
  int n() const override { return _ncols; }
  int row_start() const override { return 0; }
  int row_stop() const override { return (int)_rows.size(); }
  void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& rows, const std::vector<unsigned int>& cols) override {
    for (size_t i = 0; i < rows.size(); i++)
      for (size_t j = 0; j < cols.size(); j++) _rows[rows[i]][(int)cols[j]] += dm((unsigned)i, (unsigned)j);
  }
  void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& dof) override { this->add_matrix(dm, dof, dof); }
  void insert_row(const int row, const int ncols, const std::vector<int>& cols, double* values) override {
    for (int k = 0; k < ncols; k++) _rows[row][cols[k]] = values[k];
  }
  void add_matrix_blocked(const std::vector<double>& v, const std::vector<int>& rows, const std::vector<int>& cols) override {
    for (size_t i = 0; i < rows.size(); i++)
      for (size_t j = 0; j < cols.size(); j++) _rows[rows[i]][cols[j]] += v[i * cols.size() + j];
  }
  void add_matrix_blocked(const std::vector<double>& v, const std::vector<unsigned>& rows, const std::vector<unsigned>& cols) override {
    for (size_t i = 0; i < rows.size(); i++)
      for (size_t j = 0; j < cols.size(); j++) _rows[rows[i]][(int)cols[j]] += v[i * cols.size() + j];
  }
  void matrix_set_off_diagonal_values_blocked(const std::vector<int>&, const std::vector<int>&, const double&) override { HOSTBACKEND_ABORT("matrix_set_off_diagonal_values_blocked"); }
  void matrix_set_off_diagonal_values_blocked(const std::vector<int>&, const std::vector<int>&, const std::vector<double>&) override { HOSTBACKEND_ABORT("matrix_set_off_diagonal_values_blocked"); }
  void matrix_add(const double a, SparseMatrix& X, const char[]) override { this->add(a, X); }
  void add(const double c, SparseMatrix& B) override {       // this += c B
    const HostMatrix& b = cast(B);
    for (size_t i = 0; i < _rows.size(); i++)
      for (const auto& e : b._rows[i]) _rows[i][e.first] += c * e.second;
  }
  // this = P^T A P (MatPtAP): pattern = every structurally possible entry
  void matrix_PtAP(const SparseMatrix& mat_P, const SparseMatrix& mat_A, const bool&) override {
    const HostMatrix &P = cast(mat_P), &A = cast(mat_A);
    this->init_dims(P.n(), P.n(), P.n(), P.n());
    std::map<int, double> ap;
    for (int i = 0; i < A.m(); i++) {
      if (P._rows[i].empty()) continue;
      ap.clear();                                   // row i of A P
      for (const auto& a : A._rows[i])
        for (const auto& p : P._rows[a.first]) ap[p.first] += a.second * p.second;
      for (const auto& pi : P._rows[i])
        for (const auto& e : ap) _rows[pi.first][e.first] += pi.second * e.second;
    }
  }
  void matrix_ABC(const SparseMatrix& mat_A, const SparseMatrix& mat_B, const SparseMatrix& mat_C, const bool&) override {
    HostMatrix BC;
    product(cast(mat_B), cast(mat_C), BC);
    product(cast(mat_A), BC, *this);
  }
  void matrix_RightMatMult(const SparseMatrix& mat_A) override {      // this = this A
    HostMatrix T;
    product(*this, cast(mat_A), T);
    _rows.swap(T._rows);
    _ncols = T._ncols;
  }
  void matrix_LeftMatMult(const SparseMatrix& mat_A) override {       // this = A this
    HostMatrix T;
    product(cast(mat_A), *this, T);
    _rows.swap(T._rows);
    _ncols = T._ncols;
  }
  void matrix_get_diagonal_values(const std::vector<int>& index, std::vector<double>& value) const override {
    value.resize(index.size());
    for (size_t k = 0; k < index.size(); k++) value[k] = (*this)(index[k], index[k]);
  }
  void matrix_set_diagonal_values(NumericVector& D) override { for (int i = 0; i < m(); i++) _rows[i][i] = D(i); }
  void matrix_set_diagonal_values(const std::vector<int>& index, const double& value) override { for (int i : index) _rows[i][i] = value; }
  void matrix_set_diagonal_values(const std::vector<int>& index, const std::vector<double>& value) override {
    for (size_t k = 0; k < index.size(); k++) _rows[index[k]][index[k]] = value[k];
  }
  double l1_norm() const override {
    std::vector<double> c((size_t)_ncols, 0.0);
    for (const auto& r : _rows)
      for (const auto& e : r) c[e.first] += std::fabs(e.second);
    return c.empty() ? 0.0 : *std::max_element(c.begin(), c.end());
  }
  double linfty_norm() const override {
    double best = 0;
    for (const auto& r : _rows) {
      double s = 0;
      for (const auto& e : r) s += std::fabs(e.second);
      best = std::max(best, s);
    }
    return best;
  }
  void get_diagonal(NumericVector& dest) const override { for (int i = 0; i < m(); i++) dest.set(i, (*this)(i, i)); }
  void get_transpose(SparseMatrix& dest) const override {
    HostMatrix& T = static_cast<HostMatrix&>(dest);
    std::vector<std::map<int, double>> t((size_t)_ncols);
    for (int i = 0; i < m(); i++)
      for (const auto& e : _rows[i]) t[e.first][i] = e.second;
    const int nr = m();
    T._rows.swap(t);
    T._ncols = nr;
    T._is_initialized = true;
  }
  // MatZeroRows(A, idx, diag): the rows keep their pattern (MAT_KEEP_NONZERO_PATTERN, PetscMatrix.cpp:1076 /
  // LinearEquationSolverPetsc.cpp:428-436), the diagonal entry is set
  void mat_zero_rows(const std::vector<int>& index, const double& diagonal_value) const override {
    HostMatrix* self = const_cast<HostMatrix*>(this);
    for (int i : index) {
      for (auto& e : self->_rows[i]) e.second = 0.0;
      if (diagonal_value != 0.0 || self->_rows[i].count(i)) self->_rows[i][i] = diagonal_value;
    }
  }
  void print_personal(std::ostream& os = std::cout) const override {
    for (int i = 0; i < m(); i++)
      for (const auto& e : _rows[i]) os << i << " " << e.first << " " << e.second << "\n";
  }
  void print_hdf5(const std::string = "NULL") const override {}

  // y = A x, y = A^T x
  void mult(const std::vector<double>& x, std::vector<double>& y) const {
    y.assign(_rows.size(), 0.0);
    for (size_t i = 0; i < _rows.size(); i++) {
      double s = 0;
      for (const auto& e : _rows[i]) s += e.second * x[e.first];
      y[i] = s;
    }
  }
  void mult_transpose(const std::vector<double>& x, std::vector<double>& y) const {
    y.assign((size_t)_ncols, 0.0);
    for (size_t i = 0; i < _rows.size(); i++)
      for (const auto& e : _rows[i]) y[e.first] += e.second * x[i];
  }
  static void product(const HostMatrix& A, const HostMatrix& B, HostMatrix& C) {
    std::vector<std::map<int, double>> c((size_t)A.m());
    for (int i = 0; i < A.m(); i++)
      for (const auto& a : A._rows[i])
        for (const auto& b : B._rows[a.first]) c[i][b.first] += a.second * b.second;
    C._rows.swap(c);
    C._ncols = B._ncols;
    C._is_initialized = true;
  }

 private:
  void init_dims(const int m, const int n, const int m_l, const int n_l) {
    if (m != m_l || n != n_l) { std::fprintf(stderr, "oracle host backend: one rank only (matrix %d x %d, local %d x %d)\n", m, n, m_l, n_l); std::abort(); }
    _rows.assign((size_t)m, std::map<int, double>());
    _ncols = n;
    _is_initialized = true;
    _closed = false;
  }
  std::vector<std::map<int, double>> _rows;
  std::vector<int> _n_nz, _n_oz;
  int _ncols;
  bool _closed = false;
};

inline void HostVector::add_vector(const NumericVector& x, const SparseMatrix& A) {
  std::vector<double> y;
  HostMatrix::cast(A).mult(cast(x)._v, y);
  for (size_t i = 0; i < _v.size(); i++) _v[i] += y[i];
}
inline void HostVector::resid(const NumericVector& rhs, const NumericVector& x, const SparseMatrix& A) {
  std::vector<double> y;
  HostMatrix::cast(A).mult(cast(x)._v, y);
  for (size_t i = 0; i < _v.size(); i++) _v[i] = cast(rhs)._v[i] - y[i];
}
inline void HostVector::matrix_mult(const NumericVector& x, const SparseMatrix& A) { HostMatrix::cast(A).mult(cast(x)._v, _v); }
inline void HostVector::matrix_mult_transpose(const NumericVector& x, const SparseMatrix& A) { HostMatrix::cast(A).mult_transpose(cast(x)._v, _v); }

// ---- the level solver: what LinearEquationSolverPetsc sets up in PCMG, restated on the host -------------------------
class HostLinearEquationSolver : public LinearEquationSolver {
 public:
  HostLinearEquationSolver(const unsigned& igrid, Solution* other_solution)
      : LinearEquationSolver(igrid, other_solution), _level(igrid), _richardsonScaleFactor(0.5), _levelMax(0), _npre(1), _npost(1), _PP(nullptr),
        _bdcIndexIsInitialized(false) {}
  ~HostLinearEquationSolver() {}

  void SetTolerances(const double&, const double&, const double&, const unsigned&, const unsigned&) override {}
  void SetRichardsonScaleFactor(const double& richardsonScaleFactor) override { _richardsonScaleFactor = richardsonScaleFactor; }
  // rows that are Dirichlet (Bdc < 1.5) or whose variable is not solved (LinearEquationSolverPetsc.cpp:53-90)
  void BuildBdcIndex(const std::vector<unsigned>& variable_to_be_solved) {
    _bdcIndexIsInitialized = true;
    _bdcIndex.clear();
    std::vector<bool> included(_SolPdeIndex.size(), false);
    for (unsigned v : variable_to_be_solved) included[v] = true;
    for (unsigned k = 0; k < _SolPdeIndex.size(); k++) {
      const unsigned indexSol = _SolPdeIndex[k], soltype = _SolType[indexSol];
      const unsigned i0 = GetMeshFromLinEq()->_dofOffset[soltype][processor_id()], i1 = GetMeshFromLinEq()->_dofOffset[soltype][processor_id() + 1];
      for (unsigned i = i0; i < i1; i++)
        if (!included[k] || (*(*_Bdc)[indexSol])(i) < 1.5) _bdcIndex.push_back((int)(KKoffset[k][processor_id()] + (i - i0)));
    }
    std::sort(_bdcIndex.begin(), _bdcIndex.end());
  }
  const std::vector<int>& BdcIndex() const { return _bdcIndex; }
  // one level, no multigrid: SetPenalty, then the direct solve the reference configures for a single level (PREONLY + LU)
  void Solve(const std::vector<unsigned>& variable_to_be_solved, const bool&) override {
    if (!_bdcIndexIsInitialized) this->BuildBdcIndex(variable_to_be_solved);
    _KK->mat_zero_rows(_bdcIndex, 1.0);
    HostVector& RES = static_cast<HostVector&>(*_RES);
    for (int i : _bdcIndex) RES.data()[i] = 0.0;
    std::vector<double> x;
    direct_solve(HostMatrix::cast(*_KK), RES.data(), x);
    static_cast<HostVector&>(*_EPSC).data() = x;
    _RESC->matrix_mult(*_EPSC, *_KK);
    *_RES -= *_RESC;
    *_EPS += *_EPSC;
  }
  void MGInit(const MgSmootherType& mg_smoother_type, const unsigned& levelMax, const SolverType&) override {
    if (mg_smoother_type != MULTIPLICATIVE) HOSTBACKEND_ABORT("a multigrid cycle other than the multiplicative V-cycle");
    _levelMax = levelMax;
    _hier.assign(levelMax, nullptr);
  }
  void MGClear() override { _hier.clear(); }
  void MGSetLevel(LinearEquationSolver* LinSolver, const unsigned& levelMax, const std::vector<unsigned>& variable_to_be_solved, SparseMatrix* PP,
                  SparseMatrix*, const unsigned& npre, const unsigned& npost) override {
    HostLinearEquationSolver* top = static_cast<HostLinearEquationSolver*>(LinSolver);
    if (top->_hier.size() != (size_t)levelMax + 1) { std::fprintf(stderr, "oracle host backend: MGSetLevel before MGInit\n"); std::abort(); }
    if (!_bdcIndexIsInitialized) this->BuildBdcIndex(variable_to_be_solved);
    if (const char* dir = std::getenv("FEMUS_REF_DUMP")) this->dump_level(dir, PP);      // before the penalty: the assembled / Galerkin operator
    _KK->mat_zero_rows(_bdcIndex, 1.0);                 // SetPenalty (:428-436)
    if (_level > 0 && (this->_levelSolverType != RICHARDSON ||
                       (this->preconditioner_type() != JACOBI_PRECOND && this->preconditioner_type() != SOR_PRECOND))) {
      std::fprintf(stderr, "oracle host backend: level smoother must be Richardson + Jacobi or SOR (solver %d, preconditioner %d)\n",
                   (int)this->_levelSolverType, (int)this->preconditioner_type());
      std::abort();
    }
    _sor = this->preconditioner_type() == SOR_PRECOND;
    _PP = PP;
    _npre = npre;
    _npost = npost;
    top->_hier[_level] = this;
  }
  // one multiplicative V-cycle as outer PREONLY (:294-353): ZerosBoundaryResiduals; EPSC = V(RES); RESC = KK EPSC; RES -= RESC; EPS += EPSC
  void MGSolve(const bool) override {
    HostVector& RES = static_cast<HostVector&>(*_RES);
    if (const char* dir = std::getenv("FEMUS_REF_DUMP")) {
      if (!_resDumped) dump_array(dir, "RES", RES.data());          // the assembled residual of the first cycle
      _resDumped = true;
    }
    for (int i : _bdcIndex) RES.data()[i] = 0.0;
    std::vector<double> x;
    vcycle((int)_level, RES.data(), x);
    static_cast<HostVector&>(*_EPSC).data() = x;
    _RESC->matrix_mult(*_EPSC, *_KK);
    *_RES -= *_RESC;
    *_EPS += *_EPSC;
  }

 private:
  // ---- reference output for the golden fixtures (tests/golden/make_ref_golden.py): one raw little-endian file per array,
  // <dir>/L<level>_<name>.<i4|i8|f8>, everything read out of the reference's own objects
  template <class T>
  void dump_array(const char* dir, const char* name, const std::vector<T>& a) const {
    const char* suffix = sizeof(T) == 8 ? (std::is_floating_point<T>::value ? "f8" : "i8") : "i4";
    const std::string path = std::string(dir) + "/L" + std::to_string(_level) + "_" + name + "." + suffix;
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(a.data()), (std::streamsize)(a.size() * sizeof(T)));
  }
  void dump_csr(const char* dir, const char* name, const HostMatrix& A) const {
    std::vector<long long> rp(1, 0);
    std::vector<int> col;
    std::vector<double> val;
    for (const auto& r : A.rows()) {
      for (const auto& e : r) { col.push_back(e.first); val.push_back(e.second); }
      rp.push_back((long long)col.size());
    }
    dump_array(dir, (std::string(name) + "_rowptr").c_str(), rp);
    dump_array(dir, (std::string(name) + "_col").c_str(), col);
    dump_array(dir, (std::string(name) + "_val").c_str(), val);
    dump_array(dir, (std::string(name) + "_shape").c_str(), std::vector<int>{A.m(), A.n()});
  }
  void dump_level(const char* dir, SparseMatrix* PP) const {
    const Mesh* msh = GetMeshFromLinEq();
    const unsigned nel = msh->GetNumberOfElements(), nnode = msh->GetNumberOfNodes();
    std::vector<int> conn((size_t)nel * 27, -1), etype(nel), faces((size_t)nel * 6, 0), sysdof, info;
    for (unsigned iel = 0; iel < nel; iel++) {
      etype[iel] = msh->GetElementType(iel);
      const unsigned nn = msh->GetMeshElements()->GetElementDofNumber(iel, 2);
      for (unsigned j = 0; j < nn; j++) conn[(size_t)iel * 27 + j] = (int)msh->GetMeshElements()->GetElementDofIndex(iel, j);
      for (unsigned f = 0; f < msh->GetMeshElements()->GetElementFaceNumber(iel); f++) faces[(size_t)iel * 6 + f] = msh->GetMeshElements()->GetFaceElementIndex(iel, f);
    }
    std::vector<double> xyz((size_t)3 * nnode, 0.0);
    for (unsigned d = 0; d < msh->GetDimension(); d++)
      for (unsigned i = 0; i < nnode; i++) xyz[(size_t)d * nnode + i] = (*msh->GetTopology()->_Sol[d])(i);
    std::vector<int> dofoff;
    for (int k = 0; k < 3; k++)
      for (unsigned p = 0; p <= n_processors(); p++) dofoff.push_back((int)msh->_dofOffset[k][p]);
    // system dofs of every element and variable: [variable][element][27] (-1 padded)
    for (unsigned k = 0; k < _SolPdeIndex.size(); k++) {
      const unsigned indexSol = _SolPdeIndex[k], soltype = _SolType[indexSol];
      for (unsigned iel = 0; iel < nel; iel++) {
        const unsigned nve = msh->GetMeshElements()->GetElementDofNumber(iel, soltype);
        for (unsigned j = 0; j < 27; j++) sysdof.push_back(j < nve ? (int)GetSystemDof(indexSol, k, j, iel) : -1);
      }
    }
    std::vector<int> kkoff;
    for (const auto& row : KKoffset)
      for (unsigned v : row) kkoff.push_back((int)v);
    std::vector<double> bdc;
    for (unsigned k = 0; k < _SolPdeIndex.size(); k++) {
      const NumericVector& b = *(*_Bdc)[_SolPdeIndex[k]];
      for (int i = 0; i < b.size(); i++) bdc.push_back(b(i));
    }
    info = {(int)nel, (int)nnode, (int)msh->GetDimension(), (int)_SolPdeIndex.size(), (int)_SolType[_SolPdeIndex[0]], (int)n_processors()};
    dump_array(dir, "info", info);
    dump_array(dir, "conn", conn);
    dump_array(dir, "etype", etype);
    dump_array(dir, "face_index", faces);
    dump_array(dir, "xyz", xyz);
    dump_array(dir, "dofOffset", dofoff);
    dump_array(dir, "sysdof", sysdof);
    dump_array(dir, "KKoffset", kkoff);
    dump_array(dir, "Bdc", bdc);
    dump_array(dir, "bdcIndex", _bdcIndex);
    const HostMatrix& KK = HostMatrix::cast(*_KK);
    dump_array(dir, "n_nz", KK.n_nz());
    dump_array(dir, "n_oz", KK.n_oz());
    dump_csr(dir, "KK", KK);
    if (PP) dump_csr(dir, "PP", HostMatrix::cast(*PP));
  }
  bool _resDumped = false;

  // x = V-cycle(b) from a zero guess on level l of the hierarchy this (finest) solver owns
  void vcycle(const int l, const std::vector<double>& b, std::vector<double>& x) const {
    const HostLinearEquationSolver* L = _hier[l];
    const HostMatrix& A = HostMatrix::cast(*L->_KK);
    if (l == 0) { direct_solve(A, b, x); return; }
    const size_t n = b.size();
    std::vector<double> diag(n), r(n), t, z(n);
    for (size_t i = 0; i < n; i++) diag[i] = A((int)i, (int)i);
    const double w = L->_richardsonScaleFactor;
    x.assign(n, 0.0);
    // z = M^-1 r: Jacobi, or PCSOR's default (one local symmetric sweep, omega 1, zero guess):
    // z = (D + U)^-1 D (D + L)^-1 r
    auto precondition = [&]() {
      if (!L->_sor) { for (size_t i = 0; i < n; i++) z[i] = r[i] / diag[i]; return; }
      for (size_t i = 0; i < n; i++) {
        double s = r[i];
        for (const auto& e : A.rows()[i]) { if (e.first >= (int)i) break; s -= e.second * z[e.first]; }
        z[i] = s / diag[i];
      }
      for (size_t ii = n; ii-- > 0;) {
        double s = 0.0;
        for (auto it = A.rows()[ii].rbegin(); it != A.rows()[ii].rend() && it->first > (int)ii; ++it) s += it->second * z[it->first];
        z[ii] -= s / diag[ii];
      }
    };
    auto smooth = [&](unsigned sweeps, bool zero_guess) {
      for (unsigned s = 0; s < sweeps; s++) {
        if (zero_guess && s == 0) r = b;
        else {
          A.mult(x, t);
          for (size_t i = 0; i < n; i++) r[i] = b[i] - t[i];
        }
        precondition();
        for (size_t i = 0; i < n; i++) x[i] += w * z[i];
      }
    };
    smooth(L->_npre, true);
    A.mult(x, t);
    for (size_t i = 0; i < n; i++) r[i] = b[i] - t[i];
    const HostMatrix& P = HostMatrix::cast(*L->_PP);
    std::vector<double> bc, xc;
    P.mult_transpose(r, bc);                            // PCMGSetRestriction(..., PP): restriction = P^T (:277)
    vcycle(l - 1, bc, xc);
    P.mult(xc, t);
    for (size_t i = 0; i < n; i++) x[i] += t[i];
    smooth(L->_npost, false);
  }
  // dense LU with partial pivoting (the reference: PREONLY + LU through MUMPS on level 0)
  static void direct_solve(const HostMatrix& A, const std::vector<double>& b, std::vector<double>& x) {
    const int n = A.m();
    std::vector<double> M((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
      for (const auto& e : A.rows()[i]) M[(size_t)i * n + e.first] = e.second;
    x = b;
    std::vector<int> piv(n);
    for (int k = 0; k < n; k++) {
      int p = k;
      for (int i = k + 1; i < n; i++)
        if (std::fabs(M[(size_t)i * n + k]) > std::fabs(M[(size_t)p * n + k])) p = i;
      if (M[(size_t)p * n + k] == 0.0) { std::fprintf(stderr, "oracle host backend: singular coarse matrix\n"); std::abort(); }
      if (p != k) {
        for (int j = 0; j < n; j++) std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]);
        std::swap(x[k], x[p]);
      }
      for (int i = k + 1; i < n; i++) {
        const double f = M[(size_t)i * n + k] / M[(size_t)k * n + k];
        if (f == 0.0) continue;
        for (int j = k; j < n; j++) M[(size_t)i * n + j] -= f * M[(size_t)k * n + j];
        x[i] -= f * x[k];
      }
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = x[i];
      for (int j = i + 1; j < n; j++) s -= M[(size_t)i * n + j] * x[j];
      x[i] = s / M[(size_t)i * n + i];
    }
  }

  unsigned _level;
  double _richardsonScaleFactor;
  unsigned _levelMax, _npre, _npost;
  SparseMatrix* _PP;
  bool _sor = false;
  std::vector<int> _bdcIndex;
  bool _bdcIndexIsInitialized;
  std::vector<HostLinearEquationSolver*> _hier;      // on the finest solver: every level's solver
};

// the names the reference's factories instantiate
using PetscVector = HostVector;
using PetscMatrix = HostMatrix;
using LinearEquationSolverPetsc = HostLinearEquationSolver;
using LinearEquationSolverPetscAsm = HostLinearEquationSolver;
using LinearEquationSolverPetscFieldSplit = HostLinearEquationSolver;

}  // namespace femus
