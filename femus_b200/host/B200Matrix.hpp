// B200Matrix: FEMuS SparseMatrix on the femus_b200 device CSR.  Drop-in for PetscMatrix
// (reference src/03_algebra/01_matrices/PetscMatrix.{hpp,cpp}) behind SparseMatrix::build().
//
// Pattern.  PETSc preallocates from per-row COUNTS (init(m,n,m_l,n_l,n_nz,n_oz), SparseMatrix.hpp:73)
// and learns the columns while values are inserted; a device CSR needs the columns first.  Two ways:
//   * init_from_elements(): the exact element-coupling pattern straight from the element->dof lists
//     (what LinearEquation::GetSparsityPatternSize counts, LinearEquation.cpp:407-548), built on the
//     device.  This is what LinearEquation would call for the B200 backend.
//   * init() with counts: entries are collected on the host until the first close(), which freezes
//     the pattern and uploads it.  Later assemblies must stay inside it (as with PETSc's
//     MAT_NEW_NONZERO_ALLOCATION_ERR).  Correct for any caller, slow: it is the compatibility path.
// Values.  add_matrix_blocked()/insert_row() are staged on the host, like MatSetValues, and reach
// the device in one batch at close() (PetscMatrix.hpp:237-244).  The fused device assembly
// (b2_asm_poisson) writes the values in place and needs no staging at all.
#pragma once
#include <algorithm>
#include <map>
#include "B200Vector.hpp"
#ifdef B2_WITH_FEMUS_HEADERS
#include "DenseMatrix.hpp"
#include "Graph.hpp"
#endif

namespace femus {

class B200Matrix : public SparseMatrix {
 public:
  B200Matrix() : _A(nullptr), _frozen(false), _closed(false), _mirror_ok(false) { _m = _n = _m_l = _n_l = _ml_start = 0; }
  ~B200Matrix() { this->clear(); }
  static std::unique_ptr<SparseMatrix> build() { return std::unique_ptr<SparseMatrix>(new B200Matrix); }

  b2_csr* handle() const { return _A; }
  // changes whenever the device matrix behind this object is REPLACED (new pattern): who borrows the handle compares it
  uint64_t generation() const { return _gen; }
  bool frozen() const { return _frozen; }      // the pattern is on the device
  void touched() const { _mirror_ok = false; }      // device values were written behind our back (fused assembly)

  // ---- pattern ------------------------------------------------------------------------------
  void clear() override {
    if (_A) b2_csr_destroy(_A);
    _A = nullptr;
    _frozen = _closed = _mirror_ok = false;
    _rows.clear();
    drop_staging();
    _is_initialized = false;
  }
  void init(const int m, const int n, const int m_l, const int n_l, const std::vector<int>& n_nz,
            const std::vector<int>& n_oz) override {
    if (m_l != m || n_l != n) { std::fprintf(stderr, "femus_b200: B200Matrix::init: one rank holds its whole local matrix\n"); std::abort(); }
    this->clear();
    set_dims(m, n);
    _rows.assign((size_t)m, Row());
    for (int i = 0; i < m && i < (int)n_nz.size(); i++) _rows[i].reserve((size_t)n_nz[i] + (i < (int)n_oz.size() ? n_oz[i] : 0));
    _is_initialized = true;
  }
  void init(const int, const int, const std::vector<SparseMatrix*>&) override { B2_NOT_ON_PATH("init(nr, nc, blocks)"); }
  // exact element-coupling pattern (zeros included) from dof[nel][nve]
  void init_from_elements(const int m, const int64_t nel, const int nve, const int32_t* dof) {
    this->clear();
    set_dims(m, m);
    B2_ABORT_IF(b2_csr_create_from_elements(B200Context::get(), m, nel, nve, dof, &_A), "b2_csr_create_from_elements");
    new_generation();
    _frozen = _closed = _is_initialized = true;
  }
  void init_from_csr(const int m, const int n, const int64_t* rowptr, const int32_t* col, const double* vals) {
    this->clear();
    set_dims(m, n);
    B2_ABORT_IF(b2_csr_create(B200Context::get(), m, n, rowptr, col, vals, &_A), "b2_csr_create");
    new_generation();
    _frozen = _closed = _is_initialized = true;
  }
  // re-initialisation from per-row counts (PetscMatrix.cpp:206-298): the same staged start as init()
  void update_sparsity_pattern(int m, int n, int m_l, int n_l, const std::vector<int> n_oz, const std::vector<int> n_nz) override {
    this->init(m, n, m_l, n_l, n_nz, n_oz);
  }
#ifdef B2_WITH_FEMUS_HEADERS
  // a Graph carries the columns themselves (every row: its columns, then the off-diagonal count, PetscMatrix.cpp:301-320):
  // the exact pattern goes to the device at once
  void update_sparsity_pattern(const Graph& g) override {
    std::vector<int64_t> rp((size_t)g._m + 1, 0);
    std::vector<int32_t> col;
    for (unsigned i = 0; i < g._m; i++) {
      const size_t len = g[i].empty() ? 0 : g[i].size() - 1;
      std::vector<int32_t> c(g[i].begin(), g[i].begin() + (std::ptrdiff_t)len);
      std::sort(c.begin(), c.end());
      c.erase(std::unique(c.begin(), c.end()), c.end());
      col.insert(col.end(), c.begin(), c.end());
      rp[i + 1] = (int64_t)col.size();
    }
    this->init_from_csr((int)g._m, (int)g._n, rp.data(), col.data(), nullptr);
  }
  void update_sparsity_pattern_old(const Graph& g) override { this->update_sparsity_pattern(g); }
#else
  void update_sparsity_pattern_old(const Graph&) override { B2_NOT_ON_PATH("update_sparsity_pattern_old (needs the FEMuS Graph class)"); }
  void update_sparsity_pattern(const Graph&) override { B2_NOT_ON_PATH("update_sparsity_pattern(Graph) (needs the FEMuS Graph class)"); }
#endif

  // ---- staged element access --------------------------------------------------------------------
  void set(const int i, const int j, const double value) override {
    std::vector<int> c(1, j);
    double v = value;
    this->insert_row(i, 1, c, &v);
  }
  void add(const int i, const int j, const double value) override {
    std::vector<double> v(1, value);
    std::vector<int> r(1, i), c(1, j);
    this->add_matrix_blocked(v, r, c);
  }
  void add_matrix_blocked(const std::vector<double>& v, const std::vector<int>& rows, const std::vector<int>& cols) override {
    add_block(v.data(), rows.data(), (int)rows.size(), cols.data(), (int)cols.size());
  }
  void add_matrix_blocked(const std::vector<double>& v, const std::vector<unsigned>& rows, const std::vector<unsigned>& cols) override {
    add_block(v.data(), reinterpret_cast<const int*>(rows.data()), (int)rows.size(), reinterpret_cast<const int*>(cols.data()),
              (int)cols.size());
  }
  void insert_row(const int row, const int ncols, const std::vector<int>& cols, double* values) override {
    check_index(row, _m, "row");
    _closed = false;
    if (!_frozen) {
      Row& r = _rows[row];
      for (int k = 0; k < ncols; k++) {
        check_index(cols[k], _n, "column");
        r.set(cols[k], values[k]);
      }
      return;
    }
    _ins_rows.push_back(row);
    for (int k = 0; k < ncols; k++) { _ins_cols.push_back(cols[k]); _ins_vals.push_back(values[k]); }
    _ins_ptr.push_back((int64_t)_ins_cols.size());
  }
#ifdef B2_WITH_FEMUS_HEADERS
  // MatSetValues(ADD_VALUES) of a dense element matrix (PetscMatrix.cpp:678-696), staged like add_matrix_blocked
  void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& rows, const std::vector<unsigned int>& cols) override {
    add_block(dm.get_values().data(), reinterpret_cast<const int*>(rows.data()), (int)rows.size(), reinterpret_cast<const int*>(cols.data()),
              (int)cols.size());
  }
  void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& dof_indices) override { this->add_matrix(dm, dof_indices, dof_indices); }
#else
  void add_matrix(const DenseMatrix&, const std::vector<unsigned int>&, const std::vector<unsigned int>&) override { B2_NOT_ON_PATH("add_matrix(DenseMatrix) (needs the FEMuS DenseMatrix class)"); }
  void add_matrix(const DenseMatrix&, const std::vector<unsigned int>&) override { B2_NOT_ON_PATH("add_matrix(DenseMatrix) (needs the FEMuS DenseMatrix class)"); }
#endif
  void zero() override {
    if (_frozen) {
      drop_staging();
      B2_ABORT_IF(b2_csr_zero(_A), "b2_csr_zero");
      touched();
    } else {
      for (Row& r : _rows) std::fill(r.val.begin(), r.val.end(), 0.0);
    }
  }
  // MatAssemblyBegin/End: the first close() of a count-initialised matrix freezes the pattern
  void close() const override { const_cast<B200Matrix*>(this)->do_close(); }
  bool closed() const override { return _closed; }

  // ---- algebra -------------------------------------------------------------------------------------
  // this = P^T A P (MatPtAP, PetscMatrix.cpp:733-751).  reuse (MAT_REUSE_MATRIX): only the numeric phase, onto the pattern
  // in place (the coarse element-coupling pattern or an earlier product).  Otherwise, as MAT_INITIAL_MATRIX, the matrix is
  // rebuilt with the pattern of the product -- in an F-cycle the same level matrix is first an assembled (AMR-constrained)
  // operator and later a Galerkin product with a larger pattern (LinearImplicitSystem.cpp:347-370): P^T (A P) by two
  // general products on the device.
  void matrix_PtAP(const SparseMatrix& mat_P, const SparseMatrix& mat_A, const bool& reuse) override {
    const B200Matrix& P = cast(mat_P);
    const B200Matrix& A = cast(mat_A);
    P.close();
    A.close();
    if (reuse && _frozen && _m == P._n && _n == P._n) {
      drop_staging();
      B2_ABORT_IF(b2_csr_ptap(P._A, A._A, _A), "b2_csr_ptap");
      _closed = true;
      touched();
      return;
    }
    b2_csr *Pt = nullptr, *AP = nullptr, *C = nullptr;
    B2_ABORT_IF(b2_csr_transpose(P._A, &Pt), "b2_csr_transpose");
    B2_ABORT_IF(b2_csr_matmat(A._A, P._A, &AP), "b2_csr_matmat");
    B2_ABORT_IF(b2_csr_matmat(Pt, AP, &C), "b2_csr_matmat");
    b2_csr_destroy(AP);
    b2_csr_destroy(Pt);
    adopt(C, P._n, P._n);
  }
  // this = A B C (MatMatMatMult, PetscMatrix.cpp:833-856: KK <- RRamr KKamr PPamr on a non-homogeneous level,
  // LinearImplicitSystem.cpp:336-341): two general products on the device; with reuse the pattern is the same by construction
  void matrix_ABC(const SparseMatrix& mat_A, const SparseMatrix& mat_B, const SparseMatrix& mat_C, const bool&) override {
    const B200Matrix &A = cast(mat_A), &B = cast(mat_B), &C = cast(mat_C);
    A.close();
    B.close();
    C.close();
    b2_csr *BC = nullptr, *ABC = nullptr;
    B2_ABORT_IF(b2_csr_matmat(B._A, C._A, &BC), "b2_csr_matmat");
    B2_ABORT_IF(b2_csr_matmat(A._A, BC, &ABC), "b2_csr_matmat");
    b2_csr_destroy(BC);
    adopt(ABC, A._m, C._n);
  }
  // this = this A / this = A this (MatMatMult, PetscMatrix.cpp:858-896: _PP[ig] <- _PP[ig] _PPamr[ig-1], LinearImplicitSystem.cpp:255-258)
  void matrix_RightMatMult(const SparseMatrix& mat_A) override {
    const B200Matrix& A = cast(mat_A);
    A.close();
    this->close();
    b2_csr* C = nullptr;
    B2_ABORT_IF(b2_csr_matmat(_A, A._A, &C), "b2_csr_matmat");
    adopt(C, _m, A._n);
  }
  void matrix_LeftMatMult(const SparseMatrix& mat_A) override {
    const B200Matrix& A = cast(mat_A);
    A.close();
    this->close();
    b2_csr* C = nullptr;
    B2_ABORT_IF(b2_csr_matmat(A._A, _A, &C), "b2_csr_matmat");
    adopt(C, A._m, _n);
  }
  // this += a X (MatAXPY, PetscMatrix.cpp: matrix_add / add).  X inside this pattern: in place on the device; otherwise the
  // union pattern is formed on the host first (DIFFERENT_NONZERO_PATTERN)
  void matrix_add(const double a, SparseMatrix& X_in, const char[]) override { this->add(a, X_in); }
  void add(const double a, SparseMatrix& X_in) override {
    const B200Matrix& X = cast(X_in);
    X.close();
    this->close();
    int inside = 0;
    B2_ABORT_IF(b2_csr_pattern_contains(_A, X._A, &inside), "b2_csr_pattern_contains");
    if (!inside) {
      const std::vector<int64_t> rp = host_rowptr(), xrp = X.host_rowptr();
      const std::vector<int32_t> ci = host_col(), xci = X.host_col();
      const std::vector<double> v = host_val();
      std::vector<int64_t> urp((size_t)_m + 1, 0);
      std::vector<int32_t> uci;
      std::vector<double> uv;
      for (int i = 0; i < _m; i++) {
        int64_t p = rp[i], q = xrp[i];
        while (p < rp[i + 1] || q < xrp[i + 1]) {
          const bool take_x = p == rp[i + 1] || (q < xrp[i + 1] && xci[q] < ci[p]);
          if (take_x) { uci.push_back(xci[q++]); uv.push_back(0.0); }
          else { if (q < xrp[i + 1] && xci[q] == ci[p]) q++; uci.push_back(ci[p]); uv.push_back(v[p++]); }
        }
        urp[i + 1] = (int64_t)uci.size();
      }
      const int m = _m, n = _n;
      this->init_from_csr(m, n, urp.data(), uci.data(), uv.data());
    }
    B2_ABORT_IF(b2_csr_axpy(_A, a, X._A), "b2_csr_axpy");
    touched();
  }
  void get_transpose(SparseMatrix& dest) const override {
    this->close();
    B200Matrix& T = dynamic_cast<B200Matrix&>(dest);
    b2_csr* At = nullptr;
    B2_ABORT_IF(b2_csr_transpose(_A, &At), "b2_csr_transpose");
    const int m = _m, n = _n;       // dest may be this matrix (LinearImplicitSystem.cpp:1025 transposes _PPamr in place)
    T.adopt(At, n, m);
  }
  void get_diagonal(NumericVector& dest) const override {
    this->close();
    B200Vector& d = dynamic_cast<B200Vector&>(dest);
    B2_ABORT_IF(b2_csr_diag(_A, d.handle()), "b2_csr_diag");
    d.touched();
  }
  // MatZeroRows keeping the pattern (PetscMatrix.cpp:1073-1077)
  void mat_zero_rows(const std::vector<int>& index, const double& diagonal_value) const override {
    this->close();
    B2_ABORT_IF(b2_csr_zero_rows(_A, index.data(), (int64_t)index.size(), diagonal_value), "b2_csr_zero_rows");
    touched();
  }
  // zero the listed columns (ZeroInterpolatorDirichletNodes does it with two transposes, LinearImplicitSystem.cpp:1090-1112)
  void mat_zero_cols(const std::vector<int>& index) const {
    this->close();
    B2_ABORT_IF(b2_csr_zero_cols(_A, index.data(), (int64_t)index.size()), "b2_csr_zero_cols");
    touched();
  }
  void matrix_get_diagonal_values(const std::vector<int>& index, std::vector<double>& value) const override {
    value.resize(index.size());
    for (size_t k = 0; k < index.size(); k++) value[k] = (*this)(index[k], index[k]);
  }
  // MatDiagonalSet / MatSetValues(INSERT) on single entries (PetscMatrix.cpp:911-952): staged like every other insertion
  void matrix_set_diagonal_values(NumericVector& D) override {
    std::vector<double> d;
    D.localize(d);
    for (int i = 0; i < _m && i < (int)d.size(); i++) this->set(i, i, d[i]);
  }
  void matrix_set_diagonal_values(const std::vector<int>& index, const double& value) override {
    for (int i : index) this->set(i, i, value);
  }
  void matrix_set_diagonal_values(const std::vector<int>& index, const std::vector<double>& value) override {
    for (size_t k = 0; k < index.size(); k++) this->set(index[k], index[k], value[k]);
  }
  void matrix_set_off_diagonal_values_blocked(const std::vector<int>& rows, const std::vector<int>& cols, const double& value) override {
    for (size_t k = 0; k < rows.size(); k++) this->set(rows[k], cols[k], value);
  }
  void matrix_set_off_diagonal_values_blocked(const std::vector<int>& rows, const std::vector<int>& cols, const std::vector<double>& value) override {
    for (size_t k = 0; k < rows.size(); k++) this->set(rows[k], cols[k], value[k]);
  }
  // rebuild the matrix without the entries of magnitude <= tolerance (PetscMatrix.cpp:755-830)
  void RemoveZeroEntries(double& tolerance) override {
    const std::vector<int64_t> rp = host_rowptr();
    const std::vector<int32_t> ci = host_col();
    const std::vector<double> v = host_val();
    std::vector<int64_t> nrp((size_t)_m + 1, 0);
    std::vector<int32_t> nci;
    std::vector<double> nv;
    for (int i = 0; i < _m; i++) {
      for (int64_t k = rp[i]; k < rp[i + 1]; k++)
        if (std::fabs(v[k]) > tolerance) { nci.push_back(ci[k]); nv.push_back(v[k]); }
      nrp[i + 1] = (int64_t)nci.size();
    }
    const int m = _m, n = _n;
    this->init_from_csr(m, n, nrp.data(), nci.data(), nv.data());
  }

  // ---- inspection (host mirror of the CSR, refreshed on demand) ----------------------------------
  int m() const override { return _m; }
  int n() const override { return _n; }
  int row_start() const override { return 0; }
  int row_stop() const override { return _m; }
  int64_t nnz() const { this->close(); return b2_csr_nnz(_A); }
  double operator()(const int i, const int j) const override {
    refresh_mirror();
    const int32_t* b = _mcol.data() + _mrp[i];
    const int32_t* e = _mcol.data() + _mrp[i + 1];
    const int32_t* p = std::lower_bound(b, e, (int32_t)j);
    return (p != e && *p == j) ? _mval[p - _mcol.data()] : 0.0;
  }
  int MatGetRowM(const int i, int* cols = NULL, double* vals = NULL) override {
    refresh_mirror();
    const int64_t s = _mrp[i], e = _mrp[i + 1];
    for (int64_t k = s; k < e; k++) {
      if (cols) cols[k - s] = _mcol[k];
      if (vals) vals[k - s] = _mval[k];
    }
    return (int)(e - s);
  }
  const std::vector<int64_t>& host_rowptr() const { refresh_mirror(); return _mrp; }
  const std::vector<int32_t>& host_col() const { refresh_mirror(); return _mcol; }
  const std::vector<double>& host_val() const { refresh_mirror(); return _mval; }
  double l1_norm() const override {      // max column sum
    refresh_mirror();
    std::vector<double> c((size_t)_n, 0.0);
    for (size_t k = 0; k < _mval.size(); k++) c[_mcol[k]] += std::fabs(_mval[k]);
    return c.empty() ? 0.0 : *std::max_element(c.begin(), c.end());
  }
  double linfty_norm() const override {  // max row sum
    refresh_mirror();
    double r = 0.0;
    for (int i = 0; i < _m; i++) {
      double s = 0.0;
      for (int64_t k = _mrp[i]; k < _mrp[i + 1]; k++) s += std::fabs(_mval[k]);
      r = std::max(r, s);
    }
    return r;
  }
  void print_personal(std::ostream& os = std::cout) const override {
    refresh_mirror();
    for (int i = 0; i < _m; i++) {
      os << "row " << i << ":";
      for (int64_t k = _mrp[i]; k < _mrp[i + 1]; k++) os << " (" << _mcol[k] << ", " << _mval[k] << ")";
      os << "\n";
    }
  }
  void print_hdf5(const std::string = "NULL") const override { B2_NOT_ON_PATH("print_hdf5"); }

  static const B200Matrix& cast(const SparseMatrix& M) {
    const B200Matrix* p = dynamic_cast<const B200Matrix*>(&M);
    if (!p) { std::fprintf(stderr, "femus_b200: operand is not a B200Matrix\n"); std::abort(); }
    return *p;
  }

 private:
  // take ownership of a finished device matrix (the previous one, its staging and its mirror go)
  void adopt(b2_csr* A, int m, int n) {
    this->clear();
    set_dims(m, n);
    _A = A;
    new_generation();
    _frozen = _closed = _is_initialized = true;
  }
  void new_generation() {
    static uint64_t counter = 0;
    _gen = ++counter;
  }
  struct Row {      // unsorted (column, value) pairs of one row while the pattern is still open
    std::vector<int32_t> col;
    std::vector<double> val;
    void reserve(size_t n) { col.reserve(n); val.reserve(n); }
    size_t find(int32_t c) const { return (size_t)(std::find(col.begin(), col.end(), c) - col.begin()); }
    void add(int32_t c, double v) { const size_t k = find(c); if (k == col.size()) { col.push_back(c); val.push_back(v); } else val[k] += v; }
    void set(int32_t c, double v) { const size_t k = find(c); if (k == col.size()) { col.push_back(c); val.push_back(v); } else val[k] = v; }
  };
  void set_dims(int m, int n) { _m = _m_l = m; _n = _n_l = n; _ml_start = 0; }
  void check_index(int i, int n, const char* what) const {
    if (i < 0 || i >= n) { std::fprintf(stderr, "femus_b200: B200Matrix: %s %d out of range [0,%d)\n", what, i, n); std::abort(); }
  }
  void drop_staging() {
    _blk.clear();
    _ins_rows.clear(); _ins_cols.clear(); _ins_vals.clear();
    _ins_ptr.assign(1, 0);
  }
  void add_block(const double* v, const int* rows, int nr, const int* cols, int nc) {
    _closed = false;
    if (!_frozen) {
      for (int i = 0; i < nr; i++) {
        check_index(rows[i], _m, "row");
        Row& r = _rows[rows[i]];
        for (int j = 0; j < nc; j++) { check_index(cols[j], _n, "column"); r.add(cols[j], v[i * nc + j]); }
      }
      return;
    }
    Blocks& b = _blk[std::make_pair(nr, nc)];     // blocks of one shape go to the device in one call
    b.rows.insert(b.rows.end(), rows, rows + nr);
    b.cols.insert(b.cols.end(), cols, cols + nc);
    b.vals.insert(b.vals.end(), v, v + (size_t)nr * nc);
  }
  void do_close() {
    if (!_is_initialized) { std::fprintf(stderr, "femus_b200: B200Matrix::close() before init()\n"); std::abort(); }
    if (!_frozen) {      // freeze: sort every row, upload pattern + values
      std::vector<int64_t> rp((size_t)_m + 1, 0);
      for (int i = 0; i < _m; i++) rp[i + 1] = rp[i] + (int64_t)_rows[i].col.size();
      std::vector<int32_t> col((size_t)rp[_m]);
      std::vector<double> val((size_t)rp[_m]);
      std::vector<size_t> perm;
      for (int i = 0; i < _m; i++) {
        const Row& r = _rows[i];
        perm.resize(r.col.size());
        for (size_t k = 0; k < perm.size(); k++) perm[k] = k;
        std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return r.col[a] < r.col[b]; });
        for (size_t k = 0; k < perm.size(); k++) { col[rp[i] + k] = r.col[perm[k]]; val[rp[i] + k] = r.val[perm[k]]; }
      }
      B2_ABORT_IF(b2_csr_create(B200Context::get(), _m, _n, rp.data(), col.data(), val.data(), &_A), "b2_csr_create");
      new_generation();
      _rows.clear();
      _rows.shrink_to_fit();
      _frozen = true;
    } else {
      for (auto& kv : _blk) {
        const Blocks& b = kv.second;
        const int nr = kv.first.first, nc = kv.first.second;
        B2_ABORT_IF(b2_csr_add_blocks(_A, (int64_t)(b.rows.size() / nr), nr, nc, b.rows.data(), b.cols.data(), b.vals.data()),
                    "b2_csr_add_blocks");
      }
      if (!_ins_rows.empty())
        B2_ABORT_IF(b2_csr_set_rows(_A, (int64_t)_ins_rows.size(), _ins_rows.data(), _ins_ptr.data(), _ins_cols.data(), _ins_vals.data()),
                    "b2_csr_set_rows");
      drop_staging();
    }
    _closed = true;
    touched();
  }
  void refresh_mirror() const {
    this->close();
    if (_mirror_ok) return;
    const int64_t z = b2_csr_nnz(_A);
    _mrp.resize((size_t)_m + 1);
    _mcol.resize((size_t)z);
    _mval.resize((size_t)z);
    B2_ABORT_IF(b2_csr_get(_A, _mrp.data(), _mcol.data(), _mval.data()), "b2_csr_get");
    _mirror_ok = true;
  }

  struct Blocks { std::vector<int32_t> rows, cols; std::vector<double> vals; };
  b2_csr* _A;
  uint64_t _gen = 0;
  bool _frozen, _closed;
  std::vector<Row> _rows;
  std::map<std::pair<int, int>, Blocks> _blk;
  std::vector<int32_t> _ins_rows, _ins_cols;
  std::vector<int64_t> _ins_ptr = std::vector<int64_t>(1, 0);
  std::vector<double> _ins_vals;
  mutable std::vector<int64_t> _mrp;
  mutable std::vector<int32_t> _mcol;
  mutable std::vector<double> _mval;
  mutable bool _mirror_ok;
};

// ---- NumericVector products (PetscVector.cpp:193-247) ---------------------------------------------
inline void B200Vector::matrix_mult(const NumericVector& x, const SparseMatrix& A) {
  const B200Matrix& M = B200Matrix::cast(A);
  M.close();
  B2_ABORT_IF(b2_csr_spmv(M.handle(), dev(x), _v), "b2_csr_spmv");
  touched();
}
inline void B200Vector::matrix_mult_transpose(const NumericVector& x, const SparseMatrix& A) {
  const B200Matrix& M = B200Matrix::cast(A);
  M.close();
  B2_ABORT_IF(b2_csr_spmv_t(M.handle(), dev(x), _v), "b2_csr_spmv_t");
  touched();
}
inline void B200Vector::resid(const NumericVector& rhs, const NumericVector& x, const SparseMatrix& A) {
  const B200Matrix& M = B200Matrix::cast(A);
  M.close();
  B2_ABORT_IF(b2_csr_resid(M.handle(), dev(rhs), dev(x), _v), "b2_csr_resid");
  touched();
}
inline void B200Vector::add_vector(const NumericVector& x, const SparseMatrix& A) {
  const B200Matrix& M = B200Matrix::cast(A);
  M.close();
  B2_ABORT_IF(b2_csr_spmv_add(M.handle(), dev(x), _v), "b2_csr_spmv_add");
  touched();
}

}  // namespace femus
