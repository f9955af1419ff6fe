// Stand-alone declaration of the two abstract algebra interfaces the B200 adapters implement,
// for builds WITHOUT the FEMuS source tree (tests, the GPU box).  With -DB2_WITH_FEMUS_HEADERS the
// adapters include FEMuS's own NumericVector.hpp / SparseMatrix.hpp instead and this file is unused;
// tests/test_adapters.py compiles both ways (the second only where /root/reference exists), which
// keeps the signatures below honest.  Interfaces mirrored (reference src/03_algebra):
//   00_vectors/NumericVector.hpp:51-353, 01_matrices/SparseMatrix.hpp:48-305,
//   00_enums/algebra/SolverPackageEnum.hpp, 00_enums/algebra/01_matrices/ParalleltypeEnum.hpp.
// Only the virtual surface is declared; helper bodies that FEMuS defines in its .cpp files
// (build(), compare(), subset norms, print, hdf5 readers) are not part of the mirror.
#pragma once
#include <cstdlib>
#include <iostream>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace femus {

enum SolverPackage { PETSC_SOLVERS = 0, TRILINOS_SOLVERS, INVALID_SOLVER_PACKAGE };
enum ParallelType { AUTOMATIC = 0, SERIAL, PARALLEL, GHOSTED, INVALID_PARALLELIZATION };

class DenseVector;
class DenseSubVector;
class DenseMatrix;
class Graph;
class SparseMatrix;

class NumericVector {
 public:
  explicit NumericVector(const ParallelType t = AUTOMATIC) : _is_closed(false), _is_initialized(false), _type(t) {}
  virtual ~NumericVector() {}
  virtual std::unique_ptr<NumericVector> clone() const = 0;
  virtual void clear() { _is_closed = false; _is_initialized = false; }
  virtual void close() = 0;
  virtual void closeWithMinValues() = 0;
  virtual void init(const int, const int, const bool = false, const ParallelType = AUTOMATIC) = 0;
  virtual void init(const int, const bool = false, const ParallelType = AUTOMATIC) = 0;
  virtual void init(const int, const int, const std::vector<int>&, const bool = false, const ParallelType = AUTOMATIC) = 0;
  virtual void init(const NumericVector& other, const bool fast = false) = 0;
  virtual void set(const int i, const double value) = 0;
  virtual void add(const int i, const double value) = 0;
  virtual void zero() = 0;
  virtual NumericVector& operator=(const double s) = 0;
  virtual NumericVector& operator=(const NumericVector& V) = 0;
  virtual NumericVector& operator=(const std::vector<double>& v) = 0;
  virtual void insert(const std::vector<double>& v, const std::vector<int>& dof_indices) = 0;
  virtual void insert(const NumericVector& V, const std::vector<int>& dof_indices) = 0;
  virtual void insert(const DenseVector& V, const std::vector<int>& dof_indices) = 0;
  virtual void insert(const DenseSubVector& V, const std::vector<int>& dof_indices) = 0;
  virtual bool initialized() const { return _is_initialized; }
  virtual bool closed() const { return _is_closed; }
  ParallelType type() const { return _type; }
  virtual double min() const = 0;
  virtual double max() const = 0;
  virtual double sum() const = 0;
  virtual double l1_norm() const = 0;
  virtual double l2_norm() const = 0;
  virtual double linfty_norm() const = 0;
  virtual int size() const = 0;
  virtual int local_size() const = 0;
  virtual int first_local_index() const = 0;
  virtual int last_local_index() const = 0;
  virtual double operator()(const int i) const = 0;
  virtual double el(const int i) const { return (*this)(i); }
  virtual void get(const std::vector<int>& index, std::vector<double>& values) const = 0;
  virtual NumericVector& operator+=(const NumericVector& V) = 0;
  virtual NumericVector& operator-=(const NumericVector& V) = 0;
  NumericVector& operator*=(const double a) { this->scale(a); return *this; }
  NumericVector& operator/=(const double a) { this->scale(1. / a); return *this; }
  virtual void add(const double s) = 0;
  virtual void add(const NumericVector& V) = 0;
  virtual void add(const double a, const NumericVector& v) = 0;
  virtual void add_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof_indices) = 0;
  virtual void add_vector_blocked(const std::vector<double>& v, const std::vector<unsigned>& dof_indices) = 0;
  virtual void insert_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof_indices) = 0;
  virtual void add_vector(const std::vector<double>& v, const std::vector<int>& dof_indices) = 0;
  virtual void add_vector(const NumericVector& V, const std::vector<int>& dof_indices) = 0;
  virtual void add_vector(const NumericVector& x, const SparseMatrix& A) = 0;
  virtual void resid(const NumericVector& rhs, const NumericVector& x, const SparseMatrix& A) = 0;
  virtual void matrix_mult(const NumericVector& x, const SparseMatrix& A) = 0;
  virtual void matrix_mult_transpose(const NumericVector& x, const SparseMatrix& A) = 0;
  virtual void add_vector(const DenseVector& V, const std::vector<unsigned int>& dof_indices) = 0;
  virtual void scale(const double factor) = 0;
  virtual void abs() = 0;
  virtual double dot(const NumericVector&) const = 0;
  virtual void swap(NumericVector& v) {
    std::swap(_is_closed, v._is_closed);
    std::swap(_is_initialized, v._is_initialized);
    std::swap(_type, v._type);
  }
  virtual void localize(std::vector<double>& v_local) const = 0;
  virtual void localize(NumericVector& v_local) const = 0;
  virtual void localize(NumericVector& v_local, const std::vector<int>& send_list) const = 0;
  virtual void localize(const int first_local_idx, const int last_local_idx, const std::vector<int>& send_list) = 0;
  virtual void localize_to_one(std::vector<double>& v_local, const int proc_id = 0) const = 0;
  virtual void localize_to_all(std::vector<double>& v_local) const = 0;
  virtual void pointwise_mult(const NumericVector& vec1, const NumericVector& vec2) = 0;
  virtual void BinaryPrint(const char*) { std::abort(); }
  virtual void BinaryLoad(const char*) { std::abort(); }

 protected:
  bool _is_closed;
  bool _is_initialized;
  ParallelType _type;
};

class SparseMatrix {
 public:
  SparseMatrix() : _is_initialized(false) {}
  virtual ~SparseMatrix() {}
  virtual void clear() = 0;
  virtual void init(const int m, const int n, const int m_l, const int n_l, const std::vector<int>& n_nz,
                    const std::vector<int>& n_oz) = 0;
  virtual void init(const int nr, const int nc, const std::vector<SparseMatrix*>& P) = 0;
  virtual void set(const int i, const int j, const double value) = 0;
  virtual void add(const int i, const int j, const double value) = 0;
  virtual void zero() = 0;
  virtual void close() const = 0;
  virtual double operator()(const int i, const int j) const = 0;
  virtual int MatGetRowM(const int i_val, int* cols = NULL, double* vals = NULL) = 0;
  virtual void RemoveZeroEntries(double& tolerance) = 0;
  virtual bool initialized() const { return _is_initialized; }
  virtual bool closed() const = 0;
  virtual void update_sparsity_pattern_old(const Graph&) = 0;
  virtual void update_sparsity_pattern(const Graph&) = 0;
  virtual void update_sparsity_pattern(int m, int n, int m_l, int n_l, const std::vector<int> n_oz,
                                       const std::vector<int> n_nz) = 0;
  virtual int m() const = 0;
  virtual int n() const = 0;
  virtual int row_start() const = 0;
  virtual int row_stop() const = 0;
  virtual void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& rows, const std::vector<unsigned int>& cols) = 0;
  virtual void add_matrix(const DenseMatrix& dm, const std::vector<unsigned int>& dof_indices) = 0;
  virtual void insert_row(const int row, const int ncols, const std::vector<int>& cols, double* values) = 0;
  virtual void add_matrix_blocked(const std::vector<double>& mat_value, const std::vector<int>& rows, const std::vector<int>& cols) = 0;
  virtual void add_matrix_blocked(const std::vector<double>& mat_value, const std::vector<unsigned>& rows,
                                  const std::vector<unsigned>& cols) = 0;
  virtual void matrix_set_off_diagonal_values_blocked(const std::vector<int>& index_rows, const std::vector<int>& index_cols,
                                                      const double& value) = 0;
  virtual void matrix_set_off_diagonal_values_blocked(const std::vector<int>& index_rows, const std::vector<int>& index_cols,
                                                      const std::vector<double>& value) = 0;
  virtual void matrix_add(const double a_in, SparseMatrix& X_in, const char pattern[]) = 0;
  virtual void matrix_PtAP(const SparseMatrix& mat_P, const SparseMatrix& mat_A, const bool& reuse) = 0;
  virtual void matrix_ABC(const SparseMatrix& mat_A, const SparseMatrix& mat_B, const SparseMatrix& mat_C, const bool& reuse) = 0;
  virtual void matrix_RightMatMult(const SparseMatrix& mat_A) = 0;
  virtual void matrix_LeftMatMult(const SparseMatrix& mat_A) = 0;
  virtual void matrix_get_diagonal_values(const std::vector<int>& index, std::vector<double>& value) const = 0;
  virtual void matrix_set_diagonal_values(NumericVector& D) = 0;
  virtual void matrix_set_diagonal_values(const std::vector<int>& index, const double& value) = 0;
  virtual void matrix_set_diagonal_values(const std::vector<int>& index, const std::vector<double>& value) = 0;
  virtual void add(const double c, SparseMatrix& B) = 0;
  virtual double l1_norm() const = 0;
  virtual double linfty_norm() const = 0;
  virtual void get_diagonal(NumericVector& dest) const = 0;
  virtual void get_transpose(SparseMatrix& dest) const = 0;
  virtual void mat_zero_rows(const std::vector<int>& index, const double& diagonal_value) const = 0;
  virtual void print_personal(std::ostream& os = std::cout) const = 0;
  virtual void print_hdf5(const std::string name = "NULL") const = 0;

 protected:
  int _m, _n, _m_l, _n_l, _ml_start;
  bool _is_initialized;
};

}  // namespace femus
