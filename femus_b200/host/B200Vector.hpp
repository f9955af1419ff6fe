// B200Vector: FEMuS NumericVector on the femus_b200 device library.  Drop-in for PetscVector
// (reference src/03_algebra/00_vectors/PetscVector.{hpp,cpp}) behind NumericVector::build():
// same method names, argument meaning and error behaviour (the reference aborts on error --
// CHKERRABORT; so do we, with the library's error text).
//
//   set()/add()/add_vector_blocked()/insert()  are STAGED on the host, like VecSetValues, and reach
//                                              the device in one batch at close() (PetscVector.hpp:595-612);
//   operator()(i) / get()                      read through a host mirror refreshed after every
//                                              modification (the reference pays VecGetArray per access,
//                                              PetscVector.hpp:727-759);
//   everything else is one call into the C ABI (include/femus_b200.h), on the library stream.
// One rank per GPU holds its LOCAL vector (all dofs of its own elements); the distributed layout is
// attached with set_halo() and completes interface entries / restricts reductions to owned entries.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#ifdef B2_WITH_FEMUS_HEADERS
#include "NumericVector.hpp"
#include "SparseMatrix.hpp"
#include "DenseVector.hpp"
#include "DenseSubvector.hpp"
#else
#include "femus_iface/AlgebraBase.hpp"
#endif
#include "../../include/femus_b200.h"

namespace femus {

#define B2_ABORT_IF(status, what)                                                            \
  do {                                                                                       \
    if (status) {                                                                            \
      std::fprintf(stderr, "femus_b200: %s failed: %s\n", what, b2_last_error());            \
      std::abort();                                                                          \
    }                                                                                        \
  } while (0)
#define B2_NOT_ON_PATH(what)                                                                                     \
  do {                                                                                                           \
    std::fprintf(stderr, "femus_b200: %s is not implemented by the B200 backend (outside the in-scope path)\n", what); \
    std::abort();                                                                                                \
  } while (0)

// Process-wide device context: the role of FemusInit / PetscInitialize (FemusInit.cpp:46-73).
class B200Context {
 public:
  static b2_ctx* get() {
    static B200Context inst;
    return inst._ctx;
  }

 private:
  B200Context() : _ctx(nullptr) {
    const char* dev = std::getenv("B2_DEVICE");
    B2_ABORT_IF(b2_ctx_create(dev ? std::atoi(dev) : 0, &_ctx), "b2_ctx_create");
  }
  ~B200Context() {}      // the driver reclaims the context at process exit (FemusInit::~FemusInit likewise only finalizes)
  b2_ctx* _ctx;
};

class B200Matrix;

class B200Vector : public NumericVector {
 public:
  explicit B200Vector(const ParallelType type = AUTOMATIC) : NumericVector(type), _v(nullptr), _n(0), _mode(0), _mirror_ok(false) {}
  explicit B200Vector(const int n, const ParallelType type = AUTOMATIC) : NumericVector(type), _v(nullptr), _n(0), _mode(0), _mirror_ok(false) {
    this->init(n, n, false, type);
  }
  ~B200Vector() { this->clear(); }

  static std::unique_ptr<NumericVector> build() { return std::unique_ptr<NumericVector>(new B200Vector); }

  b2_vec* handle() const { return _v; }

  // ---- lifetime ---------------------------------------------------------------------------
  void clear() override {
    if (_v) b2_vec_destroy(_v);
    _v = nullptr;
    _n = 0;
    _stage_idx.clear();
    _stage_val.clear();
    _mode = 0;
    _mirror_ok = false;
    _is_closed = false;
    _is_initialized = false;
  }
  void init(const int n, const int n_local, const bool /*fast*/ = false, const ParallelType type = AUTOMATIC) override {
    if (n_local != n) {
      std::fprintf(stderr, "femus_b200: B200Vector::init(N=%d, n_local=%d): one rank holds its whole local vector\n", n, n_local);
      std::abort();
    }
    this->clear();
    B2_ABORT_IF(b2_vec_create(B200Context::get(), n, &_v), "b2_vec_create");
    _n = n;
    _type = type == AUTOMATIC ? SERIAL : type;
    _is_initialized = true;
    _is_closed = true;          // a fresh vector is all zeros and assembled (PetscVector::init ends with zero())
  }
  void init(const int n, const bool fast = false, const ParallelType type = AUTOMATIC) override { this->init(n, n, fast, type); }
  void init(const int n, const int n_local, const std::vector<int>& /*ghost*/, const bool fast = false,
            const ParallelType type = AUTOMATIC) override {
    this->init(n, n_local, fast, type);     // ghost entries live in the local numbering already
  }
  void init(const NumericVector& other, const bool fast = false) override { this->init(other.size(), other.local_size(), fast, other.type()); }
  std::unique_ptr<NumericVector> clone() const override {
    std::unique_ptr<NumericVector> c(new B200Vector);
    c->init(*this, true);
    *c = *this;
    return c;
  }
  // distributed layout of this rank's local vector (b2_halo_create); NULL detaches
  void set_halo(const b2_halo* h) { B2_ABORT_IF(b2_vec_set_halo(_v, h), "b2_vec_set_halo"); }

  // ---- staged element access ----------------------------------------------------------------
  void set(const int i, const double value) override { stage(1, i, value); }
  void add(const int i, const double value) override { stage(2, i, value); }
  void add_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(2, dof[k], v[k]);
  }
  void add_vector_blocked(const std::vector<double>& v, const std::vector<unsigned>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(2, (int)dof[k], v[k]);
  }
  void insert_vector_blocked(const std::vector<double>& v, const std::vector<int>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(1, dof[k], v[k]);
  }
  void add_vector(const std::vector<double>& v, const std::vector<int>& dof) override { add_vector_blocked(v, dof); }
  void add_vector(const NumericVector& V, const std::vector<int>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(2, dof[k], V((int)k));
  }
  void insert(const std::vector<double>& v, const std::vector<int>& dof) override { insert_vector_blocked(v, dof); }
  void insert(const NumericVector& V, const std::vector<int>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(1, dof[k], V((int)k));
  }
#ifdef B2_WITH_FEMUS_HEADERS
  void insert(const DenseVector& V, const std::vector<int>& dof) override {      // PetscVector.cpp:347-359
    for (size_t k = 0; k < dof.size(); k++) stage(1, dof[k], V((unsigned)k));
  }
  void insert(const DenseSubVector& V, const std::vector<int>& dof) override {
    for (size_t k = 0; k < dof.size(); k++) stage(1, dof[k], V((unsigned)k));
  }
  void add_vector(const DenseVector& V, const std::vector<unsigned int>& dof) override {      // PetscVector.cpp:197-201
    for (size_t k = 0; k < dof.size(); k++) stage(2, (int)dof[k], V((unsigned)k));
  }
#else
  void insert(const DenseVector&, const std::vector<int>&) override { B2_NOT_ON_PATH("insert(DenseVector) (needs the FEMuS dense classes)"); }
  void insert(const DenseSubVector&, const std::vector<int>&) override { B2_NOT_ON_PATH("insert(DenseSubVector) (needs the FEMuS dense classes)"); }
  void add_vector(const DenseVector&, const std::vector<unsigned int>&) override { B2_NOT_ON_PATH("add_vector(DenseVector) (needs the FEMuS dense classes)"); }
#endif
  void close() override {
    if (!_stage_idx.empty()) {
      const int64_t n = (int64_t)_stage_idx.size();
      if (_mode == 1) B2_ABORT_IF(b2_vec_set_indexed(_v, _stage_idx.data(), _stage_val.data(), n), "b2_vec_set_indexed");
      else B2_ABORT_IF(b2_vec_add_indexed(_v, _stage_idx.data(), _stage_val.data(), n), "b2_vec_add_indexed");
      _stage_idx.clear();
      _stage_val.clear();
      _mirror_ok = false;
    }
    _mode = 0;
    _is_closed = true;
  }
  void closeWithMinValues() override { this->close(); }     // one copy of every entry per rank: nothing to minimise
  double operator()(const int i) const override {
    refresh_mirror();
    return _mirror[i];
  }
  void get(const std::vector<int>& index, std::vector<double>& values) const override {
    values.resize(index.size());
    if (!index.empty())
      B2_ABORT_IF(b2_vec_get_indexed(_v, index.data(), values.data(), (int64_t)index.size()), "b2_vec_get_indexed");
  }

  // ---- whole-vector operations ----------------------------------------------------------------
  void zero() override { B2_ABORT_IF(b2_vec_zero(_v), "b2_vec_zero"); touched(); }
  NumericVector& operator=(const double s) override { B2_ABORT_IF(b2_vec_fill(_v, s), "b2_vec_fill"); touched(); return *this; }
  NumericVector& operator=(const NumericVector& V) override {
    B2_ABORT_IF(b2_vec_copy(_v, dev(V)), "b2_vec_copy");
    touched();
    return *this;
  }
  B200Vector& operator=(const B200Vector& V) { B2_ABORT_IF(b2_vec_copy(_v, V._v), "b2_vec_copy"); touched(); return *this; }
  NumericVector& operator=(const std::vector<double>& v) override {
    B2_ABORT_IF(b2_vec_put(_v, v.data(), (int64_t)v.size()), "b2_vec_put");
    touched();
    return *this;
  }
  NumericVector& operator+=(const NumericVector& V) override { this->add(1., V); return *this; }
  NumericVector& operator-=(const NumericVector& V) override { this->add(-1., V); return *this; }
  void add(const double s) override { B2_ABORT_IF(b2_vec_add_scalar(_v, s), "b2_vec_add_scalar"); touched(); }
  void add(const NumericVector& V) override { this->add(1., V); }
  void add(const double a, const NumericVector& V) override { B2_ABORT_IF(b2_vec_axpy(_v, a, dev(V)), "b2_vec_axpy"); touched(); }
  void scale(const double factor) override { B2_ABORT_IF(b2_vec_scale(_v, factor), "b2_vec_scale"); touched(); }
  void abs() override { B2_ABORT_IF(b2_vec_abs(_v), "b2_vec_abs"); touched(); }
  void pointwise_mult(const NumericVector& a, const NumericVector& b) override {
    B2_ABORT_IF(b2_vec_pointwise_mult(_v, dev(a), dev(b)), "b2_vec_pointwise_mult");
    touched();
  }
  double dot(const NumericVector& V) const override { double r; B2_ABORT_IF(b2_vec_dot(_v, dev(V), &r), "b2_vec_dot"); return r; }
  double l1_norm() const override { return norm(1); }
  double l2_norm() const override { return norm(2); }
  double linfty_norm() const override { return norm(0); }
  double sum() const override { double r; B2_ABORT_IF(b2_vec_sum(_v, &r), "b2_vec_sum"); return r; }
  double min() const override { double a, b; B2_ABORT_IF(b2_vec_minmax(_v, &a, &b), "b2_vec_minmax"); return a; }
  double max() const override { double a, b; B2_ABORT_IF(b2_vec_minmax(_v, &a, &b), "b2_vec_minmax"); return b; }
  int size() const override { return _n; }
  int local_size() const override { return _n; }
  int first_local_index() const override { return 0; }
  int last_local_index() const override { return _n; }

  // ---- matrix-vector products (PetscVector.cpp:193-247); defined in B200Matrix.hpp ------------
  void add_vector(const NumericVector& x, const SparseMatrix& A) override;          // this += A x
  void resid(const NumericVector& rhs, const NumericVector& x, const SparseMatrix& A) override;   // this = rhs - A x
  void matrix_mult(const NumericVector& x, const SparseMatrix& A) override;         // this = A x
  void matrix_mult_transpose(const NumericVector& x, const SparseMatrix& A) override;   // this = A^T x

  // ---- gathers: every rank already holds its local vector ---------------------------------------
  void localize(std::vector<double>& v_local) const override {
    v_local.resize(_n);
    if (_n) B2_ABORT_IF(b2_vec_get(_v, v_local.data(), _n), "b2_vec_get");
  }
  void localize(NumericVector& v_local) const override { v_local = *this; }
  void localize(NumericVector& v_local, const std::vector<int>&) const override { v_local = *this; }
  void localize(const int, const int, const std::vector<int>&) override {}
  void localize_to_one(std::vector<double>& v_local, const int = 0) const override { this->localize(v_local); }
  void localize_to_all(std::vector<double>& v_local) const override { this->localize(v_local); }
  void swap(NumericVector& other) override {
    B200Vector& o = dynamic_cast<B200Vector&>(other);
    NumericVector::swap(other);
    std::swap(_v, o._v);
    std::swap(_n, o._n);
    _mirror_ok = o._mirror_ok = false;
  }
  // raw fp64 dump with a one-integer header, the format of PetscVector::BinaryPrint's use in restart files
  void BinaryPrint(const char* fileName) override {
    std::vector<double> h;
    this->localize(h);
    std::ofstream f(fileName, std::ios::binary);
    const int64_t n = _n;
    f.write(reinterpret_cast<const char*>(&n), sizeof(n));
    f.write(reinterpret_cast<const char*>(h.data()), (std::streamsize)(h.size() * sizeof(double)));
  }
  void BinaryLoad(const char* fileName) override {
    std::ifstream f(fileName, std::ios::binary);
    int64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), sizeof(n));
    if (!f || n != _n) { std::fprintf(stderr, "femus_b200: BinaryLoad(%s): size mismatch\n", fileName); std::abort(); }
    std::vector<double> h((size_t)n);
    f.read(reinterpret_cast<char*>(h.data()), (std::streamsize)(h.size() * sizeof(double)));
    *this = h;
  }

  static b2_vec* dev(const NumericVector& V) {
    const B200Vector* p = dynamic_cast<const B200Vector*>(&V);
    if (!p || !p->_v) { std::fprintf(stderr, "femus_b200: operand is not an initialised B200Vector\n"); std::abort(); }
    return p->_v;
  }
  void touched() { _mirror_ok = false; }       // called by whoever writes the device data behind our back

 private:
  void stage(int mode, int i, double value) {
    if (_mode && _mode != mode) {     // PETSc: "You have already inserted values, cannot now add" without an intervening close()
      std::fprintf(stderr, "femus_b200: B200Vector: set() and add() mixed without close()\n");
      std::abort();
    }
    if (i < 0 || i >= _n) { std::fprintf(stderr, "femus_b200: B200Vector: index %d out of range [0,%d)\n", i, _n); std::abort(); }
    _mode = mode;
    _stage_idx.push_back(i);
    _stage_val.push_back(value);
    _is_closed = false;
  }
  double norm(int kind) const { double r; B2_ABORT_IF(b2_vec_norm(_v, kind, &r), "b2_vec_norm"); return r; }
  void refresh_mirror() const {
    if (_mirror_ok) return;
    _mirror.resize(_n);
    if (_n) B2_ABORT_IF(b2_vec_get(_v, _mirror.data(), _n), "b2_vec_get");
    _mirror_ok = true;
  }

  b2_vec* _v;
  int _n;
  int _mode;                         // 0 none, 1 INSERT_VALUES, 2 ADD_VALUES staged
  std::vector<int32_t> _stage_idx;
  std::vector<double> _stage_val;
  mutable std::vector<double> _mirror;
  mutable bool _mirror_ok;
};

}  // namespace femus
