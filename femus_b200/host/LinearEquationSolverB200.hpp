// LinearEquationSolverB200: the level solver of FEMuS on the femus_b200 multigrid.  Drop-in for
// LinearEquationSolverPetsc (reference src/08_algebra.../03_solvers_with_preconditioner/
// LinearEquationSolverPetsc.{hpp,cpp}): same public methods, argument meaning and call order
// (MGInit on the finest level, MGSetLevel on every level from 0 up, MGSolve on the finest,
// MGClear), same members (_KK, _RES, _EPS, _EPSC, _RESC of LinearEquation.hpp:116-125).
//
// FEMuS's abstract base LinearEquationSolver derives from LinearEquation, whose constructor needs a
// Mesh and a Solution (PETSc/MPI/HDF5 headers), so this class can only derive from it inside a full
// FEMuS build: INTEGRATION.md gives the two-line change.  Stand-alone it carries the same surface
// over B200Matrix / B200Vector and a Dirichlet flag vector in place of Solution::_Bdc.
//
// As in the reference the FINEST level owns the multigrid object (there: the PCMG inside the
// finest KSP, reached by the coarser levels through LinSolver->GetKSP(), LinearEquationSolverPetsc.cpp:230).
#pragma once
#include "B200Matrix.hpp"
#include "SystemLayout.hpp"

namespace femus {

enum MgSmootherType { FULL = 0, MULTIPLICATIVE, ADDITIVE, KASKADE };      // 00_enums MgTypeEnum.hpp
enum B200SolverType { RICHARDSON_B200 = 0, CHEBYSHEV_B200, PREONLY_B200, GMRES_B200 }; // the subset of SolvertypeEnum.hpp in scope

class LinearEquationSolverB200 {
 public:
  LinearEquationSolverB200(const unsigned& igrid) : _KK(nullptr), _RES(nullptr), _EPS(nullptr), _EPSC(nullptr), _RESC(nullptr),
        _level(igrid), _mg(nullptr), _levelMax(0), _richardsonScaleFactor(0.5), _rtol(1.e-5), _abstol(1.e-50), _dtol(1.e+5),
        _maxits(1000), _restart(30), _bdcIndexIsInitialized(false) {}
  virtual ~LinearEquationSolverB200() { this->MGClear(); this->DeletePde(); }

  // ---- LinearEquation::InitPde / DeletePde (LinearEquation.cpp:196-405): level matrix + vectors.
  // dof[nel][nve] are the system dofs of the level's elements (GetSystemDof), bdc[ndofs] the flags
  // of MultiLevelSolution::GenerateBdc (< 1.5: Dirichlet row).
  void InitPde(const int ndofs, const int64_t nel, const int nve, const int32_t* dof, const std::vector<double>& bdc) {
    this->DeletePde();
    _KK = new B200Matrix;
    _KK->init_from_elements(ndofs, nel, nve, dof);
    _RES = new B200Vector(ndofs);
    _EPS = new B200Vector(ndofs);
    _EPSC = new B200Vector(ndofs);
    _RESC = new B200Vector(ndofs);
    _bdc = bdc;
    _bdcIndexIsInitialized = false;
  }
  // The same for a system of several Lagrange variables (InitPde with _SolPdeIndex.size() > 1): rows
  // [rank][variable][dof] (KKoffset, :211-237), pattern of GetSparsityPatternSize with the coupling table
  // `pattern` (nvars x nvars, NULL = every pair, :407-548); bdc in system numbering (femus_b200::SystemBdc).
  // With this the UNCHANGED multi-variable callbacks (add_matrix_blocked of the Jacobian blocks on GetSystemDof
  // rows) assemble into the device matrix through the staged path, or b2_stokes / b2_ns write it in place.
  void InitPdeSystem(const femus_b200::MeshLevel& msh, const std::vector<int>& families, const std::vector<double>& bdc,
                     const uint8_t* pattern = nullptr) {
    this->DeletePde();
    const femus_b200::SystemLayout sys(msh, families);
    const femus_b200::HostCsr pat = femus_b200::BuildSystemSparsity(msh, sys, pattern);
    const int n = (int)sys.size();
    if ((size_t)n != bdc.size()) { std::fprintf(stderr, "femus_b200: InitPdeSystem: %zu Dirichlet flags for %d rows\n", bdc.size(), n); std::abort(); }
    _KK = new B200Matrix;
    _KK->init_from_csr(n, n, pat.rowptr.data(), pat.col.data(), nullptr);
    _RES = new B200Vector(n);
    _EPS = new B200Vector(n);
    _EPSC = new B200Vector(n);
    _RESC = new B200Vector(n);
    _bdc = bdc;
    _bdcIndexIsInitialized = false;
  }
  void DeletePde() {
    this->OnDeletePde();      // objects that BORROW _KK's handle (the subclass's block smoother) go first
    if (_coarseSolver) b2_schwarz_destroy(_coarseSolver);
    _coarseSolver = nullptr;
    delete _KK; delete _RES; delete _EPS; delete _EPSC; delete _RESC;
    _KK = nullptr;
    _RES = _EPS = _EPSC = _RESC = nullptr;
  }
  void SetResZero() { _RES->zero(); }
  void SetEpsZero() { _EPS->zero(); _EPSC->zero(); }

  // ---- LinearEquationSolver surface -----------------------------------------------------------
  void SetTolerances(const double& rtol, const double& atol, const double& divtol, const unsigned& maxits, const unsigned& restart) {
    _rtol = rtol; _abstol = atol; _dtol = divtol; _maxits = maxits; _restart = restart;
  }
  void SetRichardsonScaleFactor(const double& richardsonScaleFactor) { _richardsonScaleFactor = richardsonScaleFactor; }
  void set_solver_type(const B200SolverType st) { _levelSolverType = st; }

  // MGInit (LinearEquationSolverPetsc.cpp:185-209): called on the finest level's solver
  void MGInit(const MgSmootherType& mg_smoother_type, const unsigned& levelMax, const B200SolverType& mgSolverType) {
    if (mg_smoother_type != MULTIPLICATIVE) {
      std::fprintf(stderr, "femus_b200: MGInit: only the multiplicative V-cycle is implemented\n");
      std::abort();
    }
    this->MGClear();
    _levelMax = levelMax;
    _mgSolverType = mgSolverType;
    B2_ABORT_IF(b2_mg_create(B200Context::get(), (int)levelMax, &_mg), "b2_mg_create");
    // level 0: the reference runs PREONLY + LU (MUMPS); here Jacobi-PCG to a tight relative residual
    B2_ABORT_IF(b2_mg_set_coarse(_mg, 1.e-14, 10000), "b2_mg_set_coarse");
  }
  void MGClear() {
    if (_mg) b2_mg_destroy(_mg);
    _mg = nullptr;
  }
  // level 0 only: solve the coarsest system DIRECTLY (the reference's PREONLY + MLU_PRECOND there,
  // LinearEquationSolverPetsc.hpp:128-151) instead of the Jacobi-PCG -- required for indefinite systems
  // (velocity-pressure); at most 4096 rows.  Call before MGSetLevel.
  void SetCoarseDirect(const bool on) { _coarseDirect = on; }
  // MGSetLevel (LinearEquationSolverPetsc.cpp:213-290): called on EVERY level's solver with the
  // finest solver as first argument; PP = prolongator from level-1 (NULL on level 0), RR unused
  // (the reference passes it but restricts with PP^T, :277).
  void MGSetLevel(LinearEquationSolverB200* LinSolver, const unsigned& levelMax, const std::vector<unsigned>& /*variable_to_be_solved*/,
                  SparseMatrix* PP, SparseMatrix* /*RR*/, const unsigned& npre, const unsigned& npost) {
    if (!LinSolver->_mg || levelMax + 1 != LinSolver->_levelMax) {
      std::fprintf(stderr, "femus_b200: MGSetLevel: MGInit was not called on the finest solver for %u levels\n", levelMax + 1);
      std::abort();
    }
    this->BuildBdcIndex();
    _KK->close();
    b2_csr* P = nullptr;
    if (_level > 0) {
      const B200Matrix& Pm = B200Matrix::cast(*PP);
      Pm.close();
      P = Pm.handle();
    }
    this->SetLevelSmoother(LinSolver->_mg);
    if (_level > 0)       // KSPSetType of the level: GMRES (the reference's default) or Richardson around the level's preconditioner
      B2_ABORT_IF(b2_mg_set_level_ksp(LinSolver->_mg, (int)_level, _levelSolverType == GMRES_B200 ? 1 : 0), "b2_mg_set_level_ksp");
    if (_level == 0 && _coarseDirect) {
      if (!_coarseSolver) {
        const int64_t n = _KK->m();
        const int64_t bp[2] = {0, n}, gp[2] = {0, 1};
        std::vector<int32_t> all((size_t)n);
        for (int64_t i = 0; i < n; i++) all[i] = (int32_t)i;
        const int32_t gb[1] = {0};
        B2_ABORT_IF(b2_schwarz_create(B200Context::get(), _KK->handle(), 1, bp, all.data(), 1, gp, gb, &_coarseSolver), "b2_schwarz_create");
      }
      B2_ABORT_IF(b2_mg_set_coarse_schwarz(LinSolver->_mg, _coarseSolver), "b2_mg_set_coarse_schwarz");
    }
    // SetPenalty (:428-436) happens inside: Dirichlet rows -> identity, pattern kept
    B2_ABORT_IF(b2_mg_set_level(LinSolver->_mg, (int)_level, _KK->handle(), P, _bdcIndex.data(), (int64_t)_bdcIndex.size(), (int)npre,
                                (int)npost, _richardsonScaleFactor),
                "b2_mg_set_level");
    _KK->touched();
  }
  // MGSolve (:294-353), outer solver PREONLY: ZerosBoundaryResiduals; EPSC = Vcycle(RES);
  // RESC = KK EPSC; RES -= RESC; EPS += EPSC
  void MGSolve(const bool /*ksp_clean*/) {
    if (!_mg) { std::fprintf(stderr, "femus_b200: MGSolve on a level that does not own the multigrid (call it on the finest)\n"); std::abort(); }
    B2_ABORT_IF(b2_mg_solve(_mg, _RES->handle(), _EPS->handle()), "b2_mg_solve");
    _RES->touched();
    _EPS->touched();
  }
  int CoarseIterations() const { return _mg ? b2_mg_coarse_iterations(_mg) : 0; }

  // BuildBdcIndex (:53-90): rows with Bdc < 1.5
  void BuildBdcIndex() {
    if (_bdcIndexIsInitialized) return;
    _bdcIndex.clear();
    for (size_t i = 0; i < _bdc.size(); i++)
      if (_bdc[i] < 1.5) _bdcIndex.push_back((int32_t)i);
    _bdcIndexIsInitialized = true;
  }
  const std::vector<int32_t>& BdcIndex() { this->BuildBdcIndex(); return _bdcIndex; }
  const std::vector<double>& Bdc() const { return _bdc; }
  b2_mg* mg() const { return _mg; }

  B200Matrix* _KK;
  B200Vector *_RES, *_EPS, *_EPSC, *_RESC;

 protected:
  // called before _KK is deleted (InitPde, InitPdeSystem, DeletePde).  Not reached from this class's destructor (the
  // subclass part is gone by then): a subclass that overrides it releases its objects in its own destructor as well.
  virtual void OnDeletePde() {}
  // level smoother: set_solver_type(RICHARDSON) -> Richardson(scale)+Jacobi, set_solver_type(CHEBYSHEV) ->
  // Chebyshev+Jacobi with the backend's stated eigenvalue bounds (KSPSetType switch, :452-536).  The reference's
  // subclasses override SetPreconditioner (LinearEquationSolverPetscAsm.cpp:266); here they override this.
  virtual void SetLevelSmoother(b2_mg* mg) {
    B2_ABORT_IF(b2_mg_set_smoother(mg, (int)_level, _levelSolverType == CHEBYSHEV_B200 ? 1 : 0, 0., 0.), "b2_mg_set_smoother");
  }
  unsigned _level;

 private:
  bool _coarseDirect = false;
  b2_schwarz* _coarseSolver = nullptr;
  b2_mg* _mg;
  unsigned _levelMax;
  B200SolverType _levelSolverType = RICHARDSON_B200, _mgSolverType = PREONLY_B200;
  double _richardsonScaleFactor, _rtol, _abstol, _dtol;
  unsigned _maxits, _restart;
  std::vector<double> _bdc;
  std::vector<int32_t> _bdcIndex;
  bool _bdcIndexIsInitialized;
};

}  // namespace femus
