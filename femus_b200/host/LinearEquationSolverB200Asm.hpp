// LinearEquationSolverB200Asm: the element-block (ASM / Vanka) level solver on the femus_b200 multigrid.  Drop-in for
// LinearEquationSolverPetscAsm (reference src/08_algebra.../03_solvers_with_preconditioner/petsc_asm/
// LinearEquationSolverPetscAsm.{hpp,cpp}): same setters (SetElementBlockNumber, SetNumberOfSchurVariables), and what
// the reference does in BuildASMIndex (:91-262) + SetPreconditioner (:266-340) happens in SetLevelSmoother: element
// blocks by MeshASMPartitioning::DoPartition, overlapping index sets, PCASM basic / multiplicative with exact block
// solves, wrapped by the level's Richardson iteration (b2_schwarz_*, b2_mg_set_level_schwarz).
//
// In scope: Lagrange variables, the last NSchurVar of them Schur variables -- one variable with
// SetNumberOfSchurVariables(0) is what 001_Poisson sets for "smoother": "asm" (main.cpp:248-249); velocity-pressure
// systems with the pressure as Schur variable give Vanka blocks.  The "All"-elements standard ASM
// (SetElementBlockNumber("All", overlap)) aborts.  Stand-alone the mesh of the level (GetMeshFromLinEq() in the reference) is handed in with
// SetMesh.
#pragma once
#include "AsmPartition.hpp"
#include "GeneralMesh.hpp"
#include "LinearEquationSolverB200.hpp"

namespace femus {

enum B200PrecondType { MLU_PRECOND_B200 = 0, SOR_PRECOND_B200, ILU_PRECOND_B200 };     // the subset of PrecondtypeEnum.hpp the blocks support

class LinearEquationSolverB200Asm : public LinearEquationSolverB200 {
 public:
  LinearEquationSolverB200Asm(const unsigned& igrid)
      : LinearEquationSolverB200(igrid), _msh(nullptr), _schwarz(nullptr), _sweepOrder(1), _NSchurVar(1), _standardASM(true),
        _indexIsInitialized(false) {
    _elementBlockNumber[0] = _elementBlockNumber[1] = _elementBlockNumber[2] = 1;
  }
  ~LinearEquationSolverB200Asm() override { this->ClearIndex(); }

  // LinearEquationSolverPetscAsm.hpp:52-68
  void SetElementBlockNumber(const unsigned& block_elemet_number) {
    _elementBlockNumber[0] = _elementBlockNumber[1] = _elementBlockNumber[2] = block_elemet_number;
    _standardASM = false;
    this->ClearIndex();
  }
  void SetElementBlockNumber(const char /*all*/[], const unsigned& /*overlap*/ = 1) {
    std::fprintf(stderr, "femus_b200: SetElementBlockNumber(\"All\", overlap): the standard PETSc ASM is not implemented\n");
    std::abort();
  }
  void SetNumberOfSchurVariables(const unsigned short& NSchurVar) { _NSchurVar = NSchurVar; }
  // LinearEquationSolver::set_preconditioner_type (what SetPreconditionerFineGrids sets, LinearImplicitSystem.cpp:1236-1245):
  // the solver of every block -- MLU_PRECOND: exact, SOR_PRECOND: one SSOR iteration (001_Poisson, main.cpp:242),
  // ILU_PRECOND: ILU(0) in the block's sorted dofs (the setting of most applications)
  void set_preconditioner_type(const B200PrecondType pt) { _blockPrecond = pt; }
  // the level's mesh and the family of the unknown (what GetMeshFromLinEq() / _SolType give the reference)
  void SetMesh(const femus_b200::MeshLevel* msh, const int family) {
    _msh = msh;
    _families.assign(1, family);
    this->ClearIndex();
  }
  // a system of several variables (families in _SolPdeIndex order), the last NSchurVar of them Schur variables
  void SetMesh(const femus_b200::MeshLevel* msh, const std::vector<int>& families) {
    _msh = msh;
    _families = families;
    this->ClearIndex();
  }
  // 0: sweep the blocks in the reference's order (dependency levels), 1: in coloured order (default)
  void SetSweepOrder(const int mode) { _sweepOrder = mode; this->ClearIndex(); }
  int64_t BlockNumber() const { return _index.nblocks(); }
  int64_t GroupNumber() const { return _schwarz ? b2_schwarz_groups(_schwarz) : 0; }

 protected:
  void OnDeletePde() override { this->ClearIndex(); }      // the b2_schwarz object borrows _KK->handle()
  void SetLevelSmoother(b2_mg* mg) override {
    if (_level == 0) { LinearEquationSolverB200::SetLevelSmoother(mg); return; }     // the coarsest level is solved, not smoothed
    if (_standardASM || !_msh || _families.empty() || _NSchurVar > _families.size()) {
      std::fprintf(stderr, "femus_b200: LinearEquationSolverB200Asm needs SetMesh, SetElementBlockNumber(n) and at most as many Schur variables as variables\n");
      std::abort();
    }
    if (!_indexIsInitialized) {           // BuildASMIndex, once per mesh (the reference's _bdcIndexIsInitialized gate, :42-49)
      using namespace femus_b200;
      try {
        const SystemLayout sys(*_msh, _families);
        _index = BuildAsmIndexSystem(*_msh, sys, (int)_NSchurVar, _elementBlockNumber[2], 0);
        const HostCsr pat = BuildSystemSparsity(*_msh, sys, nullptr);
        const int64_t nb = _index.nblocks();
        std::vector<int32_t> group((size_t)nb);
        const int64_t ng = AsmSchedule(pat.nrows, pat.rowptr.data(), pat.col.data(), nb, _index.overlap_ptr.data(), _index.overlap.data(),
                                       _sweepOrder, group.data());
        std::vector<int64_t> gptr((size_t)ng + 1, 0);
        for (int64_t b = 0; b < nb; b++) gptr[group[b] + 1]++;
        for (int64_t g = 0; g < ng; g++) gptr[g + 1] += gptr[g];
        std::vector<int64_t> fill(gptr.begin(), gptr.end() - 1);
        std::vector<int32_t> gblocks((size_t)nb);
        for (int64_t b = 0; b < nb; b++) gblocks[fill[group[b]]++] = (int32_t)b;       // stable: block order kept inside a group
        B2_ABORT_IF(b2_schwarz_create(B200Context::get(), _KK->handle(), nb, _index.overlap_ptr.data(), _index.overlap.data(), ng, gptr.data(),
                                      gblocks.data(), &_schwarz),
                    "b2_schwarz_create");
      } catch (const std::exception& e) {
        std::fprintf(stderr, "femus_b200: BuildASMIndex: %s\n", e.what());
        std::abort();
      }
      _indexIsInitialized = true;
    }
    B2_ABORT_IF(b2_schwarz_set_subsolver(_schwarz, _blockPrecond == SOR_PRECOND_B200 ? 1 : (_blockPrecond == ILU_PRECOND_B200 ? 2 : 0)), "b2_schwarz_set_subsolver");
    // SetPreconditioner: the numeric phase runs inside b2_mg_set_level, after the penalty
    B2_ABORT_IF(b2_mg_set_level_schwarz(mg, (int)_level, _schwarz), "b2_mg_set_level_schwarz");
  }

 private:
  void ClearIndex() {
    if (_schwarz) b2_schwarz_destroy(_schwarz);
    _schwarz = nullptr;
    _indexIsInitialized = false;
  }
  const femus_b200::MeshLevel* _msh;
  std::vector<int> _families;
  femus_b200::AsmIndex _index;
  b2_schwarz* _schwarz;
  int _sweepOrder;
  B200PrecondType _blockPrecond = MLU_PRECOND_B200;
  unsigned _elementBlockNumber[3];
  unsigned short _NSchurVar;
  bool _standardASM;
  bool _indexIsInitialized;
};

}  // namespace femus
