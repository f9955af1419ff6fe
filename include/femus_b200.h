/* femus_b200 -- C ABI of the B200 (sm_100a) backend for the FEMuS assembly + geometric-multigrid
 * hot path.  This is the drop-in boundary: everything the FEMuS algebra plugin surface
 * (NumericVector / SparseMatrix / LinearEquationSolver, reference src/03_algebra and
 * src/08_algebra.../03_solvers_with_preconditioner) needs from a backend, as plain C.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b2_last_error() gives the text.
 *     (The reference aborts on error -- PetscMacro CHKERRABORT; the C++ adapters in
 *     femus_b200/host/ turn a non-zero status into abort() to keep that behaviour.)
 *   - handles are opaque and owned by the library; host arrays passed in are copied.
 *   - indices: 32-bit columns / dof ids (the reference asserts sizeof(PetscInt)==sizeof(int),
 *     PetscVector.hpp:536), 64-bit row pointers (256^3 Hex27 has 8.6e9 nonzeros).
 *   - all device work is issued on the context's stream; calls that return a scalar or copy to
 *     the host synchronise that stream, nothing else does.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Each group names the reference interface it replaces (paths relative to the reference root).
 */
#ifndef FEMUS_B200_H
#define FEMUS_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_ctx b2_ctx;
typedef struct b2_vec b2_vec;
typedef struct b2_csr b2_csr;
typedef struct b2_mesh b2_mesh;
typedef struct b2_asm b2_asm;
typedef struct b2_mg b2_mg;
typedef struct b2_galerkin b2_galerkin;
typedef struct b2_halo b2_halo;
typedef struct b2_schwarz b2_schwarz;
typedef struct b2_stokes b2_stokes;

const char* b2_last_error(void);
int b2_version(void);

/* ---- context: replaces FemusInit / PetscInitialize + MPI_COMM_WORLD
 *      (src/00_utils/00_application_initialization/FemusInit.cpp:46-73) ------------------- */
int b2_ctx_create(int device, b2_ctx** out);
int b2_ctx_destroy(b2_ctx* c);
int b2_ctx_sync(b2_ctx* c);
void* b2_ctx_stream(b2_ctx* c);                 /* cudaStream_t the library launches on */
int b2_ctx_device(b2_ctx* c);
/* one rank per GPU; id = 128-byte ncclUniqueId made by b2_nccl_unique_id on rank 0 and sent to
 * the other ranks by the launcher (torch.distributed / MPI / a file). */
int b2_nccl_unique_id(void* id128);
int b2_ctx_comm_init(b2_ctx* c, int nranks, int rank, const void* id128);
int b2_ctx_nranks(b2_ctx* c);
int b2_ctx_rank(b2_ctx* c);
/* bytes currently allocated by the library on the device */
int64_t b2_ctx_bytes_in_use(b2_ctx* c);
/* number of kernels launched by the library since the last reset */
int64_t b2_ctx_launch_count(b2_ctx* c, int reset);
/* CUDA-event timing on the library stream: b2_timer_start / b2_timer_stop_ms */
int b2_timer_start(b2_ctx* c);
int b2_timer_stop_ms(b2_ctx* c, double* ms);
/* per-launch CUDA-event timing of the SpMV family (tag: the b2_csr) and of the assembly kernel
 * (tag: the b2_asm); b2_ctx_profile_read sums and consumes the records of one tag */
int b2_ctx_profile(b2_ctx* c, int on);
int b2_ctx_profile_only(b2_ctx* c, const void* handle);   /* time only this handle's launches */
int b2_ctx_profile_read(b2_ctx* c, const void* handle, int* count, double* total_ms);
int b2_ctx_profile_clear(b2_ctx* c);
/* tuning knobs: "spmv_variant" = 0 (register-streaming SpMV) | 1 (TMA-staged ring, default) | 2 (staged +
 * software-pipelined gathers); "spmv_timing" = 1 makes y = A x print the consumer phase cycles of CTA 0;
 * "asm_variant" = 3 (triquadratic assembly by sum factorisation, default; plans whose tables are not tensor products
 * use 1) | 1 (FP64 tensor cores) | 0 (CUDA-core register tiles) | 2 (the table-driven kernel of the non-hexahedral
 * families, also for hexahedra) */
int b2_ctx_set_option(b2_ctx* c, const char* name, int value);
/* measured issue-rate peak of mma.sync.m8n8k4.f64 on this device, TFLOP/s (a few ms of DMMA chains) */
int b2_ctx_measure_fp64_tensor(b2_ctx* c, double* tflops);
/* the same for DFMA on the CUDA cores (the pipe of the sum-factorised assembly kernel) */
int b2_ctx_measure_fp64_fma(b2_ctx* c, double* tflops);
/* write 256 MiB of device memory (> L2 size) to evict the L2 between timed iterations */
int b2_ctx_flush_l2(b2_ctx* c);

/* ---- vectors: replaces PetscVector (src/03_algebra/00_vectors/PetscVector.{hpp,cpp}) --------
 * A vector has n_local owned entries followed by n_ghost halo entries (0 in serial). */
int b2_vec_create(b2_ctx* c, int64_t n, b2_vec** out);                 /* NumericVector::init */
int b2_vec_destroy(b2_vec* v);
int64_t b2_vec_size(const b2_vec* v);
void* b2_vec_device_ptr(b2_vec* v);
int b2_vec_zero(b2_vec* v);                                            /* zero()      VecSet(0)        PetscVector.cpp:361 */
int b2_vec_fill(b2_vec* v, double a);                                  /* operator=(double)            :449 */
int b2_vec_put(b2_vec* v, const double* host, int64_t n);              /* operator=(std::vector)       :477 */
int b2_vec_get(const b2_vec* v, double* host, int64_t n);              /* localize / get               :627 */
/* asynchronous variants on the library stream (host buffer should be pinned) */
int b2_vec_put_async(b2_vec* v, const double* host, int64_t n);
int b2_vec_get_async(const b2_vec* v, double* host, int64_t n);
/* Overlapped input upload (double buffering): b2_ctx_open_copies orders the copy stream after the
 * compute enqueued so far, b2_vec_prefetch / b2_mesh_prefetch copy on the copy stream into buffers the
 * running step does not read, b2_ctx_join_copies orders the compute stream after those copies. */
int b2_ctx_open_copies(b2_ctx* c);
int b2_ctx_join_copies(b2_ctx* c);
/* b2_ctx_mark_copies closes a batch of prefetches; b2_ctx_wait_marked makes the compute stream wait for
 * that batch only, so that a later b2_vec_fetch keeps overlapping the next step */
int b2_ctx_mark_copies(b2_ctx* c);
int b2_ctx_wait_marked(b2_ctx* c);
int b2_vec_prefetch(b2_vec* v, const double* host, int64_t n);
/* device -> host on the copy stream, after the compute enqueued so far (overlaps the next step) */
int b2_vec_fetch(const b2_vec* v, double* host, int64_t n);
int b2_vec_copy(b2_vec* dst, const b2_vec* src);                       /* operator=(NumericVector)     :429 */
int b2_vec_axpy(b2_vec* y, double a, const b2_vec* x);                 /* add(a,V)    VecAXPY          :303 */
int b2_vec_aypx(b2_vec* y, double a, const b2_vec* x);                 /* y = x + a y VecAYPX */
int b2_vec_scale(b2_vec* v, double a);                                 /* scale       VecScale         :378 */
int b2_vec_abs(b2_vec* v);                                             /* abs         VecAbs           :387 */
int b2_vec_add_scalar(b2_vec* v, double a);                            /* add(double) VecShift         :283 */
int b2_vec_pointwise_mult(b2_vec* w, const b2_vec* x, const b2_vec* y);/* pointwise_mult               :782 */
int b2_vec_dot(const b2_vec* x, const b2_vec* y, double* out);         /* dot         VecDot           :399 */
int b2_vec_norm(const b2_vec* x, int kind /*1: l1, 2: l2, 0: linf*/, double* out); /* l1/l2/linfty_norm :43-87 */
int b2_vec_sum(const b2_vec* x, double* out);                          /* sum         VecSum */
int b2_vec_minmax(const b2_vec* x, double* mn, double* mx);            /* min/max     PetscVector.hpp:777-796 */
/* staged set()/add() of PetscVector (VecSetValues INSERT/ADD, PetscVector.cpp:96-141) flushed as
 * one batch: idx/vals are host arrays in CALL ORDER; an index may repeat -- the last value set stays, added
 * values are summed in the order given (what VecSetValues does on local entries; MultiLevelSolution::GenerateBdc
 * relies on it: set(j, 2.) ... set(idof, 1.) ... set(idof, 0.) on one vector before close()). */
int b2_vec_set_indexed(b2_vec* v, const int32_t* idx, const double* vals, int64_t n);
int b2_vec_add_indexed(b2_vec* v, const int32_t* idx, const double* vals, int64_t n);
int b2_vec_fill_indexed(b2_vec* v, const int32_t* idx, int64_t n, double a); /* ZerosBoundaryResiduals LinearEquationSolverPetsc.cpp:417-424 */
int b2_vec_get_indexed(const b2_vec* v, const int32_t* idx, double* vals, int64_t n); /* get(idx,vals) */
/* dst[i] = mask[i] > thr ? src[i] : 0  -- Solution::UpdateRes (Solution.cpp:595-628) */
int b2_vec_copy_masked(b2_vec* dst, const b2_vec* src, const b2_vec* mask, double thr);

/* ---- distributed layout: replaces the MPI side of PetscVector/PetscMatrix -- off-process
 *      ADD_VALUES assembly (PetscVector.cpp:132-141, PetscMatrix.cpp:699-729 + close()), ghost update
 *      (PetscVector.hpp:604-609) and owned-entry reductions.  One rank per GPU holds all dofs of its
 *      own elements; dofs on the partition interface are held by several ranks.
 *      local_idx[n_if]  : local dof of interface entry k,
 *      packed_pos[n_if] : its position in the packed interface vector of n_packed entries, the same
 *                         on every rank that holds the dof,
 *      owned[n_local]   : 1 if this rank owns the dof (lowest rank holding it, Mesh.cpp:530-553),
 *      mult[n_local]    : number of ranks holding the dof (>= 1). */
int b2_halo_create(b2_ctx* c, int64_t n_local, int64_t n_if, const int32_t* local_idx, const int32_t* packed_pos,
                   int64_t n_packed, const uint8_t* owned, const uint8_t* mult, b2_halo** out);
int b2_halo_destroy(b2_halo* h);
/* Peer-memory form of the interface sum: ONE kernel instead of pack -> ncclAllReduce -> unpack (b2_halo.cu).  The thread
 * that owns an interface entry stores its value into the inbox of every other holder (remote stores over NVLink /
 * NVSwitch, 16-byte cells carrying the exchange number as their flag), waits for the holders' values in its own inbox
 * and sums in ascending rank order: all copies of a dof agree bit for bit, as after an allreduce; no fence, no collective.
 * Setup: every rank exports its inbox (b2_ctx_peer_export), the launcher all-gathers the 64-byte CUDA IPC handles, every
 * rank opens them (b2_ctx_peer_open); each layout then gets its per-entry holder lists (b2_halo_set_exchange;
 * femus_b200/dist.py derives them from the gathered lattice keys).  Option "halo_peer" 0 keeps the NCCL form.  The
 * coarse PCG runs the same exchange inside its persistent kernel (b2_cg.cu). */
int b2_ctx_peer_export(b2_ctx* c, int64_t slot_cells, void* handle64);
int b2_ctx_peer_open(b2_ctx* c, const void* handles /* [nranks][64] */);
int b2_ctx_peer_error(b2_ctx* c, int* err);
int b2_halo_set_exchange(b2_halo* h, const int64_t* hold_ptr, const int32_t* hold_rank, const int32_t* hold_pos, const int32_t* hold_spos);
int64_t b2_halo_owned_count(const b2_halo* h);
int64_t b2_halo_interface_count(const b2_halo* h);
/* v[interface] <- sum over the ranks holding each dof (pack, ncclAllReduce over NVLink, unpack) */
int b2_halo_sum(b2_halo* h, b2_vec* v);
/* attach a layout: dot / norms / sum of v then run over owned entries + allreduce (NULL detaches) */
int b2_vec_set_halo(b2_vec* v, const b2_halo* h);

/* ---- CSR matrices: replaces PetscMatrix (src/03_algebra/01_matrices/PetscMatrix.{hpp,cpp}) --- */
/* init(m,n,...,n_nz,n_oz) + pattern: rowptr[nrows+1] (int64), col[nnz] (int32, sorted per row);
 * vals may be NULL (zeros). */
int b2_csr_create(b2_ctx* c, int64_t nrows, int64_t ncols, const int64_t* rowptr, const int32_t* col,
                  const double* vals, b2_csr** out);
/* Pattern from element->dof lists (every (i,j) pair of every element, zeros included): replaces
 * LinearEquation::GetSparsityPatternSize + SparseMatrix::init (LinearEquation.cpp:407-548, 334-335).
 * dof[nel][nve] host array. */
int b2_csr_create_from_elements(b2_ctx* c, int64_t nrows, int64_t nel, int nve, const int32_t* dof, b2_csr** out);
int b2_csr_destroy(b2_csr* A);
int64_t b2_csr_nrows(const b2_csr* A);
int64_t b2_csr_ncols(const b2_csr* A);
int64_t b2_csr_nnz(const b2_csr* A);
int b2_csr_get(const b2_csr* A, int64_t* rowptr, int32_t* col, double* vals);   /* any may be NULL */
int b2_csr_put_vals(b2_csr* A, const double* vals);
int b2_csr_zero(b2_csr* A);                                             /* zero()  MatZeroEntries */
int b2_csr_copy_vals(b2_csr* dst, const b2_csr* src);                   /* same pattern */
/* add_matrix_blocked (PetscMatrix.cpp:699-729): nblk dense blocks, block b adds
 * vals[b][i*ncol+j] at (rows[b*nrow+i], cols[b*ncol+j]); host arrays, staged by the adapter
 * until close(). */
int b2_csr_add_blocks(b2_csr* A, int64_t nblk, int nrow, int ncol, const int32_t* rows, const int32_t* cols,
                      const double* vals);
/* insert_row (PetscMatrix.cpp:683-695), batched: row r gets vals at cols (INSERT). */
int b2_csr_set_rows(b2_csr* A, int64_t nset, const int32_t* rows, const int64_t* ptr, const int32_t* cols,
                    const double* vals);
/* mat_zero_rows / MatZeroRows keeping the pattern (PetscMatrix.cpp:1073-1077;
 * SetPenalty LinearEquationSolverPetsc.cpp:428-436) */
int b2_csr_zero_rows(b2_csr* A, const int32_t* rows, int64_t n, double diag);
/* zero the listed columns (ZeroInterpolatorDirichletNodes does it through two transposes,
 * LinearImplicitSystem.cpp:1090-1112) */
int b2_csr_zero_cols(b2_csr* A, const int32_t* cols, int64_t n);
int b2_csr_diag(const b2_csr* A, b2_vec* d);                            /* MatGetDiagonal PetscMatrix.cpp:1019-1027 */
int b2_csr_transpose(const b2_csr* A, b2_csr** At);                     /* get_transpose  :1031-1070 */
int b2_csr_spmv(const b2_csr* A, const b2_vec* x, b2_vec* y);           /* matrix_mult    PetscVector.cpp:203-212 */
int b2_csr_spmv_add(const b2_csr* A, const b2_vec* x, b2_vec* y);       /* MatMultAdd     :243-247 */
int b2_csr_spmv_t(const b2_csr* A, const b2_vec* x, b2_vec* y);         /* matrix_mult_transpose :219-228 */
int b2_csr_resid(const b2_csr* A, const b2_vec* b, const b2_vec* x, b2_vec* r); /* resid: r = b - A x  :232-247 */
/* one Richardson(omega)+Jacobi sweep  xout = xin + omega * dinv .* (b - A xin)
 * (KSPRICHARDSON + PCJACOBI, LinearEquationSolverPetsc.cpp:516-519, PetscPreconditioner.cpp:209-212) */
int b2_csr_jacobi_sweep(const b2_csr* A, const b2_vec* dinv, const b2_vec* b, const b2_vec* xin, b2_vec* xout,
                        double omega);
/* matrix_PtAP (PetscMatrix.cpp:733-751): C = P^T A P, numeric phase onto C's existing pattern
 * (the coarse element-coupling pattern). */
int b2_csr_ptap(const b2_csr* P, const b2_csr* A, b2_csr* C);
/* General products and sums of the AMR path (b2_matmat.cu).  C = A B as a NEW matrix the caller owns: matrix_RightMatMult /
 * matrix_LeftMatMult (MatMatMult, PetscMatrix.cpp:766-790; _PP[ig] <- _PP[ig] * _PPamr[ig-1], LinearImplicitSystem.cpp:255-258)
 * and, applied twice, matrix_ABC (MatMatMatMult, :755-764).  Structural zeros are kept; run-to-run bit-identical. */
int b2_csr_matmat(const b2_csr* A, const b2_csr* B, b2_csr** C);
/* Y += a X for a pattern of X inside the pattern of Y (matrix_add / add: MatAXPY, PetscMatrix.cpp:793-812); fails otherwise */
int b2_csr_axpy(b2_csr* Y, double a, const b2_csr* X);
int b2_csr_pattern_contains(const b2_csr* Y, const b2_csr* X, int* contained);
double b2_csr_last_kernel_ms(const b2_csr* A);
/* Fast path of matrix_PtAP for the geometric prolongators of BuildProlongatorMatrix
 * (LinearImplicitSystem.cpp:826-909, Dirichlet rows/columns zeroed by :1032-1120): the product is
 * formed coarse element by coarse element, C = sum_E P_E^T (W_E o A|_E) P_E, from
 *   fine_dofs[nelc][nf]   rows of Af of the fine dofs of coarse element E (lattice order),
 *   coarse_dofs[nelc][nc] rows of Ac of its coarse dofs (GetSystemDof order),
 *   ploc[nf][nc]          element prolongator (ElemType.cpp:439-532), fine_entity[nf] the 3-trit
 *                         code (low/interior/high per direction) of the sub-entity each fine point lies on,
 *   valence[nelc][27]     number of coarse elements sharing each of the 27 sub-entities of E,
 *   fine_mask / coarse_mask (may be NULL): 1 where the row / column of P is zeroed (Bdc < 1.5).
 * Equal to b2_csr_ptap(P, Af, Ac) up to summation order.  nf/nc: 125/27 or 27/8 (hexahedra). */
int b2_galerkin_create(b2_csr* Af, b2_csr* Ac, int64_t nelc, int nf, int nc, const int32_t* fine_dofs,
                       const int32_t* coarse_dofs, const double* ploc, const uint8_t* fine_entity,
                       const uint8_t* valence, const uint8_t* fine_mask, const uint8_t* coarse_mask,
                       b2_galerkin** out);
int b2_galerkin_apply(b2_galerkin* g);       /* Ac = P^T Af P (Ac is overwritten) */
int b2_galerkin_destroy(b2_galerkin* g);

/* ---- mesh + assembly: replaces the element loop of applications/001_Poisson/main.cpp:346-605
 *      (+ elem_type_3D::Jacobian, ElemType.hpp:1438-1537; MatSetValuesBlocked, VecSetValues) ---- */
/* xyz[3][nnode] (SoA), conn[nel][27] node ids of the HEX27 elements owned by this rank. */
int b2_mesh_create(b2_ctx* c, int64_t nnode, int64_t nel, const double* xyz, const int32_t* conn, b2_mesh** out);
/* re-upload coordinates / connectivity into the existing device buffers, asynchronously on the
 * library stream (either may be NULL) */
int b2_mesh_update(b2_mesh* m, const double* xyz, const int32_t* conn);
/* upload the NEXT step's mesh into shadow buffers on the copy stream; b2_mesh_swap makes them current */
int b2_mesh_prefetch(b2_mesh* m, const double* xyz, const int32_t* conn);
int b2_mesh_swap(b2_mesh* m);
int b2_mesh_destroy(b2_mesh* m);
/* Assembly plan for one unknown on one mesh: nve dofs per element (<= 27), local nodes 0..nve-1 of every
 * conn row; dof[nel][nve] = matrix row of each local node (GetSystemDof, LinearEquation.cpp:76-85);
 * tables phi/dxi/deta/dzeta [ngauss][nve] and weights[ngauss] as elem_type_3D holds them
 * (ElemType.cpp:637-740), ngauss <= 64.  The tables ARE the element type (the reference dispatches on
 * _finiteElement[ielGeom][solType], main.cpp:438): hexahedra with 8 or 27 dofs and the 64-point rule run
 * the specialised kernels, every other family of elem_type_3D (tetrahedra 4/10/15, wedges 6/15/21, other
 * rules; conn rows padded to 27) the table-driven one.  Builds the element->CSR slot map for A. */
int b2_asm_create(b2_mesh* m, b2_csr* A, int nve, const int32_t* dof, int ngauss, const double* phi,
                  const double* dxi, const double* deta, const double* dzeta, const double* weights,
                  b2_asm** out);
int b2_asm_destroy(b2_asm* p);
/* A += sum_e B_e, rhs += sum_e F_e with B_ij = nu * int grad phi_i . grad phi_j,
 * F_i = int (fsrc phi_i - nu grad phi_i . grad u).  A and rhs are NOT zeroed here (the app calls
 * myKK->zero() / SetResZero() first, main.cpp:346, LinearImplicitSystem.cpp:322). */
int b2_asm_poisson(b2_asm* p, const b2_vec* u, b2_vec* rhs, double nu, double fsrc);
/* Neumann boundary integrals of the same callback (main.cpp:495-548, elem_type_2D::JacobianSur,
 * ElemType.hpp:1330-1379): rhs[dof] += sum_g phi_i(g) * value * weight_g over the listed boundary faces
 * (element, local face 0..5, constant flux value).  phi/dxi/deta [16][nvf] and weights[16] are the tables
 * of the face element elem_type_2D("quad", family, "seventh"), face_nodes[6][9] the local nodes of the
 * hexahedron's faces (Elem.hpp `ig` table). */
int b2_asm_neumann(b2_asm* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value,
                   int nvf, const double* phi, const double* dxi, const double* deta, const double* weights,
                   const int32_t* face_nodes, b2_vec* rhs);
/* The same for the faces of ANY element type, one face kind per call: the reference selects the face element
 * _finiteElement[GetElementFaceType(iel, jface)][order_ind] (main.cpp:507-525) -- here it is the tables handed
 * in: phi/dxi/deta [ngf][nvf], weights[ngf] with ngf <= 16 Gauss points and nvf <= 9 dofs (quadrilaterals
 * 4 / 8 / 9 dofs and 16 points; triangles 3 / 6 / 7 dofs and 13 points, Triangle.hpp:60-170); face_nodes[6][9]
 * = GetLocalFaceVertexIndex of the plan's element type, -1 where a face or entry does not exist.  Fails if a
 * listed face's dof is not an element dof. */
int b2_asm_neumann_faces(b2_asm* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value,
                         int nvf, int ngf, const double* phi, const double* dxi, const double* deta, const double* weights,
                         const int32_t* face_nodes, b2_vec* rhs);
/* Fused fast path of "assemble, then matrix_PtAP" (LinearImplicitSystem.cpp:326 + 347-370): the same
 * assembly, and in the same pass gal's coarse matrix Ac = P^T A P is formed from the element matrices
 * while they are on chip, C = sum_e Pc(e)^T B_e Pc(e) with Pc(e) the element prolongator of the child
 * (ElemType.cpp:439-532) with Dirichlet rows/columns dropped as in ZeroInterpolatorDirichletNodes --
 * equal to b2_galerkin_apply(gal) after b2_asm_poisson up to summation order, without re-reading
 * the fine matrix.  gal must have been created on this plan's matrix; Ac is overwritten, A is not zeroed.
 * Fails (no fallback) if the fine elements are not the children 8*E+j of gal's coarse elements. */
int b2_asm_poisson_galerkin(b2_asm* p, b2_galerkin* gal, const b2_vec* u, b2_vec* rhs, double nu, double fsrc);
/* name of the kernel the last b2_asm_poisson / b2_asm_poisson_galerkin call on this plan launched (measurement aid) */
const char* b2_asm_kernel_name(const b2_asm* p);
/* Element-matrix Galerkin chain for the levels below: with recording on, applying `gal` through
 * b2_asm_poisson_galerkin (or through b2_galerkin_apply_from_elements) also stores the nc x nc
 * Galerkin matrix D_E of each of its coarse elements; the next-coarser plan then forms
 * Ac = sum_E Pc(E)^T D_E Pc(E) from those (5.8 KB streamed per element) instead of gathering rows of
 * its fine matrix.  Equal to b2_galerkin_apply / matrix_PtAP up to summation order. */
int b2_galerkin_record_elements(b2_galerkin* gal, int on);
int b2_galerkin_apply_from_elements(b2_galerkin* gal, b2_galerkin* finer);
double b2_asm_last_kernel_ms(const b2_asm* p);

/* ---- geometric multigrid: replaces LinearEquationSolverPetsc::{MGInit,MGSetLevel,MGSolve,MGClear}
 *      and PETSc PCMG (LinearEquationSolverPetsc.cpp:185-353) ------------------------------- */
int b2_mg_create(b2_ctx* c, int nlevels, b2_mg** out);                   /* MGInit */
/* MGSetLevel: level operator A (penalised in place: rows bdc_idx -> identity), prolongator P from
 * level-1 (NULL on level 0), Dirichlet row list, smoothing steps and Richardson scale. */
int b2_mg_set_level(b2_mg* mg, int level, b2_csr* A, b2_csr* P, const int32_t* bdc_idx, int64_t nbdc,
                    int npre, int npost, double omega);
/* distributed run: layout of the level's vectors; A is then this rank's partial operator (sum over
 * its own elements), P its local prolongator.  Call before b2_mg_set_level. */
int b2_mg_set_level_halo(b2_mg* mg, int level, b2_halo* halo);
/* Null space of a level operator (RemoveNullSpace / GetNullSpaceBase, LinearEquationSolverPetsc.cpp:357-414: levels above
 * the coarsest; MatSetNullSpace + MatSetTransposeNullSpace): nvec (copied, normalised) spans it -- 1 on the free dofs of the
 * variable flagged by MultiLevelSolution::FixSolutionAtOnePoint, e.g. the pressure of an enclosed flow.  The level's
 * Richardson smoother then projects its right-hand side and every preconditioned residual, as KSPSolve does.  NULL unsets. */
int b2_mg_set_level_nullspace(b2_mg* mg, int level, const b2_vec* nvec);
/* per-phase device timing of the cycles (measurement aid): ms[nlevels][6] = pre-smoothing, residual, restriction, coarse
 * solve, prolongation, post-smoothing, summed over the cycles since the last call */
int b2_mg_set_timing(b2_mg* mg, int on);
int b2_mg_get_timing(b2_mg* mg, double* ms);
/* smoother of a level (call before b2_mg_set_level): kind 0 = Richardson(omega)+Jacobi, 1 = Chebyshev+Jacobi
 * (KSPCHEBYSHEV + PCJACOBI, LinearEquationSolverPetsc.cpp:452-536) on [emin, emax] of D^-1 A.  emax <= 0:
 * the bounds are OUR OWN stated ones -- [0.1, 1.1] x the largest eigenvalue found by 10 power iterations
 * from a fixed start vector at b2_mg_set_level (PETSc estimates with GMRES on a random vector, which is
 * not reproducible); b2_mg_level_bounds returns the interval in use. */
int b2_mg_set_smoother(b2_mg* mg, int level, int kind, double emin, double emax);
int b2_mg_level_bounds(const b2_mg* mg, int level, double* emin, double* emax);

/* ---- steady Stokes assembly (first kernel): the callback AssembleMatrixResNS of
 * applications/003_NavierStokes/SteadyStokes/main.cpp:290-598 for three velocity components of one Lagrange family
 * (nve_v <= 27 nodes, also the geometry) and a pressure of another (nve_p <= 8) -- Taylor-Hood pairs; the equal-order
 * stabilisation (alpha != 0, :336-340) is not implemented.  A: the system matrix on the pattern of
 * LinearEquation::GetSparsityPatternSize for U, V, W, P (b2h_system_sparsity_create); elem_dofs [nel][4][27] =
 * GetSystemDof per variable (b2h_system_elem_dofs, host array); velocity tables dxi/deta/dzeta [ngauss][nve_v],
 * weights[ngauss], pressure table phi_p [ngauss][nve_p] of the element type (b2h_elem_tables).
 * b2_stokes_create also builds the element -> CSR slot map of the plan and FAILS if an element coupling is not an
 * entry of A's pattern (or a row of A is longer than 65536 entries).
 * b2_stokes_assemble: A += element blocks (IRe K on the velocity diagonal, -int dphi_i/dx_k phi1_j and its transpose),
 * rhs += F = -B sol at the current solution in system numbering; neither is zeroed (the callback zeroes KK, :372). */
int b2_stokes_create(b2_mesh* mesh, b2_csr* A, const int32_t* elem_dofs, int nve_v, int nve_p, int ngauss, const double* dxi,
                     const double* deta, const double* dzeta, const double* weights, const double* phi_p, b2_stokes** out);
int b2_stokes_assemble(b2_stokes* p, const b2_vec* sol, b2_vec* rhs, double IRe);
int b2_stokes_destroy(b2_stokes* p);
/* steady Navier-Stokes on the same plan type: the library routine src/08_equations/assemble/03_navier_stokes.hpp
 * (:305-413): RES = -aRes with aResV[k][i] = int nu grad phi_i . grad u_k + phi_i u . grad u_k - p dphi_i/dx_k,
 * aResP[i] = -int div u psi_i, and KK += the exact Newton Jacobian d aRes / d sol (the reference records the loop with
 * adept; here it is written out analytically).  b2_ns_create = b2_stokes_create + the velocity phi table. */
int b2_ns_create(b2_mesh* mesh, b2_csr* A, const int32_t* elem_dofs, int nve_v, int nve_p, int ngauss, const double* phi_v, const double* dxi,
                 const double* deta, const double* dzeta, const double* weights, const double* phi_p, b2_stokes** out);
int b2_ns_assemble(b2_stokes* p, const b2_vec* sol, b2_vec* rhs, double nu);
/* boundary pressure term of the same routine (:196-300): on the listed boundary faces (one face kind per call, tables and
 * face_nodes as in b2_asm_neumann_faces, velocity family) RES[U_k dof of face node i] -= int phi_i tau n_k, tau = the
 * prescribed boundary pressure of the face (face_value; the reference evaluates its boundary callback per Gauss point),
 * n = the unit normal of elem_type_2D::JacobianSur at the Gauss point.  Which faces qualify (normal velocity
 * component not Dirichlet, :262-266) is the caller's selection. */
int b2_ns_pressure_faces(b2_stokes* p, int64_t nfaces, const int32_t* face_elem, const int32_t* face_local, const double* face_value, int nvf,
                         int ngf, const double* phi, const double* dxi, const double* deta, const double* weights, const int32_t* face_nodes,
                         b2_vec* rhs);

/* ---- element-block (ASM / Vanka) smoother: LinearEquationSolverPetscAsm (petsc_asm/LinearEquationSolverPetscAsm.cpp)
 * What the reference sets (:266-340, PetscPreconditioner.cpp:179-184): PCASM, PC_ASM_BASIC, local type
 * PC_COMPOSITE_MULTIPLICATIVE, the overlapping index sets of BuildASMIndex (:91-262), overlap 0.  On one rank that is
 *     y = 0;  for i = 0 .. nblocks-1:   y[B_i] += A[B_i,B_i]^-1 (r - A y)[B_i].
 * Block solves are exact (MLU_PRECOND on the blocks; stated choice -- 001_Poisson's own sub-preconditioner is one
 * SSOR sweep).  blk_ptr[nblocks+1] / blk_dofs: every block's dofs, sorted (host arrays; the host layer builds them,
 * b2h_asm_create).  The SCHEDULE group_ptr[ngroups+1] / group_blocks[nblocks] lists the blocks in sweep order, cut
 * into groups of mutually independent blocks (one launch per group, one CTA per block): the dependency levels of the
 * reference's block order reproduce its sweep exactly, a colouring gives few, large groups (b2h_asm_schedule).
 * The caller guarantees the independence inside a group; everything else is checked.
 * b2_schwarz_setup: numeric phase (extract and invert the blocks of A's current values).
 * b2_schwarz_apply: y = M^-1 r.  One rank only (no interface sums). */
int b2_schwarz_create(b2_ctx* ctx, b2_csr* A, int64_t nblocks, const int64_t* blk_ptr, const int32_t* blk_dofs,
                      int64_t ngroups, const int64_t* group_ptr, const int32_t* group_blocks, b2_schwarz** out);
/* block solve (call before b2_schwarz_setup): 2 = ILU(0) of every block in its sorted dofs -- ILU_PRECOND on the
 * blocks (PCILU: levels 0, natural ordering), the fine-grid preconditioner most of the reference's applications set;
 * factors on the pattern of the blocks' rows of A, any block size; with ONE block holding every element the level
 * smoother is Richardson + ILU(0), i.e. FEMuS_DEFAULT with ILU_PRECOND;
 * 0 = exact, dense inverses of blocks of at most 4096 dofs (default);
 * 1 = one SSOR iteration on the block's rows -- PCSOR's default (local symmetric sweep, omega 1, zero guess), the
 * sub-preconditioner 001_Poisson itself selects with SetPreconditionerFineGrids(SOR_PRECOND) (main.cpp:242,
 * LinearEquationSolverPetscAsm.cpp:300-317); no storage, any block size: with ONE block holding every element the
 * level smoother is Richardson + SOR, the application's FEMuS_DEFAULT branch. */
int b2_schwarz_set_subsolver(b2_schwarz* s, int kind);
/* SSOR / ILU(0) block solves walk a block's rows in order, one warp per block.  on != 0 (before b2_schwarz_setup): the
 * rows are sorted into dependency levels of the block's triangular patterns and every warp of the CTA takes rows of a
 * level, the CTA synchronising between levels -- the same arithmetic per row, hence the same result bit for bit; for
 * large blocks (8^4 elements per block in the reference's applications, one block per level for FEMuS_DEFAULT).
 * b2_schwarz_row_levels: the longest dependency chain found by the setup. */
int b2_schwarz_set_row_levels(b2_schwarz* s, int on);
int64_t b2_schwarz_row_levels(const b2_schwarz* s);
int b2_schwarz_setup(b2_schwarz* s);
int b2_schwarz_apply(b2_schwarz* s, const b2_vec* r, b2_vec* y);
int64_t b2_schwarz_bytes(const b2_schwarz* s);      /* HBM held by the block inverses */
int64_t b2_schwarz_groups(const b2_schwarz* s);
int b2_schwarz_destroy(b2_schwarz* s);
/* level smoother = KSPRICHARDSON(scale omega of b2_mg_set_level) + this preconditioner, npre / npost iterations
 * (LinearEquationSolverPetsc.cpp:516-519 with the PC of LinearEquationSolverPetscAsm::SetPreconditioner).  s must
 * have been created on the operator handed to b2_mg_set_level for that level; b2_mg_set_level runs its numeric
 * phase after the penalty.  NULL restores the level's previous smoother kind 0. */
int b2_mg_set_level_schwarz(b2_mg* mg, int level, b2_schwarz* s);
/* level solver around the level's preconditioner (Jacobi of kind 0, or the element-block sweep): kind 0 = KSPRICHARDSON
 * with the scale of b2_mg_set_level, 1 = KSPGMRES -- the reference's DEFAULT level solver (levels >= 1: GMRES +
 * ILU_PRECOND, LinearEquationSolverPetsc.hpp:128-151; 76 applications call SetSolverFineGrids(GMRES)): left
 * preconditioning, npre / npost iterations per smoothing call, the preconditioned residual is minimised.  GMRES
 * around the one-block ILU(0) sweep is the library's default smoother. */
int b2_mg_set_level_ksp(b2_mg* mg, int level, int kind);
/* coarsest level: a DIRECT solve, what the reference runs there (PREONLY + LU, PetscPreconditioner.cpp:147-160), in
 * place of the Jacobi-PCG of b2_mg_set_coarse -- s is a b2_schwarz on the level-0 operator with ONE block holding
 * every dof and the exact block solve (at most 4096 dofs); required for indefinite (velocity-pressure) systems.  Call
 * before b2_mg_set_level(0, ...), which runs the numeric phase after the penalty.  NULL: back to the PCG. */
int b2_mg_set_coarse_schwarz(b2_mg* mg, b2_schwarz* s);
/* coarse solver: Jacobi-PCG to ||r|| <= rtol ||b|| (the reference: PREONLY + MUMPS LU,
 * PetscPreconditioner.cpp:147-160) */
int b2_mg_set_coarse(b2_mg* mg, double rtol, int maxit);
int b2_mg_vcycle(b2_mg* mg, const b2_vec* rhs, b2_vec* x);               /* one PCMG multiplicative V */
/* MGSolve with outer PREONLY: RES[bdc]=0; EPSC = Vcycle(RES); RESC = A EPSC; RES -= RESC; EPS += EPSC */
int b2_mg_solve(b2_mg* mg, b2_vec* res, b2_vec* eps);
int b2_mg_coarse_iterations(const b2_mg* mg);
int b2_mg_destroy(b2_mg* mg);                                            /* MGClear */

#ifdef __cplusplus
}
#endif
#endif
