/* femus_b200_host -- C view of the backend's HOST layer (femus_b200/host): the pieces of FEMuS
 * layers L2/L6/L8 that produce the integer inputs of the device kernels.  Used by the harness
 * (tests/bench) to drive and inspect the C++ classes; a FEMuS application would use its own
 * Mesh / MultiLevelSolution / LinearImplicitSystem objects and only the device ABI
 * (femus_b200.h).  Each entry names the reference interface it mirrors (paths relative to the
 * reference's src/).  Pure CPU: no GPU needed. */
#ifndef FEMUS_B200_HOST_H
#define FEMUS_B200_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2h_hier b2h_hier;   /* MultiLevelMesh: levels 0..nlevels-1 */
typedef struct b2h_csr b2h_csr;     /* host CSR matrix */

/* MultiLevelMesh::GenerateCoarseBoxMesh(nx,ny,nz,...,HEX27,"seventh") + RefineMesh(nlevels,...)
 * (06_mesh/01_multiple_levels/00_definition/MultiLevelMesh.cpp; MeshGeneration.cpp:790-1071;
 * MeshRefinement.cpp:188-507).  bounds6 = xmin,xmax,ymin,ymax,zmin,zmax (NULL: unit cube).
 * nprocs > 1: z-slab partition of level 0, children inherit (MeshMetisPartitioning.cpp:143-155). */
b2h_hier* b2h_hier_create(int nx, int ny, int nz, int nlevels, const double* bounds6, int nprocs);
/* The same hierarchy restricted to ONE rank of the z-slab partition: the rank's level-0 elements
 * (Mesh::_elementOffset[rank] .. [rank+1]) with locally renumbered nodes, refined locally (children
 * inherit the parent's rank, MeshMetisPartitioning.cpp:143-155).  Inside, nprocs == 1. */
b2h_hier* b2h_hier_create_local(int nx, int ny, int nz, int nlevels, const double* bounds6, int nprocs, int rank);
/* MultiLevelMesh::ReadCoarseMesh(name, "seventh", Lref) for a Gambit .neu file of 27-node hexahedra, 10-node
 * tetrahedra and 18-node wedges, also mixed (GambitIO.cpp:92-352; the face and centre nodes tetrahedra and
 * wedges lack are added as Mesh::AddBiquadraticNodesNotInMeshFile does, Mesh.cpp:1207-1333) + RefineMesh;
 * aborts on anything else, like the reference on bad input */
b2h_hier* b2h_hier_create_from_neu(const char* path, int nlevels, double Lref);
void b2h_hier_destroy(b2h_hier* h);
/* integer lattice coordinates [3][nnode] of the nodes (rank-independent node names) */
const int32_t* b2h_level_ijk(const b2h_hier* h, int l);
/* nodes on faces shared with other ranks' sub-meshes, sorted (out may be NULL: count only) */
int64_t b2h_level_interface_nodes(const b2h_hier* h, int l, int32_t* out);
int b2h_hier_nlevels(const b2h_hier* h);
int b2h_hier_nprocs(const b2h_hier* h);
int64_t b2h_level_nel(const b2h_hier* h, int l);                 /* Mesh::GetNumberOfElements */
int64_t b2h_level_nnode(const b2h_hier* h, int l);               /* Mesh::GetNumberOfNodes */
const int32_t* b2h_level_conn(const b2h_hier* h, int l);         /* elem::GetElementDofIndex, [nel][27] */
const int32_t* b2h_level_face(const b2h_hier* h, int l);         /* elem::GetFaceElementIndex, [nel][6] */
const int32_t* b2h_level_part(const b2h_hier* h, int l);         /* element -> rank */
const double* b2h_level_xyz(const b2h_hier* h, int l);           /* Mesh::GetTopology()->_Sol[0..2], [3][nnode] */
const int32_t* b2h_level_child_el(const b2h_hier* h, int l);     /* elem::GetChildElement, [nel][8] or NULL */
/* Mesh::_elementOffset [nprocs+1] and _dofOffset[k=0..2][nprocs+1] (Mesh.cpp:706-853) */
void b2h_level_offsets(const b2h_hier* h, int l, int64_t* elem_offset, int64_t* dof_offset3);
int64_t b2h_level_ndofs(const b2h_hier* h, int l, int family);
/* LinearEquation::GetSystemDof for a single-variable system (LinearEquation.cpp:76-85), [nel][nve] */
void b2h_level_system_dofs(const b2h_hier* h, int l, int family, int32_t* out);
/* MultiLevelSolution::GenerateBdc (MultiLevelSolution.cpp:725-840): dirichlet_faces7[1..6] */
void b2h_level_bdc(const b2h_hier* h, int l, int family, const int* dirichlet_faces7, double* out);

/* LinearImplicitSystem::BuildProlongatorMatrix (LinearImplicitSystem.cpp:826-909), before the
 * Dirichlet rows/columns are zeroed */
b2h_csr* b2h_prolongator_create(const b2h_hier* h, int lfine, int family);
void b2h_csr_destroy(b2h_csr* p);
int64_t b2h_csr_nrows(const b2h_csr* p);
int64_t b2h_csr_ncols(const b2h_csr* p);
int64_t b2h_csr_nnz(const b2h_csr* p);
const int64_t* b2h_csr_rowptr(const b2h_csr* p);
const int32_t* b2h_csr_col(const b2h_csr* p);
const double* b2h_csr_val(const b2h_csr* p);

/* Maps of the element-gather Galerkin product (b2_galerkin_create in femus_b200.h): number of
 * fine dofs of a coarse element (125 biquadratic / 27 linear), the dense element prolongator
 * ploc[nf][nc] (ElemType.cpp:439-532) with the entity code of every fine lattice point, and for
 * the coarse elements [e0,e1) of level lcoarse their fine dofs [nf] (through elem::GetChildElement
 * and Mesh::GetSolutionDof, as LinearImplicitSystem.cpp:761-811 walks them) and the number of
 * elements of that range at each of their 27 nodes. */
int b2h_galerkin_nf(int family);
void b2h_galerkin_element(int family, double* ploc, uint8_t* fine_entity);
int b2h_galerkin_maps(const b2h_hier* h, int lcoarse, int family, int64_t e0, int64_t e1, int32_t* fine_dofs,
                      uint8_t* valence);

/* Element types (GeomElTypeEnum: 0 HEX, 1 TET, 2 WEDGE).  b2h_level_elem_type: the type of a level, -1 if the
 * mesh mixes types (then b2h_level_elem_types gives the type of every element).  Per type: dofs per element
 * of a family (Elem.hpp NVE: hex 8/20/27, tet 4/10/15, wedge 6/15/21), Gauss points of the "seventh" rule
 * (64 / 31 / 52), elem_type_3D(type, family, "seventh") tables [ngauss][nve], weights[ngauss]
 * (ElemType.cpp:637-740), element prolongator row of (child, child-local node) (ElemType.cpp:439-532 with the
 * bases' fine2CoarseVertexMapping) and the parent face a child face lies on, -1 if interior
 * (coarse2FineFaceMapping, MeshRefinement.hpp:79-100).  conn rows hold 27 / 15 / 21 nodes (padded with -1
 * to 27), face rows 6 / 4 / 5 flags (padded with -1 to 6), child_el the 8 children of every element. */
int b2h_level_elem_type(const b2h_hier* h, int l);
void b2h_level_elem_types(const b2h_hier* h, int l, uint8_t* out);
int b2h_elem_nve(int type, int family);
int b2h_elem_ngauss(int type);
void b2h_elem_tables(int type, int family, double* phi, double* dxi, double* deta, double* dzeta, double* w);
int b2h_elem_prolongator_row(int type, int family, int child, int node, int* idx, double* val);
int b2h_elem_child_face(int type, int child, int child_face);
/* GetSystemDof of every element in rows of 27 padded with -1 (meshes of several element types), and the
 * sparsity pattern of the single-variable system built on the host (LinearEquation::GetSparsityPatternSize,
 * LinearEquation.cpp:407-548): a b2h_csr with zero values */
void b2h_level_system_dofs27(const b2h_hier* h, int l, int family, int32_t* out);
b2h_csr* b2h_sparsity_create(const b2h_hier* h, int l, int family);
/* test hook: a generated hexahedral box hierarchy refined by the general (any element type) code path */
b2h_hier* b2h_hier_create_general(int nx, int ny, int nz, int nlevels);

/* elem_type_3D("hex", family, "seventh"): tables [64][nve] and weights[64] (ElemType.cpp:637-740),
 * element prolongator row of the fine point (a,b,c) of the 5x5x5 lattice (ElemType.cpp:439-532) */
int b2h_hex_nve(int family);
void b2h_hex_tables(int family, double* phi, double* dxi, double* deta, double* dzeta, double* w);
int b2h_hex_prolongator_row(int family, int a, int b, int c, int* idx, double* val);
/* face element elem_type_2D("quad", family, "seventh") of a hexahedron: tables [16][nvf] and weights[16]
 * (nvf = 4 or 9), the local nodes of the 6 faces [6][9] (Elem.hpp `ig` table) and the boundary faces
 * of a level as (element, local face, boundary index = -(faceElementIndex+1), Elem.cpp:361-364);
 * b2h_level_boundary_faces with NULL arrays only counts. */
int b2h_face_nvf(int family);
void b2h_face_tables(int family, double* phi, double* dxi, double* deta, double* w);
void b2h_hex_face_nodes(int32_t* out);
int64_t b2h_level_boundary_faces(const b2h_hier* h, int l, int32_t* elem, int32_t* face, int32_t* bidx);
/* face elements of every 3-D element: elem_type_2D(kind, family, "seventh"), kind 0 = "quad" (4 / 8 / 9 dofs,
 * 16 points), 1 = "tri" (3 / 6 / 7 dofs, 13 points; 01_fe/2d/Triangle.hpp:60-170, quadrature_Triangle.cpp) --
 * what _finiteElement[GetElementFaceType(iel, jface)][order_ind] selects in main.cpp:507-525.  Tables
 * [ngauss][ndofs], weights[ngauss].  Per element type: local nodes of its faces [6][9] padded with -1
 * (Elem.hpp `ig` table, GetLocalFaceVertexIndex) and the face kind of face f (-1 past the last face). */
int b2h_face_kind_ngauss(int kind);
int b2h_face_kind_ndofs(int kind, int family);
void b2h_face_kind_tables(int kind, int family, double* phi, double* dxi, double* deta, double* w);
void b2h_elem_face_nodes(int type, int32_t* out);
int b2h_elem_face_kind(int type, int f);


/* ---- element-block (ASM / Vanka) smoother, one Lagrange variable, no Schur variable (001_Poisson "asm", main.cpp:234-250)
 * b2h_asm_create: MeshASMPartitioning::DoPartition (MeshASMPartitioning.cpp:89-148) with block_elems elements per
 * block, capped by the level's element count (LinearImplicitSystem.cpp:1191-1201), then
 * LinearEquationSolverPetscAsm::BuildASMIndex (LinearEquationSolverPetscAsm.cpp:91-262) for rank iproc: per block
 * its elements, the sorted "local" and "overlapping" index sets (CSR-like: ptr[nblocks+1] + entries).  NULL on bad
 * arguments (b2h_last_error).  block_type_range[3]: _blockTypeRange (solid / porous / fluid block ends).
 * b2h_asm_schedule: groups of mutually independent blocks for the multiplicative sweep of b2_schwarz_apply on the
 * operator pattern (rowptr, col): mode 0 = dependency levels of the given block order (the reference's sequential
 * sweep, exactly), mode 1 = greedy colours (the reference's sweep with the block list stably sorted by colour).
 * Writes group_of_block[nblocks], returns the number of groups, -1 on bad arguments. */
typedef struct b2h_asm b2h_asm;
b2h_asm* b2h_asm_create(const b2h_hier* h, int l, int family, int block_elems, int iproc);
/* the same for a system of nvars Lagrange variables (families[k] = 0 / 1 / 2) numbered [rank][variable][dof]
 * (LinearEquation::InitPde, GetSystemDof, LinearEquation.cpp:76-85, 211-237) whose LAST nschur variables are Schur
 * (pressure-like) variables: Vanka blocks = non-Schur dofs of one layer of near elements (elem::BuildElementNearElement,
 * Elem.cpp:493-526) + Schur dofs of the block's own elements.  b2h_system_offsets: KKoffset [nvars+1][nprocs]. */
b2h_asm* b2h_asm_create_system(const b2h_hier* h, int l, int nvars, const int* families, int nschur, int block_elems, int iproc);
int b2h_system_offsets(const b2h_hier* h, int l, int nvars, const int* families, int64_t* out);
/* systems of several Lagrange variables (host side), rows [rank][variable][dof]:
 * b2h_system_elem_dofs: GetSystemDof of every element, [nel][nvars][27] padded with -1;
 * b2h_system_sparsity_create: LinearEquation::GetSparsityPatternSize (LinearEquation.cpp:407-548), pattern = nvars x
 *   nvars coupling flags (_SparsityPattern) or NULL for all pairs; a b2h_csr with zero values;
 * b2h_system_prolongator_create: LinearImplicitSystem::BuildProlongatorMatrix (LinearImplicitSystem.cpp:826-909),
 *   variable by variable, from level lfine-1 to lfine;
 * b2h_system_bdc: GenerateBdc of every variable in system numbering, dirichlet[nvars][7] flags of the boundary sets. */
void b2h_system_elem_dofs(const b2h_hier* h, int l, int nvars, const int* families, int32_t* out);
b2h_csr* b2h_system_sparsity_create(const b2h_hier* h, int l, int nvars, const int* families, const uint8_t* pattern);
b2h_csr* b2h_system_prolongator_create(const b2h_hier* h, int lfine, int nvars, const int* families);
int b2h_system_bdc(const b2h_hier* h, int l, int nvars, const int* families, const uint8_t* dirichlet, double* out);
void b2h_asm_destroy(b2h_asm* a);
int64_t b2h_asm_nblocks(const b2h_asm* a);
void b2h_asm_block_type_range(const b2h_asm* a, int64_t* out3);
const int64_t* b2h_asm_elem_ptr(const b2h_asm* a);
const int32_t* b2h_asm_elems(const b2h_asm* a);
const int64_t* b2h_asm_local_ptr(const b2h_asm* a);
const int32_t* b2h_asm_local(const b2h_asm* a);
const int64_t* b2h_asm_overlap_ptr(const b2h_asm* a);
const int32_t* b2h_asm_overlap(const b2h_asm* a);
int64_t b2h_asm_schedule(int64_t n, const int64_t* rowptr, const int32_t* col, int64_t nblocks, const int64_t* blk_ptr,
                         const int32_t* blk_dofs, int mode, int32_t* group_of_block);
const char* b2h_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
