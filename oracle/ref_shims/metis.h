/* TEST INFRASTRUCTURE (oracle/): stand-in for <metis.h>.  The oracle build runs on ONE rank, where the reference never
 * needs a real partition; METIS_PartMeshDual here puts every element and node in part 0. */
#ifndef FEMUS_B200_ORACLE_METIS_SHIM_H
#define FEMUS_B200_ORACLE_METIS_SHIM_H
typedef int idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
enum { METIS_OK = 1, METIS_ERROR_INPUT = -2, METIS_ERROR_MEMORY = -3, METIS_ERROR = -4 };
enum { METIS_OPTION_PTYPE, METIS_OPTION_OBJTYPE, METIS_OPTION_CTYPE, METIS_OPTION_IPTYPE, METIS_OPTION_RTYPE, METIS_OPTION_DBGLVL,
       METIS_OPTION_NITER, METIS_OPTION_NCUTS, METIS_OPTION_SEED, METIS_OPTION_NO2HOP, METIS_OPTION_MINCONN, METIS_OPTION_CONTIG,
       METIS_OPTION_COMPRESS, METIS_OPTION_CCORDER, METIS_OPTION_PFACTOR, METIS_OPTION_NSEPS, METIS_OPTION_UFACTOR, METIS_OPTION_NUMBERING };
enum { METIS_PTYPE_RB, METIS_PTYPE_KWAY };
enum { METIS_CTYPE_RM, METIS_CTYPE_SHEM };
enum { METIS_IPTYPE_GROW, METIS_IPTYPE_RANDOM, METIS_IPTYPE_EDGE, METIS_IPTYPE_NODE };
static inline int METIS_SetDefaultOptions(idx_t* o) { for (int i = 0; i < METIS_NOPTIONS; i++) o[i] = -1; return METIS_OK; }
static inline int METIS_PartMeshDual(idx_t* ne, idx_t* nn, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, real_t*, idx_t*, idx_t* objval,
                                     idx_t* epart, idx_t* npart) {
  for (idx_t i = 0; i < *ne; i++) epart[i] = 0;
  for (idx_t i = 0; i < *nn; i++) npart[i] = 0;
  *objval = 0;
  return METIS_OK;
}
#endif
