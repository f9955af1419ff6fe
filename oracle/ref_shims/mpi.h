/* TEST INFRASTRUCTURE (oracle/): single-process stand-in for <mpi.h>, just enough for the reference's mesh /
 * solution / system layer to compile and run on ONE rank without an MPI installation.
 * Collectives over one rank are copies; point-to-point calls abort (never reached on one rank). */
#ifndef FEMUS_B200_ORACLE_MPI_SHIM_H
#define FEMUS_B200_ORACLE_MPI_SHIM_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL (-1)
#define MPI_IN_PLACE ((void*)-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_WTIME_IS_GLOBAL 0
/* datatype handle = size in bytes (all the shim needs) */
#define MPI_CHAR 1
#define MPI_UNSIGNED_CHAR 1
#define MPI_BYTE 1
#define MPI_PACKED 1
#define MPI_SHORT 2
#define MPI_UNSIGNED_SHORT 2
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_LONG_DOUBLE 16
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_BAND 4
#define MPI_BOR 5
#define MPI_LAND 6
#define MPI_LOR 7
#define MPI_PROD 8

#ifdef __cplusplus
extern "C" {
#endif
static inline void femus_b200_mpi_p2p_(const char* what) {
  fprintf(stderr, "oracle mpi shim: %s called on a single rank\n", what);
  abort();
}
static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Initialized(int* flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code); return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline double MPI_Wtime(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c;
  if (s != MPI_IN_PLACE) memcpy(r, s, (size_t)n * (size_t)t);
  return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  (void)root;
  return MPI_Allreduce(s, r, n, t, op, c);
}
static inline int MPI_Allgather(const void* s, int ns, MPI_Datatype ts, void* r, int nr, MPI_Datatype tr, MPI_Comm c) {
  (void)nr; (void)tr; (void)c;
  if (s != MPI_IN_PLACE) memcpy(r, s, (size_t)ns * (size_t)ts);
  return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void* s, int ns, MPI_Datatype ts, void* r, int nr, MPI_Datatype tr, MPI_Comm c) {
  return MPI_Allgather(s, ns, ts, r, nr, tr, c);
}
static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; femus_b200_mpi_p2p_("MPI_Send"); return 1;
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status* st) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)st; femus_b200_mpi_p2p_("MPI_Recv"); return 1;
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request* rq) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c; (void)rq; femus_b200_mpi_p2p_("MPI_Isend"); return 1;
}
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request* rq) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)rq; femus_b200_mpi_p2p_("MPI_Irecv"); return 1;
}
static inline int MPI_Sendrecv(const void* sb, int sn, MPI_Datatype st, int dst, int stag, void* rb, int rn, MPI_Datatype rt, int src, int rtag,
                               MPI_Comm c, MPI_Status* s) {
  (void)sb; (void)sn; (void)st; (void)dst; (void)stag; (void)rb; (void)rn; (void)rt; (void)src; (void)rtag; (void)c; (void)s;
  femus_b200_mpi_p2p_("MPI_Sendrecv"); return 1;
}
static inline int MPI_Probe(int src, int tag, MPI_Comm c, MPI_Status* s) { (void)src; (void)tag; (void)c; (void)s; femus_b200_mpi_p2p_("MPI_Probe"); return 1; }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r; (void)s; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request* r, int* flag, MPI_Status* s) { (void)r; (void)s; *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Request_free(MPI_Request* r) { *r = MPI_REQUEST_NULL; return MPI_SUCCESS; }
static inline int MPI_Get_count(const MPI_Status* s, MPI_Datatype t, int* n) { (void)s; (void)t; *n = 0; return MPI_SUCCESS; }
static inline int MPI_Pack_size(int n, MPI_Datatype t, MPI_Comm c, int* size) { (void)c; *size = n * t; return MPI_SUCCESS; }
static inline int MPI_Pack(const void* in, int n, MPI_Datatype t, void* out, int outsize, int* pos, MPI_Comm c) {
  (void)outsize; (void)c;
  memcpy((char*)out + *pos, in, (size_t)n * (size_t)t);
  *pos += n * t;
  return MPI_SUCCESS;
}
static inline int MPI_Unpack(const void* in, int insize, int* pos, void* out, int n, MPI_Datatype t, MPI_Comm c) {
  (void)insize; (void)c;
  memcpy(out, (const char*)in + *pos, (size_t)n * (size_t)t);
  *pos += n * t;
  return MPI_SUCCESS;
}
static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* out) { *out = n * t; return MPI_SUCCESS; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype* t) { (void)t; return MPI_SUCCESS; }
#ifdef __cplusplus
}
#endif
#endif
