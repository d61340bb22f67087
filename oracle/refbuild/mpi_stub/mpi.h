/*
 * Single-rank MPI stand-in, written for this repo (no MPI exists in the image).
 * It lets the UNMODIFIED host sources of the reference (src/main.cpp, comm.cpp,
 * init.cpp) compile and run as one rank so that the reference's own GPU binary can
 * produce golden outputs for the oracle (oracle/refbuild/build_ref.sh).
 * Only the ~25 entry points the reference calls are provided; every collective is the
 * identity on one rank.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef CUDNS_MPI_STUB_H_
#define CUDNS_MPI_STUB_H_
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int count; } MPI_Status;
typedef struct { long long nelem; int elsize; } MPI_Datatype;
typedef FILE *MPI_File;
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4
#define MPI_ORDER_C 56
#define MPI_SUM 1
#define MPI_MIN 2
static const MPI_Datatype MPI_DOUBLE = {1, 8};
static const MPI_Datatype MPI_FLOAT  = {1, 4};

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline double MPI_Wtime() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9*t.tv_nsec; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Cart_create(MPI_Comm, int, const int *, const int *, int, MPI_Comm *c) { *c = 0; return 0; }
static inline int MPI_Cart_coords(MPI_Comm, int, int n, int *c) { for (int i = 0; i < n; i++) c[i] = 0; return 0; }
static inline int MPI_Cart_rank(MPI_Comm, const int *, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *c) { *c = 0; return 0; }
static inline int MPI_Get_processor_name(char *n, int *l) { strcpy(n, "localhost"); *l = 9; return 0; }
static inline int MPI_Error_string(int, char *s, int *l) { strcpy(s, "stub"); *l = 4; return 0; }
static inline int MPI_Sendrecv(const void *sb, int sc, MPI_Datatype st, int, int,
                               void *rb, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) {
    memcpy(rb, sb, (size_t)sc*st.nelem*st.elsize); return 0; }
static inline int MPI_Allreduce(const void *sb, void *rb, int c, MPI_Datatype t, MPI_Op, MPI_Comm) {
    memcpy(rb, sb, (size_t)c*t.nelem*t.elsize); return 0; }
static inline int MPI_Reduce(const void *sb, void *rb, int c, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    memcpy(rb, sb, (size_t)c*t.nelem*t.elsize); return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Type_create_subarray(int nd, const int *, const int *sub, const int *, int, MPI_Datatype old, MPI_Datatype *nt) {
    long long n = 1; for (int i = 0; i < nd; i++) n *= sub[i]; nt->nelem = n*old.nelem; nt->elsize = old.elsize; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_File_open(MPI_Comm, const char *name, int, MPI_Info, MPI_File *fh) { *fh = fopen(name, "wb"); return *fh ? 0 : 1; }
static inline int MPI_File_set_view(MPI_File, MPI_Offset, MPI_Datatype, MPI_Datatype, const char *, MPI_Info) { return 0; }
static inline int MPI_File_write_all(MPI_File fh, const void *buf, int c, MPI_Datatype t, MPI_Status *) {
    if (!fh) return 1; return fwrite(buf, (size_t)t.elsize, (size_t)c*t.nelem, fh) == (size_t)c*t.nelem ? 0 : 1; }
static inline int MPI_File_close(MPI_File *fh) { if (*fh) { fclose(*fh); *fh = NULL; } return 0; }
#endif
