// Single-rank stand-ins for MPI, BLACS and ScaLAPACK/PBLAS so that the UNMODIFIED reference sources
// link into a working `dissect` for the parity oracle (oracle/_ref).  TEST INFRASTRUCTURE ONLY.
//
// On a 1 x 1 process grid every block-cyclic matrix is one local column-major array, so each PBLAS
// routine is its BLAS/LAPACK namesake on that array (sub-matrix offsets ia/ja honoured through the
// descriptor's lld), BLACS point-to-point is a FIFO inside the process (the reference always sends
// before it receives, matrix.cpp:194-201), MPI collectives are memcpy, and an MPI-IO darray view is
// the contiguous column-major file.  BLAS comes from OpenBLAS (see oracle/Makefile); with
// -DSHIM_INTERNAL_BLAS a plain blocked dsyrk/dgemm/dsymm is used instead (same answers to rounding).
#include <mpi.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>
#include <unistd.h>
#include <fcntl.h>
#include <sys/stat.h>

static size_t type_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: return 1;
    case MPI_INT: return sizeof(int);
    case MPI_LONG: return sizeof(long);
    case MPI_DOUBLE: return sizeof(double);
    default: return 8;  // darray of doubles
  }
}

struct shim_mpi_file {
  int fd;
  long long view_disp;
  long long pos;  // within the view
};

extern "C" {

// ------------------------------------------------------------------------------------------ MPI
int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm, int code) { fflush(stdout); fflush(stderr); _exit(code == 0 ? 1 : (code & 0xff ? code & 0xff : 1)); }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
int MPI_Get_processor_name(char* name, int* len) { strcpy(name, "shim"); *len = 4; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
int MPI_Gather(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, int, MPI_Datatype, int, MPI_Comm) {
  if (rbuf != sbuf) memmove(rbuf, sbuf, (size_t)scount * type_size(stype));
  return MPI_SUCCESS;
}
int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, const int*, const int* displs,
                MPI_Datatype rtype, int, MPI_Comm) {
  char* dst = (char*)rbuf + (displs ? (size_t)displs[0] * type_size(rtype) : 0);
  if (dst != sbuf) memmove(dst, sbuf, (size_t)scount * type_size(stype));
  return MPI_SUCCESS;
}
int MPI_Alltoall(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, int, MPI_Datatype, MPI_Comm) {
  if (rbuf != sbuf) memmove(rbuf, sbuf, (size_t)scount * type_size(stype));
  return MPI_SUCCESS;
}
int MPI_Alltoallv(const void* sbuf, const int* scounts, const int* sdispls, MPI_Datatype stype, void* rbuf, const int*,
                  const int* rdispls, MPI_Datatype rtype, MPI_Comm) {
  memmove((char*)rbuf + (size_t)rdispls[0] * type_size(rtype), (const char*)sbuf + (size_t)sdispls[0] * type_size(stype),
          (size_t)scounts[0] * type_size(stype));
  return MPI_SUCCESS;
}
int MPI_Comm_group(MPI_Comm, MPI_Group* g) { *g = 0; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group, int, const int*, MPI_Group* g) { *g = 0; return MPI_SUCCESS; }
int MPI_Group_rank(MPI_Group, int* rank) { *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm* c) { *c = MPI_COMM_WORLD; return MPI_SUCCESS; }
int MPI_Type_create_darray(int, int, int, const int*, const int*, const int*, const int*, int, MPI_Datatype,
                           MPI_Datatype* newtype) { *newtype = 100; return MPI_SUCCESS; }
int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype*) { return MPI_SUCCESS; }

int MPI_File_open(MPI_Comm, const char* filename, int amode, MPI_Info, MPI_File* fh) {
  int flags = 0;
  if (amode & MPI_MODE_RDONLY) flags |= O_RDONLY;
  if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
  if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
  if (amode & MPI_MODE_EXCL) flags |= O_EXCL;
  const int fd = open(filename, flags, 0644);
  if (fd < 0) return 1;
  *fh = new shim_mpi_file{fd, 0, 0};
  return MPI_SUCCESS;
}
int MPI_File_close(MPI_File* fh) {
  if (!fh || !*fh) return 1;
  const int rc = close((*fh)->fd);
  delete *fh;
  *fh = nullptr;
  return rc == 0 ? MPI_SUCCESS : 1;
}
int MPI_File_delete(const char* filename, MPI_Info) { return unlink(filename) == 0 ? MPI_SUCCESS : 1; }
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype, MPI_Datatype, const char*, MPI_Info) {
  fh->view_disp = disp;
  fh->pos = 0;
  return MPI_SUCCESS;
}
static int file_rw(MPI_File fh, long long off, void* buf, size_t bytes, bool wr) {
  char* p = (char*)buf;
  while (bytes > 0) {
    const ssize_t k = wr ? pwrite(fh->fd, p, bytes, off) : pread(fh->fd, p, bytes, off);
    if (k <= 0) return 1;
    p += k; off += k; bytes -= (size_t)k;
  }
  return MPI_SUCCESS;
}
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status*) {
  const size_t bytes = (size_t)count * type_size(type);
  const int rc = file_rw(fh, fh->view_disp + fh->pos, (void*)buf, bytes, true);
  fh->pos += (long long)bytes;
  return rc;
}
int MPI_File_write_all(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* st) {
  return MPI_File_write(fh, buf, count, type, st);
}
int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void* buf, int count, MPI_Datatype type, MPI_Status*) {
  return file_rw(fh, fh->view_disp + offset, (void*)buf, (size_t)count * type_size(type), true);
}
int MPI_File_read_all(MPI_File fh, void* buf, int count, MPI_Datatype type, MPI_Status*) {
  const size_t bytes = (size_t)count * type_size(type);
  const int rc = file_rw(fh, fh->view_disp + fh->pos, buf, bytes, false);
  fh->pos += (long long)bytes;
  return rc;
}
int MPI_Error_class(int code, int* cls) { *cls = code; return MPI_SUCCESS; }
int MPI_Error_string(int, char* s, int* len) { strcpy(s, "shim MPI error"); *len = (int)strlen(s); return MPI_SUCCESS; }

// ---------------------------------------------------------------------------------------- BLACS
void Cblacs_pinfo(int* mypnum, int* nprocs) { *mypnum = 0; *nprocs = 1; }
void Cblacs_get(int, int, int* val) { *val = 0; }
void Cblacs_gridinit(int* ctx, const char*, int, int) { *ctx = 0; }
void Cblacs_gridmap(int* ctx, int*, int, int, int) { *ctx = 0; }
void Cblacs_pcoord(int, int, int* r, int* c) { *r = 0; *c = 0; }
void Cblacs_gridinfo(int, int* nr, int* nc, int* r, int* c) { *nr = 1; *nc = 1; *r = 0; *c = 0; }
void Cblacs_gridexit(int) {}
void Cblacs_barrier(int, const char*) {}

static std::deque<std::vector<double> > g_dq;
static std::deque<std::vector<int> > g_iq;

void Cdgesd2d(int, int m, int n, double* A, int lda, int, int) {
  std::vector<double> msg((size_t)m * n);
  for (int j = 0; j < n; ++j) memcpy(&msg[(size_t)j * m], A + (size_t)j * lda, (size_t)m * sizeof(double));
  g_dq.push_back(msg);
}
void Cdgerv2d(int, int m, int n, double* A, int lda, int, int) {
  if (g_dq.empty() || g_dq.front().size() != (size_t)m * n) { fprintf(stderr, "shim: BLACS receive without matching send\n"); abort(); }
  const std::vector<double>& msg = g_dq.front();
  for (int j = 0; j < n; ++j) memcpy(A + (size_t)j * lda, &msg[(size_t)j * m], (size_t)m * sizeof(double));
  g_dq.pop_front();
}
void Cigesd2d(int, int m, int n, int* A, int lda, int, int) {
  std::vector<int> msg((size_t)m * n);
  for (int j = 0; j < n; ++j) memcpy(&msg[(size_t)j * m], A + (size_t)j * lda, (size_t)m * sizeof(int));
  g_iq.push_back(msg);
}
void Cigerv2d(int, int m, int n, int* A, int lda, int, int) {
  if (g_iq.empty() || g_iq.front().size() != (size_t)m * n) { fprintf(stderr, "shim: BLACS receive without matching send\n"); abort(); }
  const std::vector<int>& msg = g_iq.front();
  for (int j = 0; j < n; ++j) memcpy(A + (size_t)j * lda, &msg[(size_t)j * m], (size_t)m * sizeof(int));
  g_iq.pop_front();
}

// ------------------------------------------------------------------------- BLAS / LAPACK backends
#ifndef SHIM_INTERNAL_BLAS
void dsyrk_(char*, char*, int*, int*, double*, double*, int*, double*, double*, int*);
void dgemm_(char*, char*, int*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
void dsymm_(char*, char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*);
void dpotrf_(char*, int*, double*, int*, int*);
void dpotri_(char*, int*, double*, int*, int*);
void spotrf_(char*, int*, float*, int*, int*);
void spotri_(char*, int*, float*, int*, int*);
void dgetrf_(int*, int*, double*, int*, int*, int*);
void dgetri_(int*, double*, int*, int*, double*, int*, int*);
void dgeqrf_(int*, int*, double*, int*, double*, double*, int*, int*);
void dsyev_(char*, char*, int*, double*, int*, double*, double*, int*, int*);
#else
static inline bool is(char c, char u) { return c == u || c == (char)(u + 32); }
static double& el(double* a, int lda, int i, int j) { return a[(size_t)j * lda + i]; }
void dgemm_(char* ta, char* tb, int* m, int* n, int* k, double* alpha, double* a, int* lda, double* b, int* ldb,
            double* beta, double* c, int* ldc) {
  const bool tA = !is(*ta, 'N'), tB = !is(*tb, 'N');
  for (int j = 0; j < *n; ++j)
    for (int i = 0; i < *m; ++i) {
      double s = 0;
      for (int l = 0; l < *k; ++l) s += (tA ? el(a, *lda, l, i) : el(a, *lda, i, l)) * (tB ? el(b, *ldb, j, l) : el(b, *ldb, l, j));
      el(c, *ldc, i, j) = (*beta == 0 ? 0 : *beta * el(c, *ldc, i, j)) + *alpha * s;
    }
}
void dsyrk_(char* uplo, char* trans, int* n, int* k, double* alpha, double* a, int* lda, double* beta, double* c, int* ldc) {
  const bool lower = is(*uplo, 'L'), tr = !is(*trans, 'N');
  for (int j = 0; j < *n; ++j)
    for (int i = lower ? j : 0; i < (lower ? *n : j + 1); ++i) {
      double s = 0;
      for (int l = 0; l < *k; ++l) s += tr ? el(a, *lda, l, i) * el(a, *lda, l, j) : el(a, *lda, i, l) * el(a, *lda, j, l);
      el(c, *ldc, i, j) = (*beta == 0 ? 0 : *beta * el(c, *ldc, i, j)) + *alpha * s;
    }
}
static void no_lapack(const char* name) { fprintf(stderr, "shim: %s needs a LAPACK backend (built with SHIM_INTERNAL_BLAS)\n", name); abort(); }
void dsymm_(char*, char*, int*, int*, double*, double*, int*, double*, int*, double*, double*, int*) { no_lapack("dsymm"); }
void dpotrf_(char*, int*, double*, int*, int*) { no_lapack("dpotrf"); }
void dpotri_(char*, int*, double*, int*, int*) { no_lapack("dpotri"); }
void spotrf_(char*, int*, float*, int*, int*) { no_lapack("spotrf"); }
void spotri_(char*, int*, float*, int*, int*) { no_lapack("spotri"); }
void dgetrf_(int*, int*, double*, int*, int*, int*) { no_lapack("dgetrf"); }
void dgetri_(int*, double*, int*, int*, double*, int*, int*) { no_lapack("dgetri"); }
void dgeqrf_(int*, int*, double*, int*, double*, double*, int*, int*) { no_lapack("dgeqrf"); }
void dsyev_(char*, char*, int*, double*, int*, double*, double*, int*, int*) { no_lapack("dsyev"); }
#endif

// ------------------------------------------------------------------------------------ ScaLAPACK
// descriptor: {dtype, ctxt, m, n, mb, nb, rsrc, csrc, lld}
int numroc_(int* n, int*, int*, int*, int*) { return *n; }
void descinit_(int* desc, int* m, int* n, int* mb, int* nb, int* irsrc, int* icsrc, int* ictxt, int* lld, int* info) {
  desc[0] = 1; desc[1] = *ictxt; desc[2] = *m; desc[3] = *n; desc[4] = *mb; desc[5] = *nb; desc[6] = *irsrc;
  desc[7] = *icsrc; desc[8] = (*lld < 1) ? 1 : *lld;
  *info = 0;
}
static inline double* sub(double* a, int* ia, int* ja, int* desc) { return a + (size_t)(*ja - 1) * desc[8] + (*ia - 1); }
static inline float* subf(float* a, int* ia, int* ja, int* desc) { return a + (size_t)(*ja - 1) * desc[8] + (*ia - 1); }

void pdgemm_(char* ta, char* tb, int* m, int* n, int* k, double* alpha, double* a, int* ia, int* ja, int* desca,
             double* b, int* ib, int* jb, int* descb, double* beta, double* c, int* ic, int* jc, int* descc) {
  dgemm_(ta, tb, m, n, k, alpha, sub(a, ia, ja, desca), &desca[8], sub(b, ib, jb, descb), &descb[8], beta,
         sub(c, ic, jc, descc), &descc[8]);
}
void pdsymm_(char* side, char* uplo, int* m, int* n, double* alpha, double* a, int* ia, int* ja, int* desca, double* b,
             int* ib, int* jb, int* descb, double* beta, double* c, int* ic, int* jc, int* descc) {
  dsymm_(side, uplo, m, n, alpha, sub(a, ia, ja, desca), &desca[8], sub(b, ib, jb, descb), &descb[8], beta,
         sub(c, ic, jc, descc), &descc[8]);
}
void pdsyrk_(char* uplo, char* trans, int* n, int* k, double* alpha, double* a, int* ia, int* ja, int* desca,
             double* beta, double* c, int* ic, int* jc, int* descc) {
  dsyrk_(uplo, trans, n, k, alpha, sub(a, ia, ja, desca), &desca[8], beta, sub(c, ic, jc, descc), &descc[8]);
}
// C := beta*C + alpha*op(A), C is m x n
void pdgeadd_(char* trans, int* m, int* n, double* alpha, double* a, int* ia, int* ja, int* desca, double* beta,
              double* c, int* ic, int* jc, int* descc) {
  double* A = sub(a, ia, ja, desca);
  double* C = sub(c, ic, jc, descc);
  const int lda = desca[8], ldc = descc[8];
  const bool tr = !(*trans == 'N' || *trans == 'n');
  for (int j = 0; j < *n; ++j)
    for (int i = 0; i < *m; ++i) {
      const double av = tr ? A[(size_t)i * lda + j] : A[(size_t)j * lda + i];
      double& cv = C[(size_t)j * ldc + i];
      cv = (*beta == 0. ? 0. : *beta * cv) + *alpha * av;
    }
}
// C := beta*C + alpha*A^T, C is m x n, A is n x m
void pdtran_(int* m, int* n, double* alpha, double* a, int* ia, int* ja, int* desca, double* beta, double* c, int* ic,
             int* jc, int* descc) {
  char t = 'T';
  pdgeadd_(&t, m, n, alpha, a, ia, ja, desca, beta, c, ic, jc, descc);
}
void pdlaset_(char* uplo, int* m, int* n, double* alpha, double* beta, double* a, int* ia, int* ja, int* desca) {
  double* A = sub(a, ia, ja, desca);
  const int lda = desca[8];
  const bool up = (*uplo == 'U' || *uplo == 'u'), lo = (*uplo == 'L' || *uplo == 'l');
  for (int j = 0; j < *n; ++j)
    for (int i = 0; i < *m; ++i) {
      if (i == j) A[(size_t)j * lda + i] = *beta;
      else if ((up && i < j) || (lo && i > j) || (!up && !lo)) A[(size_t)j * lda + i] = *alpha;
    }
}
void pdlacpy_(char* uplo, int* m, int* n, double* a, int* ia, int* ja, int* desca, double* b, int* ib, int* jb,
              int* descb) {
  double* A = sub(a, ia, ja, desca);
  double* B = sub(b, ib, jb, descb);
  const int lda = desca[8], ldb = descb[8];
  const bool up = (*uplo == 'U' || *uplo == 'u'), lo = (*uplo == 'L' || *uplo == 'l');
  for (int j = 0; j < *n; ++j)
    for (int i = 0; i < *m; ++i)
      if ((up && i <= j) || (lo && i >= j) || (!up && !lo)) B[(size_t)j * ldb + i] = A[(size_t)j * lda + i];
}
void pdpotrf_(char* uplo, int* n, double* m, int* ia, int* ja, int* desca, int* info) { dpotrf_(uplo, n, sub(m, ia, ja, desca), &desca[8], info); }
void pdpotri_(char* uplo, int* n, double* m, int* ia, int* ja, int* desca, int* info) { dpotri_(uplo, n, sub(m, ia, ja, desca), &desca[8], info); }
void pspotrf_(char* uplo, int* n, float* m, int* ia, int* ja, int* desca, int* info) { spotrf_(uplo, n, subf(m, ia, ja, desca), &desca[8], info); }
void pspotri_(char* uplo, int* n, float* m, int* ia, int* ja, int* desca, int* info) { spotri_(uplo, n, subf(m, ia, ja, desca), &desca[8], info); }
void pdgetrf_(int* m, int* n, double* a, int* ia, int* ja, int* desca, int* ipiv, int* info) { dgetrf_(m, n, sub(a, ia, ja, desca), &desca[8], ipiv, info); }
void pdgetri_(int* n, double* a, int* ia, int* ja, int* desca, int* ipiv, double* work, int* lwork, int*, int* liwork, int* info) {
  if (*lwork == -1 || *liwork == -1) {  // workspace query
    double wq = 0; int lw = -1;
    dgetri_(n, sub(a, ia, ja, desca), &desca[8], ipiv, &wq, &lw, info);
    work[0] = wq;
    return;
  }
  dgetri_(n, sub(a, ia, ja, desca), &desca[8], ipiv, work, lwork, info);
}
void pdgeqrf_(int* m, int* n, double* a, int* ia, int* ja, int* desca, double* tau, double* work, int* lwork, int* info) {
  dgeqrf_(m, n, sub(a, ia, ja, desca), &desca[8], tau, work, lwork, info);
}
void pdsyev_(char* jobz, char* uplo, int* n, double* a, int* ia, int* ja, int* desca, double* w, double* z, int* iz,
             int* jz, int* descz, double* work, int* lwork, int* info) {
  // ScaLAPACK returns eigenvectors in Z and leaves A destroyed; LAPACK overwrites A.
  dsyev_(jobz, uplo, n, sub(a, ia, ja, desca), &desca[8], w, work, lwork, info);
  if (*lwork != -1 && (*jobz == 'V' || *jobz == 'v') && z) {
    double* A = sub(a, ia, ja, desca);
    double* Z = sub(z, iz, jz, descz);
    for (int j = 0; j < *n; ++j) memcpy(Z + (size_t)j * descz[8], A + (size_t)j * desca[8], (size_t)*n * sizeof(double));
  }
}

}  // extern "C"
