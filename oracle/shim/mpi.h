/* Single-rank MPI stand-in used ONLY to build the unmodified reference sources into the parity
 * oracle (oracle/_ref).  Test infrastructure; not part of the product.  Declares exactly the MPI
 * subset the reference tree uses; shim.cpp implements it for P = 1. */
#ifndef DGRM_ORACLE_SHIM_MPI_H
#define DGRM_ORACLE_SHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Group;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef struct shim_mpi_file* MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_LONG 3
#define MPI_DOUBLE 4

#define MPI_DISTRIBUTE_CYCLIC 1
#define MPI_ORDER_FORTRAN 1

#define MPI_MODE_RDONLY 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_EXCL 8

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Get_processor_name(char* name, int* len);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Gather(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, int rcount, MPI_Datatype rtype, int root,
               MPI_Comm comm);
int MPI_Gatherv(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, const int* rcounts, const int* displs,
                MPI_Datatype rtype, int root, MPI_Comm comm);
int MPI_Alltoall(const void* sbuf, int scount, MPI_Datatype stype, void* rbuf, int rcount, MPI_Datatype rtype,
                 MPI_Comm comm);
int MPI_Alltoallv(const void* sbuf, const int* scounts, const int* sdispls, MPI_Datatype stype, void* rbuf,
                  const int* rcounts, const int* rdispls, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Comm_group(MPI_Comm comm, MPI_Group* group);
int MPI_Group_incl(MPI_Group group, int n, const int* ranks, MPI_Group* newgroup);
int MPI_Group_rank(MPI_Group group, int* rank);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm* newcomm);
int MPI_Type_create_darray(int size, int rank, int ndims, const int* gsizes, const int* distribs, const int* dargs,
                           const int* psizes, int order, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Type_free(MPI_Datatype* type);
int MPI_File_open(MPI_Comm comm, const char* filename, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_delete(const char* filename, MPI_Info info);
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char* datarep,
                      MPI_Info info);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_File_write_all(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_File_write_at_all(MPI_File fh, MPI_Offset offset, const void* buf, int count, MPI_Datatype type,
                          MPI_Status* status);
int MPI_File_read_all(MPI_File fh, void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_Error_class(int errorcode, int* errorclass);
int MPI_Error_string(int errorcode, char* string, int* resultlen);

#ifdef __cplusplus
}
#endif
#endif
