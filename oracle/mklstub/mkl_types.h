/* TEST INFRASTRUCTURE: declarations-only stand-in for Intel MKL headers, so that the reference ABIP-QCP sources
 * (which include mkl*.h unconditionally: src/abip-qcp/include/linsys.h:14-18, cones.h:11-12) compile without MKL.
 * Only the MKL-free back end (linsys_solver = 1, vendored QDLDL) is ever exercised; the stubs abort. */
#ifndef MKLSTUB_TYPES_H
#define MKLSTUB_TYPES_H
typedef int MKL_INT;
typedef int _INTEGER_t;
typedef void *_MKL_DSS_HANDLE_t;
#define MKL_DSS_ZERO_BASED_INDEXING 131072
#define MKL_DSS_DEFAULTS 0
#define MKL_DSS_SYMMETRIC 536870976
#define MKL_DSS_INDEFINITE 134217856
#define MKL_DSS_SUCCESS 0
#define LAPACK_COL_MAJOR 102
#define LAPACK_ROW_MAJOR 101
int mklstub_dss_create(_MKL_DSS_HANDLE_t *h, MKL_INT *opt);
int mklstub_dss_delete(_MKL_DSS_HANDLE_t *h, MKL_INT *opt);
#define dss_create(handle, opt) mklstub_dss_create(&(handle), &(opt))
#define dss_delete(handle, opt) mklstub_dss_delete(&(handle), &(opt))
int dss_define_structure(_MKL_DSS_HANDLE_t h, MKL_INT sym, const MKL_INT *p, MKL_INT m, MKL_INT n, const MKL_INT *i, MKL_INT nnz);
int dss_reorder(_MKL_DSS_HANDLE_t h, MKL_INT opt, const MKL_INT *perm);
int dss_factor_real(_MKL_DSS_HANDLE_t h, MKL_INT type, const void *x);
int dss_solve_real(_MKL_DSS_HANDLE_t h, MKL_INT opt, const void *b, MKL_INT nrhs, void *x);
void PARDISO(void *pt, const MKL_INT *maxfct, const MKL_INT *mnum, const MKL_INT *mtype, const MKL_INT *phase,
             const MKL_INT *n, const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *perm, const MKL_INT *nrhs,
             MKL_INT *iparm, const MKL_INT *msglvl, void *b, void *x, MKL_INT *error);
int LAPACKE_dpotrf(int layout, char uplo, int n, double *a, int lda);
int LAPACKE_dpotrs(int layout, char uplo, int n, int nrhs, const double *a, int lda, double *b, int ldb);
#endif
