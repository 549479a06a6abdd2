/* TEST INFRASTRUCTURE: aborting stand-ins for the MKL entry points referenced by the reference's direct back ends. */
#include <stdio.h>
#include <stdlib.h>
#include "mkl_types.h"
static void die(const char *f) { fprintf(stderr, "mklstub: %s called (MKL is not available; use linsys_solver=1)\n", f); abort(); }
int mklstub_dss_create(_MKL_DSS_HANDLE_t *h, MKL_INT *opt) { die("dss_create"); return -1; }
int mklstub_dss_delete(_MKL_DSS_HANDLE_t *h, MKL_INT *opt) { die("dss_delete"); return -1; }
int dss_define_structure(_MKL_DSS_HANDLE_t h, MKL_INT sym, const MKL_INT *p, MKL_INT m, MKL_INT n, const MKL_INT *i, MKL_INT nnz) { die("dss_define_structure"); return -1; }
int dss_reorder(_MKL_DSS_HANDLE_t h, MKL_INT opt, const MKL_INT *perm) { die("dss_reorder"); return -1; }
int dss_factor_real(_MKL_DSS_HANDLE_t h, MKL_INT type, const void *x) { die("dss_factor_real"); return -1; }
int dss_solve_real(_MKL_DSS_HANDLE_t h, MKL_INT opt, const void *b, MKL_INT nrhs, void *x) { die("dss_solve_real"); return -1; }
void PARDISO(void *pt, const MKL_INT *maxfct, const MKL_INT *mnum, const MKL_INT *mtype, const MKL_INT *phase,
             const MKL_INT *n, const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *perm, const MKL_INT *nrhs,
             MKL_INT *iparm, const MKL_INT *msglvl, void *b, void *x, MKL_INT *error) { die("PARDISO"); }
int LAPACKE_dpotrf(int layout, char uplo, int n, double *a, int lda) { die("LAPACKE_dpotrf"); return -1; }
int LAPACKE_dpotrs(int layout, char uplo, int n, int nrhs, const double *a, int lda, double *b, int ldb) { die("LAPACKE_dpotrs"); return -1; }
