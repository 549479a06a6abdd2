// lp_engine.h -- internal interface between the CUDA engine (lp_engine.cu) and the host solver (lp_host.cpp).
#pragma once
#include <cstddef>
#include "../../include/abip_gpu.h"
#ifdef __CUDACC__
#include "lp_device.cuh"
#endif

// host-pointer helpers used by the linsys plugin symbols
int abipgpu_lp_spmv_host(abipgpu_lp* e, int trans, const double* x, double* y, int accumulate);
int abipgpu_lp_solve_host(abipgpu_lp* e, double* b, const double* s, long iter, int* cg_its);
ABIPGpuStats* abipgpu_lp_stats(abipgpu_lp* e);
int abipgpu_lp_dims(const abipgpu_lp* e, int* m, int* n);
int abipgpu_lp_sync(abipgpu_lp* e);
void abipgpu_lp_drop_pending(abipgpu_lp* e);
struct LpInnerArgs;
struct LpSolveArgs;
extern "C" int abipgpu_lp_is_batch(const abipgpu_lp* e);
extern "C" int abipgpu_lp_inner_loop(abipgpu_lp* e, const LpInnerArgs* L, abip_float* sc);
extern "C" int abipgpu_lp_solve_loop(abipgpu_lp* e, const LpSolveArgs* S, abip_float* sc);
extern "C" int abipgpu_lp_bb_search(abipgpu_lp* e, abip_int k, abip_float mu, int lookback, abip_float eps_cor, abip_float eps_pen, abip_float* sc);  // batch engines: forget deferred vector operations (failed solve)
int abipgpu_lp_solve_timer(abipgpu_lp* e, int stop);
extern "C" int abipgpu_lp_comm_export(abipgpu_lp* e, void* handle64);
extern "C" int abipgpu_lp_comm_connect(abipgpu_lp* e, int G, int rank, const void* handles);
extern "C" void abipgpu_lp_set_global_n(abipgpu_lp* e, long n_global);
extern "C" void abipgpu_lp_request_grid(int ctas);
extern "C" void abipgpu_lp_request_order(int on);
extern "C" int abipgpu_equilibrate(abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai, abip_float* Ax, const ABIPSettings* stgs,
                                   int device, abip_float* D, abip_float* E, abip_float* mean_norm_row_A, abip_float* mean_norm_col_A);
// lock-step batch executor (lp_engine.cu: BatchExec)
extern "C" void* abipgpu_batch_begin(int device, int capacity);
extern "C" void abipgpu_batch_attach(void* b);
extern "C" void abipgpu_batch_reserve(void* b, std::size_t device_bytes, int engines);
extern "C" int abipgpu_batch_attached();
extern "C" void abipgpu_batch_end(void* b, long* launches, long* items);
void abipgpu_lp_batch_solving(abipgpu_lp* e, int delta);
