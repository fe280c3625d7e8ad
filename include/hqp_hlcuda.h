/*
 * hqp_hlcuda.h -- C ABI of libhqphl.so: the block-diagonal BFGS update of the
 * Lagrangian Hessian on an NVIDIA B200.
 *
 * Replaces, one call for ALL diagonal blocks of Q, the per-block loop of
 *   Hqp_HL_BFGS::update      hqp/Hqp_HL_BFGS.C:216-243   (next_block / extract / insert)
 *   Hqp_HL_BFGS::update_b_Q  hqp/Hqp_HL_BFGS.C:149-213   (Powell-damped BFGS + eigenvalue control,
 *                                                          meschach/symmeig.c:174 per block)
 * The caller (an Hqp_HL module) packs the blocks it walks with
 * next_block() into one array: block b is bsize[b] x bsize[b] doubles, row-major, BOTH
 * triangles (what symsp_extract_mat hands to update_b_Q), the blocks one after the other;
 * s and u are the SQP step and gradient difference restricted to the same variables, in
 * block order (b_s.ve = s->ve + offs).  On return the blocks hold the updated Hessian
 * (upper triangle authoritative, as symsp_insert_symmat reads it).
 *
 * Plain C, host pointers unless the name ends in _dev, FP64, no exceptions; every call
 * returns a status.  There is no CPU fallback: without a CUDA device the calls return
 * HQPHL_E_CUDA.
 */
#ifndef HQP_HLCUDA_H
#define HQP_HLCUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define HQPHL_OK 0
#define HQPHL_E_ARG 1
#define HQPHL_E_UNSUPPORTED 2 /* a block too large for the shared-memory kernel (> ~110) */
#define HQPHL_E_CUDA 100      /* CUDA runtime error; see hqphl_last_error() */

const char *hqphl_last_error(void);

/* gamma: sqp_hela_gamma (>= 0: fixed damping; < 0: adapted to the step length alpha,
 * hqp/Hqp_HL_BFGS.C:163-172); eps: Hqp_HL::_eps; eigen_control: sqp_hela_eigen_control.
 * info3 (may be NULL): [0] blocks whose diagonal was shifted, [1] blocks skipped because
 * s'v or s'Qs vanished (:186-187), [2] blocks whose eigenvalue iteration hit its sweep limit. */
int hqphl_bfgs_update(int device, int nblocks, const int *bsize, double *Q, const double *s,
                      const double *u, double alpha, double gamma, double eps, int eigen_control,
                      int *info3);

/* Everything already on the device (a device-resident SQP layer, bench.py): d_qoff[b] /
 * d_voff[b] = offset of block b in d_Q / in d_s, d_u; d_info3 is accumulated into (zero
 * it first); asynchronous on cuda_stream. */
int hqphl_bfgs_update_dev(void *cuda_stream, int nblocks, int max_bsize, const int *d_bsize,
                          const long long *d_qoff, const int *d_voff, double *d_Q,
                          const double *d_s, const double *d_u, double alpha, double gamma,
                          double eps, int eigen_control, int *d_info3);

#ifdef __cplusplus
}
#endif
#endif
