/*
 * simcu_kernels.h - launch helpers exported by simcu_kernels.cu to the host
 * side of libsimulator_cuda (simcu_api.cpp).  Plain C linkage, no CUDA types.
 */
#ifndef SIMCU_KERNELS_H
#define SIMCU_KERNELS_H

#include "../../include/simulator_cuda.h"
#include "simcu_program.h"

#ifdef __cplusplus
extern "C" {
#endif

int simcuLaunchLoad(const SimcuArgs *a, int tpb, void *stream);
int simcuLaunchOp(const SimcuArgs *a, int tpb, int opOnly, void *stream);
int simcuLaunchTran(const SimcuArgs *a, int tpb, void *stream);
int simcuLaunchTranspose(const double *src, double *dst, int M, int rows,
		void *stream);
int simcuLaunchCalc(const int *code, int n, const double *consts,
		const double *vals, int nvars, double m, double *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
