/*
 * simcu_kernels.cu - the GENERIC kernels of the batched transient engine
 * (sm_100a): table-driven versions of load, operating point and transient that
 * walk the device records, the LU op-stream and the expression bytecode from
 * global memory.  They run the pilot pass (the matrices that decide the pivot
 * sequences must exist before a topology-specialised kernel can be generated,
 * simcu_api.cpp: pilotPass) and serve as an independent implementation in the
 * tests (option interpreter = 1).  Production batches run simcu_jit_kernel.cuh.
 *
 * One thread owns one instance.  All per-instance data is structure-of-arrays
 * with the instance index fastest (simcu_program.h), so every stamp, state
 * access, LU load and solution store of a warp is one coalesced line.  The
 * numeric LU works on a per-thread workspace of F + N doubles in shared memory
 * ([slot][blockDim], conflict-free) whenever it fits.  The device functions
 * themselves are shared with the specialised kernel: simcu_device.cuh.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "simcu_calc.h"
#include "simcu_program.h"
#include "simcu_kernels.h"

#include "simcu_device.cuh"

/* ------------------------------------------------------------------------ *
 * Kernels
 * ------------------------------------------------------------------------ */
extern __shared__ double simcuSmem[];

DEV double *workspaceOf(const SimcuArgs &a, int inst, int &stride)
{
	if (a.wsShared) { stride = blockDim.x; return simcuSmem + threadIdx.x; }
	stride = a.M;
	return a.wsGlobal + inst;
}

/* matrixClear + deviceLoad for every device: core/simulator.c:120-123 */
__global__ void simcuLoadKernel(SimcuArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.M) return;
	int stride;
	double *ws = workspaceOf(a, i, stride);
	Inst c(a, i, ws, stride);
	for (int s = 0; s < a.nnzA; s++) a.A[(size_t)s * a.M + i] = 0.0;
	for (int r = 0; r < a.N; r++) { a.B[(size_t)r * a.M + i] = 0.0; a.X[(size_t)r * a.M + i] = 0.0; }
	for (int k = 0; k < CI_NUM; k++) c.CI(k) = 0;
	for (int k = 0; k < CD_NUM; k++) c.CD(k) = 0.0;
	c.time = 0.0;
	for (int k = 0; k < a.ndev && !c.err; k++) devLoad(c, a.dev + k * SIMCU_DEV_STRIDE);
	if (c.err) { c.CI(CI_STATUS) = c.err; c.CI(CI_DONE) = 1; }
}

/* operating point + transient prologue: core/simulator.c:126-137,144-154 */
__global__ void simcuOpKernel(SimcuArgs a, int opOnly)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.M) return;
	int stride;
	double *ws = workspaceOf(a, i, stride);
	Inst c(a, i, ws, stride);
	if (c.CI(CI_DONE)) return;
	c.time = 0.0;
	int nfact = 0;
	struct Sync { Inst &c; DEV ~Sync() { c.CI(CI_HCOUNT) = c.hcount; c.CI(CI_NPTS) = c.npts; c.CI(CI_GRIDIDX) = c.gridIdx; } } sync{c};
	const int count = newtonSolve(c, a.ctl.itl1, nfact);
	c.CI(CI_NFACT_OP) = nfact;
	if (!c.err && count > a.ctl.itl1) c.err = SIMCU_ERR_OP;
	if (c.err) { c.CI(CI_STATUS) = c.err; c.CI(CI_DONE) = 1; return; }
	gridOutput(c, a.probeRows, 0.0, 0.0, true);
	storeSolution(c);
	c.xInWs = false;
	if (opOnly) { recordStep(c, a.histRows, 0.0); c.CI(CI_DONE) = 1; return; }
	for (int k = 0; k < a.ndev; k++) devInitStep(c, a.dev + k * SIMCU_DEV_STRIDE);
	recordStep(c, a.histRows, 0.0);
	const double tstep = a.ctl.tstep, tmax = a.ctl.tmax;
	const double maxStep = ((tstep < tmax) ? tstep : tmax) / 100;   /* simulator.c:109 */
	double thisStep = maxStep;
	for (int k = 0; k < a.ndev; k++) devNextStep(c, a.dev + k * SIMCU_DEV_STRIDE, thisStep);
	c.CD(CD_TIME) = 0.0; c.CD(CD_PREVTIME) = 0.0;
	c.CD(CD_MAXSTEP) = maxStep; c.CD(CD_THISSTEP) = thisStep; c.CD(CD_LTESTEP) = 0.0;
	c.CI(CI_ORDER) = a.ctl.maxorder;                                /* simulator.c:133 */
	if (!(0.0 < a.ctl.tstop)) c.CI(CI_DONE) = 1;
}

/* the per-timestep Newton loop: core/simulator.c:142-235, one thread per
 * instance, each with its own time, step and integration order */
__global__ void simcuTranKernel(SimcuArgs a)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.M) return;
	int stride;
	double *ws = workspaceOf(a, i, stride);
	Inst c(a, i, ws, stride);
	if (c.CI(CI_DONE)) return;
	c.hcount = c.CI(CI_HCOUNT); c.npts = c.CI(CI_NPTS); c.gridIdx = c.CI(CI_GRIDIDX);

	const SimcuControl &k = a.ctl;
	double time = c.CD(CD_TIME), prevTime = c.CD(CD_PREVTIME);
	double thisStep = c.CD(CD_THISSTEP), maxStep = c.CD(CD_MAXSTEP);
	double lteStep = c.CD(CD_LTESTEP);
	int order = c.CI(CI_ORDER);
	int nfact = c.CI(CI_NFACT), accepted = c.CI(CI_ACCEPTED);
	int rejLte = c.CI(CI_REJ_LTE), rejNc = c.CI(CI_REJ_NC);
	int attempts = c.CI(CI_ATTEMPTS);
	bool done = false;

	for (int att = 0; att < a.attemptsPerLaunch && !done; att++) {
		if ((long long)attempts >= a.maxAttempts) { c.err = SIMCU_ERR_BUDGET; break; }
		attempts++;
		time = prevTime + thisStep;                         /* simulator.c:157-161 */
		if (time > k.tstop) { thisStep = k.tstop - prevTime; time = k.tstop; }
		c.time = time; c.order = order; c.xInWs = false;

		int brk = 0;                                        /* :167-173 */
		for (int n = 0; n < a.ndev && !c.err; n++) brk |= devStep(c, a.dev + n * SIMCU_DEV_STRIDE);
		if (c.err) break;
		if (brk) { order = 1; c.order = 1; }
		for (int n = 0; n < a.ndev && !c.err; n++) devIntegrate(c, a.dev + n * SIMCU_DEV_STRIDE);
		if (c.err) break;
		if (a.stopBeforeSolve) { done = true; break; }

		const int count = newtonSolve(c, k.itl4, nfact);    /* :182 */
		if (c.err) break;
		if (count < k.itl4) {
			lteStep = k.tmax;                               /* :188-190 */
			for (int n = 0; n < a.ndev && !c.err; n++) devMinStep(c, a.dev + n * SIMCU_DEV_STRIDE, lteStep);
			if (c.err) break;
			if (lteStep < (0.9 * thisStep)) {               /* :191-198 */
				if (lteStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; break; }
				order = 1;
				thisStep = lteStep;
				rejLte++;
				continue;                                   /* X is untouched = recall */
			}
			/* accepted: :220-233, then the next pass's record + look-ahead */
			gridOutput(c, a.probeRows, prevTime, time, false);
			storeSolution(c);
			c.xInWs = false;
			order = (order < k.maxorder) ? order + 1 : order;
			prevTime = time;
			if (brk) maxStep = Min3(lteStep, 0.1 * maxStep, k.tmax);
			else maxStep = Min3(lteStep, 2 * maxStep, k.tmax);
			accepted++;
			recordStep(c, a.histRows, time);
			if (!(time < k.tstop)) { done = true; break; }
			thisStep = maxStep;                             /* :152-154 */
			for (int n = 0; n < a.ndev; n++) devNextStep(c, a.dev + n * SIMCU_DEV_STRIDE, thisStep);
		} else {                                            /* :201-214 */
			thisStep = thisStep / 8;
			if (thisStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; break; }
			order = (order > 1) ? order - 1 : order;
			maxStep = maxStep / 8;
			rejNc++;
		}
	}

	c.CD(CD_TIME) = time; c.CD(CD_PREVTIME) = prevTime;
	c.CD(CD_THISSTEP) = thisStep; c.CD(CD_MAXSTEP) = maxStep; c.CD(CD_LTESTEP) = lteStep;
	c.CI(CI_ORDER) = order; c.CI(CI_NFACT) = nfact; c.CI(CI_ACCEPTED) = accepted;
	c.CI(CI_REJ_LTE) = rejLte; c.CI(CI_REJ_NC) = rejNc; c.CI(CI_ATTEMPTS) = attempts;
	c.CI(CI_HCOUNT) = c.hcount; c.CI(CI_NPTS) = c.npts; c.CI(CI_GRIDIDX) = c.gridIdx;
	if (c.err) { c.CI(CI_STATUS) = c.err; done = true; }
	if (done) c.CI(CI_DONE) = 1;
	else atomicAdd(a.remaining, 1);
}

/* [ngrid][nprobes][M] -> [M][ngrid][nprobes] for the host-facing result */
__global__ void simcuTransposeKernel(const double *src, double *dst, int M, int rows)
{
	__shared__ double tile[32][33];
	const int m0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const int r = r0 + j, m = m0 + threadIdx.x;
		if (r < rows && m < M) tile[j][threadIdx.x] = src[(size_t)r * M + m];
	}
	__syncthreads();
	for (int j = threadIdx.y; j < 32; j += blockDim.y) {
		const int m = m0 + j, r = r0 + threadIdx.x;
		if (r < rows && m < M) dst[(size_t)m * rows + r] = tile[threadIdx.x][j];
	}
}

/* test hook: the expression interpreter on the device */
__global__ void simcuCalcKernel(const int *code, int n, const double *consts,
		const double *vals, int nvars, double m, double *out)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	double lv[CALC_MAX_VARS];
	for (int k = 0; k < nvars && k < CALC_MAX_VARS; k++) lv[k] = vals[k];
	double f, d;
	calcRun(code, n, consts, lv, -1, m, &f, &d);
	out[0] = f;
	for (int k = 0; k < nvars; k++) {
		calcRun(code, n, consts, lv, k, m, &f, &d);
		out[1 + k] = d;
	}
}

/* ---- launch helpers (called from simcu_api.cpp) --------------------------- */
static size_t smemBytes(const SimcuArgs &a, int tpb)
{
	return a.wsShared ? (size_t)(a.F + a.N) * tpb * sizeof(double) : 0;
}

extern "C" int simcuLaunchLoad(const SimcuArgs *a, int tpb, void *stream)
{
	const int blocks = (a->M + tpb - 1) / tpb;
	simcuLoadKernel<<<blocks, tpb, 0, (cudaStream_t)stream>>>(*a);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int simcuLaunchOp(const SimcuArgs *a, int tpb, int opOnly, void *stream)
{
	const int blocks = (a->M + tpb - 1) / tpb;
	const size_t sm = smemBytes(*a, tpb);
	if (sm > 48 * 1024 &&
			cudaFuncSetAttribute(simcuOpKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
		return -1;
	simcuOpKernel<<<blocks, tpb, sm, (cudaStream_t)stream>>>(*a, opOnly);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int simcuLaunchTran(const SimcuArgs *a, int tpb, void *stream)
{
	const int blocks = (a->M + tpb - 1) / tpb;
	const size_t sm = smemBytes(*a, tpb);
	if (sm > 48 * 1024 &&
			cudaFuncSetAttribute(simcuTranKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
		return -1;
	simcuTranKernel<<<blocks, tpb, sm, (cudaStream_t)stream>>>(*a);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int simcuLaunchTranspose(const double *src, double *dst, int M, int rows, void *stream)
{
	dim3 grid((M + 31) / 32, (rows + 31) / 32), block(32, 8);
	simcuTransposeKernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, dst, M, rows);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int simcuLaunchCalc(const int *code, int n, const double *consts,
		const double *vals, int nvars, double m, double *out, void *stream)
{
	simcuCalcKernel<<<1, 32, 0, (cudaStream_t)stream>>>(code, n, consts, vals, nvars, m, out);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
