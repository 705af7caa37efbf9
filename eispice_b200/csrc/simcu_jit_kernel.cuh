/*
 * simcu_jit_kernel.cuh - the production kernel: one batched transient, start
 * to finish, in ONE launch.  Compiled per netlist topology by NVRTC together
 * with simcu_device.cuh and the generated gen*() functions (SIMCU_JIT).
 *
 * Persistent grid: every warp repeatedly claims the next 32 instances from a
 * global counter and runs each of them through load -> operating point ->
 * every timestep (core/simulator.c:75-248), so a thread's whole working set -
 * stamps, device state, factors, iterate - is thread-local (registers + the
 * per-resident-thread local memory window) and HBM only sees the per-instance
 * parameters coming in and the gridded probes / statistics going out.  The
 * T-line history ring is the one state that needs dynamic addressing; it lives
 * in global memory with one column per RESIDENT thread ([record][row][slot]).
 */
#ifndef SIMCU_JIT_KERNEL_CUH
#define SIMCU_JIT_KERNEL_CUH

DEV void jitRunInstance(const SimcuArgs &a, int inst, int slot, size_t hstride, int opOnly)
{
	Inst c(a, inst, slot, hstride);
	const SimcuControl &k = a.ctl;
	const size_t M = a.M;

	int hrows[SIMCU_DIM(SIMCU_NH)], prows[SIMCU_DIM(SIMCU_NPROBES)];
	genHistRows(hrows);
	genProbeRows(prows);
	double meas[SIMCU_DIM(SIMCU_NMEASURES)];

	/* parameters of this instance: broadcast constants or its own column */
	SIMCU_UNROLL
	for (int p = 0; p < SIMCU_NPARAM; p++) {
		const int m = __ldg(&a.pmap[p]);
		c.par[p] = (m >= 0) ? a.pinst[(size_t)m * M + inst] : __ldg(&a.pconst[p]);
	}
	SIMCU_UNROLL
	for (int s = 0; s < SIMCU_NISET; s++) c.tset[s] = a.iinst[(size_t)s * M + inst];

	/* matrixClear (core/simulator.c:120) on thread-local storage */
	SIMCU_UNROLL
	for (int s = 0; s < SIMCU_NNZ; s++) c.Av[s] = 0.0;
	SIMCU_UNROLL
	for (int r = 0; r < SIMCU_N; r++) { c.Bv[r] = 0.0; c.Xv[r] = 0.0; c.Xa[r] = 0.0; }
	SIMCU_UNROLL
	for (int s = 0; s < SIMCU_SD; s++) c.sv[s] = 0.0;
	SIMCU_UNROLL
	for (int s = 0; s < SIMCU_SI; s++) c.iv[s] = 0;
	c.itN = 0; c.itAdv = false;
	SIMCU_UNROLL
	for (int j = 0; j < 4; j++) c.itT[j] = 0.0;
	SIMCU_UNROLL
	for (int j = 0; j < 3; j++) c.itH[j] = 0.0;

	double time = 0.0, prevTime = 0.0, thisStep = 0.0, maxStep = 0.0, lteStep = 0.0;
	int order = 1, brk = 0;
	int nfact = 0, nfactOp = 0, accepted = 0, rejLte = 0, rejNc = 0, pivotBreaks = 0;
	long long attempts = 0;
	bool transient = false;      /* false: the operating-point solve */

	c.time = 0.0;
	genLoad(c);                                         /* simulator.c:121-123 */

	while (!c.err) {
		if (transient) {
			if (attempts >= a.maxAttempts) { c.err = SIMCU_ERR_BUDGET; break; }
			attempts++;
			time = prevTime + thisStep;                 /* :157-161 */
			if (time > k.tstop) { thisStep = k.tstop - prevTime; time = k.tstop; }
			c.time = time; c.order = order;
			brk = genStep(c);                           /* :167-173 */
			if (c.err) break;
			if (brk) { order = 1; c.order = 1; }
			genIntegrate(c);                            /* :178-179 */
			if (c.err) break;
		}

		/* simulatorSolve (simulator.c:45-69): solve, then {linearize; solve}
		 * until every device is linear or the limit is reached */
		const int limit = transient ? k.itl4 : k.itl1;
		int count = 0;
		for (;;) {
			if (transient) { genFactorSolve1(c); nfact++; }
			else { genFactorSolve0(c); nfactOp++; }
			/* pivot breakdown in a transient attempt (the fixed sequence met a
			 * zero / non-finite pivot): reject the attempt as non-convergent, so
			 * the step is cut and Newton restarts - per instance, on device */
			if (transient && c.err == SIMCU_ERR_SINGULAR) {
				c.err = 0; count = limit; pivotBreaks++;
				break;
			}
			if (c.err) break;
			if (count++ >= limit) break;
			const int linear = genLinearize(c);
			if (c.err || linear) break;
		}
		if (c.err) break;

		if (!transient) {                               /* :126-137 */
			if (count > k.itl1) { c.err = SIMCU_ERR_OP; break; }
			gridOutput(c, prows, 0.0, 0.0, true);
			genMeasures(c, meas, 0.0, 0.0, true);
			SIMCU_UNROLL
			for (int r = 0; r < SIMCU_N; r++) c.Xa[r] = c.Xv[r];
			if (opOnly || !(0.0 < k.tstop)) { recordStep(c, hrows, 0.0); break; }
			genInitStep(c);
			recordStep(c, hrows, 0.0);                  /* :144 */
			maxStep = ((k.tstep < k.tmax) ? k.tstep : k.tmax) / 100;   /* :109 */
			thisStep = maxStep;
			genNextStep(c, thisStep);                   /* :152-154 */
			order = k.maxorder;                         /* :133 */
			transient = true;
			continue;
		}

		if (count < k.itl4) {
			lteStep = k.tmax;                           /* :188-190 */
			genMinStep(c, lteStep);
			if (c.err) break;
			if (lteStep < (0.9 * thisStep)) {           /* :191-198 */
				if (lteStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; break; }
				order = 1;
				thisStep = lteStep;
				rejLte++;
			} else {                                    /* accepted: :220-233 */
				gridOutput(c, prows, prevTime, time, false);
				genMeasures(c, meas, prevTime, time, false);
				SIMCU_UNROLL
				for (int r = 0; r < SIMCU_N; r++) c.Xa[r] = c.Xv[r];
				order = (order < k.maxorder) ? order + 1 : order;
				prevTime = time;
				if (brk) maxStep = Min3(lteStep, 0.1 * maxStep, k.tmax);
				else maxStep = Min3(lteStep, 2 * maxStep, k.tmax);
				accepted++;
				recordStep(c, hrows, time);             /* :144 of the next pass */
				if (!(time < k.tstop)) break;
				thisStep = maxStep;                     /* :152-154 */
				genNextStep(c, thisStep);
				continue;
			}
		} else {                                        /* :201-214 */
			thisStep = thisStep / 8;
			if (thisStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; break; }
			order = (order > 1) ? order - 1 : order;
			maxStep = maxStep / 8;
			rejNc++;
		}
		/* matrixRecall (:217): back to the last accepted solution */
		SIMCU_UNROLL
		for (int r = 0; r < SIMCU_N; r++) c.Xv[r] = c.Xa[r];
	}

	SIMCU_UNROLL
	for (int q = 0; q < SIMCU_NMEASURES; q++) a.outMeasure[(size_t)q * M + inst] = meas[q];
	int *ci = a.CI + inst;
	ci[(size_t)CI_STATUS * M] = c.err;
	ci[(size_t)CI_DONE * M] = 1;
	ci[(size_t)CI_ACCEPTED * M] = accepted;
	ci[(size_t)CI_REJ_LTE * M] = rejLte;
	ci[(size_t)CI_REJ_NC * M] = rejNc;
	ci[(size_t)CI_NFACT * M] = nfact;
	ci[(size_t)CI_NFACT_OP * M] = nfactOp;
	ci[(size_t)CI_NPTS * M] = c.npts;
	ci[(size_t)CI_BREAK * M] = pivotBreaks;
}

extern "C" __global__ void __launch_bounds__(SIMCU_TPB, SIMCU_MINB) simcuJitKernel(const SimcuArgs a, int opOnly)
{
	/* a.lanes < 32 spreads a batch that is too small to fill the machine over
	 * more warps (fewer instances per warp): the kernel is latency-bound, so
	 * idle schedulers cost more than idle lanes, and divergence shrinks too */
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int slot = warp * a.lanes + lane;
	const size_t hstride = (size_t)((gridDim.x * blockDim.x) >> 5) * a.lanes;
	for (;;) {
		int base = 0;
		if (lane == 0) base = atomicAdd(a.counter, a.lanes);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= a.count) break;
		if (lane < a.lanes && base + lane < a.count) {
			/* instances are processed grouped by table set (IBIS corner), so the
			 * lanes of a warp follow similar waveforms and diverge less */
			const int inst = a.order ? __ldg(&a.order[base + lane]) : base + lane;
			jitRunInstance(a, inst, slot, hstride, opOnly);
		}
		__syncwarp();
	}
}

#endif
