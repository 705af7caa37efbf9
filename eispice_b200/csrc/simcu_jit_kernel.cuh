/*
 * simcu_jit_kernel.cuh - the production kernel: one batched transient, start
 * to finish, in ONE launch.  Compiled per netlist topology by NVRTC together
 * with simcu_device.cuh and the generated gen*() functions (SIMCU_JIT).
 *
 * Persistent grid, one thread = one instance at a time.  Every LANE claims its
 * next instance from a global queue as soon as it has finished the previous one
 * (a.claimBelow decides how eagerly), so a warp never idles behind its slowest
 * instance: the lanes of a warp run attempts in lock-step - each loop trip is
 * one attempt (core/simulator.c:156-218) of every lane, or the operating point
 * of a lane that has just claimed - but at different simulated times.
 * A thread's whole working set - stamps, device state, factors, iterate - is
 * thread-local (registers + the per-resident-thread local memory window) and
 * HBM only sees the per-instance parameters coming in and the gridded probes /
 * statistics going out.  The T-line history ring is the one state that needs
 * dynamic addressing; it lives in global memory with one ring per RESIDENT
 * thread ([thread][record][1 + rows]).
 *
 * SIMCU_REPLAY (test instrumentation, option "replay"): instead of taking its
 * own accept / reject decisions the loop replays a per-instance attempt log -
 * (time, order, outcome) of every attempt the CPU oracle made - so that both
 * sides walk the same sequence of attempts and every unknown at every accepted
 * step can be compared at 1e-6 / 1e-9.  Its own decision is still evaluated
 * and disagreements are counted (CI_REPLAY_DIFF).
 */
#ifndef SIMCU_JIT_KERNEL_CUH
#define SIMCU_JIT_KERNEL_CUH

#ifndef SIMCU_REPLAY
#define SIMCU_REPLAY 0
#endif
#ifndef SIMCU_BLOCKSYNC
#define SIMCU_BLOCKSYNC 0
#endif
#ifndef SIMCU_SYNCGROUP
#define SIMCU_SYNCGROUP 0
#endif
#ifndef SIMCU_WINDOWS
#define SIMCU_WINDOWS 0
#endif
#ifndef SIMCU_SMEMTAB
#define SIMCU_SMEMTAB 0
#endif

/* statistics and status of a finished (or failed) instance; the data of an
 * instance with status != 0 is NaN from the failure time on */
DEV void jitFinish(const SimcuArgs &a, Inst &c, const double *meas, int accepted, int rejLte,
		int rejNc, int nfact, int nfactOp, int pivotBreaks, int replayDiff)
{
	const size_t M = a.M;
	const int inst = c.i;
	if (c.err) {
		const double nan = HUGE_VAL - HUGE_VAL;
		for (int g = c.gridIdx; g < a.ngrid; g++)
			for (int p = 0; p < SIMCU_NPROBES; p++)
				a.outGrid[((size_t)g * a.nprobes + p) * M + inst] = nan;
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
	ci[(size_t)CI_REPLAY_DIFF * M] = replayDiff;
	ci[(size_t)CI_LIN_DIFF * M] = c.linDiff;
	ci[(size_t)CI_GROWTH * M] = c.growth;
}

extern "C" __global__ void __launch_bounds__(SIMCU_TPB, SIMCU_MINB) simcuJitKernel(const SimcuArgs a, int opOnly)
{
	/* a.lanes < 32 spreads a batch that is too small to fill the machine over
	 * more warps (fewer instances per warp) */
#if SIMCU_SMEMTAB
	/* IBIS tables in shared memory (north_star): every block stages the tables the
	 * Newton loop looks up - one coalesced copy, then per-thread interpolation
	 * reads them at shared-memory latency instead of competing with the spilled
	 * state for L1 */
	{
		const double2 *src = reinterpret_cast<const double2 *>(a.tabStage);
		double2 *dst = reinterpret_cast<double2 *>(simcuSmemTab);
		for (int i = threadIdx.x; i < 2 * a.stagePoints; i += blockDim.x) dst[i] = __ldg(&src[i]);
		__syncthreads();
	}
#endif
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (lane >= a.lanes) return;
	const int slot = warp * a.lanes + lane;
	const size_t hstride = (size_t)((gridDim.x * blockDim.x) >> 5) * a.lanes;
	const SimcuControl &k = a.ctl;
	const size_t M = a.M;

	Inst c(a, 0, slot, hstride);
	int hrows[SIMCU_DIM(SIMCU_NH)], prows[SIMCU_DIM(SIMCU_NPROBES)];
	genHistRows(hrows);
	genProbeRows(prows);
	double meas[SIMCU_DIM(SIMCU_NMEASURES)];

	double time = 0.0, prevTime = 0.0, thisStep = 0.0, maxStep = 0.0, lteStep = 0.0;
	int order = 1, brk = 0;
	int nfact = 0, nfactOp = 0, accepted = 0, rejLte = 0, rejNc = 0, pivotBreaks = 0, replayDiff = 0;
	long long attempts = 0;
	bool transient = false;      /* false: the operating-point solve */
	bool busy = false;           /* this lane holds an instance */
	bool exhausted = false;      /* this lane found the queue empty */
	const unsigned lanesMask = (a.lanes >= 32) ? 0xffffffffu : ((1u << a.lanes) - 1u);
#if SIMCU_REPLAY
	int rp = 0, rpEnd = 0, rflag = 0, nfact0 = 0, rl = 0;
#endif

	for (;;) {
#if SIMCU_BLOCKSYNC
		/* The warps of a block start every trip together, so they walk the same
		 * code at about the same time and share its instruction-cache lines (the
		 * loop body is several times the 32 KB L1.5 I-cache; unsynchronised, every
		 * warp streams it from L2 on its own).  A warp's trip already waits for
		 * its slowest lane, so waiting for the slowest warp costs little. */
#if SIMCU_SYNCGROUP > 0
		/* ... in groups of SIMCU_SYNCGROUP warps (named barrier + OR reduction) */
		{
			int any;
			const int vote = (busy || !exhausted) ? 1 : 0;
			const int grp = (threadIdx.x >> 5) / SIMCU_SYNCGROUP;
			const int first = grp * SIMCU_SYNCGROUP * 32;
			const int cnt = min((int)blockDim.x - first, SIMCU_SYNCGROUP * 32);
			asm volatile("{ .reg .pred p, q; setp.ne.s32 q, %1, 0; bar.red.or.pred p, %2, %3, q; selp.s32 %0, 1, 0, p; }"
					: "=r"(any) : "r"(vote), "r"(grp + 1), "r"(cnt) : "memory");
			if (!any) break;
		}
#else
		if (!__syncthreads_or((busy || !exhausted) ? 1 : 0)) break;
#endif
#endif
		/* ---- claim: an idle lane takes the next instance of the queue once
		 * few enough lanes of its warp are still busy (claimBelow = 31: at once;
		 * 0: the whole warp starts its next 32 instances together).  The ballots
		 * also re-converge the warp once per trip. ---- */
		unsigned busyMask = __ballot_sync(lanesMask, busy);
		if (!busy && !exhausted && __popc(busyMask) <= a.claimBelow) {
			const int q = atomicAdd(a.counter, 1);
			if (q >= a.count) exhausted = true;
			else {
				/* instances are processed grouped by table set (IBIS corner), so the
				 * lanes of a warp follow similar waveforms and diverge less */
				const int inst = a.order ? __ldg(&a.order[q]) : q;
				c.i = inst; c.hcount = 0; c.npts = 0; c.gridIdx = 0; c.time = 0.0; c.order = 1; c.err = 0;
				c.growth = 0;
#if SIMCU_WINDOWS
				c.wTstep = a.winst[inst]; c.wTstop = a.winst[M + inst];
				c.wTmax = a.winst[2 * M + inst]; c.wMinstep = a.winst[3 * M + inst];
#endif
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
				c.cbT0 = 0.0; c.cbT1 = 0.0; c.cbDt1 = 0.0; c.cbN = 0; c.cbAdv = false;
				c.growInv = false;
				SIMCU_UNROLL
				for (int j = 0; j < 4; j++) c.itT[j] = 0.0;
				SIMCU_UNROLL
				for (int j = 0; j < 3; j++) c.itH[j] = 0.0;
				SIMCU_UNROLL
				for (int q2 = 0; q2 < SIMCU_NMEASURES; q2++) meas[q2] = HUGE_VAL - HUGE_VAL;
				time = 0.0; prevTime = 0.0; thisStep = 0.0; maxStep = 0.0; lteStep = 0.0;
				order = 1; brk = 0;
				nfact = 0; nfactOp = 0; accepted = 0; rejLte = 0; rejNc = 0; pivotBreaks = 0; replayDiff = 0;
				attempts = 0;
				transient = false;
#if SIMCU_REPLAY
				rp = __ldg(&a.replayOff[inst]); rpEnd = __ldg(&a.replayOff[inst + 1]);
				rl = __ldg(&a.replayLinOff[inst]);
				c.linDiff = 0;
#endif
				genLoad(c);                                 /* simulator.c:121-123 */
				busy = true;
			}
		}
		busyMask = __ballot_sync(lanesMask, busy);
#if !SIMCU_BLOCKSYNC
		if (busyMask == 0) break;      /* every lane is idle and the queue is empty */
#endif
		if (!busy) continue;

		bool done = (c.err != 0);    /* instance over (finished or failed) */

		/* ---- attempt prologue (simulator.c:157-179) ---- */
		if (!done && transient) {
			if (attempts >= a.maxAttempts) { c.err = SIMCU_ERR_BUDGET; done = true; }
			else {
				attempts++;
#if SIMCU_REPLAY
				time = a.replayTime[rp]; rflag = a.replayFlag[rp]; rp++;
				thisStep = time - prevTime;
				nfact0 = nfact;
#else
				time = prevTime + thisStep;                 /* :157-161 */
				if (time > W_TSTOP(c)) { thisStep = W_TSTOP(c) - prevTime; time = W_TSTOP(c); }
#endif
				c.time = time; c.order = order;
				brk = genStep(c);                           /* :167-173 */
				if (brk) { order = 1; c.order = 1; }
#if SIMCU_REPLAY
				if (order != (rflag & 15)) replayDiff++;
				order = rflag & 15; c.order = order;
#endif
				if (!c.err) genIntegrate(c);                /* :178-179 */
				if (c.err) done = true;
			}
		}

		/* ---- simulatorSolve (simulator.c:45-69): solve, then {linearize; solve}
		 * until every device is linear or the limit is reached ---- */
		const int limit = transient ? k.itl4 : k.itl1;
		int count = 0;
		if (!done) {
			for (;;) {
				/* count == 0: the first solve of this loop; later ones skip the blocks
				 * no linearize stamp reaches (same stamps, same solution) */
				if (transient) { genFactorSolve1(c, count == 0); nfact++; }
				else { genFactorSolve0(c, count == 0); nfactOp++; }
				/* pivot breakdown in a transient attempt (the fixed sequence met a
				 * zero / non-finite pivot, or one that lost all its digits): reject
				 * the attempt as non-convergent, so the step is cut and Newton
				 * restarts - per instance, on device */
				if (transient && c.err == SIMCU_ERR_SINGULAR) {
					c.err = 0; count = limit; pivotBreaks++;
					break;
				}
				if (c.err) break;
				if (count++ >= limit) break;
#if SIMCU_REPLAY
				c.linMask = (rl < __ldg(&a.replayLinOff[c.i + 1])) ? a.replayLin[rl] : 0;
				c.linBit = 0;
				rl++;
#endif
				const int linear = genLinearize(c);
				if (c.err || linear) break;
			}
			if (c.err) done = true;
		}

		/* ---- attempt epilogue ---- */
		if (!done && !transient) {                      /* :126-137 */
			if (count > k.itl1) { c.err = SIMCU_ERR_OP; done = true; }
			else {
				gridOutput(c, prows, 0.0, 0.0, true);
				genMeasures(c, meas, 0.0, 0.0, true);
				genSaveX(c);
				if (opOnly || !(0.0 < W_TSTOP(c))) { recordStep(c, hrows, 0.0); done = true; }
				else {
					genInitStep(c);
					recordStep(c, hrows, 0.0);              /* :144 */
					maxStep = ((W_TSTEP(c) < W_TMAX(c)) ? W_TSTEP(c) : W_TMAX(c)) / 100;   /* :109 */
					thisStep = maxStep;
					genNextStep(c, thisStep);               /* :152-154 */
					order = k.maxorder;                     /* :133 */
					transient = true;
#if SIMCU_REPLAY
					if (rp >= rpEnd) done = true;
#endif
				}
			}
		} else if (!done) {
			/* own decision: 0 accept, 1 LTE reject, 2 non-convergence */
			int outcome = 2;
			if (count < k.itl4) {
				lteStep = W_TMAX(c);                         /* :188-190 */
				genMinStep(c, lteStep);
				outcome = (lteStep < (0.9 * thisStep)) ? 1 : 0;      /* :191 */
			}
#if SIMCU_REPLAY
			/* the log decides; count where this engine would have decided otherwise */
			if (outcome != ((rflag >> 4) & 15) || (nfact - nfact0) != (rflag >> 8)) replayDiff++;
			outcome = (rflag >> 4) & 15;
#endif
			if (c.err) done = true;
			else if (outcome == 1) {                        /* :191-198 */
				if (!SIMCU_REPLAY && lteStep < (W_TSTEP(c) * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; done = true; }
				else {
					order = 1;
					thisStep = lteStep;
					rejLte++;
				}
			} else if (outcome == 0) {                      /* accepted: :220-233 */
				gridOutput(c, prows, prevTime, time, false);
				genMeasures(c, meas, prevTime, time, false);
				genSaveX(c);
				order = (order < k.maxorder) ? order + 1 : order;
				prevTime = time;
				if (brk) maxStep = Min3(lteStep, 0.1 * maxStep, W_TMAX(c));
				else maxStep = Min3(lteStep, 2 * maxStep, W_TMAX(c));
				accepted++;
				recordStep(c, hrows, time);                 /* :144 of the next pass */
				if (!(time < W_TSTOP(c))) done = true;
				else {
					thisStep = maxStep;                     /* :152-154 */
					genNextStep(c, thisStep);
				}
			} else {                                        /* :201-214 */
				thisStep = thisStep / 8;
				if (!SIMCU_REPLAY && thisStep < (W_TSTEP(c) * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; done = true; }
				else {
					order = (order > 1) ? order - 1 : order;
					maxStep = maxStep / 8;
					rejNc++;
				}
			}
			if (!done && outcome != 0) {
				/* matrixRecall (:217): back to the last accepted solution - at the
				 * rows anything reads before the next solve overwrites X */
				genRecallX(c);
			}
#if SIMCU_REPLAY
			if (!done && rp >= rpEnd) done = true;
#endif
		}

		if (done) {
			jitFinish(a, c, meas, accepted, rejLte, rejNc, nfact, nfactOp, pivotBreaks, replayDiff);
			busy = false;
		}
	}
}

#endif
