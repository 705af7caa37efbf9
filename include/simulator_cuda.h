/*
 * simulator_cuda.h - C-ABI of libs/simulator_cuda: the B200 batched transient
 * engine behind eispice's simulator API.
 *
 * This is the drop-in boundary of SURVEY.md section 8b.  The builder calls
 * keep the names-by-position, argument meaning, pointer-borrowing and error
 * convention of the reference's libs/simulator/include/simulator.h (cited per
 * function), so that module/simulatormodule.c can call them with the very
 * pointers it passes to simulatorAdd* today.  What is new is the instance
 * dimension: one netlist topology, numInstances independent parameter sets.
 *
 * Conventions (reference libs/log/log.h:217-268): int 0 = ok, -1 = error,
 * pointer NULL = error; a message goes to stderr.  Strings are copied; every
 * double* / table pointer is BORROWED and re-read at the start of each run
 * (reference: parameters are re-read at every deviceLoad).
 * Not thread-safe per handle; one handle per Circuit.
 *
 * No torch types, no CUDA types: plain pointers and sizes.  `stream` arguments
 * are a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef SIMULATOR_CUDA_H
#define SIMULATOR_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define SIMCU_MAJOR_VERSION 0
#define SIMCU_MINOR_VERSION 1

typedef struct _simcu simcu_;

/* per-instance outcome (reference: any of these aborts the whole run,
 * core/simulator.c:193-195,208-210, core/matrix.c:108-115; here the batch
 * continues and the instance is flagged) */
enum {
	SIMCU_OK = 0,
	SIMCU_ERR_TIMESTEP = 1,    /* "Timestep too small" */
	SIMCU_ERR_SINGULAR = 2,    /* zero / non-finite pivot ("U is singular") */
	SIMCU_ERR_NAN = 3,         /* NaN in a device evaluation */
	SIMCU_ERR_OP = 4,          /* operating point did not converge (itl1) */
	SIMCU_ERR_HISTORY = 5,     /* T-line history ring too short */
	SIMCU_ERR_BUDGET = 6       /* attempt budget exhausted */
};

typedef struct {
	long long accepted;        /* accepted timesteps, all instances */
	long long rejectedLTE;     /* attempts rejected by the LTE test */
	long long rejectedNC;      /* attempts rejected for non-convergence */
	long long factorizations;  /* numeric LU + solve count (transient) */
	long long opFactorizations;
	long long failed;          /* instances with status != SIMCU_OK */
	int numUnknowns;           /* N */
	int numNonzeros;           /* nnz(A) */
	int numFactor;             /* nnz(L+U-I) of the transient schedule */
	int numFactorOP;           /* same for the operating-point schedule */
	int stateDoubles;          /* per-instance device state (doubles) */
	int histRows;              /* values per T-line history record */
	int histRecords;           /* ring length */
	int workspaceInShared;     /* 1: LU workspace is in shared memory */
	int threadsPerBlock;
	int kernelLaunches;        /* launches of the last simcuRunDevice */
	double deviceMs;           /* device time of the last simcuRunDevice */
	double tranMs;             /* of which: the transient kernel launches */
	int jitCompiles;           /* NVRTC compilations this handle caused */
	int registersPerThread;    /* of the topology-specialised kernel */
	int localBytesPerThread;
	int gridBlocks;
	int lanesPerWarp;          /* instances per warp of the specialised kernel */
	long long pivotBreakdowns; /* transient attempts rejected for a zero pivot */
	int historyRetries;        /* instances re-run with a deeper T-line history ring */
	long long replayDiffs;     /* replay mode: attempts this engine would have decided otherwise */
	long long growthWarnings;  /* factorisations whose largest LU multiplier exceeded growth_limit:
	                            * the fixed pivot sequence (chosen on pilot instances) lost that many
	                            * digits on this instance; per instance: counter "pivot_growth" */
	double jitCompileMs;       /* wall time this handle spent compiling (NVRTC, or reading the cubin cache) */
} simcuStats_;

/* simulatorNew (simulator.h:33): r must be NULL */
simcu_ *simcuNew(simcu_ *r, int numInstances);
/* simulatorDestroy (simulator.h:31) */
int simcuDestroy(simcu_ **r);
int simcuInfo(void);
/* CUDA device for this thread's subsequent calls (one process per GPU) */
int simcuSetDevice(int device);

/* ---- builders: simulator.h:52-147, same argument lists ------------------- */
/* simulatorAddResistor (simulator.h:52-55) */
int simcuAddResistor(simcu_ *r, char *refdes, char *pNode, char *nNode,
		double *resistance);
/* simulatorAddCapacitor (simulator.h:57-60) */
int simcuAddCapacitor(simcu_ *r, char *refdes, char *pNode, char *nNode,
		double *capacitance);
/* simulatorAddInductor (simulator.h:62-65) */
int simcuAddInductor(simcu_ *r, char *refdes, char *pNode, char *nNode,
		double *inductance);
/* simulatorAddNonlinearSource (simulator.h:67-70): type 'i' 'v' 'c' */
int simcuAddNonlinearSource(simcu_ *r, char *refdes, char *pNode, char *nNode,
		char type, char *equation);
/* simulatorAddSource (simulator.h:72-75): type 'v'/'i'; stimulus 0 'p' 's' 'e'
 * 'l' 'c' 'f' 'g'; for 'l'/'c' args[0] is a double** table and args[1] an int*
 * length (math/waveform.c:143-150) */
int simcuAddSource(simcu_ *r, char *refdes, char *pNode, char *nNode,
		char type, double *dc, char stimulus, double *args[7]);
/* simulatorAddTLine (simulator.h:100-104) */
int simcuAddTLine(simcu_ *r, char *refdes, char *pNodeLeft, char *nNodeLeft,
		char *pNodeRight, char *nNodeRight, double *Z0, double *Td,
		double *loss);
/* simulatorAddVICurve (simulator.h:126-130) */
int simcuAddVICurve(simcu_ *r, char *refdes, char *pNode, char *nNode,
		double **vi, int *viLength, char viType,
		double **ta, int *taLength, char taType);
/* Not batchable (call back into the interpreter per Newton iteration /
 * unfinished upstream): always -1.  simulator.h:77-98,106-124 */
int simcuAddCallbackSource(simcu_ *r, char *refdes, char *pNode, char *nNode,
		char type, char *variables[], double values[], double derivs[],
		int numVariables, void *callback, void *private_);
int simcuAddTLineW(simcu_ *r, char *refdes, char *nodes[], int nNodes, int *M,
		double *len, double **L0, double **C0, double **R0, double **G0,
		double **Rs, double **Gd, double *fgd, double *fK);

/* ---- the instance dimension (additive) ----------------------------------- */
/* Per-instance values of a scalar parameter: name is the binding's member name
 * (R C L DC Z0 Td loss, or A0..A6 = waveform argument slot).  values[i*stride]
 * is instance i; stride 0 broadcasts.  Copied immediately.
 * T-lines also take "bypass" (0 / 1): a per-instance TOPOLOGY flag - where it is
 * 1 the section is an ideal through-connection (V_K - V_J = V_L - V_M, the same
 * current through both ports), at the operating point and in transient, so a
 * sweep over the number of sections of a line runs as one batch of the maximum
 * topology. */
int simcuSetInstanceParam(simcu_ *r, char *refdes, char *name,
		const double *values, int stride);
/* Further table sets (e.g. IBIS corners) for an existing VI device; returns
 * the set index (>= 1) or -1.  Tables are borrowed like simcuAddVICurve's. */
int simcuAddVITableSet(simcu_ *r, char *refdes, double **vi, int *viLength,
		double **ta, int *taLength);
/* Further waveform tables for an existing PWL / PWC source (stimulus 'l' / 'c'
 * of simcuAddSource); returns the set index (>= 1) or -1.  Borrowed likewise. */
int simcuAddSourceTableSet(simcu_ *r, char *refdes, double **table, int *length);
/* Which table set each instance uses (default 0), for a VI device or a PWL /
 * PWC source. Copied immediately. */
int simcuSetInstanceTableSet(simcu_ *r, char *refdes, const int *setIndex,
		int stride);
/* Per-instance analysis windows: instance i runs tran(tstep[i*stride],
 * tstop[i*stride], tmax[i*stride]) (tmax = NULL or 0: the reference's default,
 * core/simulator.c:91-94); the tstep / tstop of the run call then only define
 * the output grid.  For batches of DIFFERENT analyses of one topology - the six
 * (direction, speed) transients of an IBIS model's K tables
 * (module/ibis_parser.py:500-510).  tstep = NULL returns to one window. */
int simcuSetInstanceWindow(simcu_ *r, const double *tstep, const double *tstop, const double *tmax,
		int stride);
/* Options: reltol vntol abstol captol chgtol trtol gmin itl1 itl4 maxorder
 * minstep (control.h:34-56); hist_records, max_attempts, attempts_per_launch,
 * threads_per_block, workspace_shared (0/1), raw_max_points, raw_first,
 * raw_count, pivot_threshold, num_pilots, min_blocks_per_sm (of the specialised
 * kernel: bounds its registers, 0 = compiler's choice), lanes_per_warp (0 = automatic), keep_order (1: do not group instances by
 * table set for processing), history_retry (1: re-run instances that
 * outran the T-line history ring, with a deeper ring and 32 x the attempts of
 * an average instance), interpreter (1: run the batch on the generic
 * table-driven kernels instead of the NVRTC-specialised one; for verification),
 * claim_below (0..31: an idle lane of the specialised kernel claims its next
 * instance once at most this many lanes of its warp are still busy; 31 = at
 * once, 0 = whole warps start together), block_sync (1: the warps of a block
 * start every attempt together and share instruction-cache lines),
 * growth_limit (LU multiplier magnitude counted as pivot growth, default 1e6),
 * smem_tables_kb (0..200: every block stages that many kB of tables in shared
 * memory, V-I tables first; big-block launch shapes only; -1 = automatic: 100 kB
 * for the 256-thread shape, none for the 896-thread shape),
 * tline_prefetch (1: extra locate + L2 prefetch pass over the T-lines; measured
 * slower, off), fmad (default 0: the kernel is compiled WITHOUT a*b+c
 * contraction, i.e. with the reference's arithmetic - under the same pivot
 * sequence it then reproduces the reference algorithm bit for bit wherever no
 * libm function is involved; 1 allows contraction: 0-3 % faster, results differ
 * in the last bits and step sequences may differ),
 * solve_blocks (generated LU of the specialised kernel: 0 = all at once, 1 =
 * regrouped by independent matrix block - T-lines decouple their ports - so only
 * one block is live at a time, 2 = default: additionally the blocks no linearize
 * stamp reaches are solved only in the first solve of a Newton loop; results
 * are bit-identical in all three), shared_break_clock (default 1: one
 * break-point clock per instance instead of one per device, same decisions),
 * replay (see simcuSetReplayLog);
 * r = NULL: quiet (0/1) */
int simcuSetOption(simcu_ *r, char *name, double value);
/* Force the pivot sequence of a phase (0 = operating point, 1 = transient)
 * instead of the built-in pilot analysis: step k eliminates column pc[k] with
 * pivot row pr[k] (0-based unknowns).  This is where a build inside eispice
 * feeds perm_c / perm_r of its own libs/superlu (dgssvx, core/matrix.c:82-118;
 * SURVEY.md appendix B).  pc = NULL returns to the built-in analysis. */
int simcuSetPivots(simcu_ *r, int phase, const int *pc, const int *pr, int n);
/* Test instrumentation (option replay = 1): instead of its own accept / reject
 * decisions (core/simulator.c:186-214) instance i replays the attempts
 * [offsets[i], offsets[i+1]) of a log: times[] = the time each attempt tried,
 * flags[] = integration order | outcome << 4 | LU solves << 8, outcome 0
 * accepted, 1 rejected by the LTE test, 2 rejected for non-convergence.  The
 * log comes from the CPU oracle (oracle/eis_oracle.h: eoGetAttemptLog), so that
 * both sides take identical steps and every unknown of every accepted step can
 * be held to 1e-6 / 1e-9.  linMasks[linOffsets[i] .. linOffsets[i+1]) holds one
 * word per linearize pass of instance i (operating point first): bit k = the
 * two convergence tests of its k-th nonlinear device passed
 * (math/checklinear.c:62-77) - those threshold decisions are replayed too, and
 * the engine's own are counted where they differ (counter linearize_diff).
 * Copied immediately; offsets = NULL clears it. */
int simcuSetReplayLog(simcu_ *r, const int *offsets, const double *times, const int *flags,
		const int *linOffsets, const int *linMasks);

/* ---- runs ----------------------------------------------------------------- */
/* simulatorRunTransient (simulator.h:35-38) for ONE instance (index
 * `instance`), same result layout: *data is malloc'd [numPoints][numVariables]
 * row-major with column 0 = time, *variables is a malloc'd NULL-terminated
 * array (strings owned by the handle). Free both with simcuFree.
 * restart: the reference continues a finished transient when restart = 0
 * (core/simulator.c:111,138-140).  This engine keeps no per-instance state
 * between launches: restart = 0 after a completed transient returns -1 with a
 * message; on a handle that has not run a transient it is a normal start. */
int simcuRunTransient(simcu_ *r, double tstep, double tstop, double tmax,
		int restart, int instance, double *data[], char **variables[],
		int *numPoints, int *numVariables);
/* simulatorRunOperatingPoint (simulator.h:40-42) for one instance */
int simcuRunOperatingPoint(simcu_ *r, int instance, double *data[],
		char **variables[], int *numPoints, int *numVariables);

/* Operating point of every instance: *data is malloc'd
 * [numInstances][numProbes], *status (optional) malloc'd [numInstances]. */
int simcuRunOperatingPointBatch(simcu_ *r, char *probes[], int numProbes,
		double *data[], int *status[]);

/* Batched transient, HOST buffers in and out (the e2e path): uploads the
 * per-instance parameters, runs every instance with its own adaptive step
 * control, resamples the probes onto the grid t_g = g*tstep (g = 0..numGrid-1,
 * last point tstop) by linear interpolation and downloads the result.
 *   probes[numProbes]  result names, "v(node)" / "i(refdes)"
 *   *data              malloc'd [numInstances][numGrid][numProbes]
 *   *status            malloc'd [numInstances] (SIMCU_*)
 * stats may be NULL. */
int simcuRunTransientBatch(simcu_ *r, double tstep, double tstop, double tmax,
		int restart, char *probes[], int numProbes, double *data[],
		int *numGrid, int *status[], simcuStats_ *stats);

/* The same run in stages, for callers that keep data on the device:
 *   Prepare  one-off: compile the netlist, symbolic LU analysis on a pilot
 *            batch, allocate device state, upload tables
 *   Upload   host -> device copy of the per-instance parameters
 *   RunDevice  all kernels of one batched transient on `stream`
 *   Fetch    device -> host copy of the gridded probes ([M][numGrid][numProbes])
 *            and status; either pointer may be NULL */
int simcuPrepare(simcu_ *r, double tstep, double tstop, double tmax,
		char *probes[], int numProbes);
int simcuUpload(simcu_ *r, void *stream);
int simcuRunDevice(simcu_ *r, void *stream);
int simcuFetch(simcu_ *r, double *data, int *status, void *stream);
int simcuNumGrid(simcu_ *r);
/* device pointer of the gridded output, layout [numGrid][numProbes][M] */
void *simcuDeviceOutput(simcu_ *r);
int simcuGetStats(simcu_ *r, simcuStats_ *stats);
/* Transposes the gridded output of the last batch to the host layout
 * [numInstances][numGrid][numProbes] ON THE DEVICE (no copy out) and returns
 * device pointers to it and to the status array, for callers that move the
 * data on themselves (the NCCL gather below).  Any output may be NULL. */
int simcuTransposeOutput(simcu_ *r, void *stream, void **deviceData, void **deviceStatus,
		int *numInstances, int *valuesPerInstance);

/* ---- multi-GPU: the one collective ---------------------------
 * One process per GPU; instances are independent, so nothing is exchanged
 * while stepping.  A sweep of nranks * M instances is dealt out round-robin:
 * global instance g runs on rank g % nranks as its local instance g / nranks
 * (every rank holds the same mix of corners).  simcuGatherOutput moves the
 * finished waveforms of all ranks to `root` over NCCL (grouped ncclSend /
 * ncclRecv between device buffers, NVLink / NVSwitch) and copies them to the
 * root's host buffers once, in global instance order:
 *   hostData    root only: [nranks * M][numGrid][numProbes]
 *   hostStatus  root only: [nranks * M]
 * Every rank must hold the same M, grid and probes.  The communicator is
 * bootstrapped by the caller: rank 0 calls simcuCommUniqueId and hands the 128
 * bytes to the other ranks by any means (torch.distributed broadcast, a file);
 * then every rank calls simcuCommInit on its own GPU. */
typedef struct _simcuComm simcuComm_;
typedef struct {
	double ncclMs;             /* device time of the grouped send / recv */
	double copyMs;             /* root: device -> host copies */
	long long ncclBytes;       /* bytes this rank sent (root: received) over NCCL */
	long long copyBytes;       /* bytes copied to the host (root) */
} simcuGatherStats_;
int simcuCommVersion(void);                       /* NCCL version code, -1 without NCCL */
int simcuCommUniqueId(char *id, int bytes);       /* bytes >= 128 */
int simcuCommInit(simcuComm_ **c, int nranks, int rank, const char *id, int bytes);
int simcuCommDestroy(simcuComm_ **c);
int simcuGatherOutput(simcu_ *r, simcuComm_ *c, int root, double *hostData, int *hostStatus,
		void *stream, simcuGatherStats_ *stats);
/* The same in two halves, for callers that pipeline batches: Async transposes on
 * `stream` (the stream of simcuRunDevice) and queues the NCCL transfer and the
 * copies to the host on the communicator's own stream, then returns - the next
 * simcuRunDevice can start while the waveforms of this one are still travelling
 * (they only read buffers the next batch does not touch until ITS gather).
 * Wait blocks until the host buffers are complete.  One gather in flight per
 * communicator; host buffers must stay valid (and should be pinned) until Wait. */
int simcuGatherOutputAsync(simcu_ *r, simcuComm_ *c, int root, double *hostData, int *hostStatus,
		void *stream);
int simcuGatherWait(simcuComm_ *c, simcuGatherStats_ *stats);

/* Device-side post-processing (eispice NOTES:40 "Add post-processors"): n
 * running measures, each on one unknown probes[i] ("v(node)" / "i(refdes)"),
 * kinds[i] = 0 minimum, 1 maximum, 2 final value, 3 / 4 time of the first
 * rising / falling crossing of levels[i] (linear interpolation inside the
 * accepted step, NaN if it never crosses).  Evaluated over the accepted steps
 * inside the transient kernel; simcuFetchMeasures copies [numInstances][n] out.
 * n = 0 switches them off. */
int simcuSetMeasures(simcu_ *r, char *probes[], const int *kinds, const double *levels, int n);
int simcuFetchMeasures(simcu_ *r, double *out);
/* per-instance counters of the last batch into out[numInstances]; which is
 * status, accepted, rejected_lte, rejected_nc, factorizations,
 * op_factorizations, pivot_breakdowns, replay_diff, linearize_diff or
 * pivot_growth */
int simcuFetchCounters(simcu_ *r, char *which, int *out);
/* raw (time, X) record of one instance of the last batch, reference layout
 * (core/matrix.c:430-470): malloc'd [numPoints][numVariables], column 0 = time.
 * Needs the options raw_max_points > 0 and raw_first / raw_count covering it. */
int simcuFetchRaw(simcu_ *r, int instance, double *data[], int *numPoints,
		int *numVariables);
/* all raw records of the last batch in one copy: out is
 * [raw_max_points][numVariables][raw_count] (instance fastest), numPoints
 * [raw_count] the points each instance produced (may exceed raw_max_points:
 * then its record is truncated) */
int simcuFetchRawBlock(simcu_ *r, double *out, int *numPoints);

/* ---- introspection (host only, no GPU needed) ----------------------------- */
/* Environment: SIMCU_NVRTC / SIMCU_NCCL = library paths; SIMCU_JIT_DUMP = a
 * directory that receives every generated translation unit; SIMCU_CACHE_DIR = a
 * directory used as a cubin cache across processes (NVRTC takes 1-7 s per
 * topology; off when unset).
 * path and version of the NVRTC library the kernels were compiled with ("" before
 * the first compile).  The toolkit's own copy is preferred over one a host
 * application (PyTorch) has already loaded; override with SIMCU_NVRTC. */
const char *simcuNvrtcPath(void);
int simcuNumUnknowns(simcu_ *r);
int simcuNumNonzeros(simcu_ *r);
/* CSC pattern in the reference's order (core/node.c:56-66, matrix.c:476-497) */
int simcuGetPattern(simcu_ *r, int *rowIdx, int *colStart);
/* name of unknown k (0-based), e.g. "v(out)"; owned by the handle */
const char *simcuVariableName(simcu_ *r, int k);
/* fixed pivot sequence chosen by Prepare: phase 0 = operating point,
 * 1 = transient; step k eliminates column pc[k] with pivot row pr[k] */
int simcuGetPivots(simcu_ *r, int phase, int *pc, int *pr);
/* test hook (no GPU needed): flatten, structural symbolic analysis, source
 * generation and NVRTC compilation of the topology-specialised kernel.
 * *source (optional) receives the malloc'd translation unit; cubinBytes = NULL
 * stops after the source generation (no NVRTC). */
int simcuCompileOnly(simcu_ *r, char *probes[], int numProbes, char **source,
		int *cubinBytes);
/* test hook (no GPU needed): a malloc'd copy (simcuFree) of one of the host-side
 * arrays a launch of the specialised kernel reads, by name: "pmap", "iinst",
 * "tabDesc" (int), "pconst", "pinst", "tabP" (double), "control" (the
 * SimcuControl of csrc/simcu_program.h with the run constants of
 * core/simulator.c:91-109 resolved for tstep / tstop / tmax).  tests/ runs the
 * generated translation unit compiled for the host on them against the oracle. */
int simcuDebugProgram(simcu_ *r, const char *what, double tstep, double tstop, double tmax,
		void **data, int *bytes);
/* test hook (no GPU needed): one table ([length][2]: x, y; type 'l' linear, 'c'
 * natural cubic spline - math/piecewise.c:108-201) in the packed form the
 * kernels read, packed[length][4] = x, y, slope / second derivative, next x */
int simcuPackTable(const double *xy, int length, char type, double *packed);
/* test hook: the symbolic analysis alone.  CSC pattern + numPilots value sets
 * ([numPilots][nnz]) in, pivot sequence, workspace slots and nnz(L+U-I) out */
int simcuSymbolic(int n, const int *colStart, const int *rowIdx, int numPilots,
		const double *values, double threshold, int *pc, int *pr, int *numSlots,
		int *numFactor);
/* test hook: compile `expr` to device bytecode and evaluate it with the
 * device interpreter on `where` = 0 host build of the same code / 1 GPU.
 * Same contract as the reference's calcSolve + calcDiff per variable. */
int simcuCalcEval(const char *expr, int nvars, const char **names,
		const double *values, double minDiv, int where, double *value,
		double *derivs);

void simcuFree(void *p);

#ifdef __cplusplus
}
#endif
#endif
