/*
 * eis_oracle.h - CPU restatement of the eispice transient hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle for the B200 batched
 * transient engine: a plain-C, single-instance restatement of the algorithm in
 * the reference's libs/simulator (Newton loop, device stamps, step control),
 * libs/calculon (B-source expressions) and the numeric part of libs/superlu
 * (partial-pivoting LU, restated densely).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load it; the product never does.
 *
 * Parity status: PINNED - checked against the unmodified reference compiled
 * here (oracle/_ref, see build_ref.sh) on every golden circuit of SURVEY.md 8c
 * and against the doctest known-answer values (tests/test_oracle_*.py).
 *
 * Each function in eis_oracle.c cites the reference file:line it follows
 * (paths relative to the reference root).
 */
#ifndef EIS_ORACLE_H
#define EIS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eo_sim eo_sim;

/* per-run counters (not in the reference; used to report k = solves/step) */
typedef struct {
	long accepted;      /* accepted timesteps (rows of result - 1) */
	long rejected_lte;  /* attempts rejected by the LTE test */
	long rejected_nc;   /* attempts rejected for non-convergence */
	long factorizations;/* matrixSolve + matrixSolveAgain calls in transient */
	long op_factorizations;
} eo_stats;

eo_sim *eoNew(void);
void eoDestroy(eo_sim *s);

/* mirrors of libs/simulator/include/simulator.h:52-147, parameters by value */
int eoAddResistor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double R);
int eoAddCapacitor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double C);
int eoAddInductor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double L);
/* type 'v'/'i'; stimulus 0,'p','s','e','f','g' take args[7] (HUGE_VAL = unset);
 * 'l'/'c' take table[len][2] */
int eoAddSource(eo_sim *s, const char *refdes, const char *p, const char *n,
		char type, double dc, char stimulus, const double *args,
		const double *table, int tableLen);
int eoAddNonlinearSource(eo_sim *s, const char *refdes, const char *p,
		const char *n, char type, const char *equation);
int eoAddTLine(eo_sim *s, const char *refdes, const char *pl, const char *nl,
		const char *pr, const char *nr, double Z0, double Td, double loss);
int eoAddVICurve(eo_sim *s, const char *refdes, const char *p, const char *n,
		const double *vi, int viLen, char viType,
		const double *ta, int taLen, char taType);

/* change a scalar parameter between runs (the reference does this through the
 * pointer it holds into the Python object).  name is one of
 * R C L DC Z0 Td loss, or A0..A6 for waveform arguments. */
int eoSetParam(eo_sim *s, const char *refdes, const char *name, double v);

/* 0 = partial pivoting every factorisation (reference behaviour,
 * libs/superlu/SRC/dpivotL.c:99-147); 1 = replay a fixed pivot sequence:
 * step k eliminates column pc[k] with pivot row pr[k] (0-based unknowns).
 * phase 0 = operating point, 1 = transient. */
int eoSetPivots(eo_sim *s, int phase, const int *pc, const int *pr, int n);

/* (column, row) of every elimination step of the FIRST factorisation of a phase
 * since the simulator was created: the sequence a test hands to the engine
 * (simcuSetPivots) and back to eoSetPivots so both sides eliminate alike.
 * Returns N, or -1 if that phase has not been solved yet. */
int eoGetPivots(eo_sim *s, int phase, int *pc, int *pr);

/* Attempt log of the last transient (test instrumentation, see eis_oracle.c):
 * one entry per attempt in order - the time tried and a flag word
 * order | outcome << 4 | solves << 8, outcome 0 accepted / 1 LTE reject /
 * 2 non-convergence.  The arrays stay owned by the simulator. */
#define EO_LOG_FLAG(order, outcome, solves) ((order) | ((outcome) << 4) | ((solves) << 8))
void eoSetAttemptLog(eo_sim *s, int on);
int eoGetAttemptLog(eo_sim *s, const double **times, const int **flags);
/* one word per linearize pass of the last transient (operating point first):
 * bit k = both tolerance tests of the k-th nonlinear device (VI / B sources in
 * netlist order) passed (math/checklinear.c:62-77) - the threshold decisions at
 * reltol that a 1-ulp difference can flip.  At most 31 nonlinear devices. */
int eoGetLinearizeLog(eo_sim *s, const int **masks);

/* Stamp watch (test instrumentation): with on = 1 the runs record which A
 * entries (dense, a[row * N + col], 0-based unknowns) and which right-hand-side
 * rows differed between two consecutive solves of ONE simulatorSolve loop
 * (core/simulator.c:45-69), per phase (0 operating point, 1 transient).
 * Returns N (0: nothing recorded yet).  The engine's claim "only linearize
 * stamps change inside a Newton loop" is checked against this. */
void eoSetStampWatch(eo_sim *s, int on);
int eoGetStampWatch(eo_sim *s, int phase, const char **a, const char **b);

int eoRunTransient(eo_sim *s, double tstep, double tstop, double tmax,
		int restart, double **data, char ***variables, int *numPoints,
		int *numVariables);
int eoRunOperatingPoint(eo_sim *s, double **data, char ***variables,
		int *numPoints, int *numVariables);
void eoFree(void *p);

int eoNumUnknowns(eo_sim *s);
int eoNumNonzeros(eo_sim *s);
/* CSC pattern in the reference's order (node list sorted by col,row) */
int eoGetPattern(eo_sim *s, int *rowIdx, int *colStart);
void eoGetStats(eo_sim *s, eo_stats *st);
void eoQuiet(int q);

/* expression engine alone (differential fuzzing against calculon) */
int eoCalcEval(const char *expr, int nvars, const char **names,
		const double *values, double minDiv, double *value, double *derivs);

#ifdef __cplusplus
}
#endif
#endif
