/*
 * simcu_program.h - the compiled form of one netlist topology, shared by the
 * host builder (simcu_host.cpp) and the device kernels (simcu_kernels.cu).
 *
 * A "program" is topology only: device records, stamp slots, expression
 * bytecode, table data, the LU elimination schedule.  It is identical for all
 * instances and lives in read-only device memory.  Everything that differs per
 * instance (parameters, A, B, X, device state, history, outputs) is stored
 * structure-of-arrays with the instance index fastest: element `slot` of
 * instance `i` is at base[slot * M + i], so a warp touching one slot makes one
 * coalesced 256-byte access.
 */
#ifndef SIMCU_PROGRAM_H
#define SIMCU_PROGRAM_H

#define SIMCU_DEV_STRIDE 48   /* ints per device record */

/* device kinds, in the reference's vocabulary */
enum {
	DK_R = 0, DK_C, DK_L, DK_V, DK_I, DK_T, DK_VI, DK_BI, DK_BV, DK_BC
};

/* common record fields */
enum {
	DF_TYPE = 0, DF_K, DF_J, DF_L, DF_M, DF_ROWR, DF_ROWM, DF_SBASE, DF_IBASE,
	DF_X0 = 9            /* first type-specific field */
};
#define DF_ROWS DF_ROWM  /* T-line second branch row */

/* type-specific fields (offsets from 0) */
/* R C L */
#define DF_P0      9     /* parameter id of R / C / L */
#define DF_A0      10    /* first A slot */
/* V I */
#define DF_PDC     9
#define DF_WTYPE   10    /* 0 or 'p','g','s','e','l','c','f' */
#define DF_PARGS   11    /* first of 7 consecutive parameter ids */
#define DF_WTAB    12    /* table id for 'l'/'c' */
#define DF_SA0     13    /* V: A slots RK RJ KR JR */
#define DF_WSETSLOT 17   /* per-instance int slot holding the waveform table set, -1: set 0 */
/* T: parameter ids, 18 A slots, 6 history columns */
#define DF_PZ0     9
#define DF_PTD     10
#define DF_PLOSS   11
#define DF_TA0     12    /* RK RJ KR JR RL RM SL SM LS MS SK SJ RR SS KS JS LR MR */
#define DF_TH0     30    /* history column of R S K J L M (-1 = ground) */
#define DF_PBYPASS 36    /* parameter id of the per-instance bypass flag, -1: never bypassed */
/* VI */
#define DF_VITAB   9
#define DF_TATAB   10    /* -1: no TA table */
#define DF_NSETS   11
#define DF_SETSLOT 12    /* per-instance int slot holding the set index, -1: set 0 */
#define DF_VA0     13    /* RK RM KR MR JJ JM MJ MM */
/* B sources */
#define DF_CALC    9
#define DF_USETIME 10
#define DF_NBV     11
#define DF_BVOFF   12    /* offset into the B-variable table (4 ints per variable) */
#define DF_BA0     13    /* BI: RK RM KR MR | BV: RK RJ KR JR | BC: KK KJ JK JJ RC CR */

/* state layout (doubles, relative to DF_SBASE) */
#define INTEG_DOUBLES 22  /* y0, dydx0, ring[20] = t h x y f (4 each) */
#define CB_DOUBLES 9      /* x[3] t[3] theta[3] */
/* C / L : G Ieq integ          ints: n                */
/* V / I : dc cb pwdc           ints: cb.n pwIndex     */
/* T     : Vr Vs icK icJ icL icM cbR cbS att ints: cbR.n cbS.n hint */
/* VI    : G Ieq a lastV cb     ints: clState cb.n viIndex taIndex */
/* BI/BV : Ic In Beq lastV cb G[nbv]       ints: clState cb.n */
/* BC    : G Ieq integ Ic Cn Ceq lastV R[nbv]   ints: n clState */

/* expression bytecode (postfix) */
enum {
	OP_CONST = 0, OP_VAR, OP_NEG, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW,
	OP_FUNC, OP_GT, OP_LT, OP_GE, OP_LE, OP_EQ, OP_NE, OP_NOT, OP_AND, OP_OR,
	OP_IF
};
enum {
	FN_ABS = 0, FN_ACOSH, FN_ACOS, FN_ASINH, FN_ASIN, FN_ATANH, FN_ATAN,
	FN_COSH, FN_COS, FN_EXP, FN_LN, FN_LOG, FN_SINH, FN_SIN, FN_SQRT, FN_TAN,
	FN_URAMP, FN_U
};
#define CALC_MAX_STACK 24
#define CALC_MAX_VARS 16
#define CALC_VAR_TIME (-2)   /* variable reference meaning "time" */

/* device-side measures on a probe (simcuSetMeasures) */
enum { MS_MIN = 0, MS_MAX, MS_FINAL, MS_TCROSS_RISE, MS_TCROSS_FALL };

/* per-instance control scalars (doubles) */
enum {
	CD_TIME = 0, CD_PREVTIME, CD_THISSTEP, CD_MAXSTEP, CD_LTESTEP, CD_NUM
};
/* per-instance control scalars (ints) */
enum {
	CI_ORDER = 0, CI_STATUS, CI_DONE, CI_ACCEPTED, CI_REJ_LTE, CI_REJ_NC,
	CI_NFACT, CI_NFACT_OP, CI_NPTS, CI_HCOUNT, CI_GRIDIDX, CI_ATTEMPTS,
	CI_BREAK, CI_REPLAY_DIFF, CI_GROWTH, CI_LIN_DIFF, CI_NUM
};

/* options of the loop: reference include/control.h:34-56 */
typedef struct {
	int itl1, itl4, maxorder;
	double reltol, vntol, abstol, captol, chgtol, trtol, gmin;
	double maxAngleA, maxAngleV;
	double tstep, tstop, tmax, minstep;
} SimcuControl;

/* everything a kernel needs; passed by value */
typedef struct {
	int M;                    /* instances in this batch */
	int N, nnzA, F;           /* unknowns, nnz(A), workspace matrix slots */
	int ndev;
	SimcuControl ctl;
	/* program (read-only) */
	const int *dev;           /* [ndev][SIMCU_DEV_STRIDE] */
	const int *bvar;          /* [*][4]: row, slotA, slotB, calc variable */
	const int *lu;            /* elimination schedule (see simcu_host.cpp) */
	int luLen;
	const int *xslot;         /* [N] workspace slot holding x of unknown k */
	const int *calcOff;       /* [ncalc+1] offsets into calcCode */
	const int *calcCode;
	const double *calcConst;
	const int *calcVarOff;    /* [ncalc+1] offsets into calcVars */
	const int *calcVars;      /* row index (0 = ground), CALC_VAR_TIME */
	const double *tabP;       /* packed tables: [point][4] = x, y, m, x_next */
	const int *tabDesc;       /* [ntab][4]: first point, points, type 'l'/'c', first point of the
	                           * copy staged in shared memory (-1: not staged) */
	const double *tabStage;   /* the staged tables, contiguous, in shared-memory order */
	int stagePoints;          /* points staged (32 B each) */
	const int *histRows;      /* [nh] row index (1-based) recorded per step */
	int nh, histRecords;
	const int *probeRows;     /* [nprobes] row index (1-based) */
	int nprobes, ngrid;
	int rawMaxPoints;         /* > 0: keep raw (time, X) records ... */
	int rawFirst, rawCount;   /* ... of instances [rawFirst, rawFirst+rawCount) */
	/* parameters */
	const int *pmap;          /* [nparams] >=0: per-instance slot, <0: broadcast */
	const double *pconst;     /* [nparams] broadcast values */
	const double *pinst;      /* [ninst][M] */
	const int *iinst;         /* [niinst][M] per-instance ints (table set) */
	/* per-instance state, all [slot][M] */
	double *A, *B, *X, *Xrec, *S, *CD;
	int *IS, *CI;
	double *Ht, *Hx;          /* T-line history ring: times Ht[thread][R], values Hx[thread][R][nh] */
	double *outGrid;          /* [ngrid][nprobes][M] */
	double *outRaw;           /* [rawMaxPoints][N+1][rawCount] */
	double *outMeasure;       /* [nmeasures][M]: device-side post-processing */
	const double *measureLevel; /* [nmeasures] threshold of the crossing measures */
	double *wsGlobal;         /* [(F+N)][M] when the workspace is not in smem */
	int wsShared;
	int attemptsPerLaunch;
	long long maxAttempts;
	int *remaining;           /* device counter of unfinished instances */
	int stopBeforeSolve;      /* pilot: assemble the first attempt, then stop */
	int *counter;             /* specialised kernel: next unclaimed instance */
	int lanes;                /* active lanes per warp (power of two <= 32) */
	const int *order;         /* processing order of the instances (NULL = as given) */
	int count;                /* instances this launch processes (<= M; M stays the stride) */
	int claimBelow;           /* an idle lane claims once <= this many lanes of its warp are busy */
	/* replay of an attempt log (test instrumentation): entries
	 * [replayOff[i], replayOff[i+1]) of instance i; flag = order | outcome << 4 | solves << 8 */
	const int *replayOff;
	const double *replayTime;
	const int *replayFlag;
	/* ... and of the per-device convergence-test outcomes: one word per linearize
	 * pass, entries [replayLinOff[i], replayLinOff[i+1]) of instance i */
	const int *replayLinOff;
	const int *replayLin;
	/* per-instance analysis windows [tstep, tstop, tmax, minstep][M] (NULL: ctl's) */
	const double *winst;
} SimcuArgs;

#endif
