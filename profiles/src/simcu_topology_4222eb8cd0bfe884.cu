#define SIMCU_JIT 1
#define SIMCU_N 48
#define SIMCU_NNZ 141
#define SIMCU_SD 442
#define SIMCU_SI 56
#define SIMCU_NPARAM 36
#define SIMCU_NISET 5
#define SIMCU_NH 25
#define SIMCU_NPROBES 2
#define SIMCU_TPB 896
#define SIMCU_MINB 1
#define SIMCU_NMEASURES 0
#define SIMCU_REPLAY 0
#define SIMCU_BLOCKSYNC 1
#define SIMCU_FMAD 1   /* compiled with --fmad=true */
#define HUGE_VAL (__longlong_as_double(0x7ff0000000000000LL))
#define SIMCU_GMIN (__longlong_as_double(0x3cd203af9ee75616LL))
#define SIMCU_LMAX (__longlong_as_double(0x412e848000000000LL))
#define M_PI 3.14159265358979323846
enum { SIMCU_OK = 0, SIMCU_ERR_TIMESTEP = 1, SIMCU_ERR_SINGULAR = 2, SIMCU_ERR_NAN = 3,
	SIMCU_ERR_OP = 4, SIMCU_ERR_HISTORY = 5, SIMCU_ERR_BUDGET = 6 };

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
/* T: parameter ids, 18 A slots, 6 history columns */
#define DF_PZ0     9
#define DF_PTD     10
#define DF_PLOSS   11
#define DF_TA0     12    /* RK RJ KR JR RL RM SL SM LS MS SK SJ RR SS KS JS LR MR */
#define DF_TH0     30    /* history column of R S K J L M (-1 = ground) */
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
	const int *tabOff;        /* [ntab+1] */
	const int *tabType;       /* 'l' / 'c' */
	const double *tabP;       /* packed tables: [point][4] = x, y, m, x_next */
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
} SimcuArgs;

#endif
/*
 * simcu_calc.h - interpreter of the B-source expression bytecode.
 *
 * The reference re-runs an LALR(1) parser over the token list for every
 * evaluation, and once more per variable for each partial derivative
 * (libs/calculon/calc.c:36-102).  Here the expression is compiled once on the
 * host (simcu_host.cpp, calcCompile) to postfix bytecode; this interpreter
 * evaluates value and one partial derivative with dual numbers, with exactly
 * the semantic actions of libs/calculon/parser.lem:60-149 (including its
 * quirks: guarded division Div(x,y,gmin), ln(|x|), d acosh = sinh, u'(0) =
 * 1/gmin, uramp'(0) = 0.5).
 *
 * __host__ __device__: the same code runs in the kernels and, for the
 * simcuCalcEval test hook only, on the host.
 */
#ifndef SIMCU_CALC_H
#define SIMCU_CALC_H



#ifdef __CUDACC__
#define SIMCU_HD __host__ __device__ __forceinline__
#else
#define SIMCU_HD static inline
#endif

SIMCU_HD double calcDiv(double x, double y, double m)
{
	/* parser.lem:30-31: x/y if |y| > m, else x/(+-m) - as ONE division */
	const double den = (fabs(y) > m) ? y : ((y > 0) ? m : -m);
	return x / den;
}

SIMCU_HD double calcLn(double x) { return log(fabs(x)); }

/* one unary function on a dual number: parser.lem:109-126 */
SIMCU_HD void calcFunc(int id, double bf, double bd, double m, double *af, double *ad)
{
	switch (id) {
	case FN_ABS: *af = fabs(bf); *ad = calcDiv(bf, fabs(bf), m) * bd; break;
	case FN_ACOSH: *af = acosh(bf); *ad = sinh(bf) * bd; break;
	case FN_ACOS: *af = acos(bf); *ad = calcDiv(-1, sqrt(1 - bf * bf), m) * bd; break;
	case FN_ASINH: *af = asinh(bf); *ad = calcDiv(1, sqrt(1 + bf * bf), m) * bd; break;
	case FN_ASIN: *af = asin(bf); *ad = calcDiv(1, sqrt(1 - bf * bf), m) * bd; break;
	case FN_ATANH: *af = atanh(bf); *ad = calcDiv(1, (1 - bf * bf), m) * bd; break;
	case FN_ATAN: *af = atan(bf); *ad = calcDiv(1, (1 + bf * bf), m) * bd; break;
	case FN_COSH: *af = cosh(bf); *ad = sinh(bf) * bd; break;
	case FN_COS: *af = cos(bf); *ad = -1 * sin(bf) * bd; break;
	case FN_EXP: *af = exp(bf); *ad = exp(bf) * bd; break;
	case FN_LN: *af = calcLn(bf); *ad = calcDiv(1, bf, m) * bd; break;
	case FN_LOG: *af = log10(bf); *ad = calcDiv(1, (bf * calcLn(10.0)), m) * bd; break;
	case FN_SINH: *af = sinh(bf); *ad = cosh(bf) * bd; break;
	case FN_SIN: *af = sin(bf); *ad = cos(bf) * bd; break;
	case FN_SQRT: *af = sqrt(bf); *ad = calcDiv(1, (2 * sqrt(bf)), m) * bd; break;
	case FN_TAN: *af = tan(bf);
		*ad = calcDiv(1, cos(bf), m) * calcDiv(1, cos(bf), m) * bd; break;
	case FN_URAMP: *af = (bf > 0) ? bf : 0;
		*ad = ((bf == 0) ? 0.5 : ((bf > 0) ? 1.0 : 0)) * bd; break;
	default: /* FN_U */ *af = (bf > 0) ? 1 : 0;
		*ad = (bf == 0) ? 1 / m : 0.0; break;
	}
}

/*
 * Evaluates code[0..n) (pairs: opcode, operand).  vals[v] is the value of the
 * expression's v-th variable; dvar is the variable to differentiate against
 * (-1: none).  Returns the value in *f and the derivative in *d.
 */
SIMCU_HD void calcRun(const int *code, int n, const double *consts,
		const double *vals, int dvar, double m, double *f, double *d)
{
	double sf[CALC_MAX_STACK], sd[CALC_MAX_STACK];
	int sp = 0;
	for (int pc = 0; pc < n; pc += 2) {
		const int op = code[pc], arg = code[pc + 1];
		switch (op) {
		case OP_CONST: sf[sp] = consts[arg]; sd[sp] = 0.0; sp++; break;
		case OP_VAR: sf[sp] = vals[arg]; sd[sp] = (arg == dvar) ? 1.0 : 0.0; sp++; break;
		case OP_NEG: sf[sp - 1] = -1 * sf[sp - 1]; sd[sp - 1] = -1 * sd[sp - 1]; break;
		case OP_ADD: sp--; sf[sp - 1] = sf[sp - 1] + sf[sp]; sd[sp - 1] = sd[sp - 1] + sd[sp]; break;
		case OP_SUB: sp--; sf[sp - 1] = sf[sp - 1] - sf[sp]; sd[sp - 1] = sd[sp - 1] - sd[sp]; break;
		case OP_MUL: {
			sp--;
			const double bf = sf[sp - 1], bd = sd[sp - 1], cf = sf[sp], cd = sd[sp];
			sf[sp - 1] = bf * cf; sd[sp - 1] = bd * cf + bf * cd;
			break;
		}
		case OP_DIV: {
			sp--;
			const double bf = sf[sp - 1], bd = sd[sp - 1], cf = sf[sp], cd = sd[sp];
			sf[sp - 1] = calcDiv(bf, cf, m);
			sd[sp - 1] = calcDiv((bd * cf - bf * cd), (cf * cf), m);
			break;
		}
		case OP_POW: {
			sp--;
			const double bf = sf[sp - 1], bd = sd[sp - 1], cf = sf[sp], cd = sd[sp];
			const double p = pow(bf, cf);
			sf[sp - 1] = p;
			sd[sp - 1] = p * (bd * calcDiv(cf, bf, m) + (bf ? cd * calcLn(bf) : 0));
			break;
		}
		case OP_FUNC: calcFunc(arg, sf[sp - 1], sd[sp - 1], m, &sf[sp - 1], &sd[sp - 1]); break;
		case OP_GT: sp--; sf[sp - 1] = (sf[sp - 1] > sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_LT: sp--; sf[sp - 1] = (sf[sp - 1] < sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_GE: sp--; sf[sp - 1] = (sf[sp - 1] >= sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_LE: sp--; sf[sp - 1] = (sf[sp - 1] <= sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_EQ: sp--; sf[sp - 1] = (sf[sp - 1] == sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_NE: sp--; sf[sp - 1] = (sf[sp - 1] != sf[sp]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_NOT: sf[sp - 1] = (sf[sp - 1] == 0) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_AND: sp--; sf[sp - 1] = ((sf[sp - 1] == 1) && (sf[sp] == 1)) ? 1 : 0; sd[sp - 1] = 0.0; break;
		case OP_OR: sp--; sf[sp - 1] = ((sf[sp - 1] == 1) || (sf[sp] == 1)) ? 1 : 0; sd[sp - 1] = 0.0; break;
		default: /* OP_IF */ sf[sp - 1] = (sf[sp - 1]) ? 1 : 0; sd[sp - 1] = 0.0; break;
		}
	}
	*f = sf[0];
	*d = sd[0];
}

#endif
/*
 * simcu_device.cuh - the device library of the batched transient engine: every
 * device model, companion model, table, waveform and test of the reference's
 * libs/simulator restated for one-thread-per-instance execution.  Included by
 * the generic kernels (simcu_kernels.cu, nvcc) and, with SIMCU_JIT defined, by
 * the topology-specialised translation unit that simcu_codegen.cpp writes and
 * NVRTC compiles (sm_100a).  Semantics follow the reference file:line cited at
 * each function (paths relative to the reference root).
 */
#ifndef SIMCU_DEVICE_CUH
#define SIMCU_DEVICE_CUH

#define DEV __device__ __forceinline__
/* Heavy, state-free helpers are real functions: the specialised kernel inlines
 * one copy of every device per netlist entry, and without this its code
 * outgrows the instruction cache (measured: stall_no_inst was the top stall). */
#define DEVFN __device__ __noinline__

DEVFN double simcuAtan2Fn(double y, double x) { return atan2(y, x); }
/* atan2(+-0, x > 0) is +-0 exactly (C99 F.9.1.4): the break-point test of a
 * signal that did not move - the common case - needs no call */
DEV double simcuAtan2(double y, double x) { return (y == 0.0 && x > 0.0) ? y : simcuAtan2Fn(y, x); }
DEVFN double simcuExp(double x) { return exp(x); }

#define MaxAbs(x, y) ((fabs(x) > fabs(y)) ? fabs(x) : fabs(y))
#define Min3(x, y, z) ((x < y) ? ((z < x) ? z : x) : ((z < y) ? z : y))

#ifdef SIMCU_JIT
#define SIMCU_NH_LOOP SIMCU_NH
#define SIMCU_N_LOOP SIMCU_N
#define SIMCU_NPROBES_LOOP SIMCU_NPROBES
#else
#define SIMCU_NH_LOOP a.nh
#define SIMCU_N_LOOP a.N
#define SIMCU_NPROBES_LOOP a.nprobes
#endif

/* ------------------------------------------------------------------------ *
 * Per-thread view of one instance.  Two storage policies share every device
 * function below:
 *   generic (pilot / interpreter kernels, simcu_kernels.cu): all state in
 *     global memory, structure-of-arrays with the instance index fastest;
 *   SIMCU_JIT (topology-specialised kernel, compiled per netlist by NVRTC):
 *     all state in thread-local arrays that are only ever indexed with
 *     compile-time constants, so the compiler keeps the hot part in registers
 *     and spills the rest to local memory - which the hardware lays out exactly
 *     like the SoA above, one slot per RESIDENT thread instead of per instance.
 * S() is state addressed with constant slots; D() (generic kernels only) is the
 * part the reference indexes with a ring counter - the integrator and break-
 * point rings, which the specialised kernel stores by age instead.
 * ------------------------------------------------------------------------ */
#ifdef SIMCU_JIT
#define SIMCU_UNROLL _Pragma("unroll")
#define LDI(p) (*(p))
#define SIMCU_DIM(n) ((n) > 0 ? (n) : 1)
struct Inst {
	const SimcuArgs &a;
	int i;               /* instance */
	int hslot;           /* history ring column of this resident thread */
	size_t hstride;
	int hcount, npts, gridIdx;
	double time;
	int order;
	int err;             /* SIMCU_ERR_* of this instance (0 = ok) */
	int growth;          /* factorisations with an LU multiplier above SIMCU_LMAX */
	/* replay (test instrumentation): logged outcomes of the convergence tests of
	 * this linearize pass, next bit to consume, own decisions that differed */
	mutable int linMask, linBit, linDiff;
	mutable double Av[SIMCU_DIM(SIMCU_NNZ)], Bv[SIMCU_DIM(SIMCU_N)];
	mutable double Xv[SIMCU_DIM(SIMCU_N)], Xa[SIMCU_DIM(SIMCU_N)];
	mutable double sv[SIMCU_DIM(SIMCU_SD)];
	mutable int iv[SIMCU_DIM(SIMCU_SI)];
	/* time and step rings of the integrators: every C / L / nonlinear C of an
	 * instance is initialised and stepped at the same times, so the reference's
	 * per-device copies (integrator.c:44-57) hold identical values - one copy */
	mutable double itT[4], itH[3];
	mutable int itN;
	mutable bool itAdv;
	double par[SIMCU_DIM(SIMCU_NPARAM)];
	int tset[SIMCU_DIM(SIMCU_NISET)];

	DEV Inst(const SimcuArgs &args, int inst, int slot, size_t stride)
		: a(args), i(inst), hslot(slot), hstride(stride), hcount(0), npts(0), gridIdx(0),
		  time(0.0), order(1), err(0) {}

	DEV double P(int pid) const { return par[pid]; }
	DEV int tableSet(int slot) const { return (slot >= 0) ? tset[slot] : 0; }
	DEV double &S(int slot) const { return sv[slot]; }
	DEV int &IS(int slot) const { return iv[slot]; }
	DEV void Aplus(int slot, double d) const { if (slot >= 0) Av[slot] += d; }
	DEV void Aset(int slot, double v) const { if (slot >= 0) Av[slot] = v; }
	DEV void Bplus(int row, double d) const { if (row) Bv[row - 1] += d; }
	DEV double X(int row) const { return row ? Xv[row - 1] : 0.0; }
};
DEV const int *bvOf(const Inst &, const int *, const int *bv) { return bv; }
#else
#define SIMCU_UNROLL
#define LDI(p) __ldg(p)
struct Inst {
	const SimcuArgs &a;
	size_t M;
	int i;
	int hslot;
	size_t hstride;
	int hcount, npts, gridIdx;
	double *ws;          /* LU workspace of this thread */
	int wsStride;
	bool xInWs;          /* true: devices read the Newton iterate from ws */
	double time;
	int order;
	int err;             /* SIMCU_ERR_* of this instance (0 = ok) */

	DEV Inst(const SimcuArgs &args, int inst, double *w, int stride)
		: a(args), M(args.M), i(inst), hslot(inst), hstride(args.M), hcount(0), npts(0),
		  gridIdx(0), ws(w), wsStride(stride), xInWs(false), time(0.0), order(1), err(0) {}

	DEV double P(int pid) const
	{
		const int m = __ldg(&a.pmap[pid]);
		return (m >= 0) ? a.pinst[(size_t)m * M + i] : __ldg(&a.pconst[pid]);
	}
	DEV int tableSet(int slot) const { return (slot >= 0) ? a.iinst[(size_t)slot * M + i] : 0; }
	DEV double &S(int slot) const { return a.S[(size_t)slot * M + i]; }
	DEV double &D(int slot) const { return a.S[(size_t)slot * M + i]; }
	DEV int &IS(int slot) const { return a.IS[(size_t)slot * M + i]; }
	DEV double &CD(int k) const { return a.CD[(size_t)k * M + i]; }
	DEV int &CI(int k) const { return a.CI[(size_t)k * M + i]; }
	/* stamp primitives: node.c:86-118, row.c:85-116 (ground is a sink) */
	DEV void Aplus(int slot, double d) const { if (slot >= 0) a.A[(size_t)slot * M + i] += d; }
	DEV void Aset(int slot, double v) const { if (slot >= 0) a.A[(size_t)slot * M + i] = v; }
	DEV void Bplus(int row, double d) const { if (row) a.B[(size_t)(row - 1) * M + i] += d; }
	DEV double X(int row) const
	{
		if (!row) return 0.0;
		if (xInWs) return ws[(size_t)__ldg(&a.xslot[row - 1]) * wsStride];
		return a.X[(size_t)(row - 1) * M + i];
	}
};
DEV const int *bvOf(const Inst &c, const int *d, const int *) { return c.a.bvar + 4 * d[DF_BVOFF]; }
#endif
/* ------------------------------------------------------------------------ *
 * Piecewise tables: libs/simulator/math/piecewise.c:44-67,96-104,226-263
 * ------------------------------------------------------------------------ */
/* A table is stored packed, one 32-byte entry per point: (x_i, y_i, m_i, x_{i+1})
 * with m = slope (linear) or second derivative (cubic) and x_{i+1} = +inf for
 * the last point - so the index walk and a linear lookup cost ONE read-only
 * 32-byte fetch per visited point instead of a chain of dependent loads. */
struct Table { const double *p; int len; int type; };
struct TabEntry { double x, y, m, xn; };

DEV TabEntry tabLoad(const Table &t, int i)
{
	const double2 a = __ldg(reinterpret_cast<const double2 *>(t.p + 4 * (size_t)i));
	const double2 b = __ldg(reinterpret_cast<const double2 *>(t.p + 4 * (size_t)i) + 1);
	TabEntry e;
	e.x = a.x; e.y = a.y; e.m = b.x; e.xn = b.y;
	return e;
}

DEV Table tableOf(const SimcuArgs &a, int t)
{
	Table r;
	const int off = __ldg(&a.tabOff[t]);
	r.len = __ldg(&a.tabOff[t + 1]) - off;
	r.type = __ldg(&a.tabType[t]);
	r.p = a.tabP + 4 * (size_t)off;
	return r;
}

/* piecewise.c:44-67.  The cached index only shortens the walk; the landing
 * index depends on x0 alone.  Returns the entry at the landing index in e. */
DEV int pwSetIndex(const Table &t, int index, double x0, TabEntry &e)
{
	if ((index < 0) || (index > t.len - 1)) index = 0;
	e = tabLoad(t, index);
	while ((index >= 0) && (index < t.len - 1)) {
		if (e.x <= x0) {
			if (index == t.len - 2) return index;
			if (e.xn > x0) return index;
			index++;
		} else {
			if (index == 0) return index;
			index--;
		}
		e = tabLoad(t, index);
	}
	return index;
}

struct PwValue { double y, dydx; int index; };

/* piecewise.c:226-263 */
DEVFN PwValue pwCalcValueFn(const double *tp, int len, int type, int index, double x0)
{
	Table t;
	t.p = tp; t.len = len; t.type = type;
	PwValue r;
	TabEntry e;
	index = pwSetIndex(t, index, x0, e);
	r.index = index;
	const int i = index;
	if ((i == 0) || (i >= t.len - 2)) {
		const TabEntry first = tabLoad(t, 0), last = tabLoad(t, t.len - 1);
		if (first.x > x0) { r.y = first.y; r.dydx = 0; return r; }
		if (last.x < x0) { r.y = last.y; r.dydx = 0; return r; }
	}
	if (t.type == 'l') {
		r.y = e.y + e.m * (x0 - e.x);
		r.dydx = e.m;
	} else {
		const TabEntry n = tabLoad(t, i + 1);
		const double xi = e.x, yi = e.y;
		const double mi = e.m, mi1 = n.m;
		const double dx0 = x0 - xi, dy = n.y - yi;
		const double dx = n.x - xi;
		const double pa = dy / dx - dx * (mi1 + mi * 2.0) / 6.0;
		const double pb = 0.5 * mi;
		const double pc = (mi1 - mi) / (6.0 * dx);
		r.y = yi + dx0 * (pa + dx0 * (pb + dx0 * pc));
		r.dydx = pa + dx0 * (2 * pb + 3 * pc * dx0);
	}
	return r;
}

DEV void pwCalcValue(const Table &t, int &index, double x0, double &y0, double &dydx0)
{
	const PwValue r = pwCalcValueFn(t.p, t.len, t.type, index, x0);
	index = r.index; y0 = r.y; dydx0 = r.dydx;
}

/* piecewise.c:204-222 with its two dead branches */
DEV double pwGetNextX(const Table &t, int &index, double x0)
{
	TabEntry e;
	index = pwSetIndex(t, index, x0, e);
	if (index == t.len - 1) return HUGE_VAL;
	return e.xn;
}

/* ------------------------------------------------------------------------ *
 * Break-point test: libs/simulator/math/checkbreak.c:43-67,71-95
 * state: x[3] t[3] theta[3] at sb, n at ib
 * ------------------------------------------------------------------------ */
#ifdef SIMCU_JIT
/* Specialised kernel: the rings are stored by AGE (0 = the slot of the current
 * time point, 1 = the one before, ...) instead of by ring position n % 3, so
 * every access has a compile-time index and the state can live in registers;
 * advancing the ring counter becomes a rotation.  Same values, same tests. */
/* The reference compares theta = atan2(dx, dt) of this time point with the one
 * of the previous point (checkbreak.c:57-66).  In raw SI units dt ~ 1e-11, so
 * |theta| is within 1e-3 of pi/2 for any signal that moves, and exactly 0 for
 * one that does not: > 97 % of the calls on the IBIS circuits.  The specialised
 * kernel therefore keeps the ARGUMENTS (dx, dt) of both angles instead of the
 * angles and decides from certified bounds - small: |dx| <= dt/2, |theta| <=
 * 0.4637; big: |dx| >= 1000 dt, |theta| >= pi/2 - 0.001 - whenever they settle
 * the comparison with pi/3 (margins >= 0.05 rad); otherwise from single-precision
 * angles if those are more than 1e-3 away from the threshold, and only then from
 * the two double-precision atan2 of the reference.  The decision is the
 * reference's in every case; its cost is a handful of compares. */
DEVFN int cbAngleBreak(double dy0, double dx0, double dy1, double dx1, double maxAngle)
{
	if (maxAngle > 0.93 && maxAngle < 1.10) {
		const double a0 = fabs(dy0), a1 = fabs(dy1);
		const bool f0 = (dx0 > 0.0) && (dx0 < HUGE_VAL), f1 = (dx1 > 0.0) && (dx1 < HUGE_VAL);
		const bool s0 = f0 && (a0 <= 0.5 * dx0), s1 = f1 && (a1 <= 0.5 * dx1);
		const bool b0 = f0 && (a0 >= 1000.0 * dx0) && (a0 < HUGE_VAL);
		const bool b1 = f1 && (a1 >= 1000.0 * dx1) && (a1 < HUGE_VAL);
		if (s0 && s1) return 0;
		if (b0 && b1) return ((dy0 > 0.0) != (dy1 > 0.0)) ? 1 : 0;
		if ((s0 && b1) || (b0 && s1)) return 1;
		if (dx0 > 1e-30 && dx1 > 1e-30 && a0 < 1e30 && a1 < 1e30 && dx0 < 1e30 && dx1 < 1e30) {
			const float d = fabsf(atan2f((float)dy0, (float)dx0) - atan2f((float)dy1, (float)dx1));
			const float m = (float)maxAngle;
			if (d > m + 1e-3f) return 1;
			if (d < m - 1e-3f) return 0;
		}
	}
	return (fabs(simcuAtan2(dy0, dx0) - simcuAtan2(dy1, dx1)) > maxAngle) ? 1 : 0;
}

/* state (6 live doubles of the 9 slots): x0 x1 - | t0 t1 dt1 | - dx1 - ; 0 = the
 * slot of the current time point, 1 = the one before (the reference's third
 * ring slot is never read).  (dx1, dt1) are the arguments of the PREVIOUS point's
 * angle; they are taken over when the ring advances - x0 - x1 and t0 - t1 are
 * then exactly what the last evaluation of that point used. */
DEV void cbInitialize(const Inst &c, int sb, int ib, double ic)
{
	c.S(sb) = ic; c.S(sb + 1) = ic;
	c.S(sb + 3) = 0.0; c.S(sb + 4) = 0.0; c.S(sb + 5) = 0.0;
	c.S(sb + 7) = 0.0;
	c.IS(ib) = 0;
}

DEV int cbIsBreak(const Inst &c, int sb, int ib, double x, double maxAngle)
{
	int n = c.IS(ib);
	if (c.time > c.S(sb + 3)) {
		n++; c.IS(ib) = n;
		c.S(sb + 7) = c.S(sb) - c.S(sb + 1);         /* arguments of theta[p] */
		c.S(sb + 5) = c.S(sb + 3) - c.S(sb + 4);
		c.S(sb + 1) = c.S(sb);                       /* x */
		c.S(sb + 4) = c.S(sb + 3);                   /* t */
	}
	c.S(sb) = x;
	c.S(sb + 3) = c.time;
	/* (n - 1) % 3 < 0 only while the counter is still 0 (never in a transient
	 * attempt, time > 0): then p = slot 0 = k and both angles are atan2(0, 0) */
	if (n < 1) return 0;
	const double dy = x - c.S(sb + 1), dt = c.time - c.S(sb + 4);
	return cbAngleBreak(dy, dt, c.S(sb + 7), c.S(sb + 5), maxAngle);
}
#else
DEV void cbInitialize(const Inst &c, int sb, int ib, double ic)
{
	for (int k = 0; k < 3; k++) { c.D(sb + k) = ic; c.D(sb + 3 + k) = 0.0; }
	/* theta[] is not reset by the reference; it is zero from allocation */
	c.IS(ib) = 0;
}

DEV int cbIsBreak(const Inst &c, int sb, int ib, double x, double maxAngle)
{
	int n = c.IS(ib);
	if (c.time > c.D(sb + 3 + n % 3)) { n++; c.IS(ib) = n; }
	const int k = n % 3;
	int p = (n - 1) % 3;
	if (p < 0) p = 0;
	c.D(sb + k) = x;
	c.D(sb + 3 + k) = c.time;
	const double th = simcuAtan2(x - c.D(sb + p), c.time - c.D(sb + 3 + p));
	c.D(sb + 6 + k) = th;
	return (fabs(th - c.D(sb + 6 + p)) > maxAngle) ? 1 : 0;
}

#endif

/* ------------------------------------------------------------------------ *
 * Newton convergence test: libs/simulator/math/checklinear.c:39-91
 * state: lastV (double), state (int 0='a' 1='b' 2='c')
 * ------------------------------------------------------------------------ */
DEV int clIsLinear(const Inst &c, int sLast, int iState, double tol, double V,
		double calcedV)
{
	const double reltol = c.a.ctl.reltol;
	const double lastV = c.S(sLast);
	const double e1 = reltol * MaxAbs(lastV, V) + tol;
	const double e2 = reltol * MaxAbs(calcedV, V) + tol;
	bool passed = !(fabs(lastV - V) > e1) && !(fabs(calcedV - V) > e2);
#if defined(SIMCU_JIT) && SIMCU_REPLAY
	/* replay: these two threshold tests at reltol are what a 1-ulp difference can
	 * flip; take the oracle's outcome and count the disagreements */
	{
		const bool logged = ((c.linMask >> c.linBit) & 1) != 0;
		c.linBit++;
		if (logged != passed) c.linDiff++;
		passed = logged;
	}
#endif
	if (!passed) { c.IS(iState) = 0; c.S(sLast) = V; return 0; }
	c.S(sLast) = V;
	const int st = c.IS(iState);
	if (st == 1) { c.IS(iState) = 2; return 1; }
	if (st == 2) return 1;
	c.IS(iState) = 1;
	return 0;
}

/* ------------------------------------------------------------------------ *
 * Integrator: libs/simulator/math/integrator.c.  State at sb: y0, dydx0, then
 * the five rings t h x y f (4 each) contiguous, so the reference's negative
 * ring index at n == 1 (integrator.c:90-96) reads the same neighbours.
 * ------------------------------------------------------------------------ */
DEV double DD1(double xnp1, double xn, double hn)
{
	return (hn == 0.0) ? 0.0 : ((xnp1 - xn) / hn);
}
DEV double DD2(double xnp1, double xn, double x1, double hn, double h1)
{
	return (DD1(xnp1, xn, hn) - DD1(xn, x1, h1)) / (hn + h1);
}
DEV double DD3(double xnp1, double xn, double x1, double x2, double hn, double h1, double h2)
{
	return (DD2(xnp1, xn, x1, hn, h1) - DD2(xn, x1, x2, h1, h2)) / (hn + h1 + h2);
}

#ifdef SIMCU_JIT
/* Specialised kernel: rings by age.  AGE(k, a) is ring k (t h x y f) at the
 * slot a steps behind the current one; a = 3 of the time ring is the "next"
 * slot that holds the time of the present attempt.  The reference's negative
 * ring index while the counter is below a (integrator.c:90-96, (n-a) % 4 < 0)
 * lands on the same age of the previous ring; AGEQ reproduces that. */
#define AGE(c, sb, k, a) (c).S((sb) + 2 + (k) * 4 + (a))
#define AGEQ(c, sb, k, a, nn) (((nn) >= (a)) ? AGE(c, sb, k, a) : AGE(c, sb, (k) - 1, a))
enum { RT = 0, RH = 1, RX = 2, RY = 3, RF = 4 };

DEV void itInitialize(const Inst &c, int sb, int ib, double ic, double ydtdx)
{
	(void)ib;
	c.S(sb) = 0.0; c.S(sb + 1) = 0.0;
	c.itN = 0; c.itAdv = false;
	SIMCU_UNROLL
	for (int j = 0; j < 4; j++) {
		c.itT[j] = 0.0;
		AGE(c, sb, RY, j) = 0.0; AGE(c, sb, RX, j) = ic; AGE(c, sb, RF, j) = ydtdx;
	}
	SIMCU_UNROLL
	for (int j = 0; j < 3; j++) c.itH[j] = c.a.ctl.tstop;
}

/* once per attempt, before the devices integrate: advance the shared time ring
 * if time moved on (integrator.c:124-126) and note the step */
DEV void itBegin(const Inst &c)
{
	const double t0 = c.time;
	c.itAdv = (t0 > c.itT[3]);
	if (c.itAdv) {
		c.itN++;
		const double tNext = c.itT[3];
		c.itT[3] = c.itT[2]; c.itT[2] = c.itT[1]; c.itT[1] = c.itT[0]; c.itT[0] = tNext;
		c.itH[2] = c.itH[1]; c.itH[1] = c.itH[0];
	}
	c.itH[0] = t0 - c.itT[0];
	c.itT[3] = t0;
}

/* integrator.c:106-163 */
DEV void itIntegrate(Inst &c, int sb, int ib, double x0, double ydtdx,
		double &dydx0, double &y0)
{
	(void)ib;
	if (isnan(x0)) { c.err = SIMCU_ERR_NAN; return; }
	const int nn = c.itN;
	if (c.itAdv) {
		SIMCU_UNROLL
		for (int k = RX; k <= RF; k++) {
			AGE(c, sb, k, 2) = AGE(c, sb, k, 1); AGE(c, sb, k, 1) = AGE(c, sb, k, 0);
		}
	}
	const double h = c.itH[0];
	AGE(c, sb, RX, 0) = x0;
	const double yn = c.S(sb + 1) * x0 - c.S(sb);
	AGE(c, sb, RY, 0) = yn;
	AGE(c, sb, RF, 0) = ydtdx;
	const double fp = AGEQ(c, sb, RF, 1, nn);
	if (c.order < 2) {
		dydx0 = fp / h;
		y0 = dydx0 * x0;
	} else {
		dydx0 = 2.0 * ydtdx / h;
		double mult = ydtdx / fp;
		if (isnan(mult)) mult = 1 / c.a.ctl.gmin;
		y0 = dydx0 * x0 + mult * yn;
	}
	c.S(sb + 1) = dydx0;
	c.S(sb) = y0;
}

/* integrator.c:61-102 (LTE step recommendation; float sqrtf at :99) */
DEVFN double itNextStepFn(int order, double reltol, double chgtol, double trtol, double abstol,
		double x0, double ydtdx, double ieq, double g, double ry0, double hn, double xn, double fn,
		double f1, double x1, double h1, double f2, double x2, double h2)
{
	const double y0 = g * x0 - ieq;
	const double ey = reltol * MaxAbs(ry0, y0) + abstol;
	const double eyp = reltol * MaxAbs(MaxAbs(x0, xn) * (ydtdx / hn), chgtol);
	const double e = MaxAbs(eyp, ey);
	double h, dd;
	if (order < 2) {
		dd = DD2((ydtdx * x0), (fn * xn), (f1 * x1), hn, h1);
		dd *= 1.0 / 2.0;
		h = trtol * e / MaxAbs(dd, abstol);
	} else {
		dd = DD3((ydtdx * x0), (fn * xn), (f1 * x1), (f2 * x2), hn, h1, h2);
		dd *= 1.0 / 12.0;
		h = sqrtf(trtol * e / MaxAbs(dd, abstol));
	}
	return h;
}

DEV double itNextStep(Inst &c, int sb, int ib, double x0, double ydtdx, double abstol)
{
	if (isnan(x0)) { c.err = SIMCU_ERR_NAN; return 0.0; }
	const SimcuControl &k = c.a.ctl;
	const int nn = c.itN;
	(void)ib;
	/* below the counter the reference reads the same age of the PREVIOUS ring
	 * (AGEQ): x <- h, h <- t, f <- y */
	const double x1 = (nn >= 1) ? AGE(c, sb, RX, 1) : c.itH[1], x2 = (nn >= 2) ? AGE(c, sb, RX, 2) : c.itH[2];
	const double h1 = (nn >= 1) ? c.itH[1] : c.itT[1], h2 = (nn >= 2) ? c.itH[2] : c.itT[2];
	const double h = itNextStepFn(c.order, k.reltol, k.chgtol, k.trtol, abstol, x0, ydtdx,
			c.S(sb), c.S(sb + 1), AGE(c, sb, RY, 0), c.itH[0], AGE(c, sb, RX, 0),
			AGE(c, sb, RF, 0), AGEQ(c, sb, RF, 1, nn), x1, h1, AGEQ(c, sb, RF, 2, nn), x2, h2);
	if (isnan(h)) c.err = SIMCU_ERR_NAN;
	return h;
}
#else
#define RING(c, sb, k, j) (c).D((sb) + 2 + (k) * 4 + (j))
enum { RT = 0, RH = 1, RX = 2, RY = 3, RF = 4 };

DEV void itInitialize(const Inst &c, int sb, int ib, double ic, double ydtdx)
{
	c.D(sb) = 0.0; c.D(sb + 1) = 0.0; c.IS(ib) = 0;
	for (int j = 0; j < 4; j++) {
		RING(c, sb, RT, j) = 0.0; RING(c, sb, RH, j) = c.a.ctl.tstop;
		RING(c, sb, RY, j) = 0.0; RING(c, sb, RX, j) = ic; RING(c, sb, RF, j) = ydtdx;
	}
}

/* integrator.c:106-163 */
DEV void itIntegrate(Inst &c, int sb, int ib, double x0, double ydtdx,
		double &dydx0, double &y0)
{
	if (isnan(x0)) { c.err = SIMCU_ERR_NAN; return; }
	const double t0 = c.time;
	int nn = c.IS(ib);
	if (t0 > RING(c, sb, RT, (nn + 1) % 4)) { nn++; c.IS(ib) = nn; }
	const int n = nn % 4, p = (nn - 1) % 4;
	const double h = t0 - RING(c, sb, RT, n);
	RING(c, sb, RH, n) = h;
	RING(c, sb, RT, (nn + 1) % 4) = t0;
	RING(c, sb, RX, n) = x0;
	const double yn = c.D(sb + 1) * x0 - c.D(sb);
	RING(c, sb, RY, n) = yn;
	RING(c, sb, RF, n) = ydtdx;
	if (c.order < 2) {
		dydx0 = RING(c, sb, RF, p) / h;
		y0 = dydx0 * x0;
	} else {
		dydx0 = 2.0 * ydtdx / h;
		double mult = ydtdx / RING(c, sb, RF, p);
		if (isnan(mult)) mult = 1 / c.a.ctl.gmin;
		y0 = dydx0 * x0 + mult * yn;
	}
	c.D(sb + 1) = dydx0;
	c.D(sb) = y0;
}

/* integrator.c:61-102 (LTE step recommendation; float sqrtf at :99) */
DEV double itNextStep(Inst &c, int sb, int ib, double x0, double ydtdx, double abstol)
{
	if (isnan(x0)) { c.err = SIMCU_ERR_NAN; return 0.0; }
	const SimcuControl &k = c.a.ctl;
	const int nn = c.IS(ib);
	const int n = nn % 4, p1 = (nn - 1) % 4, p2 = (nn - 2) % 4;
	const double y0 = c.D(sb + 1) * x0 - c.D(sb);
	const double ey = k.reltol * MaxAbs(RING(c, sb, RY, n), y0) + abstol;
	const double hn = RING(c, sb, RH, n), xn = RING(c, sb, RX, n);
	const double eyp = k.reltol * MaxAbs(MaxAbs(x0, xn) * (ydtdx / hn), k.chgtol);
	const double e = MaxAbs(eyp, ey);
	double h, dd;
	if (c.order < 2) {
		dd = DD2((ydtdx * x0), (RING(c, sb, RF, n) * xn),
				(RING(c, sb, RF, p1) * RING(c, sb, RX, p1)), hn, RING(c, sb, RH, p1));
		dd *= 1.0 / 2.0;
		h = k.trtol * e / MaxAbs(dd, abstol);
	} else {
		dd = DD3((ydtdx * x0), (RING(c, sb, RF, n) * xn),
				(RING(c, sb, RF, p1) * RING(c, sb, RX, p1)),
				(RING(c, sb, RF, p2) * RING(c, sb, RX, p2)),
				hn, RING(c, sb, RH, p1), RING(c, sb, RH, p2));
		dd *= 1.0 / 12.0;
		h = sqrtf(k.trtol * e / MaxAbs(dd, abstol));
	}
	if (isnan(h)) c.err = SIMCU_ERR_NAN;
	return h;
}

#endif

/* ------------------------------------------------------------------------ *
 * Source waveforms: libs/simulator/math/waveform.c:169-338 with the default
 * substitution of :342-413 applied where the argument is read
 * ------------------------------------------------------------------------ */
DEV double wfArg(const Inst &c, int pargs, int k, double dflt)
{
	const double v = c.P(pargs + k);
	return (v == HUGE_VAL) ? dflt : v;
}

DEV double wfValue(Inst &c, const int *d, int sb, int ib)
{
	const int wtype = d[DF_WTYPE], pa = d[DF_PARGS];
	const SimcuControl &k = c.a.ctl;
	const double tstep = k.tstep, tend = k.tstop + k.tstep, fmin = 1 / tend;
	const double time = c.time;
	switch (wtype) {
	case 'p': {
		const double v1 = c.P(pa), v2 = c.P(pa + 1), td = wfArg(c, pa, 2, 0.0);
		const double tr = c.P(pa + 3), tf = wfArg(c, pa, 4, tstep);
		const double pw = wfArg(c, pa, 5, tend), per = wfArg(c, pa, 6, tend);
		const double tn = fmod(time, per);
		if (tn <= td) return v1;
		if (tn <= (tr + td)) return v1 + ((v2 - v1) / tr) * (tn - td);
		if (tn <= (tr + td + pw)) return v2;
		if (tn <= (tr + td + pw + tf)) return v2 - ((v2 - v1) / (tf)) * (tn - (tr + td + pw));
		return v1;
	}
	case 'g': {
		const double v1 = c.P(pa), v2 = c.P(pa + 1), td = wfArg(c, pa, 2, 0.0);
		double tr = c.P(pa + 3), tf = wfArg(c, pa, 4, tstep);
		const double pw = wfArg(c, pa, 5, tend), per = wfArg(c, pa, 6, tend);
		const double tn = fmod(time, per);
		tr = (tr / 0.672) * 0.281 * 2;
		tf = (tf / 0.672) * 0.281 * 2;
		tr = (tn - (tr / 0.281) * 0.672 - td) / tr;
		tf = (tn - (tf / 0.281) * 0.672 - td - pw) / tf;
		double value = v1;
		value += 0.5 * (v2 - v1) * (1 + erf(tr));
		value += 0.5 * (v1 - v2) * (1 + erf(tf));
		return value;
	}
	case 's': {
		const double vo = c.P(pa), va = c.P(pa + 1), fc = wfArg(c, pa, 2, fmin);
		const double td = wfArg(c, pa, 3, 0.0), df = wfArg(c, pa, 4, 0.0);
		if (time <= td) return vo;
		return vo + va * exp(-(time - td) * df) * sin(2 * M_PI * fc * (time - td));
	}
	case 'e': {
		const double v1 = c.P(pa), v2 = c.P(pa + 1), td1 = wfArg(c, pa, 2, 0.0);
		const double tau1 = wfArg(c, pa, 3, tstep), td2 = wfArg(c, pa, 4, tstep + td1);
		const double tau2 = wfArg(c, pa, 5, tstep);
		if (time <= td1) return v1;
		if (time <= td2) return v1 + (v2 - v1) * (1 - exp(-(time - td1) / tau1));
		return v1 + (v2 - v1) * (1 - exp(-(time - td1) / tau1)) +
			(v1 - v2) * (1 - exp(-(time - td2) / tau2));
	}
	case 'l': case 'c': {
		const Table t = tableOf(c.a, d[DF_WTAB]);
		int idx = c.IS(ib + 1);
		double v, dv;
		pwCalcValue(t, idx, time, v, dv);
		c.IS(ib + 1) = idx;
		return v;
	}
	case 'f': {
		const double vo = c.P(pa), va = c.P(pa + 1), fc = wfArg(c, pa, 2, fmin);
		const double mdi = wfArg(c, pa, 3, 0.0), fs = wfArg(c, pa, 4, fmin);
		return vo + va * sin(2 * M_PI * fc * time + mdi * sin(2 * M_PI * fs * time));
	}
	default: return 0.0;
	}
}

/* waveform.c:169-220; `next` is left untouched where the reference does */
DEV void wfNextStep(Inst &c, const int *d, int ib, double &next)
{
	const int wtype = d[DF_WTYPE], pa = d[DF_PARGS];
	const SimcuControl &k = c.a.ctl;
	const double tstep = k.tstep, tend = k.tstop + k.tstep;
	const double time = c.time;
	switch (wtype) {
	case 'p': {
		const double per = wfArg(c, pa, 6, tend);
		const double tp = fmod(time, per);
		next = wfArg(c, pa, 2, 0.0) - tp;
		if (next > 0) break;
		next += c.P(pa + 3);
		if (next > 0) break;
		next += wfArg(c, pa, 5, tend);
		if (next > 0) break;
		next += wfArg(c, pa, 4, tstep);
		if (next > 0) break;
		next = per - tp;
		break;
	}
	case 's': {
		const double td = wfArg(c, pa, 3, 0.0);
		if (time < td) next = td;            /* absolute, sic (waveform.c:197) */
		break;
	}
	case 'e': {
		const double td1 = wfArg(c, pa, 2, 0.0), td2 = wfArg(c, pa, 4, tstep + td1);
		if (time < td1) next = td1 - time;
		else if (time < td2) next = td2 - time;
		break;
	}
	case 'l': case 'c': {
		const Table t = tableOf(c.a, d[DF_WTAB]);
		int idx = c.IS(ib + 1);
		const double tp = pwGetNextX(t, idx, time);
		c.IS(ib + 1) = idx;
		if (tp != HUGE_VAL) next = tp - time;
		break;
	}
	default: break;
	}
}

/* ------------------------------------------------------------------------ *
 * Behavioural sources: devices/nonlinear_{i,v,c}.c
 * ------------------------------------------------------------------------ */
#ifdef SIMCU_JIT
/* generated per topology: value (dvar < 0) or one partial derivative of the
 * expression of calc id `id` at the current iterate */
DEV void genCalc(Inst &c, int id, int dvar, double &f, double &dv);
/* generated: all partial derivatives at once, g[calc variable] */
DEV void genCalcGrad(Inst &c, int id, double *g);
DEV void bSolve(Inst &c, const int *d, int dvar, double &f, double &dv)
{
	genCalc(c, d[DF_CALC], dvar, f, dv);
	/* calc.c:47,96: a NaN result is a hard error */
	if (isnan((dvar >= 0) ? dv : f)) c.err = SIMCU_ERR_NAN;
}
#else
DEV void bSolve(Inst &c, const int *d, int dvar, double &f, double &dv)
{
	const SimcuArgs &a = c.a;
	const int id = d[DF_CALC];
	const int c0 = __ldg(&a.calcOff[id]), c1 = __ldg(&a.calcOff[id + 1]);
	const int v0 = __ldg(&a.calcVarOff[id]), v1 = __ldg(&a.calcVarOff[id + 1]);
	double vals[CALC_MAX_VARS];
	for (int k = 0; k < v1 - v0; k++) {
		const int ref = __ldg(&a.calcVars[v0 + k]);
		vals[k] = (ref == CALC_VAR_TIME) ? c.time : c.X(ref);
	}
	calcRun(a.calcCode + c0, c1 - c0, a.calcConst, vals, dvar, a.ctl.gmin, &f, &dv);
	/* calc.c:47,96: a NaN result is a hard error */
	if (isnan((dvar >= 0) ? dv : f)) c.err = SIMCU_ERR_NAN;
}
#endif

/* state offsets of the three flavours */
DEV int bOffIc(int type) { return (type == DK_BC) ? 24 : 0; }     /* Ic In Beq lastV */
DEV int bOffG(int type) { return (type == DK_BC) ? 28 : 13; }
DEV int bOffCl(int type) { return (type == DK_BC) ? 1 : 0; }      /* int: clState */

/* deviceNonlinearCalculate: nonlinear_i.c:73-103, _v.c:71-77, _c.c:79-85 */
DEV void bCalculate(Inst &c, const int *d, bool initial)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE] + bOffIc(type);
	double In;
	if (type == DK_BI) In = c.X(d[DF_ROWR]);
	else if (type == DK_BV) In = c.X(d[DF_K]) - c.X(d[DF_J]);
	else In = c.X(d[DF_ROWM]);
	c.S(sb + 1) = In;
	if (isnan(In)) { c.err = SIMCU_ERR_NAN; return; }
	if (type == DK_BI && !initial && In == 0.0) { c.S(sb) = 0.0; return; }
	double f, dv;
	bSolve(c, d, -1, f, dv);
	c.S(sb) = f;
}

/* deviceNonlinearLoad: nonlinear_i.c:107-138, _v.c:85-116, _c.c:92-123 */
DEV void bLoad(Inst &c, const int *d, const int *bvIn)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE] + bOffIc(type);
	const int gb = d[DF_SBASE] + bOffG(type);
	const int nbv = d[DF_NBV];
	const int *bv = bvOf(c, d, bvIn);
	double eqCalc = c.S(sb);
#ifdef SIMCU_JIT
	double grad[CALC_MAX_VARS];
	genCalcGrad(c, d[DF_CALC], grad);
#endif
	SIMCU_UNROLL
	for (int k = 0; k < nbv && !c.err; k++) {
		const int row = LDI(&bv[4 * k]), sA = LDI(&bv[4 * k + 1]);
		const int sB = LDI(&bv[4 * k + 2]), cv = LDI(&bv[4 * k + 3]);
#ifdef SIMCU_JIT
		const double G = grad[cv];
		if (isnan(G)) c.err = SIMCU_ERR_NAN;     /* calc.c:96 */
#else
		double f, G;
		bSolve(c, d, cv, f, G);
#endif
		const double old = c.S(gb + k);
		c.Aplus(sA, -(G - old));
		if (type == DK_BI) c.Aplus(sB, (G - old));
		c.S(gb + k) = G;
		eqCalc -= G * c.X(row);
		if (isnan(eqCalc)) c.err = SIMCU_ERR_NAN;
	}
	const double old = c.S(sb + 2);
	if (type == DK_BI) {
		c.Bplus(d[DF_J], eqCalc - old);
		c.Bplus(d[DF_ROWM], -(eqCalc - old));
	} else {
		c.Bplus(d[DF_ROWR], eqCalc - old);
	}
	c.S(sb + 2) = eqCalc;
}

/* ------------------------------------------------------------------------ *
 * Device phases (core/device.c:45-160: visitor in insertion order)
 * ------------------------------------------------------------------------ */
DEV void devLoad(Inst &c, const int *d, const int *bv = nullptr)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	switch (type) {
	case DK_R: {                                /* resistor.c:47-70 */
		const double g = 1 / c.P(d[DF_P0]);
		c.Aplus(d[DF_A0], g); c.Aplus(d[DF_A0 + 1], -g);
		c.Aplus(d[DF_A0 + 2], -g); c.Aplus(d[DF_A0 + 3], g);
		break;
	}
	case DK_C:                                  /* capacitor.c:136-158 */
		c.S(sb) = 0; c.S(sb + 1) = 0;
		itInitialize(c, sb + 2, ib, 0.0, c.P(d[DF_P0]));
		break;
	case DK_L:                                  /* inductor.c:137-163 */
		c.S(sb) = 0; c.S(sb + 1) = 0;
		itInitialize(c, sb + 2, ib, 0.0, c.P(d[DF_P0]));
		c.Aset(d[DF_A0], 1.0); c.Aset(d[DF_A0 + 1], -1.0);
		c.Aset(d[DF_A0 + 2], 1.0); c.Aset(d[DF_A0 + 3], -1.0);
		break;
	case DK_V: case DK_I: {                     /* source_v.c:99-135, source_i.c:97-128 */
		cbInitialize(c, sb + 1, ib, 0.0);
		double dc;
		const int wtype = d[DF_WTYPE];
		if (wtype == 0) dc = c.P(d[DF_PDC]);
		else if (wtype == 'l' || wtype == 'c') {
			const Table t = tableOf(c.a, d[DF_WTAB]);   /* waveform.c:385-389 */
			int idx = 0; double dv;
			pwCalcValue(t, idx, 0.0, dc, dv);
			c.IS(ib + 1) = idx;
			c.S(sb + 10) = dc;
		} else dc = c.P(d[DF_PARGS]);               /* waveform.c:105,131,... */
		c.S(sb) = dc;
		if (type == DK_V) {
			c.Bplus(d[DF_ROWR], dc);
			c.Aset(d[DF_SA0], 1.0); c.Aset(d[DF_SA0 + 1], -1.0);
			c.Aset(d[DF_SA0 + 2], 1.0); c.Aset(d[DF_SA0 + 3], -1.0);
		} else {
			c.Bplus(d[DF_K], -dc); c.Bplus(d[DF_J], dc);
		}
		break;
	}
	case DK_T: {                                /* tline.c:208-270 */
		cbInitialize(c, sb + 6, ib, 0.0); cbInitialize(c, sb + 15, ib + 1, 0.0);
		c.IS(ib + 2) = 0;
		for (int k = 0; k < 6; k++) c.S(sb + k) = 0.0;
		const int *s = d + DF_TA0;
		/* RK RJ KR JR RL RM SL SM LS MS SK SJ RR SS KS JS LR MR */
		c.Aset(s[4], -1.0); c.Aset(s[5], 1.0); c.Aset(s[10], -1.0); c.Aset(s[11], 1.0);
		c.Aset(s[14], 1.0); c.Aset(s[15], -1.0); c.Aset(s[16], -1.0); c.Aset(s[17], 1.0);
		c.Aset(s[0], 1.0); c.Aset(s[1], -1.0); c.Aset(s[2], 1.0); c.Aset(s[3], -1.0);
		c.Aset(s[6], 1.0); c.Aset(s[7], -1.0); c.Aset(s[8], 1.0); c.Aset(s[9], -1.0);
		const double z0 = c.P(d[DF_PZ0]);
		c.Aplus(s[12], -z0); c.Aplus(s[13], -z0);
		const double loss = c.P(d[DF_PLOSS]);
		c.S(sb + 24) = (loss == HUGE_VAL) ? 1.0 : simcuExp(-loss / 2);
		break;
	}
	case DK_VI: {                               /* vicurve.c:167-221 */
		c.S(sb + 3) = 0.0; c.IS(ib) = 0;
		cbInitialize(c, sb + 4, ib + 1, 0.0);
		const int set = c.tableSet(d[DF_SETSLOT]);
		double aa = 1.0, G = 0.0, Ieq;
		int idx = 0;
		if (d[DF_TATAB] >= 0) {
			const Table ta = tableOf(c.a, d[DF_TATAB] + set);
			pwCalcValue(ta, idx, 0.0, aa, G);
			c.IS(ib + 3) = idx;
		}
		const Table vi = tableOf(c.a, d[DF_VITAB] + set);
		idx = 0;
		pwCalcValue(vi, idx, 0.0, Ieq, G);
		c.IS(ib + 2) = idx;
		G *= aa; Ieq *= aa;
		c.S(sb) = G; c.S(sb + 1) = Ieq; c.S(sb + 2) = aa;
		const int *s = d + DF_VA0;                  /* RK RM KR MR JJ JM MJ MM */
		c.Aset(s[0], 1.0); c.Aset(s[1], -1.0); c.Aset(s[2], 1.0); c.Aset(s[3], -1.0);
		c.Aplus(s[4], G); c.Aplus(s[5], -G); c.Aplus(s[6], -G); c.Aplus(s[7], G);
		c.Bplus(d[DF_J], Ieq); c.Bplus(d[DF_ROWM], -Ieq);
		break;
	}
	case DK_BI: case DK_BV: case DK_BC: {       /* nonlinear_i.c:246-290, _v.c:216-256, _c.c:303-349 */
		const int so = sb + bOffIc(type);
		c.S(so) = 0.0; c.S(so + 1) = 0.0; c.S(so + 2) = 0.0; c.S(so + 3) = 0.0;
		c.IS(ib + bOffCl(type)) = 0;
		for (int k = 0; k < d[DF_NBV]; k++) c.S(sb + bOffG(type) + k) = 0.0;
		const int *s = d + DF_BA0;
		if (type == DK_BC) {
			c.S(sb) = 0; c.S(sb + 1) = 0;
			itInitialize(c, sb + 2, ib, 0.0, 0.0);      /* icc = 0, nonlinear_c.c:427 */
			c.Aset(s[4], 1.0); c.Aset(s[5], 1.0);
		} else {
			cbInitialize(c, sb + 4, ib + 1, 0.0);
			c.Aset(s[0], 1.0); c.Aset(s[1], -1.0); c.Aset(s[2], 1.0); c.Aset(s[3], -1.0);
		}
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d, bv);
		break;
	}
	}
}

DEV int devLinearize(Inst &c, const int *d, const int *bv = nullptr)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	const SimcuControl &k = c.a.ctl;
	if (type == DK_VI) {                        /* vicurve.c:110-163 */
		const double i0 = c.X(d[DF_ROWR]);
		const double v0 = c.X(d[DF_K]) - c.X(d[DF_J]);
		if (isnan(i0) || isnan(v0)) { c.err = SIMCU_ERR_NAN; return 0; }
		double ic, gc;
		if (i0 == 0.0) {
			/* the reference leaves gc uninitialised here (vicurve.c:130-132);
			 * it only matters if the test below fails, then it is clamped */
			ic = 0.0; gc = 0.0;
		} else {
			const int set = c.tableSet(d[DF_SETSLOT]);
			const Table vi = tableOf(c.a, d[DF_VITAB] + set);
			int idx = c.IS(ib + 2);
			pwCalcValue(vi, idx, v0, ic, gc);
			c.IS(ib + 2) = idx;
			const double aa = c.S(sb + 2);
			ic *= aa; gc *= aa;
		}
		const int linear = clIsLinear(c, sb + 3, ib, k.abstol, i0, ic);
		if (!linear) {
			if (gc < k.gmin) gc = k.gmin;
			const double Ieq = (ic - gc * v0), G = gc;
			const double dG = G - c.S(sb);
			const int *s = d + DF_VA0;
			c.Aplus(s[4], dG); c.Aplus(s[5], -dG); c.Aplus(s[6], -dG); c.Aplus(s[7], dG);
			c.S(sb) = G;
			const double dI = Ieq - c.S(sb + 1);
			c.Bplus(d[DF_J], dI); c.Bplus(d[DF_ROWM], -dI);
			c.S(sb + 1) = Ieq;
		}
		return linear;
	}
	/* nonlinear_i.c:224-242, _v.c:194-212, _c.c:188-206 */
	bCalculate(c, d, false);
	if (c.err) return 0;
	const int so = sb + bOffIc(type);
	const double tol = (type == DK_BI) ? k.abstol : (type == DK_BV) ? k.vntol : k.captol;
	const int linear = clIsLinear(c, so + 3, ib + bOffCl(type), tol, c.S(so + 1), c.S(so));
	if (!linear) bLoad(c, d, bv);
	return linear;
}

DEV void devInitStep(Inst &c, const int *d)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	switch (type) {
	case DK_C:                                  /* capacitor.c:112-132 */
		itInitialize(c, sb + 2, ib, c.X(d[DF_K]) - c.X(d[DF_J]), c.P(d[DF_P0]));
		break;
	case DK_L:                                  /* inductor.c:113-133 */
		itInitialize(c, sb + 2, ib, c.X(d[DF_ROWR]), c.P(d[DF_P0]));
		break;
	case DK_BC:                                 /* nonlinear_c.c:281-299 */
		itInitialize(c, sb + 2, ib, c.X(d[DF_K]) - c.X(d[DF_J]), c.S(sb + 24));
		c.S(sb) = 0; c.S(sb + 1) = 0;
		break;
	case DK_T: {                                /* tline.c:93-122 */
		const int *s = d + DF_TA0;
		c.Aset(s[4], 0.0); c.Aset(s[5], 0.0); c.Aset(s[10], 0.0); c.Aset(s[11], 0.0);
		c.Aset(s[14], 0.0); c.Aset(s[15], 0.0); c.Aset(s[16], 0.0); c.Aset(s[17], 0.0);
		c.S(sb + 2) = c.X(d[DF_K]); c.S(sb + 3) = c.X(d[DF_J]);
		c.S(sb + 4) = c.X(d[DF_L]); c.S(sb + 5) = c.X(d[DF_M]);
		break;
	}
	default: break;
	}
}

/* T-line delayed lookup: history_interp.c:42-92, history.c:36-59.  Records
 * are kept in a ring of a.histRecords entries; `count` records exist so far,
 * record r lives in slot r % R.  Returns false if the needed record is gone. */
DEV int tlineWrap(int p, int R) { return (p >= R) ? p - R : ((p < 0) ? p + R : p); }

/* Finds the two history records that bracket (or, past the newest record,
 * precede) the delayed time tp.  times = this thread's record TIMES (dense, one
 * double per record: a walk of +-16 records stays inside one or two cache
 * lines); record q lives at position q % R.  Returns (hint, position of the
 * earlier record, position of the later one, position of the hint record + 1
 * or 0 = the needed record is gone).  One modulo per call. */
DEVFN int4 tlineLocateFn(const double *times, int R, int count, int hint, double tp)
{
	int4 r;
	r.x = hint; r.y = 0; r.z = 0; r.w = 0;
	const int lo = (count > R) ? count - R : 0;
	int i = hint;
	if (i < lo || i >= count) i = lo;
	const int i0 = i, base = i % R;
	/* time of record j in [lo, count): |j - i0| < R, so one wrap at most */
	#define HTIME(j) times[tlineWrap(base + ((j) - i0), R)]
	/* The reference walks record by record from the hint (history_interp.c:
	 * 42-92).  Accepted times increase strictly, so the record it lands on is
	 * THE largest j with t_j <= tp; a galloping + binary search finds the same j
	 * in O(log distance) dependent loads - the walk is long whenever one big
	 * step follows a burst of tiny ones, and a warp waits for its slowest lane. */
	int a, b;                                 /* t_a <= tp, and b == count or t_b > tp */
	if (HTIME(i) > tp) {
		if (i == lo) return r;                /* fell off the ring */
		b = i;
		int step = 1, j = i - 1;
		while (j > lo && HTIME(j) > tp) { b = j; j = (j - step > lo) ? j - step : lo; step <<= 1; }
		if (HTIME(j) > tp) return r;
		a = j;
	} else {
		a = i; b = count;
		int step = 1, j = i + 1, probes = 0;
		while (j < count) {
			if (HTIME(j) <= tp) { a = j; if (++probes > 1) step <<= 1; j = a + step; }
			else { b = j; break; }
		}
	}
	while (b - a > 1) {
		const int mid = a + ((b - a) >> 1);
		if (HTIME(mid) <= tp) a = mid; else b = mid;
	}
	#undef HTIME
	i = a;
	const int pos = tlineWrap(base + (i - i0), R);
	r.x = i;
	if (i + 1 < count) { r.y = pos; r.z = (pos + 1 == R) ? 0 : pos + 1; }
	else {                                    /* extrapolate from the last two */
		if (i - 1 < lo) return r;
		r.y = (pos == 0) ? R - 1 : pos - 1; r.z = pos;
	}
	r.w = pos + 1;                            /* ring position of record `hint`, + 1 */
	return r;
}

DEV double tlineInterp(const double *recP, const double *recN, int col, double tP, double tN, double tp)
{
	if (col < 0) return 0.0;
	const double xP = recP[col], xN = recN[col];
	return ((xN - xP) / (tN - tP)) * (tp - tP) + xP;
}

#ifdef SIMCU_JIT
/* First pass over the T-lines of an attempt (specialised kernel, circuits with
 * several lines): locate the delayed time of this line, remember the landing
 * record as the hint, and ask L2 for the sectors of the two records the second
 * pass (devStep) will interpolate - so the DRAM latencies of all lines overlap
 * instead of adding up, one dependent round trip after the other.  The hint only
 * shortens the search (same landing record), so results do not change. */
DEV void tlinePrefetch(Inst &c, const int *d)
{
	const int ib = d[DF_IBASE];
	const double tp = c.time - c.P(d[DF_PTD]);
	if (!(tp > 0)) return;
	const int R = c.a.histRecords, nh = SIMCU_NH;
	const double *times = c.a.Ht + (size_t)c.hslot * R;
	const double *vals = c.a.Hx + (size_t)c.hslot * R * nh;
	const int4 r = tlineLocateFn(times, R, c.hcount, c.IS(ib + 2), tp);
	if (!r.w) return;
	c.IS(ib + 2) = r.x;
	const int *h = d + DF_TH0;
	int lo = nh, hi = -1;
	SIMCU_UNROLL
	for (int k = 0; k < 6; k++) if (h[k] >= 0) { lo = (h[k] < lo) ? h[k] : lo; hi = (h[k] > hi) ? h[k] : hi; }
	if (hi < 0) return;
	const double *p0 = vals + (size_t)r.y * nh, *p1 = vals + (size_t)r.z * nh;
	asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + lo));
	asm volatile("prefetch.global.L2 [%0];" :: "l"(p1 + lo));
	if ((hi - lo) >= 4) {
		asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + hi));
		asm volatile("prefetch.global.L2 [%0];" :: "l"(p1 + hi));
	}
}
#endif

DEV int devStep(Inst &c, const int *d, const int *bv = nullptr)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	const SimcuControl &k = c.a.ctl;
	switch (type) {
	case DK_V: case DK_I: {                     /* source_v.c:72-95, source_i.c:69-93 */
		if (d[DF_WTYPE] == 0) return 0;
		const double dc = wfValue(c, d, sb, ib), old = c.S(sb);
		if (type == DK_V) c.Bplus(d[DF_ROWR], dc - old);
		else { c.Bplus(d[DF_K], -(dc - old)); c.Bplus(d[DF_J], dc - old); }
		c.S(sb) = dc;
		return cbIsBreak(c, sb + 1, ib, dc, (type == DK_V) ? k.maxAngleV : k.maxAngleA);
	}
	case DK_VI: {                               /* vicurve.c:85-106 */
		if (d[DF_TATAB] < 0) return 0;
		const int set = c.tableSet(d[DF_SETSLOT]);
		const Table ta = tableOf(c.a, d[DF_TATAB] + set);
		int idx = c.IS(ib + 3);
		double aa, dadt;
		pwCalcValue(ta, idx, c.time, aa, dadt);
		c.IS(ib + 3) = idx;
		c.S(sb + 2) = aa;
		return cbIsBreak(c, sb + 4, ib + 1, aa, k.maxAngleA);
	}
	case DK_T: {                                /* tline.c:126-204 */
		const double tp = c.time - c.P(d[DF_PTD]);
		if (isnan(tp)) { c.err = SIMCU_ERR_NAN; return 0; }
		double Ir, Is, Vk, Vj, Vl, Vm;
		if (tp <= 0) {
			Ir = 0.0; Is = 0.0;
			Vk = c.S(sb + 2); Vj = c.S(sb + 3); Vl = c.S(sb + 4); Vm = c.S(sb + 5);
		} else {
			const int *h = d + DF_TH0;              /* R S K J L M */
#ifdef SIMCU_JIT
			const int nh = SIMCU_NH, R = c.a.histRecords;
#else
			const int nh = c.a.nh, R = c.a.histRecords;
#endif
			const double *times = c.a.Ht + (size_t)c.hslot * R;
			const double *vals = c.a.Hx + (size_t)c.hslot * R * nh;
			const int4 r = tlineLocateFn(times, R, c.hcount, c.IS(ib + 2), tp);
			if (!r.w) { c.err = SIMCU_ERR_HISTORY; return 0; }
			c.IS(ib + 2) = r.x;
			const double *recP = vals + (size_t)r.y * nh, *recN = vals + (size_t)r.z * nh;
			const double tP = times[r.y], tN = times[r.z];
			Ir = tlineInterp(recP, recN, h[0], tP, tN, tp); Is = tlineInterp(recP, recN, h[1], tP, tN, tp);
			Vk = tlineInterp(recP, recN, h[2], tP, tN, tp); Vj = tlineInterp(recP, recN, h[3], tP, tN, tp);
			Vl = tlineInterp(recP, recN, h[4], tP, tN, tp); Vm = tlineInterp(recP, recN, h[5], tP, tN, tp);
		}
		const double z0 = c.P(d[DF_PZ0]), loss = c.P(d[DF_PLOSS]);
		double Vr, Vs;
		if (loss == HUGE_VAL) {
			Vr = (Vl - Vm) + z0 * Is;
			Vs = (Vk - Vj) + z0 * Ir;
		} else {
			/* exp(-loss/2), evaluated once per run at load (same value) */
			Vr = c.S(sb + 24) * ((Vl - Vm) + z0 * Is);
			Vs = c.S(sb + 24) * ((Vk - Vj) + z0 * Ir);
		}
		c.Bplus(d[DF_ROWR], Vr - c.S(sb)); c.S(sb) = Vr;
		c.Bplus(d[DF_ROWS], Vs - c.S(sb + 1)); c.S(sb + 1) = Vs;
		int brk = cbIsBreak(c, sb + 6, ib, Vr, k.maxAngleV);
		if (!brk) brk = cbIsBreak(c, sb + 15, ib + 1, Vs, k.maxAngleV);
		return brk;
	}
	case DK_BI: case DK_BV:                     /* nonlinear_i.c:198-218, _v.c:166-188 */
		if (!d[DF_USETIME]) return 0;
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d, bv);
		return cbIsBreak(c, sb + 4, ib + 1, c.S(sb + 2),
				(type == DK_BV) ? k.maxAngleV : k.maxAngleA);
	case DK_BC:                                 /* nonlinear_c.c:172-184 */
		if (!d[DF_USETIME]) return 0;
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d, bv);
		return 0;
	default: return 0;
	}
}

DEV void devIntegrate(Inst &c, const int *d)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	double G, Ieq;
	switch (type) {
	case DK_C: case DK_BC: {                    /* capacitor.c:76-108, nonlinear_c.c:232-277 */
		const double x0 = c.X(d[DF_K]) - c.X(d[DF_J]);
		const double cap = (type == DK_C) ? c.P(d[DF_P0]) : c.S(sb + 24);
		itIntegrate(c, sb + 2, ib, x0, cap, G, Ieq);
		if (c.err) return;
		const int *s = (type == DK_C) ? d + DF_A0 : d + DF_BA0;   /* KK KJ JK JJ */
		const double dG = G - c.S(sb);
		c.Aplus(s[0], dG); c.Aplus(s[1], -dG); c.Aplus(s[2], -dG); c.Aplus(s[3], dG);
		c.S(sb) = G;
		const double dI = Ieq - c.S(sb + 1);
		c.Bplus(d[DF_K], dI); c.Bplus(d[DF_J], -dI);
		c.S(sb + 1) = Ieq;
		break;
	}
	case DK_L: {                                /* inductor.c:78-109 */
		const double x0 = c.X(d[DF_ROWR]);
		itIntegrate(c, sb + 2, ib, x0, c.P(d[DF_P0]), G, Ieq);
		if (c.err) return;
		c.Aplus(d[DF_A0 + 4], -(G - c.S(sb)));
		c.S(sb) = G;
		c.Bplus(d[DF_ROWR], -(Ieq - c.S(sb + 1)));
		c.S(sb + 1) = Ieq;
		break;
	}
	default: break;
	}
}

/* device.c:106-124: smallest positive LTE recommendation */
DEV void devMinStep(Inst &c, const int *d, double &lteStep)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	const SimcuControl &k = c.a.ctl;
	double h;
	switch (type) {
	case DK_C:
		h = itNextStep(c, sb + 2, ib, c.X(d[DF_K]) - c.X(d[DF_J]), c.P(d[DF_P0]), k.vntol);
		break;
	case DK_BC:
		h = itNextStep(c, sb + 2, ib, c.X(d[DF_K]) - c.X(d[DF_J]), c.S(sb + 24), k.captol);
		break;
	case DK_L:
		h = itNextStep(c, sb + 2, ib, c.X(d[DF_ROWR]), c.P(d[DF_P0]), k.abstol);
		break;
	default: return;
	}
	if ((h > 0.0) && (h < lteStep)) lteStep = h;
}

/* device.c:128-147: look-ahead, suggestions <= minstep are ignored */
DEV void devNextStep(Inst &c, const int *d, double &thisStep)
{
	const int type = d[DF_TYPE], ib = d[DF_IBASE];
	double loc = 0.0;
	switch (type) {
	case DK_V: case DK_I:
		if (d[DF_WTYPE] == 0) return;
		wfNextStep(c, d, ib, loc);
		break;
	case DK_VI: {                               /* vicurve.c:67-81 */
		if (d[DF_TATAB] < 0) return;
		const int set = c.tableSet(d[DF_SETSLOT]);
		const Table ta = tableOf(c.a, d[DF_TATAB] + set);
		int idx = c.IS(ib + 3);
		loc = pwGetNextX(ta, idx, c.time) - c.time;
		c.IS(ib + 3) = idx;
		break;
	}
	default: return;
	}
	if ((loc > c.a.ctl.minstep) && (loc < thisStep)) thisStep = loc;
}

#ifndef SIMCU_JIT
/* ------------------------------------------------------------------------ *
 * Numeric LU + triangular solves on the fixed pattern.
 *
 * Workspace slots: [0, nnzA) = the entries of A in CSC order, [nnzA, F) = fill,
 * [F, F+N) = right-hand side by row, overwritten by the solution.  The stream
 * (simcu_host.cpp, emitSchedule) lists, per pivot k:
 *     d, yk, nU, nL, then nU x (u, xs), then nL x (l, yr, t[nU])
 * Elimination is right-looking in pivot order with the pivot stored as its
 * reciprocal (reference: SuperLU scales the column by 1/pivot,
 * libs/superlu/SRC/dpivotL.c:170-172); the forward substitution rides along as
 * an extra column; the back substitution walks the stream backwards.
 * ------------------------------------------------------------------------ */
DEV void factorSolve(Inst &c)
{
	const SimcuArgs &a = c.a;
	double *ws = c.ws;
	const size_t st = c.wsStride;
	const int nnzA = a.nnzA, F = a.F, N = a.N;
	for (int s = 0; s < nnzA; s++) ws[s * st] = a.A[(size_t)s * c.M + c.i];
	for (int s = nnzA; s < F; s++) ws[s * st] = 0.0;
	for (int r = 0; r < N; r++) ws[(F + r) * st] = a.B[(size_t)r * c.M + c.i];

	const int *lu = a.lu;
	int pos = 0;
	bool bad = false;
	for (int k = 0; k < N; k++) {
		const int d = __ldg(&lu[pos]), yk = __ldg(&lu[pos + 1]);
		const int nU = __ldg(&lu[pos + 2]), nL = __ldg(&lu[pos + 3]);
		const int *u = lu + pos + 4;
		const int *lrow = u + 2 * nU;
		const double piv = ws[d * st];
		const double inv = 1.0 / piv;
		if (piv == 0.0 || !isfinite(inv)) bad = true;
		ws[d * st] = inv;
		const double yv = ws[yk * st];
		for (int r = 0; r < nL; r++) {
			const int *e = lrow + r * (2 + nU);
			const int ls = __ldg(&e[0]), yr = __ldg(&e[1]);
			const double l = ws[ls * st] * inv;
			ws[ls * st] = l;
			if (l != 0.0) {
				for (int j = 0; j < nU; j++) {
					const int t = __ldg(&e[2 + j]);
					ws[t * st] -= l * ws[__ldg(&u[2 * j]) * st];
				}
				ws[yr * st] -= l * yv;
			}
		}
		pos += 4 + 2 * nU + nL * (2 + nU);
	}
	/* back substitution: the stream is walked backwards via per-pivot offsets
	 * stored after the forward part */
	const int *back = lu + pos;       /* [N] start offset of pivot k */
	for (int k = N - 1; k >= 0; k--) {
		const int p = __ldg(&back[k]);
		const int d = __ldg(&lu[p]), yk = __ldg(&lu[p + 1]), nU = __ldg(&lu[p + 2]);
		const int *u = lu + p + 4;
		double acc = ws[yk * st];
		for (int j = 0; j < nU; j++)
			acc -= ws[__ldg(&u[2 * j]) * st] * ws[__ldg(&u[2 * j + 1]) * st];
		ws[yk * st] = acc * ws[d * st];
	}
	if (bad) c.err = SIMCU_ERR_SINGULAR;
}

/* simulatorSolve: core/simulator.c:45-69.  Returns the reference's count. */
DEV int newtonSolve(Inst &c, int limit, int &nfact)
{
	const SimcuArgs &a = c.a;
	int count = 0;
	factorSolve(c); nfact++;
	if (c.err) return count;
	c.xInWs = true;
	while (count++ < limit) {
		int linear = 1;
		for (int k = 0; k < a.ndev && !c.err; k++) {
			const int *d = a.dev + k * SIMCU_DEV_STRIDE;
			const int type = __ldg(&d[DF_TYPE]);
			if (type >= DK_VI) linear &= devLinearize(c, d);
		}
		if (c.err || linear) break;
		factorSolve(c); nfact++;
		if (c.err) break;
	}
	return count;
}

/* copy the accepted solution out of the workspace */
DEV void storeSolution(const Inst &c)
{
	const SimcuArgs &a = c.a;
	for (int k = 0; k < a.N; k++)
		a.X[(size_t)k * c.M + c.i] = c.ws[(size_t)__ldg(&a.xslot[k]) * c.wsStride];
}

#endif /* !SIMCU_JIT */

/* matrixRecord (core/matrix.c:372-381) as far as anything reads it: the
 * T-line history ring and the optional raw record.  rows[] = history rows. */
DEV void recordStep(Inst &c, const int *rows, double tNow)
{
	const SimcuArgs &a = c.a;
	if (a.nh > 0 && a.histRecords > 0) {
		const size_t at = (size_t)c.hslot * a.histRecords + c.hcount % a.histRecords;
		a.Ht[at] = tNow;
		double *rec = a.Hx + at * SIMCU_NH_LOOP;
		SIMCU_UNROLL
		for (int j = 0; j < SIMCU_NH_LOOP; j++) rec[j] = c.X(LDI(&rows[j]));
		c.hcount++;
	}
	if (a.rawMaxPoints > 0 && c.i >= a.rawFirst && c.i < a.rawFirst + a.rawCount) {
		const int n = c.npts;
		if (n < a.rawMaxPoints) {
			double *o = a.outRaw + ((size_t)n * (SIMCU_N_LOOP + 1)) * a.rawCount + (c.i - a.rawFirst);
			o[0] = tNow;
			SIMCU_UNROLL
			for (int k = 0; k < SIMCU_N_LOOP; k++) o[(size_t)(k + 1) * a.rawCount] = c.X(k + 1);
		}
		c.npts = n + 1;
	}
}

/* the value of unknown `row` at the step just solved / at the last accepted step */
#ifdef SIMCU_JIT
DEV double xNew(const Inst &c, int row) { return row ? c.Xv[row - 1] : 0.0; }
DEV double xOld(const Inst &c, int row) { return row ? c.Xa[row - 1] : 0.0; }
#else
DEV double xNew(const Inst &c, int row)
{
	return row ? c.ws[(size_t)__ldg(&c.a.xslot[row - 1]) * c.wsStride] : 0.0;
}
DEV double xOld(const Inst &c, int row)
{
	return row ? c.a.X[(size_t)(row - 1) * c.M + c.i] : 0.0;
}
#endif

#ifdef SIMCU_JIT
/* Device-side post-processing: one running measure of an
 * unknown over the accepted steps, so a million-instance sweep can return a
 * few numbers per instance instead of whole waveforms.  kind and row are
 * compile-time constants of the specialised kernel; v is the measure's state.
 * Crossing times interpolate linearly inside the accepted step, like the
 * reference's own result access (module/circuit.py:48-67); NaN = never. */
DEV void measureStep(const Inst &c, int kind, int row, double level, double tPrev, double tNow,
		bool first, double &v)
{
	const double xN = xNew(c, row);
	switch (kind) {
	case MS_MIN: v = (first || xN < v) ? xN : v; break;
	case MS_MAX: v = (first || xN > v) ? xN : v; break;
	case MS_FINAL: v = xN; break;
	default: {
		if (first) { v = HUGE_VAL - HUGE_VAL; break; }       /* NaN: not crossed yet */
		if (v == v) break;                                   /* first crossing only */
		const double xP = xOld(c, row);
		const bool hit = (kind == MS_TCROSS_RISE) ? (xP < level && xN >= level)
			: (xP > level && xN <= level);
		if (hit) v = tPrev + ((level - xP) / (xN - xP)) * (tNow - tPrev);
		break;
	}
	}
}
#endif

/* linear interpolation of the probes onto the tstep grid, exactly what
 * numpy.interp does on the adaptive waveform: for a grid time tg with
 * tPrev <= tg < tNow: slope * (tg - tPrev) + xPrev; a knot returns its value */
DEV void gridOutput(Inst &c, const int *rows, double tPrev, double tNow, bool first)
{
	const SimcuArgs &a = c.a;
	if (a.nprobes <= 0 || a.ngrid <= 0) return;
	int g = c.gridIdx;
	const double tstep = a.ctl.tstep, tstop = a.ctl.tstop;
	while (g < a.ngrid) {
		double tg = g * tstep;
		if (g == a.ngrid - 1 || tg > tstop) tg = tstop;
		if (tg > tNow) break;
		SIMCU_UNROLL
		for (int p = 0; p < SIMCU_NPROBES_LOOP; p++) {
			const int row = LDI(&rows[p]);
			const double xN = xNew(c, row);
			double v;
			if (first || tg == tNow) v = xN;
			else {
				const double xP = xOld(c, row);
				v = ((xN - xP) / (tNow - tPrev)) * (tg - tPrev) + xP;
			}
			a.outGrid[((size_t)g * a.nprobes + p) * a.M + c.i] = v;
		}
		g++;
	}
	c.gridIdx = g;
}

#endif
/* ---- generated for this netlist by simcu_codegen.cpp ---- */
DEV void genCalc(Inst &c, int id, int dvar, double &f, double &dv)
{
	const double m = SIMCU_GMIN;
	(void)m; (void)c;
	f = 0.0; dv = 0.0;
	switch (id) {
	default: break;
	}
}

DEV void genCalcGrad(Inst &c, int id, double *g)
{
	const double m = SIMCU_GMIN;
	(void)m; (void)c; (void)g;
	switch (id) {
	default: break;
	}
}

DEV void genLoad(Inst &c)
{
	{ const int d[] = {3,1,0,0,0,2,0,0,0,0,0,-1,-1,0,-1,3,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {1,3,0,0,0,0,0,11,2,1,4,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {6,3,0,0,0,4,5,35,3,0,-1,6,0,5,12,10,11,-1,-1,-1,13,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {6,1,3,0,0,6,7,48,7,6,-1,6,1,1,17,14,15,4,16,6,18,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {6,1,3,0,0,8,9,61,11,12,18,6,2,2,22,19,20,4,21,7,23,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {6,3,0,0,0,10,11,74,15,24,30,6,3,8,26,24,25,-1,-1,-1,27,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {1,12,0,0,0,0,0,87,19,2,28,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {0,13,12,0,0,0,0,111,20,3,32,29,31,28,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {2,3,13,0,0,14,0,111,20,4,9,33,34,35,36,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {3,12,15,0,0,16,0,135,21,5,0,-1,-1,30,38,40,41,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {0,15,17,0,0,0,0,146,23,6,37,42,39,43,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,17,0,18,0,19,20,146,23,7,8,9,44,-1,50,-1,46,-1,47,-1,54,-1,45,-1,52,55,53,-1,51,-1,0,1,2,-1,3,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,18,0,21,0,22,23,171,26,10,11,12,48,-1,60,-1,56,-1,57,-1,64,-1,49,-1,62,65,63,-1,61,-1,4,5,3,-1,6,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,21,0,24,0,25,26,196,29,13,14,15,58,-1,70,-1,66,-1,67,-1,74,-1,59,-1,72,75,73,-1,71,-1,7,8,6,-1,9,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,24,0,27,0,28,29,221,32,16,17,18,68,-1,80,-1,76,-1,77,-1,84,-1,69,-1,82,85,83,-1,81,-1,10,11,9,-1,12,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,27,0,30,0,31,32,246,35,19,20,21,78,-1,90,-1,86,-1,87,-1,94,-1,79,-1,92,95,93,-1,91,-1,13,14,12,-1,15,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,30,0,33,0,34,35,271,38,22,23,24,88,-1,100,-1,96,-1,97,-1,104,-1,89,-1,102,105,103,-1,101,-1,16,17,15,-1,18,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,33,0,36,0,37,38,296,41,25,26,27,98,-1,110,-1,106,-1,107,-1,114,-1,99,-1,112,115,113,-1,111,-1,19,20,18,-1,21,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {5,36,0,39,0,40,41,321,44,28,29,30,108,-1,120,-1,117,-1,118,-1,124,-1,109,-1,122,125,123,-1,121,-1,22,23,21,-1,24,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {3,42,0,0,0,43,0,346,47,31,0,-1,-1,126,-1,127,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {1,44,0,0,0,0,0,357,49,32,128,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {6,44,0,0,0,45,46,381,50,36,-1,6,4,129,133,131,132,-1,-1,-1,134,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {1,39,0,0,0,0,0,394,54,33,116,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {0,47,39,0,0,0,0,418,55,34,136,119,135,116,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
	{ const int d[] = {2,44,47,0,0,48,0,418,55,35,130,137,138,139,140,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devLoad(c, d, bv); }
	if (c.err) return;
}

DEV void genInitStep(Inst &c)
{
	{ const int d[] = {1,3,0,0,0,0,0,11,2,1,4,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {1,12,0,0,0,0,0,87,19,2,28,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {2,3,13,0,0,14,0,111,20,4,9,33,34,35,36,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,17,0,18,0,19,20,146,23,7,8,9,44,-1,50,-1,46,-1,47,-1,54,-1,45,-1,52,55,53,-1,51,-1,0,1,2,-1,3,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,18,0,21,0,22,23,171,26,10,11,12,48,-1,60,-1,56,-1,57,-1,64,-1,49,-1,62,65,63,-1,61,-1,4,5,3,-1,6,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,21,0,24,0,25,26,196,29,13,14,15,58,-1,70,-1,66,-1,67,-1,74,-1,59,-1,72,75,73,-1,71,-1,7,8,6,-1,9,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,24,0,27,0,28,29,221,32,16,17,18,68,-1,80,-1,76,-1,77,-1,84,-1,69,-1,82,85,83,-1,81,-1,10,11,9,-1,12,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,27,0,30,0,31,32,246,35,19,20,21,78,-1,90,-1,86,-1,87,-1,94,-1,79,-1,92,95,93,-1,91,-1,13,14,12,-1,15,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,30,0,33,0,34,35,271,38,22,23,24,88,-1,100,-1,96,-1,97,-1,104,-1,89,-1,102,105,103,-1,101,-1,16,17,15,-1,18,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,33,0,36,0,37,38,296,41,25,26,27,98,-1,110,-1,106,-1,107,-1,114,-1,99,-1,112,115,113,-1,111,-1,19,20,18,-1,21,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {5,36,0,39,0,40,41,321,44,28,29,30,108,-1,120,-1,117,-1,118,-1,124,-1,109,-1,122,125,123,-1,121,-1,22,23,21,-1,24,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {1,44,0,0,0,0,0,357,49,32,128,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {1,39,0,0,0,0,0,394,54,33,116,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
	{ const int d[] = {2,44,47,0,0,48,0,418,55,35,130,137,138,139,140,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devInitStep(c, d); }
}

DEV int genStep(Inst &c)
{
	int brk = 0;
	{ const int d[] = {6,1,3,0,0,8,9,61,11,12,18,6,2,2,22,19,20,4,21,7,23,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {6,3,0,0,0,10,11,74,15,24,30,6,3,8,26,24,25,-1,-1,-1,27,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,17,0,18,0,19,20,146,23,7,8,9,44,-1,50,-1,46,-1,47,-1,54,-1,45,-1,52,55,53,-1,51,-1,0,1,2,-1,3,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,18,0,21,0,22,23,171,26,10,11,12,48,-1,60,-1,56,-1,57,-1,64,-1,49,-1,62,65,63,-1,61,-1,4,5,3,-1,6,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,21,0,24,0,25,26,196,29,13,14,15,58,-1,70,-1,66,-1,67,-1,74,-1,59,-1,72,75,73,-1,71,-1,7,8,6,-1,9,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,24,0,27,0,28,29,221,32,16,17,18,68,-1,80,-1,76,-1,77,-1,84,-1,69,-1,82,85,83,-1,81,-1,10,11,9,-1,12,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,27,0,30,0,31,32,246,35,19,20,21,78,-1,90,-1,86,-1,87,-1,94,-1,79,-1,92,95,93,-1,91,-1,13,14,12,-1,15,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,30,0,33,0,34,35,271,38,22,23,24,88,-1,100,-1,96,-1,97,-1,104,-1,89,-1,102,105,103,-1,101,-1,16,17,15,-1,18,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,33,0,36,0,37,38,296,41,25,26,27,98,-1,110,-1,106,-1,107,-1,114,-1,99,-1,112,115,113,-1,111,-1,19,20,18,-1,21,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	{ const int d[] = {5,36,0,39,0,40,41,321,44,28,29,30,108,-1,120,-1,117,-1,118,-1,124,-1,109,-1,122,125,123,-1,121,-1,22,23,21,-1,24,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; brk |= devStep(c, d, bv); }
	if (c.err) return brk;
	return brk;
}

DEV void genIntegrate(Inst &c)
{
	itBegin(c);
	{ const int d[] = {1,3,0,0,0,0,0,11,2,1,4,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
	{ const int d[] = {1,12,0,0,0,0,0,87,19,2,28,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
	{ const int d[] = {2,3,13,0,0,14,0,111,20,4,9,33,34,35,36,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
	{ const int d[] = {1,44,0,0,0,0,0,357,49,32,128,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
	{ const int d[] = {1,39,0,0,0,0,0,394,54,33,116,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
	{ const int d[] = {2,44,47,0,0,48,0,418,55,35,130,137,138,139,140,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devIntegrate(c, d); }
	if (c.err) return;
}

DEV int genLinearize(Inst &c)
{
	int linear = 1;
	{ const int d[] = {6,3,0,0,0,4,5,35,3,0,-1,6,0,5,12,10,11,-1,-1,-1,13,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; linear &= devLinearize(c, d, bv); }
	if (c.err) return 0;
	{ const int d[] = {6,1,3,0,0,6,7,48,7,6,-1,6,1,1,17,14,15,4,16,6,18,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; linear &= devLinearize(c, d, bv); }
	if (c.err) return 0;
	{ const int d[] = {6,1,3,0,0,8,9,61,11,12,18,6,2,2,22,19,20,4,21,7,23,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; linear &= devLinearize(c, d, bv); }
	if (c.err) return 0;
	{ const int d[] = {6,3,0,0,0,10,11,74,15,24,30,6,3,8,26,24,25,-1,-1,-1,27,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; linear &= devLinearize(c, d, bv); }
	if (c.err) return 0;
	{ const int d[] = {6,44,0,0,0,45,46,381,50,36,-1,6,4,129,133,131,132,-1,-1,-1,134,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; linear &= devLinearize(c, d, bv); }
	if (c.err) return 0;
	return linear;
}

DEV void genMinStep(Inst &c, double &lteStep)
{
	{ const int d[] = {1,3,0,0,0,0,0,11,2,1,4,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
	{ const int d[] = {1,12,0,0,0,0,0,87,19,2,28,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
	{ const int d[] = {2,3,13,0,0,14,0,111,20,4,9,33,34,35,36,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
	{ const int d[] = {1,44,0,0,0,0,0,357,49,32,128,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
	{ const int d[] = {1,39,0,0,0,0,0,394,54,33,116,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
	{ const int d[] = {2,44,47,0,0,48,0,418,55,35,130,137,138,139,140,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devMinStep(c, d, lteStep); }
	if (c.err) return;
}

DEV void genNextStep(Inst &c, double &thisStep)
{
	{ const int d[] = {6,1,3,0,0,8,9,61,11,12,18,6,2,2,22,19,20,4,21,7,23,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devNextStep(c, d, thisStep); }
	{ const int d[] = {6,3,0,0,0,10,11,74,15,24,30,6,3,8,26,24,25,-1,-1,-1,27,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1}; const int bv[] = {0}; (void)bv; devNextStep(c, d, thisStep); }
}

/* operating point: 48 unknowns, 171 matrix slots (30 fill), 59 L + 61 U off-diagonal entries */
DEV void genFactorSolve0(Inst &c)
{
	bool bad = false, grow = false;
	double w0 = __longlong_as_double(0x3ff0000000000000LL);
	double w1 = __longlong_as_double(0x3ff0000000000000LL);
	double w2 = __longlong_as_double(0x3ff0000000000000LL);
	double w3 = __longlong_as_double(0x3ff0000000000000LL);
	double w4 = c.Av[4];
	double w5 = __longlong_as_double(0x3ff0000000000000LL);
	double w6 = c.Av[6];
	double w7 = c.Av[7];
	double w8 = __longlong_as_double(0x3ff0000000000000LL);
	double w9 = __longlong_as_double(0x3ff0000000000000LL);
	double w10 = __longlong_as_double(0x3ff0000000000000LL);
	double w11 = __longlong_as_double(0xbff0000000000000LL);
	double w12 = __longlong_as_double(0xbff0000000000000LL);
	double w13 = c.Av[13];
	double w14 = __longlong_as_double(0x3ff0000000000000LL);
	double w15 = __longlong_as_double(0xbff0000000000000LL);
	double w16 = c.Av[16];
	double w17 = __longlong_as_double(0xbff0000000000000LL);
	double w18 = c.Av[18];
	double w19 = __longlong_as_double(0x3ff0000000000000LL);
	double w20 = __longlong_as_double(0xbff0000000000000LL);
	double w21 = c.Av[21];
	double w22 = __longlong_as_double(0xbff0000000000000LL);
	double w23 = c.Av[23];
	double w24 = __longlong_as_double(0x3ff0000000000000LL);
	double w25 = __longlong_as_double(0xbff0000000000000LL);
	double w26 = __longlong_as_double(0xbff0000000000000LL);
	double w27 = c.Av[27];
	double w28 = c.Av[28];
	double w29 = c.Av[29];
	double w30 = __longlong_as_double(0x3ff0000000000000LL);
	double w31 = c.Av[31];
	double w32 = c.Av[32];
	double w33 = __longlong_as_double(0xbff0000000000000LL);
	double w34 = __longlong_as_double(0x3ff0000000000000LL);
	double w35 = __longlong_as_double(0xbff0000000000000LL);
	double w37 = c.Av[37];
	double w38 = __longlong_as_double(0xbff0000000000000LL);
	double w39 = c.Av[39];
	double w40 = __longlong_as_double(0x3ff0000000000000LL);
	double w41 = __longlong_as_double(0xbff0000000000000LL);
	double w42 = c.Av[42];
	double w43 = c.Av[43];
	double w44 = __longlong_as_double(0x3ff0000000000000LL);
	double w45 = __longlong_as_double(0xbff0000000000000LL);
	double w46 = __longlong_as_double(0xbff0000000000000LL);
	double w47 = __longlong_as_double(0x3ff0000000000000LL);
	double w48 = __longlong_as_double(0x3ff0000000000000LL);
	double w49 = __longlong_as_double(0xbff0000000000000LL);
	double w50 = __longlong_as_double(0x3ff0000000000000LL);
	double w51 = __longlong_as_double(0xbff0000000000000LL);
	double w52 = c.Av[52];
	double w53 = __longlong_as_double(0x3ff0000000000000LL);
	double w54 = __longlong_as_double(0x3ff0000000000000LL);
	double w55 = c.Av[55];
	double w56 = __longlong_as_double(0xbff0000000000000LL);
	double w57 = __longlong_as_double(0x3ff0000000000000LL);
	double w58 = __longlong_as_double(0x3ff0000000000000LL);
	double w59 = __longlong_as_double(0xbff0000000000000LL);
	double w60 = __longlong_as_double(0x3ff0000000000000LL);
	double w61 = __longlong_as_double(0xbff0000000000000LL);
	double w62 = c.Av[62];
	double w63 = __longlong_as_double(0x3ff0000000000000LL);
	double w64 = __longlong_as_double(0x3ff0000000000000LL);
	double w65 = c.Av[65];
	double w66 = __longlong_as_double(0xbff0000000000000LL);
	double w67 = __longlong_as_double(0x3ff0000000000000LL);
	double w68 = __longlong_as_double(0x3ff0000000000000LL);
	double w69 = __longlong_as_double(0xbff0000000000000LL);
	double w70 = __longlong_as_double(0x3ff0000000000000LL);
	double w71 = __longlong_as_double(0xbff0000000000000LL);
	double w72 = c.Av[72];
	double w73 = __longlong_as_double(0x3ff0000000000000LL);
	double w74 = __longlong_as_double(0x3ff0000000000000LL);
	double w75 = c.Av[75];
	double w76 = __longlong_as_double(0xbff0000000000000LL);
	double w77 = __longlong_as_double(0x3ff0000000000000LL);
	double w78 = __longlong_as_double(0x3ff0000000000000LL);
	double w79 = __longlong_as_double(0xbff0000000000000LL);
	double w80 = __longlong_as_double(0x3ff0000000000000LL);
	double w81 = __longlong_as_double(0xbff0000000000000LL);
	double w82 = c.Av[82];
	double w83 = __longlong_as_double(0x3ff0000000000000LL);
	double w84 = __longlong_as_double(0x3ff0000000000000LL);
	double w85 = c.Av[85];
	double w86 = __longlong_as_double(0xbff0000000000000LL);
	double w87 = __longlong_as_double(0x3ff0000000000000LL);
	double w88 = __longlong_as_double(0x3ff0000000000000LL);
	double w89 = __longlong_as_double(0xbff0000000000000LL);
	double w90 = __longlong_as_double(0x3ff0000000000000LL);
	double w91 = __longlong_as_double(0xbff0000000000000LL);
	double w92 = c.Av[92];
	double w93 = __longlong_as_double(0x3ff0000000000000LL);
	double w94 = __longlong_as_double(0x3ff0000000000000LL);
	double w95 = c.Av[95];
	double w96 = __longlong_as_double(0xbff0000000000000LL);
	double w97 = __longlong_as_double(0x3ff0000000000000LL);
	double w98 = __longlong_as_double(0x3ff0000000000000LL);
	double w99 = __longlong_as_double(0xbff0000000000000LL);
	double w100 = __longlong_as_double(0x3ff0000000000000LL);
	double w101 = __longlong_as_double(0xbff0000000000000LL);
	double w102 = c.Av[102];
	double w103 = __longlong_as_double(0x3ff0000000000000LL);
	double w104 = __longlong_as_double(0x3ff0000000000000LL);
	double w105 = c.Av[105];
	double w106 = __longlong_as_double(0xbff0000000000000LL);
	double w107 = __longlong_as_double(0x3ff0000000000000LL);
	double w108 = __longlong_as_double(0x3ff0000000000000LL);
	double w109 = __longlong_as_double(0xbff0000000000000LL);
	double w110 = __longlong_as_double(0x3ff0000000000000LL);
	double w111 = __longlong_as_double(0xbff0000000000000LL);
	double w112 = c.Av[112];
	double w113 = __longlong_as_double(0x3ff0000000000000LL);
	double w114 = __longlong_as_double(0x3ff0000000000000LL);
	double w115 = c.Av[115];
	double w116 = c.Av[116];
	double w117 = __longlong_as_double(0xbff0000000000000LL);
	double w118 = __longlong_as_double(0x3ff0000000000000LL);
	double w119 = c.Av[119];
	double w120 = __longlong_as_double(0x3ff0000000000000LL);
	double w121 = __longlong_as_double(0xbff0000000000000LL);
	double w122 = c.Av[122];
	double w123 = __longlong_as_double(0x3ff0000000000000LL);
	double w124 = __longlong_as_double(0x3ff0000000000000LL);
	double w125 = c.Av[125];
	double w126 = __longlong_as_double(0x3ff0000000000000LL);
	double w127 = __longlong_as_double(0x3ff0000000000000LL);
	double w129 = __longlong_as_double(0x3ff0000000000000LL);
	double w130 = __longlong_as_double(0x3ff0000000000000LL);
	double w131 = __longlong_as_double(0x3ff0000000000000LL);
	double w132 = __longlong_as_double(0xbff0000000000000LL);
	double w133 = __longlong_as_double(0xbff0000000000000LL);
	double w134 = c.Av[134];
	double w135 = c.Av[135];
	double w136 = c.Av[136];
	double w137 = __longlong_as_double(0xbff0000000000000LL);
	double w138 = __longlong_as_double(0x3ff0000000000000LL);
	double w139 = __longlong_as_double(0xbff0000000000000LL);
	double w141 = 0.0;
	double w142 = 0.0;
	double w143 = 0.0;
	double w144 = 0.0;
	double w145 = 0.0;
	double w146 = 0.0;
	double w147 = 0.0;
	double w148 = 0.0;
	double w149 = 0.0;
	double w150 = 0.0;
	double w151 = 0.0;
	double w152 = 0.0;
	double w153 = 0.0;
	double w154 = 0.0;
	double w155 = 0.0;
	double w156 = 0.0;
	double w157 = 0.0;
	double w158 = 0.0;
	double w159 = 0.0;
	double w160 = 0.0;
	double w161 = 0.0;
	double w162 = 0.0;
	double w163 = 0.0;
	double w164 = 0.0;
	double w165 = 0.0;
	double w166 = 0.0;
	double w167 = 0.0;
	double w168 = 0.0;
	double w169 = 0.0;
	double w170 = 0.0;
	double y0 = c.Bv[0];
	double y1 = c.Bv[1];
	double y2 = c.Bv[2];
	double y3 = c.Bv[3];
	double y4 = c.Bv[4];
	double y5 = c.Bv[5];
	double y6 = c.Bv[6];
	double y7 = c.Bv[7];
	double y8 = c.Bv[8];
	double y9 = c.Bv[9];
	double y10 = c.Bv[10];
	double y11 = c.Bv[11];
	double y12 = c.Bv[12];
	double y13 = c.Bv[13];
	double y14 = c.Bv[14];
	double y15 = c.Bv[15];
	double y16 = c.Bv[16];
	double y17 = c.Bv[17];
	double y18 = c.Bv[18];
	double y19 = c.Bv[19];
	double y20 = c.Bv[20];
	double y21 = c.Bv[21];
	double y22 = c.Bv[22];
	double y23 = c.Bv[23];
	double y24 = c.Bv[24];
	double y25 = c.Bv[25];
	double y26 = c.Bv[26];
	double y27 = c.Bv[27];
	double y28 = c.Bv[28];
	double y29 = c.Bv[29];
	double y30 = c.Bv[30];
	double y31 = c.Bv[31];
	double y32 = c.Bv[32];
	double y33 = c.Bv[33];
	double y34 = c.Bv[34];
	double y35 = c.Bv[35];
	double y36 = c.Bv[36];
	double y37 = c.Bv[37];
	double y38 = c.Bv[38];
	double y39 = c.Bv[39];
	double y40 = c.Bv[40];
	double y41 = c.Bv[41];
	double y42 = c.Bv[42];
	double y43 = c.Bv[43];
	double y44 = c.Bv[44];
	double y45 = c.Bv[45];
	double y46 = c.Bv[46];
	double y47 = c.Bv[47];
	const double i0 = 1.0 / w0; bad |= (w0 == 0.0) | !isfinite(i0);
	{ const double l = w1 * i0; grow |= (fabs(l) > SIMCU_LMAX); y5 -= l * y1; }
	{ const double l = w2 * i0; grow |= (fabs(l) > SIMCU_LMAX); y7 -= l * y1; }
	const double i1 = 1.0 / w3; bad |= (w3 == 0.0) | !isfinite(i1);
	const double i2 = 1.0 / w15; bad |= (w15 == 0.0) | !isfinite(i2);
	const double i3 = 1.0 / w17; bad |= (w17 == 0.0) | !isfinite(i3);
	{ const double l = w16 * i3; grow |= (fabs(l) > SIMCU_LMAX); y2 -= l * y5; }
	const double i4 = 1.0 / w20; bad |= (w20 == 0.0) | !isfinite(i4);
	const double i5 = 1.0 / w22; bad |= (w22 == 0.0) | !isfinite(i5);
	{ const double l = w21 * i5; grow |= (fabs(l) > SIMCU_LMAX); y2 -= l * y7; }
	const double i6 = 1.0 / w126; bad |= (w126 == 0.0) | !isfinite(i6);
	const double i7 = 1.0 / w127; bad |= (w127 == 0.0) | !isfinite(i7);
	const double i8 = 1.0 / w11; bad |= (w11 == 0.0) | !isfinite(i8);
	{ const double l = w10 * i8; grow |= (fabs(l) > SIMCU_LMAX); w141 -= l * w13; y2 -= l * y4; }
	const double i9 = 1.0 / w12; bad |= (w12 == 0.0) | !isfinite(i9);
	{ const double l = w141 * i9; grow |= (fabs(l) > SIMCU_LMAX); w4 -= l * w5; y2 -= l * y3; }
	const double i10 = 1.0 / w25; bad |= (w25 == 0.0) | !isfinite(i10);
	{ const double l = w24 * i10; grow |= (fabs(l) > SIMCU_LMAX); w142 -= l * w27; y2 -= l * y10; }
	const double i11 = 1.0 / w26; bad |= (w26 == 0.0) | !isfinite(i11);
	{ const double l = w142 * i11; grow |= (fabs(l) > SIMCU_LMAX); w4 -= l * w8; y2 -= l * y9; }
	const double i12 = 1.0 / w9; bad |= (w9 == 0.0) | !isfinite(i12);
	{ const double l = w4 * i12; grow |= (fabs(l) > SIMCU_LMAX); w143 -= l * w33; y2 -= l * y13; }
	const double i13 = 1.0 / w34; bad |= (w34 == 0.0) | !isfinite(i13);
	{ const double l = w35 * i13; grow |= (fabs(l) > SIMCU_LMAX); w32 -= l * w143; y12 -= l * y2; }
	const double i14 = 1.0 / w32; bad |= (w32 == 0.0) | !isfinite(i14);
	{ const double l = w31 * i14; grow |= (fabs(l) > SIMCU_LMAX); w28 -= l * w29; y11 -= l * y12; }
	const double i15 = 1.0 / w30; bad |= (w30 == 0.0) | !isfinite(i15);
	{ const double l = w28 * i15; grow |= (fabs(l) > SIMCU_LMAX); w144 -= l * w38; y11 -= l * y15; }
	const double i16 = 1.0 / w40; bad |= (w40 == 0.0) | !isfinite(i16);
	{ const double l = w41 * i16; grow |= (fabs(l) > SIMCU_LMAX); w37 -= l * w144; y14 -= l * y11; }
	const double i17 = 1.0 / w37; bad |= (w37 == 0.0) | !isfinite(i17);
	{ const double l = w39 * i17; grow |= (fabs(l) > SIMCU_LMAX); w43 -= l * w42; y16 -= l * y14; }
	const double i18 = 1.0 / w129; bad |= (w129 == 0.0) | !isfinite(i18);
	{ const double l = w130 * i18; grow |= (fabs(l) > SIMCU_LMAX); w145 -= l * w133; y47 -= l * y44; }
	const double i19 = 1.0 / w131; bad |= (w131 == 0.0) | !isfinite(i19);
	{ const double l = w132 * i19; grow |= (fabs(l) > SIMCU_LMAX); w146 -= l * w138; y45 -= l * y43; }
	const double i20 = 1.0 / w145; bad |= (w145 == 0.0) | !isfinite(i20);
	{ const double l = w134 * i20; grow |= (fabs(l) > SIMCU_LMAX); w147 -= l * w137; y45 -= l * y47; }
	const double i21 = 1.0 / w146; bad |= (w146 == 0.0) | !isfinite(i21);
	{ const double l = w139 * i21; grow |= (fabs(l) > SIMCU_LMAX); w136 -= l * w147; y46 -= l * y45; }
	const double i22 = 1.0 / w136; bad |= (w136 == 0.0) | !isfinite(i22);
	{ const double l = w135 * i22; grow |= (fabs(l) > SIMCU_LMAX); w116 -= l * w119; y38 -= l * y46; }
	const double i23 = 1.0 / w44; bad |= (w44 == 0.0) | !isfinite(i23);
	{ const double l = w43 * i23; grow |= (fabs(l) > SIMCU_LMAX); w50 -= l * w52; w148 -= l * w46; y16 -= l * y18; }
	{ const double l = w45 * i23; grow |= (fabs(l) > SIMCU_LMAX); w149 -= l * w52; w47 -= l * w46; y19 -= l * y18; }
	const double i24 = 1.0 / w149; bad |= (w149 == 0.0) | !isfinite(i24);
	{ const double l = w50 * i24; grow |= (fabs(l) > SIMCU_LMAX); w53 -= l * w55; w148 -= l * w47; y16 -= l * y19; }
	{ const double l = w51 * i24; grow |= (fabs(l) > SIMCU_LMAX); w54 -= l * w55; w150 -= l * w47; y17 -= l * y19; }
	const double i25 = 1.0 / w53; bad |= (w53 == 0.0) | !isfinite(i25);
	{ const double l = w54 * i25; grow |= (fabs(l) > SIMCU_LMAX); w150 -= l * w148; y17 -= l * y16; }
	const double i26 = 1.0 / w48; bad |= (w48 == 0.0) | !isfinite(i26);
	{ const double l = w150 * i26; grow |= (fabs(l) > SIMCU_LMAX); w60 -= l * w62; w151 -= l * w56; y17 -= l * y21; }
	{ const double l = w49 * i26; grow |= (fabs(l) > SIMCU_LMAX); w152 -= l * w62; w57 -= l * w56; y22 -= l * y21; }
	const double i27 = 1.0 / w152; bad |= (w152 == 0.0) | !isfinite(i27);
	{ const double l = w60 * i27; grow |= (fabs(l) > SIMCU_LMAX); w63 -= l * w65; w151 -= l * w57; y17 -= l * y22; }
	{ const double l = w61 * i27; grow |= (fabs(l) > SIMCU_LMAX); w64 -= l * w65; w153 -= l * w57; y20 -= l * y22; }
	const double i28 = 1.0 / w63; bad |= (w63 == 0.0) | !isfinite(i28);
	{ const double l = w64 * i28; grow |= (fabs(l) > SIMCU_LMAX); w153 -= l * w151; y20 -= l * y17; }
	const double i29 = 1.0 / w58; bad |= (w58 == 0.0) | !isfinite(i29);
	{ const double l = w153 * i29; grow |= (fabs(l) > SIMCU_LMAX); w70 -= l * w72; w154 -= l * w66; y20 -= l * y24; }
	{ const double l = w59 * i29; grow |= (fabs(l) > SIMCU_LMAX); w155 -= l * w72; w67 -= l * w66; y25 -= l * y24; }
	const double i30 = 1.0 / w155; bad |= (w155 == 0.0) | !isfinite(i30);
	{ const double l = w70 * i30; grow |= (fabs(l) > SIMCU_LMAX); w73 -= l * w75; w154 -= l * w67; y20 -= l * y25; }
	{ const double l = w71 * i30; grow |= (fabs(l) > SIMCU_LMAX); w74 -= l * w75; w156 -= l * w67; y23 -= l * y25; }
	const double i31 = 1.0 / w73; bad |= (w73 == 0.0) | !isfinite(i31);
	{ const double l = w74 * i31; grow |= (fabs(l) > SIMCU_LMAX); w156 -= l * w154; y23 -= l * y20; }
	const double i32 = 1.0 / w68; bad |= (w68 == 0.0) | !isfinite(i32);
	{ const double l = w156 * i32; grow |= (fabs(l) > SIMCU_LMAX); w80 -= l * w82; w157 -= l * w76; y23 -= l * y27; }
	{ const double l = w69 * i32; grow |= (fabs(l) > SIMCU_LMAX); w158 -= l * w82; w77 -= l * w76; y28 -= l * y27; }
	const double i33 = 1.0 / w158; bad |= (w158 == 0.0) | !isfinite(i33);
	{ const double l = w80 * i33; grow |= (fabs(l) > SIMCU_LMAX); w83 -= l * w85; w157 -= l * w77; y23 -= l * y28; }
	{ const double l = w81 * i33; grow |= (fabs(l) > SIMCU_LMAX); w84 -= l * w85; w159 -= l * w77; y26 -= l * y28; }
	const double i34 = 1.0 / w83; bad |= (w83 == 0.0) | !isfinite(i34);
	{ const double l = w84 * i34; grow |= (fabs(l) > SIMCU_LMAX); w159 -= l * w157; y26 -= l * y23; }
	const double i35 = 1.0 / w78; bad |= (w78 == 0.0) | !isfinite(i35);
	{ const double l = w159 * i35; grow |= (fabs(l) > SIMCU_LMAX); w90 -= l * w92; w160 -= l * w86; y26 -= l * y30; }
	{ const double l = w79 * i35; grow |= (fabs(l) > SIMCU_LMAX); w161 -= l * w92; w87 -= l * w86; y31 -= l * y30; }
	const double i36 = 1.0 / w161; bad |= (w161 == 0.0) | !isfinite(i36);
	{ const double l = w90 * i36; grow |= (fabs(l) > SIMCU_LMAX); w93 -= l * w95; w160 -= l * w87; y26 -= l * y31; }
	{ const double l = w91 * i36; grow |= (fabs(l) > SIMCU_LMAX); w94 -= l * w95; w162 -= l * w87; y29 -= l * y31; }
	const double i37 = 1.0 / w93; bad |= (w93 == 0.0) | !isfinite(i37);
	{ const double l = w94 * i37; grow |= (fabs(l) > SIMCU_LMAX); w162 -= l * w160; y29 -= l * y26; }
	const double i38 = 1.0 / w88; bad |= (w88 == 0.0) | !isfinite(i38);
	{ const double l = w162 * i38; grow |= (fabs(l) > SIMCU_LMAX); w100 -= l * w102; w163 -= l * w96; y29 -= l * y33; }
	{ const double l = w89 * i38; grow |= (fabs(l) > SIMCU_LMAX); w164 -= l * w102; w97 -= l * w96; y34 -= l * y33; }
	const double i39 = 1.0 / w164; bad |= (w164 == 0.0) | !isfinite(i39);
	{ const double l = w100 * i39; grow |= (fabs(l) > SIMCU_LMAX); w103 -= l * w105; w163 -= l * w97; y29 -= l * y34; }
	{ const double l = w101 * i39; grow |= (fabs(l) > SIMCU_LMAX); w104 -= l * w105; w165 -= l * w97; y32 -= l * y34; }
	const double i40 = 1.0 / w103; bad |= (w103 == 0.0) | !isfinite(i40);
	{ const double l = w104 * i40; grow |= (fabs(l) > SIMCU_LMAX); w165 -= l * w163; y32 -= l * y29; }
	const double i41 = 1.0 / w98; bad |= (w98 == 0.0) | !isfinite(i41);
	{ const double l = w165 * i41; grow |= (fabs(l) > SIMCU_LMAX); w110 -= l * w112; w166 -= l * w106; y32 -= l * y36; }
	{ const double l = w99 * i41; grow |= (fabs(l) > SIMCU_LMAX); w167 -= l * w112; w107 -= l * w106; y37 -= l * y36; }
	const double i42 = 1.0 / w167; bad |= (w167 == 0.0) | !isfinite(i42);
	{ const double l = w110 * i42; grow |= (fabs(l) > SIMCU_LMAX); w113 -= l * w115; w166 -= l * w107; y32 -= l * y37; }
	{ const double l = w111 * i42; grow |= (fabs(l) > SIMCU_LMAX); w114 -= l * w115; w168 -= l * w107; y35 -= l * y37; }
	const double i43 = 1.0 / w113; bad |= (w113 == 0.0) | !isfinite(i43);
	{ const double l = w114 * i43; grow |= (fabs(l) > SIMCU_LMAX); w168 -= l * w166; y35 -= l * y32; }
	const double i44 = 1.0 / w108; bad |= (w108 == 0.0) | !isfinite(i44);
	{ const double l = w168 * i44; grow |= (fabs(l) > SIMCU_LMAX); w169 -= l * w117; w120 -= l * w122; y35 -= l * y39; }
	{ const double l = w109 * i44; grow |= (fabs(l) > SIMCU_LMAX); w118 -= l * w117; w170 -= l * w122; y40 -= l * y39; }
	const double i45 = 1.0 / w169; bad |= (w169 == 0.0) | !isfinite(i45);
	{ const double l = w116 * i45; grow |= (fabs(l) > SIMCU_LMAX); w121 -= l * w120; w124 -= l * w123; y38 -= l * y35; }
	{ const double l = w118 * i45; grow |= (fabs(l) > SIMCU_LMAX); w170 -= l * w120; w125 -= l * w123; y40 -= l * y35; }
	const double i46 = 1.0 / w170; bad |= (w170 == 0.0) | !isfinite(i46);
	{ const double l = w121 * i46; grow |= (fabs(l) > SIMCU_LMAX); w124 -= l * w125; y38 -= l * y40; }
	const double i47 = 1.0 / w124; bad |= (w124 == 0.0) | !isfinite(i47);
	{ double acc = y38; y38 = acc * i47; }
	{ double acc = y40; acc -= w125 * y38; y40 = acc * i46; }
	{ double acc = y35; acc -= w120 * y40; acc -= w123 * y38; y35 = acc * i45; }
	{ double acc = y39; acc -= w117 * y35; acc -= w122 * y40; y39 = acc * i44; }
	{ double acc = y32; acc -= w166 * y39; y32 = acc * i43; }
	{ double acc = y37; acc -= w115 * y32; acc -= w107 * y39; y37 = acc * i42; }
	{ double acc = y36; acc -= w112 * y37; acc -= w106 * y39; y36 = acc * i41; }
	{ double acc = y29; acc -= w163 * y36; y29 = acc * i40; }
	{ double acc = y34; acc -= w105 * y29; acc -= w97 * y36; y34 = acc * i39; }
	{ double acc = y33; acc -= w102 * y34; acc -= w96 * y36; y33 = acc * i38; }
	{ double acc = y26; acc -= w160 * y33; y26 = acc * i37; }
	{ double acc = y31; acc -= w95 * y26; acc -= w87 * y33; y31 = acc * i36; }
	{ double acc = y30; acc -= w92 * y31; acc -= w86 * y33; y30 = acc * i35; }
	{ double acc = y23; acc -= w157 * y30; y23 = acc * i34; }
	{ double acc = y28; acc -= w85 * y23; acc -= w77 * y30; y28 = acc * i33; }
	{ double acc = y27; acc -= w82 * y28; acc -= w76 * y30; y27 = acc * i32; }
	{ double acc = y20; acc -= w154 * y27; y20 = acc * i31; }
	{ double acc = y25; acc -= w75 * y20; acc -= w67 * y27; y25 = acc * i30; }
	{ double acc = y24; acc -= w72 * y25; acc -= w66 * y27; y24 = acc * i29; }
	{ double acc = y17; acc -= w151 * y24; y17 = acc * i28; }
	{ double acc = y22; acc -= w65 * y17; acc -= w57 * y24; y22 = acc * i27; }
	{ double acc = y21; acc -= w62 * y22; acc -= w56 * y24; y21 = acc * i26; }
	{ double acc = y16; acc -= w148 * y21; y16 = acc * i25; }
	{ double acc = y19; acc -= w55 * y16; acc -= w47 * y21; y19 = acc * i24; }
	{ double acc = y18; acc -= w52 * y19; acc -= w46 * y21; y18 = acc * i23; }
	{ double acc = y46; acc -= w119 * y35; y46 = acc * i22; }
	{ double acc = y45; acc -= w147 * y46; y45 = acc * i21; }
	{ double acc = y47; acc -= w137 * y46; y47 = acc * i20; }
	{ double acc = y43; acc -= w138 * y45; y43 = acc * i19; }
	{ double acc = y44; acc -= w133 * y47; y44 = acc * i18; }
	{ double acc = y14; acc -= w42 * y18; y14 = acc * i17; }
	{ double acc = y11; acc -= w144 * y14; y11 = acc * i16; }
	{ double acc = y15; acc -= w38 * y14; y15 = acc * i15; }
	{ double acc = y12; acc -= w29 * y15; y12 = acc * i14; }
	{ double acc = y2; acc -= w143 * y12; y2 = acc * i13; }
	{ double acc = y13; acc -= w33 * y12; y13 = acc * i12; }
	{ double acc = y9; acc -= w8 * y13; y9 = acc * i11; }
	{ double acc = y10; acc -= w27 * y9; y10 = acc * i10; }
	{ double acc = y3; acc -= w5 * y13; y3 = acc * i9; }
	{ double acc = y4; acc -= w13 * y3; y4 = acc * i8; }
	{ double acc = y41; y41 = acc * i7; }
	{ double acc = y42; y42 = acc * i6; }
	{ double acc = y7; y7 = acc * i5; }
	{ double acc = y8; acc -= w23 * y7; acc -= w7 * y13; y8 = acc * i4; }
	{ double acc = y5; y5 = acc * i3; }
	{ double acc = y6; acc -= w18 * y5; acc -= w6 * y13; y6 = acc * i2; }
	{ double acc = y0; acc -= w14 * y6; acc -= w19 * y8; y0 = acc * i1; }
	{ double acc = y1; y1 = acc * i0; }
	c.Xv[0] = y1;
	c.Xv[1] = y0;
	c.Xv[2] = y13;
	c.Xv[3] = y4;
	c.Xv[4] = y3;
	c.Xv[5] = y6;
	c.Xv[6] = y5;
	c.Xv[7] = y8;
	c.Xv[8] = y7;
	c.Xv[9] = y10;
	c.Xv[10] = y9;
	c.Xv[11] = y15;
	c.Xv[12] = y12;
	c.Xv[13] = y2;
	c.Xv[14] = y14;
	c.Xv[15] = y11;
	c.Xv[16] = y18;
	c.Xv[17] = y21;
	c.Xv[18] = y19;
	c.Xv[19] = y16;
	c.Xv[20] = y24;
	c.Xv[21] = y22;
	c.Xv[22] = y17;
	c.Xv[23] = y27;
	c.Xv[24] = y25;
	c.Xv[25] = y20;
	c.Xv[26] = y30;
	c.Xv[27] = y28;
	c.Xv[28] = y23;
	c.Xv[29] = y33;
	c.Xv[30] = y31;
	c.Xv[31] = y26;
	c.Xv[32] = y36;
	c.Xv[33] = y34;
	c.Xv[34] = y29;
	c.Xv[35] = y39;
	c.Xv[36] = y37;
	c.Xv[37] = y32;
	c.Xv[38] = y35;
	c.Xv[39] = y40;
	c.Xv[40] = y38;
	c.Xv[41] = y42;
	c.Xv[42] = y41;
	c.Xv[43] = y44;
	c.Xv[44] = y43;
	c.Xv[45] = y47;
	c.Xv[46] = y46;
	c.Xv[47] = y45;
	if (bad) c.err = SIMCU_ERR_SINGULAR;
	if (grow) c.growth++;
}

/* transient: 48 unknowns, 152 matrix slots (11 fill), 35 L + 37 U off-diagonal entries */
DEV void genFactorSolve1(Inst &c)
{
	bool bad = false, grow = false;
	double w0 = __longlong_as_double(0x3ff0000000000000LL);
	double w1 = __longlong_as_double(0x3ff0000000000000LL);
	double w2 = __longlong_as_double(0x3ff0000000000000LL);
	double w3 = __longlong_as_double(0x3ff0000000000000LL);
	double w4 = c.Av[4];
	double w5 = __longlong_as_double(0x3ff0000000000000LL);
	double w6 = c.Av[6];
	double w7 = c.Av[7];
	double w8 = __longlong_as_double(0x3ff0000000000000LL);
	double w9 = __longlong_as_double(0x3ff0000000000000LL);
	double w10 = __longlong_as_double(0x3ff0000000000000LL);
	double w11 = __longlong_as_double(0xbff0000000000000LL);
	double w12 = __longlong_as_double(0xbff0000000000000LL);
	double w13 = c.Av[13];
	double w14 = __longlong_as_double(0x3ff0000000000000LL);
	double w15 = __longlong_as_double(0xbff0000000000000LL);
	double w16 = c.Av[16];
	double w17 = __longlong_as_double(0xbff0000000000000LL);
	double w18 = c.Av[18];
	double w19 = __longlong_as_double(0x3ff0000000000000LL);
	double w20 = __longlong_as_double(0xbff0000000000000LL);
	double w21 = c.Av[21];
	double w22 = __longlong_as_double(0xbff0000000000000LL);
	double w23 = c.Av[23];
	double w24 = __longlong_as_double(0x3ff0000000000000LL);
	double w25 = __longlong_as_double(0xbff0000000000000LL);
	double w26 = __longlong_as_double(0xbff0000000000000LL);
	double w27 = c.Av[27];
	double w28 = c.Av[28];
	double w29 = c.Av[29];
	double w30 = __longlong_as_double(0x3ff0000000000000LL);
	double w31 = c.Av[31];
	double w32 = c.Av[32];
	double w33 = __longlong_as_double(0xbff0000000000000LL);
	double w34 = __longlong_as_double(0x3ff0000000000000LL);
	double w35 = __longlong_as_double(0xbff0000000000000LL);
	double w36 = c.Av[36];
	double w37 = c.Av[37];
	double w38 = __longlong_as_double(0xbff0000000000000LL);
	double w39 = c.Av[39];
	double w40 = __longlong_as_double(0x3ff0000000000000LL);
	double w41 = __longlong_as_double(0xbff0000000000000LL);
	double w42 = c.Av[42];
	double w43 = c.Av[43];
	double w44 = __longlong_as_double(0x3ff0000000000000LL);
	double w47 = __longlong_as_double(0x3ff0000000000000LL);
	double w48 = __longlong_as_double(0x3ff0000000000000LL);
	double w50 = __longlong_as_double(0x3ff0000000000000LL);
	double w52 = c.Av[52];
	double w54 = __longlong_as_double(0x3ff0000000000000LL);
	double w55 = c.Av[55];
	double w57 = __longlong_as_double(0x3ff0000000000000LL);
	double w58 = __longlong_as_double(0x3ff0000000000000LL);
	double w60 = __longlong_as_double(0x3ff0000000000000LL);
	double w62 = c.Av[62];
	double w64 = __longlong_as_double(0x3ff0000000000000LL);
	double w65 = c.Av[65];
	double w67 = __longlong_as_double(0x3ff0000000000000LL);
	double w68 = __longlong_as_double(0x3ff0000000000000LL);
	double w70 = __longlong_as_double(0x3ff0000000000000LL);
	double w72 = c.Av[72];
	double w74 = __longlong_as_double(0x3ff0000000000000LL);
	double w75 = c.Av[75];
	double w77 = __longlong_as_double(0x3ff0000000000000LL);
	double w78 = __longlong_as_double(0x3ff0000000000000LL);
	double w80 = __longlong_as_double(0x3ff0000000000000LL);
	double w82 = c.Av[82];
	double w84 = __longlong_as_double(0x3ff0000000000000LL);
	double w85 = c.Av[85];
	double w87 = __longlong_as_double(0x3ff0000000000000LL);
	double w88 = __longlong_as_double(0x3ff0000000000000LL);
	double w90 = __longlong_as_double(0x3ff0000000000000LL);
	double w92 = c.Av[92];
	double w94 = __longlong_as_double(0x3ff0000000000000LL);
	double w95 = c.Av[95];
	double w97 = __longlong_as_double(0x3ff0000000000000LL);
	double w98 = __longlong_as_double(0x3ff0000000000000LL);
	double w100 = __longlong_as_double(0x3ff0000000000000LL);
	double w102 = c.Av[102];
	double w104 = __longlong_as_double(0x3ff0000000000000LL);
	double w105 = c.Av[105];
	double w107 = __longlong_as_double(0x3ff0000000000000LL);
	double w108 = __longlong_as_double(0x3ff0000000000000LL);
	double w110 = __longlong_as_double(0x3ff0000000000000LL);
	double w112 = c.Av[112];
	double w114 = __longlong_as_double(0x3ff0000000000000LL);
	double w115 = c.Av[115];
	double w116 = c.Av[116];
	double w118 = __longlong_as_double(0x3ff0000000000000LL);
	double w119 = c.Av[119];
	double w120 = __longlong_as_double(0x3ff0000000000000LL);
	double w122 = c.Av[122];
	double w124 = __longlong_as_double(0x3ff0000000000000LL);
	double w125 = c.Av[125];
	double w126 = __longlong_as_double(0x3ff0000000000000LL);
	double w127 = __longlong_as_double(0x3ff0000000000000LL);
	double w128 = c.Av[128];
	double w129 = __longlong_as_double(0x3ff0000000000000LL);
	double w130 = __longlong_as_double(0x3ff0000000000000LL);
	double w131 = __longlong_as_double(0x3ff0000000000000LL);
	double w132 = __longlong_as_double(0xbff0000000000000LL);
	double w133 = __longlong_as_double(0xbff0000000000000LL);
	double w134 = c.Av[134];
	double w135 = c.Av[135];
	double w136 = c.Av[136];
	double w137 = __longlong_as_double(0xbff0000000000000LL);
	double w138 = __longlong_as_double(0x3ff0000000000000LL);
	double w139 = __longlong_as_double(0xbff0000000000000LL);
	double w140 = c.Av[140];
	double w141 = 0.0;
	double w142 = 0.0;
	double w143 = 0.0;
	double w144 = 0.0;
	double w145 = 0.0;
	double w146 = 0.0;
	double w147 = 0.0;
	double w148 = 0.0;
	double w149 = 0.0;
	double w150 = 0.0;
	double w151 = 0.0;
	double y0 = c.Bv[0];
	double y1 = c.Bv[1];
	double y2 = c.Bv[2];
	double y3 = c.Bv[3];
	double y4 = c.Bv[4];
	double y5 = c.Bv[5];
	double y6 = c.Bv[6];
	double y7 = c.Bv[7];
	double y8 = c.Bv[8];
	double y9 = c.Bv[9];
	double y10 = c.Bv[10];
	double y11 = c.Bv[11];
	double y12 = c.Bv[12];
	double y13 = c.Bv[13];
	double y14 = c.Bv[14];
	double y15 = c.Bv[15];
	double y16 = c.Bv[16];
	double y17 = c.Bv[17];
	double y18 = c.Bv[18];
	double y19 = c.Bv[19];
	double y20 = c.Bv[20];
	double y21 = c.Bv[21];
	double y22 = c.Bv[22];
	double y23 = c.Bv[23];
	double y24 = c.Bv[24];
	double y25 = c.Bv[25];
	double y26 = c.Bv[26];
	double y27 = c.Bv[27];
	double y28 = c.Bv[28];
	double y29 = c.Bv[29];
	double y30 = c.Bv[30];
	double y31 = c.Bv[31];
	double y32 = c.Bv[32];
	double y33 = c.Bv[33];
	double y34 = c.Bv[34];
	double y35 = c.Bv[35];
	double y36 = c.Bv[36];
	double y37 = c.Bv[37];
	double y38 = c.Bv[38];
	double y39 = c.Bv[39];
	double y40 = c.Bv[40];
	double y41 = c.Bv[41];
	double y42 = c.Bv[42];
	double y43 = c.Bv[43];
	double y44 = c.Bv[44];
	double y45 = c.Bv[45];
	double y46 = c.Bv[46];
	double y47 = c.Bv[47];
	const double i0 = 1.0 / w0; bad |= (w0 == 0.0) | !isfinite(i0);
	{ const double l = w1 * i0; grow |= (fabs(l) > SIMCU_LMAX); y5 -= l * y1; }
	{ const double l = w2 * i0; grow |= (fabs(l) > SIMCU_LMAX); y7 -= l * y1; }
	const double i1 = 1.0 / w3; bad |= (w3 == 0.0) | !isfinite(i1);
	const double i2 = 1.0 / w15; bad |= (w15 == 0.0) | !isfinite(i2);
	const double i3 = 1.0 / w17; bad |= (w17 == 0.0) | !isfinite(i3);
	{ const double l = w16 * i3; grow |= (fabs(l) > SIMCU_LMAX); y2 -= l * y5; }
	const double i4 = 1.0 / w20; bad |= (w20 == 0.0) | !isfinite(i4);
	const double i5 = 1.0 / w22; bad |= (w22 == 0.0) | !isfinite(i5);
	{ const double l = w21 * i5; grow |= (fabs(l) > SIMCU_LMAX); y2 -= l * y7; }
	const double i6 = 1.0 / w126; bad |= (w126 == 0.0) | !isfinite(i6);
	const double i7 = 1.0 / w127; bad |= (w127 == 0.0) | !isfinite(i7);
	const double i8 = 1.0 / w11; bad |= (w11 == 0.0) | !isfinite(i8);
	{ const double l = w10 * i8; grow |= (fabs(l) > SIMCU_LMAX); w141 -= l * w13; y2 -= l * y4; }
	const double i9 = 1.0 / w12; bad |= (w12 == 0.0) | !isfinite(i9);
	{ const double l = w141 * i9; grow |= (fabs(l) > SIMCU_LMAX); w4 -= l * w5; y2 -= l * y3; }
	const double i10 = 1.0 / w25; bad |= (w25 == 0.0) | !isfinite(i10);
	{ const double l = w24 * i10; grow |= (fabs(l) > SIMCU_LMAX); w142 -= l * w27; y2 -= l * y10; }
	const double i11 = 1.0 / w26; bad |= (w26 == 0.0) | !isfinite(i11);
	{ const double l = w142 * i11; grow |= (fabs(l) > SIMCU_LMAX); w4 -= l * w8; y2 -= l * y9; }
	const double i12 = 1.0 / w4; bad |= (w4 == 0.0) | !isfinite(i12);
	{ const double l = w9 * i12; grow |= (fabs(l) > SIMCU_LMAX); w36 -= l * w34; y13 -= l * y2; }
	const double i13 = 1.0 / w36; bad |= (w36 == 0.0) | !isfinite(i13);
	{ const double l = w35 * i13; grow |= (fabs(l) > SIMCU_LMAX); w32 -= l * w33; y12 -= l * y13; }
	const double i14 = 1.0 / w32; bad |= (w32 == 0.0) | !isfinite(i14);
	{ const double l = w31 * i14; grow |= (fabs(l) > SIMCU_LMAX); w28 -= l * w29; y11 -= l * y12; }
	const double i15 = 1.0 / w28; bad |= (w28 == 0.0) | !isfinite(i15);
	{ const double l = w30 * i15; grow |= (fabs(l) > SIMCU_LMAX); w143 -= l * w40; y15 -= l * y11; }
	const double i16 = 1.0 / w47; bad |= (w47 == 0.0) | !isfinite(i16);
	{ const double l = w48 * i16; grow |= (fabs(l) > SIMCU_LMAX); w144 -= l * w55; y21 -= l * y19; }
	const double i17 = 1.0 / w52; bad |= (w52 == 0.0) | !isfinite(i17);
	{ const double l = w50 * i17; grow |= (fabs(l) > SIMCU_LMAX); w43 -= l * w44; y16 -= l * y18; }
	const double i18 = 1.0 / w43; bad |= (w43 == 0.0) | !isfinite(i18);
	{ const double l = w42 * i18; grow |= (fabs(l) > SIMCU_LMAX); w37 -= l * w39; y14 -= l * y16; }
	const double i19 = 1.0 / w38; bad |= (w38 == 0.0) | !isfinite(i19);
	{ const double l = w37 * i19; grow |= (fabs(l) > SIMCU_LMAX); w41 -= l * w143; y14 -= l * y15; }
	const double i20 = 1.0 / w41; bad |= (w41 == 0.0) | !isfinite(i20);
	const double i21 = 1.0 / w144; bad |= (w144 == 0.0) | !isfinite(i21);
	{ const double l = w54 * i21; grow |= (fabs(l) > SIMCU_LMAX); w60 -= l * w62; y17 -= l * y21; }
	const double i22 = 1.0 / w60; bad |= (w60 == 0.0) | !isfinite(i22);
	const double i23 = 1.0 / w57; bad |= (w57 == 0.0) | !isfinite(i23);
	{ const double l = w58 * i23; grow |= (fabs(l) > SIMCU_LMAX); w145 -= l * w65; y24 -= l * y22; }
	const double i24 = 1.0 / w145; bad |= (w145 == 0.0) | !isfinite(i24);
	{ const double l = w64 * i24; grow |= (fabs(l) > SIMCU_LMAX); w70 -= l * w72; y20 -= l * y24; }
	const double i25 = 1.0 / w70; bad |= (w70 == 0.0) | !isfinite(i25);
	const double i26 = 1.0 / w67; bad |= (w67 == 0.0) | !isfinite(i26);
	{ const double l = w68 * i26; grow |= (fabs(l) > SIMCU_LMAX); w146 -= l * w75; y27 -= l * y25; }
	const double i27 = 1.0 / w146; bad |= (w146 == 0.0) | !isfinite(i27);
	{ const double l = w74 * i27; grow |= (fabs(l) > SIMCU_LMAX); w80 -= l * w82; y23 -= l * y27; }
	const double i28 = 1.0 / w80; bad |= (w80 == 0.0) | !isfinite(i28);
	const double i29 = 1.0 / w77; bad |= (w77 == 0.0) | !isfinite(i29);
	{ const double l = w78 * i29; grow |= (fabs(l) > SIMCU_LMAX); w147 -= l * w85; y30 -= l * y28; }
	const double i30 = 1.0 / w147; bad |= (w147 == 0.0) | !isfinite(i30);
	{ const double l = w84 * i30; grow |= (fabs(l) > SIMCU_LMAX); w90 -= l * w92; y26 -= l * y30; }
	const double i31 = 1.0 / w90; bad |= (w90 == 0.0) | !isfinite(i31);
	const double i32 = 1.0 / w87; bad |= (w87 == 0.0) | !isfinite(i32);
	{ const double l = w88 * i32; grow |= (fabs(l) > SIMCU_LMAX); w148 -= l * w95; y33 -= l * y31; }
	const double i33 = 1.0 / w148; bad |= (w148 == 0.0) | !isfinite(i33);
	{ const double l = w94 * i33; grow |= (fabs(l) > SIMCU_LMAX); w100 -= l * w102; y29 -= l * y33; }
	const double i34 = 1.0 / w100; bad |= (w100 == 0.0) | !isfinite(i34);
	const double i35 = 1.0 / w97; bad |= (w97 == 0.0) | !isfinite(i35);
	{ const double l = w98 * i35; grow |= (fabs(l) > SIMCU_LMAX); w149 -= l * w105; y36 -= l * y34; }
	const double i36 = 1.0 / w149; bad |= (w149 == 0.0) | !isfinite(i36);
	{ const double l = w104 * i36; grow |= (fabs(l) > SIMCU_LMAX); w110 -= l * w112; y32 -= l * y36; }
	const double i37 = 1.0 / w110; bad |= (w110 == 0.0) | !isfinite(i37);
	const double i38 = 1.0 / w107; bad |= (w107 == 0.0) | !isfinite(i38);
	{ const double l = w108 * i38; grow |= (fabs(l) > SIMCU_LMAX); w150 -= l * w115; y39 -= l * y37; }
	const double i39 = 1.0 / w150; bad |= (w150 == 0.0) | !isfinite(i39);
	{ const double l = w114 * i39; grow |= (fabs(l) > SIMCU_LMAX); w120 -= l * w122; y35 -= l * y39; }
	const double i40 = 1.0 / w120; bad |= (w120 == 0.0) | !isfinite(i40);
	const double i41 = 1.0 / w125; bad |= (w125 == 0.0) | !isfinite(i41);
	{ const double l = w124 * i41; grow |= (fabs(l) > SIMCU_LMAX); w116 -= l * w118; y38 -= l * y40; }
	const double i42 = 1.0 / w116; bad |= (w116 == 0.0) | !isfinite(i42);
	{ const double l = w119 * i42; grow |= (fabs(l) > SIMCU_LMAX); w136 -= l * w135; y46 -= l * y38; }
	const double i43 = 1.0 / w132; bad |= (w132 == 0.0) | !isfinite(i43);
	{ const double l = w131 * i43; grow |= (fabs(l) > SIMCU_LMAX); w151 -= l * w134; y43 -= l * y45; }
	const double i44 = 1.0 / w133; bad |= (w133 == 0.0) | !isfinite(i44);
	{ const double l = w151 * i44; grow |= (fabs(l) > SIMCU_LMAX); w128 -= l * w129; y43 -= l * y44; }
	const double i45 = 1.0 / w128; bad |= (w128 == 0.0) | !isfinite(i45);
	{ const double l = w130 * i45; grow |= (fabs(l) > SIMCU_LMAX); w140 -= l * w138; y47 -= l * y43; }
	const double i46 = 1.0 / w136; bad |= (w136 == 0.0) | !isfinite(i46);
	{ const double l = w137 * i46; grow |= (fabs(l) > SIMCU_LMAX); w140 -= l * w139; y47 -= l * y46; }
	const double i47 = 1.0 / w140; bad |= (w140 == 0.0) | !isfinite(i47);
	{ double acc = y47; y47 = acc * i47; }
	{ double acc = y46; acc -= w139 * y47; y46 = acc * i46; }
	{ double acc = y43; acc -= w138 * y47; y43 = acc * i45; }
	{ double acc = y44; acc -= w129 * y43; y44 = acc * i44; }
	{ double acc = y45; acc -= w134 * y44; y45 = acc * i43; }
	{ double acc = y38; acc -= w135 * y46; y38 = acc * i42; }
	{ double acc = y40; acc -= w118 * y38; y40 = acc * i41; }
	{ double acc = y35; y35 = acc * i40; }
	{ double acc = y39; acc -= w122 * y35; y39 = acc * i39; }
	{ double acc = y37; acc -= w115 * y39; y37 = acc * i38; }
	{ double acc = y32; y32 = acc * i37; }
	{ double acc = y36; acc -= w112 * y32; y36 = acc * i36; }
	{ double acc = y34; acc -= w105 * y36; y34 = acc * i35; }
	{ double acc = y29; y29 = acc * i34; }
	{ double acc = y33; acc -= w102 * y29; y33 = acc * i33; }
	{ double acc = y31; acc -= w95 * y33; y31 = acc * i32; }
	{ double acc = y26; y26 = acc * i31; }
	{ double acc = y30; acc -= w92 * y26; y30 = acc * i30; }
	{ double acc = y28; acc -= w85 * y30; y28 = acc * i29; }
	{ double acc = y23; y23 = acc * i28; }
	{ double acc = y27; acc -= w82 * y23; y27 = acc * i27; }
	{ double acc = y25; acc -= w75 * y27; y25 = acc * i26; }
	{ double acc = y20; y20 = acc * i25; }
	{ double acc = y24; acc -= w72 * y20; y24 = acc * i24; }
	{ double acc = y22; acc -= w65 * y24; y22 = acc * i23; }
	{ double acc = y17; y17 = acc * i22; }
	{ double acc = y21; acc -= w62 * y17; y21 = acc * i21; }
	{ double acc = y14; y14 = acc * i20; }
	{ double acc = y15; acc -= w143 * y14; y15 = acc * i19; }
	{ double acc = y16; acc -= w39 * y15; y16 = acc * i18; }
	{ double acc = y18; acc -= w44 * y16; y18 = acc * i17; }
	{ double acc = y19; acc -= w55 * y21; y19 = acc * i16; }
	{ double acc = y11; acc -= w40 * y14; y11 = acc * i15; }
	{ double acc = y12; acc -= w29 * y11; y12 = acc * i14; }
	{ double acc = y13; acc -= w33 * y12; y13 = acc * i13; }
	{ double acc = y2; acc -= w34 * y13; y2 = acc * i12; }
	{ double acc = y9; acc -= w8 * y2; y9 = acc * i11; }
	{ double acc = y10; acc -= w27 * y9; y10 = acc * i10; }
	{ double acc = y3; acc -= w5 * y2; y3 = acc * i9; }
	{ double acc = y4; acc -= w13 * y3; y4 = acc * i8; }
	{ double acc = y41; y41 = acc * i7; }
	{ double acc = y42; y42 = acc * i6; }
	{ double acc = y7; y7 = acc * i5; }
	{ double acc = y8; acc -= w23 * y7; acc -= w7 * y2; y8 = acc * i4; }
	{ double acc = y5; y5 = acc * i3; }
	{ double acc = y6; acc -= w18 * y5; acc -= w6 * y2; y6 = acc * i2; }
	{ double acc = y0; acc -= w14 * y6; acc -= w19 * y8; y0 = acc * i1; }
	{ double acc = y1; y1 = acc * i0; }
	c.Xv[0] = y1;
	c.Xv[1] = y0;
	c.Xv[2] = y2;
	c.Xv[3] = y4;
	c.Xv[4] = y3;
	c.Xv[5] = y6;
	c.Xv[6] = y5;
	c.Xv[7] = y8;
	c.Xv[8] = y7;
	c.Xv[9] = y10;
	c.Xv[10] = y9;
	c.Xv[11] = y11;
	c.Xv[12] = y12;
	c.Xv[13] = y13;
	c.Xv[14] = y15;
	c.Xv[15] = y14;
	c.Xv[16] = y16;
	c.Xv[17] = y19;
	c.Xv[18] = y18;
	c.Xv[19] = y21;
	c.Xv[20] = y22;
	c.Xv[21] = y17;
	c.Xv[22] = y24;
	c.Xv[23] = y25;
	c.Xv[24] = y20;
	c.Xv[25] = y27;
	c.Xv[26] = y28;
	c.Xv[27] = y23;
	c.Xv[28] = y30;
	c.Xv[29] = y31;
	c.Xv[30] = y26;
	c.Xv[31] = y33;
	c.Xv[32] = y34;
	c.Xv[33] = y29;
	c.Xv[34] = y36;
	c.Xv[35] = y37;
	c.Xv[36] = y32;
	c.Xv[37] = y39;
	c.Xv[38] = y38;
	c.Xv[39] = y35;
	c.Xv[40] = y40;
	c.Xv[41] = y42;
	c.Xv[42] = y41;
	c.Xv[43] = y43;
	c.Xv[44] = y45;
	c.Xv[45] = y44;
	c.Xv[46] = y46;
	c.Xv[47] = y47;
	if (bad) c.err = SIMCU_ERR_SINGULAR;
	if (grow) c.growth++;
}

DEV void genHistRows(int *rows)
{
	rows[0] = 19;
	rows[1] = 20;
	rows[2] = 17;
	rows[3] = 18;
	rows[4] = 22;
	rows[5] = 23;
	rows[6] = 21;
	rows[7] = 25;
	rows[8] = 26;
	rows[9] = 24;
	rows[10] = 28;
	rows[11] = 29;
	rows[12] = 27;
	rows[13] = 31;
	rows[14] = 32;
	rows[15] = 30;
	rows[16] = 34;
	rows[17] = 35;
	rows[18] = 33;
	rows[19] = 37;
	rows[20] = 38;
	rows[21] = 36;
	rows[22] = 40;
	rows[23] = 41;
	rows[24] = 39;
	(void)rows;
}

DEV void genProbeRows(int *rows)
{
	rows[0] = 39;
	rows[1] = 16;
	(void)rows;
}

DEV void genMeasures(const Inst &c, double *meas, double tPrev, double tNow, bool first)
{
	(void)c; (void)meas; (void)tPrev; (void)tNow; (void)first;
}

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
		if (!__syncthreads_or((busy || !exhausted) ? 1 : 0)) break;
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
				if (time > k.tstop) { thisStep = k.tstop - prevTime; time = k.tstop; }
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
				if (transient) { genFactorSolve1(c); nfact++; }
				else { genFactorSolve0(c); nfactOp++; }
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
				SIMCU_UNROLL
				for (int r = 0; r < SIMCU_N; r++) c.Xa[r] = c.Xv[r];
				if (opOnly || !(0.0 < k.tstop)) { recordStep(c, hrows, 0.0); done = true; }
				else {
					genInitStep(c);
					recordStep(c, hrows, 0.0);              /* :144 */
					maxStep = ((k.tstep < k.tmax) ? k.tstep : k.tmax) / 100;   /* :109 */
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
				lteStep = k.tmax;                           /* :188-190 */
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
				if (!SIMCU_REPLAY && lteStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; done = true; }
				else {
					order = 1;
					thisStep = lteStep;
					rejLte++;
				}
			} else if (outcome == 0) {                      /* accepted: :220-233 */
				gridOutput(c, prows, prevTime, time, false);
				genMeasures(c, meas, prevTime, time, false);
				SIMCU_UNROLL
				for (int r = 0; r < SIMCU_N; r++) c.Xa[r] = c.Xv[r];
				order = (order < k.maxorder) ? order + 1 : order;
				prevTime = time;
				if (brk) maxStep = Min3(lteStep, 0.1 * maxStep, k.tmax);
				else maxStep = Min3(lteStep, 2 * maxStep, k.tmax);
				accepted++;
				recordStep(c, hrows, time);                 /* :144 of the next pass */
				if (!(time < k.tstop)) done = true;
				else {
					thisStep = maxStep;                     /* :152-154 */
					genNextStep(c, thisStep);
				}
			} else {                                        /* :201-214 */
				thisStep = thisStep / 8;
				if (!SIMCU_REPLAY && thisStep < (k.tstep * 1e-9)) { c.err = SIMCU_ERR_TIMESTEP; done = true; }
				else {
					order = (order > 1) ? order - 1 : order;
					maxStep = maxStep / 8;
					rejNc++;
				}
			}
			if (!done && outcome != 0) {
				/* matrixRecall (:217): back to the last accepted solution */
				SIMCU_UNROLL
				for (int r = 0; r < SIMCU_N; r++) c.Xv[r] = c.Xa[r];
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
