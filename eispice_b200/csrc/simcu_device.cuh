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
/* Products and sums that must NOT be contracted into a*b+c.  The companion
 * models and the LTE estimate rely on exact cancellation: for a capacitor at
 * rest C*x(n+1) - C*x(n) is exactly 0 in the reference's arithmetic, so its
 * divided differences vanish and the step grows; a fused multiply-add keeps the
 * rounding error of one product (1e-16 q), the third divided difference turns
 * it into 1e-16 q / h^3, and the smaller the step the larger the spurious error
 * estimate - the step size collapses and never recovers (measured: cfg4 instance
 * 52883 of the 262 144 sweep, 32 786 accepted steps into a history overflow with
 * contraction, 3 739 steps to the end without). */
#ifdef __CUDA_ARCH__
#define MUL_RN(a, b) __dmul_rn((a), (b))
#define ADD_RN(a, b) __dadd_rn((a), (b))
#else
#define MUL_RN(a, b) ((a) * (b))
#define ADD_RN(a, b) ((a) + (b))
#endif
#define Min3(x, y, z) ((x < y) ? ((z < x) ? z : x) : ((z < y) ? z : y))

/* the analysis window: one for the whole batch (kernel constants), or - in a
 * kernel specialised with SIMCU_WINDOWS - per instance (K-table generation:
 * every (direction, speed) transient of a model has its own tstep / tstop) */
#if defined(SIMCU_JIT) && defined(SIMCU_WINDOWS) && SIMCU_WINDOWS
#define W_TSTEP(c) ((c).wTstep)
#define W_TSTOP(c) ((c).wTstop)
#define W_TMAX(c) ((c).wTmax)
#define W_MINSTEP(c) ((c).wMinstep)
#else
#define W_TSTEP(c) ((c).a.ctl.tstep)
#define W_TSTOP(c) ((c).a.ctl.tstop)
#define W_TMAX(c) ((c).a.ctl.tmax)
#define W_MINSTEP(c) ((c).a.ctl.minstep)
#endif

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
	double wTstep, wTstop, wTmax, wMinstep;      /* SIMCU_WINDOWS: this instance's window */
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
	/* ... and of the break-point tests that run at every attempt (checkbreak.c:43-67:
	 * t[] and the dt of theta[p] are the same in all of them) */
	mutable double cbT0, cbT1, cbDt1;
	mutable int cbN;
	mutable bool cbAdv;
	bool growInv;        /* pivot growth seen in the Newton-invariant blocks of this attempt */
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

#if defined(SIMCU_JIT) && defined(SIMCU_SMEMTAB) && SIMCU_SMEMTAB
/* the tables this topology looks up most (V-I tables first) are staged in
 * shared memory by every block (simcu_jit_kernel.cuh); Table.p then points into
 * the shared window, so entries are read with plain (generic) loads */
extern __shared__ double simcuSmemTab[];
#define TAB_LD(p) (*(p))
#else
#define TAB_LD(p) __ldg(p)
#endif

DEV TabEntry tabLoad(const Table &t, int i)
{
	const double2 a = TAB_LD(reinterpret_cast<const double2 *>(t.p + 4 * (size_t)i));
	const double2 b = TAB_LD(reinterpret_cast<const double2 *>(t.p + 4 * (size_t)i) + 1);
	TabEntry e;
	e.x = a.x; e.y = a.y; e.m = b.x; e.xn = b.y;
	return e;
}

DEV Table tableOf(const SimcuArgs &a, int t)
{
	Table r;
	const int4 d = __ldg(reinterpret_cast<const int4 *>(a.tabDesc) + t);     /* one 16-byte fetch */
	r.len = d.y;
	r.type = d.z;
#if defined(SIMCU_JIT) && defined(SIMCU_SMEMTAB) && SIMCU_SMEMTAB
	r.p = (d.w >= 0) ? simcuSmemTab + 4 * (size_t)d.w : a.tabP + 4 * (size_t)d.x;
#else
	r.p = a.tabP + 4 * (size_t)d.x;
#endif
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

#ifndef SIMCU_SHAREDCB
#define SIMCU_SHAREDCB 0
#endif
#if SIMCU_SHAREDCB
/* Every break-point test that runs at EVERY attempt - sources, TA tables, time
 * dependent B sources, the first test of a T-line - is initialised at load and
 * called at the same times, so their t0 / t1 / dt1 / n hold identical values:
 * one copy per instance, advanced once per step phase (cbBegin), and three
 * doubles of state (x0, x1, dx1) per test instead of six.  Same operands, same
 * tests.  Only a T-line's second test (skipped whenever the first one breaks,
 * tline.c:196-203) keeps its own clock. */
DEV void cbBegin(const Inst &c)
{
	c.cbAdv = (c.time > c.cbT0);
	if (c.cbAdv) { c.cbN++; c.cbDt1 = c.cbT0 - c.cbT1; c.cbT1 = c.cbT0; }
	c.cbT0 = c.time;
}

DEV int cbIsBreakClocked(const Inst &c, int sb, int ib, double x, double maxAngle)
{
	(void)ib;
	if (c.cbAdv) {
		c.S(sb + 7) = c.S(sb) - c.S(sb + 1);
		c.S(sb + 1) = c.S(sb);
	}
	c.S(sb) = x;
	if (c.cbN < 1) return 0;
	return cbAngleBreak(x - c.S(sb + 1), c.time - c.cbT1, c.S(sb + 7), c.cbDt1, maxAngle);
}
#else
#define cbIsBreakClocked cbIsBreak
#endif
#else
#define cbIsBreakClocked cbIsBreak
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
 * lands on the same age of the previous ring.
 * For the f ring that is age a of the y ring, read only while the counter is
 * below a - and until then that slot still holds the 0 of itInitialize (a value
 * needs a advances to get there): AGEQF reads the constant, and ages 1 and 2 of
 * the y ring are neither stored nor rotated (integrator.c never reads y[] at
 * any other index than the current one). */
#define AGE(c, sb, k, a) (c).S((sb) + 2 + (k) * 4 + (a))
#define AGEQF(c, sb, a, nn) (((nn) >= (a)) ? AGE(c, sb, RF, a) : 0.0)
enum { RT = 0, RH = 1, RX = 2, RY = 3, RF = 4 };

DEV void itInitialize(const Inst &c, int sb, int ib, double ic, double ydtdx)
{
	(void)ib;
	c.S(sb) = 0.0; c.S(sb + 1) = 0.0;
	c.itN = 0; c.itAdv = false;
	SIMCU_UNROLL
	for (int j = 0; j < 4; j++) {
		c.itT[j] = 0.0;
		AGE(c, sb, RX, j) = ic; AGE(c, sb, RF, j) = ydtdx;
	}
	AGE(c, sb, RY, 0) = 0.0;
	SIMCU_UNROLL
	for (int j = 0; j < 3; j++) c.itH[j] = W_TSTOP(c);
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
		AGE(c, sb, RX, 2) = AGE(c, sb, RX, 1); AGE(c, sb, RX, 1) = AGE(c, sb, RX, 0);
		AGE(c, sb, RF, 2) = AGE(c, sb, RF, 1); AGE(c, sb, RF, 1) = AGE(c, sb, RF, 0);
	}
	const double h = c.itH[0];
	AGE(c, sb, RX, 0) = x0;
	const double yn = ADD_RN(MUL_RN(c.S(sb + 1), x0), -c.S(sb));
	AGE(c, sb, RY, 0) = yn;
	AGE(c, sb, RF, 0) = ydtdx;
	const double fp = AGEQF(c, sb, 1, nn);
	if (c.order < 2) {
		dydx0 = fp / h;
		y0 = dydx0 * x0;
	} else {
		dydx0 = 2.0 * ydtdx / h;
		double mult = ydtdx / fp;
		if (isnan(mult)) mult = 1 / c.a.ctl.gmin;
		y0 = ADD_RN(MUL_RN(dydx0, x0), MUL_RN(mult, yn));
	}
	c.S(sb + 1) = dydx0;
	c.S(sb) = y0;
}

/* integrator.c:61-102 (LTE step recommendation; float sqrtf at :99) */
DEVFN double itNextStepFn(int order, double reltol, double chgtol, double trtol, double abstol,
		double x0, double ydtdx, double ieq, double g, double ry0, double hn, double xn, double fn,
		double f1, double x1, double h1, double f2, double x2, double h2)
{
	const double y0 = ADD_RN(MUL_RN(g, x0), -ieq);
	const double ey = ADD_RN(MUL_RN(reltol, MaxAbs(ry0, y0)), abstol);
	const double eyp = MUL_RN(reltol, MaxAbs(MUL_RN(MaxAbs(x0, xn), (ydtdx / hn)), chgtol));
	const double e = MaxAbs(eyp, ey);
	const double q0 = MUL_RN(ydtdx, x0), qn = MUL_RN(fn, xn), q1 = MUL_RN(f1, x1);
	double h, dd;
	if (order < 2) {
		dd = DD2(q0, qn, q1, hn, h1);
		dd *= 1.0 / 2.0;
		h = trtol * e / MaxAbs(dd, abstol);
	} else {
		dd = DD3(q0, qn, q1, MUL_RN(f2, x2), hn, h1, h2);
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
	 * (x <- h, h <- t; f <- y, i.e. AGEQF) */
	const double x1 = (nn >= 1) ? AGE(c, sb, RX, 1) : c.itH[1], x2 = (nn >= 2) ? AGE(c, sb, RX, 2) : c.itH[2];
	const double h1 = (nn >= 1) ? c.itH[1] : c.itT[1], h2 = (nn >= 2) ? c.itH[2] : c.itT[2];
	const double h = itNextStepFn(c.order, k.reltol, k.chgtol, k.trtol, abstol, x0, ydtdx,
			c.S(sb), c.S(sb + 1), AGE(c, sb, RY, 0), c.itH[0], AGE(c, sb, RX, 0),
			AGE(c, sb, RF, 0), AGEQF(c, sb, 1, nn), x1, h1, AGEQF(c, sb, 2, nn), x2, h2);
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
		RING(c, sb, RT, j) = 0.0; RING(c, sb, RH, j) = W_TSTOP(c);
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
	const double yn = ADD_RN(MUL_RN(c.D(sb + 1), x0), -c.D(sb));
	RING(c, sb, RY, n) = yn;
	RING(c, sb, RF, n) = ydtdx;
	if (c.order < 2) {
		dydx0 = RING(c, sb, RF, p) / h;
		y0 = dydx0 * x0;
	} else {
		dydx0 = 2.0 * ydtdx / h;
		double mult = ydtdx / RING(c, sb, RF, p);
		if (isnan(mult)) mult = 1 / c.a.ctl.gmin;
		y0 = ADD_RN(MUL_RN(dydx0, x0), MUL_RN(mult, yn));
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
	const double y0 = ADD_RN(MUL_RN(c.D(sb + 1), x0), -c.D(sb));
	const double ey = ADD_RN(MUL_RN(k.reltol, MaxAbs(RING(c, sb, RY, n), y0)), abstol);
	const double hn = RING(c, sb, RH, n), xn = RING(c, sb, RX, n);
	const double eyp = MUL_RN(k.reltol, MaxAbs(MUL_RN(MaxAbs(x0, xn), (ydtdx / hn)), k.chgtol));
	const double e = MaxAbs(eyp, ey);
	const double q0 = MUL_RN(ydtdx, x0), qn = MUL_RN(RING(c, sb, RF, n), xn);
	const double q1 = MUL_RN(RING(c, sb, RF, p1), RING(c, sb, RX, p1));
	double h, dd;
	if (c.order < 2) {
		dd = DD2(q0, qn, q1, hn, RING(c, sb, RH, p1));
		dd *= 1.0 / 2.0;
		h = k.trtol * e / MaxAbs(dd, abstol);
	} else {
		dd = DD3(q0, qn, q1, MUL_RN(RING(c, sb, RF, p2), RING(c, sb, RX, p2)),
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
	const double tstep = W_TSTEP(c), tend = W_TSTOP(c) + W_TSTEP(c), fmin = 1 / tend;
	(void)k;
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
		const Table t = tableOf(c.a, d[DF_WTAB] + c.tableSet(d[DF_WSETSLOT]));
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
	const double tstep = W_TSTEP(c), tend = W_TSTOP(c) + W_TSTEP(c);
	(void)k;
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
		const Table t = tableOf(c.a, d[DF_WTAB] + c.tableSet(d[DF_WSETSLOT]));
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
			const Table t = tableOf(c.a, d[DF_WTAB] + c.tableSet(d[DF_WSETSLOT]));   /* waveform.c:385-389 */
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
		if (d[DF_PBYPASS] >= 0 && c.P(d[DF_PBYPASS]) != 0.0) {
			/* masked section of a maximum topology (not in the
			 * reference): the two ports are joined by an ideal 0 V branch -
			 * V_K - V_J = V_L - V_M, the branch current i(ref#in1) flows through
			 * both, i(ref#in2) = 0 - at the operating point and in transient */
			c.Aset(s[0], 1.0); c.Aset(s[1], -1.0); c.Aset(s[4], -1.0); c.Aset(s[5], 1.0);
			c.Aset(s[2], 1.0); c.Aset(s[3], -1.0); c.Aset(s[16], -1.0); c.Aset(s[17], 1.0);
			c.Aset(s[6], 0.0); c.Aset(s[7], 0.0); c.Aset(s[8], 0.0); c.Aset(s[9], 0.0);
			c.Aset(s[10], 0.0); c.Aset(s[11], 0.0); c.Aset(s[14], 0.0); c.Aset(s[15], 0.0);
			c.Aplus(s[13], 1.0);
			c.S(sb + 24) = 1.0;
			break;
		}
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
		if (d[DF_PBYPASS] >= 0 && c.P(d[DF_PBYPASS]) != 0.0) break;
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
		return cbIsBreakClocked(c, sb + 1, ib, dc, (type == DK_V) ? k.maxAngleV : k.maxAngleA);
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
		return cbIsBreakClocked(c, sb + 4, ib + 1, aa, k.maxAngleA);
	}
	case DK_T: {                                /* tline.c:126-204 */
		if (d[DF_PBYPASS] >= 0 && c.P(d[DF_PBYPASS]) != 0.0) return 0;
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
		int brk = cbIsBreakClocked(c, sb + 6, ib, Vr, k.maxAngleV);
		if (!brk) brk = cbIsBreak(c, sb + 15, ib + 1, Vs, k.maxAngleV);
		return brk;
	}
	case DK_BI: case DK_BV:                     /* nonlinear_i.c:198-218, _v.c:166-188 */
		if (!d[DF_USETIME]) return 0;
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d, bv);
		return cbIsBreakClocked(c, sb + 4, ib + 1, c.S(sb + 2),
				(type == DK_BV) ? k.maxAngleV : k.maxAngleA);
	case DK_BC:                                 /* nonlinear_c.c:172-184 */
		if (!d[DF_USETIME]) return 0;
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d, bv);
		return 0;
	default: return 0;
	}
}

/* The companion model stamped at the previous attempt (the devices' G and Ieq,
 * capacitor.c:101-108) is what the integrator itself keeps as its last dydx0 and
 * y0 (integrator.c:158-160): both are 0 until the first integration and are
 * written together from the same numbers.  The specialised kernel reads the
 * integrator's pair and stores nothing twice. */
#ifdef SIMCU_JIT
#define COMPANION_OLD(c, sb, g, i) const double g = (c).S((sb) + 3), i = (c).S((sb) + 2)
#define COMPANION_KEEP(c, sb, g, i) ((void)0)
#else
#define COMPANION_OLD(c, sb, g, i) const double g = (c).S(sb), i = (c).S((sb) + 1)
#define COMPANION_KEEP(c, sb, g, i) do { (c).S(sb) = (g); (c).S((sb) + 1) = (i); } while (0)
#endif

DEV void devIntegrate(Inst &c, const int *d)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE], ib = d[DF_IBASE];
	double G, Ieq;
	switch (type) {
	case DK_C: case DK_BC: {                    /* capacitor.c:76-108, nonlinear_c.c:232-277 */
		const double x0 = c.X(d[DF_K]) - c.X(d[DF_J]);
		const double cap = (type == DK_C) ? c.P(d[DF_P0]) : c.S(sb + 24);
		COMPANION_OLD(c, sb, Gold, Iold);
		itIntegrate(c, sb + 2, ib, x0, cap, G, Ieq);
		if (c.err) return;
		const int *s = (type == DK_C) ? d + DF_A0 : d + DF_BA0;   /* KK KJ JK JJ */
		const double dG = G - Gold;
		c.Aplus(s[0], dG); c.Aplus(s[1], -dG); c.Aplus(s[2], -dG); c.Aplus(s[3], dG);
		const double dI = Ieq - Iold;
		c.Bplus(d[DF_K], dI); c.Bplus(d[DF_J], -dI);
		COMPANION_KEEP(c, sb, G, Ieq);
		break;
	}
	case DK_L: {                                /* inductor.c:78-109 */
		const double x0 = c.X(d[DF_ROWR]);
		COMPANION_OLD(c, sb, Gold, Iold);
		itIntegrate(c, sb + 2, ib, x0, c.P(d[DF_P0]), G, Ieq);
		if (c.err) return;
		c.Aplus(d[DF_A0 + 4], -(G - Gold));
		c.Bplus(d[DF_ROWR], -(Ieq - Iold));
		COMPANION_KEEP(c, sb, G, Ieq);
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
	if ((loc > W_MINSTEP(c)) && (loc < thisStep)) thisStep = loc;
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
	const double tstep = W_TSTEP(c), tstop = W_TSTOP(c);
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
