/*
 * simcu_kernels.cu - device side of the batched transient engine (sm_100a).
 *
 * One thread owns one instance for the whole run.  All per-instance data is
 * structure-of-arrays with the instance index fastest (simcu_program.h), so
 * every stamp, state access, LU load and solution store of a warp is one
 * coalesced 256-byte line.  The numeric LU works on a per-thread workspace of
 * F + N doubles that sits in SHARED memory ([slot][blockDim], conflict-free)
 * whenever it fits, so a Newton iteration costs HBM only nnzA + N loads; the
 * factors never leave the SM.  The elimination order, fill pattern and pivot
 * sequence are fixed on the host (simcu_host.cpp: symbolic analysis on a pilot
 * batch) and arrive as an op-stream that every thread of a warp walks in
 * lock-step - warp-uniform control flow, data-dependent divergence only in
 * the Newton / step-rejection loop trip counts.
 *
 * Semantics follow the reference file:line cited at each function (paths
 * relative to the reference root); tests/ compare every phase with the CPU
 * oracle.  FP64 CUDA cores only: the work is sparse, N <= ~100 and
 * memory-bound, there is nothing to give the tensor cores.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "simcu_calc.h"
#include "simcu_program.h"
#include "simcu_kernels.h"

#define DEV __device__ __forceinline__

#define MaxAbs(x, y) ((fabs(x) > fabs(y)) ? fabs(x) : fabs(y))
#define Min3(x, y, z) ((x < y) ? ((z < x) ? z : x) : ((z < y) ? z : y))

/* per-thread view of one instance */
struct Inst {
	const SimcuArgs &a;
	size_t M;
	int i;
	double *ws;          /* LU workspace of this thread */
	int wsStride;
	bool xInWs;          /* true: devices read the Newton iterate from ws */
	double time;
	int order;
	int err;             /* SIMCU_ERR_* of this instance (0 = ok) */

	DEV Inst(const SimcuArgs &args, int inst, double *w, int stride)
		: a(args), M(args.M), i(inst), ws(w), wsStride(stride), xInWs(false),
		  time(0.0), order(1), err(0) {}

	DEV double P(int pid) const
	{
		const int m = __ldg(&a.pmap[pid]);
		return (m >= 0) ? a.pinst[(size_t)m * M + i] : __ldg(&a.pconst[pid]);
	}
	DEV double &S(int slot) const { return a.S[(size_t)slot * M + i]; }
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

/* ------------------------------------------------------------------------ *
 * Piecewise tables: libs/simulator/math/piecewise.c:44-67,96-104,226-263
 * ------------------------------------------------------------------------ */
struct Table { const double *x, *y, *m; int len; int type; };

DEV Table tableOf(const SimcuArgs &a, int t)
{
	Table r;
	const int off = __ldg(&a.tabOff[t]);
	r.len = __ldg(&a.tabOff[t + 1]) - off;
	r.type = __ldg(&a.tabType[t]);
	r.x = a.tabX + off; r.y = a.tabY + off; r.m = a.tabM + off;
	return r;
}

/* the cached index only shortens the walk; the landing index depends on x0 */
DEV int pwSetIndex(const Table &t, int index, double x0)
{
	if ((index < 0) || (index > t.len - 1)) index = 0;
	while ((index >= 0) && (index < t.len - 1)) {
		if (__ldg(&t.x[index]) <= x0) {
			if (index == t.len - 2) return index;
			if (__ldg(&t.x[index + 1]) > x0) return index;
			index++;
		} else {
			if (index == 0) return index;
			index--;
		}
	}
	return index;
}

DEV void pwCalcValue(const Table &t, int &index, double x0, double &y0, double &dydx0)
{
	index = pwSetIndex(t, index, x0);
	const int i = index;
	if ((i == 0) || (i >= t.len - 2)) {
		if (__ldg(&t.x[0]) > x0) { y0 = __ldg(&t.y[0]); dydx0 = 0; return; }
		if (__ldg(&t.x[t.len - 1]) < x0) { y0 = __ldg(&t.y[t.len - 1]); dydx0 = 0; return; }
	}
	if (t.type == 'l') {
		const double mi = __ldg(&t.m[i]);
		y0 = __ldg(&t.y[i]) + mi * (x0 - __ldg(&t.x[i]));
		dydx0 = mi;
	} else {
		const double xi = __ldg(&t.x[i]), yi = __ldg(&t.y[i]);
		const double mi = __ldg(&t.m[i]), mi1 = __ldg(&t.m[i + 1]);
		const double dx0 = x0 - xi, dy = __ldg(&t.y[i + 1]) - yi;
		const double dx = __ldg(&t.x[i + 1]) - xi;
		const double pa = dy / dx - dx * (mi1 + mi * 2.0) / 6.0;
		const double pb = 0.5 * mi;
		const double pc = (mi1 - mi) / (6.0 * dx);
		y0 = yi + dx0 * (pa + dx0 * (pb + dx0 * pc));
		dydx0 = pa + dx0 * (2 * pb + 3 * pc * dx0);
	}
}

/* piecewise.c:204-222 with its two dead branches */
DEV double pwGetNextX(const Table &t, int &index, double x0)
{
	index = pwSetIndex(t, index, x0);
	if (index == t.len - 1) return HUGE_VAL;
	return __ldg(&t.x[index + 1]);
}

/* ------------------------------------------------------------------------ *
 * Break-point test: libs/simulator/math/checkbreak.c:43-67,71-95
 * state: x[3] t[3] theta[3] at sb, n at ib
 * ------------------------------------------------------------------------ */
DEV void cbInitialize(const Inst &c, int sb, int ib, double ic)
{
	for (int k = 0; k < 3; k++) { c.S(sb + k) = ic; c.S(sb + 3 + k) = 0.0; }
	/* theta[] is not reset by the reference; it is zero from allocation */
	c.IS(ib) = 0;
}

DEV int cbIsBreak(const Inst &c, int sb, int ib, double x, double maxAngle)
{
	int n = c.IS(ib);
	if (c.time > c.S(sb + 3 + n % 3)) { n++; c.IS(ib) = n; }
	const int k = n % 3;
	int p = (n - 1) % 3;
	if (p < 0) p = 0;
	c.S(sb + k) = x;
	c.S(sb + 3 + k) = c.time;
	const double th = atan2(x - c.S(sb + p), c.time - c.S(sb + 3 + p));
	c.S(sb + 6 + k) = th;
	return (fabs(th - c.S(sb + 6 + p)) > maxAngle) ? 1 : 0;
}

/* ------------------------------------------------------------------------ *
 * Newton convergence test: libs/simulator/math/checklinear.c:39-91
 * state: lastV (double), state (int 0='a' 1='b' 2='c')
 * ------------------------------------------------------------------------ */
DEV int clIsLinear(const Inst &c, int sLast, int iState, double tol, double V,
		double calcedV)
{
	const double reltol = c.a.ctl.reltol;
	const double lastV = c.S(sLast);
	double e = reltol * MaxAbs(lastV, V) + tol;
	if (fabs(lastV - V) > e) { c.IS(iState) = 0; c.S(sLast) = V; return 0; }
	e = reltol * MaxAbs(calcedV, V) + tol;
	if (fabs(calcedV - V) > e) { c.IS(iState) = 0; c.S(sLast) = V; return 0; }
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
#define RING(c, sb, k, j) (c).S((sb) + 2 + (k) * 4 + (j))
enum { RT = 0, RH = 1, RX = 2, RY = 3, RF = 4 };

DEV void itInitialize(const Inst &c, int sb, int ib, double ic, double ydtdx)
{
	c.S(sb) = 0.0; c.S(sb + 1) = 0.0; c.IS(ib) = 0;
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
	const double yn = c.S(sb + 1) * x0 - c.S(sb);
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
	c.S(sb + 1) = dydx0;
	c.S(sb) = y0;
}

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

/* integrator.c:61-102 (LTE step recommendation; float sqrtf at :99) */
DEV double itNextStep(Inst &c, int sb, int ib, double x0, double ydtdx, double abstol)
{
	if (isnan(x0)) { c.err = SIMCU_ERR_NAN; return 0.0; }
	const SimcuControl &k = c.a.ctl;
	const int nn = c.IS(ib);
	const int n = nn % 4, p1 = (nn - 1) % 4, p2 = (nn - 2) % 4;
	const double y0 = c.S(sb + 1) * x0 - c.S(sb);
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
DEV void bLoad(Inst &c, const int *d)
{
	const int type = d[DF_TYPE], sb = d[DF_SBASE] + bOffIc(type);
	const int gb = d[DF_SBASE] + bOffG(type);
	const int nbv = d[DF_NBV];
	const int *bv = c.a.bvar + 4 * d[DF_BVOFF];
	double eqCalc = c.S(sb);
	for (int k = 0; k < nbv && !c.err; k++) {
		const int row = __ldg(&bv[4 * k]), sA = __ldg(&bv[4 * k + 1]);
		const int sB = __ldg(&bv[4 * k + 2]), cv = __ldg(&bv[4 * k + 3]);
		double f, G;
		bSolve(c, d, cv, f, G);
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
DEV void devLoad(Inst &c, const int *d)
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
		break;
	}
	case DK_VI: {                               /* vicurve.c:167-221 */
		c.S(sb + 3) = 0.0; c.IS(ib) = 0;
		cbInitialize(c, sb + 4, ib + 1, 0.0);
		const int set = (d[DF_SETSLOT] >= 0) ? c.a.iinst[(size_t)d[DF_SETSLOT] * c.M + c.i] : 0;
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
		if (!c.err) bLoad(c, d);
		break;
	}
	}
}

DEV int devLinearize(Inst &c, const int *d)
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
			const int set = (d[DF_SETSLOT] >= 0) ? c.a.iinst[(size_t)d[DF_SETSLOT] * c.M + c.i] : 0;
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
	if (!linear) bLoad(c, d);
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
DEV bool histLocate(Inst &c, int ib, double tp, int &iP, int &iN)
{
	const SimcuArgs &a = c.a;
	const int count = c.CI(CI_HCOUNT), R = a.histRecords;
	const int lo = (count > R) ? count - R : 0;
	int i = c.IS(ib + 2);
	if (i < lo || i >= count) i = lo;
	#define HT(r) a.Ht[(size_t)((r) % R) * c.M + c.i]
	while (HT(i) > tp && i > lo) i--;
	if (HT(i) > tp) return false;             /* fell off the ring */
	while (i + 1 < count && HT(i + 1) <= tp) i++;
	#undef HT
	c.IS(ib + 2) = i;
	if (i + 1 < count) { iP = i; iN = i + 1; }
	else { if (i - 1 < lo) return false; iP = i - 1; iN = i; }
	return true;
}

DEV double histInterp(const Inst &c, int iP, int iN, int col, double tp)
{
	if (col < 0) return 0.0;
	const SimcuArgs &a = c.a;
	const int R = a.histRecords;
	const double tP = a.Ht[(size_t)(iP % R) * c.M + c.i], tN = a.Ht[(size_t)(iN % R) * c.M + c.i];
	const double xP = a.Hx[((size_t)(iP % R) * a.nh + col) * c.M + c.i];
	const double xN = a.Hx[((size_t)(iN % R) * a.nh + col) * c.M + c.i];
	return ((xN - xP) / (tN - tP)) * (tp - tP) + xP;
}

DEV int devStep(Inst &c, const int *d)
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
		const int set = (d[DF_SETSLOT] >= 0) ? c.a.iinst[(size_t)d[DF_SETSLOT] * c.M + c.i] : 0;
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
			int iP, iN;
			if (!histLocate(c, ib, tp, iP, iN)) { c.err = SIMCU_ERR_HISTORY; return 0; }
			const int *h = d + DF_TH0;              /* R S K J L M */
			Ir = histInterp(c, iP, iN, h[0], tp); Is = histInterp(c, iP, iN, h[1], tp);
			Vk = histInterp(c, iP, iN, h[2], tp); Vj = histInterp(c, iP, iN, h[3], tp);
			Vl = histInterp(c, iP, iN, h[4], tp); Vm = histInterp(c, iP, iN, h[5], tp);
		}
		const double z0 = c.P(d[DF_PZ0]), loss = c.P(d[DF_PLOSS]);
		double Vr, Vs;
		if (loss == HUGE_VAL) {
			Vr = (Vl - Vm) + z0 * Is;
			Vs = (Vk - Vj) + z0 * Ir;
		} else {
			Vr = exp(-loss / 2) * ((Vl - Vm) + z0 * Is);
			Vs = exp(-loss / 2) * ((Vk - Vj) + z0 * Ir);
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
		if (!c.err) bLoad(c, d);
		return cbIsBreak(c, sb + 4, ib + 1, c.S(sb + 2),
				(type == DK_BV) ? k.maxAngleV : k.maxAngleA);
	case DK_BC:                                 /* nonlinear_c.c:172-184 */
		if (!d[DF_USETIME]) return 0;
		bCalculate(c, d, true);
		if (!c.err) bLoad(c, d);
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
		const int set = (d[DF_SETSLOT] >= 0) ? c.a.iinst[(size_t)d[DF_SETSLOT] * c.M + c.i] : 0;
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

/* matrixRecord (core/matrix.c:372-381) as far as anything reads it: the
 * T-line history ring, the optional raw record, the gridded probes */
DEV void recordStep(Inst &c, double tPrev, double tNow, bool first)
{
	const SimcuArgs &a = c.a;
	if (a.nh > 0 && a.histRecords > 0) {
		const int cnt = c.CI(CI_HCOUNT), slot = cnt % a.histRecords;
		a.Ht[(size_t)slot * c.M + c.i] = tNow;
		for (int j = 0; j < a.nh; j++)
			a.Hx[((size_t)slot * a.nh + j) * c.M + c.i] = c.X(__ldg(&a.histRows[j]));
		c.CI(CI_HCOUNT) = cnt + 1;
	}
	if (a.rawMaxPoints > 0 && c.i >= a.rawFirst && c.i < a.rawFirst + a.rawCount) {
		const int n = c.CI(CI_NPTS);
		if (n < a.rawMaxPoints) {
			double *o = a.outRaw + ((size_t)n * (a.N + 1)) * a.rawCount + (c.i - a.rawFirst);
			o[0] = tNow;
			for (int k = 0; k < a.N; k++) o[(size_t)(k + 1) * a.rawCount] = c.X(k + 1);
		}
		c.CI(CI_NPTS) = n + 1;
	}
	(void)tPrev; (void)first;
}

/* linear interpolation of the probes onto the tstep grid, exactly what
 * numpy.interp does on the adaptive waveform: for a grid time tg with
 * tPrev <= tg < tNow: slope * (tg - tPrev) + xPrev; a knot returns its value */
DEV void gridOutput(Inst &c, double tPrev, double tNow, bool first)
{
	const SimcuArgs &a = c.a;
	if (a.nprobes <= 0 || a.ngrid <= 0) return;
	int g = c.CI(CI_GRIDIDX);
	const double tstep = a.ctl.tstep, tstop = a.ctl.tstop;
	while (g < a.ngrid) {
		double tg = g * tstep;
		if (g == a.ngrid - 1 || tg > tstop) tg = tstop;
		if (tg > tNow) break;
		for (int p = 0; p < a.nprobes; p++) {
			const int row = __ldg(&a.probeRows[p]);
			const double xN = row ? c.ws[(size_t)__ldg(&a.xslot[row - 1]) * c.wsStride] : 0.0;
			double v;
			if (first || tg == tNow) v = xN;
			else {
				const double xP = row ? a.X[(size_t)(row - 1) * c.M + c.i] : 0.0;
				v = ((xN - xP) / (tNow - tPrev)) * (tg - tPrev) + xP;
			}
			a.outGrid[((size_t)g * a.nprobes + p) * c.M + c.i] = v;
		}
		g++;
	}
	c.CI(CI_GRIDIDX) = g;
}

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
	const int count = newtonSolve(c, a.ctl.itl1, nfact);
	c.CI(CI_NFACT_OP) = nfact;
	if (!c.err && count > a.ctl.itl1) c.err = SIMCU_ERR_OP;
	if (c.err) { c.CI(CI_STATUS) = c.err; c.CI(CI_DONE) = 1; return; }
	gridOutput(c, 0.0, 0.0, true);
	storeSolution(c);
	c.xInWs = false;
	if (opOnly) { recordStep(c, 0.0, 0.0, true); c.CI(CI_DONE) = 1; return; }
	for (int k = 0; k < a.ndev; k++) devInitStep(c, a.dev + k * SIMCU_DEV_STRIDE);
	recordStep(c, 0.0, 0.0, true);
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
			gridOutput(c, prevTime, time, false);
			storeSolution(c);
			c.xInWs = false;
			order = (order < k.maxorder) ? order + 1 : order;
			const double tp = prevTime;
			prevTime = time;
			if (brk) maxStep = Min3(lteStep, 0.1 * maxStep, k.tmax);
			else maxStep = Min3(lteStep, 2 * maxStep, k.tmax);
			accepted++;
			recordStep(c, tp, time, false);
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
