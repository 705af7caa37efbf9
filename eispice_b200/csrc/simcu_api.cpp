/*
 * simcu_api.cpp - the C-ABI of libsimulator_cuda (include/simulator_cuda.h):
 * builder calls with the reference's argument lists, device memory of one
 * batch, and the run sequence load -> operating point -> transient kernels.
 *
 * There is no CPU execution path in here: every run call needs a CUDA device
 * and fails (-1, message on stderr) without one.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "simcu_calc.h"
#include "simcu_codegen.h"
#include "simcu_host.h"
#include "simcu_kernels.h"

using namespace simcu;

namespace {

int gQuiet = 0;

int fail(const char *fmt, ...)
{
	if (!gQuiet) {
		va_list ap;
		va_start(ap, fmt);
		std::fprintf(stderr, "simulator_cuda: ");
		std::vfprintf(stderr, fmt, ap);
		std::fprintf(stderr, "\n");
		va_end(ap);
	}
	return -1;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail("%s: %s", #call, cudaGetErrorString(e_)); } while (0)

/* one device allocation that only ever grows */
struct DBuf {
	void *p = nullptr;
	size_t bytes = 0;
	int reserve(size_t n)
	{
		if (n <= bytes && p) return 0;
		if (p) cudaFree(p);
		p = nullptr; bytes = 0;
		if (n == 0) n = 8;
		cudaError_t e = cudaMalloc(&p, n);
		if (e != cudaSuccess) {
			p = nullptr;
			return fail("cudaMalloc(%zu bytes): %s", n, cudaGetErrorString(e));
		}
		bytes = n;
		return 0;
	}
	template <class T> int upload(const std::vector<T> &v, cudaStream_t st = 0)
	{
		if (reserve(v.size() * sizeof(T))) return -1;
		if (v.empty()) return 0;
		cudaError_t e = cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T),
				cudaMemcpyHostToDevice, st);
		return (e == cudaSuccess) ? 0 : fail("upload: %s", cudaGetErrorString(e));
	}
	void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
	template <class T> T *as() const { return (T *)p; }
};

}  /* namespace */

struct _simcu {
	int M = 1;
	Netlist nl;
	bool locked = false;
	/* options (reference include/control.h:34-56 + engine knobs) */
	SimcuControl ctl;
	double minstepOpt = -1.0;
	int histRecordsOpt = 0, tpbOpt = 0, wsSharedOpt = -1;
	int rawMaxPoints = 0, rawFirst = 0, rawCount = 0;
	long long maxAttempts = 0;
	int attemptsPerLaunch = INT_MAX;
	double pivotThreshold = 0.1;
	int numPilots = 8;
	/* compiled state */
	bool compiled = false, prepared = false, ran = false;
	bool paramsChanged = false;     /* per-instance values differ from those the pivots were chosen on */
	bool ranTransient = false;      /* a transient has completed (restart = 0 would continue it) */
	int single = -1;                /* >= 0: the next launch processes this instance only */
	Program prog;
	Schedule sched[2];
	std::vector<int> forcePc[2], forcePr[2];
	std::vector<std::string> probeNames;
	std::vector<int> probeRows;
	int ngrid = 0, histRecords = 0, tpb = 64, wsShared = 1;
	int interpreter = 0;            /* 1: run the batch on the generic kernels */
	int keepOrder = 0;              /* 1: do not group instances by table set */
	int historyRetry = 0;           /* 1: re-run instances that outran the history ring */
	int minBlocks = 0;              /* __launch_bounds__ second argument */
	int maxBlocksPerSM = 0;         /* cap on resident blocks (0 = occupancy limit) */
	int minBlocksUsed = 0;
	JitConfig jitCfg;               /* what r->jit was compiled with */
	int replay = 0;                 /* test instrumentation: replay an attempt log */
	int claimBelow = 31;            /* lane claiming policy of the specialised kernel */
	int blockSync = -1;             /* warps of a block run their attempts in step (I-cache sharing); -1 = automatic */
	int blockSyncUsed = 0;
	int syncGroup = 0;              /* warps per barrier group (0 = the whole block) */
	double growthLimit = 1e6;       /* LU multiplier magnitude that counts as pivot growth */
	int tlinePrefetch = 0;          /* locate + prefetch pass over the T-lines (>= 2 lines): measured slower */
	int fmad = 0;                   /* 0: no a*b+c contraction - the reference's arithmetic (default) */
	int solveBlocks = 2;            /* generated LU: 0 all at once, 1 block by block, 2 + invariant blocks once per Newton loop */
	int sharedCb = 1;               /* one break-point clock per instance */
	std::vector<int> replayOff, replayFlag, replayLinOff, replayLin;
	std::vector<double> winTstep, winTstop, winTmax;   /* per-instance analysis windows (empty: one window) */
	std::vector<double> replayTime;
	std::vector<char> active[2];
	std::vector<double> consts[2];      /* per phase: literal value of a slot, NaN = varies */
	std::vector<int> orderKey;          /* table-set indices the processing order was built from */
	std::vector<int> measures;          /* (kind, row) pairs */
	std::vector<double> measureLevels;
	std::vector<int> jitMeasures;
	std::shared_ptr<JitKernel> jit;
	std::vector<int> jitProbes;
	int jitBlocks = 0, lanes = 32, lanesOpt = 0;
	SimcuArgs args[2];
	simcuStats_ stats;
	/* device memory */
	DBuf dTabDesc, dTabStage;
	int stagePoints = 0;            /* table points staged in shared memory by every block */
	bool stageOK = false;           /* this launch shape stages tables (one big block per SM) */
	int smemTablesKB = -1;          /* budget for that (option smem_tables_kb; 0 = tables stay global, -1 = automatic) */
	int smemTablesUsed = 0;
	DBuf dDev, dBvar, dLu[2], dXslot[2], dCalcOff, dCalcCode, dCalcConst,
		dCalcVarOff, dCalcVars, dTabP, dHistRows,
		dProbeRows, dPmap, dPconst, dPinst, dIinst, dOrder, dOrder2, dMeasure, dMeasureLevel,
		dReplayOff, dReplayTime, dReplayFlag, dReplayLinOff, dReplayLin, dWin;
	DBuf dA, dB, dX, dS, dIS, dCD, dCI, dHt, dHx, dOutGrid, dOutT, dStatusT, dOutRaw, dWs, dRemaining;
	int *hRemaining = nullptr;      /* pinned */
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
	std::vector<char *> variableNames;

	_simcu()
	{
		std::memset(&ctl, 0, sizeof ctl);
		ctl.itl1 = 100; ctl.itl4 = 10; ctl.maxorder = 2;          /* control.c:53-63 */
		ctl.reltol = 0.001; ctl.vntol = 1e-6; ctl.abstol = 1e-12;
		ctl.captol = 1e-18; ctl.chgtol = 1e-14; ctl.trtol = 1; ctl.gmin = 1e-15;
		ctl.maxAngleA = M_PI / 3; ctl.maxAngleV = M_PI / 3;       /* :67-68 */
		std::memset(&stats, 0, sizeof stats);
		std::memset(args, 0, sizeof args);
	}
	~_simcu()
	{
		DBuf *all[] = {&dDev, &dBvar, &dLu[0], &dLu[1], &dXslot[0], &dXslot[1], &dCalcOff,
			&dCalcCode, &dCalcConst, &dCalcVarOff, &dCalcVars, &dTabP, &dOrder, &dOrder2, &dMeasure,
			&dMeasureLevel, &dReplayOff, &dReplayTime, &dReplayFlag, &dReplayLinOff, &dReplayLin, &dWin, &dTabDesc, &dTabStage, &dHistRows, &dProbeRows, &dPmap, &dPconst, &dPinst, &dIinst,
			&dA, &dB, &dX, &dS, &dIS, &dCD, &dCI, &dHt, &dHx, &dOutGrid, &dOutT, &dStatusT, &dOutRaw,
			&dWs, &dRemaining};
		for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) all[i]->release();
		if (hRemaining) cudaFreeHost(hRemaining);
		if (ev0) cudaEventDestroy(ev0);
		if (ev1) cudaEventDestroy(ev1);
		if (ev2) cudaEventDestroy(ev2);
	}
};

/* ------------------------------------------------------------------------ *
 * Builders.  Same argument lists, checks and row-creation order as the
 * reference's simulatorAdd* + device*Config functions (cited per function).
 * ------------------------------------------------------------------------ */
namespace {

bool canAdd(simcu_ *r, const char *refdes)
{
	if (r == nullptr || refdes == nullptr) return false;
	if (r->locked) {      /* core/simulator.c:296 */
		fail("Can't add a device to a simulator that has run.");
		return false;
	}
	return true;
}

/* deviceNew2Pins / deviceNew4Pins: core/device.c:210-300 */
Dev makeDev(simcu_ *r, int kind, const char *refdes, int npins, char *const pins[])
{
	Dev d;
	d.kind = kind;
	d.refdes = refdes;
	for (int i = 0; i < npins; i++) d.pin[i] = r->nl.findOrAddRow('v', pins[i]);
	return d;
}

/* listAdd(devices, deviceCheckDuplicate): core/device.c:166-177 */
int commit(simcu_ *r, const Dev &d)
{
	if (r->nl.duplicate(d.refdes.c_str())) return fail("%s is listed twice", d.refdes.c_str());
	r->nl.devs.push_back(d);
	r->compiled = false; r->prepared = false;
	return 0;
}

struct BCtx { simcu_ *r; Dev *d; };

/* deviceNonlinearGetVariable: devices/nonlinear_i.c:147-192, _v.c:119-160,
 * _c.c:127-169.  Registers the Jacobian entries of one referenced unknown. */
int bGetVar(void *ctx, const char *name)
{
	BCtx *b = (BCtx *)ctx;
	Dev &d = *b->d;
	Netlist &nl = b->r->nl;
	if (!std::strcmp(name, "time")) { d.usesTime = true; return CALC_VAR_TIME + 0x40000000; }
	if (!((name[0] == 'i' || name[0] == 'v') && name[1] == '(')) {
		fail("B element variables must be i(...), v(...) or time -- not %s", name);
		return -1;
	}
	const int row = nl.findOrAddRow(0, name);
	if (row < 0) return -1;
	if (d.kind == DK_BI) { nl.addNode(d.pin[1], row); nl.addNode(d.rowM, row); }
	else nl.addNode(d.rowR, row);
	d.bvRow.push_back(row);
	return row;
}

}  /* namespace */

extern "C" {

simcu_ *simcuNew(simcu_ *r, int numInstances)
{
	if (r != nullptr) { fail("simcuNew: argument must be NULL"); return nullptr; }
	if (numInstances < 1) { fail("simcuNew: numInstances must be >= 1"); return nullptr; }
	simcu_ *s = new _simcu();
	s->M = numInstances;
	return s;
}

int simcuDestroy(simcu_ **r)
{
	if (r == nullptr || *r == nullptr) return -1;
	for (size_t i = 0; i < (*r)->variableNames.size(); i++) std::free((*r)->variableNames[i]);
	delete *r;
	*r = nullptr;
	return 0;
}

int simcuInfo(void)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	std::printf("simulator_cuda %d.%d: %d CUDA device(s)%s\n", SIMCU_MAJOR_VERSION,
			SIMCU_MINOR_VERSION, (e == cudaSuccess) ? n : 0,
			(e == cudaSuccess) ? "" : " (no driver)");
	return 0;
}

int simcuSetDevice(int device)
{
	CU(cudaSetDevice(device));
	return 0;
}

void simcuFree(void *p) { std::free(p); }

/* simulatorAddResistor: core/simulator.c:288-309, devices/resistor.c:111-144 */
int simcuAddResistor(simcu_ *r, char *refdes, char *pNode, char *nNode, double *resistance)
{
	if (!canAdd(r, refdes) || !pNode || !nNode || !resistance) return -1;
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, DK_R, refdes, 2, pins);
	d.val = resistance;
	const int K = d.pin[0], J = d.pin[1];
	r->nl.addNode(K, K); r->nl.addNode(K, J); r->nl.addNode(J, K); r->nl.addNode(J, J);
	return commit(r, d);
}

/* simulatorAddCapacitor: core/simulator.c:313-334, devices/capacitor.c:216-260 */
int simcuAddCapacitor(simcu_ *r, char *refdes, char *pNode, char *nNode, double *capacitance)
{
	if (!canAdd(r, refdes) || !pNode || !nNode || !capacitance) return -1;
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, DK_C, refdes, 2, pins);
	d.val = capacitance;
	const int K = d.pin[0], J = d.pin[1];
	r->nl.addNode(K, K); r->nl.addNode(J, K); r->nl.addNode(K, J); r->nl.addNode(J, J);
	return commit(r, d);
}

/* simulatorAddInductor: core/simulator.c:338-359, devices/inductor.c:222-269 */
int simcuAddInductor(simcu_ *r, char *refdes, char *pNode, char *nNode, double *inductance)
{
	if (!canAdd(r, refdes) || !pNode || !nNode || !inductance) return -1;
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, DK_L, refdes, 2, pins);
	d.val = inductance;
	d.rowR = r->nl.findOrAddRow('i', refdes);
	const int K = d.pin[0], J = d.pin[1], R = d.rowR;
	r->nl.addNode(R, K); r->nl.addNode(R, J); r->nl.addNode(K, R); r->nl.addNode(J, R);
	r->nl.addNode(R, R);
	return commit(r, d);
}

/* simulatorAddSource: core/simulator.c:430-457, devices/source_v.c:202-252,
 * source_i.c:195-238, math/waveform.c:95-168 */
int simcuAddSource(simcu_ *r, char *refdes, char *pNode, char *nNode, char type,
		double *dc, char stimulus, double *args[7])
{
	if (!canAdd(r, refdes) || !pNode || !nNode) return -1;
	type = (char)std::tolower((unsigned char)type);
	if (type != 'v' && type != 'i') return fail("Source type must be either i or v, not %c", type);
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, (type == 'v') ? DK_V : DK_I, refdes, 2, pins);
	if (type == 'v' && d.pin[0] == 0 && d.pin[1] == 0)
		return fail("Source %s has both nodes shorted to 0!", refdes);
	d.val = dc;
	d.wtype = stimulus;
	if (stimulus != 0) {
		if (!std::strchr("pgselcf", stimulus)) return fail("Unsupported waveform type %c", stimulus);
		if (args == nullptr) return fail("Source %s: waveform arguments missing", refdes);
		static const struct { char t; int n; } used[] = {
			{'p', 7}, {'g', 7}, {'s', 5}, {'e', 6}, {'l', 2}, {'c', 2}, {'f', 5}};
		int n = 0;
		for (size_t i = 0; i < sizeof used / sizeof used[0]; i++) if (used[i].t == stimulus) n = used[i].n;
		for (int i = 0; i < n; i++)
			if (args[i] == nullptr) return fail("Argument %i is NULL", i);
		if (stimulus == 'l' || stimulus == 'c') {
			d.wtab = (double **)args[0];
			d.wtabLen = (int *)args[1];
		} else {
			for (int i = 0; i < n; i++) d.wargs[i] = args[i];
		}
	} else if (dc == nullptr) {
		return fail("Source %s needs a DC value", refdes);
	}
	if (type == 'v') {
		d.rowR = r->nl.findOrAddRow('i', refdes);
		const int K = d.pin[0], J = d.pin[1], R = d.rowR;
		r->nl.addNode(R, K); r->nl.addNode(R, J); r->nl.addNode(K, R); r->nl.addNode(J, R);
	}
	return commit(r, d);
}

/* simulatorAddTLine: core/simulator.c:461-487, devices/tline.c:353-457 */
int simcuAddTLine(simcu_ *r, char *refdes, char *pNodeLeft, char *nNodeLeft,
		char *pNodeRight, char *nNodeRight, double *Z0, double *Td, double *loss)
{
	if (!canAdd(r, refdes) || !pNodeLeft || !nNodeLeft || !pNodeRight || !nNodeRight ||
			!Z0 || !Td) return -1;
	char *pins[4] = {pNodeLeft, nNodeLeft, pNodeRight, nNodeRight};
	Dev d = makeDev(r, DK_T, refdes, 4, pins);
	if (d.pin[0] == 0 && d.pin[1] == 0)
		return fail("T-Line %s has both input nodes shorted to 0!", refdes);
	if (d.pin[2] == 0 && d.pin[3] == 0)
		return fail("T-Line %s has both output nodes shorted to 0!", refdes);
	d.z0 = Z0; d.td = Td; d.loss = loss;
	d.rowR = r->nl.findOrAddRow('i', (std::string(refdes) + "#in1").c_str());
	d.rowM = r->nl.findOrAddRow('i', (std::string(refdes) + "#in2").c_str());
	const int K = d.pin[0], J = d.pin[1], L = d.pin[2], M = d.pin[3], R = d.rowR, S = d.rowM;
	const int rc[18][2] = {{R, K}, {R, J}, {K, R}, {J, R}, {R, L}, {R, M}, {S, L}, {S, M},
		{L, S}, {M, S}, {S, K}, {S, J}, {R, R}, {S, S}, {K, S}, {J, S}, {L, R}, {M, R}};
	for (int i = 0; i < 18; i++) r->nl.addNode(rc[i][0], rc[i][1]);
	return commit(r, d);
}

/* simulatorAddVICurve: core/simulator.c:512-538, devices/vicurve.c:298-359 */
int simcuAddVICurve(simcu_ *r, char *refdes, char *pNode, char *nNode, double **vi,
		int *viLength, char viType, double **ta, int *taLength, char taType)
{
	if (!canAdd(r, refdes) || !pNode || !nNode) return -1;
	if (!vi || !*vi || !viLength || *viLength < 1) return fail("VI curve %s needs a table", refdes);
	if (viType != 'l' && viType != 'c') return fail("VI table type must be l or c");
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, DK_VI, refdes, 2, pins);
	TableSet s; s.vi = vi; s.viLen = viLength; s.ta = ta; s.taLen = taLength;
	d.sets.push_back(s);
	d.viType = viType;
	d.taType = (taType == 'c') ? 'c' : 'l';
	d.rowR = r->nl.findOrAddRow('i', refdes);
	d.rowM = r->nl.findOrAddRow('v', refdes);
	const int K = d.pin[0], J = d.pin[1], R = d.rowR, M = d.rowM;
	r->nl.addNode(R, K); r->nl.addNode(R, M); r->nl.addNode(K, R); r->nl.addNode(M, R);
	r->nl.addNode(J, J); r->nl.addNode(J, M); r->nl.addNode(M, J); r->nl.addNode(M, M);
	return commit(r, d);
}

/* simulatorAddNonlinearSource: core/simulator.c:363-392, devices/
 * nonlinear_i.c:368-425, nonlinear_v.c:330-392, nonlinear_c.c:429-495.  The
 * reference tokenises the equation here and parses it at every evaluation; a
 * syntax error therefore surfaces at the first run there and already here. */
int simcuAddNonlinearSource(simcu_ *r, char *refdes, char *pNode, char *nNode, char type,
		char *equation)
{
	if (!canAdd(r, refdes) || !pNode || !nNode || !equation) return -1;
	type = (char)std::tolower((unsigned char)type);
	if (type != 'i' && type != 'v' && type != 'c')
		return fail("Nonlinear Source type must be either i, v or c, not %c", type);
	char *pins[2] = {pNode, nNode};
	Dev d = makeDev(r, (type == 'i') ? DK_BI : (type == 'v') ? DK_BV : DK_BC, refdes, 2, pins);
	const int K = d.pin[0], J = d.pin[1];
	Netlist &nl = r->nl;
	if (type == 'i') {
		d.rowR = nl.findOrAddRow('i', refdes);
		d.rowM = nl.findOrAddRow('v', refdes);
		nl.addNode(d.rowR, K); nl.addNode(d.rowR, d.rowM);
		nl.addNode(K, d.rowR); nl.addNode(d.rowM, d.rowR);
	} else if (type == 'v') {
		if (K == 0 && J == 0) return fail("Source %s has both input nodes shorted to 0!", refdes);
		d.rowR = nl.findOrAddRow('i', refdes);
		nl.addNode(d.rowR, K); nl.addNode(d.rowR, J);
		nl.addNode(K, d.rowR); nl.addNode(J, d.rowR);
	} else {
		nl.addNode(K, K); nl.addNode(J, K); nl.addNode(K, J); nl.addNode(J, J);
		d.rowR = nl.findOrAddRow('i', refdes);
		d.rowM = nl.findOrAddRow('v', refdes);
		nl.addNode(d.rowR, d.rowM); nl.addNode(d.rowM, d.rowR);
	}
	BCtx ctx; ctx.r = r; ctx.d = &d;
	std::string err;
	if (!calcCompile(equation, bGetVar, &ctx, d.calc, err))
		return fail("Bad B equation for %s (%s): %s", refdes, err.c_str(), equation);
	/* calc variable index of each Jacobian variable; "time" has no Jacobian */
	for (size_t i = 0; i < d.calc.varRef.size(); i++) {
		if (d.calc.varRef[i] == CALC_VAR_TIME + 0x40000000) d.calc.varRef[i] = CALC_VAR_TIME;
		else d.bvCalcVar.push_back((int)i);
	}
	return commit(r, d);
}

int simcuAddCallbackSource(simcu_ *, char *refdes, char *, char *, char, char *[], double[],
		double[], int, void *, void *)
{
	return fail("%s: callback (PyB) sources call back into the interpreter at every Newton "
			"iteration and cannot be batched", refdes ? refdes : "?");
}

int simcuAddTLineW(simcu_ *, char *refdes, char *[], int, int *, double *, double **, double **,
		double **, double **, double **, double **, double *, double *)
{
	return fail("%s: the W-element is unfinished in the reference and not supported",
			refdes ? refdes : "?");
}

/* ---- the instance dimension ------------------------------------------------ */
int simcuSetInstanceParam(simcu_ *r, char *refdes, char *name, const double *values, int stride)
{
	if (!r || !refdes || !name || !values) return -1;
	Dev *d = r->nl.find(refdes);
	if (!d) return fail("unknown device %s", refdes);
	bool ok = false;
	const std::string n(name);
	switch (d->kind) {
	case DK_R: ok = (n == "R"); break;
	case DK_C: ok = (n == "C"); break;
	case DK_L: ok = (n == "L"); break;
	case DK_T: ok = (n == "Z0" || n == "Td" || n == "loss" || n == "bypass"); break;
	case DK_V: case DK_I:
		ok = (n == "DC") || (n.size() == 2 && n[0] == 'A' && n[1] >= '0' && n[1] <= '6' &&
				d->wtype && d->wtype != 'l' && d->wtype != 'c');
		break;
	default: break;
	}
	if (!ok) return fail("unknown parameter %s.%s", refdes, name);
	const bool isNew = !d->inst.count(n);
	std::vector<double> &v = d->inst[n];
	v.resize(r->M);
	bool changed = isNew;
	for (int i = 0; i < r->M; i++) {
		const double x = values[(size_t)i * stride];
		changed = changed || std::memcmp(&v[i], &x, sizeof x) != 0;
		v[i] = x;
	}
	if (isNew) { r->compiled = false; r->prepared = false; }
	/* the pivot sequences were chosen on pilot instances of the old values */
	if (changed) r->paramsChanged = true;
	return 0;
}

int simcuAddVITableSet(simcu_ *r, char *refdes, double **vi, int *viLength, double **ta,
		int *taLength)
{
	if (!r || !refdes) return -1;
	if (r->locked) return fail("Can't add a table set to a simulator that has run.");
	Dev *d = r->nl.find(refdes);
	if (!d || d->kind != DK_VI) return fail("%s is not a VI device", refdes);
	if (!vi || !*vi || !viLength || *viLength < 1) return fail("VI curve %s needs a table", refdes);
	TableSet s; s.vi = vi; s.viLen = viLength; s.ta = ta; s.taLen = taLength;
	d->sets.push_back(s);
	r->compiled = false; r->prepared = false;
	return (int)d->sets.size() - 1;
}

int simcuAddSourceTableSet(simcu_ *r, char *refdes, double **table, int *length)
{
	if (!r || !refdes) return -1;
	if (r->locked) return fail("Can't add a table set to a simulator that has run.");
	Dev *d = r->nl.find(refdes);
	if (!d || (d->kind != DK_V && d->kind != DK_I) || (d->wtype != 'l' && d->wtype != 'c'))
		return fail("%s is not a PWL / PWC source", refdes);
	if (!table || !*table || !length || *length < 1) return fail("source %s needs a table", refdes);
	d->wsets.push_back(std::make_pair(table, length));
	r->compiled = false; r->prepared = false;
	return (int)d->wsets.size();
}

int simcuSetInstanceTableSet(simcu_ *r, char *refdes, const int *setIndex, int stride)
{
	if (!r || !refdes || !setIndex) return -1;
	Dev *d = r->nl.find(refdes);
	const bool source = d && (d->kind == DK_V || d->kind == DK_I) && (d->wtype == 'l' || d->wtype == 'c');
	if (!d || (d->kind != DK_VI && !source)) return fail("%s is not a VI device or a PWL / PWC source", refdes);
	const bool isNew = d->instSet.empty();
	d->instSet.resize(r->M);
	bool changed = isNew;
	for (int i = 0; i < r->M; i++) {
		const int x = setIndex[(size_t)i * stride];
		changed = changed || d->instSet[i] != x;
		d->instSet[i] = x;
	}
	if (isNew) { r->compiled = false; r->prepared = false; }
	if (changed) r->paramsChanged = true;
	return 0;
}

int simcuSetOption(simcu_ *r, char *name, double value)
{
	if (!name) return -1;
	const std::string n(name);
	if (n == "quiet") { gQuiet = (value != 0.0); return 0; }
	if (!r) return -1;
	SimcuControl &c = r->ctl;
	if (n == "reltol") c.reltol = value; else if (n == "vntol") c.vntol = value;
	else if (n == "abstol") c.abstol = value; else if (n == "captol") c.captol = value;
	else if (n == "chgtol") c.chgtol = value; else if (n == "trtol") c.trtol = value;
	else if (n == "gmin") c.gmin = value; else if (n == "itl1") c.itl1 = (int)value;
	else if (n == "itl4") c.itl4 = (int)value; else if (n == "maxorder") c.maxorder = (int)value;
	else if (n == "minstep") r->minstepOpt = value;
	else if (n == "hist_records") r->histRecordsOpt = (int)value;
	else if (n == "max_attempts") r->maxAttempts = (long long)value;
	else if (n == "attempts_per_launch") r->attemptsPerLaunch = (value < 1) ? INT_MAX : (int)value;
	else if (n == "threads_per_block") r->tpbOpt = (int)value;
	else if (n == "workspace_shared") r->wsSharedOpt = (int)value;
	else if (n == "raw_max_points") r->rawMaxPoints = (int)value;
	else if (n == "raw_first") r->rawFirst = (int)value;
	else if (n == "raw_count") r->rawCount = (int)value;
	else if (n == "pivot_threshold") { r->pivotThreshold = value; r->sched[0].lu.clear(); r->sched[1].lu.clear(); }
	else if (n == "num_pilots") r->numPilots = std::max(1, (int)value);
	else if (n == "interpreter") r->interpreter = (value != 0.0);
	else if (n == "lanes_per_warp") r->lanesOpt = (int)value;
	else if (n == "keep_order") r->keepOrder = (value != 0.0);
	else if (n == "history_retry") r->historyRetry = (value != 0.0);
	else if (n == "min_blocks_per_sm") { r->minBlocks = (int)value; r->jit.reset(); }
	else if (n == "max_blocks_per_sm") r->maxBlocksPerSM = (int)value;
	else if (n == "claim_below") r->claimBelow = std::max(0, std::min(31, (int)value));
	else if (n == "replay") r->replay = (value != 0.0);
	else if (n == "block_sync") r->blockSync = (value < 0.0) ? -1 : (value != 0.0);
	else if (n == "sync_group") r->syncGroup = std::max(0, std::min(15, (int)value));
	else if (n == "smem_tables_kb") r->smemTablesKB = std::max(-1, std::min(200, (int)value));
	else if (n == "growth_limit") r->growthLimit = value;
	else if (n == "tline_prefetch") r->tlinePrefetch = (value != 0.0);
	else if (n == "fmad") r->fmad = (value != 0.0);
	else if (n == "solve_blocks") r->solveBlocks = std::max(0, std::min(2, (int)value));
	else if (n == "shared_break_clock") r->sharedCb = (value != 0.0);
	else return fail("unknown option %s", name);
	r->prepared = false;
	return 0;
}

int simcuSetPivots(simcu_ *r, int phase, const int *pc, const int *pr, int n)
{
	if (!r || phase < 0 || phase > 1) return -1;
	r->forcePc[phase].clear(); r->forcePr[phase].clear();
	if (pc && pr) {
		r->forcePc[phase].assign(pc, pc + n);
		r->forcePr[phase].assign(pr, pr + n);
	}
	r->prepared = false;
	return 0;
}

int simcuSetInstanceWindow(simcu_ *r, const double *tstep, const double *tstop, const double *tmax,
		int stride)
{
	if (!r) return -1;
	r->winTstep.clear(); r->winTstop.clear(); r->winTmax.clear();
	if (tstep && tstop) {
		for (int i = 0; i < r->M; i++) {
			const double ts = tstep[(size_t)i * stride], te = tstop[(size_t)i * stride];
			if (!(ts > 0.0) || !(te > 0.0)) return fail("instance %d: tstep and tstop must be positive", i);
			double tm = tmax ? tmax[(size_t)i * stride] : 0.0;
			if (tm == 0.0) { tm = te / 50; tm = (ts < tm) ? ts : tm; }      /* core/simulator.c:91-94 */
			r->winTstep.push_back(ts); r->winTstop.push_back(te); r->winTmax.push_back(tm);
		}
	}
	r->prepared = false;
	return 0;
}

int simcuSetReplayLog(simcu_ *r, const int *offsets, const double *times, const int *flags,
		const int *linOffsets, const int *linMasks)
{
	if (!r) return -1;
	r->replayOff.clear(); r->replayTime.clear(); r->replayFlag.clear();
	r->replayLinOff.clear(); r->replayLin.clear();
	if (offsets && times && flags && linOffsets && linMasks) {
		if (linOffsets[0] != 0) return fail("replay log: linOffsets[0] must be 0");
		for (int i = 0; i < r->M; i++)
			if (linOffsets[i + 1] < linOffsets[i]) return fail("replay log: offsets must not decrease");
		r->replayLinOff.assign(linOffsets, linOffsets + r->M + 1);
		r->replayLin.assign(linMasks, linMasks + linOffsets[r->M]);
		if (offsets[0] != 0) return fail("replay log: offsets[0] must be 0");
		for (int i = 0; i < r->M; i++)
			if (offsets[i + 1] < offsets[i]) return fail("replay log: offsets must not decrease");
		r->replayOff.assign(offsets, offsets + r->M + 1);
		r->replayTime.assign(times, times + offsets[r->M]);
		r->replayFlag.assign(flags, flags + offsets[r->M]);
	}
	r->prepared = false;
	return 0;
}

}  /* extern "C" */

/* ------------------------------------------------------------------------ *
 * Preparation: flatten, symbolic analysis on a pilot batch, NVRTC, allocate
 * ------------------------------------------------------------------------ */
namespace {

int ensureCompiled(simcu_ *r)
{
	std::string err;
	if (!r->compiled) {
		if (!compileProgram(r->nl, r->M, r->prog, err)) return fail("%s", err.c_str());
		r->compiled = true;
		r->sched[0] = Schedule(); r->sched[1] = Schedule();
		r->jit.reset();
	} else if (!refreshParameters(r->nl, r->M, r->prog, err)) {
		return fail("%s", err.c_str());
	}
	return 0;
}

int uploadParameters(simcu_ *r, cudaStream_t st)
{
	const Program &p = r->prog;
	if (r->dPmap.upload(p.pmap, st) || r->dPconst.upload(p.pconst, st) ||
			r->dPinst.upload(p.pinst, st) || r->dIinst.upload(p.iinst, st) ||
			r->dTabP.upload(p.tabP, st))
		return -1;
	/* table descriptors, and the tables every block stages in shared memory: V-I
	 * tables first (looked up at every Newton iteration), then TA, then source
	 * tables, in table order, as long as they fit the budget */
	{
		const int ntab = (int)p.tabType.size();
		std::vector<int> desc((size_t)4 * std::max(1, ntab), -1);
		std::vector<double> stage;
		const size_t budget = r->stageOK ? (size_t)r->smemTablesUsed * 1024 / 32 : 0;
		for (int t = 0; t < ntab; t++) {
			desc[4 * t] = p.tabOff[t]; desc[4 * t + 1] = p.tabOff[t + 1] - p.tabOff[t];
			desc[4 * t + 2] = p.tabType[t]; desc[4 * t + 3] = -1;
		}
		for (int klass = 0; klass < 3; klass++)
			for (int t = 0; t < ntab; t++) {
				const size_t len = (size_t)desc[4 * t + 1];
				if (p.tabClass[t] != klass || stage.size() / 4 + len > budget) continue;
				desc[4 * t + 3] = (int)(stage.size() / 4);
				stage.insert(stage.end(), p.tabP.begin() + 4 * (size_t)p.tabOff[t],
						p.tabP.begin() + 4 * (size_t)p.tabOff[t + 1]);
			}
		r->stagePoints = (int)(stage.size() / 4);
		if (r->dTabDesc.upload(desc, st) || r->dTabStage.upload(stage, st)) return -1;
	}
	/* processing order: stable sort by the tuple of table-set indices */
	const int *order = nullptr;
	if (p.niinst > 0 && !r->keepOrder && r->orderKey == p.iinst && r->dOrder.p) {
		order = r->dOrder.as<int>();          /* table sets unchanged since the last upload */
	} else if (p.niinst > 0 && !r->keepOrder) {
		const int M = r->M;
		r->orderKey = p.iinst;
		std::vector<int> ord(M);
		for (int i = 0; i < M; i++) ord[i] = i;
		std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) {
			for (int k = 0; k < p.niinst; k++) {
				const int vx = p.iinst[(size_t)k * M + x], vy = p.iinst[(size_t)k * M + y];
				if (vx != vy) return vx < vy;
			}
			return false;
		});
		if (r->dOrder.upload(ord, st)) return -1;
		order = r->dOrder.as<int>();
	}
	for (int ph = 0; ph < 2; ph++) {
		SimcuArgs &a = r->args[ph];
		a.order = order;
		a.pmap = r->dPmap.as<int>(); a.pconst = r->dPconst.as<double>();
		a.pinst = r->dPinst.as<double>(); a.iinst = r->dIinst.as<int>();
		a.tabP = r->dTabP.as<double>();
		a.tabDesc = r->dTabDesc.as<int>(); a.tabStage = r->dTabStage.as<double>();
		a.stagePoints = r->stagePoints;
	}
	return 0;
}

int uploadSchedule(simcu_ *r, int ph)
{
	const Schedule &s = r->sched[ph];
	if (r->dLu[ph].upload(s.lu) || r->dXslot[ph].upload(s.xslot)) return -1;
	SimcuArgs &a = r->args[ph];
	a.lu = r->dLu[ph].as<int>(); a.luLen = (int)s.lu.size();
	a.xslot = r->dXslot[ph].as<int>();
	a.F = s.F;
	return 0;
}

/* generic kernels: threads per block and where their LU workspace lives */
int genericShape(simcu_ *r, int M, int wantTpb, int *tpbOut, int *sharedOut, DBuf &ws)
{
	const int words = std::max(r->sched[0].F, r->sched[1].F) + r->prog.N;
	int dev = 0, maxSmem = 0;
	CU(cudaGetDevice(&dev));
	CU(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
	int tpb = wantTpb > 0 ? wantTpb : 64;
	tpb = std::max(32, (tpb / 32) * 32);
	int shared = (r->wsSharedOpt < 0) ? 1 : r->wsSharedOpt;
	if (shared) {
		while (tpb > 32 && (size_t)words * tpb * 8 > (size_t)maxSmem) tpb -= 32;
		if ((size_t)words * tpb * 8 > (size_t)maxSmem) {
			if (r->wsSharedOpt == 1) return fail("LU workspace does not fit shared memory");
			shared = 0;
			tpb = wantTpb > 0 ? std::max(32, (wantTpb / 32) * 32) : 128;
		}
	}
	if (!shared && ws.reserve((size_t)words * M * 8)) return -1;
	*tpbOut = tpb; *sharedOut = shared;
	return 0;
}

/* The pilot pass: a compact batch of up to numPilots instances spread over the
 * sweep runs through the generic kernels - load, then (for the transient
 * phase) operating point and the assembly of the first attempt - and the
 * matrices it leaves behind decide the pivot sequences of both phases. */
struct Pilot {
	DBuf A, B, X, S, IS, CD, CI, Ht, Hx, pinst, iinst, remaining, ws;
};

int analyseFrom(simcu_ *r, int ph, const DBuf &dA, int P)
{
	const Program &p = r->prog;
	std::vector<double> flat((size_t)p.nnzA * P);
	CU(cudaMemcpy(flat.data(), dA.p, flat.size() * 8, cudaMemcpyDeviceToHost));
	std::vector<std::vector<double> > pilots;
	for (int q = 0; q < P; q++) {
		std::vector<double> v(p.nnzA);
		bool finite = true;
		for (int s = 0; s < p.nnzA; s++) { v[s] = flat[(size_t)s * P + q]; finite = finite && std::isfinite(v[s]); }
		/* an instance with a non-finite stamp (R = 0, NaN parameter) fails on
		 * its own later; it must not decide the pivots of the others */
		if (finite) pilots.push_back(v);
	}
	if (pilots.empty()) return fail("every pilot instance has non-finite matrix entries");
	/* a slot the flattener believes is exactly zero in this phase must be */
	std::vector<char> active = ph ? p.activeTran : p.activeOP;
	for (size_t q = 0; q < pilots.size(); q++)
		for (int s = 0; s < p.nnzA; s++)
			if (!active[s] && pilots[q][s] != 0.0) active[s] = 1;
	r->active[ph] = active;
	/* a slot the flattener believes is a literal must hold it in every pilot */
	std::vector<double> cst = ph ? p.constTran : p.constOP;
	for (size_t q = 0; q < pilots.size(); q++)
		for (int s = 0; s < p.nnzA; s++)
			if (cst[s] == cst[s] && pilots[q][s] != cst[s]) cst[s] = NAN;
	r->consts[ph] = cst;
	const bool forced = !r->forcePc[ph].empty();
	if (forced && (int)r->forcePc[ph].size() != p.N) return fail("forced pivot sequence has the wrong length");
	std::string err;
	if (!analyse(p.N, p.colStart, p.rowIdx, active, pilots, r->pivotThreshold,
			forced ? r->forcePc[ph].data() : nullptr, forced ? r->forcePr[ph].data() : nullptr,
			r->sched[ph], err))
		return fail("%s analysis: %s", ph ? "transient" : "operating-point", err.c_str());
	r->jit.reset();
	return uploadSchedule(r, ph);
}

int pilotPass(simcu_ *r, bool needOP, bool needTran)
{
	const Program &p = r->prog;
	const int M = r->M;
	int P = std::min(M, r->numPilots);
	std::vector<int> idx(P);
	for (int q = 0; q < P; q++) idx[q] = (P > 1) ? (int)((long long)q * (M - 1) / (P - 1)) : 0;
	/* every value of a categorical column (IBIS table set, T-line bypass flag) must
	 * be seen by the analysis: one fixed sequence has to serve all of them */
	{
		const size_t cap = 64;
		for (int k = 0; k < p.niinst && idx.size() < cap; k++) {
			std::set<int> seen;
			for (size_t q = 0; q < idx.size(); q++) seen.insert(p.iinst[(size_t)k * M + idx[q]]);
			for (int i = 0; i < M && idx.size() < cap; i++)
				if (seen.insert(p.iinst[(size_t)k * M + i]).second) idx.push_back(i);
		}
		for (size_t c = 0; c < p.catInst.size() && idx.size() < cap; c++) {
			const size_t k = (size_t)p.catInst[c];
			std::set<double> seen;
			for (size_t q = 0; q < idx.size(); q++) seen.insert(p.pinst[k * M + idx[q]]);
			for (int i = 0; i < M && idx.size() < cap; i++)
				if (seen.insert(p.pinst[k * M + i]).second) idx.push_back(i);
		}
		P = (int)idx.size();
	}
	std::vector<double> pin((size_t)p.ninst * P);
	std::vector<int> iin((size_t)p.niinst * P);
	for (int k = 0; k < p.ninst; k++)
		for (int q = 0; q < P; q++) pin[(size_t)k * P + q] = p.pinst[(size_t)k * M + idx[q]];
	for (int k = 0; k < p.niinst; k++)
		for (int q = 0; q < P; q++) iin[(size_t)k * P + q] = p.iinst[(size_t)k * M + idx[q]];
	Pilot b;
	const int nh = (int)p.histRows.size(), R = 8;
	const size_t m8 = (size_t)P * 8, m4 = (size_t)P * 4;
	int rc = -1;
	do {
		if (b.pinst.upload(pin) || b.iinst.upload(iin) || b.A.reserve(p.nnzA * m8) ||
				b.B.reserve(p.N * m8) || b.X.reserve(p.N * m8) ||
				b.S.reserve(std::max(1, p.stateDoubles) * m8) ||
				b.IS.reserve(std::max(1, p.stateInts) * m4) || b.CD.reserve(CD_NUM * m8) ||
				b.CI.reserve(CI_NUM * m4) || b.Ht.reserve(R * m8) ||
				b.Hx.reserve((size_t)R * std::max(1, nh) * m8) || b.remaining.reserve(sizeof(int)))
			break;
		if (cudaMemset(b.S.p, 0, std::max(1, p.stateDoubles) * m8) != cudaSuccess ||
				cudaMemset(b.IS.p, 0, std::max(1, p.stateInts) * m4) != cudaSuccess)
			break;
		SimcuArgs a = r->args[0];
		a.M = P; a.count = P; a.order = nullptr; a.pinst = b.pinst.as<double>(); a.iinst = b.iinst.as<int>();
		a.A = b.A.as<double>(); a.B = b.B.as<double>(); a.X = b.X.as<double>();
		a.S = b.S.as<double>(); a.IS = b.IS.as<int>(); a.CD = b.CD.as<double>();
		a.CI = b.CI.as<int>(); a.Ht = b.Ht.as<double>(); a.Hx = b.Hx.as<double>();
		a.histRecords = R; a.nprobes = 0; a.ngrid = 0; a.rawMaxPoints = 0;
		a.remaining = b.remaining.as<int>();
		a.attemptsPerLaunch = 1; a.stopBeforeSolve = 0;
		if (simcuLaunchLoad(&a, 32, 0)) { fail("load kernel launch failed"); break; }
		if (cudaDeviceSynchronize() != cudaSuccess) { fail("pilot load kernel failed"); break; }
		if (needOP && analyseFrom(r, 0, b.A, P)) break;
		if (needTran) {
			if (r->sched[1].lu.empty()) r->sched[1].F = p.nnzA;
			int tpb = 32, shared = 1;
			if (genericShape(r, P, 32, &tpb, &shared, b.ws)) break;
			a.lu = r->args[0].lu; a.luLen = r->args[0].luLen; a.xslot = r->args[0].xslot;
			a.F = r->args[0].F; a.wsShared = shared; a.wsGlobal = b.ws.as<double>();
			if (simcuLaunchOp(&a, tpb, 0, 0)) { fail("op kernel launch failed"); break; }
			SimcuArgs t = a;
			t.stopBeforeSolve = 1; t.F = 0; t.lu = nullptr; t.xslot = nullptr;
			if (simcuLaunchTran(&t, tpb, 0)) { fail("tran kernel launch failed"); break; }
			if (cudaDeviceSynchronize() != cudaSuccess) { fail("pilot kernels failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
			if (analyseFrom(r, 1, b.A, P)) break;
		}
		rc = 0;
	} while (0);
	DBuf *all[] = {&b.A, &b.B, &b.X, &b.S, &b.IS, &b.CD, &b.CI, &b.Ht, &b.Hx, &b.pinst, &b.iinst,
		&b.remaining, &b.ws};
	for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) all[i]->release();
	return rc;
}

/* process-wide cache of compiled topologies, keyed by the generated source */
std::map<std::string, std::shared_ptr<JitKernel> > gKernelCache;
std::mutex gKernelCacheLock;
const size_t kKernelCacheMax = 64;      /* topologies kept loaded per process */

int buildJitKernel(simcu_ *r)
{
	JitConfig cfg;
	cfg.threadsPerBlock = r->tpb; cfg.minBlocks = r->minBlocksUsed; cfg.gmin = r->ctl.gmin;
	cfg.replay = r->replay; cfg.blockSync = r->blockSyncUsed; cfg.syncGroup = r->syncGroup;
	cfg.windows = !r->winTstep.empty(); cfg.growthLimit = r->growthLimit; cfg.tlinePrefetch = r->tlinePrefetch; cfg.fmad = r->fmad;
	cfg.smemTab = (r->stagePoints > 0);
	cfg.solveBlocks = r->solveBlocks; cfg.sharedCb = r->sharedCb;
	if (r->jit && r->jitProbes == r->probeRows && r->jitCfg == cfg && r->jitMeasures == r->measures)
		return 0;
	const std::string src = assembleSource(r->prog, r->sched, r->active, r->consts, r->probeRows,
			r->measures, cfg);
	const std::string &key = src;
	std::lock_guard<std::mutex> guard(gKernelCacheLock);
	std::map<std::string, std::shared_ptr<JitKernel> >::iterator it = gKernelCache.find(key);
	if (it != gKernelCache.end()) r->jit = it->second;
	else {
		std::shared_ptr<JitKernel> k(new JitKernel());
		std::string err;
		const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		if (!jitCompile(src, *k, err, cfg.fmad != 0)) return fail("%s", err.c_str());
		r->stats.jitCompileMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		if (!jitLoad(*k, r->tpb, err)) return fail("%s", err.c_str());
		/* bounded: drop entries no handle uses any more, oldest key first */
		for (it = gKernelCache.begin(); it != gKernelCache.end() && gKernelCache.size() >= kKernelCacheMax;) {
			if (it->second.use_count() == 1) it = gKernelCache.erase(it); else ++it;
		}
		gKernelCache[key] = k;
		r->jit = k;
		r->stats.jitCompiles++;
	}
	r->jitProbes = r->probeRows; r->jitCfg = cfg;
	r->jitMeasures = r->measures;
	return 0;
}

int resetState(simcu_ *r, cudaStream_t st)
{
	const Program &p = r->prog;
	CU(cudaMemsetAsync(r->dS.p, 0, (size_t)std::max(1, p.stateDoubles) * r->M * 8, st));
	CU(cudaMemsetAsync(r->dIS.p, 0, (size_t)std::max(1, p.stateInts) * r->M * 4, st));
	return 0;
}

int prepare(simcu_ *r, double tstep, double tstop, double tmax, char *probes[], int numProbes,
		bool opOnly)
{
	if (!r) return -1;
	if (r->nl.devs.empty()) return fail("empty circuit");
	int ndevice = 0;
	if (cudaGetDeviceCount(&ndevice) != cudaSuccess || ndevice < 1)
		return fail("no CUDA device: the batched engine has no CPU path");
	if (ensureCompiled(r)) return -1;
	r->locked = true;
	const Program &p = r->prog;
	const int M = r->M;

	/* run constants: core/simulator.c:91-109 */
	SimcuControl c = r->ctl;
	if (opOnly) { tstep = 1.0; tstop = 0.0; tmax = 1.0; }
	else {
		if (!(tstep > 0.0) || !(tstop > 0.0)) return fail("tstep and tstop must be positive");
		if (tmax == 0.0) { tmax = tstop / 50; tmax = (tstep < tmax) ? tstep : tmax; }
	}
	c.tstep = tstep; c.tstop = tstop; c.tmax = tmax;
	c.minstep = (r->minstepOpt < 0) ? tstep * 5e-5 : r->minstepOpt;
	const bool runChanged = !r->prepared || c.tstep != r->ctl.tstep || c.tstop != r->ctl.tstop ||
		c.tmax != r->ctl.tmax;
	r->ctl = c;

	/* probes and output grid */
	r->probeNames.clear(); r->probeRows.clear();
	for (int i = 0; i < numProbes; i++) {
		const int row = r->nl.findRow(probes[i]);
		if (row < 1) return fail("unknown probe %s", probes[i] ? probes[i] : "(null)");
		r->probeNames.push_back(probes[i]);
		r->probeRows.push_back(row);
	}
	r->ngrid = 0;
	if (!opOnly && numProbes > 0) {
		int ng = (int)std::floor(tstop / tstep * (1.0 + 1e-12)) + 1;
		if ((ng - 1) * tstep < tstop * (1.0 - 1e-9)) ng++;
		r->ngrid = ng;
	}

	/* program upload (read by the generic kernels; tables and parameters by both) */
	if (r->dDev.upload(p.dev) || r->dBvar.upload(p.bvar) || r->dCalcOff.upload(p.calcOff) ||
			r->dCalcCode.upload(p.calcCode) || r->dCalcConst.upload(p.calcConst) ||
			r->dCalcVarOff.upload(p.calcVarOff) || r->dCalcVars.upload(p.calcVars) ||
			r->dHistRows.upload(p.histRows) || r->dProbeRows.upload(r->probeRows))
		return -1;
	const int nh = (int)p.histRows.size();
	const size_t m8 = (size_t)M * 8, m4 = (size_t)M * 4;
	if (r->dCI.reserve(CI_NUM * m4) || r->dRemaining.reserve(2 * sizeof(int)) ||
			r->dOutGrid.reserve((size_t)r->ngrid * numProbes * m8))
		return -1;
	const int nmeas = (int)r->measures.size() / 2;
	if (nmeas > 0 && r->interpreter) return fail("measures need the specialised kernel (interpreter = 0)");
	if (r->dMeasure.reserve((size_t)std::max(1, nmeas) * m8) || r->dMeasureLevel.upload(r->measureLevels))
		return -1;
	if (r->rawMaxPoints > 0) {
		if (r->rawCount < 1) { r->rawFirst = 0; r->rawCount = M; }
		if (r->rawFirst < 0 || r->rawFirst + r->rawCount > M) return fail("raw_first/raw_count out of range");
		if (r->dOutRaw.reserve((size_t)r->rawMaxPoints * (p.N + 1) * r->rawCount * 8)) return -1;
	}
	if (!r->hRemaining) CU(cudaMallocHost((void **)&r->hRemaining, sizeof(int)));
	if (!r->ev0) {
		CU(cudaEventCreate(&r->ev0)); CU(cudaEventCreate(&r->ev1)); CU(cudaEventCreate(&r->ev2));
	}

	for (int ph = 0; ph < 2; ph++) {
		SimcuArgs &a = r->args[ph];
		a.M = M; a.N = p.N; a.nnzA = p.nnzA; a.ndev = p.ndev; a.ctl = r->ctl;
		a.dev = r->dDev.as<int>(); a.bvar = r->dBvar.as<int>();
		a.calcOff = r->dCalcOff.as<int>(); a.calcCode = r->dCalcCode.as<int>();
		a.calcConst = r->dCalcConst.as<double>();
		a.calcVarOff = r->dCalcVarOff.as<int>(); a.calcVars = r->dCalcVars.as<int>();
		a.histRows = r->dHistRows.as<int>(); a.nh = nh;
		a.probeRows = r->dProbeRows.as<int>(); a.nprobes = numProbes; a.ngrid = r->ngrid;
		a.rawMaxPoints = r->rawMaxPoints; a.rawFirst = r->rawFirst; a.rawCount = r->rawCount;
		a.CI = r->dCI.as<int>();
		a.outGrid = r->dOutGrid.as<double>(); a.outRaw = r->dOutRaw.as<double>();
		a.outMeasure = r->dMeasure.as<double>(); a.measureLevel = r->dMeasureLevel.as<double>();
		a.attemptsPerLaunch = r->attemptsPerLaunch;
		a.maxAttempts = (r->maxAttempts > 0) ? r->maxAttempts : LLONG_MAX;
		a.remaining = r->dRemaining.as<int>();
		a.counter = r->dRemaining.as<int>() + 1;
		a.stopBeforeSolve = 0;
	}
	{
		/* tables are staged in shared memory only by the one-big-block-per-SM shapes
		 * (several small blocks per SM would each need their own copy) */
		int smsEarly = 148, devEarly = 0;
		if (cudaGetDevice(&devEarly) == cudaSuccess)
			cudaDeviceGetAttribute(&smsEarly, cudaDevAttrMultiProcessorCount, devEarly);
		const bool fullEarly = (long long)M >= (long long)smsEarly * 256;
		const int wsEarly = p.stateDoubles + p.nnzA + 3 * p.N + (int)p.pmap.size();
		/* automatic: 100 kB (all V-I tables of the IBIS links) for the 256-thread /
		 * 255-register shape, where L1 can spare it (cfg4 +1 %); none for the
		 * 896-thread shape, whose spilled state needs all of L1 (cfg5 -5 %) */
		r->smemTablesUsed = (r->smemTablesKB >= 0) ? r->smemTablesKB
			: ((r->tpbOpt == 0 && fullEarly && p.N > 12 && wsEarly < 600) ? 100 : 0);
		r->stageOK = !r->interpreter && r->smemTablesUsed > 0 &&
			((r->tpbOpt == 0 && fullEarly && p.N > 12) || r->tpbOpt >= 256);
	}
	if (uploadParameters(r, 0)) return -1;
	if (uploadSchedule(r, 0) || uploadSchedule(r, 1)) return -1;
	if (r->replay) {
		if (r->interpreter) return fail("replay needs the specialised kernel (interpreter = 0)");
		if ((int)r->replayOff.size() != M + 1) return fail("replay = 1 without a log (simcuSetReplayLog)");
		if (r->dReplayOff.upload(r->replayOff) || r->dReplayTime.upload(r->replayTime) ||
				r->dReplayFlag.upload(r->replayFlag) || r->dReplayLinOff.upload(r->replayLinOff) ||
				r->dReplayLin.upload(r->replayLin))
			return -1;
	}
	if (!r->winTstep.empty()) {
		if (r->interpreter) return fail("per-instance windows need the specialised kernel (interpreter = 0)");
		std::vector<double> win((size_t)4 * M);
		for (int i = 0; i < M; i++) {
			win[i] = r->winTstep[i]; win[(size_t)M + i] = r->winTstop[i]; win[(size_t)2 * M + i] = r->winTmax[i];
			win[(size_t)3 * M + i] = (r->minstepOpt < 0) ? r->winTstep[i] * 5e-5 : r->minstepOpt;
		}
		if (r->dWin.upload(win)) return -1;
	}
	for (int ph = 0; ph < 2; ph++) {
		SimcuArgs &a = r->args[ph];
		a.winst = r->winTstep.empty() ? nullptr : r->dWin.as<double>();
		a.claimBelow = r->claimBelow;
		a.replayOff = r->dReplayOff.as<int>(); a.replayTime = r->dReplayTime.as<double>();
		a.replayFlag = r->dReplayFlag.as<int>();
		a.replayLinOff = r->dReplayLinOff.as<int>(); a.replayLin = r->dReplayLin.as<int>();
	}
	int sms = 148, dev = 0;
	CU(cudaGetDevice(&dev));
	CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

	/* symbolic analysis on the pilot batch (one-off per topology and time scale) */
	const bool needOP = r->sched[0].lu.empty() || !r->forcePc[0].empty() || r->paramsChanged;
	const bool needTran = !opOnly && (r->sched[1].lu.empty() || runChanged || !r->forcePc[1].empty() ||
			r->paramsChanged);
	if ((needOP || needTran) && pilotPass(r, needOP, needTran)) return -1;
	r->paramsChanged = false;
	if (opOnly && r->sched[1].lu.empty()) {
		r->sched[1] = r->sched[0]; r->active[1] = r->active[0]; r->consts[1] = r->consts[0];
		if (uploadSchedule(r, 1)) return -1;
	}

	int slots = M;
	if (r->interpreter) {
		/* whole batch on the generic kernels: per-instance state in HBM */
		if (genericShape(r, M, r->tpbOpt, &r->tpb, &r->wsShared, r->dWs)) return -1;
		if (r->dA.reserve(p.nnzA * m8) || r->dB.reserve(p.N * m8) || r->dX.reserve(p.N * m8) ||
				r->dS.reserve(std::max(1, p.stateDoubles) * m8) ||
				r->dIS.reserve(std::max(1, p.stateInts) * m4) || r->dCD.reserve(CD_NUM * m8))
			return -1;
	} else {
		r->wsShared = 0;
		/* Launch shape (measured on B200, profiles/): the kernel is bound by the
		 * latency of its own spilled state and of the T-line history, never by
		 * HBM bandwidth, and its loop body is several times the instruction cache.
		 *  - a batch that does not fill the machine: 64-thread blocks spread over
		 *    all SMs, every register the compiler wants (one instance's latency);
		 *  - a full machine: the warps of a block start every attempt together
		 *    (block_sync) and share instruction-cache lines - one block per SM;
		 *  - a working set far beyond the register file (cfg5: 763 doubles per
		 *    instance) gains more from 28 resident warps at 72 registers than from
		 *    255 registers at 8 warps; mid-size circuits (cfg4: 434) the opposite;
		 *  - small circuits (N <= 12) in big batches: 128 registers, 16 warps. */
		const int workingSet = p.stateDoubles + p.nnzA + 3 * p.N + (int)p.pmap.size();
		const bool full = (long long)M >= (long long)sms * 256;
		int sync = r->blockSync;
		r->minBlocksUsed = r->minBlocks;
		if (r->tpbOpt > 0) {
			r->tpb = std::min(1024, std::max(32, (r->tpbOpt / 32) * 32));
			if (sync < 0) sync = 0;
		} else if (full && p.N > 12) {
			r->tpb = (workingSet >= 600) ? 896 : 256;
			if (sync < 0) sync = 1;
		} else {
			r->tpb = 64;
			if (sync < 0) sync = 0;
		}
		if (r->minBlocks == 0 && p.N <= 12 && (long long)M > (long long)sms * 4 * r->tpb)
			r->minBlocksUsed = 65536 / (128 * r->tpb);
		r->blockSyncUsed = sync;
		if (buildJitKernel(r)) return -1;
		/* active lanes per warp */
		const int wpb = r->tpb / 32;
		int perSM = r->jit->maxBlocksPerSM;
		if (r->maxBlocksPerSM > 0) perSM = std::min(perSM, r->maxBlocksPerSM);
		const long long capacity = (long long)sms * perSM * wpb;
		int lanes = 32;
		if (r->lanesOpt > 0) lanes = std::min(32, std::max(1, r->lanesOpt));
		/* (measured on B200: fewer than 32 lanes never paid off - the extra
		 * warps thrash L1 with their local-memory windows - so 32 unless asked) */
		int pow2 = 1;
		while (pow2 * 2 <= lanes) pow2 *= 2;
		r->lanes = r->blockSyncUsed ? 32 : pow2;      /* the block barrier needs every thread */
		const long long warps = ((long long)M + r->lanes - 1) / r->lanes;
		long long want = (warps + wpb - 1) / wpb;
		/* fewer blocks than SMs: spread the warps over all of them (lanes claim
		 * their instances one by one, surplus lanes stay idle) */
		if (want < sms) want = std::min<long long>(sms, warps);
		r->jitBlocks = (int)std::max(1LL, std::min(want, capacity / wpb));
		slots = r->jitBlocks * wpb * r->lanes;
		r->args[0].lanes = r->args[1].lanes = r->lanes;
	}

	/* T-line history ring: one column per instance (interpreter) or per
	 * resident thread (specialised kernel) */
	int R = 0;
	if (nh > 0 && !opOnly) {
		R = r->histRecordsOpt;
		if (R <= 0) {
			size_t freeB = 0, totalB = 0;
			CU(cudaMemGetInfo(&freeB, &totalB));
			const double per = (double)(nh + 1) * 8.0 * slots;
			R = (int)std::min(32768.0, std::max(64.0, 0.4 * (double)freeB / per));
		}
	}
	r->histRecords = R;
	if (r->dHt.reserve((size_t)R * slots * 8) || r->dHx.reserve((size_t)R * std::max(1, nh) * slots * 8)) return -1;
	for (int ph = 0; ph < 2; ph++) {
		SimcuArgs &a = r->args[ph];
		a.histRecords = R;
		a.Ht = r->dHt.as<double>(); a.Hx = r->dHx.as<double>();
		a.A = r->dA.as<double>(); a.B = r->dB.as<double>(); a.X = r->dX.as<double>();
		a.Xrec = nullptr; a.S = r->dS.as<double>(); a.CD = r->dCD.as<double>();
		a.IS = r->dIS.as<int>();
		a.wsShared = r->wsShared; a.wsGlobal = r->dWs.as<double>();
	}
	CU(cudaDeviceSynchronize());
	r->prepared = true;
	return 0;
}

int runDevice(simcu_ *r, cudaStream_t st, bool opOnly)
{
	if (!r || !r->prepared) return fail("simcuRunDevice before simcuPrepare");
	int launches = 0;
	CU(cudaEventRecord(r->ev0, st));
	if (r->interpreter) {
		if (resetState(r, st)) return -1;
		/* an instance that fails keeps NaN from the failure time on (all-ones = NaN) */
		if (r->ngrid > 0 && !r->probeRows.empty())
			CU(cudaMemsetAsync(r->dOutGrid.p, 0xff, (size_t)r->ngrid * r->probeRows.size() * r->M * 8, st));
		if (simcuLaunchLoad(&r->args[0], 128, st)) return fail("load kernel launch failed");
		if (simcuLaunchOp(&r->args[0], r->tpb, opOnly ? 1 : 0, st)) return fail("op kernel launch failed");
		launches += 2;
		CU(cudaEventRecord(r->ev2, st));
		while (!opOnly) {
			CU(cudaMemsetAsync(r->dRemaining.p, 0, sizeof(int), st));
			if (simcuLaunchTran(&r->args[1], r->tpb, st)) return fail("tran kernel launch failed");
			launches++;
			CU(cudaMemcpyAsync(r->hRemaining, r->dRemaining.p, sizeof(int), cudaMemcpyDeviceToHost, st));
			CU(cudaStreamSynchronize(st));
			if (*r->hRemaining <= 0) break;
		}
	} else {
		CU(cudaMemsetAsync(r->dRemaining.p, 0, 2 * sizeof(int), st));
		CU(cudaEventRecord(r->ev2, st));
		int op = opOnly ? 1 : 0;
		SimcuArgs a1 = r->args[1];
		a1.count = r->M;
		int blocks = r->jitBlocks;
		if (r->single >= 0) {
			/* simcuRunTransient / simcuRunOperatingPoint: one instance of the batch */
			std::vector<int> one(1, r->single);
			if (r->dOrder2.upload(one, st)) return -1;
			a1.order = r->dOrder2.as<int>(); a1.count = 1; blocks = 1;
		}
		void *params[] = {(void *)&a1, (void *)&op};
		const size_t smem = (size_t)r->stagePoints * 32;          /* staged tables */
		if (smem > 0) {
			CU(cudaFuncSetAttribute((const void *)r->jit->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			CU(cudaFuncSetAttribute((const void *)r->jit->kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
					(int)std::min<size_t>(100, (smem * 100 + 228 * 1024 - 1) / (228 * 1024))));
		}
		cudaError_t e = cudaLaunchKernel((const void *)r->jit->kernel, dim3(blocks), dim3(r->tpb),
				params, smem, st);
		if (e != cudaSuccess) return fail("specialised kernel launch failed: %s", cudaGetErrorString(e));
		launches = 1;
		r->stats.historyRetries = 0;
		if (!opOnly && r->histRecords > 0 && r->historyRetry && r->single < 0) {
			/* Instances that outran the T-line history ring (a burst of tiny steps
			 * longer than the ring inside one line delay) run again on their own:
			 * the same ring memory shared by few threads is orders of magnitude
			 * deeper.  The reference keeps the whole history and never fails here. */
			std::vector<int> status(r->M), acc(r->M), rejl(r->M), rejn(r->M);
			const int *ci = r->dCI.as<int>();
			const size_t mb = (size_t)r->M * 4;
			CU(cudaMemcpyAsync(status.data(), ci + (size_t)CI_STATUS * r->M, mb, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(acc.data(), ci + (size_t)CI_ACCEPTED * r->M, mb, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(rejl.data(), ci + (size_t)CI_REJ_LTE * r->M, mb, cudaMemcpyDeviceToHost, st));
			CU(cudaMemcpyAsync(rejn.data(), ci + (size_t)CI_REJ_NC * r->M, mb, cudaMemcpyDeviceToHost, st));
			CU(cudaStreamSynchronize(st));
			std::vector<int> again;
			long long okAttempts = 0, okCount = 0;
			for (int i = 0; i < r->M; i++) {
				if (status[i] == SIMCU_ERR_HISTORY) again.push_back(i);
				else if (status[i] == 0) { okAttempts += (long long)acc[i] + rejl[i] + rejn[i]; okCount++; }
			}
			const int n2 = (int)again.size();
			const long long slots1 = (long long)r->jitBlocks * (r->tpb / 32) * r->lanes;
			if (n2 > 0 && n2 * 4LL <= slots1) {
				const int blocks2 = (n2 + r->tpb - 1) / r->tpb;
				const long long slots2 = (long long)blocks2 * (r->tpb / 32) * r->lanes;
				const long long deeper = (long long)r->histRecords * slots1 / slots2;
				SimcuArgs a2 = r->args[1];
				a2.histRecords = (int)std::min<long long>(deeper, 1LL << 30);
				a2.count = n2;
				/* instances whose step size has collapsed would run for minutes:
				 * give the second pass 32 x the attempts of an average instance */
				if (okCount > 0)
					a2.maxAttempts = std::min(a2.maxAttempts, 32 * (okAttempts / okCount) + 1000);
				if (r->dOrder2.upload(again, st)) return -1;
				a2.order = r->dOrder2.as<int>();
				CU(cudaMemsetAsync(r->dRemaining.p, 0, 2 * sizeof(int), st));
				void *params2[] = {(void *)&a2, (void *)&op};
				e = cudaLaunchKernel((const void *)r->jit->kernel, dim3(blocks2), dim3(r->tpb), params2, smem, st);
				if (e != cudaSuccess) return fail("retry launch failed: %s", cudaGetErrorString(e));
				launches++;
				r->stats.historyRetries = n2;
			}
		}
	}
	CU(cudaEventRecord(r->ev1, st));
	CU(cudaStreamSynchronize(st));
	float ms = 0.f;
	CU(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
	r->stats.kernelLaunches = launches;
	r->stats.deviceMs = ms;
	CU(cudaEventElapsedTime(&ms, r->ev2, r->ev1));
	r->stats.tranMs = ms;
	r->ran = true;
	if (!opOnly) r->ranTransient = true;
	return 0;
}

int collectStats(simcu_ *r)
{
	const int M = r->M;
	std::vector<int> ci((size_t)CI_NUM * M);
	CU(cudaMemcpy(ci.data(), r->dCI.p, ci.size() * sizeof(int), cudaMemcpyDeviceToHost));
	simcuStats_ &s = r->stats;
	s.accepted = s.rejectedLTE = s.rejectedNC = s.factorizations = s.opFactorizations = s.failed = 0;
	s.pivotBreakdowns = 0; s.replayDiffs = 0; s.growthWarnings = 0;
	for (int i = 0; i < M; i++) {
		if (r->single >= 0 && i != r->single) continue;
		s.accepted += ci[(size_t)CI_ACCEPTED * M + i];
		s.rejectedLTE += ci[(size_t)CI_REJ_LTE * M + i];
		s.rejectedNC += ci[(size_t)CI_REJ_NC * M + i];
		s.factorizations += ci[(size_t)CI_NFACT * M + i];
		s.opFactorizations += ci[(size_t)CI_NFACT_OP * M + i];
		s.failed += (ci[(size_t)CI_STATUS * M + i] != 0);
		if (!r->interpreter) {
			s.pivotBreakdowns += ci[(size_t)CI_BREAK * M + i];
			s.replayDiffs += ci[(size_t)CI_REPLAY_DIFF * M + i];
			s.growthWarnings += ci[(size_t)CI_GROWTH * M + i];
		}
	}
	const Program &p = r->prog;
	s.numUnknowns = p.N; s.numNonzeros = p.nnzA;
	s.numFactor = r->sched[1].nL + r->sched[1].nU + p.N;
	s.numFactorOP = r->sched[0].nL + r->sched[0].nU + p.N;
	s.stateDoubles = p.stateDoubles;
	s.histRows = (int)p.histRows.size(); s.histRecords = r->histRecords;
	s.workspaceInShared = r->wsShared; s.threadsPerBlock = r->tpb;
	s.registersPerThread = (r->jit && !r->interpreter) ? r->jit->numRegs : 0;
	s.localBytesPerThread = (r->jit && !r->interpreter) ? (int)r->jit->localBytes : 0;
	s.gridBlocks = r->interpreter ? (r->M + r->tpb - 1) / r->tpb : r->jitBlocks;
	s.lanesPerWarp = r->interpreter ? 32 : r->lanes;
	return 0;
}

/* raw (time, X) record of one instance as the reference lays it out:
 * core/matrix.c:430-470, [numPoints][numVariables], column 0 = time */
int fetchRaw(simcu_ *r, int instance, double **data, int *numPoints)
{
	if (r->rawMaxPoints <= 0 || instance < r->rawFirst || instance >= r->rawFirst + r->rawCount)
		return fail("instance %d was not recorded", instance);
	const int nv = r->prog.N + 1;
	int npts = 0;
	CU(cudaMemcpy(&npts, r->dCI.as<int>() + (size_t)CI_NPTS * r->M + instance, sizeof(int),
			cudaMemcpyDeviceToHost));
	*numPoints = npts;
	if (npts > r->rawMaxPoints) return 1;       /* caller enlarges and reruns */
	double *out = (double *)std::malloc(sizeof(double) * (size_t)std::max(1, npts) * nv);
	if (!out) return fail("out of memory");
	if (npts > 0)
		CU(cudaMemcpy2D(out, sizeof(double),
				r->dOutRaw.as<double>() + (instance - r->rawFirst),
				(size_t)r->rawCount * sizeof(double), sizeof(double), (size_t)npts * nv,
				cudaMemcpyDeviceToHost));
	*data = out;
	return 0;
}

int variablesOf(simcu_ *r, char ***variables, int *numVariables)
{
	const int nv = r->prog.N + 1;
	if (r->variableNames.empty()) {
		r->variableNames.push_back(strdup("time"));
		for (int k = 1; k <= r->prog.N; k++) r->variableNames.push_back(strdup(r->nl.rowName[k].c_str()));
	}
	if (numVariables) *numVariables = nv;
	if (variables) {
		char **v = (char **)std::malloc(sizeof(char *) * (nv + 1));
		if (!v) return fail("out of memory");
		for (int k = 0; k < nv; k++) v[k] = r->variableNames[k];
		v[nv] = nullptr;
		*variables = v;
	}
	return 0;
}

int runSingle(simcu_ *r, bool opOnly, double tstep, double tstop, double tmax, int instance,
		double *data[], char **variables[], int *numPoints, int *numVariables)
{
	if (!r || !data || !numPoints) return -1;
	if (instance < 0 || instance >= r->M) return fail("instance out of range");
	const int saveMax = r->rawMaxPoints, saveFirst = r->rawFirst, saveCount = r->rawCount;
	int cap = opOnly ? 4 : 8192;
	int rc = -1;
	for (int attempt = 0; attempt < 2; attempt++) {
		r->rawMaxPoints = cap; r->rawFirst = instance; r->rawCount = 1;
		r->prepared = false;
		if (prepare(r, tstep, tstop, tmax, nullptr, 0, opOnly)) break;
		r->single = r->interpreter ? -1 : instance;     /* launch that instance alone */
		const int rcRun = runDevice(r, 0, opOnly);
		if (rcRun) { r->single = -1; break; }
		int status = 0;
		if (cudaMemcpy(&status, r->dCI.as<int>() + (size_t)CI_STATUS * r->M + instance,
				sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) break;
		if (status != 0) {
			static const char *const why[] = {"", "Timestep too small", "U is singular",
				"NaN in a device evaluation", "operating point failed to converge",
				"T-line history ring too short", "attempt budget exhausted"};
			fail("%s", why[(status >= 0 && status <= 6) ? status : 0]);
			break;
		}
		rc = fetchRaw(r, instance, data, numPoints);
		if (rc != 1) break;
		cap = *numPoints + 16;
		rc = -1;
	}
	r->rawMaxPoints = saveMax; r->rawFirst = saveFirst; r->rawCount = saveCount;
	r->prepared = false;
	if (rc != 0) { r->single = -1; return -1; }
	collectStats(r);
	r->single = -1;
	r->ran = false;       /* the batch outputs were not produced by this launch */
	return variablesOf(r, variables, numVariables);
}

}  /* namespace */

extern "C" {

int simcuPrepare(simcu_ *r, double tstep, double tstop, double tmax, char *probes[], int numProbes)
{
	return prepare(r, tstep, tstop, tmax, probes, numProbes, false);
}

int simcuUpload(simcu_ *r, void *stream)
{
	if (!r || !r->prepared) return fail("simcuUpload before simcuPrepare");
	if (r->paramsChanged) {
		/* new per-instance values: the pivot sequences are re-derived on pilots of
		 * the new sweep (and the kernel rebuilt if they changed) */
		std::vector<std::string> names = r->probeNames;
		std::vector<char *> probes;
		for (size_t i = 0; i < names.size(); i++) probes.push_back(&names[i][0]);
		return prepare(r, r->ctl.tstep, r->ctl.tstop, r->ctl.tmax, probes.data(), (int)probes.size(), false);
	}
	std::string err;
	if (!refreshParameters(r->nl, r->M, r->prog, err)) return fail("%s", err.c_str());
	return uploadParameters(r, (cudaStream_t)stream);
}

int simcuRunDevice(simcu_ *r, void *stream) { return runDevice(r, (cudaStream_t)stream, false); }

int simcuTransposeOutput(simcu_ *r, void *stream, void **deviceData, void **deviceStatus,
		int *numInstances, int *valuesPerInstance)
{
	if (!r || !r->ran) return fail("simcuTransposeOutput before simcuRunDevice");
	cudaStream_t st = (cudaStream_t)stream;
	const int M = r->M, rows = r->ngrid * (int)r->probeRows.size();
	if (rows > 0) {
		if (r->dOutT.reserve((size_t)rows * M * 8)) return -1;
		if (simcuLaunchTranspose(r->dOutGrid.as<double>(), r->dOutT.as<double>(), M, rows, st))
			return fail("transpose kernel launch failed");
	}
	if (deviceData) *deviceData = rows > 0 ? r->dOutT.p : nullptr;
	if (deviceStatus) {
		/* a snapshot: the next batch may already be running when this one travels */
		if (r->dStatusT.reserve((size_t)M * 4)) return -1;
		CU(cudaMemcpyAsync(r->dStatusT.p, r->dCI.as<int>() + (size_t)CI_STATUS * M, (size_t)M * 4,
				cudaMemcpyDeviceToDevice, st));
		*deviceStatus = r->dStatusT.p;
	}
	if (numInstances) *numInstances = M;
	if (valuesPerInstance) *valuesPerInstance = rows;
	return 0;
}

int simcuFetch(simcu_ *r, double *data, int *status, void *stream)
{
	if (!r || !r->ran) return fail("simcuFetch before simcuRunDevice");
	cudaStream_t st = (cudaStream_t)stream;
	const int M = r->M, rows = r->ngrid * (int)r->probeRows.size();
	if (data && rows > 0) {
		if (simcuTransposeOutput(r, stream, nullptr, nullptr, nullptr, nullptr)) return -1;
		CU(cudaMemcpyAsync(data, r->dOutT.p, (size_t)rows * M * 8, cudaMemcpyDeviceToHost, st));
	}
	if (status)
		CU(cudaMemcpyAsync(status, r->dCI.as<int>() + (size_t)CI_STATUS * M, (size_t)M * 4,
				cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

int simcuNumGrid(simcu_ *r) { return r ? r->ngrid : -1; }
void *simcuDeviceOutput(simcu_ *r) { return r ? r->dOutGrid.p : nullptr; }

int simcuGetStats(simcu_ *r, simcuStats_ *stats)
{
	if (!r || !stats) return -1;
	if (r->ran && collectStats(r)) return -1;
	*stats = r->stats;
	return 0;
}

int simcuSetMeasures(simcu_ *r, char *probes[], const int *kinds, const double *levels, int n)
{
	if (!r || n < 0 || (n > 0 && (!probes || !kinds))) return -1;
	std::vector<int> m;
	std::vector<double> lv;
	for (int i = 0; i < n; i++) {
		const int row = r->nl.findRow(probes[i]);
		if (row < 1) return fail("unknown probe %s", probes[i] ? probes[i] : "(null)");
		if (kinds[i] < MS_MIN || kinds[i] > MS_TCROSS_FALL) return fail("unknown measure kind %d", kinds[i]);
		m.push_back(kinds[i]); m.push_back(row);
		lv.push_back(levels ? levels[i] : 0.0);
	}
	r->measures = m; r->measureLevels = lv;
	r->prepared = false;
	return 0;
}

int simcuFetchMeasures(simcu_ *r, double *out)
{
	if (!r || !r->ran || !out) return -1;
	const int n = (int)r->measures.size() / 2, M = r->M;
	if (n == 0) return 0;
	std::vector<double> tmp((size_t)n * M);
	CU(cudaMemcpy(tmp.data(), r->dMeasure.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
	for (int i = 0; i < M; i++)
		for (int q = 0; q < n; q++) out[(size_t)i * n + q] = tmp[(size_t)q * M + i];
	return 0;
}

int simcuFetchCounters(simcu_ *r, char *which, int *out)
{
	if (!r || !r->ran || !which || !out) return -1;
	static const struct { const char *name; int row; } rows[] = {
		{"status", CI_STATUS}, {"accepted", CI_ACCEPTED}, {"rejected_lte", CI_REJ_LTE},
		{"rejected_nc", CI_REJ_NC}, {"factorizations", CI_NFACT}, {"op_factorizations", CI_NFACT_OP},
		{"pivot_breakdowns", CI_BREAK}, {"replay_diff", CI_REPLAY_DIFF}, {"pivot_growth", CI_GROWTH},
		{"linearize_diff", CI_LIN_DIFF}};
	for (size_t i = 0; i < sizeof rows / sizeof rows[0]; i++)
		if (!std::strcmp(which, rows[i].name)) {
			CU(cudaMemcpy(out, r->dCI.as<int>() + (size_t)rows[i].row * r->M, (size_t)r->M * 4,
					cudaMemcpyDeviceToHost));
			return 0;
		}
	return fail("unknown counter %s", which);
}

int simcuFetchRaw(simcu_ *r, int instance, double *data[], int *numPoints, int *numVariables)
{
	if (!r || !r->ran || !data || !numPoints) return -1;
	const int rc = fetchRaw(r, instance, data, numPoints);
	if (rc == 1) return fail("raw record overflow: %d points, raw_max_points = %d", *numPoints, r->rawMaxPoints);
	if (numVariables) *numVariables = r->prog.N + 1;
	return rc;
}

int simcuFetchRawBlock(simcu_ *r, double *out, int *numPoints)
{
	if (!r || !r->ran || !out || !numPoints) return -1;
	if (r->rawMaxPoints <= 0 || r->rawCount < 1) return fail("no raw record was kept (raw_max_points)");
	CU(cudaMemcpy(numPoints, r->dCI.as<int>() + (size_t)CI_NPTS * r->M + r->rawFirst,
			(size_t)r->rawCount * sizeof(int), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(out, r->dOutRaw.p, (size_t)r->rawMaxPoints * (r->prog.N + 1) * r->rawCount * 8,
			cudaMemcpyDeviceToHost));
	return 0;
}

int simcuRunTransientBatch(simcu_ *r, double tstep, double tstop, double tmax, int restart,
		char *probes[], int numProbes, double *data[], int *numGrid, int *status[],
		simcuStats_ *stats)
{
	if (!r) return -1;
	/* core/simulator.c:111,138-140: restart = 0 after a finished transient would
	 * continue it from the state it ended in.  The batch engine keeps no
	 * per-instance state between launches, so this is refused, not ignored;
	 * restart = 0 on a simulator that has not run is a normal start (:111). */
	if (!restart && r->ranTransient)
		return fail("continuing a finished transient (restart = 0) is not supported: "
				"per-instance state is not kept between launches");
	if (prepare(r, tstep, tstop, tmax, probes, numProbes, false)) return -1;
	if (runDevice(r, 0, false)) return -1;
	const size_t n = (size_t)r->M * r->ngrid * numProbes;
	double *out = data ? (double *)std::malloc(sizeof(double) * std::max<size_t>(1, n)) : nullptr;
	int *st = status ? (int *)std::malloc(sizeof(int) * r->M) : nullptr;
	if ((data && !out) || (status && !st)) { std::free(out); std::free(st); return fail("out of memory"); }
	if (simcuFetch(r, out, st, nullptr)) { std::free(out); std::free(st); return -1; }
	if (data) *data = out;
	if (status) *status = st;
	if (numGrid) *numGrid = r->ngrid;
	if (stats) return simcuGetStats(r, stats);
	return 0;
}

/* Operating point of every instance: the requested
 * unknowns come back through the measure path (kind "final" at the OP). */
int simcuRunOperatingPointBatch(simcu_ *r, char *probes[], int numProbes, double *data[],
		int *status[])
{
	if (!r || numProbes < 1 || !probes || !data) return -1;
	const std::vector<int> saveM = r->measures;
	const std::vector<double> saveL = r->measureLevels;
	std::vector<int> kinds(numProbes, MS_FINAL);
	int rc = simcuSetMeasures(r, probes, kinds.data(), nullptr, numProbes);
	double *out = nullptr;
	int *st = nullptr;
	if (!rc) rc = prepare(r, 0.0, 0.0, 0.0, nullptr, 0, true);
	if (!rc) rc = runDevice(r, 0, true);
	if (!rc) {
		out = (double *)std::malloc(sizeof(double) * (size_t)r->M * numProbes);
		st = status ? (int *)std::malloc(sizeof(int) * r->M) : nullptr;
		if (!out || (status && !st)) rc = fail("out of memory");
	}
	if (!rc) rc = simcuFetchMeasures(r, out);
	if (!rc && st && cudaMemcpy(st, r->dCI.as<int>() + (size_t)CI_STATUS * r->M, (size_t)r->M * 4,
			cudaMemcpyDeviceToHost) != cudaSuccess)
		rc = fail("status download failed");
	r->measures = saveM; r->measureLevels = saveL; r->prepared = false;
	if (rc) { std::free(out); std::free(st); return -1; }
	*data = out;
	if (status) *status = st;
	return 0;
}

int simcuRunTransient(simcu_ *r, double tstep, double tstop, double tmax, int restart,
		int instance, double *data[], char **variables[], int *numPoints, int *numVariables)
{
	if (!r) return -1;
	if (!restart && r->ranTransient)      /* see simcuRunTransientBatch */
		return fail("continuing a finished transient (restart = 0) is not supported: "
				"per-instance state is not kept between launches");
	return runSingle(r, false, tstep, tstop, tmax, instance, data, variables, numPoints, numVariables);
}

int simcuRunOperatingPoint(simcu_ *r, int instance, double *data[], char **variables[],
		int *numPoints, int *numVariables)
{
	return runSingle(r, true, 0.0, 0.0, 0.0, instance, data, variables, numPoints, numVariables);
}

/* ---- introspection (host only) --------------------------------------------- */
const char *simcuNvrtcPath(void) { return jitNvrtcPath(); }
int simcuNumUnknowns(simcu_ *r) { return r ? (int)r->nl.rowName.size() - 1 : -1; }
int simcuNumNonzeros(simcu_ *r) { return r ? (int)r->nl.nodes.size() : -1; }

int simcuGetPattern(simcu_ *r, int *rowIdx, int *colStart)
{
	if (!r || !rowIdx || !colStart) return -1;
	const int N = (int)r->nl.rowName.size() - 1;
	for (int c = 0; c <= N; c++) colStart[c] = 0;
	int k = 0;
	for (std::set<std::pair<int, int> >::const_iterator it = r->nl.nodes.begin();
			it != r->nl.nodes.end(); ++it, ++k) {
		rowIdx[k] = it->second - 1;
		colStart[it->first]++;
	}
	for (int c = 0; c < N; c++) colStart[c + 1] += colStart[c];
	return 0;
}

const char *simcuVariableName(simcu_ *r, int k)
{
	if (!r || k < 0 || k + 1 >= (int)r->nl.rowName.size()) return nullptr;
	return r->nl.rowName[k + 1].c_str();
}

int simcuGetPivots(simcu_ *r, int phase, int *pc, int *pr)
{
	if (!r || phase < 0 || phase > 1 || r->sched[phase].pc.empty()) return -1;
	const Schedule &s = r->sched[phase];
	for (size_t k = 0; k < s.pc.size(); k++) { pc[k] = s.pc[k]; pr[k] = s.pr[k]; }
	return 0;
}

/* test hook: everything Prepare does short of touching a GPU - flatten, a
 * structural symbolic analysis on random pilot values (or the forced pivots),
 * source generation and the NVRTC compile for sm_100a */
int simcuCompileOnly(simcu_ *r, char *probes[], int numProbes, char **source, int *cubinBytes)
{
	if (!r) return -1;
	if (ensureCompiled(r)) return -1;
	const Program &p = r->prog;
	std::vector<int> probeRows;
	for (int i = 0; i < numProbes; i++) {
		const int row = r->nl.findRow(probes[i]);
		if (row < 1) return fail("unknown probe %s", probes[i] ? probes[i] : "(null)");
		probeRows.push_back(row);
	}
	Schedule sched[2];
	std::vector<char> active[2] = {p.activeOP, p.activeTran};
	unsigned long long seed = 0x9E3779B97F4A7C15ULL;
	for (int ph = 0; ph < 2; ph++) {
		std::vector<std::vector<double> > pilots(2, std::vector<double>(p.nnzA, 0.0));
		for (size_t q = 0; q < pilots.size(); q++)
			for (int sl = 0; sl < p.nnzA; sl++) {
				seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
				if (active[ph][sl]) pilots[q][sl] = 0.5 + (double)(seed >> 40) / (double)(1 << 24);
			}
		const bool forced = (int)r->forcePc[ph].size() == p.N;
		std::string err;
		if (!analyse(p.N, p.colStart, p.rowIdx, active[ph], pilots, r->pivotThreshold,
				forced ? r->forcePc[ph].data() : nullptr, forced ? r->forcePr[ph].data() : nullptr,
				sched[ph], err))
			return fail("%s", err.c_str());
	}
	const int tpb = r->tpbOpt > 0 ? std::max(32, (r->tpbOpt / 32) * 32) : 64;
	const std::vector<double> consts[2] = {p.constOP, p.constTran};
	JitConfig cfg;
	cfg.threadsPerBlock = tpb; cfg.minBlocks = r->minBlocks; cfg.gmin = r->ctl.gmin; cfg.replay = r->replay;
	cfg.blockSync = (r->blockSync > 0); cfg.syncGroup = r->syncGroup; cfg.growthLimit = r->growthLimit;
	cfg.windows = !r->winTstep.empty(); cfg.fmad = r->fmad; cfg.smemTab = (r->smemTablesKB > 0); cfg.tlinePrefetch = r->tlinePrefetch;
	cfg.solveBlocks = r->solveBlocks; cfg.sharedCb = r->sharedCb;
	const std::string src = assembleSource(p, sched, active, consts, probeRows, r->measures, cfg);
	if (source) *source = strdup(src.c_str());
	if (!cubinBytes) return 0;          /* the generated source only */
	JitKernel k;
	std::string err;
	if (!jitCompile(src, k, err, cfg.fmad != 0)) return fail("%s", err.c_str());
	*cubinBytes = (int)k.cubin.size();
	return 0;
}

/* test hook (no GPU needed): the host-side arrays a launch of the specialised
 * kernel reads - what Prepare / Upload copy to the device - by name: "pmap" (int),
 * "pconst", "pinst", "tabP" (double), "iinst", "tabDesc" (int), "control" (one
 * SimcuControl with the run constants of core/simulator.c:91-109 resolved for this
 * tstep / tstop / tmax).  *data receives a malloc'd copy (simcuFree), *bytes its
 * size.  tests/ runs the generated translation unit compiled for the host on
 * these arrays and compares it with the oracle; the product never reads them
 * back. */
int simcuDebugProgram(simcu_ *r, const char *what, double tstep, double tstop, double tmax,
		void **data, int *bytes)
{
	if (!r || !what || !data || !bytes) return -1;
	if (ensureCompiled(r)) return -1;
	const Program &p = r->prog;
	const std::string w(what);
	const void *src = nullptr;
	size_t n = 0;
	std::vector<int> desc;
	SimcuControl c = r->ctl;
	if (w == "pmap") { src = p.pmap.data(); n = p.pmap.size() * sizeof(int); }
	else if (w == "pconst") { src = p.pconst.data(); n = p.pconst.size() * sizeof(double); }
	else if (w == "pinst") { src = p.pinst.data(); n = p.pinst.size() * sizeof(double); }
	else if (w == "iinst") { src = p.iinst.data(); n = p.iinst.size() * sizeof(int); }
	else if (w == "tabP") { src = p.tabP.data(); n = p.tabP.size() * sizeof(double); }
	else if (w == "tabDesc") {
		const int ntab = (int)p.tabType.size();
		desc.assign((size_t)4 * std::max(1, ntab), -1);
		for (int t = 0; t < ntab; t++) {
			desc[4 * t] = p.tabOff[t]; desc[4 * t + 1] = p.tabOff[t + 1] - p.tabOff[t];
			desc[4 * t + 2] = p.tabType[t]; desc[4 * t + 3] = -1;
		}
		src = desc.data(); n = desc.size() * sizeof(int);
	} else if (w == "control") {
		if (!(tstep > 0.0) || !(tstop > 0.0)) return fail("tstep and tstop must be positive");
		if (tmax == 0.0) { tmax = tstop / 50; tmax = (tstep < tmax) ? tstep : tmax; }
		c.tstep = tstep; c.tstop = tstop; c.tmax = tmax;
		c.minstep = (r->minstepOpt < 0) ? tstep * 5e-5 : r->minstepOpt;
		src = &c; n = sizeof c;
	} else return fail("simcuDebugProgram: unknown array %s", what);
	*data = std::malloc(n ? n : 1);
	if (!*data) return -1;
	if (n) std::memcpy(*data, src, n);
	*bytes = (int)n;
	return 0;
}

/* test hook: the packed form of one table ([length][4]: x, y, slope or second
 * derivative, x of the next point) exactly as the kernels read it */
int simcuPackTable(const double *xy, int length, char type, double *packed)
{
	if (!xy || !packed || length < 1) return fail("simcuPackTable: bad arguments");
	std::vector<double> out;
	std::string err;
	if (!packTable(xy, length, type, out, err)) return fail("%s", err.c_str());
	std::memcpy(packed, out.data(), out.size() * sizeof(double));
	return 0;
}

/* test hook: the symbolic analysis alone on caller-supplied matrices */
int simcuSymbolic(int n, const int *colStart, const int *rowIdx, int numPilots,
		const double *values, double threshold, int *pc, int *pr, int *numSlots,
		int *numFactor)
{
	if (n < 1 || !colStart || !rowIdx || !values || numPilots < 1) return -1;
	const int nnz = colStart[n];
	std::vector<int> cs(colStart, colStart + n + 1), ri(rowIdx, rowIdx + nnz);
	std::vector<std::vector<double> > pilots(numPilots);
	for (int q = 0; q < numPilots; q++) pilots[q].assign(values + (size_t)q * nnz, values + (size_t)(q + 1) * nnz);
	Schedule s;
	std::string err;
	if (!analyse(n, cs, ri, std::vector<char>(), pilots, threshold, nullptr, nullptr, s, err))
		return fail("%s", err.c_str());
	for (int k = 0; k < n; k++) { if (pc) pc[k] = s.pc[k]; if (pr) pr[k] = s.pr[k]; }
	if (numSlots) *numSlots = s.F;
	if (numFactor) *numFactor = s.nL + s.nU + n;
	return 0;
}

namespace {
struct EvalCtx { int n; const char **names; };
int evalGetVar(void *ctx, const char *name)
{
	EvalCtx *e = (EvalCtx *)ctx;
	for (int i = 0; i < e->n; i++) if (!std::strcmp(name, e->names[i])) return i;
	return -1;
}
}

int simcuCalcEval(const char *expr, int nvars, const char **names, const double *values,
		double minDiv, int where, double *value, double *derivs)
{
	if (!expr || !value || (nvars > 0 && (!names || !values || !derivs))) return -1;
	EvalCtx ctx; ctx.n = nvars; ctx.names = names;
	CalcProg prog;
	std::string err;
	const int saveQuiet = gQuiet;
	if (!calcCompile(expr, evalGetVar, &ctx, prog, err)) return -1;
	const int nv = (int)prog.varNames.size();
	std::vector<double> vals(std::max(1, nv)), out(1 + std::max(1, nv), 0.0);
	for (int i = 0; i < nv; i++) vals[i] = values[prog.varRef[i]];
	if (where == 0) {
		double f, d;
		calcRun(prog.code.data(), (int)prog.code.size(), prog.consts.data(), vals.data(), -1, minDiv, &f, &d);
		out[0] = f;
		for (int i = 0; i < nv; i++) {
			calcRun(prog.code.data(), (int)prog.code.size(), prog.consts.data(), vals.data(), i, minDiv, &f, &d);
			out[1 + i] = d;
		}
	} else {
		DBuf dCode, dConst, dVals, dOut;
		int rc = 0;
		if (dCode.upload(prog.code) || dConst.upload(prog.consts) || dVals.upload(vals) ||
				dOut.reserve(out.size() * 8))
			rc = -1;
		if (!rc && simcuLaunchCalc(dCode.as<int>(), (int)prog.code.size(), dConst.as<double>(),
				dVals.as<double>(), nv, minDiv, dOut.as<double>(), nullptr))
			rc = -1;
		if (!rc && cudaMemcpy(out.data(), dOut.p, out.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
			rc = -1;
		dCode.release(); dConst.release(); dVals.release(); dOut.release();
		if (rc) { gQuiet = saveQuiet; return fail("device evaluation failed"); }
	}
	*value = out[0];
	for (int i = 0; i < nvars; i++) derivs[i] = NAN;
	for (int i = 0; i < nv; i++) derivs[prog.varRef[i]] = out[1 + i];
	/* calc.c:47,96: a NaN result is an error */
	if (std::isnan(out[0])) return -1;
	for (int i = 0; i < nv; i++) if (std::isnan(out[1 + i])) return -1;
	return 0;
}

}  /* extern "C" */
