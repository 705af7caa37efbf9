/*
 * simcu_codegen.cpp - writes the topology-specialised part of the production
 * kernel: the netlist as literal device records fed to the hand-written device
 * library (simcu_device.cuh), the B-source expressions as straight-line code
 * (value and partial derivatives, same operations as the bytecode interpreter
 * of simcu_calc.h / the reference's parser.lem actions), and the numeric LU +
 * triangular solves of both phases fully unrolled on named scalars, so the
 * factors live in registers and never touch memory.
 *
 * The text produced here is concatenated with simcu_program.h, simcu_calc.h,
 * simcu_device.cuh and simcu_jit_kernel.cuh and compiled for sm_100a by NVRTC
 * (simcu_jit.cpp).  Nothing in it depends on the number of instances, on
 * parameter values or on the analysis window.
 */
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/simulator_cuda.h"
#include "simcu_codegen.h"

namespace simcu {
namespace {

struct Out {
	std::string s;
	void f(const char *fmt, ...)
	{
		char buf[1024];
		va_list ap;
		va_start(ap, fmt);
		const int n = std::vsnprintf(buf, sizeof buf, fmt, ap);
		va_end(ap);
		if (n >= (int)sizeof buf) {
			std::vector<char> big(n + 1);
			va_start(ap, fmt);
			std::vsnprintf(big.data(), big.size(), fmt, ap);
			va_end(ap);
			s += big.data();
		} else s += buf;
	}
};

/* exact literal of a double (also inf / nan) */
std::string lit(double v)
{
	unsigned long long bits;
	std::memcpy(&bits, &v, sizeof bits);
	char buf[64];
	std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxLL)", bits);
	return buf;
}

/* ---- B-source expressions -------------------------------------------------- */
struct Dual { std::string f, d; };   /* d empty = structurally zero derivative */

/* One evaluation of calc `id` as SSA statements: dvar < 0 value only, else the
 * partial derivative against calc variable dvar.  Mirrors calcRun() of
 * simcu_calc.h operation by operation; a derivative that is structurally zero
 * (constants, other variables, comparison results) is dropped instead of being
 * multiplied through, which differs from the interpreter only when an
 * intermediate value is non-finite. */
void emitCalc(Out &o, const Program &p, int id, int dvar, const char *ind)
{
	const int c0 = p.calcOff[id], c1 = p.calcOff[id + 1];
	std::vector<Dual> st;
	int t = 0;
	auto tmp = [&](const std::string &expr) {
		char name[32];
		std::snprintf(name, sizeof name, "t%d", t++);
		o.f("%sconst double %s = %s;\n", ind, name, expr.c_str());
		return std::string(name);
	};
	for (int pc = c0; pc < c1; pc += 2) {
		const int op = p.calcCode[pc], arg = p.calcCode[pc + 1];
		switch (op) {
		case OP_CONST: { Dual a; a.f = tmp(lit(p.calcConst[arg])); st.push_back(a); break; }
		case OP_VAR: {
			Dual a;
			char nm[32];
			std::snprintf(nm, sizeof nm, "v%d", arg);
			a.f = nm;
			if (arg == dvar) a.d = tmp("1.0");
			st.push_back(a);
			break;
		}
		case OP_NEG: {
			Dual &b = st.back();
			b.f = tmp("-1 * " + b.f);
			if (!b.d.empty()) b.d = tmp("-1 * " + b.d);
			break;
		}
		case OP_ADD: case OP_SUB: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			const char *sgn = (op == OP_ADD) ? " + " : " - ";
			b.f = tmp(b.f + sgn + c.f);
			if (!b.d.empty() && !c.d.empty()) b.d = tmp(b.d + sgn + c.d);
			else if (!c.d.empty()) b.d = (op == OP_ADD) ? c.d : tmp("0.0 - " + c.d);
			break;
		}
		case OP_MUL: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			std::string d;
			if (!b.d.empty() && !c.d.empty()) d = tmp(b.d + " * " + c.f + " + " + b.f + " * " + c.d);
			else if (!b.d.empty()) d = tmp(b.d + " * " + c.f);
			else if (!c.d.empty()) d = tmp(b.f + " * " + c.d);
			b.f = tmp(b.f + " * " + c.f);
			b.d = d;
			break;
		}
		case OP_DIV: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			std::string d;
			if (!b.d.empty() || !c.d.empty()) {
				std::string num;
				if (!b.d.empty() && !c.d.empty()) num = b.d + " * " + c.f + " - " + b.f + " * " + c.d;
				else if (!b.d.empty()) num = b.d + " * " + c.f;
				else num = "0.0 - " + b.f + " * " + c.d;
				d = tmp("calcDiv((" + num + "), (" + c.f + " * " + c.f + "), m)");
			}
			b.f = tmp("calcDiv(" + b.f + ", " + c.f + ", m)");
			b.d = d;
			break;
		}
		case OP_POW: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			const std::string pw = tmp("pow(" + b.f + ", " + c.f + ")");
			std::string d;
			if (!b.d.empty() || !c.d.empty()) {
				const std::string left = b.d.empty() ? std::string()
					: b.d + " * calcDiv(" + c.f + ", " + b.f + ", m)";
				const std::string right = c.d.empty() ? std::string()
					: "((" + b.f + " != 0.0) ? " + c.d + " * calcLn(" + b.f + ") : 0.0)";
				if (!left.empty() && !right.empty()) d = tmp(pw + " * (" + left + " + " + right + ")");
				else d = tmp(pw + " * (" + (left.empty() ? right : left) + ")");
			}
			b.f = pw;
			b.d = d;
			break;
		}
		case OP_FUNC: {
			Dual &b = st.back();
			char fn[32], dn[32];
			std::snprintf(fn, sizeof fn, "t%d", t++);
			std::snprintf(dn, sizeof dn, "t%d", t++);
			o.f("%sdouble %s, %s; calcFunc(%d, %s, %s, m, &%s, &%s);\n", ind, fn, dn, arg,
					b.f.c_str(), b.d.empty() ? "0.0" : b.d.c_str(), fn, dn);
			/* u(x): the derivative 1/m at exactly 0 does not depend on x' */
			const bool keep = !b.d.empty() || (arg == FN_U && dvar >= 0);
			b.f = fn;
			b.d = keep ? std::string(dn) : std::string();
			break;
		}
		case OP_GT: case OP_LT: case OP_GE: case OP_LE: case OP_EQ: case OP_NE: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			const char *rel = (op == OP_GT) ? ">" : (op == OP_LT) ? "<" : (op == OP_GE) ? ">="
				: (op == OP_LE) ? "<=" : (op == OP_EQ) ? "==" : "!=";
			b.f = tmp("(" + b.f + " " + rel + " " + c.f + ") ? 1.0 : 0.0");
			b.d.clear();
			break;
		}
		case OP_NOT: {
			Dual &b = st.back();
			b.f = tmp("(" + b.f + " == 0.0) ? 1.0 : 0.0");
			b.d.clear();
			break;
		}
		case OP_AND: case OP_OR: {
			const Dual c = st.back(); st.pop_back();
			Dual &b = st.back();
			b.f = tmp("((" + b.f + " == 1.0) " + ((op == OP_AND) ? "&&" : "||") + " (" + c.f +
					" == 1.0)) ? 1.0 : 0.0");
			b.d.clear();
			break;
		}
		default: { /* OP_IF */
			Dual &b = st.back();
			b.f = tmp("(" + b.f + " != 0.0) ? 1.0 : 0.0");
			b.d.clear();
			break;
		}
		}
	}
	o.f("%sf = %s; dv = %s;\n", ind, st.back().f.c_str(),
			(dvar >= 0 && !st.back().d.empty()) ? st.back().d.c_str() : "0.0");
}

/* Value and ALL partial derivatives of calc `id` in one pass: the value chain
 * is emitted once, each derivative chain only where it is structurally
 * non-zero.  Per variable this is operation for operation what emitCalc()
 * writes for that variable alone. */
struct DualN { std::string f; std::vector<std::string> d; };

void emitCalcGrad(Out &o, const Program &p, int id, const char *ind)
{
	const int c0 = p.calcOff[id], c1 = p.calcOff[id + 1];
	const int v0 = p.calcVarOff[id], nv = p.calcVarOff[id + 1] - v0;
	std::vector<char> want(nv, 0);
	for (int v = 0; v < nv; v++) want[v] = (p.calcVars[v0 + v] != CALC_VAR_TIME);
	std::vector<DualN> st;
	int t = 0;
	auto tmp = [&](const std::string &expr) {
		char name[32];
		std::snprintf(name, sizeof name, "t%d", t++);
		o.f("%sconst double %s = %s;\n", ind, name, expr.c_str());
		return std::string(name);
	};
	auto fresh = [&]() { DualN a; a.d.assign(nv, std::string()); return a; };
	for (int pc = c0; pc < c1; pc += 2) {
		const int op = p.calcCode[pc], arg = p.calcCode[pc + 1];
		switch (op) {
		case OP_CONST: { DualN a = fresh(); a.f = tmp(lit(p.calcConst[arg])); st.push_back(a); break; }
		case OP_VAR: {
			DualN a = fresh();
			char nm[32];
			std::snprintf(nm, sizeof nm, "v%d", arg);
			a.f = nm;
			if (want[arg]) a.d[arg] = tmp("1.0");
			st.push_back(a);
			break;
		}
		case OP_NEG: {
			DualN &b = st.back();
			b.f = tmp("-1 * " + b.f);
			for (int v = 0; v < nv; v++) if (!b.d[v].empty()) b.d[v] = tmp("-1 * " + b.d[v]);
			break;
		}
		case OP_ADD: case OP_SUB: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			const char *sgn = (op == OP_ADD) ? " + " : " - ";
			b.f = tmp(b.f + sgn + c.f);
			for (int v = 0; v < nv; v++) {
				if (!b.d[v].empty() && !c.d[v].empty()) b.d[v] = tmp(b.d[v] + sgn + c.d[v]);
				else if (!c.d[v].empty()) b.d[v] = (op == OP_ADD) ? c.d[v] : tmp("0.0 - " + c.d[v]);
			}
			break;
		}
		case OP_MUL: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			for (int v = 0; v < nv; v++) {
				std::string d;
				if (!b.d[v].empty() && !c.d[v].empty())
					d = tmp(b.d[v] + " * " + c.f + " + " + b.f + " * " + c.d[v]);
				else if (!b.d[v].empty()) d = tmp(b.d[v] + " * " + c.f);
				else if (!c.d[v].empty()) d = tmp(b.f + " * " + c.d[v]);
				b.d[v] = d;
			}
			b.f = tmp(b.f + " * " + c.f);
			break;
		}
		case OP_DIV: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			std::string den;
			for (int v = 0; v < nv; v++) {
				std::string d;
				if (!b.d[v].empty() || !c.d[v].empty()) {
					std::string num;
					if (!b.d[v].empty() && !c.d[v].empty())
						num = b.d[v] + " * " + c.f + " - " + b.f + " * " + c.d[v];
					else if (!b.d[v].empty()) num = b.d[v] + " * " + c.f;
					else num = "0.0 - " + b.f + " * " + c.d[v];
					if (den.empty()) den = tmp(c.f + " * " + c.f);
					d = tmp("calcDiv((" + num + "), " + den + ", m)");
				}
				b.d[v] = d;
			}
			b.f = tmp("calcDiv(" + b.f + ", " + c.f + ", m)");
			break;
		}
		case OP_POW: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			const std::string pw = tmp("pow(" + b.f + ", " + c.f + ")");
			std::string ratio, lnb;
			for (int v = 0; v < nv; v++) {
				std::string d;
				if (!b.d[v].empty() || !c.d[v].empty()) {
					std::string left, right;
					if (!b.d[v].empty()) {
						if (ratio.empty()) ratio = tmp("calcDiv(" + c.f + ", " + b.f + ", m)");
						left = b.d[v] + " * " + ratio;
					}
					if (!c.d[v].empty()) {
						if (lnb.empty()) lnb = tmp("calcLn(" + b.f + ")");
						right = "((" + b.f + " != 0.0) ? " + c.d[v] + " * " + lnb + " : 0.0)";
					}
					if (!left.empty() && !right.empty()) d = tmp(pw + " * (" + left + " + " + right + ")");
					else d = tmp(pw + " * (" + (left.empty() ? right : left) + ")");
				}
				b.d[v] = d;
			}
			b.f = pw;
			break;
		}
		case OP_FUNC: {
			DualN &b = st.back();
			/* the value and the derivative factor once, then one product per
			 * variable: calcFunc(id, x, 1) returns f'(x) * 1 */
			char fn[32], gn[32];
			std::snprintf(fn, sizeof fn, "t%d", t++);
			std::snprintf(gn, sizeof gn, "t%d", t++);
			bool any = false;
			for (int v = 0; v < nv; v++) any = any || !b.d[v].empty();
			if (arg == FN_U) {
				o.f("%sdouble %s, %s; calcFunc(%d, %s, 0.0, m, &%s, &%s);\n", ind, fn, gn, arg,
						b.f.c_str(), fn, gn);
				for (int v = 0; v < nv; v++) b.d[v] = want[v] ? std::string(gn) : std::string();
			} else if (!any) {
				o.f("%sdouble %s, %s; calcFunc(%d, %s, 0.0, m, &%s, &%s);\n", ind, fn, gn, arg,
						b.f.c_str(), fn, gn);
			} else {
				/* keep the interpreter's rounding: it evaluates factor * x' with
				 * the factor computed first, which is what factor * 1.0 * x' is
				 * not in general - so emit one calcFunc per live variable and
				 * let the compiler merge the shared transcendental calls */
				bool first = true;
				for (int v = 0; v < nv; v++) {
					if (b.d[v].empty()) continue;
					char fv[32], dv[32];
					std::snprintf(fv, sizeof fv, "t%d", t++);
					std::snprintf(dv, sizeof dv, "t%d", t++);
					o.f("%sdouble %s, %s; calcFunc(%d, %s, %s, m, &%s, &%s);\n", ind, fv, dv, arg,
							b.f.c_str(), b.d[v].c_str(), fv, dv);
					if (first) { o.f("%sconst double %s = %s;\n", ind, fn, fv); first = false; }
					b.d[v] = dv;
				}
			}
			b.f = fn;
			break;
		}
		case OP_GT: case OP_LT: case OP_GE: case OP_LE: case OP_EQ: case OP_NE: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			const char *rel = (op == OP_GT) ? ">" : (op == OP_LT) ? "<" : (op == OP_GE) ? ">="
				: (op == OP_LE) ? "<=" : (op == OP_EQ) ? "==" : "!=";
			b.f = tmp("(" + b.f + " " + rel + " " + c.f + ") ? 1.0 : 0.0");
			b.d.assign(nv, std::string());
			break;
		}
		case OP_NOT: {
			DualN &b = st.back();
			b.f = tmp("(" + b.f + " == 0.0) ? 1.0 : 0.0");
			b.d.assign(nv, std::string());
			break;
		}
		case OP_AND: case OP_OR: {
			const DualN c = st.back(); st.pop_back();
			DualN &b = st.back();
			b.f = tmp("((" + b.f + " == 1.0) " + ((op == OP_AND) ? "&&" : "||") + " (" + c.f +
					" == 1.0)) ? 1.0 : 0.0");
			b.d.assign(nv, std::string());
			break;
		}
		default: { /* OP_IF */
			DualN &b = st.back();
			b.f = tmp("(" + b.f + " != 0.0) ? 1.0 : 0.0");
			b.d.assign(nv, std::string());
			break;
		}
		}
	}
	for (int v = 0; v < nv; v++)
		if (want[v]) o.f("%sg[%d] = %s;\n", ind, v, st.back().d[v].empty() ? "0.0" : st.back().d[v].c_str());
}

void emitCalcAll(Out &o, const Program &p)
{
	o.f("DEV void genCalc(Inst &c, int id, int dvar, double &f, double &dv)\n{\n");
	/* gmin (the guard of every division, parser.lem:30-31) is a compile-time
	 * constant here, so divisions of constants by constants fold away */
	o.f("\tconst double m = SIMCU_GMIN;\n\t(void)m; (void)c;\n\tf = 0.0; dv = 0.0;\n\tswitch (id) {\n");
	const int ncalc = (int)p.calcOff.size() - 1;
	for (int id = 0; id < ncalc; id++) {
		const int v0 = p.calcVarOff[id], v1 = p.calcVarOff[id + 1];
		o.f("\tcase %d: {\n", id);
		for (int k = v0; k < v1; k++) {
			if (p.calcVars[k] == CALC_VAR_TIME) o.f("\t\tconst double v%d = c.time;\n", k - v0);
			else o.f("\t\tconst double v%d = c.X(%d);\n", k - v0, p.calcVars[k]);
		}
		o.f("\t\tswitch (dvar) {\n");
		for (int dv = -1; dv < v1 - v0; dv++) {
			if (dv >= 0 && p.calcVars[v0 + dv] == CALC_VAR_TIME) continue;
			o.f("\t\tcase %d: {\n", dv);
			emitCalc(o, p, id, dv, "\t\t\t");
			o.f("\t\t\tbreak;\n\t\t}\n");
		}
		o.f("\t\tdefault: break;\n\t\t}\n\t\tbreak;\n\t}\n");
	}
	o.f("\tdefault: break;\n\t}\n}\n\n");
	o.f("DEV void genCalcGrad(Inst &c, int id, double *g)\n{\n");
	o.f("\tconst double m = SIMCU_GMIN;\n\t(void)m; (void)c; (void)g;\n\tswitch (id) {\n");
	for (int id = 0; id < ncalc; id++) {
		const int v0 = p.calcVarOff[id], v1 = p.calcVarOff[id + 1];
		o.f("\tcase %d: {\n", id);
		for (int k = v0; k < v1; k++) {
			if (p.calcVars[k] == CALC_VAR_TIME) o.f("\t\tconst double v%d = c.time;\n", k - v0);
			else o.f("\t\tconst double v%d = c.X(%d);\n", k - v0, p.calcVars[k]);
		}
		emitCalcGrad(o, p, id, "\t\t");
		o.f("\t\tbreak;\n\t}\n");
	}
	o.f("\tdefault: break;\n\t}\n}\n\n");
}

/* ---- device phases ---------------------------------------------------------- */
void emitRecord(Out &o, const Program &p, int k)
{
	const int *rec = &p.dev[(size_t)k * SIMCU_DEV_STRIDE];
	o.f("\t{ const int d[] = {");
	for (int j = 0; j < SIMCU_DEV_STRIDE; j++) o.f("%s%d", j ? "," : "", rec[j]);
	o.f("};");
	/* the Jacobian-variable table is a separate literal array: its loop bound
	 * comes from d[], which must fold first */
	const int kind = rec[DF_TYPE];
	o.f(" const int bv[] = {");
	if ((kind == DK_BI || kind == DK_BV || kind == DK_BC) && rec[DF_NBV] > 0) {
		const int nbv = rec[DF_NBV], off = rec[DF_BVOFF];
		for (int j = 0; j < 4 * nbv; j++) o.f("%s%d", j ? "," : "", p.bvar[(size_t)4 * off + j]);
	} else o.f("0");
	o.f("}; (void)bv;");
}

bool inPhase(const int *rec, int phase)
{
	const int kind = rec[DF_TYPE];
	const bool wave = (kind == DK_V || kind == DK_I) && rec[DF_WTYPE] != 0;
	const bool ta = (kind == DK_VI) && rec[DF_TATAB] >= 0;
	const bool bsrc = (kind == DK_BI || kind == DK_BV || kind == DK_BC);
	switch (phase) {
	case 0: return true;                                                /* load */
	case 1: return kind == DK_C || kind == DK_L || kind == DK_BC || kind == DK_T;   /* initStep */
	case 2: return wave || ta || kind == DK_T || (bsrc && rec[DF_USETIME]);          /* step */
	case 3: return kind == DK_C || kind == DK_L || kind == DK_BC;                   /* integrate */
	case 4: return kind == DK_VI || bsrc;                                            /* linearize */
	case 5: return kind == DK_C || kind == DK_L || kind == DK_BC;                   /* minStep */
	default: return wave || ta;                                                      /* nextStep */
	}
}

/* devices whose step phase runs the break-point test (checkbreak.c:43-67) */
bool stepHasBreakTest(const int *rec)
{
	const int kind = rec[DF_TYPE];
	if ((kind == DK_V || kind == DK_I) && rec[DF_WTYPE] != 0) return true;
	if (kind == DK_VI && rec[DF_TATAB] >= 0) return true;
	if (kind == DK_T) return true;
	return (kind == DK_BI || kind == DK_BV) && rec[DF_USETIME];
}

void emitPhases(Out &o, const Program &p, bool tlinePrefetch, bool sharedCb)
{
	static const char *const head[7] = {
		"DEV void genLoad(Inst &c)\n{\n",
		"DEV void genInitStep(Inst &c)\n{\n",
		"DEV int genStep(Inst &c)\n{\n\tint brk = 0;\n",
		"DEV void genIntegrate(Inst &c)\n{\n",
		"DEV int genLinearize(Inst &c)\n{\n\tint linear = 1;\n",
		"DEV void genMinStep(Inst &c, double &lteStep)\n{\n",
		"DEV void genNextStep(Inst &c, double &thisStep)\n{\n"};
	static const char *const call[7] = {
		" devLoad(c, d, bv); }\n", " devInitStep(c, d); }\n", " brk |= devStep(c, d, bv); }\n",
		" devIntegrate(c, d); }\n", " linear &= devLinearize(c, d, bv); }\n",
		" devMinStep(c, d, lteStep); }\n", " devNextStep(c, d, thisStep); }\n"};
	static const char *const onErr[7] = {
		"\tif (c.err) return;\n", "", "\tif (c.err) return brk;\n", "\tif (c.err) return;\n",
		"\tif (c.err) return 0;\n", "\tif (c.err) return;\n", ""};
	static const char *const tail[7] = {"}\n\n", "}\n\n", "\treturn brk;\n}\n\n", "}\n\n",
		"\treturn linear;\n}\n\n", "}\n\n", "}\n\n"};
	for (int ph = 0; ph < 7; ph++) {
		o.f("%s", head[ph]);
		if (ph == 3) {          /* integrate: the shared time ring advances once */
			bool any = false;
			for (int k = 0; k < p.ndev; k++) any = any || inPhase(&p.dev[(size_t)k * SIMCU_DEV_STRIDE], 3);
			if (any) o.f("\titBegin(c);\n");
		}
		if (ph == 2 && sharedCb) {    /* step: the shared break-point clock advances once */
			bool any = false;
			for (int k = 0; k < p.ndev; k++) any = any || stepHasBreakTest(&p.dev[(size_t)k * SIMCU_DEV_STRIDE]);
			if (any) o.f("\tcbBegin(c);\n");
		}
		if (ph == 2 && tlinePrefetch) {
			/* step: with several T-lines, a first pass locates every line's delayed
			 * time and prefetches the records, so their DRAM latencies overlap */
			int nT = 0;
			for (int k = 0; k < p.ndev; k++) nT += (p.dev[(size_t)k * SIMCU_DEV_STRIDE + DF_TYPE] == DK_T);
			if (nT >= 2 && !p.histRows.empty())
				for (int k = 0; k < p.ndev; k++) {
					if (p.dev[(size_t)k * SIMCU_DEV_STRIDE + DF_TYPE] != DK_T) continue;
					emitRecord(o, p, k);
					o.f(" tlinePrefetch(c, d); }\n");
				}
		}

		for (int k = 0; k < p.ndev; k++) {
			if (!inPhase(&p.dev[(size_t)k * SIMCU_DEV_STRIDE], ph)) continue;
			emitRecord(o, p, k);
			o.f("%s%s", call[ph], onErr[ph]);
		}
		o.f("%s", tail[ph]);
	}
}

/* ---- numeric LU + solves, straight-line --------------------------------------- */

/* A slots and right-hand-side rows that the linearize phase can change, i.e.
 * everything that may differ between two solves of ONE simulatorSolve loop
 * (core/simulator.c:45-69): the stamps of the V-I curves (vicurve.c:110-163)
 * and of the B sources (nonlinear_*.c).  Conservative: every slot and row the
 * record of such a device names. */
void newtonTouched(const Program &p, std::vector<char> &slot, std::vector<char> &row)
{
	slot.assign(p.nnzA, 0);
	row.assign(p.N + 1, 0);
	auto markSlot = [&](int sl) { if (sl >= 0 && sl < p.nnzA) slot[sl] = 1; };
	auto markRow = [&](int r) { if (r >= 1 && r <= p.N) row[r] = 1; };
	for (int k = 0; k < p.ndev; k++) {
		const int *rec = &p.dev[(size_t)k * SIMCU_DEV_STRIDE];
		const int kind = rec[DF_TYPE];
		const bool bsrc = (kind == DK_BI || kind == DK_BV || kind == DK_BC);
		if (kind != DK_VI && !bsrc) continue;
		markRow(rec[DF_K]); markRow(rec[DF_J]); markRow(rec[DF_L]); markRow(rec[DF_M]);
		markRow(rec[DF_ROWR]); markRow(rec[DF_ROWM]);
		if (kind == DK_VI) for (int j = 0; j < 8; j++) markSlot(rec[DF_VA0 + j]);
		else {
			for (int j = 0; j < 6; j++) markSlot(rec[DF_BA0 + j]);
			for (int j = 0; j < rec[DF_NBV]; j++) {
				const int *bv = &p.bvar[(size_t)4 * (rec[DF_BVOFF] + j)];
				markRow(bv[0]); markSlot(bv[1]); markSlot(bv[2]);
			}
		}
	}
}

/* The matrix of a circuit with T-lines falls apart into independent blocks (a
 * line couples its two ports only through the history, i.e. the right-hand side:
 * tline.c:104-111), and the pivot sequence wanders between them.  `blocks`
 * regroups the straight-line code block by block - loads, elimination, back
 * substitution and the stores of one block before the next one starts - so that
 * only one block's entries are live at a time (at 72 registers per thread the
 * all-at-once order spills ~90 doubles per solve).  Every operation keeps its
 * operands and its place among the operations of ITS block, so the results are
 * bit-identical.  blocks == 2 additionally runs the blocks no linearize stamp
 * reaches only in the FIRST solve of a Newton loop: their stamps, right-hand
 * side and therefore solution cannot change before the loop ends (cfg5: the
 * seven inner nodes of the 8-section line, 21 of 48 unknowns). */
void emitFactorSolve(Out &o, const Program &p, const Schedule &s, const std::vector<char> &active,
		const std::vector<double> &cst, int phase, bool guard, int blocks)
{
	const int N = p.N, F = s.F;
	o.f("/* %s: %d unknowns, %d matrix slots (%d fill), %d L + %d U off-diagonal entries */\n",
			phase ? "transient" : "operating point", N, F, F - p.nnzA, s.nL, s.nU);
	/* where each pivot step starts; connected components over slots (0..F-1) and
	 * right-hand-side entries (F..F+N-1) */
	std::vector<int> start(N), uf(F + N);
	for (int v = 0; v < F + N; v++) uf[v] = v;
	auto find = [&](int v) { while (uf[v] != v) { uf[v] = uf[uf[v]]; v = uf[v]; } return v; };
	auto join = [&](int a, int b) { a = find(a); b = find(b); if (a != b) uf[(a < b) ? b : a] = (a < b) ? a : b; };
	std::vector<char> used(F, 0);
	{
		int pos = 0;
		for (int k = 0; k < N; k++) {
			start[k] = pos;
			const int d = s.lu[pos], yk = s.lu[pos + 1], nU = s.lu[pos + 2], nL = s.lu[pos + 3];
			used[d] = 1;
			join(d, yk);
			for (int j = 0; j < nU; j++) {
				used[s.lu[pos + 4 + 2 * j]] = 1;
				join(d, s.lu[pos + 4 + 2 * j]); join(d, s.lu[pos + 4 + 2 * j + 1]);
			}
			const int *lrow = &s.lu[pos + 4 + 2 * nU];
			for (int r = 0; r < nL; r++) {
				const int *e = lrow + r * (2 + nU);
				used[e[0]] = 1;
				join(d, e[0]); join(d, e[1]);
				for (int j = 0; j < nU; j++) { used[e[2 + j]] = 1; join(d, e[2 + j]); }
			}
			pos += 4 + 2 * nU + nL * (2 + nU);
		}
	}
	if (blocks <= 0) for (int v = 0; v < F + N; v++) uf[v] = 0;     /* one block */
	/* blocks in the order of their first pivot step; which ones a linearize pass reaches */
	std::vector<int> blockOf(F + N, -1), roots;
	for (int k = 0; k < N; k++) {
		const int r = find(s.lu[start[k]]);
		if (blockOf[r] < 0) { blockOf[r] = (int)roots.size(); roots.push_back(r); }
	}
	for (int v = 0; v < F + N; v++) blockOf[v] = blockOf[find(v)];
	const int nb = (int)roots.size();
	std::vector<char> touched(nb, 1);
	if (blocks >= 2) {
		std::vector<char> tslot, trow;
		newtonTouched(p, tslot, trow);
		o.f("/* linearize can change A slots:");
		for (int sl = 0; sl < p.nnzA; sl++) if (tslot[sl]) o.f(" %d", sl);
		o.f("; B rows:");
		for (int r = 1; r <= N; r++) if (trow[r]) o.f(" %d", r);
		o.f(" */\n");
		touched.assign(nb, 0);
		for (int sl = 0; sl < p.nnzA; sl++) if (tslot[sl] && used[sl] && blockOf[sl] >= 0) touched[blockOf[sl]] = 1;
		for (int r = 1; r <= N; r++) if (trow[r] && blockOf[F + r - 1] >= 0) touched[blockOf[F + r - 1]] = 1;
	}
	int nInvariant = 0;
	for (int b = 0; b < nb; b++) nInvariant += !touched[b];
	o.f("/* %d independent block%s, %d of them not reached by any linearize stamp */\n", nb, nb == 1 ? "" : "s",
			(blocks >= 2) ? nInvariant : 0);
	o.f("DEV void genFactorSolve%d(Inst &c, const bool first)\n{\n\tbool bad = false, grow = false;\n\t(void)first;\n",
			phase);
	auto emitBlock = [&](int b, const char *growVar) {
		o.f("\t{\n");
		for (int sl = 0; sl < F; sl++) {
			if (!used[sl] || blockOf[sl] != b) continue;
			/* +-1 incidence entries are literals: reciprocals of such pivots, and the
			 * multipliers built from them, fold at compile time (1/1 and x*1 are
			 * exact, so the results do not change), and the stamp needs no register */
			if (sl < p.nnzA && active[sl] && cst[sl] == cst[sl])
				o.f("\tdouble w%d = %s;\n", sl, lit(cst[sl]).c_str());
			else if (sl < p.nnzA && active[sl]) o.f("\tdouble w%d = c.Av[%d];\n", sl, sl);
			else o.f("\tdouble w%d = 0.0;\n", sl);
		}
		for (int r = 0; r < N; r++) if (blockOf[F + r] == b) o.f("\tdouble y%d = c.Bv[%d];\n", r, r);
		for (int k = 0; k < N; k++) {
			const int pos = start[k];
			const int d = s.lu[pos], yk = s.lu[pos + 1] - F, nU = s.lu[pos + 2], nL = s.lu[pos + 3];
			if (blockOf[d] != b) continue;
			const int *u = &s.lu[pos + 4];
			const int *lrow = u + 2 * nU;
			o.f("\tconst double i%d = 1.0 / w%d; bad |= (w%d == 0.0) | !isfinite(i%d);\n", k, d, d, k);
			for (int r = 0; r < nL; r++) {
				const int *e = lrow + r * (2 + nU);
				/* growth guard: the fixed sequence was chosen on pilot instances; a
				 * multiplier this large means the pivot has lost that many digits */
				o.f("\t{ const double l = w%d * i%d;", e[0], k);
				if (guard) o.f(" %s |= (fabs(l) > SIMCU_LMAX);", growVar);
				for (int j = 0; j < nU; j++) o.f(" w%d -= l * w%d;", e[2 + j], u[2 * j]);
				o.f(" y%d -= l * y%d; }\n", e[1] - F, yk);
			}
		}
		/* back substitution, reverse pivot order; the solution of column pc[k]
		 * replaces the right-hand side of its pivot row */
		for (int k = N - 1; k >= 0; k--) {
			const int q = start[k];
			if (blockOf[s.lu[q]] != b) continue;
			const int yk = s.lu[q + 1] - F, nU = s.lu[q + 2];
			const int *u = &s.lu[q + 4];
			o.f("\t{ double acc = y%d;", yk);
			for (int j = 0; j < nU; j++) o.f(" acc -= w%d * y%d;", u[2 * j], u[2 * j + 1] - F);
			o.f(" y%d = acc * i%d; }\n", yk, k);
		}
		for (int col = 0; col < N; col++)
			if (blockOf[s.xslot[col]] == b) o.f("\tc.Xv[%d] = y%d;\n", col, s.xslot[col] - F);
		o.f("\t}\n");
	};
	if (blocks >= 2 && nInvariant > 0) {
		/* growth seen in these blocks counts for every solve of the loop, as if they
		 * had been factorised again */
		o.f("\tif (first) {\n\tbool growInv = false;\n");
		for (int b = 0; b < nb; b++) if (!touched[b]) emitBlock(b, "growInv");
		o.f("\tc.growInv = growInv;\n\t}\n\tgrow |= c.growInv;\n");
	}
	for (int b = 0; b < nb; b++) if (touched[b] || blocks < 2) emitBlock(b, "grow");
	o.f("\tif (bad) c.err = SIMCU_ERR_SINGULAR;\n\tif (grow) c.growth++;\n}\n\n");
}

/* The last accepted solution (matrixRecall, core/matrix.c:385-392) is only ever
 * read - between a rejected attempt and the next solve, or as the "previous
 * point" of the output interpolation - at the rows the step / integrate phases
 * and the probes use: C / L / nonlinear-C terminals, the variables of
 * time-dependent B sources, probe and measure rows.  Only those are saved on
 * accept and restored on reject (cfg5: 10 of 48 unknowns). */
void emitSaveRecall(Out &o, const Program &p, const std::vector<int> &probeRows,
		const std::vector<int> &measures)
{
	std::vector<char> need(p.N + 1, 0);
	for (int k = 0; k < p.ndev; k++) {
		const int *rec = &p.dev[(size_t)k * SIMCU_DEV_STRIDE];
		const int kind = rec[DF_TYPE];
		const bool bsrc = (kind == DK_BI || kind == DK_BV || kind == DK_BC);
		if (kind == DK_C || kind == DK_BC) { need[rec[DF_K]] = 1; need[rec[DF_J]] = 1; }
		if (kind == DK_L) need[rec[DF_ROWR]] = 1;
		if (bsrc && rec[DF_USETIME]) {
			need[rec[DF_K]] = 1; need[rec[DF_J]] = 1; need[rec[DF_ROWR]] = 1;
			if (rec[DF_ROWM] > 0) need[rec[DF_ROWM]] = 1;
			for (int j = 0; j < rec[DF_NBV]; j++) need[p.bvar[(size_t)4 * (rec[DF_BVOFF] + j)]] = 1;
			const int id = rec[DF_CALC];
			for (int v = p.calcVarOff[id]; v < p.calcVarOff[id + 1]; v++)
				if (p.calcVars[v] > 0) need[p.calcVars[v]] = 1;
		}
	}
	for (size_t j = 0; j < probeRows.size(); j++) need[probeRows[j]] = 1;
	for (size_t q = 0; 2 * q + 1 < measures.size(); q++) need[measures[2 * q + 1]] = 1;
	o.f("DEV void genSaveX(const Inst &c)\n{\n");
	for (int r = 1; r <= p.N; r++) if (need[r]) o.f("\tc.Xa[%d] = c.Xv[%d];\n", r - 1, r - 1);
	o.f("\t(void)c;\n}\n\n");
	o.f("DEV void genRecallX(const Inst &c)\n{\n");
	for (int r = 1; r <= p.N; r++) if (need[r]) o.f("\tc.Xv[%d] = c.Xa[%d];\n", r - 1, r - 1);
	o.f("\t(void)c;\n}\n\n");
}

void emitRows(Out &o, const char *name, const std::vector<int> &rows)
{
	o.f("DEV void %s(int *rows)\n{\n", name);
	for (size_t j = 0; j < rows.size(); j++) o.f("\trows[%zu] = %d;\n", j, rows[j]);
	o.f("\t(void)rows;\n}\n\n");
}

}  /* namespace */

std::string generatePrelude(const Program &p, int numProbes, int numMeasures, const JitConfig &cfg)
{
	const int threadsPerBlock = cfg.threadsPerBlock, minBlocks = cfg.minBlocks;
	const double gmin = cfg.gmin;
	Out o;
	o.f("#define SIMCU_JIT 1\n");
	o.f("#define SIMCU_N %d\n#define SIMCU_NNZ %d\n#define SIMCU_SD %d\n#define SIMCU_SI %d\n",
			p.N, p.nnzA, p.stateDoubles, p.stateInts);
	o.f("#define SIMCU_NPARAM %d\n#define SIMCU_NISET %d\n#define SIMCU_NH %d\n",
			(int)p.pmap.size(), p.niinst, (int)p.histRows.size());
	o.f("#define SIMCU_NPROBES %d\n#define SIMCU_TPB %d\n#define SIMCU_MINB %d\n", numProbes,
			threadsPerBlock, minBlocks > 0 ? minBlocks : 1);
	o.f("#define SIMCU_NMEASURES %d\n", numMeasures);
	o.f("#define SIMCU_REPLAY %d\n", cfg.replay ? 1 : 0);
	o.f("#define SIMCU_BLOCKSYNC %d\n#define SIMCU_SYNCGROUP %d\n", cfg.blockSync ? 1 : 0, cfg.syncGroup);
	o.f("#define SIMCU_WINDOWS %d\n#define SIMCU_SMEMTAB %d\n", cfg.windows ? 1 : 0, cfg.smemTab ? 1 : 0);
	o.f("#define SIMCU_FMAD %d   /* compiled with --fmad=%s */\n", cfg.fmad ? 1 : 0, cfg.fmad ? "true" : "false");
	o.f("#define SIMCU_SHAREDCB %d\n", cfg.sharedCb ? 1 : 0);
	o.f("#define HUGE_VAL (__longlong_as_double(0x7ff0000000000000LL))\n");
	o.f("#define SIMCU_GMIN (%s)\n", lit(gmin).c_str());
	o.f("#define SIMCU_LMAX (%s)\n", lit(cfg.growthLimit).c_str());
	o.f("#define M_PI 3.14159265358979323846\n");
	o.f("enum { SIMCU_OK = %d, SIMCU_ERR_TIMESTEP = %d, SIMCU_ERR_SINGULAR = %d, SIMCU_ERR_NAN = %d,\n"
			"\tSIMCU_ERR_OP = %d, SIMCU_ERR_HISTORY = %d, SIMCU_ERR_BUDGET = %d };\n\n",
			SIMCU_OK, SIMCU_ERR_TIMESTEP, SIMCU_ERR_SINGULAR, SIMCU_ERR_NAN, SIMCU_ERR_OP,
			SIMCU_ERR_HISTORY, SIMCU_ERR_BUDGET);
	return o.s;
}

std::string generateTopology(const Program &p, const Schedule sched[2],
		const std::vector<char> active[2], const std::vector<double> consts[2],
		const std::vector<int> &probeRows, const std::vector<int> &measures, const JitConfig &cfg)
{
	Out o;
	const bool growthGuard = cfg.growthLimit > 0.0;
	o.f("/* ---- generated for this netlist by simcu_codegen.cpp ---- */\n");
	emitCalcAll(o, p);
	emitPhases(o, p, cfg.tlinePrefetch != 0, cfg.sharedCb != 0);
	emitFactorSolve(o, p, sched[0], active[0], consts[0], 0, growthGuard, cfg.solveBlocks);
	emitFactorSolve(o, p, sched[1], active[1], consts[1], 1, growthGuard, cfg.solveBlocks);
	emitSaveRecall(o, p, probeRows, measures);
	emitRows(o, "genHistRows", p.histRows);
	emitRows(o, "genProbeRows", probeRows);
	/* measures: (kind, row) pairs; the level comes from the kernel arguments */
	o.f("DEV void genMeasures(const Inst &c, double *meas, double tPrev, double tNow, bool first)\n{\n");
	for (size_t q = 0; 2 * q + 1 < measures.size(); q++)
		o.f("\tmeasureStep(c, %d, %d, __ldg(&c.a.measureLevel[%zu]), tPrev, tNow, first, meas[%zu]);\n",
				measures[2 * q], measures[2 * q + 1], q, q);
	o.f("\t(void)c; (void)meas; (void)tPrev; (void)tNow; (void)first;\n}\n\n");
	return o.s;
}

}  /* namespace simcu */
