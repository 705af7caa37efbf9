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

#include <math.h>

#include "simcu_program.h"

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
