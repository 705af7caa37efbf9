/*
 * eis_oracle.c - CPU restatement of the eispice transient hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see eis_oracle.h).  Plain C99, one instance, no
 * dependencies.  Written from the reference's behaviour, not its code: flat
 * arrays instead of linked lists, a recursive-descent expression evaluator
 * instead of the generated LALR parser, dense partial-pivoting LU instead of
 * SuperLU.  Every block cites the reference file:line whose behaviour it
 * restates (paths relative to the reference root).
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eis_oracle.h"

static int g_quiet = 0;
void eoQuiet(int q) { g_quiet = q; }

#define FAIL(...) do { if (!g_quiet) { fprintf(stderr, "eis_oracle: " __VA_ARGS__); \
	fputc('\n', stderr); } return -1; } while (0)
#define TRY(x) do { if ((x)) return -1; } while (0)

#define MaxAbs(x, y) ((fabs(x) > fabs(y)) ? fabs(x) : fabs(y))

/* ------------------------------------------------------------------------ *
 * Options: libs/simulator/core/control.c:53-74, include/control.h:34-64
 * ------------------------------------------------------------------------ */
typedef struct {
	int itl1, itl4;
	double reltol, vntol, captol, abstol, chgtol, trtol, minstep, gmin;
	int maxorder;
	double maxAngleA, maxAngleV;
	double tstop, tstep;
	int integratorOrder;
	double time;
} Control;

static void controlInit(Control *c)
{
	c->itl1 = 100; c->itl4 = 10;
	c->reltol = 0.001; c->vntol = 1e-6; c->abstol = 1e-12;
	c->captol = 1e-18; c->chgtol = 1e-14; c->trtol = 1;
	c->minstep = -1.0; c->gmin = 1e-15; c->maxorder = 2;
	c->maxAngleA = M_PI / 3; c->maxAngleV = M_PI / 3;
	c->tstop = 0.0; c->tstep = 0.0; c->integratorOrder = 1; c->time = 0.0;
}

/* ------------------------------------------------------------------------ *
 * Expression engine: libs/calculon (tokenizer.re:117-242, parser.lem:52-149)
 * ------------------------------------------------------------------------ */
enum {
	T_END, T_FUNC, T_IF, T_PLUS, T_MINUS, T_TIMES, T_DIVIDE, T_POWER,
	T_LPAREN, T_RPAREN, T_GT, T_LT, T_EQ, T_NOT, T_AND, T_OR, T_CONST, T_VAR
};
enum {
	F_ABS, F_ACOSH, F_ACOS, F_ASINH, F_ASIN, F_ATANH, F_ATAN, F_COSH, F_COS,
	F_EXP, F_LN, F_LOG, F_SINH, F_SIN, F_SQRT, F_TAN, F_URAMP, F_U
};
/* tokenizer.re:121-139; order matters only for readability, re2c takes the
 * longest match and "name(" forms never tie with each other */
static const char *funcNames[] = {
	"abs(", "acosh(", "acos(", "asinh(", "asin(", "atanh(", "atan(", "cosh(",
	"cos(", "exp(", "ln(", "log(", "sinh(", "sin(", "sqrt(", "tan(", "uramp(",
	"u(", NULL
};

typedef struct { int type; int id; double constant; } Token;

typedef int (*calcGetVar)(void *ctx, const char *name);

typedef struct {
	Token *tok; int ntok;
	int nvars;            /* distinct variables, first-appearance order */
	char **varNames;
	int *varRef;          /* what the owner's callback returned */
} Calc;

static int isAlpha(int c)
{
	return isalpha(c) || c == '_' || c == ':' || c == '#' || c == '@';
}
static int isText(int c) { return isAlpha(c) || isdigit(c); }

static int ciPrefix(const char *s, const char *lit)
{
	for (; *lit; s++, lit++)
		if (tolower((unsigned char)*s) != *lit) return 0;
	return 1;
}

static void calcPush(Calc *c, int type, int id, double k)
{
	c->tok = realloc(c->tok, sizeof(Token) * (c->ntok + 1));
	c->tok[c->ntok].type = type;
	c->tok[c->ntok].id = id;
	c->tok[c->ntok].constant = k;
	c->ntok++;
}

/* calc.c:104-121 (localGetVarFunc): one entry per distinct name */
static int calcVar(Calc *c, const char *name, calcGetVar gv, void *ctx)
{
	int i, ref;
	for (i = 0; i < c->nvars; i++)
		if (!strcmp(c->varNames[i], name)) return i;
	ref = gv(ctx, name);
	if (ref < 0) return -1;
	c->varNames = realloc(c->varNames, sizeof(char *) * (c->nvars + 1));
	c->varRef = realloc(c->varRef, sizeof(int) * (c->nvars + 1));
	c->varNames[c->nvars] = strdup(name);
	c->varRef[c->nvars] = ref;
	return c->nvars++;
}

static void calcFree(Calc *c)
{
	int i;
	if (c == NULL) return;
	for (i = 0; i < c->nvars; i++) free(c->varNames[i]);
	free(c->varNames); free(c->varRef); free(c->tok); free(c);
}

/* tokenizer.re:99-257 */
static Calc *calcNew(const char *buf, calcGetVar gv, void *ctx)
{
	Calc *c = calloc(1, sizeof(Calc));
	const char *p = buf;
	char name[256];

	for (;;) {
		const char *s = p;
		int i, n;
		if (*p == ' ' || *p == '\t') { p++; continue; }
		if (*p == '\n' || *p == '\r') {       /* CONTINUE = EOL+ "+" */
			const char *q = p;
			while (*q == '\n' || *q == '\r') q++;
			if (*q == '+') { p = q + 1; continue; }
			goto bad;
		}
		if (*p == 0) { calcPush(c, T_END, 0, 0); return c; }
		/* functions and if( : case-insensitive, include the '(' */
		for (i = 0; funcNames[i]; i++)
			if (ciPrefix(p, funcNames[i])) break;
		if (funcNames[i]) {
			/* longest match: "v(" style variables never collide with these,
			 * but "u(" must not eat "uramp(" (handled by table order) */
			calcPush(c, T_FUNC, i, 0); p += strlen(funcNames[i]); continue;
		}
		if (ciPrefix(p, "if(")) { calcPush(c, T_IF, 0, 0); p += 3; continue; }
		switch (*p) {
		case '+': calcPush(c, T_PLUS, 0, 0); p++; continue;
		case '-': calcPush(c, T_MINUS, 0, 0); p++; continue;
		case '*': calcPush(c, T_TIMES, 0, 0); p++; continue;
		case '/': calcPush(c, T_DIVIDE, 0, 0); p++; continue;
		case '^': calcPush(c, T_POWER, 0, 0); p++; continue;
		case '(': calcPush(c, T_LPAREN, 0, 0); p++; continue;
		case ')': calcPush(c, T_RPAREN, 0, 0); p++; continue;
		case '>': calcPush(c, T_GT, 0, 0); p++; continue;
		case '<': calcPush(c, T_LT, 0, 0); p++; continue;
		case '=': calcPush(c, T_EQ, 0, 0); p++; continue;
		case '!': calcPush(c, T_NOT, 0, 0); p++; continue;
		case '&': calcPush(c, T_AND, 0, 0); p++; continue;
		case '|': calcPush(c, T_OR, 0, 0); p++; continue;
		default: break;
		}
		/* FLOAT [unit letters]: tokenizer.re:153-163 */
		if (isdigit((unsigned char)*p) ||
				(*p == '.' && isdigit((unsigned char)p[1]))) {
			const char *q = p;
			double v, mult = 1.0;
			char num[128];
			while (isdigit((unsigned char)*q)) q++;
			if (*q == '.' && isdigit((unsigned char)q[1])) {
				q++;
				while (isdigit((unsigned char)*q)) q++;
			}
			if (*q == 'e' || *q == 'E') {
				const char *e = q + 1;
				if (*e == '-' || *e == '+') e++;
				if (isdigit((unsigned char)*e)) {
					while (isdigit((unsigned char)*e)) e++;
					q = e;
				}
			}
			n = (int)(q - p);
			if (n > 127) goto bad;
			memcpy(num, p, n); num[n] = 0;
			v = atof(num);
			/* unit rules in file order: T G k u n p f Meg mil m (none) */
			if (ciPrefix(q, "t")) mult = 1e12;
			else if (ciPrefix(q, "g")) mult = 1e9;
			else if (ciPrefix(q, "k")) mult = 1e3;
			else if (ciPrefix(q, "u")) mult = 1e-6;
			else if (ciPrefix(q, "n")) mult = 1e-9;
			else if (ciPrefix(q, "p")) mult = 1e-12;
			else if (ciPrefix(q, "f")) mult = 1e-15;
			else if (ciPrefix(q, "meg")) mult = 1e6;
			else if (ciPrefix(q, "mil")) mult = 2.54e-5;
			else if (ciPrefix(q, "m")) mult = 1e-3;
			while (isAlpha((unsigned char)*q)) q++;
			calcPush(c, T_CONST, 0, (mult == 1.0) ? v : v * mult);
			p = q; continue;
		}
		/* v(a,b) -> ( v(a) - v(b) ) : tokenizer.re:165-212 */
		if ((*p == 'v' || *p == 'V') && p[1] == '(') {
			const char *q = p + 2, *a0 = q, *a1, *b0, *b1;
			while (isText((unsigned char)*q)) q++;
			a1 = q;
			if (a1 > a0 && *q == ',') {
				b0 = ++q;
				while (isText((unsigned char)*q)) q++;
				b1 = q;
				if (b1 > b0 && *q == ')') {
					int va, vb;
					if (a1 - a0 > 200 || b1 - b0 > 200) goto bad;
					sprintf(name, "v(%.*s)", (int)(a1 - a0), a0);
					va = calcVar(c, name, gv, ctx);
					if (va < 0) goto bad;
					calcPush(c, T_LPAREN, 0, 0);
					calcPush(c, T_VAR, va, 0);
					calcPush(c, T_MINUS, 0, 0);
					sprintf(name, "v(%.*s)", (int)(b1 - b0), b0);
					vb = calcVar(c, name, gv, ctx);
					if (vb < 0) goto bad;
					calcPush(c, T_VAR, vb, 0);
					calcPush(c, T_RPAREN, 0, 0);
					p = q + 1; continue;
				}
			}
		}
		/* v(x) / i(x) : tokenizer.re:213-228 */
		if ((tolower((unsigned char)*p) == 'v' || tolower((unsigned char)*p) == 'i')
				&& p[1] == '(') {
			const char *q = p + 2, *a0 = q;
			while (isText((unsigned char)*q)) q++;
			if (q > a0 && *q == ')') {
				int v;
				if (q - a0 > 200) goto bad;
				sprintf(name, "%c(%.*s)", tolower((unsigned char)*p),
						(int)(q - a0), a0);
				v = calcVar(c, name, gv, ctx);
				if (v < 0) goto bad;
				calcPush(c, T_VAR, v, 0);
				p = q + 1; continue;
			}
		}
		/* bare VARIABLE : tokenizer.re:229-239 */
		if (isAlpha((unsigned char)*p)) {
			const char *q = p;
			int v;
			while (isText((unsigned char)*q)) q++;
			if (q - p > 200) goto bad;
			sprintf(name, "%.*s", (int)(q - p), p);
			v = calcVar(c, name, gv, ctx);
			if (v < 0) goto bad;
			calcPush(c, T_VAR, v, 0);
			p = q; continue;
		}
		(void)s;
		goto bad;
	}
bad:
	calcFree(c);
	return NULL;
}

/* dual-number evaluation restating parser.lem:60-149.  dvar < 0: value mode. */
typedef struct { double f, d; } Dual;
typedef struct {
	const Calc *c; int pos; const double *vals; int dvar; double m; int err;
} PS;

#define Ln(x) log(fabs(x))
static double Div(double x, double y, double m)
{
	return (fabs(y) > m) ? (x / y) : ((y > 0) ? x / m : x / (-m));
}

static Dual pExpr(PS *s);
static int pEval(PS *s, double *res);

static Dual pFunc(PS *s, int id, Dual b)
{
	Dual a; double m = s->m;
	switch (id) {
	case F_ABS: a.f = fabs(b.f); a.d = Div(b.f, fabs(b.f), m) * b.d; break;
	case F_ACOSH: a.f = acosh(b.f); a.d = sinh(b.f) * b.d; break; /* sic, :110 */
	case F_ACOS: a.f = acos(b.f); a.d = Div(-1, sqrt(1 - b.f * b.f), m) * b.d; break;
	case F_ASINH: a.f = asinh(b.f); a.d = Div(1, sqrt(1 + b.f * b.f), m) * b.d; break;
	case F_ASIN: a.f = asin(b.f); a.d = Div(1, sqrt(1 - b.f * b.f), m) * b.d; break;
	case F_ATANH: a.f = atanh(b.f); a.d = Div(1, (1 - b.f * b.f), m) * b.d; break;
	case F_ATAN: a.f = atan(b.f); a.d = Div(1, (1 + b.f * b.f), m) * b.d; break;
	case F_COSH: a.f = cosh(b.f); a.d = sinh(b.f) * b.d; break;
	case F_COS: a.f = cos(b.f); a.d = -1 * sin(b.f) * b.d; break;
	case F_EXP: a.f = exp(b.f); a.d = exp(b.f) * b.d; break;
	case F_LN: a.f = Ln(b.f); a.d = Div(1, b.f, m) * b.d; break;
	case F_LOG: a.f = log10(b.f); a.d = Div(1, (b.f * Ln(10.0)), m) * b.d; break;
	case F_SINH: a.f = sinh(b.f); a.d = cosh(b.f) * b.d; break;
	case F_SIN: a.f = sin(b.f); a.d = cos(b.f) * b.d; break;
	case F_SQRT: a.f = sqrt(b.f); a.d = Div(1, (2 * sqrt(b.f)), m) * b.d; break;
	case F_TAN: a.f = tan(b.f);
		a.d = Div(1, cos(b.f), m) * Div(1, cos(b.f), m) * b.d; break;
	case F_URAMP: a.f = (b.f > 0) ? b.f : 0;
		a.d = ((b.f == 0) ? 0.5 : ((b.f > 0) ? 1.0 : 0)) * b.d; break;
	default: /* F_U */ a.f = (b.f > 0) ? 1 : 0;
		a.d = (b.f == 0) ? 1 / m : 0.0; break;
	}
	return a;
}

static Dual pPrimary(PS *s)
{
	Dual a = {0, 0};
	const Token *t = &s->c->tok[s->pos];
	switch (t->type) {
	case T_VAR:
		a.f = s->vals[t->id]; a.d = (t->id == s->dvar) ? 1.0 : 0.0;
		s->pos++; return a;
	case T_CONST:
		a.f = t->constant; s->pos++; return a;
	case T_LPAREN:
		s->pos++; a = pExpr(s);
		if (s->err || s->c->tok[s->pos].type != T_RPAREN) { s->err = 1; return a; }
		s->pos++; return a;
	case T_FUNC:
		s->pos++; a = pExpr(s);
		if (s->err || s->c->tok[s->pos].type != T_RPAREN) { s->err = 1; return a; }
		s->pos++; return pFunc(s, t->id, a);
	case T_IF: {
		double r = 0;
		s->pos++;
		if (pEval(s, &r) || s->c->tok[s->pos].type != T_RPAREN) { s->err = 1; return a; }
		s->pos++; a.f = r ? 1 : 0; a.d = 0.0; return a;
	}
	default:
		s->err = 1; return a;
	}
}

/* implicit multiplication "a(b)" binds tightest (rule precedence RPAREN) */
static Dual pPostfix(PS *s)
{
	Dual a = pPrimary(s);
	while (!s->err && s->c->tok[s->pos].type == T_LPAREN) {
		Dual b, r;
		s->pos++; b = pExpr(s);
		if (s->err || s->c->tok[s->pos].type != T_RPAREN) { s->err = 1; return a; }
		s->pos++;
		r.f = a.f * b.f; r.d = a.d * b.f + a.f * b.d; a = r;
	}
	return a;
}

/* unary minus has precedence NOT, above POWER (parser.lem:56,66) */
static Dual pUnary(PS *s)
{
	if (s->c->tok[s->pos].type == T_MINUS) {
		Dual a;
		s->pos++; a = pUnary(s);
		a.f = -1 * a.f; a.d = -1 * a.d; return a;
	}
	/* unary plus exists only in diff mode (parser.lem:102); calcSolve always
	 * runs first and fails on it, so it is a syntax error here */
	return pPostfix(s);
}

static Dual pPower(PS *s)          /* '^' is left-associative (parser.lem:55) */
{
	Dual a = pUnary(s);
	while (!s->err && s->c->tok[s->pos].type == T_POWER) {
		Dual b, r;
		s->pos++; b = pUnary(s);
		r.f = pow(a.f, b.f);
		r.d = pow(a.f, b.f) * (a.d * Div(b.f, a.f, s->m) +
				(a.f ? b.d * Ln(a.f) : 0));
		a = r;
	}
	return a;
}

static Dual pTerm(PS *s)
{
	Dual a = pPower(s);
	while (!s->err) {
		int ty = s->c->tok[s->pos].type;
		Dual b, r;
		if (ty == T_TIMES) {
			s->pos++; b = pPower(s);
			r.f = a.f * b.f; r.d = a.d * b.f + a.f * b.d; a = r;
		} else if (ty == T_DIVIDE) {
			s->pos++; b = pPower(s);
			r.f = Div(a.f, b.f, s->m);
			r.d = Div((a.d * b.f - a.f * b.d), (b.f * b.f), s->m); a = r;
		} else break;
	}
	return a;
}

static Dual pExpr(PS *s)
{
	Dual a = pTerm(s);
	while (!s->err) {
		int ty = s->c->tok[s->pos].type;
		Dual b;
		if (ty == T_PLUS) {
			s->pos++; b = pTerm(s); a.f = a.f + b.f; a.d = a.d + b.d;
		} else if (ty == T_MINUS) {
			s->pos++; b = pTerm(s); a.f = a.f - b.f; a.d = a.d - b.d;
		} else break;
	}
	return a;
}

/* boolean sub-grammar, parser.lem:129-149; returns non-zero on syntax error */
static int pEval(PS *s, double *res)
{
	const Token *tk = s->c->tok;
	int save = s->pos;
	Dual a, b;

	if (tk[s->pos].type == T_NOT) {
		double r;
		s->pos++;
		if (pEval(s, &r)) return -1;
		*res = (r == 0) ? 1 : 0; return 0;
	}
	if (tk[s->pos].type == T_LPAREN) {
		double r1, r2;
		s->pos++;
		if (!pEval(s, &r1) && tk[s->pos].type == T_RPAREN) {
			s->pos++;
			if ((tk[s->pos].type == T_AND && tk[s->pos + 1].type == T_AND) ||
					(tk[s->pos].type == T_OR && tk[s->pos + 1].type == T_OR)) {
				int isAnd = (tk[s->pos].type == T_AND);
				s->pos += 2;
				if (tk[s->pos].type != T_LPAREN) return -1;
				s->pos++;
				if (pEval(s, &r2) || tk[s->pos].type != T_RPAREN) return -1;
				s->pos++;
				*res = isAnd ? (((r1 == 1) && (r2 == 1)) ? 1 : 0)
						: (((r1 == 1) || (r2 == 1)) ? 1 : 0);
				return 0;
			}
			*res = r1; return 0;
		}
		s->pos = save; s->err = 0;   /* it was "(expr) cmp expr" after all */
	}
	a = pExpr(s);
	if (s->err) return -1;
	switch (tk[s->pos].type) {
	case T_GT:
		s->pos++;
		if (tk[s->pos].type == T_EQ) {
			s->pos++; b = pExpr(s); if (s->err) return -1;
			*res = (a.f >= b.f) ? 1 : 0; return 0;
		}
		b = pExpr(s); if (s->err) return -1;
		*res = (a.f > b.f) ? 1 : 0; return 0;
	case T_LT:
		s->pos++;
		if (tk[s->pos].type == T_EQ) {
			s->pos++; b = pExpr(s); if (s->err) return -1;
			*res = (a.f <= b.f) ? 1 : 0; return 0;
		}
		b = pExpr(s); if (s->err) return -1;
		*res = (a.f < b.f) ? 1 : 0; return 0;
	case T_EQ:
		if (tk[s->pos + 1].type != T_EQ) return -1;
		s->pos += 2; b = pExpr(s); if (s->err) return -1;
		*res = (a.f == b.f) ? 1 : 0; return 0;
	case T_NOT:
		if (tk[s->pos + 1].type != T_EQ) return -1;
		s->pos += 2; b = pExpr(s); if (s->err) return -1;
		*res = (a.f != b.f) ? 1 : 0; return 0;
	default:
		return -1;
	}
}

/* calc.c:36-52 (calcSolve) / :78-102 (calcDiff) */
static int calcRun(const Calc *c, const double *vals, int dvar, double m,
		double *out)
{
	PS s;
	Dual r;
	s.c = c; s.pos = 0; s.vals = vals; s.dvar = dvar; s.m = m; s.err = 0;
	r = pExpr(&s);
	if (s.err || c->tok[s.pos].type != T_END) FAIL("calc: syntax error");
	*out = (dvar >= 0) ? r.d : r.f;
	if (isnan(*out)) FAIL("calc: result is NaN");
	return 0;
}

/* standalone evaluation for fuzzing against the reference's calculon */
typedef struct { int n; const char **names; int *used; } FuzzCtx;
static int fuzzGetVar(void *ctx, const char *name)
{
	FuzzCtx *f = ctx; int i;
	for (i = 0; i < f->n; i++)
		if (!strcmp(name, f->names[i])) { f->used[i] = 1; return i; }
	return -1;
}

int eoCalcEval(const char *expr, int nvars, const char **names,
		const double *values, double minDiv, double *value, double *derivs)
{
	FuzzCtx f; Calc *c; double *vals; int i, rc = 0;
	f.n = nvars; f.names = names; f.used = calloc(nvars + 1, sizeof(int));
	c = calcNew(expr, fuzzGetVar, &f);
	if (c == NULL) { free(f.used); return -1; }
	vals = calloc(c->nvars + 1, sizeof(double));
	for (i = 0; i < c->nvars; i++) vals[i] = values[c->varRef[i]];
	if (calcRun(c, vals, -1, minDiv, value)) rc = -1;
	for (i = 0; i < nvars; i++) derivs[i] = NAN;
	for (i = 0; rc == 0 && i < c->nvars; i++)
		if (calcRun(c, vals, i, minDiv, &derivs[c->varRef[i]])) rc = -1;
	free(vals); free(f.used); calcFree(c);
	return rc;
}

/* ------------------------------------------------------------------------ *
 * Piecewise tables: libs/simulator/math/piecewise.c
 * ------------------------------------------------------------------------ */
typedef struct { char type; double *xy; double *m; int len; } Piecewise;

static Piecewise *pwNew(const double *xy, int len, char type)
{
	Piecewise *r;
	if (xy == NULL || len < 1 || (type != 'c' && type != 'l')) return NULL;
	r = calloc(1, sizeof(Piecewise));
	r->type = type; r->len = len;
	r->xy = malloc(sizeof(double) * 2 * len);
	memcpy(r->xy, xy, sizeof(double) * 2 * len);
	return r;
}
static void pwFree(Piecewise *r)
{
	if (r) { free(r->xy); free(r->m); free(r); }
}
#define PX(r, i) ((r)->xy[2 * (i)])
#define PY(r, i) ((r)->xy[2 * (i) + 1])

/* piecewise.c:44-67: the cached index only speeds the walk up, the landing
 * index is a function of x0 alone */
static void pwSetIndex(const Piecewise *r, int *index, double x0)
{
	if ((*index < 0) || (*index > r->len - 1)) *index = 0;
	while ((*index >= 0) && (*index < r->len - 1)) {
		if (PX(r, *index) <= x0) {
			if (*index == r->len - 2) return;
			if (PX(r, *index + 1) > x0) return;
			(*index)++;
		} else {
			if (*index == 0) return;
			(*index)--;
		}
	}
}

/* piecewise.c:183-201 (linear) and :108-181 (natural cubic spline) */
static int pwInitialize(Piecewise *r)
{
	int i, n = r->len - 1;
	for (i = 0; i < r->len - 1; i++)
		if (PX(r, i) > PX(r, i + 1)) FAIL("piecewise: x not sorted");
	free(r->m);
	r->m = malloc(sizeof(double) * r->len);
	if (r->type == 'l') {
		for (i = 0; i < r->len; i++)
			r->m[i] = (i == r->len - 1) ? 0
				: (PY(r, i + 1) - PY(r, i)) / (PX(r, i + 1) - PX(r, i));
		return 0;
	}
	if (r->len == 2) { r->m[0] = 0.0; r->m[1] = 0.0; return 0; }
	{
		double *h = malloc(sizeof(double) * n), *b = malloc(sizeof(double) * n);
		double *u = malloc(sizeof(double) * n), *v = malloc(sizeof(double) * n);
		int bad = 0;
		for (i = 0; i < n; i++) {
			h[i] = PX(r, i + 1) - PX(r, i);
			b[i] = 6.0 * (PY(r, i + 1) - PY(r, i)) / h[i];
		}
		u[1] = 2.0 * h[0] + 2.0 * h[1];
		if (u[1] == 0.0) bad = 1;
		v[1] = b[1] - b[0];
		for (i = 2; i < n && !bad; i++) {
			u[i] = 2.0 * h[i] + 2.0 * h[i - 1] - (h[i - 1] * h[i - 1] / u[i - 1]);
			if (u[i] == 0.0) bad = 1;
			v[i] = b[i] - b[i - 1] - (h[i - 1] * v[i - 1] / u[i - 1]);
		}
		r->m[0] = 0.0; r->m[n] = 0.0;
		for (i = n - 1; i > 0 && !bad; i--)
			r->m[i] = (v[i] - h[i] * r->m[i + 1]) / u[i];
		free(h); free(b); free(u); free(v);
		if (bad) FAIL("piecewise: singular spline");
	}
	return 0;
}

/* piecewise.c:204-222.  The "== length-1" and "index == 0" (pointer compare)
 * branches of the reference are dead, so it is always the next table x. */
static void pwGetNextX(const Piecewise *r, int *index, double x0, double *x)
{
	pwSetIndex(r, index, x0);
	if (*index == r->len - 1) *x = HUGE_VAL;
	else *x = PX(r, *index + 1);
}

/* piecewise.c:226-263 */
static int pwCalcValue(const Piecewise *r, int *index, double x0, double *y0,
		double *dydx0)
{
	int i;
	if (r->m == NULL) FAIL("piecewise: not initialised");
	pwSetIndex(r, index, x0);
	i = *index;
	if ((i == 0) || (i >= r->len - 2)) {
		if (PX(r, 0) > x0) { *y0 = PY(r, 0); *dydx0 = 0; return 0; }
		if (PX(r, r->len - 1) < x0) { *y0 = PY(r, r->len - 1); *dydx0 = 0; return 0; }
	}
	if (r->type == 'l') {
		*y0 = PY(r, i) + r->m[i] * (x0 - PX(r, i));
		*dydx0 = r->m[i];
	} else {
		double dx0 = x0 - PX(r, i), dy = PY(r, i + 1) - PY(r, i);
		double dx = PX(r, i + 1) - PX(r, i);
		double a = dy / dx - dx * (r->m[i + 1] + r->m[i] * 2.0) / 6.0;
		double b = 0.5 * r->m[i];
		double c = (r->m[i + 1] - r->m[i]) / (6.0 * dx);
		*y0 = PY(r, i) + dx0 * (a + dx0 * (b + dx0 * c));
		*dydx0 = a + dx0 * (2 * b + 3 * c * dx0);
	}
	return 0;
}

/* ------------------------------------------------------------------------ *
 * Break-point test: libs/simulator/math/checkbreak.c:43-67,71-95
 * ------------------------------------------------------------------------ */
typedef struct {
	char units; double maxAngle; double x[3], t[3], theta[3]; int n;
} CheckBreak;

static void cbInitialize(CheckBreak *r, const Control *c, double ic)
{
	int i;
	r->maxAngle = (r->units == 'V') ? c->maxAngleV : c->maxAngleA;
	for (i = 0; i < 3; i++) { r->x[i] = ic; r->t[i] = 0.0; }
	/* theta[] is deliberately NOT reset (calloc'ed once, checkbreak.c:128) */
	r->n = 0;
}

static int cbIsBreak(CheckBreak *r, const Control *c, double x)
{
	int n, p;
	if (c->time > r->t[r->n % 3]) r->n++;
	n = r->n % 3;
	p = (r->n - 1) % 3;
	if (p < 0) p = 0; /* unreachable: time > 0 in transient forces n >= 1 */
	r->x[n] = x;
	r->t[n] = c->time;
	r->theta[n] = atan2(r->x[n] - r->x[p], r->t[n] - r->t[p]);
	return (fabs(r->theta[n] - r->theta[p]) > r->maxAngle) ? 1 : 0;
}

/* ------------------------------------------------------------------------ *
 * Newton convergence test: libs/simulator/math/checklinear.c:39-91
 * ------------------------------------------------------------------------ */
typedef struct { char state; char units; double lastV; } CheckLinear;

static void clInitialize(CheckLinear *r, double ic) { r->state = 'a'; r->lastV = ic; }

/* outcome of the two tolerance tests of the last clIsLinear call (1 = both
 * passed); read by the attempt log (test instrumentation) */
static int g_clPassed;

static int clIsLinear(CheckLinear *r, const Control *c, double V, double calcedV)
{
	double tol = (r->units == 'V') ? c->vntol
		: (r->units == 'A') ? c->abstol : c->captol;
	double e = c->reltol * MaxAbs(r->lastV, V) + tol;
	g_clPassed = 0;
	if (fabs(r->lastV - V) > e) { r->state = 'a'; r->lastV = V; return 0; }
	e = c->reltol * MaxAbs(calcedV, V) + tol;
	if (fabs(calcedV - V) > e) { r->state = 'a'; r->lastV = V; return 0; }
	g_clPassed = 1;
	r->lastV = V;
	if (r->state == 'b') { r->state = 'c'; return 1; }
	if (r->state == 'c') return 1;
	r->state = 'b';
	return 0;
}

/* ------------------------------------------------------------------------ *
 * Integrator (companion models + LTE): libs/simulator/math/integrator.c
 * The five rings live in ONE array in the struct's order t,h,x,y,f so that
 * the reference's negative "(n-2)%N" index at n==1 (integrator.c:90-96 with
 * struct :44-57) lands on the same neighbours: f[-1]=y[3], x[-1]=h[3],
 * h[-1]=t[3].
 * ------------------------------------------------------------------------ */
#define NI 4
typedef struct {
	char units; double abstol;
	const double *ydtdx;
	double y0, dydx0;
	double ring[5 * NI];
	int n;
} Integrator;
#define I_T(r, k) ((r)->ring[0 * NI + (k)])
#define I_H(r, k) ((r)->ring[1 * NI + (k)])
#define I_X(r, k) ((r)->ring[2 * NI + (k)])
#define I_Y(r, k) ((r)->ring[3 * NI + (k)])
#define I_F(r, k) ((r)->ring[4 * NI + (k)])

static double DD1(double xnp1, double xn, double hn)
{
	return (hn == 0.0) ? 0.0 : ((xnp1 - xn) / hn);
}
static double DD2(double xnp1, double xn, double x1, double hn, double h1)
{
	return (DD1(xnp1, xn, hn) - DD1(xn, x1, h1)) / (hn + h1);
}
static double DD3(double xnp1, double xn, double x1, double x2, double hn,
		double h1, double h2)
{
	return (DD2(xnp1, xn, x1, hn, h1) - DD2(xn, x1, x2, h1, h2)) / (hn + h1 + h2);
}

/* integrator.c:167-199 */
static int itInitialize(Integrator *r, const Control *c, double ic,
		const double *ydtdx)
{
	int i;
	r->ydtdx = ydtdx;
	r->y0 = 0.0; r->dydx0 = 0.0; r->n = 0;
	for (i = 0; i < NI; i++) {
		I_T(r, i) = 0.0; I_H(r, i) = c->tstop; I_Y(r, i) = 0.0;
		I_X(r, i) = ic; I_F(r, i) = *ydtdx;
	}
	if (r->units == 'V') r->abstol = c->vntol;
	else if (r->units == 'A') r->abstol = c->abstol;
	else if (r->units == 'F') r->abstol = c->captol;
	else FAIL("integrator: bad units");
	return 0;
}

/* integrator.c:106-163 */
static int itIntegrate(Integrator *r, const Control *c, double x0,
		double *dydx0, double *y0)
{
	double t0 = c->time, mult;
	int n, p;
	if (isnan(x0)) FAIL("integrator: NaN input");
	if (t0 > I_T(r, (r->n + 1) % NI)) r->n++;
	n = r->n % NI; p = (r->n - 1) % NI;
	I_H(r, n) = t0 - I_T(r, n);
	I_T(r, (r->n + 1) % NI) = t0;
	I_X(r, n) = x0;
	I_Y(r, n) = r->dydx0 * x0 - r->y0;
	I_F(r, n) = *r->ydtdx;
	if (c->integratorOrder < 2) {
		*dydx0 = I_F(r, p) / I_H(r, n);
		*y0 = (*dydx0) * I_X(r, n);
	} else {
		*dydx0 = 2.0 * I_F(r, n) / I_H(r, n);
		mult = I_F(r, n) / I_F(r, p);
		if (isnan(mult)) mult = 1 / c->gmin;
		*y0 = (*dydx0) * I_X(r, n) + mult * I_Y(r, n);
	}
	r->dydx0 = *dydx0;
	r->y0 = *y0;
	return 0;
}

/* integrator.c:61-102 */
static int itNextStep(Integrator *r, const Control *c, double x0, double *h)
{
	double y0, ey, eyp, e, dd;
	int n = r->n % NI, p1 = (r->n - 1) % NI, p2 = (r->n - 2) % NI;
	if (isnan(x0)) FAIL("integrator: NaN input");
	y0 = r->dydx0 * x0 - r->y0;
	ey = c->reltol * MaxAbs(I_Y(r, n), y0) + r->abstol;
	eyp = c->reltol * MaxAbs(MaxAbs(x0, I_X(r, n)) * (*r->ydtdx / I_H(r, n)),
			c->chgtol);
	e = MaxAbs(eyp, ey);
	if (c->integratorOrder < 2) {
		dd = DD2(((*r->ydtdx) * x0), (I_F(r, n) * I_X(r, n)),
				(I_F(r, p1) * I_X(r, p1)), I_H(r, n), I_H(r, p1));
		dd *= 1.0 / 2.0;
		*h = c->trtol * e / MaxAbs(dd, r->abstol);
	} else {
		dd = DD3(((*r->ydtdx) * x0), (I_F(r, n) * I_X(r, n)),
				(I_F(r, p1) * I_X(r, p1)), (I_F(r, p2) * I_X(r, p2)),
				I_H(r, n), I_H(r, p1), I_H(r, p2));
		dd *= 1.0 / 12.0;
		*h = sqrtf(c->trtol * e / MaxAbs(dd, r->abstol)); /* float, :99 */
	}
	if (isnan(*h)) FAIL("integrator: NaN step");
	return 0;
}

/* ------------------------------------------------------------------------ *
 * Source waveforms: libs/simulator/math/waveform.c
 * ------------------------------------------------------------------------ */
typedef struct {
	char type;
	double arg[7];      /* as given (HUGE_VAL = unset) */
	double eff[7];      /* after default substitution, waveform.c:342-413 */
	Piecewise *pw; int index; double dc;
} Waveform;

static int wfInitialize(Waveform *w, const Control *c)
{
	double tstep = c->tstep, tstop = c->tstop + c->tstep, fmin = 1 / tstop;
	double dvdt;
	int i;
	if (tstep == HUGE_VAL || tstop == HUGE_VAL) FAIL("waveform: bad tstep");
	for (i = 0; i < 7; i++) w->eff[i] = w->arg[i];
	switch (w->type) {
	case 'p': case 'g':      /* td tf pw per (tr is not defaulted) */
		if (w->arg[2] == HUGE_VAL) w->eff[2] = 0.0;
		if (w->arg[4] == HUGE_VAL) w->eff[4] = tstep;
		if (w->arg[5] == HUGE_VAL) w->eff[5] = tstop;
		if (w->arg[6] == HUGE_VAL) w->eff[6] = tstop;
		break;
	case 's':
		if (w->arg[2] == HUGE_VAL) w->eff[2] = fmin;
		if (w->arg[3] == HUGE_VAL) w->eff[3] = 0.0;
		if (w->arg[4] == HUGE_VAL) w->eff[4] = 0.0;
		break;
	case 'e':
		if (w->arg[2] == HUGE_VAL) w->eff[2] = 0.0;
		if (w->arg[3] == HUGE_VAL) w->eff[3] = tstep;
		if (w->arg[4] == HUGE_VAL) w->eff[4] = tstep + w->eff[2];
		if (w->arg[5] == HUGE_VAL) w->eff[5] = tstep;
		break;
	case 'l': case 'c':
		TRY(pwInitialize(w->pw));
		TRY(pwCalcValue(w->pw, &w->index, 0.0, &w->dc, &dvdt));
		break;
	case 'f':
		if (w->arg[2] == HUGE_VAL) w->eff[2] = fmin;
		if (w->arg[3] == HUGE_VAL) w->eff[3] = 0.0;
		if (w->arg[4] == HUGE_VAL) w->eff[4] = fmin;
		break;
	default: break;
	}
	return 0;
}

/* the value a source holds at the operating point: waveform.c:97-165 */
static double wfDC(const Waveform *w)
{
	if (w->type == 'l' || w->type == 'c') return w->dc;
	return w->arg[0];
}

/* waveform.c:169-220; *nextStep is left untouched where the reference does */
static void wfNextStep(Waveform *w, const Control *c, double *nextStep)
{
	double tp, time = c->time;
	const double *a = w->eff;
	switch (w->type) {
	case 'p':
		tp = fmod(time, a[6]);
		*nextStep = a[2] - tp;
		if (*nextStep > 0) break;
		*nextStep += a[3];
		if (*nextStep > 0) break;
		*nextStep += a[5];
		if (*nextStep > 0) break;
		*nextStep += a[4];
		if (*nextStep > 0) break;
		*nextStep = a[6] - tp;
		break;
	case 's':
		if (time < a[3]) *nextStep = a[3];  /* absolute, sic (:197) */
		break;
	case 'e':
		if (time < a[2]) *nextStep = a[2] - time;
		else if (time < a[4]) *nextStep = a[4] - time;
		break;
	case 'l': case 'c':
		pwGetNextX(w->pw, &w->index, time, &tp);
		if (tp != HUGE_VAL) *nextStep = tp - time;
		break;
	default: break;
	}
}

/* Cephes-style erf is only needed for the Gauss waveform; libm's erf agrees
 * to the last few ulp (waveform.c:283-284, math/netlib.c:719-723). */
static int wfCalcValue(Waveform *w, const Control *c, double *value)
{
	double time = c->time, tn, dvdt;
	const double *a = w->eff;
	*value = 0.0;
	switch (w->type) {
	case 'p': {
		double v1 = a[0], v2 = a[1], td = a[2], tr = a[3], tf = a[4], pw = a[5];
		tn = fmod(time, a[6]);
		if (tn <= td) *value = v1;
		else if (tn <= (tr + td)) *value = v1 + ((v2 - v1) / tr) * (tn - td);
		else if (tn <= (tr + td + pw)) *value = v2;
		else if (tn <= (tr + td + pw + tf))
			*value = v2 - ((v2 - v1) / (tf)) * (tn - (tr + td + pw));
		else *value = v1;
		break;
	}
	case 'g': {
		double v1 = a[0], v2 = a[1], td = a[2], tr = a[3], tf = a[4], pw = a[5];
		double vo, va;
		tn = fmod(time, a[6]);
		tr = (tr / 0.672) * 0.281 * 2;
		tf = (tf / 0.672) * 0.281 * 2;
		tr = (tn - (tr / 0.281) * 0.672 - td) / tr;
		tf = (tn - (tf / 0.281) * 0.672 - td - pw) / tf;
		vo = erf(tr); va = erf(tf);
		*value = v1;
		*value += 0.5 * (v2 - v1) * (1 + vo);
		*value += 0.5 * (v1 - v2) * (1 + va);
		break;
	}
	case 's':
		if (time <= a[3]) *value = a[0];
		else *value = a[0] + a[1] * exp(-(time - a[3]) * a[4]) *
			sin(2 * M_PI * a[2] * (time - a[3]));
		break;
	case 'e': {
		double v1 = a[0], v2 = a[1], td1 = a[2], tau1 = a[3], td2 = a[4], tau2 = a[5];
		if (time <= td1) *value = v1;
		else if (time <= td2) *value = v1 + (v2 - v1) * (1 - exp(-(time - td1) / tau1));
		else *value = v1 + (v2 - v1) * (1 - exp(-(time - td1) / tau1)) +
			(v1 - v2) * (1 - exp(-(time - td2) / tau2));
		break;
	}
	case 'l': case 'c':
		TRY(pwCalcValue(w->pw, &w->index, time, value, &dvdt));
		break;
	case 'f':
		*value = a[0] + a[1] * sin(2 * M_PI * a[2] * time +
				a[3] * sin(2 * M_PI * a[4] * time));
		break;
	default: FAIL("waveform: bad type");
	}
	return 0;
}

/* ------------------------------------------------------------------------ *
 * Devices
 * ------------------------------------------------------------------------ */
enum { D_R, D_C, D_L, D_V, D_I, D_T, D_VI, D_BI, D_BV, D_BC };

typedef struct { int row; double G; } BVar;   /* variable_ in nonlinear_*.c */

typedef struct {
	int type;
	char *refdes;
	int pin[4];
	int rowR, rowM, rowS;        /* extra rows (branch / internal) */
	/* parameters (by value; the reference holds pointers) */
	double val;                  /* R, C, L, DC */
	double Z0, Td, loss;
	int bypass;             /* engine extension under test: the section is an ideal through-connection */
	/* state */
	double G, Ieq;               /* C: G,Ieq  L: R,Veq  VI: G,Ieq  BC: cap G,Ieq */
	double a;                    /* VI multiplier */
	double dc;                   /* sources: present value */
	double Vr, Vs, ic[6];        /* T-line: Ir Is Vk Vj Vl Vm */
	int hint;                    /* history search start */
	Integrator integ;
	CheckBreak cb, cb2;
	CheckLinear cl;
	Waveform *wave;
	Piecewise *vi, *ta; int viIndex, taIndex;
	/* behavioural */
	Calc *calc; int usesTime; int nbv; BVar *bv;
	double Ic, In, Beq, BeqCalc; /* Ic/Vc/Cc, In/Vn/Cn, Ieq/Veq/Ceq, *Calc */
} Dev;

typedef struct { int row, col; } NodeIdx;

struct eo_sim {
	Control ctl;
	char **rowName; int nrows;      /* row 0 = ground "v(0)" */
	NodeIdx *nodes; int nnodes;     /* excluding the ground node */
	Dev **dev; int ndev;            /* insertion order */
	int locked;
	int N;
	double *A, *B, *X;              /* dense A[N*N] row-major over unknowns */
	double *W, *y;                  /* LU work */
	int *rdone, *cdone, *prow, *pcol;
	double *pinv;
	/* history: matrix.c:372-392, history.c */
	double *htime; double *hX; int npts, hcap;
	/* fixed pivot replay (tests only) */
	int *fixPc[2], *fixPr[2];
	int phase;
	eo_stats st;
	/* attempt log (tests only): one entry per transient attempt, in order */
	int logOn, logN, logCap;
	double *logTime; int *logFlag;
	/* per linearize pass (simulatorSolve :57-58), operating point included: bit k =
	 * the tolerance tests of the k-th nonlinear device passed (checklinear.c:62-77) */
	int linN, linCap; int *linMask;
	/* stamp watch (tests only): which A entries / B rows differed between two
	 * consecutive solves of ONE simulatorSolve loop, per phase */
	/* pivots of the FIRST factorisation of each phase of the last run (tests only) */
	int *firstPc[2], *firstPr[2];
	int firstSeen[2];
	int watchOn;
	double *watchA, *watchB;        /* stamps at the previous solve of this loop */
	char *changedA[2], *changedB[2];
};

/* --- rows & pattern: core/row.c:46-66,163-204, core/matrix.c:292-346 ------ */
static int rowFindOrAdd(eo_sim *s, char type, const char *name)
{
	int i;
	char *full;
	for (i = 0; i < s->nrows; i++) {
		const char *rn = s->rowName[i];
		if (type == 0) {
			if (!strcmp(rn, name)) return i;
		} else if (rn[0] == type && strlen(name) == strlen(rn) - 3 &&
				!strncmp(&rn[2], name, strlen(name))) {
			return i;
		}
	}
	if (type == 0) {
		size_t L = strlen(name);
		if (L < 4) return -1;
		if (name[0] != 'i' && name[0] != 'v' && name[0] != 'I' && name[0] != 'V')
			return -1;
		full = strdup(name);
		full[0] = tolower((unsigned char)full[0]);
	} else {
		full = malloc(strlen(name) + 4);
		sprintf(full, "%c(%s)", type, name);
	}
	s->rowName = realloc(s->rowName, sizeof(char *) * (s->nrows + 1));
	s->rowName[s->nrows] = full;
	return s->nrows++;
}

static void nodeAdd(eo_sim *s, int row, int col)
{
	int i;
	if (row == 0 || col == 0) return;
	for (i = 0; i < s->nnodes; i++)
		if (s->nodes[i].row == row && s->nodes[i].col == col) return;
	s->nodes = realloc(s->nodes, sizeof(NodeIdx) * (s->nnodes + 1));
	s->nodes[s->nnodes].row = row; s->nodes[s->nnodes].col = col;
	s->nnodes++;
}

/* stamp primitives: node.c:86-118, row.c:85-116 (ground is a sink) */
static void aPlus(eo_sim *s, int row, int col, double d)
{
	if (row && col) s->A[(row - 1) * s->N + (col - 1)] += d;
}
static void aSet(eo_sim *s, int row, int col, double v)
{
	if (row && col) s->A[(row - 1) * s->N + (col - 1)] = v;
}
static void bPlus(eo_sim *s, int row, double d) { if (row) s->B[row - 1] += d; }
static double xGet(const eo_sim *s, int row) { return row ? s->X[row - 1] : 0.0; }

static Dev *devNew(eo_sim *s, int type, const char *refdes, int npins,
		const char **pins)
{
	Dev *d;
	int i;
	if (s == NULL || refdes == NULL) return NULL;
	if (s->locked) {
		if (!g_quiet) fprintf(stderr, "eis_oracle: can't add a device to a "
				"simulator that has run\n");
		return NULL;
	}
	for (i = 0; i < npins; i++) if (pins[i] == NULL) return NULL;
	/* duplicate refdes is checked when the device is appended to the list
	 * (device.c:166-177); rows created before that stay, as in the reference */
	s->dev = realloc(s->dev, sizeof(Dev *) * (s->ndev + 1));
	d = calloc(1, sizeof(Dev));
	s->dev[s->ndev] = d;       /* becomes visible once devCommit bumps ndev */
	d->type = type;
	d->refdes = strdup(refdes);
	for (i = 0; i < npins; i++) d->pin[i] = rowFindOrAdd(s, 'v', pins[i]);
	return d;
}

static int devCommit(eo_sim *s, Dev *d)
{
	int i;
	for (i = 0; i < s->ndev; i++)
		if (!strcmp(s->dev[i]->refdes, d->refdes))
			FAIL("%s is listed twice", d->refdes);
	s->ndev++;
	return 0;
}

eo_sim *eoNew(void)
{
	eo_sim *s = calloc(1, sizeof(eo_sim));
	controlInit(&s->ctl);
	s->rowName = malloc(sizeof(char *));
	s->rowName[0] = strdup("v(0)");
	s->nrows = 1;
	return s;
}

static void devFree(Dev *d)
{
	free(d->refdes);
	if (d->wave) { pwFree(d->wave->pw); free(d->wave); }
	pwFree(d->vi); pwFree(d->ta);
	calcFree(d->calc); free(d->bv);
}

void eoDestroy(eo_sim *s)
{
	int i;
	if (s == NULL) return;
	for (i = 0; i < s->ndev; i++) { devFree(s->dev[i]); free(s->dev[i]); }
	for (i = 0; i < s->nrows; i++) free(s->rowName[i]);
	free(s->rowName); free(s->nodes); free(s->dev);
	free(s->A); free(s->B); free(s->X); free(s->W); free(s->y);
	free(s->rdone); free(s->cdone); free(s->prow); free(s->pcol); free(s->pinv);
	free(s->htime); free(s->hX);
	for (i = 0; i < 2; i++) { free(s->fixPc[i]); free(s->fixPr[i]); }
	free(s->logTime); free(s->logFlag); free(s->linMask);
	free(s->watchA); free(s->watchB);
	for (i = 0; i < 2; i++) { free(s->changedA[i]); free(s->changedB[i]); }
	for (i = 0; i < 2; i++) { free(s->firstPc[i]); free(s->firstPr[i]); }
	free(s);
}
void eoFree(void *p) { free(p); }

/* resistor.c:111-144 */
int eoAddResistor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double R)
{
	const char *pins[2] = { p, n };
	Dev *d = devNew(s, D_R, refdes, 2, pins);
	if (!d) return -1;
	d->val = R;
	nodeAdd(s, d->pin[0], d->pin[0]); nodeAdd(s, d->pin[0], d->pin[1]);
	nodeAdd(s, d->pin[1], d->pin[0]); nodeAdd(s, d->pin[1], d->pin[1]);
	return devCommit(s, d);
}

/* capacitor.c:216-260 */
int eoAddCapacitor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double C)
{
	const char *pins[2] = { p, n };
	Dev *d = devNew(s, D_C, refdes, 2, pins);
	if (!d) return -1;
	d->val = C;
	d->integ.units = 'V';
	itInitialize(&d->integ, &s->ctl, 0.0, &d->val);
	nodeAdd(s, d->pin[0], d->pin[0]); nodeAdd(s, d->pin[1], d->pin[0]);
	nodeAdd(s, d->pin[0], d->pin[1]); nodeAdd(s, d->pin[1], d->pin[1]);
	return devCommit(s, d);
}

/* inductor.c:222-269 */
int eoAddInductor(eo_sim *s, const char *refdes, const char *p, const char *n,
		double L)
{
	const char *pins[2] = { p, n };
	Dev *d = devNew(s, D_L, refdes, 2, pins);
	if (!d) return -1;
	d->val = L;
	d->integ.units = 'A';
	itInitialize(&d->integ, &s->ctl, 0.0, &d->val);
	d->rowR = rowFindOrAdd(s, 'i', refdes);
	nodeAdd(s, d->rowR, d->pin[0]); nodeAdd(s, d->rowR, d->pin[1]);
	nodeAdd(s, d->pin[0], d->rowR); nodeAdd(s, d->pin[1], d->rowR);
	nodeAdd(s, d->rowR, d->rowR);
	return devCommit(s, d);
}

/* simulator.c:430-457, source_v.c:202-252, source_i.c:195-238.  The reference
 * appends a source to the device list BEFORE configuring it. */
int eoAddSource(eo_sim *s, const char *refdes, const char *p, const char *n,
		char type, double dc, char stimulus, const double *args,
		const double *table, int tableLen)
{
	const char *pins[2] = { p, n };
	Dev *d;
	int i;
	type = tolower((unsigned char)type);
	d = devNew(s, (type == 'v') ? D_V : D_I, refdes, 2, pins);
	if (!d) return -1;
	TRY(devCommit(s, d));
	if (type != 'v' && type != 'i') FAIL("source type must be i or v");
	if (type == 'v' && d->pin[0] == 0 && d->pin[1] == 0)
		FAIL("Source %s has both nodes shorted to 0!", refdes);
	d->val = dc;
	if (stimulus != 0) {
		Waveform *w;
		if (!strchr("pgselcf", stimulus)) FAIL("bad waveform type");
		w = calloc(1, sizeof(Waveform));
		w->type = stimulus;
		if (stimulus == 'l' || stimulus == 'c') {
			w->pw = pwNew(table, tableLen, stimulus);
			if (w->pw == NULL) { free(w); FAIL("bad table"); }
		} else {
			if (args == NULL) { free(w); FAIL("missing waveform args"); }
			for (i = 0; i < 7; i++) w->arg[i] = args[i];
		}
		d->wave = w;
	}
	d->cb.units = (type == 'v') ? 'V' : 'A';
	cbInitialize(&d->cb, &s->ctl, 0.0);
	if (type == 'v') {
		d->rowR = rowFindOrAdd(s, 'i', refdes);
		nodeAdd(s, d->rowR, d->pin[0]); nodeAdd(s, d->rowR, d->pin[1]);
		nodeAdd(s, d->pin[0], d->rowR); nodeAdd(s, d->pin[1], d->rowR);
	}
	return 0;
}

/* tline.c:353-457 */
int eoAddTLine(eo_sim *s, const char *refdes, const char *pl, const char *nl,
		const char *pr, const char *nr, double Z0, double Td, double loss)
{
	const char *pins[4] = { pl, nl, pr, nr };
	char nm[300];
	Dev *d = devNew(s, D_T, refdes, 4, pins);
	int K, J, L, M, R, S;
	if (!d) return -1;
	if (d->pin[0] == 0 && d->pin[1] == 0)
		FAIL("T-Line %s has both input nodes shorted to 0!", refdes);
	if (d->pin[2] == 0 && d->pin[3] == 0)
		FAIL("T-Line %s has both output nodes shorted to 0!", refdes);
	d->Z0 = Z0; d->Td = Td; d->loss = loss;
	snprintf(nm, sizeof nm, "%s#in1", refdes);
	d->rowR = rowFindOrAdd(s, 'i', nm);
	snprintf(nm, sizeof nm, "%s#in2", refdes);
	d->rowS = rowFindOrAdd(s, 'i', nm);
	K = d->pin[0]; J = d->pin[1]; L = d->pin[2]; M = d->pin[3];
	R = d->rowR; S = d->rowS;
	nodeAdd(s, R, K); nodeAdd(s, R, J); nodeAdd(s, K, R); nodeAdd(s, J, R);
	nodeAdd(s, R, L); nodeAdd(s, R, M); nodeAdd(s, S, L); nodeAdd(s, S, M);
	nodeAdd(s, L, S); nodeAdd(s, M, S); nodeAdd(s, S, K); nodeAdd(s, S, J);
	nodeAdd(s, R, R); nodeAdd(s, S, S); nodeAdd(s, K, S); nodeAdd(s, J, S);
	nodeAdd(s, L, R); nodeAdd(s, M, R);
	d->cb.units = 'V'; d->cb2.units = 'V';
	cbInitialize(&d->cb, &s->ctl, 0.0); cbInitialize(&d->cb2, &s->ctl, 0.0);
	return devCommit(s, d);
}

/* vicurve.c:298-359 */
int eoAddVICurve(eo_sim *s, const char *refdes, const char *p, const char *n,
		const double *vi, int viLen, char viType,
		const double *ta, int taLen, char taType)
{
	const char *pins[2] = { p, n };
	Dev *d = devNew(s, D_VI, refdes, 2, pins);
	int K, J, R, M;
	if (!d) return -1;
	d->vi = pwNew(vi, viLen, viType);
	if (d->vi == NULL) FAIL("bad VI table");
	if (ta != NULL && taLen > 0) {
		d->ta = pwNew(ta, taLen, taType);
		if (d->ta == NULL) FAIL("bad TA table");
	}
	d->cl.units = 'A'; clInitialize(&d->cl, 0.0);
	d->cb.units = 'A'; cbInitialize(&d->cb, &s->ctl, 0.0);
	d->rowR = rowFindOrAdd(s, 'i', refdes);
	d->rowM = rowFindOrAdd(s, 'v', refdes);
	K = d->pin[0]; J = d->pin[1]; R = d->rowR; M = d->rowM;
	nodeAdd(s, R, K); nodeAdd(s, R, M); nodeAdd(s, K, R); nodeAdd(s, M, R);
	nodeAdd(s, J, J); nodeAdd(s, J, M); nodeAdd(s, M, J); nodeAdd(s, M, M);
	return devCommit(s, d);
}

/* deviceNonlinearGetVariable: nonlinear_i.c:147-192, _v.c:119-160, _c.c:127-169 */
typedef struct { eo_sim *s; Dev *d; } BCtx;
static int bGetVar(void *ctx, const char *name)
{
	BCtx *b = ctx;
	Dev *d = b->d;
	eo_sim *s = b->s;
	int row;
	if (!strcmp("time", name)) { d->usesTime = 1; return 0x40000000; }
	if (!((name[0] == 'i' || name[0] == 'v') && name[1] == '(')) {
		if (!g_quiet) fprintf(stderr, "eis_oracle: B element variables must be "
				"i(...), v(...) or time -- not %s\n", name);
		return -1;
	}
	row = rowFindOrAdd(s, 0, name);
	if (row < 0) return -1;
	if (d->type == D_BI) { nodeAdd(s, d->pin[1], row); nodeAdd(s, d->rowM, row); }
	else nodeAdd(s, d->rowR, row);
	d->bv = realloc(d->bv, sizeof(BVar) * (d->nbv + 1));
	d->bv[d->nbv].row = row; d->bv[d->nbv].G = 0.0;
	d->nbv++;
	return row;
}

/* simulator.c:363-392; nonlinear_i.c:368-425, _v.c:330-392, _c.c:429-495 */
int eoAddNonlinearSource(eo_sim *s, const char *refdes, const char *p,
		const char *n, char type, const char *equation)
{
	const char *pins[2] = { p, n };
	Dev *d;
	BCtx ctx;
	int K, J;
	type = tolower((unsigned char)type);
	if (equation == NULL) return -1;
	d = devNew(s, (type == 'i') ? D_BI : (type == 'v') ? D_BV : D_BC, refdes, 2, pins);
	if (!d) return -1;
	if (type != 'i' && type != 'v' && type != 'c')
		FAIL("Nonlinear Source type must be either i, v or c, not %c", type);
	K = d->pin[0]; J = d->pin[1];
	if (type == 'i') {
		d->cl.units = 'A'; d->cb.units = 'A';
		d->rowR = rowFindOrAdd(s, 'i', refdes);
		d->rowM = rowFindOrAdd(s, 'v', refdes);
		nodeAdd(s, d->rowR, K); nodeAdd(s, d->rowR, d->rowM);
		nodeAdd(s, K, d->rowR); nodeAdd(s, d->rowM, d->rowR);
	} else if (type == 'v') {
		if (K == 0 && J == 0)
			FAIL("Source %s has both input nodes shorted to 0!", refdes);
		d->cl.units = 'V'; d->cb.units = 'V';
		d->rowR = rowFindOrAdd(s, 'i', refdes);
		nodeAdd(s, d->rowR, K); nodeAdd(s, d->rowR, J);
		nodeAdd(s, K, d->rowR); nodeAdd(s, J, d->rowR);
	} else {
		static const double icc = 0.0;   /* nonlinear_c.c:427 */
		d->cl.units = 'F';
		d->integ.units = 'F';
		itInitialize(&d->integ, &s->ctl, 0.0, &icc);
		nodeAdd(s, K, K); nodeAdd(s, J, K); nodeAdd(s, K, J); nodeAdd(s, J, J);
		d->rowR = rowFindOrAdd(s, 'i', refdes);
		d->rowM = rowFindOrAdd(s, 'v', refdes);   /* rowC */
		nodeAdd(s, d->rowR, d->rowM); nodeAdd(s, d->rowM, d->rowR);
	}
	clInitialize(&d->cl, 0.0);
	if (type != 'c') cbInitialize(&d->cb, &s->ctl, 0.0);
	ctx.s = s; ctx.d = d;
	d->calc = calcNew(equation, bGetVar, &ctx);
	if (d->calc == NULL) {
		/* the device was never appended, but release what it owns */
		free(d->refdes); free(d->bv); d->refdes = NULL; d->bv = NULL;
		FAIL("Bad B equation: %s", equation);
	}
	return devCommit(s, d);
}

int eoSetParam(eo_sim *s, const char *refdes, const char *name, double v)
{
	int i;
	for (i = 0; i < s->ndev; i++) {
		Dev *d = s->dev[i];
		if (strcmp(d->refdes, refdes)) continue;
		if (!strcmp(name, "R") || !strcmp(name, "C") || !strcmp(name, "L") ||
				!strcmp(name, "DC")) { d->val = v; return 0; }
		if (!strcmp(name, "Z0")) { d->Z0 = v; return 0; }
		if (!strcmp(name, "Td")) { d->Td = v; return 0; }
		if (!strcmp(name, "loss")) { d->loss = v; return 0; }
		if (!strcmp(name, "bypass")) { d->bypass = (v != 0.0); return 0; }
		if (name[0] == 'A' && name[1] >= '0' && name[1] <= '6' && !name[2] &&
				d->wave) { d->wave->arg[name[1] - '0'] = v; return 0; }
		FAIL("unknown parameter %s.%s", refdes, name);
	}
	FAIL("unknown device %s", refdes);
}

/* ------------------------------------------------------------------------ *
 * Behavioural helpers (shared by the three B flavours)
 * ------------------------------------------------------------------------ */
static int bSolve(eo_sim *s, Dev *d, int dvar, double *out)
{
	double vals[64];
	const Calc *c = d->calc;
	int i;
	if (c->nvars > 64) FAIL("too many variables in B equation");
	for (i = 0; i < c->nvars; i++)
		vals[i] = (c->varRef[i] == 0x40000000) ? s->ctl.time : xGet(s, c->varRef[i]);
	return calcRun(c, vals, dvar, s->ctl.gmin, out);
}

/* index into calc variables of the k-th row-variable (skipping "time") */
static int bCalcVarOf(const Dev *d, int k)
{
	int i, seen = 0;
	for (i = 0; i < d->calc->nvars; i++) {
		if (d->calc->varRef[i] == 0x40000000) continue;
		if (seen == k) return i;
		seen++;
	}
	return -1;
}

/* deviceNonlinearLoad(+LoadVariable): nonlinear_i.c:107-138, _v.c:85-116,
 * _c.c:92-123 */
static int bLoad(eo_sim *s, Dev *d)
{
	int k;
	d->BeqCalc = d->Ic;
	for (k = 0; k < d->nbv; k++) {
		double G = 0.0;
		BVar *v = &d->bv[k];
		TRY(bSolve(s, d, bCalcVarOf(d, k), &G));
		if (d->type == D_BI) {
			aPlus(s, d->pin[1], v->row, -(G - v->G));
			aPlus(s, d->rowM, v->row, (G - v->G));
		} else {
			aPlus(s, d->rowR, v->row, -(G - v->G));
		}
		v->G = G;
		d->BeqCalc -= G * xGet(s, v->row);
		if (isnan(d->BeqCalc)) FAIL("B source: NaN");
	}
	if (d->type == D_BI) {
		bPlus(s, d->pin[1], d->BeqCalc - d->Beq);
		bPlus(s, d->rowM, -(d->BeqCalc - d->Beq));
	} else {
		bPlus(s, d->rowR, d->BeqCalc - d->Beq);
	}
	d->Beq = d->BeqCalc;
	return 0;
}

/* deviceNonlinearCalculate: nonlinear_i.c:85-103 (with the In==0 shortcut),
 * :73-81 (initial, without it), _v.c:71-77, _c.c:79-85 */
static int bCalculate(eo_sim *s, Dev *d, int initial)
{
	if (d->type == D_BI) {
		d->In = xGet(s, d->rowR);
		if (isnan(d->In)) FAIL("B source: NaN");
		if (!initial && d->In == 0.0) { d->Ic = 0.0; return 0; }
	} else if (d->type == D_BV) {
		d->In = xGet(s, d->pin[0]) - xGet(s, d->pin[1]);
		if (isnan(d->In)) FAIL("B source: NaN");
	} else {
		d->In = xGet(s, d->rowM);
		if (isnan(d->In)) FAIL("B source: NaN");
	}
	return bSolve(s, d, -1, &d->Ic);
}

/* ------------------------------------------------------------------------ *
 * Device phases (core/device.c:45-160 dispatch, insertion order)
 * ------------------------------------------------------------------------ */
static int devLoad(eo_sim *s, Dev *d)
{
	const Control *c = &s->ctl;
	int K = d->pin[0], J = d->pin[1];
	double g;
	switch (d->type) {
	case D_R:                                   /* resistor.c:47-70 */
		g = 1 / d->val;
		aPlus(s, K, K, g); aPlus(s, K, J, -g); aPlus(s, J, K, -g); aPlus(s, J, J, g);
		break;
	case D_C:                                   /* capacitor.c:136-158 */
		d->G = 0; d->Ieq = 0;
		break;
	case D_L:                                   /* inductor.c:137-163 */
		d->G = 0; d->Ieq = 0;
		aSet(s, d->rowR, K, 1.0); aSet(s, d->rowR, J, -1.0);
		aSet(s, K, d->rowR, 1.0); aSet(s, J, d->rowR, -1.0);
		break;
	case D_V:                                   /* source_v.c:99-135 */
		cbInitialize(&d->cb, c, 0.0);
		if (d->wave) TRY(wfInitialize(d->wave, c));
		d->dc = d->wave ? wfDC(d->wave) : d->val;
		bPlus(s, d->rowR, d->dc);
		aSet(s, d->rowR, K, 1.0); aSet(s, d->rowR, J, -1.0);
		aSet(s, K, d->rowR, 1.0); aSet(s, J, d->rowR, -1.0);
		break;
	case D_I:                                   /* source_i.c:97-128 */
		cbInitialize(&d->cb, c, 0.0);
		if (d->wave) TRY(wfInitialize(d->wave, c));
		d->dc = d->wave ? wfDC(d->wave) : d->val;
		bPlus(s, K, -d->dc); bPlus(s, J, d->dc);
		break;
	case D_T: {                                 /* tline.c:208-270 */
		int L = d->pin[2], M = d->pin[3], R = d->rowR, S = d->rowS, i;
		cbInitialize(&d->cb, c, 0.0); cbInitialize(&d->cb2, c, 0.0);
		d->hint = 0;
		d->Vr = 0.0; d->Vs = 0.0;
		for (i = 0; i < 6; i++) d->ic[i] = 0.0;
		if (d->bypass) {
			/* masked section (not in the reference; mirrors simcu_device.cuh): the
			 * two ports are joined by an ideal 0 V branch, V_K - V_J = V_L - V_M with
			 * the branch current i(ref#in1) through both; i(ref#in2) = 0 */
			aSet(s, R, L, -1.0); aSet(s, R, M, 1.0); aSet(s, S, K, 0.0); aSet(s, S, J, 0.0);
			aSet(s, K, S, 0.0); aSet(s, J, S, 0.0); aSet(s, L, R, -1.0); aSet(s, M, R, 1.0);
			aSet(s, R, K, 1.0); aSet(s, R, J, -1.0); aSet(s, K, R, 1.0); aSet(s, J, R, -1.0);
			aSet(s, S, L, 0.0); aSet(s, S, M, 0.0); aSet(s, L, S, 0.0); aSet(s, M, S, 0.0);
			aPlus(s, S, S, 1.0);
			break;
		}
		aSet(s, R, L, -1.0); aSet(s, R, M, 1.0); aSet(s, S, K, -1.0); aSet(s, S, J, 1.0);
		aSet(s, K, S, 1.0); aSet(s, J, S, -1.0); aSet(s, L, R, -1.0); aSet(s, M, R, 1.0);
		aSet(s, R, K, 1.0); aSet(s, R, J, -1.0); aSet(s, K, R, 1.0); aSet(s, J, R, -1.0);
		aSet(s, S, L, 1.0); aSet(s, S, M, -1.0); aSet(s, L, S, 1.0); aSet(s, M, S, -1.0);
		aPlus(s, R, R, -d->Z0); aPlus(s, S, S, -d->Z0);
		break;
	}
	case D_VI: {                                /* vicurve.c:167-221 */
		int R = d->rowR, M = d->rowM;
		clInitialize(&d->cl, 0.0);
		cbInitialize(&d->cb, c, 0.0);
		if (d->ta) {
			TRY(pwInitialize(d->ta));
			TRY(pwCalcValue(d->ta, &d->taIndex, 0.0, &d->a, &d->G));
		} else d->a = 1.0;
		TRY(pwInitialize(d->vi));
		TRY(pwCalcValue(d->vi, &d->viIndex, 0.0, &d->Ieq, &d->G));
		d->G *= d->a; d->Ieq *= d->a;
		aSet(s, R, K, 1.0); aSet(s, R, M, -1.0); aSet(s, K, R, 1.0); aSet(s, M, R, -1.0);
		aPlus(s, J, J, d->G); aPlus(s, J, M, -d->G);
		aPlus(s, M, J, -d->G); aPlus(s, M, M, d->G);
		bPlus(s, J, d->Ieq); bPlus(s, M, -d->Ieq);
		break;
	}
	case D_BI:                                  /* nonlinear_i.c:246-290 */
		clInitialize(&d->cl, 0.0); cbInitialize(&d->cb, c, 0.0);
		d->Ic = d->In = d->Beq = d->BeqCalc = 0.0;
		{ int k; for (k = 0; k < d->nbv; k++) d->bv[k].G = 0.0; }
		aSet(s, d->rowR, K, 1.0); aSet(s, d->rowR, d->rowM, -1.0);
		aSet(s, K, d->rowR, 1.0); aSet(s, d->rowM, d->rowR, -1.0);
		TRY(bCalculate(s, d, 1)); TRY(bLoad(s, d));
		break;
	case D_BV:                                  /* nonlinear_v.c:216-256 */
		clInitialize(&d->cl, 0.0); cbInitialize(&d->cb, c, 0.0);
		d->Ic = d->In = d->Beq = d->BeqCalc = 0.0;
		{ int k; for (k = 0; k < d->nbv; k++) d->bv[k].G = 0.0; }
		aSet(s, d->rowR, K, 1.0); aSet(s, d->rowR, J, -1.0);
		aSet(s, K, d->rowR, 1.0); aSet(s, J, d->rowR, -1.0);
		TRY(bCalculate(s, d, 1)); TRY(bLoad(s, d));
		break;
	case D_BC:                                  /* nonlinear_c.c:303-349 */
		clInitialize(&d->cl, 0.0);
		d->Ic = d->In = d->Beq = d->BeqCalc = 0.0;
		{ int k; for (k = 0; k < d->nbv; k++) d->bv[k].G = 0.0; }
		aSet(s, d->rowR, d->rowM, 1.0); aSet(s, d->rowM, d->rowR, 1.0);
		TRY(bCalculate(s, d, 1)); TRY(bLoad(s, d));
		break;
	}
	return 0;
}

static int devLinearize(eo_sim *s, Dev *d, int *linear)
{
	const Control *c = &s->ctl;
	if (d->type == D_VI) {                      /* vicurve.c:110-163 */
		int J = d->pin[1], M = d->rowM;
		double i0 = xGet(s, d->rowR), v0, ic, gc;
		if (isnan(i0)) FAIL("VI: NaN");
		v0 = xGet(s, d->pin[0]) - xGet(s, J);
		if (isnan(v0)) FAIL("VI: NaN");
		if (i0 == 0.0) { ic = 0.0; gc = 0.0; }
		else {
			TRY(pwCalcValue(d->vi, &d->viIndex, v0, &ic, &gc));
			ic *= d->a; gc *= d->a;
		}
		*linear = clIsLinear(&d->cl, c, i0, ic);
		if (!*linear) {
			double G, Ieq;
			if (gc < c->gmin) gc = c->gmin;
			Ieq = (ic - gc * v0); G = gc;
			aPlus(s, J, J, G - d->G); aPlus(s, J, M, -(G - d->G));
			aPlus(s, M, J, -(G - d->G)); aPlus(s, M, M, G - d->G);
			d->G = G;
			bPlus(s, J, Ieq - d->Ieq); bPlus(s, M, -(Ieq - d->Ieq));
			d->Ieq = Ieq;
		}
		return 0;
	}
	if (d->type == D_BI || d->type == D_BV || d->type == D_BC) {
		/* nonlinear_i.c:224-242, _v.c:194-212, _c.c:188-206 */
		TRY(bCalculate(s, d, 0));
		*linear = clIsLinear(&d->cl, c, d->In, d->Ic);
		if (!*linear) TRY(bLoad(s, d));
	}
	return 0;
}

static int devInitStep(eo_sim *s, Dev *d)
{
	const Control *c = &s->ctl;
	switch (d->type) {
	case D_C:                                   /* capacitor.c:112-132 */
		return itInitialize(&d->integ, c, xGet(s, d->pin[0]) - xGet(s, d->pin[1]),
				&d->val);
	case D_L:                                   /* inductor.c:113-133 */
		return itInitialize(&d->integ, c, xGet(s, d->rowR), &d->val);
	case D_BC:                                  /* nonlinear_c.c:281-299 */
		TRY(itInitialize(&d->integ, c, xGet(s, d->pin[0]) - xGet(s, d->pin[1]),
				&d->Ic));
		d->G = 0; d->Ieq = 0;
		return 0;
	case D_T: {                                 /* tline.c:93-122 */
		int K = d->pin[0], J = d->pin[1], L = d->pin[2], M = d->pin[3];
		int R = d->rowR, S = d->rowS;
		if (d->bypass) return 0;
		aSet(s, R, L, 0.0); aSet(s, R, M, 0.0); aSet(s, S, K, 0.0); aSet(s, S, J, 0.0);
		aSet(s, K, S, 0.0); aSet(s, J, S, 0.0); aSet(s, L, R, 0.0); aSet(s, M, R, 0.0);
		d->ic[0] = 0.0; d->ic[1] = 0.0;
		d->ic[2] = xGet(s, K); d->ic[3] = xGet(s, J);
		d->ic[4] = xGet(s, L); d->ic[5] = xGet(s, M);
		return 0;
	}
	default: return 0;
	}
}

/* history lookup: history_interp.c:42-92, history.c:36-59 */
static int histLocate(eo_sim *s, Dev *d, double tp, int *iP, int *iN)
{
	int i = d->hint;
	if (i < 0 || i >= s->npts) i = 0;
	while (s->htime[i] > tp && i > 0) i--;
	while (i + 1 < s->npts && s->htime[i + 1] <= tp) i++;
	d->hint = i;
	if (i + 1 < s->npts) { *iP = i; *iN = i + 1; }
	else {
		if (i - 1 < 0) FAIL("T-line: history too short to extrapolate");
		*iP = i - 1; *iN = i;
	}
	return 0;
}
static double histInterp(const eo_sim *s, int iP, int iN, int row, double tp)
{
	double tP, tN, xP, xN;
	if (row == 0) return 0;
	tP = s->htime[iP]; tN = s->htime[iN];
	xP = s->hX[(size_t)iP * s->N + row - 1]; xN = s->hX[(size_t)iN * s->N + row - 1];
	return ((xN - xP) / (tN - tP)) * (tp - tP) + xP;
}

static int devStep(eo_sim *s, Dev *d, int *brk)
{
	const Control *c = &s->ctl;
	double dc;
	switch (d->type) {
	case D_V:                                   /* source_v.c:72-95 */
		if (!d->wave) return 0;
		TRY(wfCalcValue(d->wave, c, &dc));
		bPlus(s, d->rowR, dc - d->dc);
		d->dc = dc;
		*brk = cbIsBreak(&d->cb, c, dc);
		return 0;
	case D_I:                                   /* source_i.c:69-93 */
		if (!d->wave) return 0;
		TRY(wfCalcValue(d->wave, c, &dc));
		bPlus(s, d->pin[0], -(dc - d->dc));
		bPlus(s, d->pin[1], dc - d->dc);
		d->dc = dc;
		*brk = cbIsBreak(&d->cb, c, dc);
		return 0;
	case D_VI: {                                /* vicurve.c:85-106 */
		double dadt;
		if (!d->ta) return 0;
		TRY(pwCalcValue(d->ta, &d->taIndex, c->time, &d->a, &dadt));
		*brk = cbIsBreak(&d->cb, c, d->a);
		return 0;
	}
	case D_T: {                                 /* tline.c:126-204 */
		double tp = c->time - d->Td, Ir, Is, Vk, Vj, Vl, Vm, Vr, Vs;
		if (d->bypass) return 0;
		if (isnan(tp)) FAIL("T-line: NaN");
		if (tp <= 0) {
			Ir = d->ic[0]; Is = d->ic[1]; Vk = d->ic[2]; Vj = d->ic[3];
			Vl = d->ic[4]; Vm = d->ic[5];
		} else {
			int iP, iN;
			TRY(histLocate(s, d, tp, &iP, &iN));
			Ir = histInterp(s, iP, iN, d->rowR, tp);
			Is = histInterp(s, iP, iN, d->rowS, tp);
			Vk = histInterp(s, iP, iN, d->pin[0], tp);
			Vj = histInterp(s, iP, iN, d->pin[1], tp);
			Vl = histInterp(s, iP, iN, d->pin[2], tp);
			Vm = histInterp(s, iP, iN, d->pin[3], tp);
		}
		if (d->loss == HUGE_VAL) {
			Vr = (Vl - Vm) + (d->Z0) * Is;
			Vs = (Vk - Vj) + (d->Z0) * Ir;
		} else {
			Vr = exp(-(d->loss) / 2) * ((Vl - Vm) + (d->Z0) * Is);
			Vs = exp(-(d->loss) / 2) * ((Vk - Vj) + (d->Z0) * Ir);
		}
		bPlus(s, d->rowR, Vr - d->Vr); d->Vr = Vr;
		bPlus(s, d->rowS, Vs - d->Vs); d->Vs = Vs;
		*brk = cbIsBreak(&d->cb, c, Vr);
		if (!*brk) *brk = cbIsBreak(&d->cb2, c, Vs);
		return 0;
	}
	case D_BI: case D_BV:                       /* nonlinear_i.c:198-218, _v.c:166-188 */
		if (!d->usesTime) return 0;
		TRY(bCalculate(s, d, 1)); TRY(bLoad(s, d));
		*brk = cbIsBreak(&d->cb, c, d->Beq);
		return 0;
	case D_BC:                                  /* nonlinear_c.c:172-184 */
		if (!d->usesTime) return 0;
		TRY(bCalculate(s, d, 1)); TRY(bLoad(s, d));
		return 0;
	default: return 0;
	}
}

static int devIntegrate(eo_sim *s, Dev *d)
{
	const Control *c = &s->ctl;
	double G, Ieq, x0;
	int K = d->pin[0], J = d->pin[1];
	switch (d->type) {
	case D_C: case D_BC:                        /* capacitor.c:76-108, nonlinear_c.c:232-277 */
		x0 = xGet(s, K) - xGet(s, J);
		TRY(itIntegrate(&d->integ, c, x0, &G, &Ieq));
		aPlus(s, K, K, G - d->G); aPlus(s, K, J, -(G - d->G));
		aPlus(s, J, K, -(G - d->G)); aPlus(s, J, J, G - d->G);
		d->G = G;
		bPlus(s, K, Ieq - d->Ieq); bPlus(s, J, -(Ieq - d->Ieq));
		d->Ieq = Ieq;
		return 0;
	case D_L:                                   /* inductor.c:78-109 */
		x0 = xGet(s, d->rowR);
		TRY(itIntegrate(&d->integ, c, x0, &G, &Ieq));
		aPlus(s, d->rowR, d->rowR, -(G - d->G));
		d->G = G;
		bPlus(s, d->rowR, -(Ieq - d->Ieq));
		d->Ieq = Ieq;
		return 0;
	default: return 0;
	}
}

static int devMinStep(eo_sim *s, Dev *d, double *minStep, int *has)
{
	const Control *c = &s->ctl;
	*has = 1;
	switch (d->type) {
	case D_C: case D_BC:                        /* capacitor.c:54-72, nonlinear_c.c:210-228 */
		return itNextStep(&d->integ, c, xGet(s, d->pin[0]) - xGet(s, d->pin[1]), minStep);
	case D_L:                                   /* inductor.c:56-74 */
		return itNextStep(&d->integ, c, xGet(s, d->rowR), minStep);
	default: *has = 0; return 0;
	}
}

static int devNextStep(eo_sim *s, Dev *d, double *nextStep, int *has)
{
	const Control *c = &s->ctl;
	*has = 1;
	switch (d->type) {
	case D_V: case D_I:                         /* source_v.c:54-68 */
		if (d->wave) wfNextStep(d->wave, c, nextStep);
		return 0;
	case D_VI:                                  /* vicurve.c:67-81 */
		if (d->ta) {
			double nextTime;
			pwGetNextX(d->ta, &d->taIndex, c->time, &nextTime);
			*nextStep = nextTime - c->time;
		}
		return 0;
	default: *has = 0; return 0;
	}
}

/* ------------------------------------------------------------------------ *
 * Linear solve.  Restates the NUMERIC behaviour of dgssvx as configured in
 * core/matrix.c:240-249 (no equilibration, no refinement, partial pivoting
 * with u = 1.0, libs/superlu/SRC/dpivotL.c:99-147: largest magnitude in the
 * column, first one wins ties, zero column = "U is singular").  Dense
 * right-looking elimination, pivots stored as reciprocals
 * (dpivotL.c:170-172 scales the column by 1/pivot).  The column order is the
 * natural one unless a fixed (column,row) sequence is replayed; SuperLU's
 * COLAMD order only changes rounding.
 * ------------------------------------------------------------------------ */
static int luSolve(eo_sim *s)
{
	int N = s->N, k, r, cidx;
	const int *fpc = s->fixPc[s->phase], *fpr = s->fixPr[s->phase];
	double *W = s->W, *y = s->y;

	memcpy(W, s->A, sizeof(double) * N * N);
	memcpy(y, s->B, sizeof(double) * N);
	memset(s->rdone, 0, sizeof(int) * N);
	memset(s->cdone, 0, sizeof(int) * N);

	for (k = 0; k < N; k++) {
		int col = fpc ? fpc[k] : k, prow = -1;
		double inv;
		if (fpr) prow = fpr[k];
		else {
			double pivmax = 0.0;
			for (r = 0; r < N; r++) {
				double v;
				if (s->rdone[r]) continue;
				v = fabs(W[r * N + col]);
				if (v > pivmax) { pivmax = v; prow = r; }
			}
			if (prow < 0 || pivmax == 0.0) FAIL("SuperLU: U is singular");
		}
		if (W[prow * N + col] == 0.0) FAIL("SuperLU: U is singular");
		s->pcol[k] = col; s->prow[k] = prow;
		s->rdone[prow] = 1; s->cdone[col] = 1;
		inv = 1.0 / W[prow * N + col];
		s->pinv[k] = inv;
		for (r = 0; r < N; r++) {
			double l;
			if (s->rdone[r]) continue;
			l = W[r * N + col] * inv;
			W[r * N + col] = l;
			if (l == 0.0) continue;
			for (cidx = 0; cidx < N; cidx++) {
				if (s->cdone[cidx]) continue;
				W[r * N + cidx] -= l * W[prow * N + cidx];
			}
			y[r] -= l * y[prow];
		}
	}
	for (k = N - 1; k >= 0; k--) {
		double acc = y[s->prow[k]];
		int j;
		for (j = k + 1; j < N; j++)
			acc -= W[s->prow[k] * N + s->pcol[j]] * s->X[s->pcol[j]];
		s->X[s->pcol[k]] = acc * s->pinv[k];
	}
	if (s->phase) s->st.factorizations++; else s->st.op_factorizations++;
	if (!s->firstSeen[s->phase]) {       /* eoGetPivots */
		if (s->firstPc[s->phase] == NULL) {
			s->firstPc[s->phase] = malloc(sizeof(int) * (N + 1));
			s->firstPr[s->phase] = malloc(sizeof(int) * (N + 1));
		}
		memcpy(s->firstPc[s->phase], s->pcol, sizeof(int) * N);
		memcpy(s->firstPr[s->phase], s->prow, sizeof(int) * N);
		s->firstSeen[s->phase] = 1;
	}
	return 0;
}

int eoSetPivots(eo_sim *s, int phase, const int *pc, const int *pr, int n)
{
	if (phase < 0 || phase > 1) return -1;
	free(s->fixPc[phase]); free(s->fixPr[phase]);
	s->fixPc[phase] = s->fixPr[phase] = NULL;
	if (pc == NULL || pr == NULL) return 0;
	s->fixPc[phase] = malloc(sizeof(int) * n);
	s->fixPr[phase] = malloc(sizeof(int) * n);
	memcpy(s->fixPc[phase], pc, sizeof(int) * n);
	memcpy(s->fixPr[phase], pr, sizeof(int) * n);
	return 0;
}

/* simulatorSolve: core/simulator.c:45-69.  Returns the reference's count. */
/* stamp watch: first = 1 remembers the stamps of the loop's first solve, later
 * calls mark what changed since the previous solve (test instrumentation: the
 * engine solves the matrix blocks nothing here can reach only once per loop) */
static void watchStamps(eo_sim *s, int first)
{
	int N = s->N, i;
	if (!s->watchOn) return;
	if (s->watchA == NULL) {
		s->watchA = malloc(sizeof(double) * N * N + 8); s->watchB = malloc(sizeof(double) * N + 8);
		for (i = 0; i < 2; i++) {
			s->changedA[i] = calloc((size_t)N * N + 1, 1); s->changedB[i] = calloc(N + 1, 1);
		}
	}
	if (!first) {
		for (i = 0; i < N * N; i++)
			if (memcmp(&s->A[i], &s->watchA[i], sizeof(double))) s->changedA[s->phase][i] = 1;
		for (i = 0; i < N; i++)
			if (memcmp(&s->B[i], &s->watchB[i], sizeof(double))) s->changedB[s->phase][i] = 1;
	}
	memcpy(s->watchA, s->A, sizeof(double) * N * N);
	memcpy(s->watchB, s->B, sizeof(double) * N);
}

static int simSolve(eo_sim *s, int limit)
{
	int count = 0, linear, i;
	if (limit < 1) return -1;
	watchStamps(s, 1);
	TRY(luSolve(s));
	while (count++ < limit) {
		int mask = 0, bit = 0;
		linear = 1;
		for (i = 0; i < s->ndev; i++) {
			int loc = 1;
			Dev *d = s->dev[i];
			if (d->type == D_VI || d->type == D_BI || d->type == D_BV ||
					d->type == D_BC) {
				TRY(devLinearize(s, d, &loc));
				linear = linear && loc;
				if (bit < 31) mask |= g_clPassed << bit;
				bit++;
			}
		}
		if (s->logOn) {
			if (s->linN == s->linCap) {
				s->linCap = s->linCap ? 2 * s->linCap : 8192;
				s->linMask = realloc(s->linMask, sizeof(int) * s->linCap);
			}
			s->linMask[s->linN++] = mask;
		}
		if (linear) break;
		watchStamps(s, 0);
		TRY(luSolve(s));
	}
	return count;
}

/* ------------------------------------------------------------------------ *
 * matrix init / record / recall / results: core/matrix.c
 * ------------------------------------------------------------------------ */
static int nodeCmp(const void *a, const void *b)
{
	const NodeIdx *x = a, *y = b;
	if (x->col != y->col) return x->col - y->col;
	return x->row - y->row;
}

static int matInitialize(eo_sim *s)
{
	int N = s->nrows - 1;
	s->N = N;
	qsort(s->nodes, s->nnodes, sizeof(NodeIdx), nodeCmp);
	s->A = calloc((size_t)N * N + 1, sizeof(double));
	s->W = calloc((size_t)N * N + 1, sizeof(double));
	s->B = calloc(N + 1, sizeof(double)); s->X = calloc(N + 1, sizeof(double));
	s->y = calloc(N + 1, sizeof(double)); s->pinv = calloc(N + 1, sizeof(double));
	s->rdone = calloc(N + 1, sizeof(int)); s->cdone = calloc(N + 1, sizeof(int));
	s->prow = calloc(N + 1, sizeof(int)); s->pcol = calloc(N + 1, sizeof(int));
	return 0;
}

static void matClear(eo_sim *s)
{
	memset(s->A, 0, sizeof(double) * s->N * s->N);
	memset(s->X, 0, sizeof(double) * s->N);
	memset(s->B, 0, sizeof(double) * s->N);
	s->npts = 0;
}

static void matRecord(eo_sim *s, double time)
{
	if (s->npts == s->hcap) {
		s->hcap = s->hcap ? 2 * s->hcap : 1024;
		s->htime = realloc(s->htime, sizeof(double) * s->hcap);
		s->hX = realloc(s->hX, sizeof(double) * (size_t)s->hcap * (s->N + 1));
	}
	s->htime[s->npts] = time;
	memcpy(&s->hX[(size_t)s->npts * s->N], s->X, sizeof(double) * s->N);
	s->npts++;
}

static void matRecall(eo_sim *s)
{
	memcpy(s->X, &s->hX[(size_t)(s->npts - 1) * s->N], sizeof(double) * s->N);
}

/* matrix.c:430-470: [numPoints][numVariables], column 0 = time, numVariables
 * counts the time column (the reference counts the ground row instead) */
static int matGetSolution(eo_sim *s, double **data, char ***variables,
		int *numPoints, int *numVariables)
{
	int i, nv = s->N + 1;
	*numVariables = nv;
	*numPoints = s->npts;
	*data = malloc(sizeof(double) * (size_t)s->npts * nv + 8);
	for (i = 0; i < s->npts; i++) {
		(*data)[(size_t)i * nv] = s->htime[i];
		memcpy(&(*data)[(size_t)i * nv + 1], &s->hX[(size_t)i * s->N],
				sizeof(double) * s->N);
	}
	if (variables) {
		*variables = malloc(sizeof(char *) * (nv + 1));
		(*variables)[0] = "time";
		for (i = 1; i < s->nrows; i++) (*variables)[i] = s->rowName[i];
		(*variables)[nv] = NULL;
	}
	return 0;
}

int eoNumUnknowns(eo_sim *s) { return s->nrows - 1; }
int eoNumNonzeros(eo_sim *s) { return s->nnodes; }
int eoGetPattern(eo_sim *s, int *rowIdx, int *colStart)
{
	int i, N = s->nrows - 1;
	qsort(s->nodes, s->nnodes, sizeof(NodeIdx), nodeCmp);
	for (i = 0; i <= N; i++) colStart[i] = 0;
	for (i = 0; i < s->nnodes; i++) {
		rowIdx[i] = s->nodes[i].row - 1;
		colStart[s->nodes[i].col] = i + 1;
	}
	for (i = 1; i <= N; i++)
		if (colStart[i] < colStart[i - 1]) colStart[i] = colStart[i - 1];
	return 0;
}
void eoGetStats(eo_sim *s, eo_stats *st) { *st = s->st; }

/* ------------------------------------------------------------------------ *
 * Analyses: core/simulator.c:75-282
 * ------------------------------------------------------------------------ */
#define Min3(x, y, z) ((x < y) ? ((z < x) ? z : x) : ((z < y) ? z : y))

static int opPrologue(eo_sim *s)
{
	int i, linCount;
	if (!s->locked) { TRY(matInitialize(s)); s->locked = -1; }
	matClear(s);
	s->phase = 0;
	for (i = 0; i < s->ndev; i++) TRY(devLoad(s, s->dev[i]));
	linCount = simSolve(s, s->ctl.itl1);
	if ((linCount < 0) || (linCount > s->ctl.itl1))
		FAIL("operating point failed to converge");
	return 0;
}

int eoRunOperatingPoint(eo_sim *s, double **data, char ***variables,
		int *numPoints, int *numVariables)
{
	if (s == NULL) return -1;
	TRY(opPrologue(s));
	matRecord(s, 0.0);
	return matGetSolution(s, data, variables, numPoints, numVariables);
}

/* The attempt log: what core/simulator.c:156-218 decided for every attempt of
 * the last transient - the time it tried, the integration order it used (after
 * the break-point reset, :172), how many LU solves its Newton loop took and
 * whether it was accepted (0), rejected by the LTE test (1, :191-198) or for
 * non-convergence (2, :201-214).  The GPU engine replays it (option "replay")
 * so that both sides take the same steps and can be compared value by value. */
static void logAttempt(eo_sim *s, double time, int order, int outcome, int solves)
{
	if (!s->logOn) return;
	if (s->logN == s->logCap) {
		s->logCap = s->logCap ? 2 * s->logCap : 4096;
		s->logTime = realloc(s->logTime, sizeof(double) * s->logCap);
		s->logFlag = realloc(s->logFlag, sizeof(int) * s->logCap);
	}
	s->logTime[s->logN] = time;
	s->logFlag[s->logN] = EO_LOG_FLAG(order, outcome, solves);
	s->logN++;
}

void eoSetAttemptLog(eo_sim *s, int on) { s->logOn = on; s->logN = 0; s->linN = 0; }

/* the (column, row) sequence of the first factorisation of a phase in the runs so
 * far (partial pivoting unless a fixed sequence is being replayed) */
int eoGetPivots(eo_sim *s, int phase, int *pc, int *pr)
{
	if (phase < 0 || phase > 1 || !s->firstSeen[phase]) return -1;
	memcpy(pc, s->firstPc[phase], sizeof(int) * s->N);
	memcpy(pr, s->firstPr[phase], sizeof(int) * s->N);
	return s->N;
}

void eoSetStampWatch(eo_sim *s, int on) { s->watchOn = on; }

int eoGetStampWatch(eo_sim *s, int phase, const char **a, const char **b)
{
	if (phase < 0 || phase > 1 || s->changedA[0] == NULL) return 0;
	if (a) *a = s->changedA[phase];
	if (b) *b = s->changedB[phase];
	return s->N;
}

int eoGetLinearizeLog(eo_sim *s, const int **masks)
{
	if (masks) *masks = s->linMask;
	return s->linN;
}

int eoGetAttemptLog(eo_sim *s, const double **times, const int **flags)
{
	if (times) *times = s->logTime;
	if (flags) *flags = s->logFlag;
	return s->logN;
}

int eoRunTransient(eo_sim *s, double tstep, double tstop, double tmax,
		int restart, double **data, char ***variables, int *numPoints,
		int *numVariables)
{
	Control *c;
	int linCount = 0, breakPoint = 0, i, usedOrder;
	long solves0;
	double maxStep, thisStep = 0.0, lteStep = 0.0, prevTime = 0.0, oldMinstep;

	if (s == NULL) return -1;
	c = &s->ctl;
	memset(&s->st, 0, sizeof s->st);
	s->logN = 0; s->linN = 0;

	if (tmax == 0.0) {                          /* :91-94 */
		tmax = tstop / 50;
		tmax = (tstep < tmax) ? tstep : tmax;
	}
	c->tstep = tstep; c->tstop = tstop;
	oldMinstep = c->minstep;
	if (c->minstep < 0) c->minstep = tstep * 5e-5;
	maxStep = ((tstep < tmax) ? tstep : tmax) / 100;   /* :109 */

	if (restart || (c->time == 0.0)) {
		TRY(opPrologue(s));
		for (i = 0; i < s->ndev; i++) TRY(devInitStep(s, s->dev[i]));
		c->integratorOrder = c->maxorder;
		c->time = 0.0;
		prevTime = 0.0;
	} else {
		prevTime = c->time;
	}
	s->phase = 1;

	while (c->time < tstop) {
		matRecord(s, c->time);                  /* :144 */
		thisStep = maxStep;                     /* :152-154, device.c:128-147 */
		for (i = 0; i < s->ndev; i++) {
			double loc = 0.0; int has;
			TRY(devNextStep(s, s->dev[i], &loc, &has));
			if (has && (loc > c->minstep) && (loc < thisStep)) thisStep = loc;
		}
		while (1) {
			c->time = prevTime + thisStep;      /* :157-161 */
			if (c->time > tstop) { thisStep = tstop - prevTime; c->time = tstop; }
			breakPoint = 0;                     /* :167-173 */
			for (i = 0; i < s->ndev; i++) {
				int loc = 0;
				TRY(devStep(s, s->dev[i], &loc));
				breakPoint = breakPoint || loc;
			}
			if (breakPoint) c->integratorOrder = 1;
			usedOrder = c->integratorOrder;
			solves0 = s->st.factorizations;
			for (i = 0; i < s->ndev; i++) TRY(devIntegrate(s, s->dev[i]));
			linCount = simSolve(s, c->itl4);    /* :182 */
			if (linCount < 0) return -1;
			if (linCount < c->itl4) {
				lteStep = tmax;                 /* :188-200, device.c:106-124 */
				for (i = 0; i < s->ndev; i++) {
					double loc = 0.0; int has;
					TRY(devMinStep(s, s->dev[i], &loc, &has));
					if (has && (loc > 0.0) && (loc < lteStep)) lteStep = loc;
				}
				if (lteStep < (0.9 * thisStep)) {
					logAttempt(s, c->time, usedOrder, 1, (int)(s->st.factorizations - solves0));
					if (lteStep < (tstep * 1e-9))
						FAIL("Timestep %es is too Small at %es", lteStep, c->time);
					c->integratorOrder = 1;
					thisStep = lteStep;
					s->st.rejected_lte++;
				} else {
					logAttempt(s, c->time, usedOrder, 0, (int)(s->st.factorizations - solves0));
					break;
				}
			} else {                            /* :201-214 */
				logAttempt(s, c->time, usedOrder, 2, (int)(s->st.factorizations - solves0));
				thisStep = thisStep / 8;
				if (thisStep < (tstep * 1e-9))
					FAIL("Timestep %es is too Small at %es", thisStep, c->time);
				c->integratorOrder = (c->integratorOrder > 1)
					? c->integratorOrder - 1 : c->integratorOrder;
				maxStep = maxStep / 8;
				s->st.rejected_nc++;
			}
			matRecall(s);                       /* :217 */
		}
		c->integratorOrder = (c->integratorOrder < c->maxorder)   /* :222 */
			? c->integratorOrder + 1 : c->integratorOrder;
		prevTime = c->time;
		if (breakPoint) maxStep = Min3(lteStep, 0.1 * maxStep, tmax);
		else maxStep = Min3(lteStep, 2 * maxStep, tmax);
		s->st.accepted++;
	}
	matRecord(s, c->time);                      /* :238 */
	c->minstep = oldMinstep;
	return matGetSolution(s, data, variables, numPoints, numVariables);
}
