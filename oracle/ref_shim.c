/*
 * Link-time shim for building the UNMODIFIED reference engine without a
 * Fortran toolchain (test infrastructure only - never linked into the product).
 *
 * The reference's SuperLU build expects a handful of Fortran BLAS-1 / LAPACK
 * symbols (libs/superlu/Makefile:17 sets -DAdd_, so names carry a trailing
 * underscore).  None of them is reached on the transient hot path
 * (matrix.c:240-249 switches off equilibration, refinement, condition number
 * and pivot growth); they are only referenced by code that is linked in.
 * The W-element / complex LAPACK symbols (math/netlib.c:31-58) are reachable
 * only from tline_w.c, which is out of scope; they abort loudly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

int lsame_(char *a, char *b)
{
	int ca = *a, cb = *b;
	if (ca >= 'a' && ca <= 'z') ca -= 32;
	if (cb >= 'a' && cb <= 'z') cb -= 32;
	return ca == cb;
}

int xerbla_(char *name, int *info)
{
	fprintf(stderr, "ref_shim: xerbla %.6s info=%d\n", name, *info);
	return 0;
}

double dasum_(int *n, double *x, int *incx)
{
	double s = 0.0;
	int i;
	for (i = 0; i < *n; i++) s += fabs(x[i * (*incx)]);
	return s;
}

int dcopy_(int *n, double *x, int *incx, double *y, int *incy)
{
	int i;
	for (i = 0; i < *n; i++) y[i * (*incy)] = x[i * (*incx)];
	return 0;
}

int daxpy_(int *n, double *a, double *x, int *incx, double *y, int *incy)
{
	int i;
	for (i = 0; i < *n; i++) y[i * (*incy)] += (*a) * x[i * (*incx)];
	return 0;
}

int idamax_(int *n, double *x, int *incx)
{
	int i, best = 0;
	double m = -1.0;
	for (i = 0; i < *n; i++) {
		double v = fabs(x[i * (*incx)]);
		if (v > m) { m = v; best = i; }
	}
	return (*n > 0) ? best + 1 : 0;
}

#define ABORT_STUB(name) \
	void name(void) { fprintf(stderr, "ref_shim: " #name " is not available " \
		"(no Fortran toolchain; W-element path is out of scope)\n"); abort(); }

ABORT_STUB(dtrsv_)
ABORT_STUB(dgemv_)
ABORT_STUB(dgemm_)
ABORT_STUB(dtrsm_)
ABORT_STUB(dgesv_)
ABORT_STUB(dgetrf_)
ABORT_STUB(dgetri_)
ABORT_STUB(rpoly_)
ABORT_STUB(zgees_)
ABORT_STUB(zgemm_)
ABORT_STUB(zgesv_)

/* The reference expects its host application to define the log streams
 * (libs/log/log.h:137-144). NULL means stderr / stdout. */
#ifdef REF_SHIM_NO_LOG_STREAMS   /* the Python binding defines them itself */
extern FILE *fderr;
extern FILE *fdlog;
#else
FILE *fderr = NULL;
FILE *fdlog = NULL;
#endif

/* Test helper: silence (1) or restore (0) the reference's log output. */
void ref_shim_quiet(int quiet)
{
	static FILE *devnull = NULL;
	if (quiet) {
		if (devnull == NULL) devnull = fopen("/dev/null", "w");
		fderr = devnull;
		fdlog = devnull;
	} else {
		fderr = NULL;
		fdlog = NULL;
	}
}

/* Test helper for differential fuzzing of the expression engine: evaluates
 * `expr` with the reference's calculon (libs/calculon/calc.h:33-41) at the
 * given variable values.  value <- calcSolve, derivs[i] <- calcDiff(names[i]).
 * Returns 0 on success, -1 on a parse / evaluation error. Variables that do not
 * occur in the expression get derivative NaN. */
typedef struct _calc calc_;
typedef double **(*calcGetVarPtr_)(char *varName, void *private);
int calcSolve(calc_ *r, double *solution);
int calcDiff(calc_ *r, char *variable, double *solution);
int calcDestroy(calc_ **r);
calc_ *calcNew(calc_ *r, char *buffer, calcGetVarPtr_ getVarFunc,
		void *private, double *minDiv);

#include <string.h>

typedef struct {
	int n;
	const char **names;
	double **slots;   /* slots[i] points at the i-th value */
	int *used;
} shimVars_;

static double **shimGetVar(char *name, void *private)
{
	shimVars_ *v = (shimVars_ *)private;
	int i;
	for (i = 0; i < v->n; i++) {
		if (!strcmp(name, v->names[i])) {
			v->used[i] = 1;
			return &v->slots[i];
		}
	}
	return NULL;
}

int ref_calc_eval(const char *expr, int nvars, const char **names,
		const double *values, double minDiv, double *value, double *derivs)
{
	shimVars_ v;
	calc_ *calc = NULL;
	double *vals, md = minDiv;
	char *buf;
	int i, rc = 0;

	vals = malloc(sizeof(double) * (nvars + 1));
	v.n = nvars;
	v.names = names;
	v.slots = malloc(sizeof(double *) * (nvars + 1));
	v.used = calloc(nvars + 1, sizeof(int));
	for (i = 0; i < nvars; i++) { vals[i] = values[i]; v.slots[i] = &vals[i]; }
	buf = malloc(strlen(expr) + 1);
	strcpy(buf, expr);

	calc = calcNew(NULL, buf, shimGetVar, &v, &md);
	if (calc == NULL) { rc = -1; goto done; }
	/* the reference checks isnan(*solution) after every token, so the
	 * caller-side initial value matters (the devices start from 0.0) */
	*value = 0.0;
	if (calcSolve(calc, value)) { rc = -1; goto done; }
	for (i = 0; i < nvars; i++) {
		derivs[i] = NAN;
		if (v.used[i]) {
			derivs[i] = 0.0;
			char *nm = malloc(strlen(names[i]) + 1);
			strcpy(nm, names[i]);
			if (calcDiff(calc, nm, &derivs[i])) rc = -1;
			free(nm);
		}
	}
done:
	if (calc != NULL) calcDestroy(&calc);
	free(buf); free(vals); free(v.slots); free(v.used);
	return rc;
}
