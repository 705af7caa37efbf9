/*
 * ref_probe.c - link-time probe of the UNMODIFIED reference build (test
 * infrastructure only): reads the permutations SuperLU's dgssvx leaves behind.
 *
 * The reference keeps perm_c / perm_r in a struct private to core/matrix.c
 * (:57-78) and passes them to dgssvx on every factorisation (:104,136).  The
 * reference library is linked with -Wl,--wrap=dgssvx (build_ref.sh), so the
 * call in matrix.o lands here, goes straight on to the real dgssvx, and - only
 * while capturing is switched on - the two permutations of every full
 * factorisation (options->Fact == DOFACT: the first solve of the operating
 * point and of every transient attempt, core/matrix.c:82-118) are copied out.
 * No reference source is modified or copied; with capturing off the wrapper
 * adds one branch per call.
 *
 * After dgssvx (libs/superlu/SRC/dgssvx.c:500-534, sp_preorder.c:16-21):
 *   perm_c[i] = j   column i of A is column j of A*Pc, etree post-order included
 *   perm_r[i] = j   row i of A is row j of Pr*A = the pivot row of step j
 * so the engine's sequence "step k eliminates column pc[k] with pivot row
 * pr[k]" (include/simulator_cuda.h: simcuSetPivots) is the inverse of both.
 */
#include <stdlib.h>
#include <string.h>
#include <superlu.h>

void __real_dgssvx(superlu_options_t *, SuperMatrix *, int *, int *, int *, char *, double *,
		double *, SuperMatrix *, SuperMatrix *, void *, int, SuperMatrix *, SuperMatrix *,
		double *, double *, double *, double *, mem_usage_t *, SuperLUStat_t *, int *);

static int probeOn = 0, probeN = 0, probeCount = 0, probeCap = 0, probeMax = 0;
static int *probeData = NULL;      /* [count][2][n]: perm_c, perm_r */

void __wrap_dgssvx(superlu_options_t *options, SuperMatrix *A, int *perm_c, int *perm_r,
		int *etree, char *equed, double *R, double *C, SuperMatrix *L, SuperMatrix *U,
		void *work, int lwork, SuperMatrix *B, SuperMatrix *X, double *rpg, double *rcond,
		double *ferr, double *berr, mem_usage_t *mem_usage, SuperLUStat_t *stat, int *info)
{
	const int full = (options->Fact == DOFACT);
	__real_dgssvx(options, A, perm_c, perm_r, etree, equed, R, C, L, U, work, lwork, B, X,
			rpg, rcond, ferr, berr, mem_usage, stat, info);
	if (!probeOn || !full || *info != 0) return;
	if (probeCount >= probeMax) return;
	if (probeN != A->ncol) {
		free(probeData); probeData = NULL;
		probeN = A->ncol; probeCount = 0; probeCap = 0;
	}
	if (probeCount == probeCap) {
		probeCap = probeCap ? 2 * probeCap : 64;
		probeData = realloc(probeData, sizeof(int) * 2 * (size_t)probeN * probeCap);
	}
	memcpy(probeData + (size_t)2 * probeN * probeCount, perm_c, sizeof(int) * probeN);
	memcpy(probeData + (size_t)2 * probeN * probeCount + probeN, perm_r, sizeof(int) * probeN);
	probeCount++;
}

/* start (maxRecords > 0) / stop (0) capturing; starting clears what was kept */
void ref_probe_capture(int maxRecords)
{
	probeOn = (maxRecords > 0);
	probeMax = maxRecords;
	if (probeOn) probeCount = 0;
}

/* number of full factorisations captured, unknowns in *n */
int ref_probe_count(int *n)
{
	if (n) *n = probeN;
	return probeCount;
}

/* the k-th captured factorisation as the engine's sequence: pc[j] / pr[j] =
 * column / pivot row (0-based unknowns) eliminated at step j */
int ref_probe_sequence(int k, int *pc, int *pr)
{
	int i;
	const int *c, *r;
	if (k < 0 || k >= probeCount) return -1;
	c = probeData + (size_t)2 * probeN * k;
	r = c + probeN;
	for (i = 0; i < probeN; i++) { pc[c[i]] = i; pr[r[i]] = i; }
	return 0;
}
