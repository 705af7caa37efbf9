/*
 * simcu_host.h - host side of libsimulator_cuda that needs no GPU: the netlist
 * builder (rows, structural non-zeros, devices in insertion order), the
 * B-source expression compiler, the flattening of a netlist into the device
 * program of simcu_program.h, and the symbolic LU analysis that turns a pivot
 * sequence into the elimination op-stream the kernels walk.
 *
 * Reference behaviour mirrored here: core/row.c, core/node.c, core/matrix.c
 * (row / pattern creation order), devices/*.c *Config functions (which rows and
 * entries each device adds), libs/calculon/tokenizer.re + parser.lem (the
 * expression language).  Each function cites its lines in simcu_host.cpp.
 */
#ifndef SIMCU_HOST_H
#define SIMCU_HOST_H

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "simcu_program.h"

namespace simcu {

/* ---- B-source expressions ------------------------------------------------ */
struct CalcProg {
	std::vector<int> code;            /* (opcode, operand) pairs, postfix */
	std::vector<double> consts;       /* OP_CONST operands index this */
	std::vector<std::string> varNames;/* distinct, first-appearance order */
	std::vector<int> varRef;          /* what the owner's callback returned */
	int maxStack = 0;
};
typedef int (*CalcGetVar)(void *ctx, const char *name);   /* < 0: reject */
bool calcCompile(const char *expr, CalcGetVar gv, void *ctx, CalcProg &out,
		std::string &err);

/* ---- netlist -------------------------------------------------------------- */
struct TableSet { double **vi; int *viLen; double **ta; int *taLen; };

struct Dev {
	int kind = 0;
	std::string refdes;
	int pin[4] = {0, 0, 0, 0};        /* rows; 0 = ground */
	int rowR = 0, rowM = 0;           /* branch / internal rows (T: in1, in2) */
	const double *val = nullptr;      /* R C L, or a source's DC (borrowed) */
	const double *z0 = nullptr, *td = nullptr, *loss = nullptr;
	char wtype = 0;
	const double *wargs[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	double **wtab = nullptr; int *wtabLen = nullptr;
	std::vector<std::pair<double **, int *> > wsets;   /* further PWL / PWC tables (set 1, 2, ...) */
	std::vector<TableSet> sets; char viType = 'l', taType = 'l';
	CalcProg calc; bool usesTime = false;
	std::vector<int> bvRow, bvCalcVar;
	std::map<std::string, std::vector<double> > inst;   /* per-instance values */
	std::vector<int> instSet;                            /* per-instance table set */
};

struct Netlist {
	std::vector<std::string> rowName;             /* [0] = "v(0)" */
	std::set<std::pair<int, int> > nodes;         /* (col, row): CSC order */
	std::vector<Dev> devs;
	Netlist() { rowName.push_back("v(0)"); }
	int findOrAddRow(char type, const char *name);   /* type 0: full "v(x)" name */
	int findRow(const char *fullName) const;
	void addNode(int row, int col) { if (row && col) nodes.insert(std::make_pair(col, row)); }
	Dev *find(const char *refdes);
	bool duplicate(const char *refdes) const;
};

/* ---- compiled program (host copy of what SimcuArgs points at) ------------- */
struct Program {
	int N = 0, nnzA = 0, ndev = 0;
	std::vector<int> colStart, rowIdx;            /* CSC pattern, 0-based */
	std::vector<int> dev;                         /* [ndev][SIMCU_DEV_STRIDE] */
	std::vector<int> bvar;
	std::vector<int> calcOff, calcCode, calcVarOff, calcVars;
	std::vector<double> calcConst;
	std::vector<int> tabOff, tabType;
	std::vector<int> tabClass;                    /* 0 V-I table (looked up every Newton iteration), 1 TA, 2 source */
	std::vector<double> tabP;                     /* [point][4] = x, y, m, x_next */
	std::vector<int> histRows;
	std::vector<int> pmap; std::vector<double> pconst;
	std::vector<double> pinst;                    /* [ninst][M] */
	std::vector<int> iinst;                       /* [niinst][M] */
	int ninst = 0, niinst = 0;
	std::vector<int> catInst;                     /* pinst columns that are flags (T-line bypass): pilots cover every value */
	int stateDoubles = 0, stateInts = 0;
	std::vector<char> activeOP, activeTran;       /* [nnzA] slot can be non-zero */
	/* slots that hold the same compile-time constant for every instance in a
	 * phase (+-1 incidence entries of branch rows): [nnzA], NaN = not constant */
	std::vector<double> constOP, constTran;
	int maxCalcStack = 0;
};

/* one table in the packed form the kernels read, appended to out: per point
 * (x, y, m, x of the next point); type 'l' linear or 'c' natural cubic spline */
bool packTable(const double *xy, int n, char type, std::vector<double> &out, std::string &err);
/* flatten; M = instances.  false + err on failure */
bool compileProgram(const Netlist &nl, int M, Program &p, std::string &err);
/* re-read only the scalar parameters and tables (they are borrowed) */
bool refreshParameters(const Netlist &nl, int M, Program &p, std::string &err);

/* ---- symbolic LU ----------------------------------------------------------- */
struct Schedule {
	std::vector<int> pc, pr;          /* step k: column pc[k], pivot row pr[k] */
	std::vector<int> lu;              /* op-stream, see simcu_kernels.cu factorSolve */
	std::vector<int> xslot;           /* [N] workspace slot of x_k after the solve */
	int F = 0;                        /* matrix slots incl. fill */
	int nL = 0, nU = 0;               /* off-diagonal factor entries */
	long long flops = 0;              /* multiply-adds of one factor + solve */
	double worstRatio = 1.0;          /* min over steps/pilots of |pivot| / max|column| */
};
/* Chooses (unless forced) a pivot sequence by Markowitz count under threshold
 * partial pivoting judged on every pilot matrix, then emits the op-stream.
 * pilots[p][slot] = value of A slot for pilot p; active[slot] = 0 marks slots
 * that are exactly zero for every instance in this phase. */
bool analyse(int N, const std::vector<int> &colStart, const std::vector<int> &rowIdx,
		const std::vector<char> &active,
		const std::vector<std::vector<double> > &pilots, double threshold,
		const int *forcePc, const int *forcePr, Schedule &out, std::string &err);

}  /* namespace simcu */
#endif
