/*
 * simcu_host.cpp - netlist builder, expression compiler, program flattening
 * and symbolic LU analysis of libsimulator_cuda.  No CUDA in this file.
 *
 * Reference lines are relative to the reference root.
 */
#include "simcu_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace simcu {

/* ========================================================================== *
 * Expression compiler.
 *
 * The reference tokenises a B-source equation once (libs/calculon/
 * tokenizer.re:99-257) and replays the tokens through an LALR(1) parser whose
 * actions evaluate on the fly (parser.lem:60-149).  Here the same token
 * language is parsed once, by precedence climbing that reproduces the
 * grammar's %left table (parser.lem:52-58), into postfix bytecode for the
 * interpreter in simcu_calc.h.
 * ========================================================================== */
namespace {

enum Tk {
	TK_END, TK_VAR, TK_CONST, TK_PLUS, TK_MINUS, TK_TIMES, TK_DIVIDE, TK_POWER,
	TK_LPAREN, TK_RPAREN, TK_FUNC, TK_IF, TK_GT, TK_LT, TK_EQ, TK_NOT, TK_AND,
	TK_OR
};
struct Token { int type; int id; double value; };

/* tokenizer.re:121-139; every pattern ends in '(' so at most one matches */
const char *const kFuncs[] = {
	"abs(", "acosh(", "acos(", "asinh(", "asin(", "atanh(", "atan(", "cosh(",
	"cos(", "exp(", "ln(", "log(", "sinh(", "sin(", "sqrt(", "tan(", "uramp(",
	"u(", nullptr
};

inline bool isAlphaCh(int c)   /* ALPHA, tokenizer.re:43 */
{
	return std::isalpha(c) || c == '_' || c == ':' || c == '#' || c == '@';
}
inline bool isTextCh(int c) { return isAlphaCh(c) || std::isdigit(c); }

bool startsNoCase(const char *s, const char *lit)
{
	for (; *lit; s++, lit++)
		if (std::tolower((unsigned char)*s) != *lit) return false;
	return true;
}

struct Lexer {
	CalcProg &prog;
	CalcGetVar gv; void *ctx;
	std::vector<Token> tok;
	std::string err;

	Lexer(CalcProg &p, CalcGetVar g, void *c) : prog(p), gv(g), ctx(c) {}

	void push(int type, int id = 0, double v = 0.0)
	{
		Token t; t.type = type; t.id = id; t.value = v; tok.push_back(t);
	}

	/* calc.c:104-121: one variable per distinct name */
	int variable(const std::string &name)
	{
		for (size_t i = 0; i < prog.varNames.size(); i++)
			if (prog.varNames[i] == name) return (int)i;
		const int ref = gv(ctx, name.c_str());
		if (ref < 0) { err = "unknown variable " + name; return -1; }
		prog.varNames.push_back(name);
		prog.varRef.push_back(ref);
		return (int)prog.varNames.size() - 1;
	}

	bool run(const char *p)
	{
		for (;;) {
			const unsigned char ch = (unsigned char)*p;
			if (ch == ' ' || ch == '\t') { p++; continue; }
			if (ch == '\n' || ch == '\r') {             /* CONTINUE = EOL+ "+" */
				const char *q = p;
				while (*q == '\n' || *q == '\r') q++;
				if (*q != '+') { err = "unexpected end of line"; return false; }
				p = q + 1; continue;
			}
			if (ch == 0) { push(TK_END); return true; }
			int f = 0;
			while (kFuncs[f] && !startsNoCase(p, kFuncs[f])) f++;
			if (kFuncs[f]) { push(TK_FUNC, f); p += std::strlen(kFuncs[f]); continue; }
			if (startsNoCase(p, "if(")) { push(TK_IF); p += 3; continue; }
			int single = -1;
			switch (ch) {
			case '+': single = TK_PLUS; break;   case '-': single = TK_MINUS; break;
			case '*': single = TK_TIMES; break;  case '/': single = TK_DIVIDE; break;
			case '^': single = TK_POWER; break;  case '(': single = TK_LPAREN; break;
			case ')': single = TK_RPAREN; break; case '>': single = TK_GT; break;
			case '<': single = TK_LT; break;     case '=': single = TK_EQ; break;
			case '!': single = TK_NOT; break;    case '&': single = TK_AND; break;
			case '|': single = TK_OR; break;
			default: break;
			}
			if (single >= 0) { push(single); p++; continue; }
			if (std::isdigit(ch) || (ch == '.' && std::isdigit((unsigned char)p[1]))) {
				p = number(p);
				if (!p) return false;
				continue;
			}
			if ((ch == 'v' || ch == 'V' || ch == 'i' || ch == 'I') && p[1] == '(') {
				const char *q = nodeRef(p);
				if (!q && !err.empty()) return false;
				if (q) { p = q; continue; }
			}
			if (isAlphaCh(ch)) {                         /* VARIABLE, :229-239 */
				const char *q = p;
				while (isTextCh((unsigned char)*q)) q++;
				const int v = variable(std::string(p, q));
				if (v < 0) return false;
				push(TK_VAR, v);
				p = q; continue;
			}
			err = std::string("unexpected character '") + (char)ch + "'";
			return false;
		}
	}

	/* FLOAT [unit] UNITS, tokenizer.re:153-163.  re2c takes the longest match
	 * and, on a tie, the first rule in the file: T G k u n p f Meg mil m */
	const char *number(const char *p)
	{
		const char *q = p;
		while (std::isdigit((unsigned char)*q)) q++;
		if (*q == '.' && std::isdigit((unsigned char)q[1])) {
			q++;
			while (std::isdigit((unsigned char)*q)) q++;
		}
		if (*q == 'e' || *q == 'E') {
			const char *e = q + 1;
			if (*e == '-' || *e == '+') e++;
			if (std::isdigit((unsigned char)*e)) {
				while (std::isdigit((unsigned char)*e)) e++;
				q = e;
			}
		}
		const std::string text(p, q);
		const double v = std::atof(text.c_str());
		static const struct { const char *s; double m; } units[] = {
			{"t", 1e12}, {"g", 1e9}, {"k", 1e3}, {"u", 1e-6}, {"n", 1e-9},
			{"p", 1e-12}, {"f", 1e-15}, {"meg", 1e6}, {"mil", 2.54e-5}, {"m", 1e-3}
		};
		double value = v;
		for (size_t i = 0; i < sizeof units / sizeof units[0]; i++)
			if (startsNoCase(q, units[i].s)) { value = v * units[i].m; break; }
		while (isAlphaCh((unsigned char)*q)) q++;
		push(TK_CONST, 0, value);
		return q;
	}

	/* 'v(' TEXT+ ',' TEXT+ ')' -> ( v(a) - v(b) )  (:165-212) and
	 * ('v'|'i') '(' TEXT+ ')'                      (:213-228).
	 * Returns nullptr without an error when neither form matches. */
	const char *nodeRef(const char *p)
	{
		const char kind = (char)std::tolower((unsigned char)*p);
		const char *a0 = p + 2, *q = a0;
		while (isTextCh((unsigned char)*q)) q++;
		if (q == a0) return nullptr;
		const std::string a(a0, q);
		if (*q == ')') {
			const int v = variable(std::string(1, kind) + "(" + a + ")");
			if (v < 0) return nullptr;
			push(TK_VAR, v);
			return q + 1;
		}
		if (kind == 'v' && *q == ',') {
			const char *b0 = q + 1, *e = b0;
			while (isTextCh((unsigned char)*e)) e++;
			if (e == b0 || *e != ')') return nullptr;
			const int va = variable("v(" + a + ")");
			if (va < 0) return nullptr;
			push(TK_LPAREN); push(TK_VAR, va); push(TK_MINUS);
			const int vb = variable("v(" + std::string(b0, e) + ")");
			if (vb < 0) return nullptr;
			push(TK_VAR, vb); push(TK_RPAREN);
			return e + 1;
		}
		return nullptr;
	}
};

/* Grammar of parser.lem.  Precedence, low to high (parser.lem:52-58):
 * + -  <  * /  <  ^  <  unary minus (NOT)  <  function tokens  <  ( ) .
 * All binary operators are left-associative, so 2^3^2 = 64; unary minus binds
 * tighter than ^, so -2^2 = 4; "a(b)" is a product that binds tightest. */
struct Parser {
	const std::vector<Token> &tok;
	CalcProg &prog;
	size_t pos = 0;
	bool bad = false;

	Parser(const std::vector<Token> &t, CalcProg &p) : tok(t), prog(p) {}

	int peek(size_t ahead = 0) const
	{
		const size_t i = pos + ahead;
		return (i < tok.size()) ? tok[i].type : (int)TK_END;
	}
	void emit(int op, int arg = 0) { prog.code.push_back(op); prog.code.push_back(arg); }
	bool expect(int type)
	{
		if (bad || peek() != type) { bad = true; return false; }
		pos++;
		return true;
	}

	void primary()
	{
		const Token &t = tok[pos];
		switch (t.type) {
		case TK_VAR: emit(OP_VAR, t.id); pos++; return;
		case TK_CONST:
			prog.consts.push_back(t.value);
			emit(OP_CONST, (int)prog.consts.size() - 1); pos++; return;
		case TK_LPAREN: pos++; sum(); expect(TK_RPAREN); return;
		case TK_FUNC: pos++; sum(); if (expect(TK_RPAREN)) emit(OP_FUNC, t.id); return;
		case TK_IF: pos++; condition(); if (expect(TK_RPAREN)) emit(OP_IF); return;
		default: bad = true; return;
		}
	}
	void product()      /* expr LPAREN expr RPAREN, parser.lem:68 */
	{
		primary();
		while (!bad && peek() == TK_LPAREN) {
			pos++; sum();
			if (expect(TK_RPAREN)) emit(OP_MUL);
		}
	}
	void unary()        /* MINUS expr [NOT], :66; no unary plus in value mode */
	{
		if (peek() == TK_MINUS) { pos++; unary(); emit(OP_NEG); return; }
		product();
	}
	void power()
	{
		unary();
		while (!bad && peek() == TK_POWER) { pos++; unary(); emit(OP_POW); }
	}
	void term()
	{
		power();
		while (!bad) {
			const int t = peek();
			if (t == TK_TIMES) { pos++; power(); emit(OP_MUL); }
			else if (t == TK_DIVIDE) { pos++; power(); emit(OP_DIV); }
			else break;
		}
	}
	void sum()
	{
		term();
		while (!bad) {
			const int t = peek();
			if (t == TK_PLUS) { pos++; term(); emit(OP_ADD); }
			else if (t == TK_MINUS) { pos++; term(); emit(OP_SUB); }
			else break;
		}
	}

	/* eval, parser.lem:129-138.  "(" may open a parenthesised condition or an
	 * arithmetic operand of a comparison; try the former, fall back to the
	 * latter. */
	void condition()
	{
		if (peek() == TK_NOT) { pos++; condition(); emit(OP_NOT); return; }
		if (peek() == TK_LPAREN) {
			const size_t savePos = pos, saveCode = prog.code.size();
			const size_t saveConst = prog.consts.size();
			pos++;
			condition();
			if (!bad && peek() == TK_RPAREN) {
				pos++;
				const int t = peek();
				if ((t == TK_AND || t == TK_OR) && peek(1) == t) {
					pos += 2;
					if (!expect(TK_LPAREN)) return;
					condition();
					if (expect(TK_RPAREN)) emit(t == TK_AND ? OP_AND : OP_OR);
				}
				return;
			}
			pos = savePos; bad = false;
			prog.code.resize(saveCode); prog.consts.resize(saveConst);
		}
		sum();
		if (bad) return;
		switch (peek()) {
		case TK_GT:
			pos++;
			if (peek() == TK_EQ) { pos++; sum(); emit(OP_GE); } else { sum(); emit(OP_GT); }
			return;
		case TK_LT:
			pos++;
			if (peek() == TK_EQ) { pos++; sum(); emit(OP_LE); } else { sum(); emit(OP_LT); }
			return;
		case TK_EQ:
			if (peek(1) != TK_EQ) { bad = true; return; }
			pos += 2; sum(); emit(OP_EQ); return;
		case TK_NOT:
			if (peek(1) != TK_EQ) { bad = true; return; }
			pos += 2; sum(); emit(OP_NE); return;
		default: bad = true; return;
		}
	}
};

}  /* namespace */

bool calcCompile(const char *expr, CalcGetVar gv, void *ctx, CalcProg &out,
		std::string &err)
{
	out = CalcProg();
	if (expr == nullptr || gv == nullptr) { err = "null expression"; return false; }
	Lexer lex(out, gv, ctx);
	if (!lex.run(expr)) { err = lex.err; return false; }
	Parser ps(lex.tok, out);
	ps.sum();
	if (ps.bad || ps.peek() != TK_END) { err = "syntax error"; return false; }
	/* stack depth of the postfix program */
	int sp = 0, top = 0;
	for (size_t i = 0; i < out.code.size(); i += 2) {
		switch (out.code[i]) {
		case OP_CONST: case OP_VAR: sp++; break;
		case OP_NEG: case OP_FUNC: case OP_NOT: case OP_IF: break;
		default: sp--; break;
		}
		top = std::max(top, sp);
	}
	out.maxStack = top;
	if (sp != 1) { err = "syntax error"; return false; }
	if (top > CALC_MAX_STACK) { err = "expression too deeply nested"; return false; }
	if ((int)out.varNames.size() > CALC_MAX_VARS) { err = "too many variables"; return false; }
	return true;
}

/* ========================================================================== *
 * Netlist: rows and structural non-zeros.
 * core/row.c:46-66,163-204 (names "v(<node>)" / "i(<refdes>)", first-mention
 * order, row 0 = ground), core/matrix.c:292-346, core/node.c:56-66 (entries
 * kept sorted by column then row; anything touching ground is dropped).
 * ========================================================================== */
int Netlist::findRow(const char *fullName) const
{
	for (size_t i = 0; i < rowName.size(); i++)
		if (rowName[i] == fullName) return (int)i;
	return -1;
}

int Netlist::findOrAddRow(char type, const char *name)
{
	std::string full;
	if (type == 0) {
		full = name;
		const int hit = findRow(full.c_str());
		if (hit >= 0) return hit;
		if (full.size() < 4) return -1;
		const char c = full[0];
		if (c != 'i' && c != 'v' && c != 'I' && c != 'V') return -1;
		full[0] = (char)std::tolower((unsigned char)c);
	} else {
		full = std::string(1, type) + "(" + name + ")";
		const int hit = findRow(full.c_str());
		if (hit >= 0) return hit;
	}
	rowName.push_back(full);
	return (int)rowName.size() - 1;
}

Dev *Netlist::find(const char *refdes)
{
	for (size_t i = 0; i < devs.size(); i++)
		if (devs[i].refdes == refdes) return &devs[i];
	return nullptr;
}

bool Netlist::duplicate(const char *refdes) const
{
	for (size_t i = 0; i < devs.size(); i++)
		if (devs[i].refdes == refdes) return true;
	return false;
}

/* ========================================================================== *
 * Flattening a netlist into the device program
 * ========================================================================== */
/* One table in the packed form the kernels read: per point (x, y, m, x of the
 * next point), m = slope to the next point (linear, piecewise.c:183-201) or
 * second derivative of the natural cubic spline (piecewise.c:108-181), appended
 * to out. */
bool packTable(const double *xy, int n, char type, std::vector<double> &out, std::string &err)
{
	if (type == 'c' && n < 2) { err = "a spline table needs at least two points"; return false; }
	std::vector<double> x(n), y(n), m(n, 0.0);
	for (int i = 0; i < n; i++) { x[i] = xy[2 * i]; y[i] = xy[2 * i + 1]; }
	for (int i = 0; i + 1 < n; i++)
		if (x[i] > x[i + 1]) { err = "table x values must be sorted"; return false; }
	if (type == 'l') {
		for (int i = 0; i + 1 < n; i++) m[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
	} else if (n > 2) {
		const int last = n - 1;
		std::vector<double> h(last), b(last), u(last, 0.0), v(last, 0.0);
		for (int i = 0; i < last; i++) {
			h[i] = x[i + 1] - x[i];
			b[i] = 6.0 * (y[i + 1] - y[i]) / h[i];
		}
		u[1] = 2.0 * h[0] + 2.0 * h[1];
		v[1] = b[1] - b[0];
		bool singular = (u[1] == 0.0);
		for (int i = 2; i < last && !singular; i++) {
			u[i] = 2.0 * h[i] + 2.0 * h[i - 1] - (h[i - 1] * h[i - 1] / u[i - 1]);
			singular = (u[i] == 0.0);
			v[i] = b[i] - b[i - 1] - (h[i - 1] * v[i - 1] / u[i - 1]);
		}
		if (singular) { err = "spline table is singular"; return false; }
		for (int i = last - 1; i > 0; i--) m[i] = (v[i] - h[i] * m[i + 1]) / u[i];
	}
	for (int i = 0; i < n; i++) {
		out.push_back(x[i]); out.push_back(y[i]); out.push_back(m[i]);
		out.push_back((i + 1 < n) ? x[i + 1] : HUGE_VAL);
	}
	return true;
}

namespace {

struct Flattener {
	const Netlist &nl;
	int M;
	Program &p;
	std::map<std::pair<int, int>, int> slotOf;    /* (row, col) -> A slot */
	std::string err;

	Flattener(const Netlist &n, int m, Program &prog) : nl(n), M(m), p(prog) {}

	int slot(int row, int col) const
	{
		if (!row || !col) return -1;
		std::map<std::pair<int, int>, int>::const_iterator it =
			slotOf.find(std::make_pair(row, col));
		return (it == slotOf.end()) ? -1 : it->second;
	}
	void mark(std::vector<char> &mask, int s) { if (s >= 0) mask[s] = 1; }
	void both(int s) { mark(p.activeOP, s); mark(p.activeTran, s); }
	/* every mention of a slot by a device; a slot keeps a constant only while
	 * exactly one mention exists and that one sets a literal */
	std::vector<int> refs;
	void ref(int s) { if (s >= 0) refs[s]++; }
	void lit(int s, double op, double tran)
	{
		if (s < 0) return;
		refs[s]++;
		p.constOP[s] = op; p.constTran[s] = tran;
	}

	/* one scalar parameter: per-instance row if the device has one */
	int param(const Dev &d, const char *name, const double *src)
	{
		const int pid = (int)p.pmap.size();
		std::map<std::string, std::vector<double> >::const_iterator it = d.inst.find(name);
		if (it != d.inst.end()) {
			if (!std::strcmp(name, "bypass")) p.catInst.push_back(p.ninst);
			p.pmap.push_back(p.ninst);
			p.pconst.push_back(0.0);
			p.pinst.insert(p.pinst.end(), it->second.begin(), it->second.end());
			p.ninst++;
		} else {
			p.pmap.push_back(-1);
			p.pconst.push_back(src ? *src : HUGE_VAL);
		}
		return pid;
	}

	/* piecewise.c:183-201 (linear slopes) and :108-181 (natural cubic spline) */
	bool table(double **src, const int *len, char type, int &id, int klass = 2)
	{
		if (!src || !*src || !len || *len < 1) { err = "missing table"; return false; }
		id = (int)p.tabType.size();
		if (!packTable(*src, *len, type, p.tabP, err)) return false;
		p.tabType.push_back(type);
		p.tabClass.push_back(klass);
		p.tabOff.push_back((int)p.tabP.size() / 4);
		return true;
	}

	/* parameters and tables, in device order; fills rec when it is given */
	bool parameters(bool writeRecords)
	{
		p.pmap.clear(); p.pconst.clear(); p.pinst.clear(); p.iinst.clear(); p.catInst.clear();
		p.ninst = 0; p.niinst = 0;
		p.tabOff.assign(1, 0); p.tabType.clear(); p.tabClass.clear();
		p.tabP.clear();
		for (size_t k = 0; k < nl.devs.size(); k++) {
			const Dev &d = nl.devs[k];
			int *rec = writeRecords ? &p.dev[k * SIMCU_DEV_STRIDE] : nullptr;
			int a = -1, b = -1, c = -1;
			switch (d.kind) {
			case DK_R: a = param(d, "R", d.val); if (rec) rec[DF_P0] = a; break;
			case DK_C: a = param(d, "C", d.val); if (rec) rec[DF_P0] = a; break;
			case DK_L: a = param(d, "L", d.val); if (rec) rec[DF_P0] = a; break;
			case DK_V: case DK_I: {
				a = param(d, "DC", d.val);
				int first = -1, tab = -1, wsetSlot = -1;
				if (d.wtype == 'l' || d.wtype == 'c') {
					if (!table(d.wtab, d.wtabLen, d.wtype, tab)) return false;
					for (size_t s = 0; s < d.wsets.size(); s++) {
						int id;
						if (!table(d.wsets[s].first, d.wsets[s].second, d.wtype, id)) return false;
					}
					if (!d.instSet.empty()) {
						for (size_t i = 0; i < d.instSet.size(); i++)
							if (d.instSet[i] < 0 || d.instSet[i] > (int)d.wsets.size()) {
								err = "table set index out of range for " + d.refdes;
								return false;
							}
						wsetSlot = p.niinst++;
						p.iinst.insert(p.iinst.end(), d.instSet.begin(), d.instSet.end());
					}
				} else if (d.wtype) {
					for (int i = 0; i < 7; i++) {
						char nm[3] = {'A', (char)('0' + i), 0};
						const int pid = param(d, nm, d.wargs[i]);
						if (i == 0) first = pid;
					}
				}
				if (rec) {
					rec[DF_PDC] = a; rec[DF_WTYPE] = d.wtype; rec[DF_PARGS] = first; rec[DF_WTAB] = tab;
					rec[DF_WSETSLOT] = wsetSlot;
				}
				break;
			}
			case DK_T:
				a = param(d, "Z0", d.z0); b = param(d, "Td", d.td); c = param(d, "loss", d.loss);
				if (rec) { rec[DF_PZ0] = a; rec[DF_PTD] = b; rec[DF_PLOSS] = c; }
				if (d.inst.count("bypass")) {      /* masked maximum topology: per-instance flag */
					const int m = param(d, "bypass", nullptr);
					if (rec) rec[DF_PBYPASS] = m;
				}
				break;
			case DK_VI: {
				int viTab = -1, taTab = -1;
				const bool hasTa = d.sets[0].ta && *d.sets[0].ta && d.sets[0].taLen && *d.sets[0].taLen > 0;
				for (size_t s = 0; s < d.sets.size(); s++) {
					int id;
					if (!table(d.sets[s].vi, d.sets[s].viLen, d.viType, id, 0)) return false;
					if (s == 0) viTab = id;
				}
				for (size_t s = 0; hasTa && s < d.sets.size(); s++) {
					int id;
					if (!table(d.sets[s].ta, d.sets[s].taLen, d.taType, id, 1)) {
						err = "every table set of " + d.refdes + " needs a TA table";
						return false;
					}
					if (s == 0) taTab = id;
				}
				int setSlot = -1;
				if (!d.instSet.empty()) {
					for (size_t i = 0; i < d.instSet.size(); i++)
						if (d.instSet[i] < 0 || d.instSet[i] >= (int)d.sets.size()) {
							err = "table set index out of range for " + d.refdes;
							return false;
						}
					setSlot = p.niinst++;
					p.iinst.insert(p.iinst.end(), d.instSet.begin(), d.instSet.end());
				}
				if (rec) {
					rec[DF_VITAB] = viTab; rec[DF_TATAB] = taTab;
					rec[DF_NSETS] = (int)d.sets.size(); rec[DF_SETSLOT] = setSlot;
				}
				break;
			}
			default: break;
			}
		}
		return true;
	}

	bool run()
	{
		p = Program();
		p.N = (int)nl.rowName.size() - 1;
		p.ndev = (int)nl.devs.size();
		if (p.N < 1 || p.ndev < 1) { err = "empty circuit"; return false; }
		/* CSC pattern: matrix.c:476-549 */
		p.colStart.assign(p.N + 1, 0);
		for (std::set<std::pair<int, int> >::const_iterator it = nl.nodes.begin();
				it != nl.nodes.end(); ++it) {
			slotOf[std::make_pair(it->second, it->first)] = (int)p.rowIdx.size();
			p.rowIdx.push_back(it->second - 1);
			p.colStart[it->first]++;
		}
		for (int c = 0; c < p.N; c++) p.colStart[c + 1] += p.colStart[c];
		p.nnzA = (int)p.rowIdx.size();
		p.activeOP.assign(p.nnzA, 0); p.activeTran.assign(p.nnzA, 0);
		p.constOP.assign(p.nnzA, NAN); p.constTran.assign(p.nnzA, NAN);
		refs.assign(p.nnzA, 0);
		p.dev.assign((size_t)p.ndev * SIMCU_DEV_STRIDE, -1);
		p.calcOff.assign(1, 0); p.calcVarOff.assign(1, 0);

		/* T-line history columns: every row a T-line looks up in the past */
		std::map<int, int> histCol;
		for (size_t k = 0; k < nl.devs.size(); k++) {
			const Dev &d = nl.devs[k];
			if (d.kind != DK_T) continue;
			const int rows[6] = {d.rowR, d.rowM, d.pin[0], d.pin[1], d.pin[2], d.pin[3]};
			for (int j = 0; j < 6; j++)
				if (rows[j] && !histCol.count(rows[j])) {
					const int col = (int)p.histRows.size();
					histCol[rows[j]] = col;
					p.histRows.push_back(rows[j]);
				}
		}

		int sbase = 0, ibase = 0;
		for (size_t k = 0; k < nl.devs.size(); k++) {
			const Dev &d = nl.devs[k];
			int *rec = &p.dev[k * SIMCU_DEV_STRIDE];
			const int K = d.pin[0], J = d.pin[1], L = d.pin[2], Mm = d.pin[3];
			const int R = d.rowR, S = d.rowM;
			rec[DF_TYPE] = d.kind;
			rec[DF_K] = K; rec[DF_J] = J; rec[DF_L] = L; rec[DF_M] = Mm;
			rec[DF_ROWR] = R; rec[DF_ROWM] = S;
			rec[DF_SBASE] = sbase; rec[DF_IBASE] = ibase;
			switch (d.kind) {
			case DK_R: {                       /* resistor.c:134-141 */
				const int s[4] = {slot(K, K), slot(K, J), slot(J, K), slot(J, J)};
				for (int j = 0; j < 4; j++) { rec[DF_A0 + j] = s[j]; both(s[j]); ref(s[j]); }
				break;
			}
			case DK_C: {                       /* capacitor.c:250-257 */
				const int s[4] = {slot(K, K), slot(K, J), slot(J, K), slot(J, J)};
				for (int j = 0; j < 4; j++) { rec[DF_A0 + j] = s[j]; mark(p.activeTran, s[j]); ref(s[j]); }
				sbase += 2 + INTEG_DOUBLES; ibase += 1;
				break;
			}
			case DK_L: {                       /* inductor.c:251-266 */
				const int s[5] = {slot(R, K), slot(R, J), slot(K, R), slot(J, R), slot(R, R)};
				static const double pm[4] = {1.0, -1.0, 1.0, -1.0};
				for (int j = 0; j < 4; j++) { rec[DF_A0 + j] = s[j]; both(s[j]); lit(s[j], pm[j], pm[j]); }
				rec[DF_A0 + 4] = s[4]; mark(p.activeTran, s[4]); ref(s[4]);
				sbase += 2 + INTEG_DOUBLES; ibase += 1;
				break;
			}
			case DK_V: {                       /* source_v.c:240-249 */
				const int s[4] = {slot(R, K), slot(R, J), slot(K, R), slot(J, R)};
				static const double pm[4] = {1.0, -1.0, 1.0, -1.0};
				for (int j = 0; j < 4; j++) { rec[DF_SA0 + j] = s[j]; both(s[j]); lit(s[j], pm[j], pm[j]); }
				sbase += 2 + CB_DOUBLES; ibase += 2;
				break;
			}
			case DK_I:
				sbase += 2 + CB_DOUBLES; ibase += 2;
				break;
			case DK_T: {                       /* tline.c:397-444 */
				/* RK RJ KR JR RL RM SL SM LS MS SK SJ RR SS KS JS LR MR */
				const int s[18] = {
					slot(R, K), slot(R, J), slot(K, R), slot(J, R), slot(R, L), slot(R, Mm),
					slot(S, L), slot(S, Mm), slot(L, S), slot(Mm, S), slot(S, K), slot(S, J),
					slot(R, R), slot(S, S), slot(K, S), slot(J, S), slot(L, R), slot(Mm, R)};
				static const char inTran[18] = {1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 1, 1, 0, 0, 0, 0};
				/* devLoad (tline.c:236-267) and devInitStep (:104-111) literals */
				static const double atOP[18] = {1, -1, 1, -1, -1, 1, 1, -1, 1, -1, -1, 1, 0, 0, 1, -1, -1, 1};
				/* a section that some instances bypass (ideal through-connection, see
				 * devLoad) keeps all 18 entries in both patterns, none of them a literal */
				const bool maskable = d.inst.count("bypass") != 0;
				for (int j = 0; j < 18; j++) {
					rec[DF_TA0 + j] = s[j];
					mark(p.activeOP, s[j]);
					if (inTran[j] || maskable) mark(p.activeTran, s[j]);
					if (j == 12 || j == 13 || maskable) ref(s[j]);          /* RR, SS = -Z0 */
					else lit(s[j], atOP[j], inTran[j] ? atOP[j] : 0.0);
				}
				const int rows[6] = {R, S, K, J, L, Mm};
				for (int j = 0; j < 6; j++) rec[DF_TH0 + j] = rows[j] ? histCol[rows[j]] : -1;
				sbase += 6 + 2 * CB_DOUBLES + 1; ibase += 3;
				break;
			}
			case DK_VI: {                      /* vicurve.c:333-356 */
				const int s[8] = {slot(R, K), slot(R, S), slot(K, R), slot(S, R),
					slot(J, J), slot(J, S), slot(S, J), slot(S, S)};
				static const double pm[4] = {1.0, -1.0, 1.0, -1.0};
				for (int j = 0; j < 8; j++) {
					rec[DF_VA0 + j] = s[j]; both(s[j]);
					if (j < 4) lit(s[j], pm[j], pm[j]); else ref(s[j]);
				}
				sbase += 4 + CB_DOUBLES; ibase += 4;
				break;
			}
			case DK_BI: case DK_BV: case DK_BC: {
				rec[DF_CALC] = (int)p.calcOff.size() - 1;
				rec[DF_USETIME] = d.usesTime ? 1 : 0;
				rec[DF_NBV] = (int)d.bvRow.size();
				rec[DF_BVOFF] = (int)p.bvar.size() / 4;
				const int base = (int)p.calcConst.size();
				for (size_t i = 0; i < d.calc.code.size(); i += 2) {
					const int op = d.calc.code[i], arg = d.calc.code[i + 1];
					p.calcCode.push_back(op);
					p.calcCode.push_back(op == OP_CONST ? arg + base : arg);
				}
				p.calcConst.insert(p.calcConst.end(), d.calc.consts.begin(), d.calc.consts.end());
				p.calcOff.push_back((int)p.calcCode.size());
				for (size_t i = 0; i < d.calc.varRef.size(); i++) p.calcVars.push_back(d.calc.varRef[i]);
				p.calcVarOff.push_back((int)p.calcVars.size());
				p.maxCalcStack = std::max(p.maxCalcStack, d.calc.maxStack);
				for (size_t i = 0; i < d.bvRow.size(); i++) {
					const int x = d.bvRow[i];
					int sA, sB = -1;
					if (d.kind == DK_BI) { sA = slot(J, x); sB = slot(S, x); }   /* nonlinear_i.c:163-188 */
					else sA = slot(R, x);                                        /* _v.c:134-156, _c.c:142-164 */
					both(sA); both(sB); ref(sA); ref(sB);
					p.bvar.push_back(x); p.bvar.push_back(sA); p.bvar.push_back(sB);
					p.bvar.push_back(d.bvCalcVar[i]);
				}
				if (d.kind == DK_BI) {         /* nonlinear_i.c:402-417: RK RM KR MR */
					const int s[4] = {slot(R, K), slot(R, S), slot(K, R), slot(S, R)};
					static const double pm[4] = {1.0, -1.0, 1.0, -1.0};
					for (int j = 0; j < 4; j++) { rec[DF_BA0 + j] = s[j]; both(s[j]); lit(s[j], pm[j], pm[j]); }
					sbase += 4 + CB_DOUBLES + (int)d.bvRow.size(); ibase += 2;
				} else if (d.kind == DK_BV) {  /* nonlinear_v.c:371-384: RK RJ KR JR */
					const int s[4] = {slot(R, K), slot(R, J), slot(K, R), slot(J, R)};
					static const double pm[4] = {1.0, -1.0, 1.0, -1.0};
					for (int j = 0; j < 4; j++) { rec[DF_BA0 + j] = s[j]; both(s[j]); lit(s[j], pm[j], pm[j]); }
					sbase += 4 + CB_DOUBLES + (int)d.bvRow.size(); ibase += 2;
				} else {                       /* nonlinear_c.c:465-485: KK KJ JK JJ RC CR */
					const int s[6] = {slot(K, K), slot(K, J), slot(J, K), slot(J, J),
						slot(R, S), slot(S, R)};
					for (int j = 0; j < 4; j++) { rec[DF_BA0 + j] = s[j]; mark(p.activeTran, s[j]); ref(s[j]); }
					for (int j = 4; j < 6; j++) { rec[DF_BA0 + j] = s[j]; both(s[j]); lit(s[j], 1.0, 1.0); }
					sbase += 2 + INTEG_DOUBLES + 4 + (int)d.bvRow.size(); ibase += 2;
				}
				break;
			}
			default: err = "unknown device kind"; return false;
			}
		}
		p.stateDoubles = sbase; p.stateInts = ibase;
		for (int sl = 0; sl < p.nnzA; sl++)
			if (refs[sl] != 1) { p.constOP[sl] = NAN; p.constTran[sl] = NAN; }
		return parameters(true);
	}
};

}  /* namespace */

bool compileProgram(const Netlist &nl, int M, Program &p, std::string &err)
{
	Flattener f(nl, M, p);
	if (!f.run()) { err = f.err; return false; }
	return true;
}

bool refreshParameters(const Netlist &nl, int M, Program &p, std::string &err)
{
	Flattener f(nl, M, p);
	const size_t nparams = p.pmap.size(), ntab = p.tabType.size();
	if (!f.parameters(true)) { err = f.err; return false; }
	if (p.pmap.size() != nparams || p.tabType.size() != ntab) {
		err = "netlist changed after compilation";
		return false;
	}
	return true;
}

/* ========================================================================== *
 * Symbolic LU analysis.
 *
 * The reference calls SuperLU's dgssvx with partial pivoting (u = 1.0) on every
 * Newton iteration (core/matrix.c:82-150, libs/superlu/SRC/dpivotL.c:99-147).
 * The batched engine fixes one (column, pivot row) sequence per phase for all
 * instances.  It is chosen here on pilot matrices: at each step the candidate
 * with the smallest Markowitz count (r-1)(c-1) among those whose magnitude is
 * at least `threshold` times the largest entry of their column, in EVERY
 * pilot; ties go to the numerically stronger candidate.  A sequence obtained
 * elsewhere (e.g. perm_c / perm_r of libs/superlu) can be
 * forced instead.
 * ========================================================================== */
bool analyse(int N, const std::vector<int> &colStart, const std::vector<int> &rowIdx,
		const std::vector<char> &active,
		const std::vector<std::vector<double> > &pilots, double threshold,
		const int *forcePc, const int *forcePr, Schedule &out, std::string &err)
{
	const int nnzA = (int)rowIdx.size();
	const int P = (int)pilots.size();
	out = Schedule();
	std::vector<int> slotAt((size_t)N * N, -1);     /* [r][c] -> slot */
	std::vector<char> S((size_t)N * N, 0);          /* structurally non-zero */
	for (int c = 0; c < N; c++)
		for (int s = colStart[c]; s < colStart[c + 1]; s++) {
			slotAt[(size_t)rowIdx[s] * N + c] = s;
			S[(size_t)rowIdx[s] * N + c] = active.empty() ? 1 : active[s];
		}

	out.pc.assign(N, -1); out.pr.assign(N, -1);
	if (forcePc && forcePr) {
		for (int k = 0; k < N; k++) { out.pc[k] = forcePc[k]; out.pr[k] = forcePr[k]; }
	} else {
		if (P < 1) { err = "symbolic analysis needs a pilot matrix"; return false; }
		std::vector<std::vector<double> > W(P, std::vector<double>((size_t)N * N, 0.0));
		for (int q = 0; q < P; q++)
			for (int c = 0; c < N; c++)
				for (int s = colStart[c]; s < colStart[c + 1]; s++)
					W[q][(size_t)rowIdx[s] * N + c] = pilots[q][s];
		std::vector<char> T(S), rdone(N, 0), cdone(N, 0);
		std::vector<double> colMax((size_t)P * N);
		for (int k = 0; k < N; k++) {
			std::vector<int> rcount(N, 0), ccount(N, 0);
			for (int r = 0; r < N; r++) {
				if (rdone[r]) continue;
				for (int c = 0; c < N; c++)
					if (!cdone[c] && T[(size_t)r * N + c]) { rcount[r]++; ccount[c]++; }
			}
			for (int q = 0; q < P; q++)
				for (int c = 0; c < N; c++) {
					double m = 0.0;
					for (int r = 0; r < N; r++)
						if (!rdone[r]) m = std::max(m, std::fabs(W[q][(size_t)r * N + c]));
					colMax[(size_t)q * N + c] = m;
				}
			int br = -1, bc = -1; long bestCost = 0; double bestRatio = -1.0; bool bestOk = false;
			for (int c = 0; c < N; c++) {
				if (cdone[c]) continue;
				for (int r = 0; r < N; r++) {
					if (rdone[r] || !T[(size_t)r * N + c]) continue;
					double ratio = 1.0;
					for (int q = 0; q < P; q++) {
						const double m = colMax[(size_t)q * N + c];
						const double v = std::fabs(W[q][(size_t)r * N + c]);
						ratio = std::min(ratio, (m > 0.0 && std::isfinite(v)) ? v / m : 0.0);
					}
					const bool ok = ratio >= threshold;
					const long cost = (long)(rcount[r] - 1) * (ccount[c] - 1);
					bool better;
					if (br < 0) better = true;
					else if (ok != bestOk) better = ok;
					else if (ok) better = (cost < bestCost) || (cost == bestCost && ratio > bestRatio);
					else better = ratio > bestRatio;
					if (better) { br = r; bc = c; bestCost = cost; bestRatio = ratio; bestOk = ok; }
				}
			}
			if (br < 0 || bestRatio <= 0.0) {
				err = "matrix is structurally or numerically singular (no pivot at step "
					+ std::to_string(k) + ")";
				return false;
			}
			out.pc[k] = bc; out.pr[k] = br;
			out.worstRatio = std::min(out.worstRatio, bestRatio);
			rdone[br] = 1; cdone[bc] = 1;
			for (int q = 0; q < P; q++) {
				std::vector<double> &w = W[q];
				const double inv = 1.0 / w[(size_t)br * N + bc];
				for (int r = 0; r < N; r++) {
					if (rdone[r]) continue;
					const double l = w[(size_t)r * N + bc] * inv;
					if (l == 0.0) continue;
					for (int c = 0; c < N; c++)
						if (!cdone[c]) w[(size_t)r * N + c] -= l * w[(size_t)br * N + c];
				}
			}
			for (int r = 0; r < N; r++) {
				if (rdone[r] || !T[(size_t)r * N + bc]) continue;
				for (int c = 0; c < N; c++)
					if (!cdone[c] && T[(size_t)br * N + c]) T[(size_t)r * N + c] = 1;
			}
		}
	}

	/* validate the sequence */
	{
		std::vector<char> seenR(N, 0), seenC(N, 0);
		for (int k = 0; k < N; k++) {
			const int c = out.pc[k], r = out.pr[k];
			if (c < 0 || c >= N || r < 0 || r >= N || seenC[c] || seenR[r]) {
				err = "pivot sequence is not a permutation";
				return false;
			}
			seenC[c] = 1; seenR[r] = 1;
		}
	}

	/* symbolic replay: fill slots, then the op-stream */
	std::vector<int> stepOfCol(N), rowOfCol(N);
	for (int k = 0; k < N; k++) { stepOfCol[out.pc[k]] = k; rowOfCol[out.pc[k]] = out.pr[k]; }
	int F = nnzA;
	struct Step { std::vector<int> ucols, lrows; };
	std::vector<Step> steps(N);
	{
		std::vector<char> rdone(N, 0), cdone(N, 0);
		for (int k = 0; k < N; k++) {
			const int c = out.pc[k], pr = out.pr[k];
			if (!S[(size_t)pr * N + c]) {
				err = "pivot (" + std::to_string(pr) + "," + std::to_string(c) +
					") is structurally zero";
				return false;
			}
			rdone[pr] = 1; cdone[c] = 1;
			Step &st = steps[k];
			for (int c2 = 0; c2 < N; c2++)
				if (!cdone[c2] && S[(size_t)pr * N + c2]) st.ucols.push_back(c2);
			std::sort(st.ucols.begin(), st.ucols.end(),
					[&](int a, int b) { return stepOfCol[a] < stepOfCol[b]; });
			for (int r = 0; r < N; r++)
				if (!rdone[r] && S[(size_t)r * N + c]) st.lrows.push_back(r);
			for (size_t i = 0; i < st.lrows.size(); i++)
				for (size_t j = 0; j < st.ucols.size(); j++) {
					const size_t at = (size_t)st.lrows[i] * N + st.ucols[j];
					S[at] = 1;
					if (slotAt[at] < 0) slotAt[at] = F++;
				}
		}
	}
	out.F = F;
	std::vector<int> back(N);
	for (int k = 0; k < N; k++) {
		const int c = out.pc[k], pr = out.pr[k];
		const Step &st = steps[k];
		const int nU = (int)st.ucols.size(), nL = (int)st.lrows.size();
		back[k] = (int)out.lu.size();
		out.lu.push_back(slotAt[(size_t)pr * N + c]);
		out.lu.push_back(F + pr);
		out.lu.push_back(nU);
		out.lu.push_back(nL);
		for (int j = 0; j < nU; j++) {
			out.lu.push_back(slotAt[(size_t)pr * N + st.ucols[j]]);
			out.lu.push_back(F + rowOfCol[st.ucols[j]]);
		}
		for (int i = 0; i < nL; i++) {
			const int r = st.lrows[i];
			out.lu.push_back(slotAt[(size_t)r * N + c]);
			out.lu.push_back(F + r);
			for (int j = 0; j < nU; j++) out.lu.push_back(slotAt[(size_t)r * N + st.ucols[j]]);
		}
		out.nU += nU; out.nL += nL;
		out.flops += (long long)nL * (nU + 2) + nU + 1;
	}
	out.lu.insert(out.lu.end(), back.begin(), back.end());
	out.xslot.assign(N, 0);
	for (int c = 0; c < N; c++) out.xslot[c] = F + rowOfCol[c];
	return true;
}

}  /* namespace simcu */
