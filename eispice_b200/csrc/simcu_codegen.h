/*
 * simcu_codegen.h - source generation for the topology-specialised kernel
 * (simcu_codegen.cpp) and its NVRTC compilation / loading (simcu_jit.cpp).
 */
#ifndef SIMCU_CODEGEN_H
#define SIMCU_CODEGEN_H

#include <string>
#include <vector>

#include "simcu_host.h"

namespace simcu {

/* compile-time choices of one specialised kernel besides the topology itself */
struct JitConfig {
	int threadsPerBlock = 64;
	int minBlocks = 0;        /* __launch_bounds__ second argument (0 = 1) */
	double gmin = 1e-15;      /* guard of every calculon division, folded into the code */
	int replay = 0;           /* 1: replay an attempt log instead of own step control (tests) */
	int blockSync = 0;        /* 1: the warps of a block start every attempt together */
	int syncGroup = 0;        /* > 0: ... in groups of this many warps (named barriers) */
	int windows = 0;          /* 1: per-instance analysis windows (tstep, tstop, tmax, minstep) */
	int smemTab = 0;          /* 1: hot tables staged in shared memory */
	double growthLimit = 1e6; /* LU multipliers above this are counted (pivot growth guard; <= 0: off) */
	int tlinePrefetch = 0;    /* circuits with >= 2 T-lines: locate + prefetch pass before the step */
	int fmad = 0;             /* 0: no a*b+c contraction: the reference's (and the oracle's) arithmetic */
	int solveBlocks = 2;      /* LU + solves emitted block by block (1), blocks no linearize stamp reaches
	                           * only in the first solve of a Newton loop (2); 0: all at once */
	int sharedCb = 1;         /* one break-point clock per instance instead of one per device */
	bool operator==(const JitConfig &o) const
	{
		return threadsPerBlock == o.threadsPerBlock && minBlocks == o.minBlocks && gmin == o.gmin &&
			replay == o.replay && blockSync == o.blockSync && syncGroup == o.syncGroup && windows == o.windows && smemTab == o.smemTab && growthLimit == o.growthLimit &&
			tlinePrefetch == o.tlinePrefetch && fmad == o.fmad && solveBlocks == o.solveBlocks &&
			sharedCb == o.sharedCb;
	}
};

/* #defines of the topology's sizes; first part of the translation unit */
std::string generatePrelude(const Program &p, int numProbes, int numMeasures, const JitConfig &cfg);
/* device records, expressions and the unrolled LU of both phases */
std::string generateTopology(const Program &p, const Schedule sched[2],
		const std::vector<char> active[2], const std::vector<double> consts[2],
		const std::vector<int> &probeRows, const std::vector<int> &measures, const JitConfig &cfg);
/* prelude + embedded library headers + topology + kernel */
std::string assembleSource(const Program &p, const Schedule sched[2],
		const std::vector<char> active[2], const std::vector<double> consts[2],
		const std::vector<int> &probeRows, const std::vector<int> &measures,
		const JitConfig &cfg);

/* A compiled kernel.  compile() needs no GPU (NVRTC only); load() does. */
struct JitKernel {
	std::vector<char> cubin;
	std::string log;
	void *library = nullptr;      /* cudaLibrary_t */
	void *kernel = nullptr;       /* cudaKernel_t */
	int numRegs = 0;
	size_t localBytes = 0;
	int maxBlocksPerSM = 0;
	~JitKernel();
};
/* Returns false and fills err on failure.  If the environment variable
 * SIMCU_JIT_DUMP names a directory, the translation unit is also written there
 * (and compiled under that path, so profilers can show the source). */
bool jitCompile(const std::string &source, JitKernel &k, std::string &err, bool fmad = true);
bool jitLoad(JitKernel &k, int threadsPerBlock, std::string &err);
/* where libnvrtc was found ("" before the first compile) */
const char *jitNvrtcPath();
void jitSetNvrtcPath(const char *path);

}  /* namespace simcu */
#endif
