/*
 * simcu_jit.cpp - compiles the topology-specialised translation unit with
 * NVRTC for sm_100a and loads the cubin through the CUDA runtime's library API.
 * libnvrtc is opened with dlopen at the first compile, so libsimulator_cuda.so
 * itself links against nothing but the (static) CUDA runtime and loads on a
 * machine without a driver; compiling needs no GPU either, loading does.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "simcu_codegen.h"

namespace simcu {

namespace {

/* simcu_program.h, simcu_calc.h, simcu_device.cuh, simcu_jit_kernel.cuh as raw
 * string literals (written by the Makefile, #include lines removed) */
#include "simcu_embed.inc"

struct Nvrtc {
	void *handle = nullptr;
	std::string path;
	decltype(&nvrtcCreateProgram) createProgram = nullptr;
	decltype(&nvrtcCompileProgram) compileProgram = nullptr;
	decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
	decltype(&nvrtcGetProgramLogSize) getLogSize = nullptr;
	decltype(&nvrtcGetProgramLog) getLog = nullptr;
	decltype(&nvrtcGetCUBINSize) getCubinSize = nullptr;
	decltype(&nvrtcGetCUBIN) getCubin = nullptr;
	decltype(&nvrtcGetErrorString) errorString = nullptr;
};
Nvrtc gNvrtc;
std::string gNvrtcHint;

bool openNvrtc(std::string &err)
{
	if (gNvrtc.handle) return true;
	std::vector<std::string> names;
	if (!gNvrtcHint.empty()) names.push_back(gNvrtcHint);
	if (const char *e = std::getenv("SIMCU_NVRTC")) names.push_back(e);
	/* The toolkit's own NVRTC first, by PATH: a process that has imported PyTorch
	 * already holds torch's bundled libnvrtc.so.12 (an older 12.x), which a bare
	 * soname lookup would return - measured 6 % (cfg5) to 30 % (cfg4) slower code
	 * for this kernel than the 12.9 compiler. */
	if (const char *e = std::getenv("CUDA_HOME")) names.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
	names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
	names.push_back("/usr/local/cuda/lib64/libnvrtc.so");
	names.push_back("libnvrtc.so.12");
	names.push_back("libnvrtc.so");
	for (size_t i = 0; i < names.size() && !gNvrtc.handle; i++) {
		gNvrtc.handle = dlopen(names[i].c_str(), RTLD_NOW | RTLD_LOCAL);
		if (gNvrtc.handle) gNvrtc.path = names[i];
	}
	if (!gNvrtc.handle) { err = "cannot load libnvrtc (set SIMCU_NVRTC to its path)"; return false; }
#define SYM(member, name) \
	gNvrtc.member = (decltype(gNvrtc.member))dlsym(gNvrtc.handle, name); \
	if (!gNvrtc.member) { err = std::string("libnvrtc lacks ") + name; return false; }
	SYM(createProgram, "nvrtcCreateProgram")
	SYM(compileProgram, "nvrtcCompileProgram")
	SYM(destroyProgram, "nvrtcDestroyProgram")
	SYM(getLogSize, "nvrtcGetProgramLogSize")
	SYM(getLog, "nvrtcGetProgramLog")
	SYM(getCubinSize, "nvrtcGetCUBINSize")
	SYM(getCubin, "nvrtcGetCUBIN")
	SYM(errorString, "nvrtcGetErrorString")
#undef SYM
	{
		typedef nvrtcResult (*versionFn)(int *, int *);
		versionFn v = (versionFn)dlsym(gNvrtc.handle, "nvrtcVersion");
		int major = 0, minor = 0;
		if (v && v(&major, &minor) == NVRTC_SUCCESS) {
			char buf[32];
			std::snprintf(buf, sizeof buf, " (%d.%d)", major, minor);
			gNvrtc.path += buf;
		}
	}
	return true;
}

}  /* namespace */

const char *jitNvrtcPath() { return gNvrtc.path.c_str(); }
void jitSetNvrtcPath(const char *path) { gNvrtcHint = path ? path : ""; }

std::string assembleSource(const Program &p, const Schedule sched[2],
		const std::vector<char> active[2], const std::vector<double> consts[2],
		const std::vector<int> &probeRows, const std::vector<int> &measures,
		const JitConfig &cfg)
{
	std::string s = generatePrelude(p, (int)probeRows.size(), (int)measures.size() / 2, cfg);
	s += kSrcProgram;
	s += kSrcCalc;
	s += kSrcDevice;
	s += generateTopology(p, sched, active, consts, probeRows, measures, cfg);
	s += kSrcJitKernel;
	return s;
}

JitKernel::~JitKernel()
{
	if (library) cudaLibraryUnload((cudaLibrary_t)library);
}

namespace {

/* Optional cubin cache on disk (environment variable SIMCU_CACHE_DIR, off when
 * unset): NVRTC takes 1-7 s per topology, which a process pays again every time
 * it starts.  The key is the whole translation unit, the contraction flag and
 * the build stamp of this library. */
std::string cachePathOf(const std::string &source, bool fmad)
{
	const char *dir = std::getenv("SIMCU_CACHE_DIR");
	if (!dir || !*dir) return std::string();
	unsigned long long h = 1469598103934665603ULL;
	const std::string stamp = std::string(__DATE__ " " __TIME__) + (fmad ? " fmad" : " nofmad");
	for (size_t i = 0; i < source.size(); i++) { h ^= (unsigned char)source[i]; h *= 1099511628211ULL; }
	for (size_t i = 0; i < stamp.size(); i++) { h ^= (unsigned char)stamp[i]; h *= 1099511628211ULL; }
	char name[64];
	std::snprintf(name, sizeof name, "/simcu_%016llx_%zu.cubin", h, source.size());
	return std::string(dir) + name;
}

}  /* namespace */

bool jitCompile(const std::string &source, JitKernel &k, std::string &err, bool fmad)
{
	const std::string cached = cachePathOf(source, fmad);
	if (!cached.empty()) {
		if (FILE *f = std::fopen(cached.c_str(), "rb")) {
			std::fseek(f, 0, SEEK_END);
			const long n = std::ftell(f);
			std::fseek(f, 0, SEEK_SET);
			if (n > 0) {
				k.cubin.resize((size_t)n);
				if (std::fread(k.cubin.data(), 1, (size_t)n, f) != (size_t)n) k.cubin.clear();
			}
			std::fclose(f);
			if (!k.cubin.empty()) { k.log = "cubin from " + cached; return true; }
		}
	}
	if (!openNvrtc(err)) return false;
	std::string name = "simcu_topology.cu";
	if (const char *dir = std::getenv("SIMCU_JIT_DUMP")) {
		char file[64];
		std::snprintf(file, sizeof file, "/simcu_topology_%016zx.cu", std::hash<std::string>()(source));
		name = std::string(dir) + file;
		if (FILE *f = std::fopen(name.c_str(), "w")) { std::fwrite(source.data(), 1, source.size(), f); std::fclose(f); }
	}
	nvrtcProgram prog = nullptr;
	nvrtcResult rc = gNvrtc.createProgram(&prog, source.c_str(), name.c_str(), 0, nullptr, nullptr);
	if (rc != NVRTC_SUCCESS) { err = gNvrtc.errorString(rc); return false; }
	std::vector<const char *> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
		fmad ? "--fmad=true" : "--fmad=false"};
	rc = gNvrtc.compileProgram(prog, (int)opts.size(), opts.data());
	size_t logSize = 0;
	gNvrtc.getLogSize(prog, &logSize);
	if (logSize > 1) {
		k.log.resize(logSize);
		gNvrtc.getLog(prog, &k.log[0]);
	}
	if (rc != NVRTC_SUCCESS) {
		err = std::string("NVRTC: ") + gNvrtc.errorString(rc) + "\n" + k.log;
		gNvrtc.destroyProgram(&prog);
		return false;
	}
	size_t n = 0;
	gNvrtc.getCubinSize(prog, &n);
	k.cubin.resize(n);
	gNvrtc.getCubin(prog, k.cubin.data());
	gNvrtc.destroyProgram(&prog);
	if (!cached.empty()) {          /* write-then-rename: concurrent ranks compile the same source */
		char tmp[32];
		std::snprintf(tmp, sizeof tmp, ".%ld.tmp", (long)getpid());
		const std::string part = cached + tmp;
		if (FILE *f = std::fopen(part.c_str(), "wb")) {
			const bool ok = std::fwrite(k.cubin.data(), 1, k.cubin.size(), f) == k.cubin.size();
			std::fclose(f);
			if (!ok || std::rename(part.c_str(), cached.c_str()) != 0) std::remove(part.c_str());
		}
	}
	return true;
}

bool jitLoad(JitKernel &k, int threadsPerBlock, std::string &err)
{
	cudaLibrary_t lib = nullptr;
	cudaError_t e = cudaLibraryLoadData(&lib, k.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
	if (e != cudaSuccess) { err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e); return false; }
	k.library = lib;
	cudaKernel_t kern = nullptr;
	e = cudaLibraryGetKernel(&kern, lib, "simcuJitKernel");
	if (e != cudaSuccess) { err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e); return false; }
	k.kernel = kern;
	/* no shared memory is used: give the whole SM array to L1, where the
	 * spilled part of the thread-local state lives */
	if (cudaFuncSetAttribute((const void *)kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0)
			!= cudaSuccess)
		cudaGetLastError();
	cudaFuncAttributes attr;
	if (cudaFuncGetAttributes(&attr, (const void *)kern) == cudaSuccess) {
		k.numRegs = attr.numRegs;
		k.localBytes = attr.localSizeBytes;
	} else cudaGetLastError();
	int blocks = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, (const void *)kern, threadsPerBlock, 0)
			== cudaSuccess && blocks > 0) {
		k.maxBlocksPerSM = blocks;
	} else {
		cudaGetLastError();
		const int regs = k.numRegs > 0 ? k.numRegs : 255;
		const int perBlock = ((regs * 32 + 255) / 256 * 256) * (threadsPerBlock / 32);
		k.maxBlocksPerSM = std::max(1, std::min(32, 65536 / std::max(1, perBlock)));
	}
	return true;
}

}  /* namespace simcu */
