/*
 * simcu_comm.cpp - the one collective of the engine: gather the finished probe
 * waveforms of all ranks to the root GPU over NCCL (NVLink 5 / NVSwitch) and
 * copy them to the host once.  Instances are independent, so
 * nothing is exchanged while stepping.
 *
 * One process per GPU.  The sweep is dealt out round-robin: global instance g
 * lives on rank g % nranks as local instance g / nranks, so every rank gets the
 * same mix of IBIS corners / parameter ranges and the ranks finish together.
 * Every rank transposes its own [grid][probe][M] block to [M][grid][probe] on
 * its GPU (simcuTransposeKernel), the non-root ranks ncclSend it, the root
 * ncclRecv's all blocks into one device buffer (grouped, so all peers stream in
 * concurrently through the switch) and de-interleaves them into the host array
 * with one strided DMA per rank - no host bounce on the senders, no gather to
 * ranks that do not need the data.
 *
 * libnccl is opened with dlopen at the first use (the copy PyTorch has already
 * loaded is found by its soname), so libsimulator_cuda.so itself still links
 * against nothing but the static CUDA runtime.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/simulator_cuda.h"
#include "simcu_kernels.h"

namespace {

struct Nccl {
	void *handle = nullptr;
	decltype(&ncclGetUniqueId) getUniqueId = nullptr;
	decltype(&ncclCommInitRank) commInitRank = nullptr;
	decltype(&ncclCommDestroy) commDestroy = nullptr;
	decltype(&ncclGroupStart) groupStart = nullptr;
	decltype(&ncclGroupEnd) groupEnd = nullptr;
	decltype(&ncclSend) send = nullptr;
	decltype(&ncclRecv) recv = nullptr;
	decltype(&ncclGetErrorString) errorString = nullptr;
	decltype(&ncclGetVersion) getVersion = nullptr;
};
Nccl gNccl;

int commFail(const char *what, const char *why)
{
	std::fprintf(stderr, "simulator_cuda: %s: %s\n", what, why);
	return -1;
}

int openNccl()
{
	if (gNccl.handle) return 0;
	std::vector<std::string> names;
	if (const char *e = std::getenv("SIMCU_NCCL")) names.push_back(e);
	names.push_back("libnccl.so.2");
	names.push_back("libnccl.so");
	for (size_t i = 0; i < names.size() && !gNccl.handle; i++)
		gNccl.handle = dlopen(names[i].c_str(), RTLD_NOW | RTLD_GLOBAL);
	if (!gNccl.handle) return commFail("NCCL", "cannot load libnccl (set SIMCU_NCCL to its path)");
#define SYM(member, name) \
	gNccl.member = (decltype(gNccl.member))dlsym(gNccl.handle, name); \
	if (!gNccl.member) return commFail("libnccl lacks", name);
	SYM(getUniqueId, "ncclGetUniqueId")
	SYM(commInitRank, "ncclCommInitRank")
	SYM(commDestroy, "ncclCommDestroy")
	SYM(groupStart, "ncclGroupStart")
	SYM(groupEnd, "ncclGroupEnd")
	SYM(send, "ncclSend")
	SYM(recv, "ncclRecv")
	SYM(errorString, "ncclGetErrorString")
	SYM(getVersion, "ncclGetVersion")
#undef SYM
	return 0;
}

#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
	return commFail(#call, gNccl.errorString(r_)); } while (0)
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return commFail(#call, cudaGetErrorString(e_)); } while (0)

}  /* namespace */

struct _simcuComm {
	ncclComm_t comm = nullptr;
	int nranks = 1, rank = 0;
	void *dGather = nullptr;      /* root: [nranks][M][rows] doubles, then [nranks][M] ints */
	size_t gatherBytes = 0;
	int *hStatus = nullptr;       /* root, pinned: [nranks][M] */
	size_t hStatusBytes = 0;
	cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t evReady = nullptr;   /* compute stream: the transposed block is complete */
	cudaStream_t side = nullptr;     /* NCCL + copies run here, beside the next batch */
	/* the gather in flight (simcuGatherOutputAsync ... simcuGatherWait) */
	bool pending = false, pendingRoot = false;
	int pendingM = 0;
	int *pendingStatus = nullptr;
	long long pendingNccl = 0, pendingCopy = 0;
};

extern "C" {

int simcuCommVersion(void)
{
	int v = 0;
	if (openNccl()) return -1;
	return (gNccl.getVersion(&v) == ncclSuccess) ? v : -1;
}

int simcuCommUniqueId(char *id, int bytes)
{
	if (!id || bytes < (int)sizeof(ncclUniqueId)) return commFail("simcuCommUniqueId", "needs a 128-byte buffer");
	if (openNccl()) return -1;
	ncclUniqueId u;
	NC(gNccl.getUniqueId(&u));
	std::memcpy(id, &u, sizeof u);
	return 0;
}

int simcuCommInit(simcuComm_ **out, int nranks, int rank, const char *id, int bytes)
{
	if (!out || !id || bytes < (int)sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
		return commFail("simcuCommInit", "bad arguments");
	if (openNccl()) return -1;
	ncclUniqueId u;
	std::memcpy(&u, id, sizeof u);
	simcuComm_ *c = new _simcuComm();
	c->nranks = nranks; c->rank = rank;
	ncclResult_t rc = gNccl.commInitRank(&c->comm, nranks, u, rank);
	if (rc != ncclSuccess) { delete c; return commFail("ncclCommInitRank", gNccl.errorString(rc)); }
	for (int i = 0; i < 3; i++) cudaEventCreate(&c->ev[i]);
	cudaEventCreateWithFlags(&c->evReady, cudaEventDisableTiming);
	cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
	*out = c;
	return 0;
}

int simcuCommDestroy(simcuComm_ **c)
{
	if (!c || !*c) return -1;
	if ((*c)->dGather) cudaFree((*c)->dGather);
	if ((*c)->hStatus) cudaFreeHost((*c)->hStatus);
	for (int i = 0; i < 3; i++) if ((*c)->ev[i]) cudaEventDestroy((*c)->ev[i]);
	if ((*c)->evReady) cudaEventDestroy((*c)->evReady);
	if ((*c)->side) cudaStreamDestroy((*c)->side);
	if ((*c)->comm) gNccl.commDestroy((*c)->comm);
	delete *c;
	*c = nullptr;
	return 0;
}

int simcuGatherOutputAsync(simcu_ *r, simcuComm_ *c, int root, double *hostData, int *hostStatus,
		void *stream)
{
	if (!r || !c || root < 0 || root >= c->nranks) return commFail("simcuGatherOutput", "bad arguments");
	if (c->pending) return commFail("simcuGatherOutputAsync", "the previous gather has not been waited for");
	cudaStream_t st = c->side;
	int M = 0, rows = 0;
	const double *block = nullptr;
	const int *status = nullptr;
	/* transpose on the caller's (compute) stream; everything after it runs on the
	 * side stream, so the caller can launch its next batch at once */
	if (simcuTransposeOutput(r, stream, (void **)&block, (void **)&status, &M, &rows)) return -1;
	CUC(cudaEventRecord(c->evReady, (cudaStream_t)stream));
	CUC(cudaStreamWaitEvent(c->side, c->evReady, 0));
	const size_t count = (size_t)M * rows, blockBytes = count * 8;
	const int W = c->nranks;
	const bool isRoot = (c->rank == root);
	if (isRoot) {
		const size_t need = blockBytes * W + (size_t)M * 4 * W;
		if (need > c->gatherBytes) {
			if (c->dGather) cudaFree(c->dGather);
			c->dGather = nullptr; c->gatherBytes = 0;
			CUC(cudaMalloc(&c->dGather, need));
			c->gatherBytes = need;
		}
		if ((size_t)M * 4 * W > c->hStatusBytes) {
			if (c->hStatus) cudaFreeHost(c->hStatus);
			c->hStatus = nullptr; c->hStatusBytes = 0;
			CUC(cudaMallocHost((void **)&c->hStatus, (size_t)M * 4 * W));
			c->hStatusBytes = (size_t)M * 4 * W;
		}
	}
	CUC(cudaEventRecord(c->ev[0], st));
	if (W > 1) {
		NC(gNccl.groupStart());
		if (isRoot) {
			char *base = (char *)c->dGather;
			int *sbase = (int *)(base + blockBytes * W);
			for (int p = 0; p < W; p++) {
				if (p == root) continue;
				if (count) NC(gNccl.recv(base + blockBytes * p, count, ncclDouble, p, c->comm, st));
				NC(gNccl.recv(sbase + (size_t)M * p, (size_t)M, ncclInt32, p, c->comm, st));
			}
		} else {
			if (count) NC(gNccl.send(block, count, ncclDouble, root, c->comm, st));
			NC(gNccl.send(status, (size_t)M, ncclInt32, root, c->comm, st));
		}
		NC(gNccl.groupEnd());
	}
	CUC(cudaEventRecord(c->ev[1], st));
	if (isRoot) {
		/* de-interleave while copying out: global instance g = local * W + rank */
		const size_t rowBytes = (size_t)rows * 8;
		char *base = (char *)c->dGather;
		const int *sbase = (const int *)(base + blockBytes * W);
		for (int p = 0; p < W && hostData && rowBytes; p++) {
			const void *src = (p == root) ? (const void *)block : (const void *)(base + blockBytes * p);
			CUC(cudaMemcpy2DAsync((char *)hostData + rowBytes * p, rowBytes * W, src, rowBytes, rowBytes,
					(size_t)M, cudaMemcpyDeviceToHost, st));
		}
		if (hostStatus) {
			for (int p = 0; p < W; p++) {
				const int *src = (p == root) ? status : sbase + (size_t)M * p;
				CUC(cudaMemcpyAsync(c->hStatus + (size_t)M * p, src, (size_t)M * 4, cudaMemcpyDeviceToHost, st));
			}
		}
	}
	CUC(cudaEventRecord(c->ev[2], st));
	c->pending = true; c->pendingRoot = isRoot; c->pendingM = M;
	c->pendingStatus = isRoot ? hostStatus : nullptr;
	c->pendingNccl = isRoot ? (long long)(blockBytes + (size_t)M * 4) * (W - 1)
		: (W > 1 ? (long long)(blockBytes + (size_t)M * 4) : 0);
	c->pendingCopy = isRoot ? (long long)(blockBytes + (size_t)M * 4) * W : 0;
	return 0;
}

int simcuGatherWait(simcuComm_ *c, simcuGatherStats_ *stats)
{
	if (!c) return -1;
	if (!c->pending) return 0;
	CUC(cudaEventSynchronize(c->ev[2]));
	c->pending = false;
	const int W = c->nranks, M = c->pendingM;
	if (c->pendingRoot && c->pendingStatus)
		for (int p = 0; p < W; p++)
			for (int i = 0; i < M; i++) c->pendingStatus[(size_t)i * W + p] = c->hStatus[(size_t)M * p + i];
	if (stats) {
		float a = 0.f, b = 0.f;
		cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
		cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
		stats->ncclMs = a; stats->copyMs = b;
		stats->ncclBytes = c->pendingNccl;
		stats->copyBytes = c->pendingCopy;
	}
	return 0;
}

int simcuGatherOutput(simcu_ *r, simcuComm_ *c, int root, double *hostData, int *hostStatus,
		void *stream, simcuGatherStats_ *stats)
{
	if (simcuGatherOutputAsync(r, c, root, hostData, hostStatus, stream)) return -1;
	return simcuGatherWait(c, stats);
}

}  /* extern "C" */
