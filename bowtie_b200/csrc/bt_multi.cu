/*
 * bt_multi.cu — the path's only collective: the sum of the five summary counters over the ranks.
 *
 * The reference keeps one set of counters per process (HitSink: numAligned, numUnaligned, numMaxed, numReported,
 * numReportedPaired; hit.h:169-175, printed by hit.h:303-337) because all its worker threads share a sink.  With one process per
 * GPU and the reads sharded across them, the summary line needs their sum: one ncclAllReduce over five 64-bit words.  NCCL is
 * bound at call time (dlopen) so that single-GPU users do not need it installed; the communicator is the caller's.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string>
#include "../../include/bowtie_b200.h"

int bt_internal_fail(const std::string &m);

typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);   /* ncclAllReduce(sendbuff, recvbuff, count, datatype, op, comm, stream) */
typedef const char *(*nccl_errstr_fn)(int);

extern "C" int bt_counters_allreduce(void *nccl_comm, uint64_t counters[5], void *stream) {
	if (!nccl_comm || !counters) return bt_internal_fail("bt_counters_allreduce: null argument");
	static void *lib = nullptr;
	static nccl_allreduce_fn allreduce = nullptr;
	static nccl_errstr_fn errstr = nullptr;
	if (!lib) {
		/* the process usually has NCCL loaded already (torch.distributed, or the caller's own link): look there first */
		allreduce = (nccl_allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
		if (allreduce) { lib = RTLD_DEFAULT; errstr = (nccl_errstr_fn)dlsym(RTLD_DEFAULT, "ncclGetErrorString"); }
		else {
			lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
			if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
			if (!lib) return bt_internal_fail(std::string("bt_counters_allreduce: NCCL is not available (") + dlerror() + ")");
			allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
			errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
			if (!allreduce) { lib = nullptr; return bt_internal_fail("bt_counters_allreduce: libnccl has no ncclAllReduce"); }
		}
	}
	unsigned long long *d = nullptr;
	cudaStream_t st = (cudaStream_t)stream;
	if (cudaMalloc((void **)&d, 5 * sizeof *d) != cudaSuccess) return bt_internal_fail("bt_counters_allreduce: cudaMalloc failed");
	int rc = 0;
	if (cudaMemcpyAsync(d, counters, 5 * sizeof *d, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = bt_internal_fail("bt_counters_allreduce: copy failed");
	if (!rc) {
		const int r = allreduce(d, d, 5, /* ncclUint64 */ 5, /* ncclSum */ 0, nccl_comm, st);
		if (r != 0) rc = bt_internal_fail(std::string("bt_counters_allreduce: ncclAllReduce: ") + (errstr ? errstr(r) : "error"));
	}
	if (!rc && (cudaMemcpyAsync(counters, d, 5 * sizeof *d, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) rc = bt_internal_fail("bt_counters_allreduce: copy back failed");
	cudaFree(d);
	return rc;
}
