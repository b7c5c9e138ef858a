/*
 * bt_io.cu — the C ABI of the device I/O path: FASTQ text in, formatted hits out, with the search in between
 * on the same device buffers.  The per-element code is bt_io.cuh, the driver bt_io_run.h; this file is the CUDA backend
 * (CUB reduce / select / scan: library code off the search path) and the entry points.
 */
#include <cuda_runtime.h>
#include <utility>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <string>
#include <vector>

#include "bt_io_run.h"

int bt_internal_fail(const std::string &m);                           /* bt_lib.cu: sets bt_last_error() */
struct bt_index;
int bt_internal_context_device(bt_context_t *cx);                     /* bt_lib.cu */
bt_index_t *bt_internal_context_index(bt_context_t *cx);

template <class F>
__global__ void bio_each_kernel(uint64_t n, F f) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
struct BioIsNlVal { const char *t; __host__ __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return t[i] == '\n'; } };
struct BioIsNlPred { const char *t; __host__ __device__ __forceinline__ bool operator()(uint32_t i) const { return t[i] == '\n'; } };

struct BioCuda {
	cudaStream_t st = nullptr; bt_context_t *cx = nullptr;
	cudaStream_t st_hi = nullptr;        /* highest priority: record cutting (small kernels) must not queue behind other chunks' search kernels, which fill the machine */
	void *tmp = nullptr; size_t tmp_bytes = 0;
	unsigned long long *d_num = nullptr;
	char *pinned = nullptr; size_t pinned_cap = 0;
	cudaError_t e = cudaSuccess;
	int sms = 148;
	void note(cudaError_t x) { if (x != cudaSuccess && e == cudaSuccess) e = x; }
	void *alloc(size_t bytes) { void *p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { note(cudaGetLastError()); return nullptr; } return p; }
	void release(void *p) { cudaFree(p); }
	bool reserve(size_t bytes) {
		if (bytes <= tmp_bytes) return true;
		cudaFree(tmp); tmp = nullptr; tmp_bytes = 0;
		if (cudaMalloc(&tmp, bytes) != cudaSuccess) { note(cudaGetLastError()); return false; }
		tmp_bytes = bytes; return true;
	}
	void h2d(void *d, const void *s, size_t bytes) { if (bytes) note(cudaMemcpyAsync(d, s, bytes, cudaMemcpyHostToDevice, st)); }
	void d2h(void *d, const void *s, size_t bytes) { if (bytes) { note(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToHost, st)); note(cudaStreamSynchronize(st)); } }
	void zero(void *d, size_t bytes) { if (bytes) note(cudaMemsetAsync(d, 0, bytes, st)); }
	void sync() { note(cudaStreamSynchronize(st)); }
	template <class F> void each(uint64_t n, F f) {
		if (!n) return;
		const uint64_t want = (n + 255) / 256, cap = (uint64_t)sms * 16;
		bio_each_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(n, f);
		note(cudaGetLastError());
	}
	uint64_t count_nl(const char *text, uint64_t n) {
		if (!d_num && cudaMalloc((void **)&d_num, 16) != cudaSuccess) { note(cudaGetLastError()); return 0; }
		thrust::counting_iterator<uint64_t> idx(0);
		thrust::transform_iterator<BioIsNlVal, thrust::counting_iterator<uint64_t>> it(idx, BioIsNlVal{ text });
		size_t bytes = 0;
		uint32_t *d_out = (uint32_t *)d_num;
		note(cub::DeviceReduce::Sum(nullptr, bytes, it, d_out, (long long)n, st));
		if (!reserve(bytes)) return 0;
		note(cub::DeviceReduce::Sum(tmp, bytes, it, d_out, (long long)n, st));
		uint32_t m = 0;
		d2h(&m, d_out, 4);
		return m;
	}
	void positions_nl(const char *text, uint64_t n, uint32_t *out) {
		thrust::counting_iterator<uint32_t> idx(0u);
		size_t bytes = 0;
		note(cub::DeviceSelect::If(nullptr, bytes, idx, out, d_num + 1, (long long)n, BioIsNlPred{ text }, st));
		if (!reserve(bytes)) return;
		note(cub::DeviceSelect::If(tmp, bytes, idx, out, d_num + 1, (long long)n, BioIsNlPred{ text }, st));
	}
	void excl_scan(uint32_t *a, uint64_t n) {
		size_t bytes = 0;
		note(cub::DeviceScan::ExclusiveSum(nullptr, bytes, a, a, (long long)n, st));
		if (!reserve(bytes)) return;
		note(cub::DeviceScan::ExclusiveSum(tmp, bytes, a, a, (long long)n, st));
	}
	int align(const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out) {
		if (bt_context_align_device(cx, pol, in, out, st)) return 1;
		return bt_context_join(cx, st);
	}
	char *out_host(size_t bytes, std::vector<char> &) {
		if (bytes > pinned_cap) {
			if (pinned) cudaFreeHost(pinned);
			pinned = nullptr; pinned_cap = 0;
			const size_t want = bytes + bytes / 4 + (1 << 20);
			if (cudaMallocHost((void **)&pinned, want) != cudaSuccess) { note(cudaGetLastError()); return nullptr; }
			pinned_cap = want;
		}
		return pinned;
	}
	void destroy() { cudaFree(tmp); cudaFree(d_num); if (pinned) cudaFreeHost(pinned); if (st) cudaStreamDestroy(st); if (st_hi) cudaStreamDestroy(st_hi); }
};

struct bt_io { BioPipe<BioCuda> pipe; int device = 0; };

static int io_fail(bt_io *io, const char *what) {
	std::string m = std::string(what) + ": " + (io->pipe.err.empty() ? "failed" : io->pipe.err);
	if (io->pipe.be.e != cudaSuccess) m += std::string(" (CUDA: ") + cudaGetErrorString(io->pipe.be.e) + ")";
	io->pipe.err.clear(); io->pipe.be.e = cudaSuccess;
	return bt_internal_fail(m);
}

extern "C" int bt_io_create(bt_context_t *cx, bt_io_t **out) {
	if (!cx || !out) return bt_internal_fail("bt_io_create: null argument");
	*out = nullptr;
	const int dev = bt_internal_context_device(cx);
	if (cudaSetDevice(dev) != cudaSuccess) return bt_internal_fail("bt_io_create: cudaSetDevice failed");
	bt_io *io = new bt_io();
	io->device = dev; io->pipe.be.cx = cx;
	cudaDeviceGetAttribute(&io->pipe.be.sms, cudaDevAttrMultiProcessorCount, dev);
	int prLo = 0, prHi = 0;
	cudaDeviceGetStreamPriorityRange(&prLo, &prHi);                          /* numerically lower = higher priority */
	if (cudaStreamCreateWithFlags(&io->pipe.be.st, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithPriority(&io->pipe.be.st_hi, cudaStreamNonBlocking, prHi) != cudaSuccess || !io->pipe.init()) { io->pipe.destroy(); io->pipe.be.destroy(); delete io; return bt_internal_fail("bt_io_create: CUDA resource allocation failed"); }
	bt_index_t *ix = bt_internal_context_index(cx);
	bt_index_info_t info;
	bt_index_info(ix, &info);
	(void)info;
	*out = io;
	return 0;
}
extern "C" void bt_io_free(bt_io_t *io) {
	if (!io) return;
	cudaSetDevice(io->device);
	io->pipe.be.sync();
	io->pipe.destroy(); io->pipe.be.destroy();
	delete io;
}
extern "C" int bt_io_parse_fastq(bt_io_t *io, const char *text, uint64_t nbytes, uint32_t global_seed, uint32_t max_reads, uint32_t *nreads, uint64_t *consumed, int *irregular) {
	if (!io || (!text && nbytes) || !nreads || !consumed || !irregular) return bt_internal_fail("bt_io_parse_fastq: null argument");
	if (cudaSetDevice(io->device) != cudaSuccess) return bt_internal_fail("bt_io_parse_fastq: cudaSetDevice failed");
	/* the whole parse runs on the high-priority stream and ends with a synchronising read-back, so the search (on the io's ordinary
	 * stream) always starts after it */
	std::swap(io->pipe.be.st, io->pipe.be.st_hi);
	const bool ok = io->pipe.parse(text, nbytes, global_seed, max_reads, nreads, consumed, irregular);
	io->pipe.be.sync();
	std::swap(io->pipe.be.st, io->pipe.be.st_hi);
	if (!ok || io->pipe.be.e != cudaSuccess) return io_fail(io, "bt_io_parse_fastq");
	return 0;
}
extern "C" int bt_io_align_format(bt_io_t *io, const bt_policy_t *pol, const bt_io_format_t *fmt, const char **out_text, uint64_t *out_bytes, uint64_t counters[4]) {
	if (!io || !pol || !fmt || !out_text || !out_bytes || !counters) return bt_internal_fail("bt_io_align_format: null argument");
	if (cudaSetDevice(io->device) != cudaSuccess) return bt_internal_fail("bt_io_align_format: cudaSetDevice failed");
	if (pol->paired || pol->all_hits || pol->sample_max) return bt_internal_fail("bt_io_align_format: paired-end, -a and -M output are formatted by the caller");
	if (pol->khits == 0 || pol->khits > 16) return bt_internal_fail("bt_io_align_format: -k must be 1..16 on the device output path");
	BioPipe<BioCuda> &p = io->pipe;
	if (!p.have_names || p.names_full != (fmt->full_ref != 0)) {
		bt_index_t *ix = bt_internal_context_index(p.be.cx);
		bt_index_info_t info;
		if (bt_index_info(ix, &info)) return 1;
		std::vector<std::string> names;
		for (uint32_t i = 0; i < info.n_refs; i++) { const char *nm = bt_index_refname(ix, i); names.push_back(nm ? nm : std::to_string(i)); }
		if (!p.set_names(names, fmt->full_ref != 0)) return io_fail(io, "bt_io_align_format");
	}
	BioFmt f;
	f.sam = fmt->sam ? 1u : 0u; f.khits = pol->khits; f.mhits = pol->mhits; f.strata = pol->strata ? 1u : 0u; f.noUnal = fmt->no_unal ? 1u : 0u;
	f.noQnameTrunc = fmt->no_qname_trunc ? 1u : 0u; f.offBase = (uint32_t)fmt->off_base; f.mapq = fmt->mapq; f.slots = pol->khits; f.recWords = 0;
	if (!p.align_format(pol, f, out_text, out_bytes, counters) || p.be.e != cudaSuccess) {
		const bool nc = p.not_covered && p.be.e == cudaSuccess;
		io_fail(io, "bt_io_align_format");
		return nc ? 2 : 1;                                            /* 2: nothing produced, the caller formats this batch itself */
	}
	return 0;
}
