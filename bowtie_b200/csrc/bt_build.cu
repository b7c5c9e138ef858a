/*
 * bt_build.cu — bt_index_build: index construction with the suffix sort on the device.
 *
 * The CUDA backend of bt_build_sa.cuh: radix sort / scan / compaction from CUB (library code, like cuBLAS for a GEMM — none of
 * this is on the search path) and one grid-stride kernel that applies a functor per element (bt_build_sa.cuh holds the functors:
 * key construction, group bookkeeping, BWT / side packing / occ / ftab histogram).  bt_build.h reads FASTA and writes the files.
 * Memory: the first sort holds 2 x (8 + 4) bytes per suffix (double-buffered), later rounds only touch the suffixes still in groups.
 */
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string>
#include <vector>

#include "bt_build_sa.cuh"
#include "../../include/bowtie_b200.h"

int bt_internal_fail(const std::string &m);                           /* bt_lib.cu: sets bt_last_error() */

template <class F>
__global__ void bsa_each_kernel(uint64_t n, F f) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

struct BsaMaxOp { __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

struct BsaCuda {
	cudaError_t st = cudaSuccess;
	void *tmp = nullptr; size_t tmp_bytes = 0;
	unsigned long long *d_count = nullptr;
	int sms = 148;
	bool verbose = getenv("BT_BUILD_VERBOSE") != nullptr; double t_last = 0;
	static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
	void mark(const char *what) {
		if (!verbose) return;
		cudaDeviceSynchronize();
		const double t = now();
		if (t_last > 0) fprintf(stderr, "[bt_index_build] %-40s %7.2f s\n", what, t - t_last);
		t_last = t;
	}
	void note(cudaError_t e) { if (e != cudaSuccess && st == cudaSuccess) st = e; }
	bool ok(std::string *err) {
		note(cudaDeviceSynchronize());
		if (st == cudaSuccess) return true;
		if (err) *err = std::string("CUDA error while building the index: ") + cudaGetErrorString(st);
		return false;
	}
	bool reserve(size_t bytes) {
		if (bytes <= tmp_bytes) return true;
		cudaFree(tmp); tmp = nullptr; tmp_bytes = 0;
		if (cudaMalloc(&tmp, bytes) != cudaSuccess) { note(cudaGetLastError()); return false; }
		tmp_bytes = bytes;
		return true;
	}
	template <class T> T *alloc(uint64_t n) {
		void *p = nullptr;
		if (cudaMalloc(&p, (size_t)(n ? n : 1) * sizeof(T)) != cudaSuccess) { note(cudaGetLastError()); return nullptr; }
		return (T *)p;
	}
	void release(void *p) { cudaFree(p); }
	void upload(void *d, const void *s, uint64_t bytes) { if (bytes) note(cudaMemcpy(d, s, (size_t)bytes, cudaMemcpyHostToDevice)); }
	void download(void *d, const void *s, uint64_t bytes) { if (bytes) note(cudaMemcpy(d, s, (size_t)bytes, cudaMemcpyDeviceToHost)); }
	void copy(void *d, const void *s, uint64_t bytes) { if (bytes) note(cudaMemcpy(d, s, (size_t)bytes, cudaMemcpyDeviceToDevice)); }
	void zero(void *d, uint64_t bytes) { if (bytes) note(cudaMemset(d, 0, (size_t)bytes)); }
	template <class F> void each(uint64_t n, F f) {
		if (!n) return;
		const uint64_t want = (n + 255) / 256, cap = (uint64_t)sms * 16;
		bsa_each_kernel<<<(unsigned)(want < cap ? want : cap), 256>>>(n, f);
		note(cudaGetLastError());
	}
	void sort_pairs(uint64_t *k0, uint64_t *k1, uint32_t *v0, uint32_t *v1, uint64_t n, int bits, uint64_t **kres, uint32_t **vres) {
		cub::DoubleBuffer<uint64_t> keys(k0, k1);
		cub::DoubleBuffer<uint32_t> vals(v0, v1);
		size_t bytes = 0;
		note(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys, vals, (long long)n, 0, bits));
		if (reserve(bytes)) note(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys, vals, (long long)n, 0, bits));
		*kres = keys.Current(); *vres = vals.Current();
	}
	void max_scan(uint32_t *a, uint64_t n) {
		size_t bytes = 0;
		note(cub::DeviceScan::InclusiveScan(nullptr, bytes, a, a, BsaMaxOp(), (long long)n));
		if (!reserve(bytes)) return;
		note(cub::DeviceScan::InclusiveScan(tmp, bytes, a, a, BsaMaxOp(), (long long)n));
	}
	void sum_scan(uint32_t *a, uint64_t n) {
		size_t bytes = 0;
		note(cub::DeviceScan::InclusiveSum(nullptr, bytes, a, a, (long long)n));
		if (!reserve(bytes)) return;
		note(cub::DeviceScan::InclusiveSum(tmp, bytes, a, a, (long long)n));
	}
	uint64_t select(const uint32_t *in, const uint8_t *flags, uint32_t *out, uint64_t n) {
		if (!d_count && cudaMalloc((void **)&d_count, sizeof *d_count) != cudaSuccess) { note(cudaGetLastError()); return 0; }
		size_t bytes = 0;
		thrust::counting_iterator<uint32_t> idx(0u);
		if (in) note(cub::DeviceSelect::Flagged(nullptr, bytes, in, flags, out, d_count, (long long)n));
		else note(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flags, out, d_count, (long long)n));
		if (!reserve(bytes)) return 0;
		if (in) note(cub::DeviceSelect::Flagged(tmp, bytes, in, flags, out, d_count, (long long)n));
		else note(cub::DeviceSelect::Flagged(tmp, bytes, idx, flags, out, d_count, (long long)n));
		unsigned long long m = 0;
		note(cudaMemcpy(&m, d_count, sizeof m, cudaMemcpyDeviceToHost));
		return st == cudaSuccess ? (uint64_t)m : 0;
	}
	~BsaCuda() { cudaFree(tmp); cudaFree(d_count); }
};

static int build_on_device(const BtRefInfo &R, const char *out_base, int off_rate, int ftab_chars, int device) {
	if (off_rate < 0 || off_rate > 31 || ftab_chars < 1 || ftab_chars > 15) return bt_internal_fail("bt_index_build: off_rate must be 0..31 and ftab_chars 1..15");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return bt_internal_fail("bt_index_build: no CUDA device (this library has no CPU path)"); }
	if (device < 0 || device >= ndev) return bt_internal_fail("bt_index_build: bad device ordinal");
	if (cudaSetDevice(device) != cudaSuccess) return bt_internal_fail(std::string("bt_index_build: cudaSetDevice: ") + cudaGetErrorString(cudaGetLastError()));
	BsaCuda be;
	cudaDeviceGetAttribute(&be.sms, cudaDevAttrMultiProcessorCount, device);
	BtBuildParams P; P.offRate = off_rate; P.ftabChars = ftab_chars;
	std::string err;
	be.t_last = BsaCuda::now();
	if (!bt_build_all_on(be, R, out_base, P, err)) { if (be.st != cudaSuccess) err = std::string("CUDA error while building the index: ") + cudaGetErrorString(be.st) + (err.empty() ? "" : " (" + err + ")"); return bt_internal_fail(err); }
	if (!be.ok(&err)) return bt_internal_fail(err);
	return 0;
}

extern "C" int bt_index_build(const char *const *fasta_paths, uint32_t n_paths, const char *out_base, int off_rate, int ftab_chars, int device) {
	if (!fasta_paths || !n_paths || !out_base) return bt_internal_fail("bt_index_build: null argument");
	std::vector<std::string> files;
	for (uint32_t i = 0; i < n_paths; i++) files.push_back(fasta_paths[i] ? fasta_paths[i] : "");
	BtRefInfo R; std::string err;
	if (!bt_build_read_fasta(files, false, R, err)) return bt_internal_fail(err);
	return build_on_device(R, out_base, off_rate, ftab_chars, device);
}

extern "C" int bt_index_build_text(const uint8_t *text, uint64_t text_len, const bt_ref_record_t *recs, uint32_t n_recs, const char *const *names, uint32_t n_names,
                                   const char *out_base, int off_rate, int ftab_chars, int device) {
	if (!text || !recs || !n_recs || !names || !out_base) return bt_internal_fail("bt_index_build_text: null argument");
	BtRefInfo R; std::string err;
	R.text = text; R.textLen = text_len;
	for (uint32_t i = 0; i < n_recs; i++) {
		R.recs.push_back({ recs[i].off, recs[i].len, (uint8_t)(recs[i].first ? 1 : 0) });
		if (recs[i].first) R.plens.push_back(0);
		if (R.plens.empty()) return bt_internal_fail("bt_index_build_text: the first record must start a sequence");
		R.plens.back() += recs[i].off + recs[i].len;
	}
	for (uint32_t i = 0; i < n_names; i++) R.names.push_back(names[i] ? names[i] : "");
	if (!bt_build_check_ref(R, err)) return bt_internal_fail(err);
	return build_on_device(R, out_base, off_rate, ftab_chars, device);
}
