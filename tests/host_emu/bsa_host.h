/* Test-only: a host backend for bt_build_sa.cuh (std:: algorithms instead of CUB) and a comparison-sort cross-check.
 * Used by build_emu.cpp and by the emulation shim's bt_index_build; the product runs the same templates on the GPU (bt_build.cu). */
#pragma once
#include <algorithm>
#include <numeric>
#include <stdlib.h>
#include "../../bowtie_b200/csrc/bt_build_sa.cuh"

/* the backend interface of bt_build_sa.cuh over std:: algorithms */
struct BsaHost {
	template <class T> T *alloc(uint64_t n) { return (T *)malloc((size_t)(n ? n : 1) * sizeof(T)); }
	void release(void *p) { free(p); }
	bool ok(std::string *) { return true; }
	void mark(const char *) { }
	void upload(void *d, const void *s, uint64_t bytes) { memcpy(d, s, (size_t)bytes); }
	void download(void *d, const void *s, uint64_t bytes) { memcpy(d, s, (size_t)bytes); }
	void copy(void *d, const void *s, uint64_t bytes) { memcpy(d, s, (size_t)bytes); }
	void zero(void *d, uint64_t bytes) { memset(d, 0, (size_t)bytes); }
	template <class F> void each(uint64_t n, F f) { for (uint64_t i = 0; i < n; i++) f(i); }
	void sort_pairs(uint64_t *k0, uint64_t *k1, uint32_t *v0, uint32_t *v1, uint64_t n, int bits, uint64_t **kres, uint32_t **vres) {
		const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
		std::vector<uint64_t> idx((size_t)n);
		std::iota(idx.begin(), idx.end(), 0ull);
		std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return (k0[a] & mask) < (k0[b] & mask); });
		for (uint64_t i = 0; i < n; i++) { k1[i] = k0[idx[(size_t)i]]; v1[i] = v0[idx[(size_t)i]]; }
		if (rand() & 1) { memcpy(k0, k1, (size_t)n * 8); memcpy(v0, v1, (size_t)n * 4); memset(k1, 0xee, (size_t)n * 8); memset(v1, 0xee, (size_t)n * 4); *kres = k0; *vres = v0; }   /* either buffer may hold the result */
		else { memset(k0, 0xee, (size_t)n * 8); memset(v0, 0xee, (size_t)n * 4); *kres = k1; *vres = v1; }
	}
	void max_scan(uint32_t *a, uint64_t n) { for (uint64_t i = 1; i < n; i++) if (a[i] < a[i - 1]) a[i] = a[i - 1]; }
	void sum_scan(uint32_t *a, uint64_t n) { for (uint64_t i = 1; i < n; i++) a[i] += a[i - 1]; }
	uint64_t select(const uint32_t *in, const uint8_t *flags, uint32_t *out, uint64_t n) {
		uint64_t m = 0;
		for (uint64_t i = 0; i < n; i++) if (flags[i]) out[m++] = in ? in[i] : (uint32_t)i;
		return m;
	}
};

/* BT_BUILD_CROSSCHECK: the prefix-doubling suffix array against a comparison sort */
struct BsaHostChecked : BsaHost { };
static bool host_sa_crosscheck(const uint8_t *s, uint32_t len, std::string *err) {
	BsaHost be;
	uint32_t *sa = nullptr;
	if (!bt_suffix_sort(be, s, len, &sa, err)) return false;
	std::vector<uint32_t> ref((size_t)len + 1);
	std::iota(ref.begin(), ref.end(), 0u);
	std::sort(ref.begin(), ref.end(), [&](uint32_t a, uint32_t b) {
		if (a == b) return false;
		const uint32_t la = len - a, lb = len - b, m = la < lb ? la : lb;
		const int c = memcmp(s + a, s + b, m);
		if (c) return c < 0;
		return la > lb;                                                /* the end of the text is greater than any character */
	});
	const bool same = memcmp(ref.data(), sa, ((size_t)len + 1) * 4) == 0;
	free(sa);
	if (!same) { *err = "prefix doubling and the comparison sort disagree"; return false; }
	return true;
}

static bool host_build_all(const std::vector<std::string> &fasta, const std::string &base, const BtBuildParams &P, std::string &err) {
	BtRefInfo R;
	if (!bt_build_read_fasta(fasta, P.nsToAs, R, err)) return false;
	if (getenv("BT_BUILD_CROSSCHECK") && !host_sa_crosscheck(R.text, (uint32_t)R.textLen, &err)) return false;
	BsaHost be;
	return bt_build_all_on(be, R, base, P, err);
}
