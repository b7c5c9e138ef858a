/* Test-only: a host backend for bt_build_sa.cuh (std:: algorithms instead of CUB) and a comparison-sort cross-check.
 * Used by build_emu.cpp and by the emulation shim's bt_index_build; the product sorts on the GPU (bt_build.cu). */
#pragma once
#include <algorithm>
#include <numeric>
#include <stdlib.h>
#include "../../bowtie_b200/csrc/bt_build_sa.cuh"

/* shared with the device path: what to read off a finished suffix array */
static bool sa_to_result(const uint8_t *s, uint32_t len, const std::vector<uint32_t> &sa, int offRate, int ftabChars, BtSuffixResult *out) {
	out->bwt.assign((size_t)len + 1, 0);
	out->offs.clear(); out->absorb.clear();
	const uint32_t K = (uint32_t)ftabChars;
	uint32_t run = 0;
	for (uint64_t row = 0; row <= len; row++) {
		const uint32_t p = sa[row];
		if (p == 0) out->zOff = (uint32_t)row; else out->bwt[row] = s[p - 1];
		if ((row & ((1ull << offRate) - 1)) == 0) out->offs.push_back(p);
		if (len - p < K) run++;
		else if (run) { uint32_t v = 0; for (uint32_t i = 0; i < K; i++) v = (v << 2) | s[p + i]; out->absorb.push_back({ v, run }); run = 0; }
	}
	if (run) out->absorb.push_back({ (uint32_t)(1ull << (2 * K)), run });
	return true;
}

static bool host_sort(const uint8_t *s, uint32_t len, int offRate, int ftabChars, BtSuffixResult *out, void *, std::string *) {
	std::vector<uint32_t> sa((size_t)len + 1);
	std::iota(sa.begin(), sa.end(), 0u);
	std::sort(sa.begin(), sa.end(), [&](uint32_t a, uint32_t b) {
		if (a == b) return false;
		const uint32_t la = len - a, lb = len - b, m = la < lb ? la : lb;
		const int c = memcmp(s + a, s + b, m);
		if (c) return c < 0;
		return la > lb;                                                /* the end of the text is greater than any character */
	});
	return sa_to_result(s, len, sa, offRate, ftabChars, out);
}

/* the backend interface of bt_build_sa.cuh over std:: algorithms */
struct BsaHost {
	template <class T> T *alloc(uint64_t n) { return (T *)malloc((size_t)(n ? n : 1) * sizeof(T)); }
	void release(void *p) { free(p); }
	void upload(void *d, const void *s, uint64_t bytes) { memcpy(d, s, (size_t)bytes); }
	void download(void *d, const void *s, uint64_t bytes) { memcpy(d, s, (size_t)bytes); }
	template <class F> void each(uint64_t n, F f) { for (uint64_t i = 0; i < n; i++) f(i); }
	void sort_pairs(const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, uint64_t n) {
		std::vector<uint64_t> idx((size_t)n);
		std::iota(idx.begin(), idx.end(), 0ull);
		std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return kin[a] < kin[b]; });
		for (uint64_t i = 0; i < n; i++) { kout[i] = kin[idx[(size_t)i]]; vout[i] = vin[idx[(size_t)i]]; }
	}
	void max_scan(uint32_t *a, uint64_t n) { for (uint64_t i = 1; i < n; i++) if (a[i] < a[i - 1]) a[i] = a[i - 1]; }
	uint64_t select(const uint32_t *in, const uint8_t *flags, uint32_t *out, uint64_t n) {
		uint64_t m = 0;
		for (uint64_t i = 0; i < n; i++) if (flags[i]) out[m++] = in ? in[i] : (uint32_t)i;
		return m;
	}
};

static bool doubling_sort(const uint8_t *s, uint32_t len, int offRate, int ftabChars, BtSuffixResult *out, void *, std::string *err) {
	BsaHost be;
	if (!bt_suffix_sort(be, s, len, offRate, ftabChars, out, err)) return false;
	if (getenv("BT_BUILD_CROSSCHECK")) {                              /* against the comparison sort above */
		BtSuffixResult ref;
		host_sort(s, len, offRate, ftabChars, &ref, NULL, NULL);
		if (ref.bwt != out->bwt || ref.zOff != out->zOff || ref.offs != out->offs || ref.absorb != out->absorb) { *err = "prefix doubling and the comparison sort disagree"; return false; }
	}
	return true;
}

