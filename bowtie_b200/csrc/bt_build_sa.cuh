/*
 * bt_build_sa.cuh — the suffix sort behind bt_build.h, written against a small backend interface.
 *
 * The reference sorts suffixes block by block on the CPU with a difference-cover sample (blockwise_sa.h, diff_sample.h,
 * multikey_qsort.h) because it was designed for machines that cannot hold the suffix array; a B200 holds text, suffix array
 * and inverse for a 4-Gbp text many times over, so this is plain prefix doubling (Manber-Myers with Larsson-Sadakane's
 * "leave sorted groups alone") made of radix sorts, scans and compactions over flat arrays:
 *
 *   round 0   key(i) = the first 21 characters of suffix i, 3 bits each (A C G T = 0..3, end of text = 4: the end compares
 *             greater than every character, see bt_build.h); sort (key, i); a group = a run of equal keys; its rank = the
 *             row of its first member; ISA[i] = rank of i's group.  Groups of one are final.
 *   round h   for the rows still in groups of several: key = (rank of the group) << 32 | ISA[i + h]; sort; write back into
 *             the same rows; split groups where keys differ; drop the rows that ended up alone.  h doubles.
 *
 * The end-of-text character is unique, so two different suffixes never tie through it and i + h never runs past the text
 * for a suffix that still shares a group.  Everything after the sort that needs the suffix array — BWT characters, zOff,
 * the SA sample, the rows of the suffixes shorter than ftabChars — is read off on the device too (BsaExtract), so only
 * len + 1 bytes and the sample travel back.
 *
 * Backends: BsaCuda (bt_lib.cu; cub::DeviceRadixSort / DeviceScan / DeviceSelect and one-line kernels) is the product;
 * tests/host_emu/build_emu.cpp instantiates the same template with std:: algorithms to check the algorithm without a GPU.
 * A backend provides:
 *   T *alloc<T>(n), release(p), upload(dst, src, bytes), download(dst, src, bytes)
 *   each(n, functor)                                  functor(i) for i in [0, n)
 *   sort_pairs(kin, kout, vin, vout, n)               by the 64-bit key, ascending
 *   max_scan(a, n)                                    inclusive, in place
 *   select(in, flags, out, n) -> count                stable compaction of in[i] where flags[i]; in may be NULL = the index i
 */
#pragma once
#include "bt_build.h"

#ifdef __CUDACC__
#define BSA_FN __host__ __device__ __forceinline__
#else
#define BSA_FN inline
#endif

#define BSA_H0 21                                                     /* characters in the first key */

struct BsaInitKey {
	const uint8_t *s; uint64_t len; uint64_t *key; uint32_t *val;
	BSA_FN void operator()(uint64_t i) const {
		uint64_t k = 0;
		for (int j = 0; j < BSA_H0; j++) { const uint64_t p = i + (uint64_t)j; const uint64_t c = p < len ? s[p] : (p == len ? 4u : 0u); k = (k << 3) | c; }
		key[i] = k; val[i] = (uint32_t)i;
	}
};
/* heads of the groups of a sorted key array: out[p] = p where a group starts, else 0 (a max-scan turns it into the rank) */
struct BsaHeads {
	const uint64_t *key; uint32_t *out;
	BSA_FN void operator()(uint64_t p) const { out[p] = (p == 0 || key[p] != key[p - 1]) ? (uint32_t)p : 0u; }
};
struct BsaSetIsa {                                                    /* ISA[SA[p]] = rank[p] */
	const uint32_t *sa, *rank; uint32_t *isa;
	BSA_FN void operator()(uint64_t p) const { isa[sa[p]] = rank[p]; }
};
struct BsaActive0 {                                                   /* rows whose group has more than one member */
	const uint32_t *rank; uint64_t n; uint8_t *act;
	BSA_FN void operator()(uint64_t p) const { const bool head = rank[p] == (uint32_t)p, nextHead = (p + 1 == n) || rank[p + 1] == (uint32_t)(p + 1); act[p] = !(head && nextHead); }
};
struct BsaRoundKey {
	const uint32_t *slots, *sa, *isa; uint64_t h, len; uint64_t *key; uint32_t *sfx;
	BSA_FN void operator()(uint64_t k) const {
		const uint32_t x = sa[slots[k]];
		const uint64_t y = (uint64_t)x + h;
		key[k] = ((uint64_t)isa[x] << 32) | (y <= len ? isa[y] : 0u);
		sfx[k] = x;
	}
};
struct BsaRoundHeads {                                                /* write the sorted suffixes back and mark the new group heads by their row */
	const uint64_t *key; const uint32_t *sfx, *slots; uint32_t *sa, *out;
	BSA_FN void operator()(uint64_t k) const { sa[slots[k]] = sfx[k]; out[k] = (k == 0 || key[k] != key[k - 1]) ? slots[k] : 0u; }
};
struct BsaRoundIsa {
	const uint32_t *sfx, *rank; uint32_t *isa;
	BSA_FN void operator()(uint64_t k) const { isa[sfx[k]] = rank[k]; }
};
struct BsaRoundActive {
	const uint32_t *rank, *slots; uint64_t m; uint8_t *act;
	BSA_FN void operator()(uint64_t k) const { const bool head = rank[k] == slots[k], nextHead = (k + 1 == m) || rank[k + 1] == slots[k + 1]; act[k] = !(head && nextHead); }
};
/* what buildToDisk reads off the suffix array row by row (ebwt.h:4119-4185) */
struct BsaExtract {
	const uint8_t *s; const uint32_t *sa; uint64_t len; uint32_t offMask; int offRate; uint32_t K;
	uint8_t *bwt; uint32_t *offs, *zoff, *shortRow;                   /* shortRow[j] = row of the suffix of length j < K */
	BSA_FN void operator()(uint64_t row) const {
		const uint32_t p = sa[row];
		bwt[row] = p ? s[p - 1] : 0;
		if (p == 0) *zoff = (uint32_t)row;
		if ((row & offMask) == 0) offs[row >> offRate] = p;
		if (len - p < K) shortRow[len - p] = (uint32_t)row;
	}
};

template <class B>
static bool bt_suffix_sort(B &be, const uint8_t *text_host, uint32_t len, int offRate, int ftabChars, BtSuffixResult *out, std::string *err) {
	const uint64_t n = (uint64_t)len + 1;
	uint8_t *s = be.template alloc<uint8_t>(len ? len : 1);
	uint64_t *kin = be.template alloc<uint64_t>(n), *kout = be.template alloc<uint64_t>(n);
	uint32_t *vin = be.template alloc<uint32_t>(n), *sa = be.template alloc<uint32_t>(n);
	if (!s || !kin || !kout || !vin || !sa) { if (err) *err = "Error: out of device memory in the suffix sort"; return false; }
	be.upload(s, text_host, len);
	be.each(n, BsaInitKey{ s, len, kin, vin });
	be.sort_pairs(kin, kout, vin, sa, n);
	be.release(kin); be.release(vin);
	uint32_t *rank = be.template alloc<uint32_t>(n), *isa = be.template alloc<uint32_t>(n);
	uint8_t *act = be.template alloc<uint8_t>(n);
	if (!rank || !isa || !act) { if (err) *err = "Error: out of device memory in the suffix sort"; return false; }
	be.each(n, BsaHeads{ kout, rank });
	be.release(kout);
	be.max_scan(rank, n);
	be.each(n, BsaSetIsa{ sa, rank, isa });
	be.each(n, BsaActive0{ rank, n, act });
	uint32_t *slots = be.template alloc<uint32_t>(n);
	if (!slots) { if (err) *err = "Error: out of device memory in the suffix sort"; return false; }
	uint64_t m = be.select((const uint32_t *)NULL, act, slots, n);
	be.release(rank); be.release(act);
	uint64_t h = BSA_H0;
	while (m > 0) {
		uint64_t *k2 = be.template alloc<uint64_t>(m), *k2s = be.template alloc<uint64_t>(m);
		uint32_t *sfx = be.template alloc<uint32_t>(m), *sfxs = be.template alloc<uint32_t>(m), *nr = be.template alloc<uint32_t>(m), *slots2 = be.template alloc<uint32_t>(m);
		uint8_t *act2 = be.template alloc<uint8_t>(m);
		if (!k2 || !k2s || !sfx || !sfxs || !nr || !slots2 || !act2) { if (err) *err = "Error: out of device memory in the suffix sort"; return false; }
		be.each(m, BsaRoundKey{ slots, sa, isa, h, len, k2, sfx });
		be.sort_pairs(k2, k2s, sfx, sfxs, m);
		be.each(m, BsaRoundHeads{ k2s, sfxs, slots, sa, nr });
		be.max_scan(nr, m);
		be.each(m, BsaRoundIsa{ sfxs, nr, isa });
		be.each(m, BsaRoundActive{ nr, slots, m, act2 });
		const uint64_t m2 = be.select(slots, act2, slots2, m);
		be.release(k2); be.release(k2s); be.release(sfx); be.release(sfxs); be.release(nr); be.release(act2); be.release(slots);
		slots = slots2; m = m2;
		h *= 2;
		if (h > 2 * n + BSA_H0) { if (err) *err = "internal error: suffix sort did not converge"; return false; }
	}
	be.release(slots); be.release(isa);

	/* read the results off the suffix array */
	const uint32_t K = (uint32_t)ftabChars;
	const uint64_t offsLen = (n + (1ull << offRate) - 1) >> offRate;
	uint8_t *bwt = be.template alloc<uint8_t>(n);
	uint32_t *offs = be.template alloc<uint32_t>(offsLen), *small = be.template alloc<uint32_t>(K + 1);
	if (!bwt || !offs || !small) { if (err) *err = "Error: out of device memory in the suffix sort"; return false; }
	be.each(n, BsaExtract{ s, sa, len, (uint32_t)((1u << offRate) - 1), offRate, K, bwt, offs, small + K, small });
	out->bwt.resize(n); out->offs.resize(offsLen);
	be.download(out->bwt.data(), bwt, n);
	be.download(out->offs.data(), offs, offsLen * 4);
	std::vector<uint32_t> sm(K + 1);
	be.download(sm.data(), small, (K + 1) * 4);
	out->zOff = sm[K];
	/* runs of rows that hold suffixes shorter than ftabChars, and the k-mer of the suffix in the row after each run */
	std::vector<uint32_t> rows;
	for (uint32_t j = 0; j < K && j <= len; j++) rows.push_back(sm[j]);
	for (size_t a = 0; a < rows.size(); a++) for (size_t b = a + 1; b < rows.size(); b++) if (rows[b] < rows[a]) { const uint32_t t = rows[a]; rows[a] = rows[b]; rows[b] = t; }
	out->absorb.clear();
	for (size_t a = 0; a < rows.size();) {
		size_t b = a;
		while (b + 1 < rows.size() && rows[b + 1] == rows[b] + 1) b++;
		const uint64_t next = (uint64_t)rows[b] + 1;
		uint32_t kmer = (uint32_t)(1ull << (2 * K));
		if (next <= len) {
			uint32_t p = 0;
			be.download(&p, sa + next, 4);
			kmer = 0;
			for (uint32_t i = 0; i < K; i++) kmer = (kmer << 2) | text_host[p + i];
		}
		out->absorb.push_back({ kmer, (uint32_t)(b - a + 1) });
		a = b + 1;
	}
	be.release(bwt); be.release(offs); be.release(small); be.release(sa); be.release(s);
	return true;
}
