/*
 * bt_build_sa.cuh — the device side of index construction, written against a small backend interface.
 *
 * Part 1, the suffix sort.  The reference sorts suffixes block by block on the CPU with a difference-cover sample
 * (blockwise_sa.h, diff_sample.h, multikey_qsort.h) because it was designed for machines that cannot hold the suffix array;
 * a B200 holds text, suffix array and inverse for a 4-Gbp text many times over, so this is plain prefix doubling
 * (Manber-Myers with Larsson-Sadakane's "leave sorted groups alone") made of radix sorts, scans and compactions over flat arrays:
 *
 *   round 0   key(i) = the first 21 characters of suffix i, 3 bits each (A C G T = 0..3, end of text = 4: the end compares
 *             greater than every character, see bt_build.h); sort (key, i); a group = a run of equal keys; its rank = the
 *             row of its first member; ISA[i] = rank of i's group.  Groups of one are final.
 *   round h   for the rows still in groups of several: key = (rank of the group) << 32 | ISA[i + h]; sort; write back into
 *             the same rows; split groups where keys differ; drop the rows that ended up alone.  h doubles.
 *
 * The end-of-text character is unique, so two different suffixes never tie through it and i + h never runs past the text
 * for a suffix that still shares a group.
 *
 * Part 2, everything buildToDisk (ebwt.h:3985-4388) derives from the suffix array, also per element on the device: the BWT
 * character of every row, zOff, the SA sample, the rows of the suffixes shorter than ftabChars, the per-side character counts
 * (-> occ words by a prefix sum), the packed side pairs, the ftab k-mer histogram.  Only the finished file images travel back.
 *
 * Backends: BsaCuda (bt_build.cu; cub::DeviceRadixSort / DeviceScan / DeviceSelect and one grid-stride kernel per functor) is
 * the product; tests/host_emu/bsa_host.h instantiates the same templates with std:: algorithms to check them without a GPU.
 * A backend provides:
 *   T *alloc<T>(n), release(p), upload(dst, src, bytes), download(dst, src, bytes), copy(dst, src, bytes), zero(p, bytes)
 *   ok(&err) -> no error so far (synchronises); mark(label) progress / timing hook
 *   each(n, functor)                                  functor(i) for i in [0, n)
 *   sort_pairs(k0, k1, v0, v1, n, bits, &kres, &vres) by the low `bits` bits of the 64-bit key, ascending; either buffer pair may
 *                                                     be clobbered, *kres / *vres say which one holds the result
 *   max_scan(a, n), sum_scan(a, n)                    inclusive, in place
 *   select(in, flags, out, n) -> count                stable compaction of in[i] where flags[i]; in may be NULL = the index i
 */
#pragma once
#include "bt_build.h"

#ifdef __CUDACC__
#define BSA_FN __host__ __device__ __forceinline__
#else
#define BSA_FN inline
#endif
#if defined(__CUDA_ARCH__)
#define BSA_ATOMIC_INC(p) atomicAdd((p), 1u)
#else
#define BSA_ATOMIC_INC(p) (++*(p))
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

/* Suffix array of the device-resident text s[0, len) (codes 0..3); rows 0..len, the empty suffix last.  The caller releases *sa_out. */
template <class B>
static bool bt_suffix_sort(B &be, const uint8_t *s, uint32_t len, uint32_t **sa_out, std::string *err) {
	const uint64_t n = (uint64_t)len + 1;
	const char *oom = "Error: out of device memory in the suffix sort";
	uint64_t *k0 = be.template alloc<uint64_t>(n), *k1 = be.template alloc<uint64_t>(n);
	uint32_t *v0 = be.template alloc<uint32_t>(n), *v1 = be.template alloc<uint32_t>(n);
	if (!k0 || !k1 || !v0 || !v1) { if (err) *err = oom; return false; }
	be.each(n, BsaInitKey{ s, len, k0, v0 });
	uint64_t *ks = nullptr; uint32_t *sa = nullptr;
	be.sort_pairs(k0, k1, v0, v1, n, 3 * BSA_H0, &ks, &sa);
	be.mark("sort: first keys (21 characters)");
	be.release(sa == v0 ? v1 : v0);
	uint32_t *rank = be.template alloc<uint32_t>(n);
	if (!rank) { if (err) *err = oom; return false; }
	be.each(n, BsaHeads{ ks, rank });
	be.release(k0); be.release(k1);
	be.max_scan(rank, n);
	uint32_t *isa = be.template alloc<uint32_t>(n);
	uint8_t *act = be.template alloc<uint8_t>(n);
	if (!isa || !act) { if (err) *err = oom; return false; }
	be.each(n, BsaSetIsa{ sa, rank, isa });
	be.each(n, BsaActive0{ rank, n, act });
	uint32_t *slots = be.template alloc<uint32_t>(n);
	if (!slots) { if (err) *err = oom; return false; }
	uint64_t m = be.select((const uint32_t *)NULL, act, slots, n);
	be.release(rank); be.release(act);
	if (m < n) {                                                      /* keep only what the rounds need */
		uint32_t *sl = be.template alloc<uint32_t>(m);
		if (!sl) { if (err) *err = oom; return false; }
		be.copy(sl, slots, m * 4);
		be.release(slots); slots = sl;
	}
	uint64_t h = BSA_H0;
	while (m > 0) {
		uint64_t *a0 = be.template alloc<uint64_t>(m), *a1 = be.template alloc<uint64_t>(m);
		uint32_t *b0 = be.template alloc<uint32_t>(m), *b1 = be.template alloc<uint32_t>(m), *nr = be.template alloc<uint32_t>(m), *slots2 = be.template alloc<uint32_t>(m);
		uint8_t *act2 = be.template alloc<uint8_t>(m);
		if (!a0 || !a1 || !b0 || !b1 || !nr || !slots2 || !act2) { if (err) *err = oom; return false; }
		be.each(m, BsaRoundKey{ slots, sa, isa, h, len, a0, b0 });
		uint64_t *k2s = nullptr; uint32_t *sfxs = nullptr;
		be.sort_pairs(a0, a1, b0, b1, m, 64, &k2s, &sfxs);
		be.each(m, BsaRoundHeads{ k2s, sfxs, slots, sa, nr });
		be.max_scan(nr, m);
		be.each(m, BsaRoundIsa{ sfxs, nr, isa });
		be.each(m, BsaRoundActive{ nr, slots, m, act2 });
		const uint64_t m2 = be.select(slots, act2, slots2, m);
		be.release(a0); be.release(a1); be.release(b0); be.release(b1); be.release(nr); be.release(act2); be.release(slots);
		slots = slots2; m = m2;
		be.mark("sort: doubling round");
		h *= 2;
		if (h > 2 * n + BSA_H0) { if (err) *err = "internal error: suffix sort did not converge"; return false; }
	}
	be.release(slots); be.release(isa);
	*sa_out = sa;
	return true;
}

/* ---- part 2: what buildToDisk reads off the suffix array -------------------------------------------------------------- */

/* The mirror index's text: every record (unambiguous stretch) reversed in place (REF_READ_REVERSE_EACH, ref_read.h:247-253). */
struct BsaMirror {
	const uint8_t *in; uint8_t *out; const uint64_t *recStart; uint32_t nrec;     /* recStart[0..nrec]: where each record's characters begin in the text */
	BSA_FN void operator()(uint64_t i) const {
		uint32_t lo = 0, hi = nrec;                                   /* the record that holds i: recStart[lo] <= i < recStart[lo + 1] */
		while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (recStart[mid] <= i) lo = mid; else hi = mid; }
		out[recStart[lo] + (recStart[lo + 1] - 1 - i)] = in[i];
	}
};
/* X.4.ebwt: the unambiguous characters, 2 bits each, first character in the low bits (filebuf.h:537-590) */
struct BsaPack2 {
	const uint8_t *s; uint64_t len; uint8_t *out;
	BSA_FN void operator()(uint64_t i) const {
		uint32_t b = 0;
		for (uint32_t j = 0; j < 4; j++) { const uint64_t p = 4 * i + j; if (p < len) b |= (uint32_t)s[p] << (2 * j); }
		out[i] = (uint8_t)b;
	}
};
/* per row (ebwt.h:4119-4185) */
struct BsaExtract {
	const uint8_t *s; const uint32_t *sa; uint64_t len; uint32_t offMask; int offRate; uint32_t K;
	uint8_t *bwt; uint32_t *offs, *zoff, *shortRow;                   /* shortRow[j] = row of the suffix of length j < K */
	BSA_FN void operator()(uint64_t row) const {
		const uint32_t p = sa[row];
		bwt[row] = p ? s[p - 1] : 0;                                  /* the row of suffix 0 is stored as an 'A' ... */
		if (p == 0) *zoff = (uint32_t)row;
		if ((row & offMask) == 0) offs[row >> offRate] = p;
		if (len - p < K) shortRow[len - p] = (uint32_t)row;
	}
};
/* characters per side of 224 rows: rows past the text are padding 'A's and are counted, the row of suffix 0 is not (ebwt.h:4197-4201) */
struct BsaSideCount {
	const uint8_t *bwt; uint64_t len; const uint32_t *zoff; uint32_t *c0, *c1, *c2, *c3;
	BSA_FN void operator()(uint64_t side) const {
		const uint64_t base = side * 224, z = *zoff;
		uint32_t n1 = 0, n2 = 0, n3 = 0;
		for (uint32_t k = 0; k < 224; k++) {
			const uint64_t row = base + k;
			if (row > len) break;
			const uint32_t ch = bwt[row];
			n1 += ch == 1; n2 += ch == 2; n3 += ch == 3;
		}
		uint32_t n0 = 224 - n1 - n2 - n3;                             /* the As of the text rows + the padding rows */
		if (z >= base && z < base + 224) n0--;
		c0[side] = n0; c1[side] = n1; c2[side] = n2; c3[side] = n3;
	}
};
/* one byte of the ebwt[] image (buildToDisk ebwt.h:4203-4281): side pairs of 2 x 64 bytes; the even ("backward") side holds its
 * 224 rows from its last byte to its first, high bit pair first, then u32 occ[A], occ[C] counted up to the end of that side (the
 * pair's midpoint); the odd ("forward") side holds its rows in natural order, then occ[G], occ[T] of that same midpoint. */
struct BsaPackSides {
	const uint8_t *bwt; uint64_t len; const uint32_t *i0, *i1, *i2, *i3; uint8_t *out;    /* i*: inclusive prefix sums of the side counts */
	BSA_FN uint32_t ch(uint64_t row) const { return row <= len ? bwt[row] : 0u; }
	BSA_FN void operator()(uint64_t B) const {
		const uint64_t side = B >> 6; const uint32_t o = (uint32_t)(B & 63); const bool fw = side & 1;
		uint32_t v;
		if (o < 56) {
			const uint64_t r0 = side * 224 + 4ull * (fw ? o : 55 - o);
			v = fw ? (ch(r0) | (ch(r0 + 1) << 2) | (ch(r0 + 2) << 4) | (ch(r0 + 3) << 6))
			       : (ch(r0 + 3) | (ch(r0 + 2) << 2) | (ch(r0 + 1) << 4) | (ch(r0) << 6));
		} else {
			const uint32_t w = (o - 56) >> 2, b = (o - 56) & 3;
			const uint32_t word = fw ? (w ? i3[side - 1] : i2[side - 1]) : (w ? i1[side] : i0[side]);
			v = (word >> (8 * b)) & 0xffu;
		}
		out[B] = (uint8_t)v;
	}
};
/* ftab histogram (ebwt.h:4143-4174): hist[kmer + 1] = how many suffixes start with that k-mer */
struct BsaKmerHist {
	const uint8_t *s; uint32_t K; uint32_t *hist;
	BSA_FN void operator()(uint64_t i) const {
		uint32_t v = 0;
		for (uint32_t j = 0; j < K; j++) v = (v << 2) | s[i + j];
		BSA_ATOMIC_INC(hist + (size_t)v + 1);
	}
};

/* One index (forward or mirror) from its device-resident text: the pieces of X.1.ebwt / X.2.ebwt that depend on the suffix array. */
template <class B>
static bool bt_build_index_parts(B &be, const uint8_t *s, uint32_t len, const BtBuildParams &P, BtIndexParts *out, std::string *err) {
	const char *oom = "Error: out of device memory while building the index";
	const uint64_t n = (uint64_t)len + 1;
	uint32_t *sa = nullptr;
	if (!bt_suffix_sort(be, s, len, &sa, err)) return false;
	const uint32_t K = (uint32_t)P.ftabChars;
	const uint64_t offsLen = (n + (1ull << P.offRate) - 1) >> P.offRate;
	uint8_t *bwt = be.template alloc<uint8_t>(n);
	uint32_t *offs = be.template alloc<uint32_t>(offsLen), *small = be.template alloc<uint32_t>(K + 1);
	if (!bwt || !offs || !small) { if (err) *err = oom; return false; }
	be.each(n, BsaExtract{ s, sa, len, (uint32_t)((1u << P.offRate) - 1), P.offRate, K, bwt, offs, small + K, small });
	out->offs.resize(offsLen);
	be.download(out->offs.data(), offs, offsLen * 4);
	be.release(offs);
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
			uint32_t p = 0; uint8_t km[16];
			be.download(&p, sa + next, 4);
			be.download(km, s + p, K);                                /* (a suffix of at least K characters: p + K <= len) */
			kmer = 0;
			for (uint32_t i = 0; i < K; i++) kmer = (kmer << 2) | km[i];
		}
		out->absorb.push_back({ kmer, (uint32_t)(b - a + 1) });
		a = b + 1;
	}
	be.release(sa);
	be.mark("BWT, SA sample");
	/* sides */
	const uint32_t sideBwtSz = 56;
	const uint32_t bwtSz = len / 4 + 1;
	const uint64_t numSides = 2ull * ((bwtSz + 2 * sideBwtSz - 1) / (2 * sideBwtSz)), ebwtTotLen = numSides * 64;
	uint32_t *c[4];
	for (int k = 0; k < 4; k++) { c[k] = be.template alloc<uint32_t>(numSides); if (!c[k]) { if (err) *err = oom; return false; } }
	be.each(numSides, BsaSideCount{ bwt, len, small + K, c[0], c[1], c[2], c[3] });
	for (int k = 0; k < 4; k++) be.sum_scan(c[k], numSides);
	uint8_t *img = be.template alloc<uint8_t>(ebwtTotLen);
	if (!img) { if (err) *err = oom; return false; }
	be.each(ebwtTotLen, BsaPackSides{ bwt, len, c[0], c[1], c[2], c[3], img });
	out->ebwt.resize(ebwtTotLen);
	be.download(out->ebwt.data(), img, ebwtTotLen);
	be.release(img); be.release(bwt); be.release(small);
	be.mark("side pairs + occ");
	/* fchr (ebwt.h:4296-4315): the BWT holds every character of the text once, plus the padding As */
	uint32_t tot[4];
	for (int k = 0; k < 4; k++) { be.download(&tot[k], c[k] + (numSides - 1), 4); be.release(c[k]); }
	tot[0] -= (uint32_t)(numSides * 224 - n);
	out->fchr[0] = 0;
	for (int k = 0; k < 4; k++) out->fchr[k + 1] = out->fchr[k] + tot[k];
	/* ftab histogram */
	const uint64_t ftabLen = (1ull << (2 * K)) + 1;
	uint32_t *hist = be.template alloc<uint32_t>(ftabLen);
	if (!hist) { if (err) *err = oom; return false; }
	be.zero(hist, ftabLen * 4);
	if (len >= K) be.each((uint64_t)len - K + 1, BsaKmerHist{ s, K, hist });
	out->ftab.resize(ftabLen);
	be.download(out->ftab.data(), hist, ftabLen * 4);
	be.release(hist);
	be.mark("ftab histogram");
	return true;
}

/* The whole build from the parsed reference: X.3/X.4, then the forward and the mirror index. */
template <class B>
static bool bt_build_all_on(B &be, const BtRefInfo &R, const std::string &base, const BtBuildParams &P, std::string &err) {
	const char *oom = "Error: out of device memory while building the index";
	const uint64_t len = R.textLen;
	if (len == 0 || len > 0xfffffffeull) { err = "Error: the joined reference must have 1 .. 2^32-2 characters"; return false; }
	uint8_t *text = be.template alloc<uint8_t>(len);
	if (!text) { err = oom; return false; }
	be.upload(text, R.text, len);
	{
		const uint64_t n4 = (len + 3) / 4;
		uint8_t *p4 = be.template alloc<uint8_t>(n4);
		if (!p4) { err = oom; return false; }
		be.each(n4, BsaPack2{ text, len, p4 });
		std::vector<uint8_t> o4((size_t)n4);
		be.download(o4.data(), p4, n4);
		be.release(p4);
		if (!be.ok(&err)) return false;
		if (!bt_build_write_ref(base, R, o4, err)) return false;
		be.mark("text uploaded, X.3 / X.4 written");
	}
	for (int mirror = 0; mirror < 2; mirror++) {
		const uint8_t *s = text; uint8_t *rev = nullptr;
		if (mirror) {
			std::vector<uint64_t> st;
			uint64_t at = 0;
			for (const BtRefRecord &r : R.recs) if (r.len) { st.push_back(at); at += r.len; }
			st.push_back(at);
			uint64_t *d_st = be.template alloc<uint64_t>(st.size());
			rev = be.template alloc<uint8_t>(len);
			if (!d_st || !rev) { err = oom; return false; }
			be.upload(d_st, st.data(), st.size() * 8);
			be.each(len, BsaMirror{ text, rev, d_st, (uint32_t)(st.size() - 1) });
			be.release(d_st);
			be.release(text); text = nullptr;
			s = rev;
		}
		BtIndexParts S;
		if (!bt_build_index_parts(be, s, (uint32_t)len, P, &S, &err)) return false;
		if (!be.ok(&err)) return false;
		if (rev) be.release(rev);
		const std::string b = base + (mirror ? ".rev" : "");
		if (!bt_build_write_index(b + ".1.ebwt", b + ".2.ebwt", R, (uint32_t)len, P, S, err)) return false;
		be.mark(mirror ? "mirror index files written" : "forward index files written");
	}
	if (text) be.release(text);
	return true;
}
