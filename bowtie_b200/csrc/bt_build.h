/*
 * bt_build.h — index construction: everything of `bowtie-build` except the suffix sort.
 *
 * the hg19-scale configurations need an index builder that does not take CPU-hours.  The reference builds
 * X.{1,2,3,4}.ebwt and X.rev.{1,2}.ebwt in ebwt_build.cpp:303-480 (driver), ref_read.cpp:10-141,202-273 (FASTA -> records),
 * ebwt.h:3825-3983 (joinToDisk), 582-611 (szsToDisk), 3602-3665 (header) and 3985-4388 (buildToDisk: BWT sides, occ words,
 * SA sample, ftab/eftab).  Here the split is:
 *   host (this file, plain C++): FASTA -> unambiguous stretches ("records"), joined text, X.3/X.4, file headers, names, the
 *        ftab histogram, packing of the BWT into side pairs with their occ words, writing the files;
 *   device (bt_build_sa.cuh): the suffix sort of the joined text and what is read off it row by row — the BWT character of
 *        every row, the row of suffix 0 (zOff), every 2^offRate-th suffix-array entry, and where the suffixes shorter than
 *        ftabChars fall.  The host part sees it through BtSuffixOracle and does not care who sorted.
 * Files are byte-identical to the reference's for the same input and -o/-t (the parity test); construction-only options
 * (--bmax, --dcv, --threads, --seed, --packed) have no counterpart because they do not change the output.
 *
 * Suffix order (multikey_qsort.h:25-28, blockwise_sa.h): the end of the text compares GREATER than every character, so
 * the empty suffix is the last row and a suffix is greater than any suffix it is a proper prefix of.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>

struct BtRefRecord { uint32_t off, len; uint8_t first; };          /* RefRecord (ref_read.h:57-88): `off` gap characters, then `len` unambiguous ones */

struct BtRefInfo {
	std::vector<BtRefRecord> recs;
	std::vector<uint32_t> plens;                                   /* length of every sequence that has unambiguous characters, gaps included */
	std::vector<std::string> names;
	std::vector<uint8_t> text;                                     /* the joined unambiguous characters, codes 0..3, in file order */
};

struct BtBuildParams { int offRate = 5, ftabChars = 10, lineRate = 6, linesPerSide = 1; bool nsToAs = false; };

/* What the suffix sort has to deliver for one text (length len, rows 0..len). */
struct BtSuffixResult {
	std::vector<uint8_t> bwt;                                      /* len + 1: text[SA[row] - 1], 0 for the row of suffix 0 */
	uint32_t zOff = 0;                                             /* the row of suffix 0 */
	std::vector<uint32_t> offs;                                    /* SA[row] for every row with the low offRate bits clear */
	/* suffixes shorter than ftabChars ("absorbed" into the ftab transition that follows them, ebwt.h:4146-4174): for each run
	 * of such rows, the ftabChars-mer of the next longer suffix in row order (or 4^ftabChars if none follows) and the run length */
	std::vector<std::pair<uint32_t, uint32_t>> absorb;
};
typedef bool (*BtSuffixOracle)(const uint8_t *text, uint32_t len, int offRate, int ftabChars, BtSuffixResult *out, void *ctx, std::string *err);

static inline int bt_dna_cat(int c) {                              /* dna4Cat (alphabet.cpp:3-25): 1 = ACGT, 2 = IUPAC ambiguity code or '-', 0 = anything else */
	switch (c) {
	case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1;
	case 'B': case 'D': case 'H': case 'K': case 'M': case 'N': case 'R': case 'S': case 'V': case 'W': case 'X': case 'Y':
	case 'b': case 'd': case 'h': case 'k': case 'm': case 'n': case 'r': case 's': case 'v': case 'w': case 'x': case 'y': case '-': return 2;
	default: return 0;
	}
}
static inline uint8_t bt_dna_code(int c) { switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; } }

/* FASTA -> records.  One pass instead of the reference's two (fastaRefReadSizes, then fastaRefReadAppend per record); the
 * cases where those two disagree with each other or with common sense — empty files, sequences without a single unambiguous
 * character, '#' comment lines — are refused rather than imitated. */
static inline bool bt_build_read_fasta(const std::vector<std::string> &files, bool nsToAs, BtRefInfo &R, std::string &err) {
	if (nsToAs) { err = "Error: --ntoa is not supported"; return false; }   /* (the reference's two passes treat it differently) */
	R.recs.clear(); R.plens.clear(); R.names.clear(); R.text.clear();
	uint64_t unambigTot = 0;
	for (const std::string &fn : files) {
		gzFile f = gzopen(fn.c_str(), "rb");
		if (!f) { err = "Error: could not open " + fn; return false; }
		gzbuffer(f, 1 << 20);
		std::vector<char> buf(1 << 22);
		size_t len = 0, pos = 0;
		int lastByte = -1;
		auto get = [&]() -> int { if (pos >= len) { int n = gzread(f, buf.data(), (unsigned)buf.size()); if (n <= 0) return -1; len = (size_t)n; pos = 0; } return lastByte = (unsigned char)buf[pos++]; };
		int c = get();
		while (c >= 0 && (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f')) c = get();
		if (c < 0) { gzclose(f); err = "Error: empty reference file " + fn; return false; }
		if (c != '>') { gzclose(f); err = "Reference file does not seem to be a FASTA file"; return false; }
		while (c == '>') {
			std::string name;
			c = get();
			while (c >= 0 && c != '\n' && c != '\r') { name.push_back((char)c); c = get(); }
			while (c == '\n' || c == '\r') c = get();
			if (c == '#') { gzclose(f); err = "Error: comment lines in the reference are not supported"; return false; }
			/* the sequence: alternating runs of gap characters and unambiguous characters up to the next '>' */
			uint32_t both = 0, unambig = 0, off = 0, run = 0; bool first = true;
			const size_t rec0 = R.recs.size();
			while (c >= 0 && c != '>') {
				int cat = bt_dna_cat(c);
				if (nsToAs && cat == 2) { c = 'A'; cat = 1; }
				if (cat == 1) { R.text.push_back(bt_dna_code(c)); run++; }
				else if (cat == 2) {
					if (run) { R.recs.push_back({ off, run, (uint8_t)first }); first = false; both += off + run; unambig += run; off = 0; run = 0; }
					off++;
				}
				c = get();
			}
			if (run) { R.recs.push_back({ off, run, (uint8_t)first }); first = false; both += off + run; unambig += run; off = 0; }
			else if (off == 1 && !first && c < 0 && bt_dna_cat(lastByte) == 2) {
				/* a single gap character as the very last byte of a file is consumed with the stretch before it and the reader is
				 * at EOF before a record can be made of it (ref_read.cpp:123-127,220): it is in neither the records nor plen */
			} else if (off) {
				if (first) { gzclose(f); err = "Error: reference sequence \"" + name + "\" has no unambiguous characters (not supported)"; return false; }
				R.recs.push_back({ off, 0, 0 }); both += off;                /* a trailing gap is a record of its own */
			}
			if (R.recs.size() == rec0) { gzclose(f); err = "Error: empty reference sequence \"" + name + "\" (not supported)"; return false; }
			unambigTot += unambig;
			if (unambigTot > 0xfffffffeull) { gzclose(f); err = "Error: Reference sequence has more than 2^32-1 characters!  Please try to\nbuild a large index instead using the appropiate options."; return false; }
			if (name.empty()) name = std::to_string(R.names.size());     /* ebwt.h:3904-3909 */
			R.names.push_back(name);
			R.plens.push_back(both);
		}
		gzclose(f);
	}
	if (R.text.empty()) { err = "Error: No unambiguous stretches of characters in the input.  Aborting..."; return false; }
	return true;
}

static inline void bt_put_u32(std::vector<uint8_t> &o, uint32_t v) { o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)(v >> 16)); o.push_back((uint8_t)(v >> 24)); }
static inline bool bt_write_file(const std::string &path, const std::vector<uint8_t> &a, const std::vector<uint8_t> *b, std::string &err) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) { err = "Could not open index file for writing: \"" + path + "\"\nPlease make sure the directory exists and that permissions allow writing by\nBowtie."; return false; }
	bool ok = a.empty() || fwrite(a.data(), 1, a.size(), f) == a.size();
	if (ok && b && !b->empty()) ok = fwrite(b->data(), 1, b->size(), f) == b->size();
	ok = (fclose(f) == 0) && ok;
	if (!ok) err = "Error writing " + path;
	return ok;
}

/* X.3.ebwt (the records) and X.4.ebwt (the unambiguous characters, 2 bits each, first character in the low bits):
 * ebwt_build.cpp:361-391, filebuf.h:537-590 */
static inline bool bt_build_write_ref(const std::string &base, const BtRefInfo &R, std::string &err) {
	std::vector<uint8_t> o3;
	bt_put_u32(o3, 1); bt_put_u32(o3, (uint32_t)R.recs.size());
	for (const BtRefRecord &r : R.recs) { bt_put_u32(o3, r.off); bt_put_u32(o3, r.len); o3.push_back(r.first); }
	std::vector<uint8_t> o4((R.text.size() + 3) / 4, 0);
	for (size_t i = 0; i < R.text.size(); i++) o4[i >> 2] |= (uint8_t)(R.text[i] << ((i & 3) << 1));
	return bt_write_file(base + ".3.ebwt", o3, NULL, err) && bt_write_file(base + ".4.ebwt", o4, NULL, err);
}

/* The text of one index: the joined records as they are (forward index) or every record reversed in place (the mirror
 * index, REF_READ_REVERSE_EACH: ref_read.h:247-253, ebwt_build.cpp:77) */
static inline void bt_build_text(const BtRefInfo &R, bool mirror, std::vector<uint8_t> &s) {
	s = R.text;
	if (!mirror) return;
	size_t at = 0;
	for (const BtRefRecord &r : R.recs) {
		for (size_t i = 0, j = r.len; i + 1 < j; i++, j--) { const uint8_t t = s[at + i]; s[at + i] = s[at + j - 1]; s[at + j - 1] = t; }
		at += r.len;
	}
}

/* X.1.ebwt / X.2.ebwt (or X.rev.1 / X.rev.2) from the text and what the suffix sort delivered. */
static inline bool bt_build_write_index(const std::string &path1, const std::string &path2, const BtRefInfo &R, const std::vector<uint8_t> &s,
                                        const BtBuildParams &P, const BtSuffixResult &S, std::string &err) {
	const uint32_t len = (uint32_t)s.size();
	/* EbwtParams::init (ebwt.h:138-184) */
	const uint32_t sideSz = (1u << P.lineRate) * (uint32_t)P.linesPerSide, sideBwtSz = sideSz - 8, sideBwtLen = sideBwtSz * 4;
	const uint32_t bwtSz = len / 4 + 1;
	const uint32_t numSidePairs = (bwtSz + 2 * sideBwtSz - 1) / (2 * sideBwtSz), numSides = numSidePairs * 2;
	const uint64_t ebwtTotLen = (uint64_t)numSides * sideSz;
	const uint32_t offsLen = (uint32_t)(((uint64_t)len + 1 + (1ull << P.offRate) - 1) >> P.offRate);
	const uint64_t ftabLen = (1ull << (2 * P.ftabChars)) + 1;
	const uint32_t eftabLen = (uint32_t)P.ftabChars * 2;
	if (S.bwt.size() != (size_t)len + 1 || S.offs.size() != offsLen) { err = "internal error: suffix-sort result has the wrong shape"; return false; }

	std::vector<uint8_t> o1;
	o1.reserve((size_t)ebwtTotLen + ftabLen * 4 + 4096);
	/* header (ebwt.h:3611-3624) */
	bt_put_u32(o1, 1); bt_put_u32(o1, len); bt_put_u32(o1, (uint32_t)P.lineRate); bt_put_u32(o1, (uint32_t)P.linesPerSide);
	bt_put_u32(o1, (uint32_t)P.offRate); bt_put_u32(o1, (uint32_t)P.ftabChars); bt_put_u32(o1, (uint32_t)-1);
	/* plen[], rstarts[] (joinToDisk ebwt.h:3865-3881, szsToDisk ebwt.h:582-611; per-record reversal leaves them as they are) */
	bt_put_u32(o1, (uint32_t)R.plens.size());
	for (uint32_t p : R.plens) bt_put_u32(o1, p);
	uint32_t nFrag = 0;
	for (const BtRefRecord &r : R.recs) if (r.len) nFrag++;
	bt_put_u32(o1, nFrag);
	{
		uint32_t seq = 0, off = 0, tot = 0;
		for (const BtRefRecord &r : R.recs) {
			if (r.len == 0) continue;
			if (r.first) { off = 0; seq++; }
			off += r.off;
			bt_put_u32(o1, tot); bt_put_u32(o1, seq - 1); bt_put_u32(o1, off);
			tot += r.len; off += r.len;
		}
	}
	/* the BWT in side pairs (buildToDisk ebwt.h:4101-4281): a backward side — filled from its last byte to its first, high
	 * bit pair first — then u32 occ[A], occ[C] at the pair's midpoint; a forward side in natural order, then occ[G], occ[T] of
	 * that same midpoint.  Rows past the text are padding 'A's and are counted; the row of suffix 0 holds an 'A' that is not. */
	{
		const size_t base = o1.size();
		o1.resize(base + (size_t)ebwtTotLen, 0);
		uint8_t *e = o1.data() + base;
		uint32_t occ[4] = { 0, 0, 0, 0 }, save[2] = { 0, 0 };
		uint64_t row = 0;
		for (uint32_t side = 0; side < numSides; side++) {
			uint8_t *sd = e + (size_t)side * sideSz;
			const bool fw = side & 1;
			for (uint32_t k = 0; k < sideBwtLen; k++, row++) {
				uint32_t ch = 0; bool count = true;
				if (row <= len) { ch = S.bwt[row]; if (row == S.zOff) { ch = 0; count = false; } }
				if (count) occ[ch]++;
				const uint32_t byte = k >> 2, bp = k & 3;
				if (fw) sd[byte] |= (uint8_t)(ch << (bp << 1));
				else sd[sideBwtSz - 1 - byte] |= (uint8_t)(ch << ((3 - bp) << 1));
			}
			const uint32_t w0 = fw ? save[0] : occ[0], w1 = fw ? save[1] : occ[1];
			for (int b = 0; b < 4; b++) { sd[sideBwtSz + b] = (uint8_t)(w0 >> (8 * b)); sd[sideBwtSz + 4 + b] = (uint8_t)(w1 >> (8 * b)); }
			if (!fw) { save[0] = occ[2]; save[1] = occ[3]; }
		}
	}
	bt_put_u32(o1, S.zOff);
	/* fchr (ebwt.h:4296-4315): characters of the text, cumulative */
	{
		uint32_t cnt[4] = { 0, 0, 0, 0 };
		for (uint32_t i = 0; i < len; i++) cnt[s[i]]++;
		uint32_t acc = 0;
		bt_put_u32(o1, 0);
		for (int i = 0; i < 4; i++) { acc += cnt[i]; bt_put_u32(o1, acc); }
	}
	/* ftab / eftab (ebwt.h:4143-4174, 4317-4352): ftab[k] = first row of the suffixes that start with k-mer k; where suffixes
	 * shorter than ftabChars sit between two k-mers' blocks the entry points (x ^ 0xffffffff) at an eftab pair (lo, hi) */
	{
		std::vector<uint32_t> ftab((size_t)ftabLen, 0);
		const uint32_t K = (uint32_t)P.ftabChars;
		if (len >= K) {
			const uint64_t mask = (1ull << (2 * K)) - 1;
			uint64_t v = 0;
			for (uint32_t i = 0; i < K - 1; i++) v = (v << 2) | s[i];
			for (uint32_t i = K - 1; i < len; i++) { v = ((v << 2) | s[i]) & mask; ftab[(size_t)v + 1]++; }
		}
		std::vector<uint32_t> absorb_at, absorb_n;
		for (auto &a : S.absorb) { absorb_at.push_back(a.first); absorb_n.push_back(a.second); }
		std::vector<uint32_t> eftab(eftabLen, 0);
		uint32_t ecur = 0, hiPrev = 0;                                 /* hiPrev = ftabHi(i - 1) */
		size_t ai = 0;
		for (uint64_t i = 1; i < ftabLen; i++) {
			const uint32_t lo = ftab[(size_t)i] + hiPrev;
			while (ai < absorb_at.size() && absorb_at[ai] < i) ai++;       /* (sorted by k-mer; entries at 0 cannot occur) */
			if (ai < absorb_at.size() && absorb_at[ai] == i) {
				const uint32_t hi = lo + absorb_n[ai];
				if (ecur * 2 + 1 >= eftabLen) { err = "internal error: eftab overflow"; return false; }
				eftab[ecur * 2] = lo; eftab[ecur * 2 + 1] = hi;
				ftab[(size_t)i] = ecur ^ 0xffffffffu; ecur++;
				hiPrev = hi;
			} else { ftab[(size_t)i] = lo; hiPrev = lo; }
		}
		if (hiPrev != len + 1) { err = "internal error: ftab does not cover the suffix array"; return false; }
		const size_t at = o1.size();
		o1.resize(at + ((size_t)ftabLen + eftabLen) * 4);
		memcpy(o1.data() + at, ftab.data(), (size_t)ftabLen * 4);            /* little-endian host */
		memcpy(o1.data() + at + (size_t)ftabLen * 4, eftab.data(), (size_t)eftabLen * 4);
	}
	/* names (ebwt.h:803-811) */
	for (const std::string &n : R.names) { o1.insert(o1.end(), n.begin(), n.end()); o1.push_back('\n'); }
	o1.push_back(0);

	std::vector<uint8_t> o2;
	bt_put_u32(o2, 1);
	const size_t at2 = o2.size();
	o2.resize(at2 + (size_t)offsLen * 4);
	memcpy(o2.data() + at2, S.offs.data(), (size_t)offsLen * 4);
	return bt_write_file(path1, o1, NULL, err) && bt_write_file(path2, o2, NULL, err);
}

/* Whole build: X.3/X.4, then the forward and the mirror index. */
static inline bool bt_build_all(const std::vector<std::string> &fasta, const std::string &base, const BtBuildParams &P, BtSuffixOracle sort, void *ctx, std::string &err) {
	BtRefInfo R;
	if (!bt_build_read_fasta(fasta, P.nsToAs, R, err)) return false;
	if (!bt_build_write_ref(base, R, err)) return false;
	for (int mirror = 0; mirror < 2; mirror++) {
		std::vector<uint8_t> s;
		bt_build_text(R, mirror != 0, s);
		BtSuffixResult S;
		if (!sort(s.data(), (uint32_t)s.size(), P.offRate, P.ftabChars, &S, ctx, &err)) return false;
		const std::string b = base + (mirror ? ".rev" : "");
		if (!bt_build_write_index(b + ".1.ebwt", b + ".2.ebwt", R, s, P, S, err)) return false;
	}
	return true;
}
