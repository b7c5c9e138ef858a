/*
 * bt_io.cuh — read ingest and hit formatting on the device, written against a small backend interface
 * like bt_build_sa.cuh: per-element functors + scans + one compaction, so that the same code runs on the GPU (bt_io.cu) and, for
 * the CPU test suite, over a host backend (tests/host_emu).
 *
 * f1 — FASTQ text -> the fields of Read the search consumes (pat.cpp:858-975 FastqPatternSource::parse for well-formed records;
 *      read.h:118-132; genRandSeed pat.cpp:21-57): newline positions (compaction) -> record table (4 lines per record, validated)
 *      -> base codes, Phred+33 qualities, offsets, per-read seeds.  Anything that is not a plain 4-line record with equal
 *      sequence / quality lengths, letters or '.' in the sequence and qualities >= '!' is NOT handled here: the parser reports the
 *      first irregular record and the caller hands everything from there on to the host parser, whose behaviour is the specification.
 * f2 — hit records -> output text: the default format (VerboseHitSink::append, hit.cpp:176-240) and SAM (SAMHitSink::append /
 *      reportUnOrMax, sam.cpp:57-257), unpaired reads, HitSinkPerThread::finishRead's -k / -m arithmetic (hit.h:741-786), in read
 *      order: a length pass, a prefix sum, a write pass (ordered compaction).
 */
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BIO_FN __host__ __device__ __forceinline__
#else
#define BIO_FN inline
#endif
#if defined(__CUDA_ARCH__)
#define BIO_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define BIO_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define BIO_ATOMIC_OR(p, v) atomicOr((p), (v))
#define BIO_ATOMIC_ADD64(p, v) atomicAdd((p), (unsigned long long)(v))
#else
#define BIO_ATOMIC_MIN(p, v) do { if ((v) < *(p)) *(p) = (v); } while (0)
#define BIO_ATOMIC_MAX(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define BIO_ATOMIC_OR(p, v) (*(p) |= (v))
#define BIO_ATOMIC_ADD64(p, v) (*(p) += (unsigned long long)(v))
#endif

#define BIO_HIT_HDR 5

struct BioRec { uint32_t name_off, name_len, seq_off, qual_off, len; };      /* one FASTQ record inside the text chunk */

BIO_FN uint32_t bio_alpha_code(uint32_t c) {                                   /* asc2dna for letters and '.', 255 for anything else (alphabet.cpp) */
	switch (c) {
	case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
	case '.': return 4;
	default: return ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) ? 4u : 255u;
	}
}

/* ---- f1 ------------------------------------------------------------------------------------------------------------------ */
struct BioMarkNl { const char *text; uint8_t *flag; BIO_FN void operator()(uint64_t i) const { flag[i] = text[i] == '\n'; } };

/* record r = lines 4r .. 4r+3 of the chunk (which starts at a record) */
struct BioRecords {
	const char *text; const uint32_t *nl; BioRec *rec; uint32_t *len; uint32_t *firstBad, *maxLen;
	BIO_FN void operator()(uint64_t r) const {
		const uint32_t s0 = r ? nl[4 * r - 1] + 1 : 0, e0 = nl[4 * r], e1 = nl[4 * r + 1], e2 = nl[4 * r + 2], e3 = nl[4 * r + 3];
		BioRec x;
		x.name_off = s0 + 1; x.name_len = e0 - s0 - 1; x.seq_off = e0 + 1; x.len = e1 - e0 - 1; x.qual_off = e2 + 1;
		const uint32_t qlen = e3 - e2 - 1;
		bool ok = e0 > s0 && text[s0] == '@' && text[e1 + 1] == '+' && x.len == qlen && x.len >= 4 && x.len < 1024;   /* (shorter reads: the host path prints the reference's warnings) */
		if (ok && (text[e0 - 1] == '\r' || text[e1 - 1] == '\r' || text[e3 - 1] == '\r')) ok = false;
		if (ok && x.name_len == 0) ok = false;                                 /* (the reference names such reads by their ordinal: host path) */
		rec[r] = x; len[r] = ok ? x.len : 0;
		if (!ok) BIO_ATOMIC_MIN(firstBad, (uint32_t)r); else BIO_ATOMIC_MAX(maxLen, x.len);
	}
};
struct BioAnyFlag { const uint32_t *flags; uint32_t *any; BIO_FN void operator()(uint64_t i) const { if (flags[i]) BIO_ATOMIC_OR(any, flags[i]); } };
struct BioWiden { const uint32_t *in; uint64_t *out; BIO_FN void operator()(uint64_t i) const { out[i] = in[i]; } };
struct BioConvert {
	const char *text; const BioRec *rec; const uint32_t *off; uint8_t *seq, *qual; uint32_t *seeds; uint32_t gseed; uint32_t *firstBad;
	BIO_FN void operator()(uint64_t r) const {
		const BioRec x = rec[r];
		uint32_t rseed = (gseed + 101u) * 59u * 61u * 67u * 71u * 73u * 79u * 83u;                /* genRandSeed (pat.cpp:21-57) */
		bool ok = true;
		const uint32_t o = off[r];
		for (uint32_t i = 0; i < x.len; i++) {
			const uint32_t code = bio_alpha_code((unsigned char)text[x.seq_off + i]);
			if (code == 255u) { ok = false; break; }
			seq[o + i] = (uint8_t)code;
			rseed ^= code << ((i & 15) << 1);
		}
		for (uint32_t i = 0; ok && i < x.len; i++) {
			const uint32_t q = (unsigned char)text[x.qual_off + i];
			if (q < 33) { ok = false; break; }
			qual[o + i] = (uint8_t)q;
			rseed ^= q << ((i & 3) << 3);
		}
		for (uint32_t i = 0; i < x.name_len; i++) {
			const uint32_t ch = (unsigned char)text[x.name_off + i];
			if (ch == '\r') ok = false;                                          /* (the reference's name ends at the first CR: host path) */
			rseed ^= ch << ((i & 3) << 3);
		}
		seeds[r] = rseed;
		if (!ok) BIO_ATOMIC_MIN(firstBad, (uint32_t)r);
	}
};

/* ---- f2 ------------------------------------------------------------------------------------------------------------------ */
struct BioFmt {                 /* the options of the two formats that the device path provides (everything else: host formatter) */
	uint32_t sam, khits, mhits, strata, noUnal, noQnameTrunc, offBase, mapq, slots, recWords;
};
struct BioNames { const char *buf; const uint32_t *off; };                       /* reference names as printed: buf[off[t], off[t+1]) */

BIO_FN uint32_t bio_digits(uint64_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; n++; } return n; }
BIO_FN uint32_t bio_put_uint(char *d, uint64_t v) {
	const uint32_t n = bio_digits(v);
	for (uint32_t i = n; i > 0; i--) { d[i - 1] = (char)('0' + v % 10); v /= 10; }
	return n;
}
BIO_FN uint32_t bio_qname_len(const char *text, const BioRec &x, uint32_t noTrunc) {
	if (noTrunc) return x.name_len;
	uint32_t n = 0;
	while (n < x.name_len) { const char c = text[x.name_off + n]; if (c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r') break; n++; }
	return n;
}
/* mismatches of one record sorted by offset from the 5' end (insertion sort; a record holds at most recWords - 5 of them) */
BIO_FN uint32_t bio_sorted_mms(const uint32_t *w, uint32_t cap, uint32_t *mm) {
	uint32_t n = w[4] < cap ? w[4] : cap;
	for (uint32_t i = 0; i < n; i++) {
		const uint32_t v = w[BIO_HIT_HDR + i];
		uint32_t j = i;
		while (j > 0 && (mm[j - 1] & 0xffffu) > (v & 0xffffu)) { mm[j] = mm[j - 1]; j--; }
		mm[j] = v;
	}
	return n;
}
#define BIO_MM_MAX 32

/* One functor for both passes: `out == NULL` computes the length of read r's output, else writes it at pos[r]. */
struct BioFormat {
	const char *text; const BioRec *rec; const uint8_t *seq; const uint32_t *off; const uint32_t *found, *hits; BioNames names; BioFmt f;
	uint32_t *lens; const uint32_t *pos; char *out; unsigned long long *cnt;     /* cnt: aligned, unaligned, maxed, reported */

	BIO_FN uint32_t put_seq(char *d, uint32_t r, uint32_t len, bool fw) const {
		const uint8_t *s = seq + off[r];
		if (d) { for (uint32_t i = 0; i < len; i++) { const uint32_t c = fw ? s[i] : s[len - 1 - i]; d[i] = "ACGTN"[fw ? c : (c < 4 ? (c ^ 3u) : 4u)]; } }
		return len;
	}
	BIO_FN uint32_t put_qual(char *d, const BioRec &x, bool fw) const {
		if (d) { for (uint32_t i = 0; i < x.len; i++) d[i] = text[x.qual_off + (fw ? i : x.len - 1 - i)]; }
		return x.len;
	}
	BIO_FN uint32_t put_ref(char *d, uint32_t tidx) const {
		const uint32_t a = names.off[tidx], b = names.off[tidx + 1];
		if (d) for (uint32_t i = a; i < b; i++) d[i - a] = names.buf[i];
		return b - a;
	}
	/* VerboseHitSink::append (hit.cpp:176-240) */
	BIO_FN uint32_t line_default(char *d, uint32_t r, const BioRec &x, const uint32_t *w, uint32_t oms) const {
		uint32_t n = 0;
#define PUTC(c) do { if (d) d[n] = (c); n++; } while (0)
		const bool fw = (w[3] >> 24) & 1;
		if (d) for (uint32_t i = 0; i < x.name_len; i++) d[n + i] = text[x.name_off + i];
		n += x.name_len; PUTC('\t'); PUTC(fw ? '+' : '-'); PUTC('\t');
		n += put_ref(d ? d + n : 0, w[0]); PUTC('\t');
		{ const uint64_t v = (uint64_t)w[1] + f.offBase; if (d) bio_put_uint(d + n, v); n += bio_digits(v); } PUTC('\t');
		n += put_seq(d ? d + n : 0, r, x.len, fw); PUTC('\t');
		n += put_qual(d ? d + n : 0, x, fw); PUTC('\t');
		if (d) bio_put_uint(d + n, oms); n += bio_digits(oms); PUTC('\t');
		uint32_t mm[BIO_MM_MAX];
		const uint32_t nmm = bio_sorted_mms(w, f.recWords - BIO_HIT_HDR < BIO_MM_MAX ? f.recWords - BIO_HIT_HDR : BIO_MM_MAX, mm);
		const uint8_t *s = seq + off[r];
		for (uint32_t i = 0; i < nmm; i++) {
			const uint32_t p = mm[i] & 0xffffu, refc = (mm[i] >> 16) & 3u;
			if (i) PUTC(',');
			if (d) bio_put_uint(d + n, p); n += bio_digits(p);
			const uint32_t c = s[p];
			PUTC(':'); PUTC("ACGT"[refc]); PUTC('>'); PUTC(fw ? "ACGTN"[c] : "ACGTN"[c < 4 ? (c ^ 3u) : 4u]);
		}
		PUTC('\n');
		return n;
	}
	/* SAMHitSink::append (sam.cpp:129-257), unpaired */
	BIO_FN uint32_t line_sam(char *d, uint32_t r, const BioRec &x, const uint32_t *w, uint32_t xms) const {
		uint32_t n = 0;
		const bool fw = (w[3] >> 24) & 1;
		const uint32_t ql = bio_qname_len(text, x, f.noQnameTrunc);
		if (d) for (uint32_t i = 0; i < ql; i++) d[n + i] = text[x.name_off + i];
		n += ql; PUTC('\t');
		if (fw) PUTC('0'); else { PUTC('1'); PUTC('6'); }
		PUTC('\t'); n += put_ref(d ? d + n : 0, w[0]); PUTC('\t');
		{ const uint64_t v = (uint64_t)w[1] + 1; if (d) bio_put_uint(d + n, v); n += bio_digits(v); } PUTC('\t');
		if (d) bio_put_uint(d + n, f.mapq); n += bio_digits(f.mapq); PUTC('\t');
		if (d) bio_put_uint(d + n, x.len); n += bio_digits(x.len); PUTC('M');
		PUTC('\t'); PUTC('*'); PUTC('\t'); PUTC('0'); PUTC('\t'); PUTC('0'); PUTC('\t');
		n += put_seq(d ? d + n : 0, r, x.len, fw); PUTC('\t');
		n += put_qual(d ? d + n : 0, x, fw);
		PUTC('\t'); PUTC('X'); PUTC('A'); PUTC(':'); PUTC('i'); PUTC(':');
		{ const uint32_t st = (w[3] >> 16) & 0xffu; if (d) bio_put_uint(d + n, st); n += bio_digits(st); }
		PUTC('\t'); PUTC('M'); PUTC('D'); PUTC(':'); PUTC('Z'); PUTC(':');
		uint32_t mm[BIO_MM_MAX];
		const uint32_t nmm = bio_sorted_mms(w, f.recWords - BIO_HIT_HDR < BIO_MM_MAX ? f.recWords - BIO_HIT_HDR : BIO_MM_MAX, mm);
		/* mismatch offsets count from the 5' end; MD runs along the reference: 5'->3' for a '+' hit, reversed for a '-' hit */
		uint32_t prev = 0, nm = 0;
		for (uint32_t k = 0; k < nmm; k++) {
			const uint32_t e = fw ? mm[k] : mm[nmm - 1 - k];
			uint32_t p = e & 0xffffu;
			if (p >= x.len) continue;
			if (!fw) p = x.len - 1 - p;
			const uint32_t run = p - prev;
			if (d) bio_put_uint(d + n, run); n += bio_digits(run);
			PUTC("ACGT"[(e >> 16) & 3u]);
			prev = p + 1; nm++;
		}
		{ const uint32_t run = x.len - prev; if (d) bio_put_uint(d + n, run); n += bio_digits(run); }
		PUTC('\t'); PUTC('N'); PUTC('M'); PUTC(':'); PUTC('i'); PUTC(':');
		if (d) bio_put_uint(d + n, nm); n += bio_digits(nm);
		if (xms > 0) { PUTC('\t'); PUTC('X'); PUTC('M'); PUTC(':'); PUTC('i'); PUTC(':'); if (d) bio_put_uint(d + n, xms); n += bio_digits(xms); }
		PUTC('\n');
		return n;
	}
	/* SAMHitSink::reportUnOrMax (sam.cpp:57-124), unpaired, un == true */
	BIO_FN uint32_t line_sam_unal(char *d, uint32_t r, const BioRec &x) const {
		uint32_t n = 0;
		const uint32_t ql = bio_qname_len(text, x, f.noQnameTrunc);
		if (d) for (uint32_t i = 0; i < ql; i++) d[n + i] = text[x.name_off + i];
		n += ql;
		const char *mid = "\t4\t*\t0\t0\t*\t*\t0\t0\t";
		for (uint32_t i = 0; mid[i]; i++) PUTC(mid[i]);
		n += put_seq(d ? d + n : 0, r, x.len, true); PUTC('\t');
		n += put_qual(d ? d + n : 0, x, true);
		const char *tail = "\tXM:i:0\n";
		for (uint32_t i = 0; tail[i]; i++) PUTC(tail[i]);
		return n;
	}
#undef PUTC
	BIO_FN void operator()(uint64_t r64) const {
		const uint32_t r = (uint32_t)r64;
		const BioRec x = rec[r];
		const uint32_t fnd = found[r];
		char *d = out ? out + pos[r] : 0;
		uint32_t n = 0;
		/* HitSinkPerThread::finishRead (hit.h:741-786) */
		const bool maxed = fnd > f.mhits, unal = fnd == 0;
		if (maxed || unal) {
			if (f.sam && !f.noUnal && unal) n += line_sam_unal(d, r, x);        /* (maxed reads: -M is not provided here; without it they print nothing) */
			if (!out) { if (maxed) BIO_ATOMIC_ADD64(cnt + 2, 1); else BIO_ATOMIC_ADD64(cnt + 1, 1); }
		} else {
			const uint32_t nrep = fnd < f.khits ? fnd : f.khits;
			for (uint32_t s = 0; s < nrep && s < f.slots; s++) {
				const uint32_t *w = hits + ((uint64_t)r * f.slots + s) * f.recWords;
				const uint32_t oms = f.strata ? fnd - 1 : w[2];                 /* NBestFirstStratHitSinkPerThread::finishReadImpl (hit.h:1099-1108) */
				n += f.sam ? line_sam(d ? d + n : 0, r, x, w, nrep) : line_default(d ? d + n : 0, r, x, w, oms);
			}
			if (!out) { BIO_ATOMIC_ADD64(cnt + 0, 1); BIO_ATOMIC_ADD64(cnt + 3, nrep); }
		}
		if (!out) lens[r] = n;
	}
};
