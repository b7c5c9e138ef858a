/*
 * bt_io_run.h — the driver of the device I/O path (f1 + search + f2) over a backend, shared by bt_io.cu (CUDA) and the test-only
 * host backend of tests/host_emu (same functors, std:: loops).  A backend provides
 *   void *alloc(bytes), release(p)                    "device" memory
 *   h2d(dst, src, bytes), d2h(dst, src, bytes)        copies on the object's stream (d2h: complete on return)
 *   zero(p, bytes)
 *   each(n, functor)
 *   uint64_t count_nl(text, n)                        number of '\n' in text[0, n)            (complete on return)
 *   void positions_nl(text, n, out)                   their offsets, ascending
 *   void excl_scan(a, n)                              in place, uint32
 *   int  align(pol, batch, hits)                      bt_context_align_device on the object's context + join
 *   void sync()
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "bt_io.cuh"
#include "../../include/bowtie_b200.h"

template <class T> struct BioBuf {
	T *p = nullptr; size_t cap = 0;
	template <class B> bool need(B &be, size_t n) {
		if (n <= cap) return true;
		if (p) be.release(p);
		p = nullptr; cap = 0;
		const size_t want = n + n / 4 + 64;
		p = (T *)be.alloc(want * sizeof(T));
		if (!p) return false;
		cap = want;
		return true;
	}
};

template <class B>
struct BioPipe {
	B be;
	BioBuf<char> text, out; BioBuf<uint32_t> nl, len, seeds, found, flags, hits, olen; BioBuf<BioRec> rec; BioBuf<uint64_t> offs; BioBuf<uint8_t> seq, qual;
	BioBuf<char> names; BioBuf<uint32_t> nameOff;
	uint32_t *d_misc = nullptr;            /* [0] firstBad */
	unsigned long long *d_cnt = nullptr;   /* 4 counters */
	std::vector<char> host_out;            /* (the CUDA backend replaces this with pinned memory through out_host()) */
	uint32_t nreads = 0, maxlen = 0, mm_cap = 0, slots = 0;
	uint64_t nbases = 0;
	bool names_full = false, have_names = false;
	std::string err;

	bool init() {
		d_misc = (uint32_t *)be.alloc(16); d_cnt = (unsigned long long *)be.alloc(4 * sizeof(unsigned long long));
		return d_misc && d_cnt;
	}
	void destroy() {
		void *ps[] = { text.p, out.p, nl.p, len.p, seeds.p, found.p, flags.p, hits.p, olen.p, rec.p, offs.p, seq.p, qual.p, names.p, nameOff.p, d_misc, d_cnt };
		for (void *p : ps) if (p) be.release(p);
	}
	/* reference names as the formats print them (up to the first whitespace unless --fullref) */
	bool set_names(const std::vector<std::string> &refnames, bool full) {
		std::string cat; std::vector<uint32_t> off(1, 0);
		for (const std::string &n : refnames) { size_t k = full ? n.size() : strcspn(n.c_str(), " \t"); cat.append(n, 0, k); off.push_back((uint32_t)cat.size()); }
		if (!names.need(be, cat.size() + 1) || !nameOff.need(be, off.size())) return false;
		be.h2d(names.p, cat.data(), cat.size()); be.h2d(nameOff.p, off.data(), off.size() * 4);
		names_full = full; have_names = true;
		return true;
	}
	/* f1.  Parses the complete, regular records at the start of text[0, nbytes) — at most max_reads, and never the chunk's last
	 * complete record (the caller keeps it for the next chunk or the host parser, which owns the end-of-file rules). */
	bool parse(const char *host_text, uint64_t nbytes, uint32_t gseed, uint32_t max_reads, uint32_t *n_out, uint64_t *consumed, int *irregular) {
		*n_out = 0; *consumed = 0; *irregular = 0; nreads = 0;
		if (nbytes == 0) return true;
		if (nbytes > 0xfffffff0ull) { err = "bt_io_parse_fastq: chunk too large"; return false; }
		if (!text.need(be, nbytes)) { err = "bt_io: out of device memory"; return false; }
		be.h2d(text.p, host_text, nbytes);
		const uint64_t m = be.count_nl(text.p, nbytes);
		if (m > nbytes / 2 + 4) { *irregular = 1; return true; }               /* (a regular record has 4 newlines in >= 9 bytes) */
		if (m < 8) return true;
		if (!nl.need(be, m)) { err = "bt_io: out of device memory"; return false; }
		be.positions_nl(text.p, nbytes, nl.p);
		uint64_t nrec = m / 4 - 1;
		if (nrec > max_reads) nrec = max_reads;
		if (nrec == 0) return true;
		if (!rec.need(be, nrec) || !len.need(be, nrec + 1) || !offs.need(be, nrec + 1) || !seeds.need(be, nrec)) { err = "bt_io: out of device memory"; return false; }
		const uint32_t none = 0xffffffffu, init[3] = { none, 0, 0 };
		be.h2d(d_misc, init, 12);
		be.each(nrec, BioRecords{ text.p, nl.p, rec.p, len.p, d_misc, d_misc + 2 });
		be.zero(len.p + nrec, 4);
		uint32_t firstBad = none;
		be.d2h(&firstBad, d_misc, 4);
		if (firstBad != none) { *irregular = 1; nrec = firstBad; if (nrec == 0) return true; be.zero(len.p + nrec, 4); }
		be.excl_scan(len.p, nrec + 1);
		uint32_t total = 0;
		be.d2h(&total, len.p + nrec, 4);
		if (!seq.need(be, (size_t)total + 1) || !qual.need(be, (size_t)total + 1)) { err = "bt_io: out of device memory"; return false; }
		be.each(nrec, BioConvert{ text.p, rec.p, len.p, seq.p, qual.p, seeds.p, gseed, d_misc });
		be.d2h(&firstBad, d_misc, 4);
		if (firstBad != none && firstBad < nrec) {                              /* an odd character: everything from that record on is the host parser's */
			*irregular = 1; nrec = firstBad;
			if (nrec == 0) return true;
			be.d2h(&total, len.p + nrec, 4);
		}
		be.each(nrec + 1, BioWiden{ len.p, offs.p });
		/* bytes consumed = up to and including the newline that ends record nrec - 1; longest read for the search's scratch */
		uint32_t lastNl = 0;
		be.d2h(&lastNl, nl.p + (4 * nrec - 1), 4);
		*consumed = (uint64_t)lastNl + 1;
		nreads = (uint32_t)nrec; nbases = total; *n_out = nreads;
		be.d2h(&maxlen, d_misc + 2, 4);                                        /* longest read of the chunk (an upper bound when the chunk was cut short) */
		return true;
	}
	/* search + f2; returns the formatted text (host memory owned by the pipe) and the four sink counters */
	bool not_covered = false;        /* set when align_format gave up on a batch the device formatter does not cover (nothing was produced) */
	bool align_format(const bt_policy_t *pol, const BioFmt &fmt_in, const char **out_text, uint64_t *out_bytes, uint64_t counters[4]) {
		not_covered = false;
		const uint32_t max_read_len = maxlen ? maxlen : 1;
		*out_text = nullptr; *out_bytes = 0;
		for (int k = 0; k < 4; k++) counters[k] = 0;
		if (nreads == 0) return true;
		if (!have_names) { err = "bt_io: reference names not set"; return false; }
		BioFmt fmt = fmt_in;
		slots = fmt.slots;
		uint32_t cap = 12;
		for (;;) {
			mm_cap = cap;
			const uint32_t rw = BIO_HIT_HDR + mm_cap;
			fmt.recWords = rw;
			if (!found.need(be, nreads) || !flags.need(be, nreads) || !hits.need(be, (size_t)nreads * slots * rw) || !olen.need(be, (size_t)nreads + 1)) { err = "bt_io: out of device memory"; return false; }
			be.zero(found.p, (size_t)nreads * 4); be.zero(flags.p, (size_t)nreads * 4);
			bt_read_batch_t in; memset(&in, 0, sizeof in);
			in.nreads = nreads; in.seq = seq.p; in.qual = qual.p; in.offs = offs.p; in.seeds = seeds.p; in.max_len = max_read_len;
			bt_hit_batch_t ho = { found.p, flags.p, hits.p, slots, mm_cap };
			if (be.align(pol, &in, &ho)) { err = bt_last_error(); return false; }
			/* a read with more mismatches than the records hold (possible only with zero-penalty qualities): once more with room for all */
			be.h2d(d_misc + 1, "\0\0\0\0", 4);
			be.each(nreads, BioAnyFlag{ flags.p, d_misc + 1 });
			uint32_t any = 0;
			be.d2h(&any, d_misc + 1, 4);
			if (any & ~(uint32_t)(BT_OVF_MM | BT_OVF_HITS)) { err = "bt_io: search scratch exhausted for a read"; return false; }
			if (!any) break;
			const uint32_t full = max_read_len < BIO_MM_MAX ? max_read_len : BIO_MM_MAX;
			if ((any & BT_OVF_HITS) || cap >= full) { not_covered = true; err = "bt_io: a read has more mismatches (or hits) than the device formatter handles: the host output path formats this batch"; return false; }
			cap = full;
		}
		be.zero(d_cnt, 4 * sizeof(unsigned long long));
		BioFormat F{ text.p, rec.p, seq.p, len.p, found.p, hits.p, BioNames{ names.p, nameOff.p }, fmt, olen.p, nullptr, nullptr, d_cnt };
		be.each(nreads, F);
		be.zero(olen.p + nreads, 4);
		be.excl_scan(olen.p, (uint64_t)nreads + 1);
		uint32_t total = 0;
		be.d2h(&total, olen.p + nreads, 4);
		unsigned long long c[4];
		be.d2h(c, d_cnt, sizeof c);
		for (int k = 0; k < 4; k++) counters[k] = c[k];
		if (total) {
			if (!out.need(be, total)) { err = "bt_io: out of device memory"; return false; }
			F.lens = nullptr; F.pos = olen.p; F.out = out.p;
			be.each(nreads, F);
			char *h = be.out_host(total, host_out);
			if (!h) { err = "bt_io: out of host memory"; return false; }
			be.d2h(h, out.p, total);
			*out_text = h;
		}
		*out_bytes = total;
		return true;
	}
};
