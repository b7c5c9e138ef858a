/*
 * bt_ctxq.cuh — lane contexts in shared memory.
 *
 * In the queue-driven kernel a read's search state (BtLane) is not tied to a thread: it lives packed in
 * shared memory (39 words, structure-of-arrays so that a warp's accesses spread over the banks) and is
 * loaded into the registers of whichever warp lane picked it from a work queue, advanced by one transition
 * (fast) or one chain of rare transitions, and stored back.  This is what lets every warp execute lanes
 * that are all in the same class of state.
 */
#pragma once
#include "bt_core.cuh"

#define BT_CTX_WORDS 39

/* word w of context c lives at ctx[w * nctx + c] */
#define CW(w) ctx[(size_t)(w) * nctx + id]

BT_FN void bt_ctx_store(const BtLane &L, uint32_t *ctx, uint32_t nctx, uint32_t id) {
	CW(0) = L.K->rid; CW(1) = L.K->seed; CW(2) = L.K->found; CW(3) = L.K->step; CW(4) = L.nit; CW(5) = L.K->rnd; CW(6) = L.K->numBts;
	CW(7) = L.K->mut0; CW(8) = L.K->mut1; CW(9) = L.K->mut2;
	CW(10) = L.top; CW(11) = L.bot; CW(12) = L.ltop; CW(13) = L.lbot; CW(14) = L.eligibleSz; CW(15) = L.eltop; CW(16) = L.elbot;
	CW(17) = L.K->bttop; CW(18) = L.K->btbot; CW(19) = L.K->rep_top; CW(20) = L.K->rep_bot; CW(21) = L.K->rep_r; CW(22) = L.K->rep_i; CW(23) = L.crow; CW(24) = L.cjumps;
	CW(25) = L.rlen | (L.qlen << 10) | ((L.flags & 63u) << 20) | ((L.K->hasN & 1u) << 26) | ((L.K->done & 1u) << 27) | ((L.ret & 1u) << 28) |
	         ((L.K->bailed & 1u) << 29) | ((L.ebwtSel & 1u) << 30) | ((L.K->fw & 1u) << 31);
	CW(26) = L.pc | (L.K->ph << 5) | (L.lfk << 9) | ((L.considerQuals & 1u) << 12) | ((L.halfAndHalf & 1u) << 13) | ((L.reportPartials & 3u) << 14) |
	         ((L.K->reportExacts & 1u) << 16) | ((L.K->nmuts & 3u) << 17) | ((L.K->disableFtab & 1u) << 19) | ((L.c & 7u) << 20) | ((L.curIsAlt & 1u) << 23) |
	         ((L.curIsElig & 1u) << 24) | ((L.curOverrides & 1u) << 25) | ((L.f_bdm & 1u) << 26) | ((L.f_must & 1u) << 27) | ((L.f_invHH & 1u) << 28) |
	         ((L.f_invExact & 1u) << 29) | ((L.elignore & 1u) << 30);
	CW(27) = L.depth5 | (L.depth3 << 10) | (L.K->unrev0 << 20) | ((L.elcint & 3u) << 30);
	CW(28) = L.K->rev1_0 | (L.rev2_0 << 10) | (L.rev3_0 << 20) | ((L.K->bt_j & 3u) << 30);
	CW(29) = L.stackDepth | (L.K->depth << 10) | (L.d << 20) | ((L.K->rep_site & 3u) << 30);
	CW(30) = L.unrevOff | (L.K->oneRevOff << 10) | (L.K->twoRevOff << 20);
	CW(31) = L.K->threeRevOff | (L.eli << 10) | (L.rowd0 << 20);
	CW(32) = L.ham;
	CW(33) = (L.altNum & 0xffffu) | (L.eligibleNum << 16);
	CW(34) = (L.K->iham & 0xffu) | ((L.elham & 0xffu) << 8) | ((L.lowAltQual & 0xffu) << 16) | ((L.q & 0xffu) << 24);
	CW(35) = L.K->bt_i | (L.K->rep_sd << 10) | ((L.K->rep_stratum & 0xfu) << 20) ;
	CW(36) = (L.K->btham & 0xffffu) | (L.K->rep_cost << 16);
	CW(37) = (L.K->npart & 0xffffu) | (L.K->pal_i << 16);
	CW(38) = L.rowbase;
}

BT_FN void bt_ctx_load(BtLane &L, const uint32_t *ctx, uint32_t nctx, uint32_t id) {
	L.K->rid = CW(0); L.K->seed = CW(1); L.K->found = CW(2); L.K->step = CW(3); L.nit = CW(4); L.K->rnd = CW(5); L.K->numBts = CW(6);
	L.K->mut0 = CW(7); L.K->mut1 = CW(8); L.K->mut2 = CW(9);
	L.top = CW(10); L.bot = CW(11); L.ltop = CW(12); L.lbot = CW(13); L.eligibleSz = CW(14); L.eltop = CW(15); L.elbot = CW(16);
	L.K->bttop = CW(17); L.K->btbot = CW(18); L.K->rep_top = CW(19); L.K->rep_bot = CW(20); L.K->rep_r = CW(21); L.K->rep_i = CW(22); L.crow = CW(23); L.cjumps = CW(24);
	uint32_t w = CW(25);
	L.rlen = w & 1023u; L.qlen = (w >> 10) & 1023u; L.flags = (w >> 20) & 63u; L.K->hasN = (w >> 26) & 1u; L.K->done = (w >> 27) & 1u; L.ret = (w >> 28) & 1u;
	L.K->bailed = (w >> 29) & 1u; L.ebwtSel = (w >> 30) & 1u; L.K->fw = w >> 31;
	w = CW(26);
	L.pc = w & 31u; L.K->ph = (w >> 5) & 15u; L.lfk = (w >> 9) & 7u; L.considerQuals = (w >> 12) & 1u; L.halfAndHalf = (w >> 13) & 1u; L.reportPartials = (w >> 14) & 3u;
	L.K->reportExacts = (w >> 16) & 1u; L.K->nmuts = (w >> 17) & 3u; L.K->disableFtab = (w >> 19) & 1u; L.c = (w >> 20) & 7u; L.curIsAlt = (w >> 23) & 1u;
	L.curIsElig = (w >> 24) & 1u; L.curOverrides = (w >> 25) & 1u; L.f_bdm = (w >> 26) & 1u; L.f_must = (w >> 27) & 1u; L.f_invHH = (w >> 28) & 1u;
	L.f_invExact = (w >> 29) & 1u; L.elignore = (w >> 30) & 1u;
	w = CW(27); L.depth5 = w & 1023u; L.depth3 = (w >> 10) & 1023u; L.K->unrev0 = (w >> 20) & 1023u; L.elcint = w >> 30;
	w = CW(28); L.K->rev1_0 = w & 1023u; L.rev2_0 = (w >> 10) & 1023u; L.rev3_0 = (w >> 20) & 1023u; L.K->bt_j = w >> 30;
	w = CW(29); L.stackDepth = w & 1023u; L.K->depth = (w >> 10) & 1023u; L.d = (w >> 20) & 1023u; L.K->rep_site = w >> 30;
	w = CW(30); L.unrevOff = w & 1023u; L.K->oneRevOff = (w >> 10) & 1023u; L.K->twoRevOff = (w >> 20) & 1023u;
	w = CW(31); L.K->threeRevOff = w & 1023u; L.eli = (w >> 10) & 1023u; L.rowd0 = (w >> 20) & 1023u;
	L.ham = CW(32);
	w = CW(33); L.altNum = w & 0xffffu; L.eligibleNum = w >> 16;
	w = CW(34); L.K->iham = w & 0xffu; L.elham = (w >> 8) & 0xffu; L.lowAltQual = (w >> 16) & 0xffu; L.q = w >> 24;
	w = CW(35); L.K->bt_i = w & 1023u; L.K->rep_sd = (w >> 10) & 1023u; L.K->rep_stratum = (w >> 20) & 0xfu;
	w = CW(36); L.K->btham = w & 0xffffu; L.K->rep_cost = w >> 16;
	w = CW(37); L.K->npart = w & 0xffffu; L.K->pal_i = w >> 16;
	L.rowbase = CW(38);
	L.viewRev = (L.ebwtSel == 0) ? !L.K->fw : L.K->fw;
	L.viewComp = !L.K->fw;
}
#undef CW

/* ---- checkpoint slots ---------------------------------------------------------------------------------------------------------
 * A slot is everything a suspended read owns: the packed lane state above (word-major in P.slot_ctx), its writable copy of the read
 * (mutated by seedlings) and a private scratch set (rows / elims / frames / seedlings) with the capacities of the later passes.
 * The main pass suspends a read by copying what is live of its per-thread scratch into a fresh slot; the tail's turns that follow work in
 * the slot's own scratch, so suspending again costs only the 39 state words. */
BT_FN void bt_slot_scratch(const BtKParams &P, uint32_t slot, BtScratch &S) {
	S.rows = P.slot_rows + (size_t)slot * P.slot_R * 2;
	S.elims = P.slot_elims + (size_t)slot * P.slot_R;
	S.frames = P.slot_frames + (size_t)slot * P.slot_FCAP;
	S.partials = P.slot_partials + (size_t)slot * P.slot_PCAP;
}

/* First suspension (main pass): per-thread scratch S -> slot.  Live rows are [0, rowbase + qlen - rowd0) of the current frame's
 * allocation (bt_blk_frame_enter reserved them), live frames [0, stackDepth] (BTLOOP writes mm_pos / mm_refc of the frame it is
 * about to push before reporting through it), live seedlings [0, npart). */
BT_FN void bt_slot_save_new(const BtLane &L, const BtKParams &P, const BtScratch &S, uint32_t slot) {
	BtScratch D; bt_slot_scratch(P, slot, D);
	uint32_t nrows = L.rowbase + (L.qlen > L.rowd0 ? L.qlen - L.rowd0 : 0u);
	if (nrows > P.R) nrows = P.R;
	if (nrows > P.slot_R) nrows = P.slot_R;
	for (uint32_t i = 0; i < 2 * nrows; i++) D.rows[i] = S.rows[i];
	for (uint32_t i = 0; i < nrows; i++) D.elims[i] = S.elims[i];
	uint32_t nfr = L.stackDepth + 1; if (nfr > P.FCAP) nfr = P.FCAP; if (nfr > P.slot_FCAP) nfr = P.slot_FCAP;
	for (uint32_t i = 0; i < nfr; i++) D.frames[i] = S.frames[i];
	uint32_t np = L.K->npart; if (np > P.PCAP) np = P.PCAP; if (np > P.slot_PCAP) np = P.slot_PCAP;
	for (uint32_t i = 0; i < np; i++) D.partials[i] = S.partials[i];
	uint8_t *st = P.slot_stage + (size_t)slot * 2 * P.slot_stage_len;
	for (uint32_t i = 0; i < L.rlen; i++) { st[i] = L.rseq[i]; st[P.slot_stage_len + i] = L.rqual[i]; }
	bt_ctx_store(L, P.slot_ctx, P.nslot, slot);
}

/* Resume: the lane continues the slot's read in the slot's scratch. */
BT_FN void bt_slot_resume(BtLane &L, const BtKParams &P, BtScratch &S, uint32_t slot) {
	bt_ctx_load(L, P.slot_ctx, P.nslot, slot);
	bt_slot_scratch(P, slot, S);
	L.rseq = P.slot_stage + (size_t)slot * 2 * P.slot_stage_len; L.rqual = L.rseq + P.slot_stage_len;
	L.qualThresh = P.pol.mode == 0 ? 0xffffffffu : P.pol.qualThresh;
	L.K->maxBts = P.pol.mode == 0 ? 0xffffffffu : P.pol.maxBts;
	L.maqPenalty = P.pol.mode == 0 ? 1u : (uint32_t)P.pol.maqRound;
}
